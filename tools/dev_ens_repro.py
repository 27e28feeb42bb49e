"""Developer probe: a B = 8 tile with warm start and recycled bases on a small mesh (argv: scale recycle)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from heatflow_b200 import _lib
if os.environ.get("HF_DEV_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["HF_DEV_LIB"])
from helpers import build_case, make_solver
from heatflow_b200 import problem
scale = float(sys.argv[1]); rec = int(sys.argv[2]); nv = int(sys.argv[3]) if len(sys.argv) > 3 else 6
c = build_case("geballe_with_diamond", scale)
tag = int(c.tags[[m.name for m in c.mats].index("p_sample")])
s = make_solver(c, warm=1.0, ordering="auto", recycle=rec)
ks = [2.0, 2.0, 10.0, 10.0, 50.0, 50.0, 70.0, 90.0][:nv]
fw = [2e-6, 5e-5] * 4
try:
    s.ens_create(ks, [problem.gaussian_coeff(f) for f in fw[:nv]], tag)
    print("path", s.ens_path())
    hist, iters = s.ens_run(c.amps, c.ic, [0, len(c.nodes) // 2])
    print("ok iters", iters.tolist()[-8:], "finite", np.isfinite(hist).all(), "retries", s.stats()["retries"])
except Exception as e:
    print("FAILED:", e)
