"""Developer timing of the ensemble kernels: python tools/dev_ens_timing.py [cfg] [size_scale] [steps] [B ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
if os.environ.get("HF_DEV_LIB"):
    from heatflow_b200 import _lib as _l
    _l.LIB_PATH = os.path.abspath(os.environ["HF_DEV_LIB"])
from helpers import build_case, make_solver  # noqa: E402
from heatflow_b200 import problem  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "geballe_with_diamond"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 100
Bs = [int(a) for a in sys.argv[4:]] or [4, 8, 16, 32]
c = build_case(name, scale)
n = len(c.nodes)
tag = int(c.tags[[m.name for m in c.mats].index("p_sample")])
s = make_solver(c, warm=float(os.environ.get("HF_WARM", "1")), ordering=os.environ.get("HF_ORD", "hilbert"),
                recycle=int(os.environ.get("HF_RECYCLE", "0")))
for B in Bs:
    ks = np.logspace(0, 2, 64)[20:20 + B]
    fw = np.logspace(-6, -4, 64)[10:10 + B]
    s.set_state(np.full(n, c.ic))
    t0 = time.time()
    s.ens_create(ks, [problem.gaussian_coeff(f) for f in fw], tag)
    t_create = time.time() - t0
    s.ens_run(c.amps[20:23], c.ic, [0])                     # warm-up (graph capture)
    s.ens_destroy()
    s.set_state(np.full(n, c.ic))
    s.ens_create(ks, [problem.gaussian_coeff(f) for f in fw], tag)
    t0 = time.time()
    hist, iters = s.ens_run(c.amps[:steps], c.ic, [0, 1])
    wall = time.time() - t0
    ms = s.stats()["run_ms"]
    tot = int(iters.sum())
    print(f"{name} scale={scale} N={n} B={B}: create={t_create:.3f}s steps={steps} iters={tot} max={iters.max()} "
          f"dev={ms / 1e3:.3f}s wall={wall:.3f}s -> {B / (ms / 1e3):.2f} sims/s, {ms * 1e3 / tot:.2f} us/iter, "
          f"{ms * 1e3 / tot / B:.2f} us/iter/sim, {B * n * steps / (ms / 1e3) / 1e6:.1f} MDOF-steps/s, "
          f"eff. {(96 + 145 / B) * n * B / (ms * 1e-3 / tot) / 1e9:.0f} GB/s", flush=True)
    s.ens_destroy()
s.close()
