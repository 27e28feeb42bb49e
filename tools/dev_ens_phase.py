"""Phase clocks of the batched on-chip ensemble kernel (library built with -DHF_PHASE_TIMING, HF_DEV_LIB=path):
python tools/dev_ens_phase.py [scale]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from heatflow_b200 import _lib
if os.environ.get("HF_DEV_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["HF_DEV_LIB"])
from helpers import build_case, make_solver
from heatflow_b200 import problem
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
c = build_case("geballe_with_diamond", scale)
tag = int(c.tags[[m.name for m in c.mats].index("p_sample")])
n = len(c.nodes)
fws = np.logspace(-6, -4, 64)
cf = [problem.gaussian_coeff(f) for f in fws[10:14]]
names = ["spmv+dots", "arrive", "halo fetch", "wait+ctl", "update(+check)", "sync", "-", "-"]
for nh in ("1", "2"):
    os.environ["HF_ENS_NH"] = nh
    s = make_solver(c, warm=0.0, ordering="hilbert", recycle=0)
    s.ens_create([10.0] * 4, cf, tag)
    s.ens_run(c.amps[:8], c.ic, [0])
    G = 160
    buf = np.zeros(G * 2 * 8, np.int64)
    _lib.check(s._L.hf_debug_phase_times(s._h, None, buf.size))
    s.set_profile(True)
    _, iters = s.ens_run(c.amps[8:9], c.ic, [0])
    ms, _ = s.solve_profile()
    _lib.check(s._L.hf_debug_phase_times(s._h, _lib.ptr(buf), buf.size))
    its = int(iters[0])
    a = buf.reshape(G, 2, 8).astype(float) / max(1, its)
    live = a[:, 0, :].sum(axis=1) > 0
    print(f"NH={nh}: path {s.ens_path()}, {its} iterations in {ms*1e3:.0f} us = {ms*1e3/its:.2f} us/iteration, {live.sum()} CTAs; cycles per iteration (mean / max over CTAs)")
    for w in (0, 1):
        print(f" warp {w}: " + "  ".join(f"{nm} {a[live, w, i].mean():.0f}/{a[live, w, i].max():.0f}" for i, nm in enumerate(names) if nm != "-"),
              f" total {a[live, w, :].sum(axis=1).mean():.0f}")
    _lib.check(s._L.hf_debug_phase_times(s._h, None, 0))
    s.ens_destroy(); s.close()
