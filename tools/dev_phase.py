"""Phase clocks of the pipelined on-chip kernel (library built with -DHF_PHASE_TIMING): python tools/dev_phase.py"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import build_case
from bench import configured_solver
from heatflow_b200 import _lib
if os.environ.get("HF_DEV_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["HF_DEV_LIB"])
c = build_case("geballe_with_diamond", 1.0)
s = configured_solver(c, 0, 1e-14, warm=1.0, mode=3, recycle=0)
n, _ = s.sizes()
s.set_state(np.full(n, c.ic)); s.run(c.amps[:8], c.ic, c.coeff, [0])
G = 160
buf = np.zeros(G * 2 * 8, np.int64)
_lib.check(s._L.hf_debug_phase_times(s._h, None, buf.size))
its, _ = s.step(c.amps[8], c.ic, c.coeff)
_lib.check(s._L.hf_debug_phase_times(s._h, _lib.ptr(buf), buf.size))
a = buf.reshape(G, 2, 8).astype(float) / max(1, its)
live = a[:, 0, :].sum(axis=1) > 0
names = ["dots+arrive", "spmv", "publish+fetch", "wait", "update", "sync", "-", "loop head"]
print(f"{its} iterations, {live.sum()} CTAs; cycles per iteration (mean / max over CTAs)")
for w in (0, 1):
    print(f" warp {w}: " + "  ".join(f"{nm} {a[live, w, i].mean():.0f}/{a[live, w, i].max():.0f}" for i, nm in enumerate(names) if nm != "-"),
          f" total {a[live, w, :].sum(axis=1).mean():.0f}")
s.close()
