"""Per-iteration time of the on-chip PCG kernels (mode 3 = pipelined where available, mode 4 = classic):
python tools/dev_patch.py [cfg] [scale] [steps] [recycle]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
if os.environ.get("HF_DEV_LIB"):
    from heatflow_b200 import _lib as _l
    _l.LIB_PATH = os.path.abspath(os.environ["HF_DEV_LIB"])
from helpers import build_case
from bench import configured_solver
cfg = sys.argv[1] if len(sys.argv) > 1 else "geballe_with_diamond"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 25
rec = int(sys.argv[4]) if len(sys.argv) > 4 else 128
c = build_case(cfg, scale)
out = {}
for mode in ((4, 3) if not os.environ.get("HF_ONLY_PIPE") else (3,)):
    s = configured_solver(c, 0, 1e-14, warm=1.0, mode=mode, recycle=rec)
    n, nnz = s.sizes()
    s.set_state(np.full(n, c.ic)); s.run(c.amps[:steps], c.ic, c.coeff, [0])
    s.set_profile(True)
    s.set_state(np.full(n, c.ic))
    _, iters, _ = s.run(c.amps[:steps], c.ic, c.coeff, [0])
    ms, _ = s.solve_profile()
    st = s.stats()
    out[mode] = s.get_state()
    print(f"mode {mode} path {s.solver_path()}: N={n} iterations={int(iters.sum())} solve {ms:.3f} ms = {ms*1e3/max(1,iters.sum()):.3f} us/iteration; "
          f"run {st['run_ms']:.3f} ms = {n*steps/st['run_ms']/1e3:.1f} M DOF-steps/s; retries {st['retries']}", flush=True)
    s.close()
if 4 in out:
    print("max rel diff between the two kernels:", np.abs(out[3] / out[4] - 1).max())
