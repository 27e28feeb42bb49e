"""Run time and PCG iterations of a whole simulation against the recycle-basis size: python tools/dev_caps.py [cfg] [scale]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import build_case
from bench import configured_solver
cfg = sys.argv[1] if len(sys.argv) > 1 else "geballe_with_diamond"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
c = build_case(cfg, scale)
for cap in (0, 8, 16, 24, 32, 48, 64, 96, 128):
    s = configured_solver(c, 0, 1e-14, warm=1.0, recycle=cap)
    n, _ = s.sizes()
    for rep in range(2):
        s.set_state(np.full(n, c.ic))
        s.set_profile(rep == 1)
        _, iters, _ = s.run(c.amps, c.ic, c.coeff, [0])
    ms, _ = s.solve_profile()
    run = s.stats()["run_ms"]
    print(f"cap {cap:3d}: run {run:7.2f} ms, solves {ms:7.2f} ms, other {run - ms:6.2f} ms, iterations {int(iters.sum()):6d}, "
          f"last 10 steps {iters[-10:].tolist()}", flush=True)
    s.close()
