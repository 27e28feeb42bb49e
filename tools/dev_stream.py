"""Per-iteration time of the streaming PCG kernels (modes 1 and 2) inside real solves: python tools/dev_stream.py [scale] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from helpers import build_case
from bench import configured_solver, iter_bytes
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.35
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = build_case("konopkova", scale)
res = {}
for mode in (1, 2):
    s = configured_solver(c, 0, 1e-14, mode=mode)
    n, nnz = s.sizes()
    s.set_state(np.full(n, c.ic))
    s.run(c.amps[10:11], c.ic, c.coeff, [0])
    s.set_profile(True)
    _, iters, _ = s.run(c.amps[11:11 + steps], c.ic, c.coeff, [0])
    ms, launches = s.solve_profile()
    us = ms * 1e3 / iters.sum()
    res[mode] = s.get_state()
    print(f"mode {mode}: N={n} iterations={iters.tolist()} solve {ms:.2f} ms, {us:.2f} us/iteration, "
          f"{iter_bytes(n, nnz) / us / 1e3:.0f} GB/s algorithmic, launches {launches}, path {s.solver_path()}", flush=True)
    s.close()
print("modes 1 and 2 bit-identical:", np.array_equal(res[1], res[2]), np.abs(res[1] / res[2] - 1).max())
