"""Developer probe: cost of the per-step pieces of run_no_diamond's loop."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import build_case, make_solver
c = build_case("geballe_no_diamond", 1.0)
s = make_solver(c, warm=1.0, recycle=64)
n = len(c.nodes)
s.run(c.amps[:12], c.ic, c.coeff, [0, 5])
for k in range(3):
    t0 = time.time(); h, it, f = s.run(c.amps[12 + k:13 + k], c.ic, c.coeff, [0, 5], keep_fields=True); t1 = time.time()
    g = s.project_gradient(); t2 = time.time()
    u = s.get_state(); t3 = time.time()
    print(f"run(1 step, fields) {1e3*(t1-t0):.2f} ms [{int(it.sum())} its], project_gradient {1e3*(t2-t1):.2f} ms, get_state {1e3*(t3-t2):.2f} ms")
import ctypes as C
it = C.c_int32()
