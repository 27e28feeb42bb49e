"""Developer probe: PCG iterations per time step with the recycled initial guess."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import build_case, make_solver
name, scale, cap = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
c = build_case(name, scale)
s = make_solver(c, warm=float(sys.argv[4]) if len(sys.argv) > 4 else 1.0, recycle=cap)
hist, iters, _ = s.run(c.amps, c.ic, c.coeff, [0])
print(name, scale, "N", len(c.nodes), "cap", cap, "total", int(iters.sum()))
print(" ".join(str(int(i)) for i in iters))
print("amps", " ".join(f"{a:.0f}" for a in c.amps[:40]))
