"""Small target for ncu: a few time steps of a config (argv: cfg scale steps [mode [mesher]])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import build_case, make_solver
name, scale, steps = sys.argv[1], float(sys.argv[2]), int(sys.argv[3])
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
method = sys.argv[5] if len(sys.argv) > 5 else None
c = build_case(name, scale, method=method)
s = make_solver(c, mode=mode)
k0 = 20
_, iters, _ = s.run(c.amps[k0:k0 + steps], c.ic, c.coeff, [0])
print("done", s.sizes(), "path", s.solver_path(), "PCG iterations per step", iters.tolist())
