"""Small ncu target for the ensemble kernels (argv: cfg scale steps B)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import build_case, make_solver
from heatflow_b200 import problem
name, scale, steps, B = sys.argv[1], float(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
c = build_case(name, scale)
s = make_solver(c, ordering="hilbert")
tag = int(c.tags[[m.name for m in c.mats].index("p_sample")])
s.ens_create(np.logspace(0, 2, 64)[20:20 + B], [problem.gaussian_coeff(f) for f in np.logspace(-6, -4, 64)[10:10 + B]], tag)
s.ens_run(c.amps[20:20 + steps], c.ic, [0])
print("done", s.sizes())
