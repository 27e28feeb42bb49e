import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import build_case, make_solver
c = build_case("geballe_no_diamond", 8.0)
for mode in (1, 2):
    s = make_solver(c, mode=mode)
    s.run(c.amps[:25], c.ic, c.coeff, [])
    for rep in range(3):
        g = s.project_gradient()
        print("mode", mode, "rep", rep, "iters", s.last_projection_iters, "nan z/r", np.isnan(g[:,0]).sum(), np.isnan(g[:,1]).sum(), "max", np.nanmax(np.abs(g), axis=0))
    s.close()
