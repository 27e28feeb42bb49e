"""Developer probe: two independent simulations on two contexts / streams from two host threads."""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
if os.environ.get("HF_DEV_LIB"):
    from heatflow_b200 import _lib as _l
    _l.LIB_PATH = os.path.abspath(os.environ["HF_DEV_LIB"])
from helpers import build_case, make_solver
from heatflow_b200 import problem
nthreads = int(sys.argv[1]); nsim = int(sys.argv[2]); mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
c = build_case("geballe_with_diamond", 1.0)
n = len(c.nodes)
from heatflow_b200.solver import HeatSolver
def mk():
    s = HeatSolver(0)
    s.set_sharing(2 if nthreads > 1 else 1)
    s.set_mesh(c.nodes, c.tris, c.cell_tag); s.set_materials(c.tags, c.kappa_t, c.rhoc_t)
    s.set_bcs(c.bc_dofs, c.bc_value, c.gauss_slot, c.gauss_r); s.build_operator(c.dt, True)
    s.set_solver(rtol=1e-14, warm=1.0, mode=mode); s.set_recycle(128)
    return s
solvers = [mk() for _ in range(nthreads)]
fw = np.logspace(-6, -4, 64)
def work(t):
    s = solvers[t]
    for j in range(nsim):
        s.set_state(np.full(n, c.ic))
        s.run(c.amps, c.ic, problem.gaussian_coeff(fw[(7 * t + 3 * j) % 64]), [0, n // 2])
for rep in range(2):
    t0 = time.time()
    th = [threading.Thread(target=work, args=(t,)) for t in range(nthreads)]
    [x.start() for x in th]; [x.join() for x in th]
    dt = time.time() - t0
    print(f"mode={mode} path={solvers[0].solver_path()} threads={nthreads} sims={nthreads*nsim} wall={dt:.3f}s -> {nthreads*nsim/dt:.1f} sims/s", flush=True)
