"""Developer probe: where the device set-up of a sweep group goes (cProfile of Simulation2D construction)."""
import cProfile, pstats, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import load_cfg
from heatflow_b200 import problem
from heatflow_b200.runners import Simulation2D, prepare_mesh
cfg = load_cfg("geballe_with_diamond")
tmp = tempfile.mkdtemp()
prepare_mesh(cfg, problem.stack_with_diamond, tmp, rebuild_mesh=True)
Simulation2D(cfg, problem.stack_with_diamond, tmp, rebuild_mesh=False, device=0).close()      # CUDA context, module load
def go():
    t0 = time.time()
    a = Simulation2D(cfg, problem.stack_with_diamond, tmp, rebuild_mesh=False, device=0, sharing=2)
    t1 = time.time()
    b = Simulation2D(cfg, problem.stack_with_diamond, tmp, rebuild_mesh=False, device=0, sharing=2)
    t2 = time.time()
    print(f"first {t1 - t0:.2f}s second {t2 - t1:.2f}s")
    a.close(); b.close()
pr = cProfile.Profile(); pr.enable(); go(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
