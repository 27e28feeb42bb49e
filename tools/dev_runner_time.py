"""Developer probe: loop time of the drop-in runners at the cfgs' own mesh sizes (XDMF + CSV outputs on)."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.chdir(ROOT)
from helpers import load_cfg
import run_with_diamond, run_no_diamond
for name, mod in (("geballe_with_diamond", run_with_diamond), ("geballe_no_diamond", run_no_diamond)):
    cfg = load_cfg(name)
    with tempfile.TemporaryDirectory() as d:
        t0 = time.time()
        mod.run_simulation(cfg, os.path.join(d, "mesh"), rebuild_mesh=True, visualize_mesh=False,
                           output_folder=os.path.join(d, "out"), watcher_points={"pside": (0.0, 0.0), "oside": (1e-6, 0.0)},
                           write_xdmf=True, suppress_print=False)
        print(f"### {name}: total {time.time() - t0:.2f} s", flush=True)
