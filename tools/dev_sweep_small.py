"""Developer probe: the 2 x 3 sweep of tests/test_gpu_ensemble.py::test_parameter_sweep_outputs_match_oracle (argv: mode batch)."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import yaml
from helpers import load_cfg
import parameter_sweep as psw
mode, batch = sys.argv[1], int(sys.argv[2])
cfg = load_cfg("geballe_with_diamond")
for m in cfg["mats"].values():
    m["mesh"] = float(m["mesh"]) * 8.0
tmp = tempfile.mkdtemp()
path = os.path.join(tmp, "base.yaml")
yaml.safe_dump(cfg, open(path, "w"))
width = float(cfg["mats"]["p_sample"]["z"])
res, failed = psw.run_parameter_sweep(path, os.path.join(tmp, "out"), (2e-6, 5e-5), (2.0, 50.0), (width, width), (2, 3, 1),
                                      base_mesh_folder=os.path.join(tmp, "meshes"), batch=batch, mode=mode)
print("RESULT ok", len(res), "failed", [(f["run_name"], f["error"]) for f in failed])
