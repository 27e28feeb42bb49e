"""Developer probe: wall clock of run_parameter_sweep on the cfg's own mesh per engine (argv: n_fwhm n_k [mode:batch ...])."""
import os, sys, time, tempfile, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import yaml
from helpers import load_cfg
import parameter_sweep as psw
n_f, n_k = int(sys.argv[1]), int(sys.argv[2])
modes = sys.argv[3:] or ["auto:16", "ensemble:4"]
cfg = load_cfg("geballe_with_diamond")
tmp = tempfile.mkdtemp(prefix="hf_sweep_")
cfg_path = os.path.join(tmp, "base.yaml")
with open(cfg_path, "w") as f:
    yaml.safe_dump(cfg, f)
width = float(cfg["mats"]["p_sample"]["z"])
meshes = os.path.join(tmp, "meshes")
psw.run_parameter_sweep(cfg_path, os.path.join(tmp, "warm"), (1e-6, 1e-4), (1.0, 100.0), (width, width), (1, 2, 1), base_mesh_folder=meshes)
for spec in modes:
    mode, batch = spec.split(":")
    out = os.path.join(tmp, "out_" + mode + batch)
    t0 = time.perf_counter()
    res, failed = psw.run_parameter_sweep(cfg_path, out, (1e-6, 1e-4), (1.0, 100.0), (width, width), (n_f, n_k, 1),
                                          base_mesh_folder=meshes, mode=mode, batch=int(batch))
    dt = time.perf_counter() - t0
    print(f"SWEEP mode={mode} batch={batch}: {n_f * n_k} variants in {dt:.2f} s = {n_f * n_k / dt:.1f} sims/s, ok {len(res)} failed {len(failed)}", file=sys.stderr, flush=True)
shutil.rmtree(tmp, ignore_errors=True)
