"""torchrun worker of test_sweep_host.py: run_parameter_sweep's multi-rank flow over gloo, world_size 2 (CPU).

The device part (``_group_on_device``) is replaced by a stand-in that produces a known watcher history per variant
and feeds it to the real output writer; everything around it is the product code: mesh built by rank 0, tiles dealt
to the ranks, every rank writing the run folders of its own variants, ONE final gather, summary files on rank 0.
Second sweep: the stand-in raises on rank 1 - that rank's variants must come back as failed and rank 0 must not hang.
"""
import os
import sys

import numpy as np
import pandas as pd
import torch.distributed as dist
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from heatflow_b200 import parameter_sweep as psw, sweep  # noqa: E402
from helpers import load_cfg  # noqa: E402

STATE = {"fail_rank": None}


def fake_hist(combo, S):
    return np.column_stack((np.full(S, combo["k"]) + np.arange(S), np.full(S, combo["fwhm"] * 1e6) + np.arange(S)))


def fake_group(base_config, combinations, mesh_folder, batch, device, tiles, suppress_print, engine="auto", output_dir=None,
               names=None, share=None, on_ready=None):
    rank = dist.get_rank()
    assert os.path.isfile(os.path.join(mesh_folder, "mesh.msh"))         # rank 0 built it and marked it ready
    if STATE["fail_rank"] == rank:
        raise RuntimeError(f"device lost on rank {rank}")
    S = int(base_config["timing"]["num_steps"])
    step_t = (np.arange(S) + 1) * float(base_config["timing"]["t_final"]) / S
    writer = psw._OutputWriter(output_dir, base_config, combinations, step_t, names)
    idx = np.concatenate([np.asarray(t, dtype=np.int64) for t in tiles]) if len(tiles) else np.zeros(0, np.int64)
    for i in idx:
        writer.submit(int(i), fake_hist(combinations[int(i)], S), 7, 0.01, None)
    errors = writer.close()
    return idx, None, np.full(len(idx), 7, dtype=np.int64), np.full(len(idx), 0.01), errors, step_t


def main():
    tmp = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world, _ = sweep.dist_info()
    assert world == 2
    psw._group_on_device = fake_group
    psw._visible_gpus = lambda: 1
    cfg = load_cfg("geballe_no_diamond")
    for m in cfg["mats"].values():
        m["mesh"] = float(m["mesh"]) * 16.0
    cfg_path = os.path.join(tmp, "base.yaml")
    if rank == 0:
        with open(cfg_path, "w") as f:
            yaml.safe_dump(cfg, f)
    dist.barrier()
    width = float(cfg["mats"]["p_sample"]["z"])
    S = int(cfg["timing"]["num_steps"])
    for name, fail_rank in (("a", None), ("b", 1)):
        STATE["fail_rank"] = fail_rank
        out = os.path.join(tmp, "out_" + name)
        results, failed = psw.run_parameter_sweep(cfg_path, out, (1e-6, 1e-4), (1.0, 100.0), (width, width), (3, 4, 1),
                                                  base_mesh_folder=os.path.join(tmp, "meshes"), batch=2)
        dist.barrier()
        if rank != 0:
            assert results == [] and failed == []
            continue
        combos, _, _, _ = psw.create_parameter_grid((1e-6, 1e-4), (1.0, 100.0), (width, width), (3, 4, 1))
        tiles = sweep.plan_tiles([c["k"] for c in combos], 2, 2)
        owner = {int(i): r for r in range(2) for t in tiles[r] for i in t}
        assert len(results) + len(failed) == 12
        if fail_rank is None:
            assert failed == [] and [r["run_id"] for r in results] == list(range(1, 13))
        else:
            assert sorted(r["run_id"] - 1 for r in failed) == sorted(i for i, r in owner.items() if r == 1)
            assert all("device lost on rank 1" in r["error"] for r in failed)
            assert os.path.isfile(os.path.join(out, "failed_runs.csv"))
        assert len(pd.read_csv(os.path.join(out, "successful_runs.csv"))) == len(results)
        for r in results:                                     # folders written by BOTH ranks, contents per variant
            df = pd.read_csv(os.path.join(r["output_dir"], "watcher_points.csv"))
            combo = combos[r["run_id"] - 1]
            assert list(df.columns) == ["time", "pside", "oside"] and len(df) == S
            assert np.allclose(df[["pside", "oside"]].to_numpy(), fake_hist(combo, S))
            used = yaml.safe_load(open(os.path.join(r["output_dir"], "used_config.yaml")))
            assert used == psw.modify_config_for_parameters(cfg, combo["fwhm"], combo["k"], combo["width"])
        assert {owner[r["run_id"] - 1] for r in results} == ({0, 1} if fail_rank is None else {0})
    if rank == 0:
        with open(os.path.join(tmp, "ok.txt"), "w") as f:
            f.write("ok")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
