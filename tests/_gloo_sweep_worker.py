"""torchrun worker of test_sweep_host.py: the sweep's final gather over gloo, world_size 2 (CPU)."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from heatflow_b200 import sweep  # noqa: E402


def fake_hist(i, S, W):
    return (1000.0 * i + np.arange(S)[:, None] + 0.25 * np.arange(W)[None, :]).astype(np.float64)


def main():
    out_file = sys.argv[1]
    dist.init_process_group("gloo")
    rank, world, _ = sweep.dist_info()
    assert world == 2 and rank == dist.get_rank()
    P, S, W, B = 37, 5, 2, 4
    k = np.linspace(100.0, 1.0, P)                       # descending: the plan must sort
    tiles = sweep.plan_tiles(k, B, world)
    mine = tiles[rank]
    idx = np.concatenate(mine)
    hist = np.stack([fake_hist(i, S, W) for i in idx])
    iters = 10 * idx + rank
    secs = 0.5 + idx.astype(np.float64)
    errors = {int(idx[0]): f"boom on rank {rank}"}
    hist[0] = np.nan
    got = sweep.gather_results(P, S, W, idx, hist, iters, secs, errors)
    if rank != 0:
        assert got is None
    else:
        H, it, sc, err = got
        owner = np.empty(P, dtype=int)
        for r in range(world):
            owner[np.concatenate(tiles[r])] = r
        first = {int(np.concatenate(tiles[r])[0]) for r in range(world)}
        assert set(err) == first and all(err[i] == f"boom on rank {owner[i]}" for i in first)
        for i in range(P):
            if i in first:
                assert np.all(np.isnan(H[i]))
            else:
                assert np.array_equal(H[i], fake_hist(i, S, W))
            assert it[i] == 10 * i + owner[i] and sc[i] == 0.5 + i
        with open(out_file, "w") as f:
            f.write("ok")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
