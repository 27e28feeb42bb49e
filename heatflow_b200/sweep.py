"""Execution engine of ``parameter_sweep``: tiles of B variants per GPU, variants sharded across ranks.

The reference runs one OS process per parameter set, each re-reading the mesh, re-assembling and
re-factorising (parameter_sweep.py:123-192, :436-438).  Here the variants of one width group
(same mesh) differ only in the sample conductivity and the Gaussian heating width, so a rank

  * sets the mesh / pattern / base operator up once (``Simulation2D``),
  * advances its variants tile by tile - engine ``'ensemble'``: ``batch`` simulations at once through
    the batched multi-RHS kernels (``hf_ens_create`` / ``hf_ens_run``), which read the operator once
    for the whole tile (meshes that stream from HBM); engine ``'serial'``: one simulation after the
    other through the single-simulation path, re-assembling only when the conductivity changes
    (meshes that fit on chip: the persistent kernel is latency-bound, so batching buys nothing and
    every simulation keeps the recycled-initial-guess speed-up); ``'auto'`` picks by
    ``HeatSolver.on_chip()`` - and
  * hands its ``[P_local, S, n_watch]`` watcher histories to rank 0 in ONE final gather - the only
    collective of the sweep.

Sharding: variants are ordered by conductivity (similar PCG iteration counts inside a tile), cut
into tiles of ``batch`` and dealt round-robin to the ranks, so every rank sees the same mix of
cheap and expensive tiles.
"""
from __future__ import annotations

import os
import time

import numpy as np

from . import problem


def plan_tiles(k_values, batch, world_size):
    """Tiles of variant indices per rank.

    Returns ``tiles[rank] = [index arrays]``.  Every variant appears exactly once; tiles hold at
    most ``batch`` variants, are contiguous in ascending-k order (ties keep input order) and are
    dealt round-robin over the ranks.
    """
    if batch < 1:
        raise ValueError("batch must be >= 1")
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    order = np.argsort(np.asarray(k_values, dtype=np.float64), kind="stable")
    cut = [order[i:i + batch] for i in range(0, len(order), batch)]
    return [cut[r::world_size] for r in range(world_size)]


def run_tiles_serial(sims, fwhm, k, tiles, watch_nodes, sample_name="p_sample"):
    """``run_tiles`` through the single-simulation path: same arguments, same return value
    (iters = PCG iterations of the variant itself, seconds = wall time of the variant).

    ``sims``: one ``Simulation2D`` or a list of them on the same device.  With two (each planned with
    ``sharing=2``) the variants are pulled from a common queue by two host threads, one context and stream
    each: the on-chip kernels of the two simulations are co-resident and hide each other's reduction
    latency (the C calls release the GIL)."""
    import threading
    sims = list(sims) if isinstance(sims, (list, tuple)) else [sims]
    S, W = sims[0].num_steps, len(watch_nodes)
    todo = [int(i) for tile in tiles for i in np.asarray(tile, dtype=np.int64)]
    out, errors, lock, cursor = {}, {}, threading.Lock(), [0]
    k_arr, f_arr = np.asarray(k, dtype=np.float64), np.asarray(fwhm, dtype=np.float64)

    def worker(sim):
        s = sim.solver
        k_now = None
        u0 = np.full(sim.n_dofs, sim.ic_temp)
        while True:
            with lock:
                if cursor[0] >= len(todo):
                    return
                i = todo[cursor[0]]
                cursor[0] += 1
            t0 = time.time()
            try:
                ki = float(k_arr[i])
                if k_now != ki:
                    k_now = None                      # a failed re-assembly must not be mistaken for this k
                    sim.set_conductivity(sample_name, ki)
                    k_now = ki
                s.set_state(u0)
                hist, iters, _ = s.run(sim.amps, sim.ic_temp, problem.gaussian_coeff(float(f_arr[i])), watch_nodes)
                its = int(iters.sum())
            except Exception as exc:                   # recorded per run, as the reference does
                with lock:
                    errors[i] = str(exc)
                hist = np.full((S, W), np.nan)
                its = -1
            with lock:
                out[i] = (hist, its, time.time() - t0)

    if len(sims) == 1:
        worker(sims[0])
    else:
        threads = [threading.Thread(target=worker, args=(sim,)) for sim in sims]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if not todo:
        return (np.zeros(0, np.int64), np.zeros((0, S, W)), np.zeros(0, np.int64), np.zeros(0), errors)
    return (np.asarray(todo, dtype=np.int64), np.stack([out[i][0] for i in todo]).reshape(len(todo), S, W),
            np.asarray([out[i][1] for i in todo], dtype=np.int64), np.asarray([out[i][2] for i in todo]), errors)


def run_tiles(sim, fwhm, k, tiles, watch_nodes, sample_name="p_sample", engine="ensemble", extra_sims=()):
    """Advance every tile on ``sim.solver``'s device (``extra_sims``: further simulations on the same
    device for the concurrent serial engine).

    ``fwhm`` / ``k``: arrays over ALL variants; ``tiles``: index arrays owned by this rank.
    Returns (indices [P_local], hist [P_local, S, n_watch], iters [P_local] (PCG iterations of the
    variant's tile, summed over steps), seconds [P_local] (tile wall time / tile size), errors
    {index: message}).
    """
    if engine not in ("auto", "ensemble", "serial"):
        raise ValueError("engine must be 'auto', 'ensemble' or 'serial'")
    s = sim.solver
    if engine == "serial" or (engine == "auto" and s.on_chip()):
        return run_tiles_serial([sim, *extra_sims], fwhm, k, tiles, watch_nodes, sample_name)
    S, W = sim.num_steps, len(watch_nodes)
    idx_all, hist_all, it_all, sec_all, errors = [], [], [], [], {}
    for tile in tiles:
        tile = np.asarray(tile, dtype=np.int64)
        t0 = time.time()
        try:
            s.set_state(np.full(sim.n_dofs, sim.ic_temp))
            s.ens_create(np.asarray(k)[tile], [problem.gaussian_coeff(f) for f in np.asarray(fwhm)[tile]],
                         sim.sample_tag(sample_name))
            hist, iters = s.ens_run(sim.amps, sim.ic_temp, watch_nodes)
            its = int(iters.sum())
        except Exception as exc:                       # recorded per run, as the reference does
            for i in tile:
                errors[int(i)] = str(exc)
            hist = np.full((len(tile), S, W), np.nan)
            its = -1
        finally:
            try:
                s.ens_destroy()
            except Exception:
                pass
        dt_tile = (time.time() - t0) / max(1, len(tile))
        idx_all.append(tile)
        hist_all.append(hist)
        it_all.append(np.full(len(tile), its, dtype=np.int64))
        sec_all.append(np.full(len(tile), dt_tile))
    if not idx_all:
        return (np.zeros(0, np.int64), np.zeros((0, S, W)), np.zeros(0, np.int64), np.zeros(0), errors)
    return (np.concatenate(idx_all), np.concatenate(hist_all), np.concatenate(it_all), np.concatenate(sec_all), errors)


# ----------------------------------------------------------------------------------------
# the single collective of the sweep
# ----------------------------------------------------------------------------------------
def dist_info():
    """(rank, world_size, local_rank) from torch.distributed if initialised, else the torchrun env."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size(), int(os.environ.get("LOCAL_RANK", dist.get_rank()))
    except ImportError:
        pass
    return 0, 1, 0


def gather_results(n_total, S, W, idx, hist, iters, secs, errors):
    """Final gather to rank 0.  Returns (hist [P,S,W], iters [P], secs [P], errors) on rank 0 and
    ``None`` elsewhere.  Works on NCCL (device tensors) and gloo (CPU tensors)."""
    rank, world, _ = dist_info()
    if world == 1:
        out = np.full((n_total, S, W), np.nan)
        it = np.full(n_total, -1, dtype=np.int64)
        sc = np.zeros(n_total)
        out[idx], it[idx], sc[idx] = hist, iters, secs
        return out, it, sc, dict(errors)
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    # fixed-size payload per rank: [cap, 3 + S*W] (index, iterations, seconds, history); unused rows index -1
    cap_t = torch.tensor([len(idx)], dtype=torch.int64, device=dev)
    dist.all_reduce(cap_t, op=dist.ReduceOp.MAX)
    cap = int(cap_t.item())
    pay = torch.full((cap, 3 + S * W), -1.0, dtype=torch.float64)
    if len(idx):
        pay[:len(idx), 0] = torch.from_numpy(idx.astype(np.float64))
        pay[:len(idx), 1] = torch.from_numpy(iters.astype(np.float64))
        pay[:len(idx), 2] = torch.from_numpy(np.asarray(secs, dtype=np.float64))
        pay[:len(idx), 3:] = torch.from_numpy(np.ascontiguousarray(hist).reshape(len(idx), S * W))
    pay = pay.to(dev)
    bucket = [torch.empty_like(pay) for _ in range(world)] if rank == 0 else None
    dist.gather(pay, bucket, dst=0)
    err_list = [None] * world if rank == 0 else None
    dist.gather_object(dict(errors), err_list, dst=0)
    if rank != 0:
        return None
    out = np.full((n_total, S, W), np.nan)
    it = np.full(n_total, -1, dtype=np.int64)
    sc = np.zeros(n_total)
    for t in bucket:
        a = t.cpu().numpy()
        a = a[a[:, 0] >= 0]
        ii = a[:, 0].astype(np.int64)
        out[ii] = a[:, 3:].reshape(len(ii), S, W)
        it[ii] = a[:, 1].astype(np.int64)
        sc[ii] = a[:, 2]
    merged = {}
    for e in err_list:
        merged.update(e)
    return out, it, sc, merged
