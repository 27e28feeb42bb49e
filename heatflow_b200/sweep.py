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
  * writes the run folders of its own variants while the GPU works on the next ones (``on_done``), and
  * hands its per-variant summary (and, on request, the ``[P_local, S, n_watch]`` watcher histories) to rank 0
    in ONE final gather - the only collective of the sweep.

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


def run_tiles_serial(sims, fwhm, k, tiles, watch_nodes, sample_name="p_sample", on_done=None):
    """``run_tiles`` through the single-simulation path: same arguments, same return value
    (iters = PCG iterations of the variant itself, seconds = wall time of the variant).

    ``sims``: one ``Simulation2D`` or a list of them on the same device.  With two (each planned with
    ``sharing=2``) the variants are pulled from a common queue by two host threads, one context and stream
    each: the on-chip kernels of the two simulations are co-resident and hide each other's reduction
    latency (the C calls release the GIL).  ``on_done(i, hist_i, iters_i, seconds_i, error_or_None)`` is called
    from the worker thread as soon as variant ``i`` has finished."""
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
            if on_done is not None:
                on_done(i, hist, its, out[i][2], errors.get(i))

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


def run_tiles(sim, fwhm, k, tiles, watch_nodes, sample_name="p_sample", engine="ensemble", extra_sims=(), on_done=None):
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
        return run_tiles_serial([sim, *extra_sims], fwhm, k, tiles, watch_nodes, sample_name, on_done)
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
        if on_done is not None:
            for pos, i in enumerate(tile):
                on_done(int(i), hist[pos], its, dt_tile, errors.get(int(i)))
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
    """Final gather to rank 0 - ONE collective (``gather_object`` of one small dict per rank; pickled numpy arrays,
    so it works on NCCL and gloo alike).  ``hist`` may be ``None`` when every rank has already written its own run
    folders and rank 0 only needs the summary.  Returns (hist [P,S,W] | None, iters [P], secs [P], errors) on rank 0
    and ``None`` elsewhere."""
    rank, world, _ = dist_info()
    mine = {"idx": np.asarray(idx, dtype=np.int64), "iters": np.asarray(iters, dtype=np.int64),
            "secs": np.asarray(secs, dtype=np.float64), "errors": dict(errors),
            "hist": None if hist is None else np.ascontiguousarray(hist, dtype=np.float64).reshape(len(idx), S, W)}
    if world == 1:
        parts = [mine]
    else:
        import torch.distributed as dist
        parts = [None] * world if rank == 0 else None
        dist.gather_object(mine, parts, dst=0)
        if rank != 0:
            return None
    out = None if hist is None else np.full((n_total, S, W), np.nan)
    it = np.full(n_total, -1, dtype=np.int64)
    sc = np.zeros(n_total)
    merged = {}
    for part in parts:
        ii = part["idx"]
        if out is not None and part["hist"] is not None:
            out[ii] = part["hist"]
        it[ii], sc[ii] = part["iters"], part["secs"]
        merged.update(part["errors"])
    return out, it, sc, merged
