#!/usr/bin/env python3
"""Parameter sweep over laser FWHM, sample conductivity and sample width on the GPU.

Drop-in for the reference ``parameter_sweep.py`` (same functions, arguments, CLI flags, run-folder
naming and summary files: parameter_sweep.py:56-121 watchers, :123-192 single run, :195-268 grid /
cfg editing, :290-540 driver, :543-604 CLI).  What differs is how the runs execute:

  reference                                      here
  ---------------------------------------------  ---------------------------------------------
  one spawned CPU process per parameter set,     variants of a width group (same mesh) advance in
  each re-reading the mesh, re-assembling and    tiles of ``batch`` simulations through the batched
  LU-factorising (``mp.Pool``)                   multi-RHS CUDA kernels (``mode='ensemble'``) or, when
                                                 the mesh fits on chip, one after the other on the
                                                 resident mesh (``mode='serial'``; ``'auto'``, the
                                                 default, picks); tiles are sharded over the GPUs
                                                 (torchrun ranks, or ``num_processes`` spawned GPU
                                                 workers); every rank writes the run folders of its own
                                                 variants while its GPU works on the next ones, and one
                                                 final gather brings the per-run summary to rank 0
  ---------------------------------------------  ---------------------------------------------
  ``mode='per_run'`` (forced by ``write_xdmf=True``) keeps the reference's behaviour of calling
  ``run_simulation`` once per parameter set, which also writes the XDMF and gradient CSV files.

The runner follows the cfg: ``run_with_diamond`` when the cfg has diamond anvils (``p_diam``),
``run_no_diamond`` otherwise (the reference imports ``run_no_diamond`` only, but its own
``get_watcher_points`` already handles both stacks, parameter_sweep.py:85-103).
"""
import argparse
import copy
import itertools
import json
import multiprocessing as mp
import os
import queue
import sys
import threading
import time
from datetime import datetime

import numpy as np
import pandas as pd
import yaml

from . import problem, sweep
from .runners import Simulation2D, suppress_output


def set_single_thread():
    """Set environment variables to ensure single-threaded host libraries (parameter_sweep.py:46-53)."""
    for var in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'VECLIB_MAXIMUM_THREADS',
                'NUMEXPR_NUM_THREADS', 'BLIS_NUM_THREADS'):
        os.environ[var] = '1'


def initialize_worker():
    """Worker initialiser (parameter_sweep.py:56-66; there is no MPI to initialise here)."""
    set_single_thread()


def get_watcher_points(cfg):
    """Watcher points halfway through the iridium coupler layers, ``{'pside': (z, r), 'oside': (z, r)}``
    (parameter_sweep.py:69-120)."""
    z_sample = float(cfg['mats']['p_sample']['z'])
    z_ins_pside = float(cfg['mats']['p_ins']['z'])
    z_ins_oside = float(cfg['mats']['o_ins']['z'])
    z_coupler = float(cfg['mats']['p_coupler']['z'])
    if 'p_diam' in cfg['mats']:
        z_diam = float(cfg['mats']['p_diam']['z'])
        mesh_zmin = -(z_sample / 2) - z_ins_pside - z_coupler - z_diam
        mesh_zmax = (z_sample / 2) + z_ins_oside + z_coupler + z_diam
        bnd_p_ins_end = mesh_zmin + z_diam + z_ins_pside
        bnd_o_ins_start = mesh_zmax - z_diam - z_ins_oside
    else:
        mesh_zmin = -(z_sample / 2) - z_ins_pside - z_coupler
        mesh_zmax = (z_sample / 2) + z_ins_oside + z_coupler
        bnd_p_ins_end = mesh_zmin + z_ins_pside
        bnd_o_ins_start = mesh_zmax - z_ins_oside
    return {'pside': (bnd_p_ins_end + z_coupler / 2, 0.0), 'oside': (bnd_o_ins_start - z_coupler / 2, 0.0)}


def _runner_for(cfg):
    if 'p_diam' in cfg['mats']:
        from .run_with_diamond import run_simulation
        return run_simulation, problem.stack_with_diamond
    from .run_no_diamond import run_simulation
    return run_simulation, problem.stack_no_diamond


def run_name_for(fwhm, k, width):
    """Run folder name (parameter_sweep.py:145)."""
    return f"fwhm_{fwhm:.2e}_k_{k:.2f}_width_{width:.2e}".replace('+', '').replace('-0', '-')


def run_single_simulation(args):
    """One parameter set through ``run_simulation`` (parameter_sweep.py:123-192).

    args = (combo, base_config, mesh_folder, output_dir, write_xdmf, suppress_print, run_id)
    """
    set_single_thread()
    combo, base_config, mesh_folder, output_dir, write_xdmf, suppress_print, run_id = args
    fwhm, k, width = combo['fwhm'], combo['k'], combo['width']
    run_name = run_name_for(fwhm, k, width)
    run_output_dir = os.path.join(output_dir, run_name)
    config = modify_config_for_parameters(base_config, fwhm, k, width)
    watcher_points = get_watcher_points(config)
    result = {'run_id': run_id, 'run_name': run_name, 'fwhm': fwhm, 'k': k, 'width': width,
              'output_dir': run_output_dir}
    try:
        start_time = time.time()
        run_simulation, _ = _runner_for(config)
        run_simulation(cfg=config, mesh_folder=mesh_folder, rebuild_mesh=False, visualize_mesh=False,
                       output_folder=run_output_dir, watcher_points=watcher_points, write_xdmf=write_xdmf,
                       suppress_print=suppress_print)
        result.update(runtime=time.time() - start_time, status='success', error=None)
    except Exception as e:
        result.update(runtime=0.0, status='failed', error=str(e))
    return result


def create_parameter_grid(fwhm_range, k_range, width_range, num_points):
    """Log-spaced FWHM and k, linearly spaced width, grouped by width (parameter_sweep.py:195-237)."""
    fwhm_min, fwhm_max = fwhm_range
    k_min, k_max = k_range
    width_min, width_max = width_range
    num_fwhm, num_k, num_width = num_points
    fwhm_vals = np.logspace(np.log10(fwhm_min), np.log10(fwhm_max), num_fwhm)
    k_vals = np.logspace(np.log10(k_min), np.log10(k_max), num_k)
    width_vals = np.linspace(width_min, width_max, num_width)
    parameter_combinations = []
    for width in width_vals:
        for fwhm, k in itertools.product(fwhm_vals, k_vals):
            parameter_combinations.append({'fwhm': fwhm, 'k': k, 'width': width})
    return parameter_combinations, fwhm_vals, k_vals, width_vals


def modify_config_for_parameters(base_config, fwhm, k, width):
    """Copy of ``base_config`` with heating.fwhm, mats.p_sample.k and mats.p_sample.z replaced
    (parameter_sweep.py:240-268)."""
    config = copy.deepcopy(base_config)
    config['heating']['fwhm'] = float(fwhm)
    config['mats']['p_sample']['k'] = float(k)
    config['mats']['p_sample']['z'] = float(width)
    return config


def get_mesh_folder_for_width(base_mesh_folder, width):
    """``<base>/width_<w>`` (parameter_sweep.py:271-288)."""
    width_str = f"{width:.3e}".replace('+', '').replace('-0', '-')
    return os.path.join(base_mesh_folder, f"width_{width_str}")


# ----------------------------------------------------------------------------------------
# ensemble execution of one width group
# ----------------------------------------------------------------------------------------
class _OutputWriter:
    """Writes the run folders (``used_config.yaml`` + ``watcher_points.csv``, parameter_sweep.py:145-181 via
    run_simulation) of finished variants on a host thread, so the files of one variant are written while the GPU
    advances the next ones.  ``submit`` is the ``on_done`` callback of the sweep engine."""

    def __init__(self, output_dir, base_config, combinations, step_t, names):
        self.output_dir, self.base_config, self.combinations = output_dir, base_config, combinations
        self.step_t, self.names = np.asarray(step_t), list(names)
        self.errors = {}
        self._q = queue.Queue()
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()

    def submit(self, i, hist, iters, seconds, error):
        if error is None:
            self._q.put((int(i), np.array(hist, copy=True)))

    def _loop(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            i, hist = item
            try:
                if not np.all(np.isfinite(hist)):
                    raise FloatingPointError('non-finite watcher history')
                _write_run_outputs(self.output_dir, self.base_config, self.combinations[i], self.step_t, hist, self.names)
            except Exception as exc:                   # recorded per run, as the reference does
                self.errors[i] = str(exc)

    def close(self):
        self._q.put(None)
        self._t.join()
        return self.errors


def _group_on_device(base_config, combinations, mesh_folder, batch, device, tiles, suppress_print, engine="auto",
                     output_dir=None, names=None, share=None, on_ready=None):
    """Set the group's mesh up on ``device``, run this rank's tiles and (``output_dir`` given) write their run
    folders.  Returns (idx, hist, iters, secs, errors, step_times).  ``share`` = (rank, world): when the serial
    engine runs the group, the rank takes every world-th variant of the k-sorted list instead of its tiles - tiles
    of 16 cut a conductivity's heating widths into a narrow and a wide half, and the wide halves (more PCG
    iterations) all land on the odd ranks (measured: 17.7 s against 19.7 s on two GPUs)."""
    cfg0 = modify_config_for_parameters(base_config, combinations[0]['fwhm'], combinations[0]['k'], combinations[0]['width'])
    _, stack = _runner_for(cfg0)
    t_setup = time.time()
    timing = bool(os.environ.get("HF_SWEEP_TIMING"))
    # serial engine on an on-chip mesh: two simulations share the SMs (hf_set_sharing) when the mesh still fits with
    # half the registers / shared memory per CTA - planned for that from the start (re-planning costs 0.3 s)
    n_mine = sum(len(t) for t in tiles)
    want_pair = engine in ("auto", "serial") and (n_mine > 1 or share is not None)
    with suppress_output(suppress_print):
        sim = Simulation2D(cfg0, stack, mesh_folder, rebuild_mesh=False, device=device, sharing=2 if want_pair else 1)
    extra = []
    writer = None
    try:
        if want_pair:
            with suppress_output(suppress_print):
                if sim.solver.on_chip():
                    extra.append(Simulation2D(cfg0, stack, mesh_folder, rebuild_mesh=False, device=device, sharing=2))
                else:
                    sim.set_sharing(1)
        watch = sim.watcher_nodes(list(get_watcher_points(cfg0).values()))
        fwhm = np.array([c['fwhm'] for c in combinations])
        k = np.array([c['k'] for c in combinations])
        if share is not None and (engine == "serial" or (engine == "auto" and sim.solver.on_chip())):
            tiles = [serial_share(k, share[0], share[1])]
            n_mine = len(tiles[0])
        if output_dir is not None:
            writer = _OutputWriter(output_dir, base_config, combinations, sim.step_t.copy(), names)
        t_run = time.time()
        if on_ready is not None:
            on_ready()                                  # multi-rank sweeps: NCCL warm-up overlapping the first simulations
        idx, hist, iters, secs, errors = sweep.run_tiles(sim, fwhm, k, tiles, watch, engine=engine, extra_sims=extra,
                                                         on_done=writer.submit if writer else None)
        t_write = time.time()
        if writer is not None:
            errors = dict(errors)
            errors.update(writer.close())
            writer = None
        if timing:
            print(f"[sweep timing] device {device}: set-up {t_run - t_setup:.2f}s, {n_mine} simulations {t_write - t_run:.2f}s, "
                  f"output drain {time.time() - t_write:.2f}s", file=sys.stderr, flush=True)
        return idx, hist, iters, secs, errors, sim.step_t.copy()
    finally:
        if writer is not None:
            writer.close()
        sim.close()
        for e in extra:
            e.close()


def serial_share(k_values, rank, world):
    """Variants of ``rank`` for the serial engine: position p of the k-sorted list goes to rank (p + p // world) % world
    - one variant per rank out of every ``world`` consecutive ones, the offset rotating so that no rank always gets
    the widest heating profile of its group."""
    order = np.argsort(np.asarray(k_values, dtype=np.float64), kind="stable")
    pos = np.arange(len(order))
    return order[(pos + pos // world) % world == rank]


def _rank_share(base_config, combinations, mesh_folder, batch, device, tiles, suppress_print, engine, output_dir, names,
                share=None):
    """``_group_on_device`` that never raises: a failure outside the per-variant handlers (set-up, out of memory)
    marks every variant of this rank as failed, so that the rank still takes part in the final gather."""
    mine = np.concatenate([np.asarray(t, dtype=np.int64) for t in tiles]) if len(tiles) else np.zeros(0, np.int64)
    warm = []

    def on_ready():
        warm.append(_warm_up_collectives(device))

    try:
        idx, _hist, iters, secs, errors, _ = _group_on_device(base_config, combinations, mesh_folder, batch, device, tiles,
                                                            suppress_print, engine, output_dir, names, share=share,
                                                            on_ready=on_ready if share is not None else None)
        return idx, iters, secs, errors
    except Exception as exc:
        if share is not None:
            mine = serial_share([c['k'] for c in combinations], share[0], share[1]) if engine == "serial" else mine
        return mine, np.full(len(mine), -1, dtype=np.int64), np.zeros(len(mine)), {int(i): str(exc) for i in mine}
    finally:
        if share is not None:
            if not warm:                                # set-up failed before the warm-up: every rank still takes part
                on_ready()
            warm[0].join()


def _mesh_marker(mesh_folder):
    """Marker file next to (not inside) the mesh folder, which keeps the reference's two files only."""
    return os.path.normpath(mesh_folder) + ".ready"


def _wait_for_mesh(mesh_folder, mesh_file, mesh_cfg_file, timeout_s=7200.0):
    """Ranks > 0: block until rank 0 has marked the group's mesh files complete (``<mesh folder>.ready``)."""
    marker = _mesh_marker(mesh_folder)
    t0 = time.time()
    while not (os.path.exists(marker) and os.path.exists(mesh_file) and os.path.exists(mesh_cfg_file)):
        if time.time() - t0 > timeout_s:
            raise TimeoutError(f"mesh files of rank 0 did not appear in {mesh_folder} within {timeout_s:.0f} s")
        time.sleep(0.02)


def _warm_up_collectives(local_rank):
    """First use of a NCCL communicator costs ~1.3 s (measured): do it on a side thread while the first simulations
    run, so the final gather - the sweep's one data-carrying collective - finds it ready.  (Started during the device
    set-up instead, it slowed that down from 0.6 s to 3.6 s.)"""
    def work():
        try:
            import torch
            import torch.distributed as dist
            if not (dist.is_available() and dist.is_initialized()):
                return                               # spawned workers without a process group: nothing to warm up
            if dist.get_backend() == "nccl":
                torch.cuda.set_device(local_rank)
            # the same collective as the final gather (sweep.gather_results), so that its connections exist
            bucket = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
            dist.gather_object({"warm": dist.get_rank()}, bucket, dst=0)
        except Exception:
            pass                                     # the gather itself will report a broken process group
    t = threading.Thread(target=work, daemon=True)
    t.start()
    return t


def _device_worker(args):
    """Spawned GPU worker (one per device) for sweeps launched without torchrun."""
    set_single_thread()
    return _rank_share(*args)


def _write_run_outputs(output_dir, base_config, combo, step_t, hist, names):
    run_dir = os.path.join(output_dir, run_name_for(combo['fwhm'], combo['k'], combo['width']))
    os.makedirs(run_dir, exist_ok=True)
    config = modify_config_for_parameters(base_config, combo['fwhm'], combo['k'], combo['width'])
    with open(os.path.join(run_dir, 'used_config.yaml'), 'w') as f:
        yaml.dump(config, f, Dumper=getattr(yaml, 'CSafeDumper', yaml.SafeDumper))
    df = pd.DataFrame({'time': step_t})
    for w, name in enumerate(names):
        df[name] = hist[:, w]
    df.to_csv(os.path.join(run_dir, 'watcher_points.csv'), index=False)
    return run_dir


def _visible_gpus():
    from . import _lib
    return int(_lib.load().hf_device_count())


def run_parameter_sweep(base_config_path, output_dir, fwhm_range, k_range, width_range, num_points,
                        base_mesh_folder="meshes", write_xdmf=False, suppress_print=True, num_processes=None,
                        mode="auto", batch=16):
    """Run the sweep (parameter_sweep.py:290-540).  Returns ``(results, failed_runs)`` on rank 0
    (``([], [])`` on the other torchrun ranks).

    ``num_processes``: GPU worker processes when not launched under torchrun (default: every
    visible GPU).  ``mode``: 'ensemble' (batched kernels, ``batch`` variants per tile), 'serial' (one
    simulation after the other on the resident mesh), 'auto' (serial when the mesh fits on chip, else
    ensemble) or 'per_run' (one ``run_simulation`` call per parameter set, as the reference).
    """
    set_single_thread()
    if mode not in ("auto", "ensemble", "serial", "per_run"):
        raise ValueError("mode must be 'auto', 'ensemble', 'serial' or 'per_run'")
    if write_xdmf:
        mode = "per_run"                       # field output needs every state on the host
    if not 1 <= int(batch) <= 32:
        raise ValueError("batch must be in [1, 32]")
    rank, world, local_rank = sweep.dist_info()

    with open(base_config_path, 'r') as f:
        base_config = yaml.safe_load(f)
    parameter_combinations, fwhm_vals, k_vals, width_vals = create_parameter_grid(fwhm_range, k_range, width_range, num_points)

    n_gpus = _visible_gpus()
    if n_gpus < 1:
        raise RuntimeError("parameter_sweep: no CUDA device (heatflow_b200 has no CPU fallback)")
    n_workers = 1 if world > 1 else max(1, min(n_gpus, num_processes or n_gpus))

    if rank == 0:
        os.makedirs(output_dir, exist_ok=True)
        sweep_metadata = {
            'base_config': base_config_path, 'fwhm_range': fwhm_range, 'k_range': k_range, 'width_range': width_range,
            'num_points': num_points, 'fwhm_values': fwhm_vals.tolist(), 'k_values': k_vals.tolist(),
            'width_values': width_vals.tolist(), 'total_runs': len(parameter_combinations),
            'num_processes': world if world > 1 else n_workers, 'timestamp': datetime.now().isoformat(),
            'execution': {'mode': mode, 'batch': int(batch), 'gpus': world if world > 1 else n_workers},
            'watcher_points': {
                'description': 'Temperature monitoring points positioned halfway through iridium coupler layers',
                'locations': {'pside': 'Center of p-side iridium coupler (r=0)', 'oside': 'Center of o-side iridium coupler (r=0)'},
                'coordinates': 'Relative to mesh geometry, calculated for each parameter combination'}}
        with open(os.path.join(output_dir, 'sweep_metadata.json'), 'w') as f:
            json.dump(sweep_metadata, f, indent=2)

    width_groups = {}
    for combo in parameter_combinations:
        width_groups.setdefault(combo['width'], []).append(combo)

    results, failed_runs = [], []
    total_completed = 0
    say = print if rank == 0 else (lambda *a, **k: None)
    say(f"Starting parameter sweep with {len(parameter_combinations)} total runs")
    say(f"Parameters: {len(fwhm_vals)} FWHM values, {len(k_vals)} k values, {len(width_vals)} width values")
    say(f"Grouped into {len(width_groups)} width groups for mesh reuse")
    say(f"Using {world if world > 1 else n_workers} GPU(s), mode={mode}, batch={batch}")
    say(f"Output directory: {output_dir}")
    say("Watcher points: Temperature monitoring at iridium coupler centers (pside, oside)")
    say("-" * 80)

    names = list(get_watcher_points(base_config).keys())
    # tiled modes: per-variant summaries of this rank over all width groups, gathered ONCE at the end
    my_idx, my_iters, my_secs, my_errors = [], [], [], {}
    group_offset, offset = [], 0
    t_sweep = time.time()
    warm = None
    for width_idx, (width, combinations) in enumerate(width_groups.items()):
        say(f"\nProcessing width group {width_idx + 1}/{len(width_groups)}: width = {width:.2e} m")
        say(f"  {len(combinations)} runs for this width")
        mesh_folder = get_mesh_folder_for_width(base_mesh_folder, width)
        mesh_file = os.path.join(mesh_folder, 'mesh.msh')
        mesh_cfg_file = os.path.join(mesh_folder, 'mesh_cfg.yaml')
        if rank == 0:
            os.makedirs(mesh_folder, exist_ok=True)
            if not (os.path.exists(mesh_file) and os.path.exists(mesh_cfg_file)):
                if os.path.exists(_mesh_marker(mesh_folder)):
                    os.remove(_mesh_marker(mesh_folder))
                say(f"  Building new mesh for width {width:.2e} m")
                config = modify_config_for_parameters(base_config, combinations[0]['fwhm'], combinations[0]['k'], width)
                _, stack = _runner_for(config)
                from .runners import prepare_mesh
                with suppress_output(suppress_print):
                    prepare_mesh(config, stack, mesh_folder, rebuild_mesh=True)
            else:
                say(f"  Reusing existing mesh for width {width:.2e} m")
            with open(_mesh_marker(mesh_folder), 'w') as f:      # both files are complete
                f.write("ok\n")
        elif world > 1:
            # the other ranks wait for rank 0's mesh files through the file system - no collective, so the NCCL
            # communicator is set up in the background (below) while the devices are already working
            _wait_for_mesh(mesh_folder, mesh_file, mesh_cfg_file)
        if os.environ.get("HF_SWEEP_TIMING"):
            print(f"[sweep timing] rank {rank}: mesh ready after {time.time() - t_sweep:.2f}s", file=sys.stderr, flush=True)

        if mode == "per_run":
            # the reference's own scheme: one run_simulation per parameter set (this rank's share)
            mine = list(range(len(combinations)))[rank::world]
            local = [run_single_simulation((combinations[i], base_config, mesh_folder, output_dir, write_xdmf, suppress_print,
                                            total_completed + i + 1)) for i in mine]
            if world > 1:
                import torch.distributed as dist
                bucket = [None] * world if rank == 0 else None
                dist.gather_object(local, bucket, dst=0)
                local = [r for part in bucket for r in part] if rank == 0 else []
            for result in sorted(local, key=lambda r: r['run_id']):
                total_completed += 1
                (results if result['status'] == 'success' else failed_runs).append(result)
                _report(say, result, total_completed, len(parameter_combinations))
            if rank != 0:
                total_completed += len(combinations)
            continue

        tiles = sweep.plan_tiles([c['k'] for c in combinations], int(batch), world if world > 1 else n_workers)
        say(f"  Starting {len(combinations)} simulations in {sum(len(t) for t in tiles)} tile(s) of <= {batch}...")
        t_group = time.time()
        if world > 1 or n_workers == 1:
            parts = [_rank_share(base_config, combinations, mesh_folder, batch, local_rank if world > 1 else 0, tiles[rank],
                                 suppress_print, mode, output_dir, names, (rank, world) if world > 1 else None)]
        else:
            if mp.get_start_method(allow_none=True) != 'spawn':
                try:
                    mp.set_start_method('spawn', force=True)
                except RuntimeError:
                    pass
            jobs = [(base_config, combinations, mesh_folder, batch, d, tiles[d], suppress_print, mode, output_dir, names,
                     (d, n_workers)) for d in range(n_workers)]
            with mp.Pool(processes=n_workers, initializer=initialize_worker) as pool:
                parts = pool.map(_device_worker, jobs)
        for idx_p, it_p, sec_p, err_p in parts:
            my_idx.append(np.asarray(idx_p, dtype=np.int64) + offset)
            my_iters.append(np.asarray(it_p, dtype=np.int64))
            my_secs.append(np.asarray(sec_p, dtype=np.float64))
            my_errors.update({int(i) + offset: e for i, e in err_p.items()})
        say(f"  group finished on rank 0 in {time.time() - t_group:.2f}s")
        group_offset.append((offset, combinations))
        offset += len(combinations)

    if mode != "per_run":
        cat = lambda parts, dt: np.concatenate(parts) if parts else np.zeros(0, dt)
        if os.environ.get("HF_SWEEP_TIMING"):
            print(f"[sweep timing] rank {rank}: at the final gather after {time.time() - t_sweep:.2f}s", file=sys.stderr, flush=True)
        gathered = sweep.gather_results(offset, 0, 0, cat(my_idx, np.int64), None, cat(my_iters, np.int64),
                                        cat(my_secs, np.float64), my_errors)       # the single collective of the sweep
        if os.environ.get("HF_SWEEP_TIMING"):
            print(f"[sweep timing] rank {rank}: gathered after {time.time() - t_sweep:.2f}s", file=sys.stderr, flush=True)
        if rank == 0:
            _, iters, secs, errors = gathered
            for off, combinations in group_offset:
                for i, combo in enumerate(combinations):
                    total_completed += 1
                    run_name = run_name_for(combo['fwhm'], combo['k'], combo['width'])
                    result = {'run_id': total_completed, 'run_name': run_name, 'fwhm': combo['fwhm'], 'k': combo['k'],
                              'width': combo['width'], 'output_dir': os.path.join(output_dir, run_name)}
                    g = off + i
                    if g in errors or iters[g] < 0:
                        result.update(runtime=0.0, status='failed', error=errors.get(g, 'run did not report back'))
                        failed_runs.append(result)
                    else:
                        result.update(runtime=float(secs[g]), status='success', error=None)
                        results.append(result)
                    _report(say, result, total_completed, len(parameter_combinations))
            say(f"\nall groups finished in {time.time() - t_sweep:.2f}s")
    if os.environ.get("HF_SWEEP_TIMING"):
        print(f"[sweep timing] rank {rank}: gather + summary done after {time.time() - t_sweep:.2f}s", file=sys.stderr, flush=True)

    if rank != 0:
        return [], []
    results_df = pd.DataFrame(results)
    if not results_df.empty:
        results_df.to_csv(os.path.join(output_dir, 'successful_runs.csv'), index=False)
    failed_df = pd.DataFrame(failed_runs)
    if not failed_df.empty:
        failed_df.to_csv(os.path.join(output_dir, 'failed_runs.csv'), index=False)

    print("\n" + "=" * 80)
    print("PARAMETER SWEEP COMPLETE")
    print("=" * 80)
    print(f"Total runs: {len(parameter_combinations)}")
    print(f"Successful: {len(results)}")
    print(f"Failed: {len(failed_runs)}")
    print(f"Results saved to: {output_dir}")
    if results:
        avg_runtime = np.mean([r['runtime'] for r in results])
        total_runtime = sum(r['runtime'] for r in results)
        print(f"Average runtime per simulation: {avg_runtime:.2f}s")
        print(f"Total simulation time: {total_runtime:.2f}s")
    wall = time.time() - t_sweep
    print(f"Sweep wall time: {wall:.2f}s ({len(parameter_combinations) / max(wall, 1e-9):.2f} simulations/s on "
          f"{world if world > 1 else n_workers} GPU(s), mesh handling, run folders and the final gather included)")
    return results, failed_runs


def _report(say, result, done, total):
    head = f"[{done}/{total}] FWHM={result['fwhm']:.2e}m, k={result['k']:.2f}W/m/K, width={result['width']:.2e}m"
    if result['status'] == 'success':
        say(f"  ✓ {head} - Completed in {result['runtime']:.2f}s")
    else:
        say(f"  ✗ {head} - Failed: {result['error']}")


def main():
    """Command-line interface (parameter_sweep.py:543-604) plus --mode / --batch."""
    parser = argparse.ArgumentParser(description='Parameter sweep for heatflow simulations')
    parser.add_argument('--config', type=str, required=True, help='Path to base configuration file')
    parser.add_argument('--output-dir', type=str, required=True, help='Directory to save all results')
    parser.add_argument('--fwhm-range', type=float, nargs=2, default=[1e-6, 1e-4], help='FWHM range in meters (min max)')
    parser.add_argument('--k-range', type=float, nargs=2, default=[1.0, 100.0],
                        help='Thermal conductivity range in W/m/K (min max)')
    parser.add_argument('--width-range', type=float, nargs=2, default=[1e-6, 10e-6],
                        help='Sample width range in meters (min max)')
    parser.add_argument('--num-points', type=int, nargs=3, default=[5, 5, 3],
                        help='Number of points for each parameter (fwhm k width)')
    parser.add_argument('--mesh-folder', type=str, default='meshes', help='Base directory for mesh storage')
    parser.add_argument('--write-xdmf', action='store_true', help='Write XDMF output files (forces --mode per_run)')
    parser.add_argument('--verbose', action='store_true', help='Show detailed output during simulations')
    parser.add_argument('--num-processes', type=int, default=None, help='Number of GPU workers (default: all visible GPUs)')
    parser.add_argument('--mode', choices=['auto', 'ensemble', 'serial', 'per_run'], default='auto')
    parser.add_argument('--batch', type=int, default=16, help='Variants per ensemble tile (1..32)')
    args = parser.parse_args()

    if any(x <= 0 for x in args.num_points):
        parser.error("Number of points must be positive")
    if args.fwhm_range[0] <= 0 or args.fwhm_range[1] <= 0:
        parser.error("FWHM range must be positive")
    if args.k_range[0] <= 0 or args.k_range[1] <= 0:
        parser.error("Thermal conductivity range must be positive")
    if args.width_range[0] <= 0 or args.width_range[1] <= 0:
        parser.error("Width range must be positive")
    if args.num_processes is not None and args.num_processes <= 0:
        parser.error("Number of processes must be positive")

    # under torchrun every rank runs this script; join the process group for the final gather
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    try:
        run_parameter_sweep(base_config_path=args.config, output_dir=args.output_dir, fwhm_range=tuple(args.fwhm_range),
                            k_range=tuple(args.k_range), width_range=tuple(args.width_range),
                            num_points=tuple(args.num_points), base_mesh_folder=args.mesh_folder,
                            write_xdmf=args.write_xdmf, suppress_print=not args.verbose, num_processes=args.num_processes,
                            mode=args.mode, batch=args.batch)
    finally:
        if int(os.environ.get("WORLD_SIZE", "1")) > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()
