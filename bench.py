#!/usr/bin/env python
"""Benchmark of the heat-conduction hot path (contract: see the round prompt / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A *step* is one backward-Euler time step (Gaussian BC update, RHS, Jacobi-PCG solve) of
``cfgs/geballe_with_diamond.yaml`` on the mesh at the cfg's own sizes (N ~ 1.4e5 dofs).  Both arms run the same
window: W warm-up steps from T = 300 K (steps 0..W-1 of the simulation), then K timed steps (W..W+K-1).
At N GPUs every rank runs the same simulation - no data-path collective, one final gather of the watcher
histories ("weak").
``value`` = DOF-timesteps/s of the whole job with the state resident in HBM, device-timed;
``e2e`` = the same through the C-ABI with host buffers (state upload, amplitudes in, watcher
histories and final field out) inside the timed region.

Extra objects on the JSON line:
  roofline      dominant kernel of the timed region.  At the cfg's own size the mesh fits on chip and the whole
                solve runs in ``k_pcg_pipe`` / ``k_pcg_patch``: bound by the latency of its grid reduction, so the object says
                ``bound: "latency"`` and reports its real DRAM rate and the time per PCG iteration.
  roofline_1m   ``k_pcg_stream`` (persistent streaming kernel, one cooperative launch per solve) on the >= 1 M-dof
                refinement of ``cfgs/konopkova.yaml`` (BASELINE config #4), timed inside real solves;
                ``roofline_1m_launch_per_iteration``: the same with ``k_pcg_iter`` (round-1 scheme).
  roofline_4m   the same on a 4.3 M-dof refinement: 580 MB per iteration, far beyond the L2.
  roofline_ens_1m  ``k_ens_iter`` (batched multi-RHS ensemble, 16 sweep variants per tile) on the same >= 1 M-dof mesh.
  mid_mesh / mid_mesh_220k  the headline cfg refined to 3.8e5 / 2.2e5 dofs (the sizes of the reference's own gmsh meshes).
  sweep         BASELINE config #5 through ``run_parameter_sweep``: 128 variants per GPU, run folders written;
                ``config.sweep_sims_per_s`` / ``config.sweep_cpu_sims_per_s`` repeat the two numbers.
  cpu_baseline  the scipy sparse-LU oracle on this host (1 core), same steps; ``parity_check`` compares the GPU
                result of the timed steps with it and fails the run above 1e-10.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORKLOAD = "geballe_with_diamond"
METRIC = "DOF-timesteps/sec, geballe_with_diamond"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons during the timed region, sampled every 20 ms through NVML
    (the same counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints; a piped
    `nvidia-smi -lms` loses its buffered lines when it is terminated)."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, index=0):
        self.index, self.sm, self.reasons, self.sm_max = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None

    def _visible_to_physical(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self._visible_to_physical())
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self._stop.is_set():
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                try:
                    mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
                self._stop.wait(0.02)
            pynvml.nvmlShutdown()
        except Exception:
            pass

    def __enter__(self):
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=2)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "how": "NVML, every 20 ms over the timed passes"}


def variant_params(case, index):
    """Sweep variant `index` of the 64 x 64 (k, fwhm) grid of config #5 (parameter_sweep.py:221-222)."""
    if index == 0:
        return None                                    # the cfg's own k_sample / fwhm
    ks = np.logspace(0.0, 2.0, 64)
    fw = np.logspace(-6.0, -4.0, 64)
    return float(ks[(index * 7) % 64]), float(fw[(index * 11) % 64])


def build(index, size_scale=1.0):
    from helpers import build_case
    from heatflow_b200 import problem
    c = build_case(WORKLOAD, size_scale)
    v = variant_params(c, index)
    if v is not None:
        k, fwhm = v
        c.kappa_t = c.kappa_t.copy()
        c.kappa_t[[m.name for m in c.mats].index("p_sample")] = k
        c.kappa_c = c.kappa_t[c.cell_tag - 1]
        c.fwhm, c.coeff = fwhm, problem.gaussian_coeff(fwhm)
    return c


def oracle_loop(c, steps, warmup, watch=None):
    """CPU arm: the scipy sparse-LU oracle on this host, 1 thread (SuperLU is sequential).  Times steps
    warmup .. warmup+steps-1 of the simulation; returns the watcher values of those steps as well."""
    from helpers import make_oracle
    t0 = time.perf_counter()
    O = make_oracle(c)
    t_asm = time.perf_counter() - t0
    t0 = time.perf_counter()
    O.factorize()
    t_fac = time.perf_counter() - t0
    for s in range(warmup):
        O.step((s + 1) * c.dt)
    hist = []
    t0 = time.perf_counter()
    for s in range(warmup, warmup + steps):
        u = O.step((s + 1) * c.dt)
        if watch is not None:
            hist.append(u[watch])
    t_loop = time.perf_counter() - t0
    return t_loop, t_asm, t_fac, O, np.array(hist)


def large_cpu_baseline(case, steps=3):
    """CPU arm on BASELINE config #4 (1.16 M dofs): one splu factorisation (~1 min, 4 GB) + `steps` time steps."""
    t_loop, t_asm, t_fac, _, _ = oracle_loop(case, steps, 0)
    n = len(case.nodes)
    return {"value": n * steps / t_loop, "unit": "DOF-timesteps/s", "cores": 1, "kind": "port",
            "sample": f"{steps} time steps on the N={n} mesh, scipy splu factorised once outside the loop "
                      f"(assembly {t_asm:.1f} s, factorisation {t_fac:.1f} s, {t_loop / steps:.2f} s per step)"}


def _reference_worker(job):
    index, steps, warmup = job
    os.environ["OMP_NUM_THREADS"] = "1"
    c = build(index)
    t_loop, t_asm, t_fac, _, _ = oracle_loop(c, steps, warmup)
    return len(c.nodes), t_loop, t_asm, t_fac


def _cpu_sweep_worker(job):
    """One sweep variant the way the reference runs it (parameter_sweep.py:123-192): assemble, factorise, time loop."""
    index, steps = job
    os.environ["OMP_NUM_THREADS"] = "1"
    c = build(index + 1)                                  # mesh re-used per width group in the reference: not timed
    t0 = time.perf_counter()
    oracle_loop(c, steps, 0)
    return time.perf_counter() - t0


def cpu_sweep_baseline(steps):
    """Sweep throughput of the CPU arm: one single-threaded process per variant on every host core (bounded sample)."""
    import multiprocessing as mp
    procs = max(1, min(os.cpu_count() or 1, 16))
    try:
        with mp.get_context("spawn").Pool(processes=procs) as pool:
            secs = pool.map_async(_cpu_sweep_worker, [(i, steps) for i in range(procs)]).get(timeout=240)
    except Exception as exc:                                # a reported baseline must never hang or fail the bench
        return {"sims_per_s": None, "cores": procs, "kind": "port", "sample": f"not measured: {type(exc).__name__}"}
    return {"sims_per_s": procs / max(secs), "cores": procs, "kind": "port",
            "sample": f"{procs} variants in {procs} single-threaded processes at once, each assembling, factorising (scipy splu) "
                      f"and running {steps} steps (mesh generation not timed); slowest process {max(secs):.1f} s"}


def run_reference(args, rank, world):
    """CPU arm: the oracle port (scipy sparse LU) on the box's host cores.  Like our arm at N GPUs it runs N
    independent simulations (sweep variant = index), one single-threaded process each - the reference's own
    execution model (parameter_sweep.py:46-66, :436-438: an mp.Pool of single-threaded workers)."""
    if rank != 0:
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    nsim = max(1, args.gpus)
    c0 = build(0)
    W = max(args.warmup, 3)
    steps = min(args.steps, c0.num_steps - W)
    jobs = [(0, steps, W) for _ in range(nsim)]        # the same simulation in every process, as in our arm
    if nsim == 1:
        res = [_reference_worker(jobs[0])]
    else:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(processes=min(nsim, os.cpu_count() or 1)) as pool:
            res = pool.map(_reference_worker, jobs)
    n = res[0][0]
    t_loop = max(r[1] for r in res)                    # the job ends when its slowest simulation does
    value = nsim * n * steps / t_loop
    cores = min(nsim, os.cpu_count() or 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "DOF-timesteps/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": W, "ms_per_step": 1e3 * t_loop / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD}: {nsim} independent copies of the same simulation, N={n} dofs each, cfg mesh sizes, "
                               f"in-repo mesher; time steps {W}..{W + steps - 1} of the run (the first {W} are the warm-up)"},
        "cpu_baseline": {"value": value, "unit": "DOF-timesteps/s", "cores": cores, "kind": "port",
                         "sample": f"{steps} time steps after {W} warm-up steps per simulation, {nsim} simulation(s) in "
                                   f"{cores} single-threaded process(es); scipy splu factorised once outside the timed loop "
                                   f"(assembly {res[0][2]:.2f} s, factorisation {res[0][3]:.2f} s); host has {os.cpu_count()} cores"},
        "e2e": {"value": value, "unit": "DOF-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# algorithmic bytes of one PCG iteration: fp64 values + 16-bit local columns,
# slice pointers, x r p q read and written once
def iter_bytes(n, nnz):
    return 10.0 * nnz + 4.0 * n / 32 + 64.0 * n


def configured_solver(case, device, rtol, warm=0.0, mode=0, ordering="auto", recycle=0, sharing=1):
    from heatflow_b200.solver import HeatSolver
    s = HeatSolver(device)
    s.set_sharing(sharing)
    s.set_ordering(ordering)
    s.set_mesh(case.nodes, case.tris, case.cell_tag)
    s.set_materials(case.tags, case.kappa_t, case.rhoc_t)
    s.set_bcs(case.bc_dofs, case.bc_value, case.gauss_slot, case.gauss_r)
    s.build_operator(case.dt, True)
    s.set_solver(rtol=rtol, warm=warm, mode=mode)
    s.set_recycle(recycle)
    return s


def streaming_roofline(case, device, rtol, peak, peak_src, steps, traffic=None, mode=2):
    """The streaming PCG kernel timed inside real solves: CUDA events on the solver stream around the PCG solve of
    every time step (hf_set_profile), divided by the PCG iterations performed inside the brackets.  mode 2: the
    persistent kernel k_pcg_stream (one cooperative launch per solve, the auto choice); mode 1: k_pcg_iter, one
    launch per iteration with the host polling for convergence."""
    s = configured_solver(case, device, rtol, mode=mode)
    n, nnz = s.sizes()
    s.set_state(np.full(n, case.ic))
    k0 = min(10, max(0, case.num_steps - steps - 1))
    s.run(case.amps[k0:k0 + 1], case.ic, case.coeff, [0])               # warm-up: graphs captured
    s.set_profile(True)
    _, iters, _ = s.run(case.amps[k0 + 1:k0 + 1 + steps], case.ic, case.coeff, [0])
    solve_ms, launches = s.solve_profile()
    s.set_profile(False)
    # per PCG ITERATION: the early-exit launches the host queues past convergence move no data, so dividing by
    # launches would flatter the kernel; their (small) cost is charged to the iterations instead
    us = solve_ms * 1e3 / max(1, int(iters.sum()))
    alg = iter_bytes(n, nnz)
    ms_flushed, _ = s.bench_kernels(reps=20, flush_l2=True)
    s.close()
    ach = alg / (us * 1e-6) / 1e9
    return {"bound": "hbm", "kernel": "k_pcg_stream" if mode == 2 else "k_pcg_iter", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic, "algorithmic_bytes_per_launch": alg, "launch_us": us, "launches": int(launches),
            "pcg_iterations": int(iters.sum()), "n_dofs": n, "nnz": nnz,
            "isolated_launch_us_l2_flushed": ms_flushed * 1e3, "isolated_frac_l2_flushed": alg / (ms_flushed * 1e-3) / 1e9 / peak,
            "peak_source": peak_src,
            "how": f"CUDA events on the solver stream around the PCG solves of {steps} time steps / PCG iterations performed "
                   "(mode 2: grid barriers included; mode 1: launch gaps and the early-exit launches queued past "
                   "convergence count against it); isolated_*: single k_pcg_iter launches (the same per-iteration work) "
                   "with the L2 flushed in between"}


def ensemble_roofline(case, device, rtol, peak, peak_src, B=16, steps=2, traffic=None):
    """north_star (c) on a mesh that streams from HBM: `k_ens_iter` advancing B sweep variants at once (one launch per PCG
    iteration of the whole tile), timed inside real solves.  Algorithmic bytes per launch: per dof
    and variant x z w p read + written (64 B), per row the diagonal pair (16 B), per non-zero two fp64 value arrays
    and a 16-bit local column (18 B) + 4 B row pointer, shared by the B variants."""
    from heatflow_b200 import problem
    s = configured_solver(case, device, rtol, warm=0.0, ordering="hilbert")
    n, nnz = s.sizes()
    s.set_state(np.full(n, case.ic))
    tag = int(case.tags[[m.name for m in case.mats].index("p_sample")])
    ks = np.logspace(0, 2, 64)[20:20 + B]
    cf = [problem.gaussian_coeff(f) for f in np.logspace(-6, -4, 64)[10:10 + B]]
    s.ens_create(ks, cf, tag)
    k0 = min(10, max(0, case.num_steps - steps - 1))
    s.ens_run(case.amps[k0:k0 + 1], case.ic, [0])                       # warm-up: graphs captured
    s.set_profile(True)
    _, iters = s.ens_run(case.amps[k0 + 1:k0 + 1 + steps], case.ic, [0])
    solve_ms, _ = s.solve_profile()
    s.set_profile(False)
    path = s.ens_path()
    s.ens_destroy()
    s.close()
    its = max(1, int(iters.sum()))
    us = solve_ms * 1e3 / its
    alg = n * (64.0 * B + 16.0) + 18.0 * nnz + 4.0 * n
    ach = alg / (us * 1e-6) / 1e9
    return {"bound": "hbm", "kernel": "k_ens_iter" if path == 1 else "k_ens_patch", "variants": B, "achieved": ach, "peak": peak,
            "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "algorithmic_bytes_per_launch": alg, "launch_us": us,
            "pcg_iterations": its, "n_dofs": n, "nnz": nnz, "us_per_iteration_and_variant": us / B,
            "dof_iterations_per_s": n * B / (us * 1e-6), "peak_source": peak_src,
            "how": f"CUDA events around the PCG solves of {steps} time steps of a {B}-variant tile / PCG iterations of the tile "
                   "(host polls between graph chunks and launches queued past convergence count against it)"}


def large_mesh_run(case, device, rtol, warm, recycle):
    """BASELINE config #4: the whole konopkova run on the >= 1 M-dof mesh with the runner defaults."""
    s = configured_solver(case, device, rtol, warm=warm, recycle=recycle)
    n, nnz = s.sizes()
    path_id = s.solver_path()
    s.set_state(np.full(n, case.ic))
    s.run(case.amps[:2], case.ic, case.coeff, [0])                      # warm-up: graphs captured
    s.set_state(np.full(n, case.ic))
    _, iters, _ = s.run(case.amps, case.ic, case.coeff, [0])
    ms = s.stats()["run_ms"]
    s.close()
    path = {1: "streaming kernel (launch per iteration)", 2: "persistent streaming kernel", 3: "on-chip patch kernel",
            4: "on-chip patch kernel (pipelined CG)"}[path_id]
    return {"workload": f"{case.name} refined, N={n} dofs, nnz={nnz}, {case.num_steps} steps, {path}, "
                        f"recycled initial guess {recycle} vectors", "value": n * case.num_steps / (ms * 1e-3),
            "unit": "DOF-timesteps/s", "ms_per_step": ms / case.num_steps, "pcg_iterations_total": int(iters.sum())}


def sweep_leg(args, rank, world, local_rank, barrier, per_gpu=128):
    """BASELINE config #5 through the real entry point: ``run_parameter_sweep`` (tiles sharded over the ranks, every
    rank writes the run folders of its own variants, one final gather) on ``per_gpu`` variants per GPU of the
    (fwhm, k) grid, the cfg's own mesh and all of its 100 steps.  Wall clock, max over ranks, outputs included;
    the mesh is generated by an untimed two-variant warm-up sweep (the reference re-uses meshes the same way)."""
    import shutil
    import tempfile
    import yaml
    from helpers import load_cfg
    import parameter_sweep as psw
    cfg = load_cfg(WORKLOAD)                                  # absolute heating-file path
    box = [tempfile.mkdtemp(prefix="hf_sweep_") if rank == 0 else None]
    if world > 1:
        import torch.distributed as dist
        dist.broadcast_object_list(box, src=0)
    tmp = box[0]
    cfg_path = os.path.join(tmp, "base.yaml")
    if rank == 0:
        with open(cfg_path, "w") as f:
            yaml.safe_dump(cfg, f)
    barrier()
    width = float(cfg["mats"]["p_sample"]["z"])
    meshes = os.path.join(tmp, "meshes")
    n_f, n_k = 16, (per_gpu // 16) * world
    psw.run_parameter_sweep(cfg_path, os.path.join(tmp, "warm"), (1e-6, 1e-4), (1.0, 100.0), (width, width), (1, 2 * world, 1),
                            base_mesh_folder=meshes)
    barrier()
    t0 = time.perf_counter()
    results, failed = psw.run_parameter_sweep(cfg_path, os.path.join(tmp, "out"), (1e-6, 1e-4), (1.0, 100.0), (width, width),
                                              (n_f, n_k, 1), base_mesh_folder=meshes)
    secs = time.perf_counter() - t0
    barrier()
    n_ok, n_bad = len(results), len(failed)
    n_dirs = len(os.listdir(os.path.join(tmp, "out"))) if rank == 0 else 0
    barrier()
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True)
    return secs, n_f * n_k, n_ok, n_bad, n_dirs, int(cfg["timing"]["num_steps"])


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from scipy.spatial import cKDTree
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    # stdout carries exactly one JSON line: everything else that writes to fd 1 (NCCL prints its version banner
    # there at NCCL_DEBUG=VERSION / WARN, the sweep prints its progress) goes to stderr for the whole run
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    c = build(0)                       # the SAME simulation on every rank: efficiency measures the machine, not variant imbalance
    n = len(c.nodes)
    W = max(args.warmup, 3)
    steps = min(args.steps, c.num_steps - W)
    s = configured_solver(c, local_rank, args.rtol, warm=args.warm_start, recycle=args.recycle)
    _, nnz = s.sizes()
    tree = cKDTree(c.nodes)
    watch = np.array([tree.query(p)[1] for p in [(c.heating_z + 0.5 * 6.2e-8, 0.0), (0.951e-6, 0.0)]], dtype=np.int32)
    u0 = np.full(n, c.ic)
    amps_w, amps_t = c.amps[:W], c.amps[W:W + steps]        # warm-up steps 0..W-1, timed steps W..W+steps-1 (both arms)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # warm-up: the first W steps of the simulation (kernels planned, clocks up); u_W is where the timed steps start
    s.set_state(u0)
    s.run(amps_w, c.ic, c.coeff, watch)
    u_w = s.get_state()
    s.set_state(u_w)
    s.run(amps_t, c.ic, c.coeff, watch)                      # one untimed pass over the timed steps as well

    with ClockSampler(local_rank) as clocks:
        # ---- timed region 1: state resident in HBM, device-timed (CUDA events on the solver stream)
        s.set_state(u_w)
        launches0 = s.stats()["launches"]
        barrier()
        hist, iters, _ = s.run(amps_t, c.ic, c.coeff, watch)
        barrier()
        st = s.stats()
        dev_ms = st["run_ms"]
        launches = st["launches"] - launches0

        # ---- timed region 2: end to end through the C-ABI with host buffers
        barrier()
        t0 = time.perf_counter()
        s.set_state(u_w)                                       # H2D: N*8 bytes
        hist2, iters2, _ = s.run(amps_t, c.ic, c.coeff, watch)  # H2D amps, D2H watcher history
        final = s.get_state()                                  # D2H: N*8 bytes
        e2e_s = time.perf_counter() - t0
        barrier()

        # ---- kernel timing pass for the roofline: CUDA events around every PCG solve (hf_set_profile)
        s.set_profile(True)
        s.set_state(u_w)
        _, iters_p, _ = s.run(amps_t, c.ic, c.coeff, watch)
        solve_ms, solve_launches = s.solve_profile()
        prof_run_ms = s.stats()["run_ms"]
        s.set_profile(False)
        for _ in range(4):                                     # a few more passes so that the sampler sees the load
            s.set_state(u_w)
            s.run(amps_t, c.ic, c.coeff, watch)
    path = s.solver_path()
    kernel = {1: "k_pcg_iter", 2: "k_pcg_stream", 3: "k_pcg_patch", 4: "k_pcg_pipe"}[path]
    retries = s.stats()["retries"]

    # ---- sweep leg (config #5) through run_parameter_sweep
    sweep = None
    if not args.skip_sweep:
        s.close()
        s = None
        sw_s, sw_n, sw_ok, sw_bad, sw_dirs, sw_steps = sweep_leg(args, rank, world, local_rank, barrier, args.sweep_per_gpu)
        t_sw = torch.tensor([sw_s], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_sw, op=dist.ReduceOp.MAX)
        sweep = {"sims_per_s": sw_n / float(t_sw[0]), "variants": sw_n, "variants_per_gpu": sw_n // world, "steps": sw_steps,
                 "seconds": float(t_sw[0]), "successful": sw_ok, "failed": sw_bad, "run_folders_written": sw_dirs,
                 "dof_timesteps_per_s": sw_n * n * sw_steps / float(t_sw[0]),
                 "how": "run_parameter_sweep(cfg, (1e-6, 1e-4), (1, 100), width, (16, 8 x n_gpus, 1)) with its defaults: wall "
                        "clock of the call (max over ranks) incl. mesh load, device set-up, operator re-assembly per "
                        "conductivity, used_config.yaml + watcher_points.csv of every run (written by the rank that ran "
                        "it) and the final gather; the mesh file was generated by an untimed warm-up sweep"}

    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gathered = [torch.zeros_like(torch.from_numpy(hist).cuda()) for _ in range(world)] if rank == 0 else None
        dist.gather(torch.from_numpy(hist).cuda(), gathered, dst=0)     # the single final gather
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        if s is not None:
            s.close()
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    traffic = {}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass
    share = solve_ms / prof_run_ms if prof_run_ms > 0 else None
    its_total = max(1, int(iters_p.sum()))
    if path >= 3:
        # On chip: operator and vectors live in registers / shared memory for the whole solve.  Not an HBM-bound kernel:
        # its limit is the latency of one grid-wide reduction per PCG iteration, so the line reports the DRAM rate it
        # really sustains and the time per PCG iteration; the HBM-bound kernel of the code base is in roofline_1m.
        us = solve_ms * 1e3 / steps
        dram = traffic.get(kernel)
        ach = (dram / (us * 1e-6) / 1e9) if dram else None
        roof = {"bound": "latency", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": (ach / peak) if ach else None, "traffic": dram, "launch_us": us,
                "pcg_iterations_per_launch": its_total / steps, "us_per_pcg_iteration": solve_ms * 1e3 / its_total,
                "hbm_equivalent_gbs": iter_bytes(n, nnz) * its_total / (solve_ms * 1e-3) / 1e9,
                "share_of_step_time": share, "peak_source": peak_src,
                "note": "latency-bound on-chip kernel (one cooperative launch per solve): `achieved` is its real DRAM rate "
                        "(ncu dram bytes of one launch / launch time, both in this object), which is tiny by design; "
                        "hbm_equivalent_gbs = bytes a streaming PCG iteration would move (10 nnz + 64 N) x iterations / "
                        "time, for comparison with roofline_1m only",
                "how": "CUDA events on the solver stream around every solver launch of a separate pass over the same "
                       "steps (hf_set_profile), summed / launches"}
    else:
        alg = iter_bytes(n, nnz)
        us = solve_ms * 1e3 / its_total
        roof = {"bound": "hbm", "kernel": kernel, "achieved": alg / (us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                "traffic": traffic.get(kernel), "algorithmic_bytes_per_launch": alg, "launch_us": us,
                "share_of_step_time": share, "peak_source": peak_src,
                "how": "CUDA events on the solver stream around the PCG solve of every time step / PCG iterations inside"}
        roof["frac"] = roof["achieved"] / peak
    line = {
        "metric": METRIC, "value": world * n * steps / (dev_ms * 1e-3), "unit": "DOF-timesteps/s", "n_gpus": world,
        "steps": steps, "warmup": W, "ms_per_step": dev_ms / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{WORKLOAD}: the same simulation on every GPU, N={n} dofs, nnz={nnz}, cfg mesh sizes, in-repo "
                               f"mesher; time steps {W}..{W + steps - 1} of the run (the first {W} are the warm-up, as in the "
                               f"reference arm), every timed pass starts from the host copy of the state after step {W - 1}; "
                               f"rtol={args.rtol:g}, warm start {args.warm_start:g}, recycled initial guess {args.recycle} "
                               f"vectors (runner defaults)",
                   "l2": "working set (~20 MB) is smaller than L2 and lives on chip; the >= 1 M-dof roofline run has a "
                         "~145 MB working set (> 126 MB L2)",
                   "pcg_iterations_total": int(iters.sum()), "pcg_iterations_max": int(iters.max()),
                   "solver_kernel": kernel, "runs_repeated_on_streaming_kernel": retries},
        "e2e": {"value": world * n * steps / (e2e_ms * 1e-3), "unit": "DOF-timesteps/s",
                "h2d_bytes_per_step": (n * 8 + steps * 8 + len(watch) * 4) / steps,
                "d2h_bytes_per_step": (n * 8 + steps * len(watch) * 8 + steps * 4) / steps},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": roof,
    }
    if sweep is not None:
        line["sweep"] = sweep
        line["config"]["sweep_sims_per_s"] = sweep["sims_per_s"]
        line["config"]["sweep_variants"] = sweep["variants"]
    # >= 1 M-dof mesh (north_star target for the SpMV roofline): BASELINE config #4, konopkova cfg refined x 0.35
    cl = None
    if not args.skip_large:
        from helpers import build_case
        cl = build_case("konopkova", 0.35)
        line["roofline_1m"] = streaming_roofline(cl, local_rank, args.rtol, peak, peak_src, steps=3,
                                                 traffic=traffic.get("k_pcg_stream_1m"))
        line["roofline_1m_launch_per_iteration"] = streaming_roofline(cl, local_rank, args.rtol, peak, peak_src, steps=2,
                                                                     traffic=traffic.get("k_pcg_iter_1m"), mode=1)
        line["roofline_ens_1m"] = ensemble_roofline(cl, local_rank, args.rtol, peak, peak_src, traffic=traffic.get("k_ens_iter_b16_1m"))
        line["konopkova_1m"] = large_mesh_run(cl, local_rank, args.rtol, args.warm_start, min(args.recycle, 64))
        # DRAM-honest: 4.3 M dofs, 580 MB per PCG iteration - nothing survives in the 126 MB L2 between iterations
        # (bandwidth measurement only: the fast 'rows' mesher - Delaunay of 4.3 M points would take minutes)
        line["roofline_4m"] = streaming_roofline(build_case("konopkova", 0.18, method="rows"), local_rank, args.rtol, peak, peak_src, steps=2,
                                                 traffic=traffic.get("k_pcg_stream_4m"))
        # the size of the reference's own gmsh meshes (2.1e5 - 4.3e5 nodes): still on chip
        line["mid_mesh"] = large_mesh_run(build_case(WORKLOAD, 0.6), local_rank, args.rtol, args.warm_start, args.recycle)
        line["mid_mesh_220k"] = large_mesh_run(build_case(WORKLOAD, 0.8), local_rank, args.rtol, args.warm_start, args.recycle)
    # CPU baseline on this host (bounded sample) and the parity check against it: the oracle runs the same steps
    if not args.skip_cpu:
        t_loop, t_asm, t_fac, O, ohist = oracle_loop(c, steps, W, watch)
        line["cpu_baseline"] = {"value": n * steps / t_loop, "unit": "DOF-timesteps/s", "cores": 1, "kind": "port",
                                "sample": f"time steps {W}..{W + steps - 1} of the same simulation after {W} warm-up steps, scipy "
                                          f"splu (SuperLU, sequential) factorised once outside the loop "
                                          f"(assembly {t_asm:.2f} s, factorisation {t_fac:.2f} s); host has {os.cpu_count()} cores"}
        err_hist = float(np.abs(hist / ohist - 1).max())
        err_field = float(np.abs(final / O.u - 1).max())
        line["parity_check"] = {"max_rel_err_watcher_history_vs_oracle": err_hist, "max_rel_err_final_field_vs_oracle": err_field,
                                "tolerance": 1e-10, "e2e_equals_device_run": bool(np.array_equal(hist, hist2)),
                                "what": f"GPU watcher histories of the timed steps and the final field (N={n}) against the "
                                        "scipy-LU oracle run over the same steps"}
        if not (err_hist <= 1e-10 and err_field <= 1e-10):
            raise RuntimeError(f"bench.py: GPU result differs from the oracle: {line['parity_check']}")
        if sweep is not None:
            cb = cpu_sweep_baseline(sweep["steps"])
            line["sweep"]["cpu_baseline"] = cb
            line["config"]["sweep_cpu_sims_per_s"] = cb["sims_per_s"]
        if cl is not None and world == 1:
            line["konopkova_1m"]["cpu_baseline"] = large_cpu_baseline(cl)
    else:
        line["parity_check"] = {"e2e_equals_device_run": bool(np.array_equal(hist, hist2))}
    json_out.write(json.dumps(line) + "\n")
    json_out.flush()
    if s is not None:
        s.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rtol", type=float, default=1e-14)
    ap.add_argument("--warm-start", type=float, default=1.0)
    ap.add_argument("--recycle", type=int, default=128, help="hf_set_recycle vectors (the runners' default)")
    ap.add_argument("--skip-large", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-sweep", action="store_true")
    ap.add_argument("--sweep-per-gpu", type=int, default=128, help="variants per GPU of the sweep leg (multiple of 16)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
