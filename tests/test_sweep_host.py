"""CPU tests of the parameter_sweep host logic: grid, naming, tiling / sharding plan, final gather (gloo)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from heatflow_b200 import sweep
from helpers import load_cfg

HERE = os.path.dirname(os.path.abspath(__file__))


def ps():
    import parameter_sweep          # root-level shim, as a user of the reference imports it
    return parameter_sweep


def test_shim_exports_reference_api():
    m = ps()
    for name in ("set_single_thread", "initialize_worker", "get_watcher_points", "run_single_simulation",
                 "create_parameter_grid", "modify_config_for_parameters", "get_mesh_folder_for_width",
                 "run_parameter_sweep", "main"):
        assert callable(getattr(m, name)), name


def test_parameter_grid_matches_reference_generators():
    combos, f, k, w = ps().create_parameter_grid((1e-6, 1e-4), (1.0, 100.0), (1e-6, 10e-6), (5, 5, 3))
    assert np.array_equal(f, np.logspace(-6, -4, 5)) and np.array_equal(k, np.logspace(0, 2, 5))
    assert np.array_equal(w, np.linspace(1e-6, 10e-6, 3))
    assert len(combos) == 75
    # grouped by width first, then fwhm-major / k-minor (itertools.product order)
    assert [c["width"] for c in combos[:25]] == [w[0]] * 25
    assert [c["k"] for c in combos[:5]] == list(k) and all(c["fwhm"] == f[0] for c in combos[:5])
    assert combos[5]["fwhm"] == f[1]


def test_run_and_mesh_folder_names():
    from heatflow_b200.parameter_sweep import run_name_for
    assert run_name_for(1e-6, 1.0, 1.84e-6) == "fwhm_1.00e-6_k_1.00_width_1.84e-6"
    assert run_name_for(1.3e-5, 31.62, 1e-5) == "fwhm_1.30e-5_k_31.62_width_1.00e-5"
    assert run_name_for(2.5e+10, 100.0, 1.0e-12) == "fwhm_2.50e10_k_100.00_width_1.00e-12"
    assert ps().get_mesh_folder_for_width("meshes", 1.84e-6) == os.path.join("meshes", "width_1.840e-6")


def test_modify_config_and_watchers():
    m = ps()
    base = load_cfg("geballe_with_diamond")
    cfg = m.modify_config_for_parameters(base, np.float64(2e-5), np.float64(7.5), np.float64(2.5e-6))
    assert type(cfg["heating"]["fwhm"]) is float and cfg["heating"]["fwhm"] == 2e-5
    assert cfg["mats"]["p_sample"]["k"] == 7.5 and cfg["mats"]["p_sample"]["z"] == 2.5e-6
    assert float(base["mats"]["p_sample"]["z"]) != 2.5e-6          # the base cfg is left alone
    wp = m.get_watcher_points(cfg)
    zs, zc = 2.5e-6, float(cfg["mats"]["p_coupler"]["z"])
    zpi, zoi, zd = (float(cfg["mats"][n]["z"]) for n in ("p_ins", "o_ins", "p_diam"))
    zmin = -(zs / 2) - zpi - zc - zd
    zmax = (zs / 2) + zoi + zc + zd
    assert list(wp) == ["pside", "oside"]
    assert wp["pside"] == (zmin + zd + zpi + zc / 2, 0.0) and wp["oside"] == (zmax - zd - zoi - zc / 2, 0.0)
    nd = m.get_watcher_points(load_cfg("geballe_no_diamond"))
    c = load_cfg("geballe_no_diamond")["mats"]
    zmin = -(float(c["p_sample"]["z"]) / 2) - float(c["p_ins"]["z"]) - float(c["p_coupler"]["z"])
    assert nd["pside"] == (zmin + float(c["p_ins"]["z"]) + float(c["p_coupler"]["z"]) / 2, 0.0)


@pytest.mark.parametrize("P,B,R", [(1, 16, 1), (37, 4, 2), (4096, 16, 8), (5, 32, 8), (64, 1, 3)])
def test_plan_tiles_partition(P, B, R):
    rng = np.random.default_rng(P)
    k = rng.uniform(1.0, 100.0, P)
    tiles = sweep.plan_tiles(k, B, R)
    assert len(tiles) == R
    flat = [t for r in tiles for t in r]
    assert sorted(np.concatenate(flat).tolist()) == list(range(P))           # each variant exactly once
    assert all(1 <= len(t) <= B for t in flat)
    for t in flat:                                                           # a tile is a contiguous k-range
        assert np.all(np.diff(k[t]) >= 0)
    counts = [sum(len(t) for t in r) for r in tiles]
    assert max(counts) - min(counts) <= B                                    # balanced to one tile
    n_tiles = [len(r) for r in tiles]
    assert max(n_tiles) - min(n_tiles) <= 1


def test_plan_tiles_rejects_bad_arguments():
    with pytest.raises(ValueError):
        sweep.plan_tiles([1.0], 0, 1)
    with pytest.raises(ValueError):
        sweep.plan_tiles([1.0], 4, 0)
    assert sweep.plan_tiles([], 4, 2) == [[], []]


def test_gather_results_single_rank():
    idx = np.array([2, 0])
    hist = np.arange(2 * 3 * 2, dtype=float).reshape(2, 3, 2)
    H, it, sc, err = sweep.gather_results(3, 3, 2, idx, hist, np.array([7, 9]), np.array([0.1, 0.2]), {0: "x"})
    assert np.array_equal(H[2], hist[0]) and np.array_equal(H[0], hist[1]) and np.all(np.isnan(H[1]))
    assert it.tolist() == [9, -1, 7] and err == {0: "x"}


def test_final_gather_gloo_world_size_2(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ok = tmp_path / "ok.txt"
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(HERE, "_gloo_sweep_worker.py"), str(ok)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert ok.read_text() == "ok"


def test_parameter_sweep_driver_gloo_world_size_2(tmp_path):
    # the sweep driver under torchrun, 2 ranks, gloo: per-rank output writing, one final gather, a failing rank
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(HERE, "_gloo_psweep_worker.py"), str(tmp_path)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert (tmp_path / "ok.txt").read_text() == "ok"


class _FakeSolver:
    """Stands in for HeatSolver in the host-side tests of the serial sweep engine (no GPU)."""

    def __init__(self, log):
        self.log, self.k = log, None

    def set_state(self, u0):
        self.u0 = float(u0[0])

    def run(self, amps, t_ic, coeff, watch_nodes):
        import time as _t
        _t.sleep(0.002)
        if coeff < -1e17:
            raise RuntimeError("diverged")
        S, W = len(amps), len(watch_nodes)
        return np.full((S, W), self.k * 1000.0 + coeff), np.full(S, 7, dtype=np.int32), None


class _FakeSim:
    def __init__(self, log):
        self.solver = _FakeSolver(log)
        self.num_steps, self.n_dofs, self.ic_temp = 5, 11, 300.0
        self.amps = np.arange(5.0)
        self.log = log

    def set_conductivity(self, name, k):
        self.log.append((id(self), k))
        self.solver.k = k


@pytest.mark.parametrize("workers", [1, 2])
def test_serial_engine_queue_results_and_errors(workers):
    # variants are pulled from one queue by `workers` threads; results come back in tile order whatever thread
    # ran them, a failing variant is recorded and does not stop the others, k changes re-assemble per worker
    from heatflow_b200 import problem
    log = []
    sims = [_FakeSim(log) for _ in range(workers)]
    k = np.array([1.0, 1.0, 2.0, 2.0, 2.0, 3.0, 3.0])
    fwhm = np.array([1e-6, 2e-6, 3e-6, 4e-6, 5e-6, 6e-6, 1e-9])       # the last one "diverges" (huge |coeff|)
    tiles = sweep.plan_tiles(k, 3, 1)[0]
    idx, hist, iters, secs, errors = sweep.run_tiles_serial(sims, fwhm, k, tiles, [4, 9])
    assert idx.tolist() == [int(i) for t in tiles for i in t] and hist.shape == (7, 5, 2)
    for pos, i in enumerate(idx):
        if i == 6:
            assert i in errors and "diverged" in errors[i] and np.isnan(hist[pos]).all() and iters[pos] == -1
        else:
            assert np.allclose(hist[pos], k[i] * 1000.0 + problem.gaussian_coeff(fwhm[i])) and iters[pos] == 35
    assert set(errors) == {6} and np.all(secs > 0)
    # every worker re-assembles only when the conductivity it sees changes
    per_worker = {}
    for wid, kv in log:
        per_worker.setdefault(wid, []).append(kv)
    assert all(len(v) <= 3 and all(a != b for a, b in zip(v, v[1:])) for v in per_worker.values())
    # empty share
    out = sweep.run_tiles_serial(sims, fwhm, k, [], [4, 9])
    assert out[0].size == 0 and out[1].shape == (0, 5, 2)


def test_serial_engine_share_is_a_partition_with_every_conductivity_on_every_rank():
    # _group_on_device(share=(rank, world)): the serial engine deals the k-sorted variants one by one, so every rank
    # sees every conductivity with a mix of narrow and wide heating profiles (tiles of 16 put all wide ones on odd ranks)
    fw, ks = np.logspace(-6, -4, 32), np.logspace(0, 2, 32)
    k = np.array([kk for f in fw for kk in ks])
    f = np.array([ff for ff in fw for _ in ks])
    from heatflow_b200.parameter_sweep import serial_share
    shares = [serial_share(k, r, 4) for r in range(4)]
    assert sorted(np.concatenate(shares).tolist()) == list(range(len(k)))
    means = [np.log(f[s]).mean() for s in shares]
    assert max(means) - min(means) < 0.2                       # tiles of 16: the difference is ln(10) = 2.3
    for s in shares:
        assert len(np.unique(k[s])) == 32


def test_mesh_hand_over_marker_and_wait(tmp_path):
    # ranks > 0 wait for rank 0's mesh through the file system: marker next to the folder, both files required
    import threading
    import time
    from heatflow_b200 import parameter_sweep as psw
    folder = tmp_path / "meshes" / "width_1"
    folder.mkdir(parents=True)
    mesh, cfg = str(folder / "mesh.msh"), str(folder / "mesh_cfg.yaml")
    assert psw._mesh_marker(str(folder)) == str(folder) + ".ready"
    with pytest.raises(TimeoutError):
        psw._wait_for_mesh(str(folder), mesh, cfg, timeout_s=0.1)
    def rank0():
        time.sleep(0.1)
        open(cfg, "w").write("a: 1\n")
        open(mesh, "w").write("$MeshFormat\n")
        open(psw._mesh_marker(str(folder)), "w").write("ok\n")
    t = threading.Thread(target=rank0)
    t.start()
    t0 = time.time()
    psw._wait_for_mesh(str(folder), mesh, cfg, timeout_s=10.0)
    t.join()
    assert 0.05 < time.time() - t0 < 5.0 and sorted(os.listdir(folder)) == ["mesh.msh", "mesh_cfg.yaml"]
