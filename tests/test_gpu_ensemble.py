"""GPU parity tests of the batched multi-RHS ensemble (parameter_sweep path) against the scipy oracle."""
import json
import os

import numpy as np
import pandas as pd
import pytest
import yaml

from heatflow_b200 import _lib, problem
from helpers import build_case, load_cfg, make_solver
from oracle import heat_oracle as ho

pytestmark = pytest.mark.gpu
RTOL_FIELD = 1e-10   # north_star tolerance for temperature histories


def oracle_variant(c, k_sample, fwhm):
    kap_t = c.kappa_t.copy()
    kap_t[[m.name for m in c.mats].index("p_sample")] = k_sample
    return ho.Oracle2D(c.nodes, c.tris, c.rhoc_c, kap_t[c.cell_tag - 1], c.dt, c.oracle_bcs, c.ic, fwhm, c.heat_t, c.heat_T)


def sample_tag(c):
    return int(c.tags[[m.name for m in c.mats].index("p_sample")])


@pytest.fixture(scope="module")
def wd():
    return build_case("geballe_with_diamond", 8.0)


@pytest.mark.parametrize("ks,fw", [
    ([3.8], [1.3e-5]),                                                    # B = 1
    ([1.0, 10.0, 100.0], [1e-6, 1.3e-5, 1e-4]),                           # 3 variants padded to B = 4
    (list(np.logspace(0, 2, 8)), list(np.logspace(-6, -4, 8)[::-1])),     # B = 8
])
def test_ensemble_matches_oracle_per_variant(wd, ks, fw):
    c = wd
    S = 25
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 3e-8, 0.0), (0.95e-6, 0.0), (0.0, 5e-6)])
    s = make_solver(c, ordering="hilbert" if len(ks) != 3 else "given")
    s.ens_create(ks, [problem.gaussian_coeff(f) for f in fw], sample_tag(c))
    hist, iters = s.ens_run(c.amps[:S], c.ic, watch)
    u = s.ens_get_state()
    assert hist.shape == (len(ks), S, 3) and u.shape == (len(ks), len(c.nodes)) and iters.shape == (S,)
    assert np.all(iters[5:] > 0)
    for b, (k, f) in enumerate(zip(ks, fw)):
        O = oracle_variant(c, k, f)
        ohist, _ = O.run(S, watch)
        assert np.abs(hist[b] / ohist - 1).max() <= RTOL_FIELD, (b, k, f)
        assert np.abs(u[b] / O.u - 1).max() <= RTOL_FIELD, (b, k, f)
    s.ens_destroy()
    s.close()


def test_ensemble_equals_single_simulation_path(wd):
    c = wd
    S = 30
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 3e-8, 0.0), (0.95e-6, 0.0)])
    k0 = float(c.kappa_t[[m.name for m in c.mats].index("p_sample")])
    s = make_solver(c)
    h1, _, _ = s.run(c.amps[:S], c.ic, c.coeff, watch)
    u1 = s.get_state()
    s.set_state(np.full(len(c.nodes), c.ic))
    s.ens_create([k0, k0], [c.coeff, c.coeff], sample_tag(c))
    h2, _ = s.ens_run(c.amps[:S], c.ic, watch)
    u2 = s.ens_get_state()
    assert np.array_equal(h2[0], h2[1]) and np.array_equal(u2[0], u2[1])     # identical variants, identical bits
    assert np.abs(h2[0] / h1 - 1).max() <= 1e-11
    assert np.abs(u2[0] / u1 - 1).max() <= 1e-11
    # a second run continues from the ensemble state (steps S .. 2S-1 of the curve)
    h3, _ = s.ens_run(c.amps[S:2 * S], c.ic, watch)
    O = oracle_variant(c, k0, c.fwhm)
    oh, _ = O.run(2 * S, watch)
    assert np.abs(h3[0] / oh[S:] - 1).max() <= RTOL_FIELD
    s.close()


def test_ensemble_argument_errors(wd):
    c = wd
    s = make_solver(c)
    with pytest.raises(_lib.HeatflowError):
        s.ens_run(c.amps[:2], c.ic, [0])                    # no ensemble yet
    for ks in ([], [1.0] * 33, [0.0], [-1.0]):
        with pytest.raises(_lib.HeatflowError):
            s.ens_create(ks, [c.coeff] * len(ks), sample_tag(c))
    with pytest.raises(_lib.HeatflowError):
        s.ens_create([1.0], [c.coeff], 999)                 # not a material tag
    with pytest.raises(ValueError):
        s.ens_create([1.0, 2.0], [c.coeff], sample_tag(c))
    s.ens_create([1.0], [c.coeff], sample_tag(c))
    with pytest.raises(_lib.HeatflowError):
        s.ens_run(c.amps[:2], c.ic, [len(c.nodes)])         # watch node out of range
    s.close()


def coarse_cfg_file(tmp_path, name, factor):
    cfg = load_cfg(name)
    for m in cfg["mats"].values():
        m["mesh"] = float(m["mesh"]) * factor
    path = tmp_path / "base.yaml"
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    return cfg, str(path)


@pytest.mark.parametrize("name,factor,mode", [("geballe_with_diamond", 8.0, "ensemble"), ("geballe_no_diamond", 6.0, "ensemble"),
                                              ("geballe_with_diamond", 8.0, "auto"), ("geballe_no_diamond", 6.0, "serial")])
def test_parameter_sweep_outputs_match_oracle(tmp_path, name, factor, mode):
    import parameter_sweep as psw
    from heatflow_b200.mesh_and_materials import read_msh
    cfg, cfg_path = coarse_cfg_file(tmp_path, name, factor)
    out, meshes = str(tmp_path / "sweep"), str(tmp_path / "meshes")
    width = float(cfg["mats"]["p_sample"]["z"])
    results, failed = psw.run_parameter_sweep(cfg_path, out, (2e-6, 5e-5), (2.0, 50.0), (width, width), (2, 3, 1),
                                              base_mesh_folder=meshes, batch=4, mode=mode)
    assert failed == [] and len(results) == 6
    meta = json.load(open(os.path.join(out, "sweep_metadata.json")))
    assert meta["execution"]["mode"] == mode
    assert meta["total_runs"] == 6 and meta["k_values"] == np.logspace(np.log10(2.0), np.log10(50.0), 3).tolist()
    ok = pd.read_csv(os.path.join(out, "successful_runs.csv"))
    assert list(ok.columns) == ["run_id", "run_name", "fwhm", "k", "width", "output_dir", "runtime", "status", "error"]
    assert not os.path.exists(os.path.join(out, "failed_runs.csv"))
    mesh_folder = psw.get_mesh_folder_for_width(meshes, width)
    assert sorted(os.listdir(mesh_folder)) == ["mesh.msh", "mesh_cfg.yaml"]
    nodes, tris, tag, _ = read_msh(os.path.join(mesh_folder, "mesh.msh"))
    with_diamond = "p_diam" in cfg["mats"]
    S = int(cfg["timing"]["num_steps"])
    dt = float(cfg["timing"]["t_final"]) / S
    ht, hT = ho.load_heating(cfg["heating"]["file"])
    for r in results:
        run_dir = os.path.join(out, r["run_name"])
        assert r["output_dir"] == run_dir and sorted(os.listdir(run_dir)) == ["used_config.yaml", "watcher_points.csv"]
        used = yaml.safe_load(open(os.path.join(run_dir, "used_config.yaml")))
        assert used == psw.modify_config_for_parameters(cfg, r["fwhm"], r["k"], r["width"])
        mats, _, info = (problem.stack_with_diamond if with_diamond else problem.stack_no_diamond)(used)
        kap = np.array([m.properties["k"] for m in mats])[tag - 1]
        rc = np.array([m.properties["rho_cv"] for m in mats])[tag - 1]
        zc = next(m for m in mats if m.name == "p_coupler").boundaries[0]
        bcs = [(ho.locate_row_dofs(nodes, "left"), "const"), (ho.locate_row_dofs(nodes, "right"), "const"),
               (ho.locate_row_dofs(nodes, "top"), "const"),
               (ho.locate_row_dofs(nodes, "x", coord=zc, length=2 * info["r_sample"], center=0.0), "gauss")]
        O = ho.Oracle2D(nodes, tris, rc, kap, dt, bcs, float(used["heating"]["ic_temp"]), float(used["heating"]["fwhm"]), ht, hT)
        watch = ho.nearest_nodes(nodes, list(psw.get_watcher_points(used).values()))
        ohist, _ = O.run(S, watch)
        df = pd.read_csv(os.path.join(run_dir, "watcher_points.csv"))
        assert list(df.columns) == ["time", "pside", "oside"] and len(df) == S
        assert np.allclose(df["time"], (np.arange(S) + 1) * dt, rtol=1e-15)
        assert np.abs(df[["pside", "oside"]].to_numpy() / ohist - 1).max() <= RTOL_FIELD, r["run_name"]
    # second call reuses the mesh; per_run mode produces the same watcher histories through run_simulation
    out2 = str(tmp_path / "sweep2")
    res2, failed2 = psw.run_parameter_sweep(cfg_path, out2, (2e-6, 5e-5), (2.0, 50.0), (width, width), (1, 2, 1),
                                            base_mesh_folder=meshes, mode="per_run")
    assert failed2 == [] and len(res2) == 2
    for r in res2:
        a = pd.read_csv(os.path.join(out2, r["run_name"], "watcher_points.csv"))
        b = pd.read_csv(os.path.join(out, r["run_name"], "watcher_points.csv"))
        assert np.abs(a[["pside", "oside"]].to_numpy() / b[["pside", "oside"]].to_numpy() - 1).max() <= 1e-11


def test_config5_corner_variants_at_full_size_match_oracle(tmp_path):
    # BASELINE config #5 spot check: the four corner variants (k = 1 / 100 W/m/K, fwhm = 1e-6 / 1e-4 m) of the 64 x 64
    # sweep, on the cfg's own mesh (1.4e5 dofs) and all 100 steps, through run_parameter_sweep with its defaults -
    # every watcher history against the LU oracle
    import parameter_sweep as psw
    from heatflow_b200.mesh_and_materials import read_msh
    cfg, cfg_path = coarse_cfg_file(tmp_path, "geballe_with_diamond", 1.0)      # cfg sizes; absolute heating-file path
    out, meshes = str(tmp_path / "sweep"), str(tmp_path / "meshes")
    width = float(cfg["mats"]["p_sample"]["z"])
    results, failed = psw.run_parameter_sweep(cfg_path, out, (1e-6, 1e-4), (1.0, 100.0), (width, width), (2, 2, 1),
                                              base_mesh_folder=meshes)
    assert failed == [] and len(results) == 4
    nodes, tris, tag, _ = read_msh(os.path.join(psw.get_mesh_folder_for_width(meshes, width), "mesh.msh"))
    assert len(nodes) > 130_000
    S = int(cfg["timing"]["num_steps"])
    dt = float(cfg["timing"]["t_final"]) / S
    for r in results:
        used = yaml.safe_load(open(os.path.join(out, r["run_name"], "used_config.yaml")))
        mats, _, info = problem.stack_with_diamond(used)
        kap = np.array([m.properties["k"] for m in mats])[tag - 1]
        rc = np.array([m.properties["rho_cv"] for m in mats])[tag - 1]
        zc = next(m for m in mats if m.name == "p_coupler").boundaries[0]
        bcs = [(ho.locate_row_dofs(nodes, "left"), "const"), (ho.locate_row_dofs(nodes, "right"), "const"),
               (ho.locate_row_dofs(nodes, "top"), "const"),
               (ho.locate_row_dofs(nodes, "x", coord=zc, length=2 * info["r_sample"], center=0.0), "gauss")]
        ht, hT = ho.load_heating(cfg["heating"]["file"])
        O = ho.Oracle2D(nodes, tris, rc, kap, dt, bcs, float(used["heating"]["ic_temp"]), float(used["heating"]["fwhm"]), ht, hT)
        watch = ho.nearest_nodes(nodes, list(psw.get_watcher_points(used).values()))
        ohist, _ = O.run(S, watch)
        df = pd.read_csv(os.path.join(out, r["run_name"], "watcher_points.csv"))
        assert np.abs(df[["pside", "oside"]].to_numpy() / ohist - 1).max() <= RTOL_FIELD, r["run_name"]


@pytest.mark.parametrize("ks,fw,cap", [
    ([1.0, 10.0, 100.0], [1e-6, 1.3e-5, 1e-4], 64),                        # padded tile, full history
    (list(np.logspace(0, 2, 16)), list(np.logspace(-6, -4, 16)[::-1]), 64),  # B = 16
    (list(np.logspace(0, 2, 8)), list(np.logspace(-6, -4, 8)), 5),         # basis frozen after 5 solves
])
def test_ensemble_with_recycled_initial_guess_matches_oracle(wd, ks, fw, cap):
    # per-variant recycled bases (hf_set_recycle applies to the ensemble too): same answers, fewer iterations
    c = wd
    S = 40
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 3e-8, 0.0), (0.95e-6, 0.0), (0.0, 5e-6)])
    coeffs = [problem.gaussian_coeff(f) for f in fw]
    s0 = make_solver(c, warm=1.0, ordering="hilbert")
    s0.ens_create(ks, coeffs, sample_tag(c))
    _, it0 = s0.ens_run(c.amps[:S], c.ic, watch)
    s0.close()
    s = make_solver(c, warm=1.0, ordering="hilbert", recycle=cap)
    s.ens_create(ks, coeffs, sample_tag(c))
    hist, iters = s.ens_run(c.amps[:S], c.ic, watch)
    u = s.ens_get_state()
    for b, (k, f) in enumerate(zip(ks, fw)):
        O = oracle_variant(c, k, f)
        ohist, _ = O.run(S, watch)
        assert np.abs(hist[b] / ohist - 1).max() <= RTOL_FIELD, (b, k, f)
        assert np.abs(u[b] / O.u - 1).max() <= RTOL_FIELD, (b, k, f)
    if cap >= S:
        assert iters.sum() < 0.6 * it0.sum()
    s.ens_destroy()
    s.close()


def test_two_simulations_sharing_the_gpu_match_sequential_runs(wd):
    # hf_set_sharing(2): two contexts driven from two host threads run their on-chip kernels concurrently;
    # every simulation must give the answer it gives alone (1e-10 vs the oracle, ~1e-12 between kernels)
    import threading
    c = wd
    S = 40
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 3e-8, 0.0), (0.95e-6, 0.0)])
    fws = [2e-6, 7e-6, 1.3e-5, 4e-5, 9e-5, 3e-6]
    ref = make_solver(c, warm=1.0, recycle=64)
    assert ref.on_chip()
    alone = []
    for f in fws:
        ref.set_state(np.full(len(c.nodes), c.ic))
        h, _, _ = ref.run(c.amps[:S], c.ic, problem.gaussian_coeff(f), watch)
        alone.append(h)
    ref.close()
    from heatflow_b200.solver import HeatSolver
    pair = []
    for _ in range(2):
        s = HeatSolver(0)
        s.set_sharing(2)
        s.set_mesh(c.nodes, c.tris, c.cell_tag)
        s.set_materials(c.tags, c.kappa_t, c.rhoc_t)
        s.set_bcs(c.bc_dofs, c.bc_value, c.gauss_slot, c.gauss_r)
        s.build_operator(c.dt, True)
        s.set_solver(rtol=1e-14, warm=1.0)
        s.set_recycle(64)
        assert s.on_chip()
        pair.append(s)
    got = {}

    def work(s, mine):
        for i in mine:
            s.set_state(np.full(len(c.nodes), c.ic))
            got[i], _, _ = s.run(c.amps[:S], c.ic, problem.gaussian_coeff(fws[i]), watch)

    th = [threading.Thread(target=work, args=(pair[j], range(j, len(fws), 2))) for j in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    for i, f in enumerate(fws):
        assert np.abs(got[i] / alone[i] - 1).max() <= 1e-11, i
    O = oracle_variant(c, float(c.kappa_t[[m.name for m in c.mats].index("p_sample")]), fws[2])
    oh, _ = O.run(S, watch)
    assert np.abs(got[2] / oh - 1).max() <= RTOL_FIELD
    with pytest.raises(_lib.HeatflowError):
        pair[0].set_sharing(3)
    for s in pair:
        s.close()
