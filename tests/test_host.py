"""CPU tests of the host-side mirror of the reference interface (BCs, problem set-up, XDMF, runners)."""
import io
import contextlib
import os

import numpy as np
import pytest
import yaml

from heatflow_b200 import fem, problem
from heatflow_b200.dirichlet_bc import RowDirichletBC, resolve_last_wins
from heatflow_b200.io_utilities.xdmf_extract import extract_point_timeseries_xdmf
from heatflow_b200.io_utilities.xdmf_utils import XDMFFile, init_xdmf, save_params
from heatflow_b200.mesh_and_materials import Domain
from heatflow_b200.run_no_diamond_1d import extract_1d_submesh_from_2d
from heatflow_b200.mesh_and_materials.mesh import MeshTags
from helpers import build_case, load_cfg
from oracle import heat_oracle as ho


@pytest.fixture(scope="module")
def case():
    return build_case("geballe_no_diamond", 8.0)


def test_row_bc_matches_oracle_location(case):
    V = case.V
    for loc, kw in (("left", {}), ("right", {}), ("top", {}), ("bottom", {}),
                    ("x", dict(coord=case.heating_z, length=2 * case.r_sample, center=0.0)),
                    ("x", dict(coord=case.heating_z, length=1e-5)),           # default centre = domain middle
                    ("y", dict(coord=0.0, length=2.2e-5))):
        bc = RowDirichletBC(V, loc, value=1.0, **kw)
        want = ho.locate_row_dofs(case.nodes, loc, **{k: v for k, v in kw.items()})
        assert np.array_equal(bc.row_dofs, want), loc
        assert np.array_equal(bc.dof_coords[:, :2], case.nodes[want])
    outer = RowDirichletBC(V, "outer", value=0.0)
    union = np.unique(np.concatenate([ho.locate_row_dofs(case.nodes, s) for s in ("left", "right", "bottom", "top")]))
    assert np.array_equal(outer.row_dofs, union)
    with pytest.raises(ValueError):
        RowDirichletBC(V, "x")
    with pytest.raises(ValueError):
        RowDirichletBC(V, "diagonal")
    with pytest.raises(RuntimeError):
        RowDirichletBC(V, "x", coord=1.0)                                    # nothing there


def test_row_bc_update_and_constant(case):
    bc = RowDirichletBC(case.V, "x", coord=case.heating_z, length=2 * case.r_sample, center=0.0,
                        value=lambda x, y, t: 2.0 * t + y)
    bc.update(3.0)
    assert np.allclose(bc._g.x.array[bc.row_dofs], 6.0 + case.nodes[bc.row_dofs, 1])
    assert np.all(np.delete(bc._g.x.array, bc.row_dofs) == 0.0)
    const = RowDirichletBC.constant(case.V, "left", 7.5)
    assert np.all(const._g.x.array[const.row_dofs] == 7.5) and const.bc.dofs is const.row_dofs
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        RowDirichletBC.describe_row_bcs([bc, "not a bc", const])
    assert "Row BC #0" in out.getvalue() and "Row BC #2" in out.getvalue() and f"n = {bc.row_dofs.size} DOFs" in out.getvalue()


def test_device_bc_arrays_last_wins(case):
    owner = resolve_last_wins(len(case.nodes), case.bcs)
    dofs, value, gslot, gr = problem.device_bc_arrays(len(case.nodes), case.bcs, case.bcs[3], case.nodes)
    assert np.array_equal(dofs, np.flatnonzero(owner >= 0)) and np.all(np.diff(dofs) > 0)
    assert np.array_equal(dofs[gslot], np.flatnonzero(owner == 3))
    assert np.array_equal(gr, case.nodes[dofs[gslot], 1])
    assert np.all(value == case.ic)                                          # t = 0: everything at ic_temp
    O = ho.Oracle2D(case.nodes, case.tris, case.rhoc_c, case.kappa_c, case.dt, case.oracle_bcs, case.ic, case.fwhm,
                    case.heat_t, case.heat_T)
    assert np.array_equal(dofs, O.bc_dofs) and np.array_equal(np.sort(dofs[gslot]), np.sort(O.gauss_dofs))


def test_stacks_follow_the_reference_geometry():
    cfg = load_cfg("geballe_with_diamond")
    mats, bounds, info = problem.stack_with_diamond(cfg)
    assert [m.name for m in mats] == problem.WITH_DIAMOND_ORDER
    box = {m.name: m.boundaries for m in mats}
    assert bounds[3] == pytest.approx(80e-6) and bounds[0] == pytest.approx(-(0.92e-6 + 3.2e-6 + 6.2e-8 + 40e-6))
    assert box["p_sample"][0] == pytest.approx(-0.92e-6) and box["p_sample"][1] == pytest.approx(0.92e-6)
    assert box["p_coupler"][1] == pytest.approx(box["p_sample"][0]) and box["o_coupler"][0] == pytest.approx(box["p_sample"][1], abs=1e-20)
    assert box["g_ins"][2:] == pytest.approx([20e-6, 25e-6]) and box["gasket"][2:] == pytest.approx([25e-6, 80e-6])
    assert mats[3].properties == {"rho_cv": 5164.0 * 1158.0, "k": 3.8} and mats[3].mesh_size == 0.04e-6
    assert isinstance(cfg["mats"]["g_ins"]["r"], str)                        # PyYAML quirk the float() calls absorb
    nd, b2, _ = problem.stack_no_diamond(load_cfg("geballe_no_diamond"))
    assert [m.name for m in nd] == problem.NO_DIAMOND_ORDER and b2[3] == pytest.approx(40e-6)
    assert max(m.boundaries[3] for m in nd) == pytest.approx(20e-6)          # materials stop at r = 20 um


def test_heating_amplitudes_vectorised_matches_scalar(case):
    t = (np.arange(case.num_steps) + 1) * case.dt
    want = [ho.heating_amplitude(x, case.heat_t, case.heat_T, case.ic) for x in t]
    assert np.array_equal(case.amps, want)
    with pytest.raises(ValueError):
        problem.read_heating_curve(os.path.join(os.path.dirname(case.cfg["heating"]["file"]), "konopkova_pside.csv"))


def test_xdmf_roundtrip(tmp_path, case):
    dom = Domain(case.arrays)
    x = init_xdmf(dom, str(tmp_path), "output")

    class F:
        name = "Temperature (K)"
        x = fem._Vector(len(case.nodes))
    f = F()
    fields = []
    for k, t in enumerate([0.0, 1e-7, 2e-7]):
        f.x.array[:] = 300.0 + k + case.nodes[:, 1] * 1e6
        fields.append(f.x.array.copy())
        x.write_function(f, t)
    x.close()
    text = open(tmp_path / "output.xdmf").read()
    assert 'Version="3.0"' in text and 'CollectionType="Temporal"' in text and 'TopologyType="Triangle"' in text
    assert text.count("<Time Value=") == 3 and 'Name="Temperature (K)"' in text and "xi:include" in text
    q = [(case.nodes[5, 0], case.nodes[5, 1]), (case.nodes[100, 0], case.nodes[100, 1])]
    times, data = extract_point_timeseries_xdmf(str(tmp_path / "output.xdmf"), "Temperature (K)", q)
    assert np.array_equal(times, [0.0, 1e-7, 2e-7])
    assert np.array_equal(data, np.array([[fl[5] for fl in fields], [fl[100] for fl in fields]]))
    save_params(str(tmp_path), {"a": 1, "b": "x"})
    assert open(tmp_path / "params.txt").read() == "a = 1\nb = x\n"
    with pytest.raises(RuntimeError):
        XDMFFile(None, str(tmp_path / "o2.xdmf"), "w").write_function(f, 0.0)


def test_axis_submesh_matches_oracle(case):
    dom = Domain(case.arrays)
    with contextlib.redirect_stdout(io.StringIO()):
        d1, tags1, maps = extract_1d_submesh_from_2d(dom, MeshTags(case.cell_tag))
    z, cells, tags, verts = ho.extract_axis_submesh(case.nodes, case.tris, case.cell_tag)
    assert np.array_equal(d1.geometry.x[:, 0], z) and np.array_equal(d1.cells, cells)
    assert np.array_equal(tags1.values, tags) and np.array_equal(maps[1], verts)
    far = Domain(case.arrays)
    far.geometry.x = far.geometry.x + np.array([0.0, 1.0, 0.0])
    with pytest.raises(ValueError):
        extract_1d_submesh_from_2d(far, MeshTags(case.cell_tag))


def test_runner_file_contract(tmp_path):
    import run_no_diamond
    import run_with_diamond
    cfg = load_cfg("geballe_no_diamond")
    with pytest.raises(FileNotFoundError, match="mesh.msh, mesh_cfg.yaml"):
        run_no_diamond.run_simulation(cfg, str(tmp_path / "nomesh"), rebuild_mesh=False, suppress_print=True)
    # rebuild_mesh writes both files before the GPU is needed
    folder = tmp_path / "mesh"
    cfg["mats"]["p_ins"]["mesh"] = cfg["mats"]["o_ins"]["mesh"] = "0.4e-6"
    cfg["mats"]["p_sample"]["mesh"] = cfg["mats"]["p_coupler"]["mesh"] = cfg["mats"]["o_coupler"]["mesh"] = "0.3e-6"
    from heatflow_b200.runners import prepare_mesh
    with contextlib.redirect_stdout(io.StringIO()):
        mats, info, dom, tags, tagmap = prepare_mesh(cfg, problem.stack_no_diamond, str(folder), True)
    assert sorted(os.listdir(folder)) == ["mesh.msh", "mesh_cfg.yaml"]
    mesh_cfg = yaml.safe_load(open(folder / "mesh_cfg.yaml"))
    assert mesh_cfg["material_tags"] == {"p_ins": 1, "p_coupler": 2, "p_sample": 3, "o_coupler": 4, "o_ins": 5}
    assert mesh_cfg["mats"] == cfg["mats"] and cfg["material_tags"] == {}
    os.remove(folder / "mesh_cfg.yaml")
    with pytest.raises(FileNotFoundError, match="mesh_cfg.yaml"):
        prepare_mesh(cfg, problem.stack_no_diamond, str(folder), False)
    assert run_with_diamond.run_simulation.__code__.co_varnames[:8] == (
        "cfg", "mesh_folder", "rebuild_mesh", "visualize_mesh", "output_folder", "watcher_points", "write_xdmf", "suppress_print")


def test_bench_reference_arm_prints_one_contract_line():
    # the CPU arm of bench.py runs without a GPU: one JSON line with the keys the round contract names
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("DOF-timesteps/sec") and d["unit"] == "DOF-timesteps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_reference_arm_under_torchrun_world_size_2():
    # N > 1: launched like the driver does; rank 0 alone prints the line (N simulations on N cores), rank 1 exits 0
    import json
    import socket
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
           "--warmup", "1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["cpu_baseline"]["cores"] == 2 and d["scaling"] == "weak"
    assert "2 independent cop" in d["config"]["workload"] and d["value"] > 0 and d["warmup"] == 3


def test_cfgs_of_baseline_configs_1_and_2():
    # configs[0] (geballe_1d.yaml) is the no-diamond stack with 50 steps; geballe_no_diamond_read_flux.yaml is the same
    # file - the GPU tests run geballe_1d.yaml through run_no_diamond + run_1d (tests/test_gpu_runners.py)
    import yaml
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    load = lambda n: yaml.safe_load(open(os.path.join(root, "cfgs", n + ".yaml")))
    one_d, flux, nd = load("geballe_1d"), load("geballe_no_diamond_read_flux"), load("geballe_no_diamond")
    assert one_d == flux
    assert one_d["timing"]["num_steps"] == 50 and nd["timing"]["num_steps"] == 40
    nd["timing"]["num_steps"] = 50
    assert one_d == nd
