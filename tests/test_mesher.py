"""CPU tests of the in-repo mesher and the MSH 4.1 reader/writer (replacing gmsh, mesh.py:81-195)."""
import io
import contextlib

import numpy as np
import pytest

from heatflow_b200 import problem
from heatflow_b200.mesh_and_materials import Material, Mesh, read_msh, triangulate_rectangles, write_msh
from helpers import load_cfg


def _edges(tris):
    e = np.concatenate([tris[:, [0, 1]], tris[:, [1, 2]], tris[:, [2, 0]]])
    return np.sort(e, axis=1)


def _signed_area(nodes, tris):
    a, b, c = nodes[tris[:, 0]], nodes[tris[:, 1]], nodes[tris[:, 2]]
    return 0.5 * ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]))


@pytest.fixture(scope="module", params=[("geballe_with_diamond", 4.0, "quadtree"), ("geballe_no_diamond", 4.0, "quadtree"),
                                        ("geballe_with_diamond", 4.0, "rows"), ("geballe_no_diamond", 4.0, "rows")],
                ids=lambda p: f"{p[0]}-{p[2]}")
def meshed(request):
    name, scale, method = request.param
    cfg = load_cfg(name)
    stack = problem.stack_with_diamond if "p_diam" in cfg["mats"] else problem.stack_no_diamond
    mats, bounds, _ = stack(cfg)
    m = triangulate_rectangles([x.boundaries for x in mats], [x.mesh_size for x in mats], size_scale=scale, method=method)
    m.method = method
    return mats, m, scale


def _angles_and_aspect(nodes, tris):
    """(smallest angle, largest angle) in degrees and longest edge / height on it, per triangle."""
    p = nodes[tris]
    a = np.linalg.norm(p[:, 1] - p[:, 2], axis=1)
    b = np.linalg.norm(p[:, 0] - p[:, 2], axis=1)
    c = np.linalg.norm(p[:, 0] - p[:, 1], axis=1)
    A = np.arccos(np.clip((b * b + c * c - a * a) / (2 * b * c), -1, 1))
    B = np.arccos(np.clip((a * a + c * c - b * b) / (2 * a * c), -1, 1))
    ang = np.degrees(np.stack([A, B, np.pi - A - B], axis=1))
    longest = np.maximum.reduce([a, b, c])
    return ang.min(axis=1), ang.max(axis=1), longest / (2.0 * np.abs(_signed_area(nodes, tris)) / longest)


@pytest.mark.parametrize("name", ["geballe_no_diamond", "geballe_with_diamond", "konopkova"])
def test_mesh_quality_bound_at_cfg_sizes(name):
    # the graded transition from 0.02 um (coupler) to 10 um (diamond, gasket) cells at the cfgs' own
    # sizes keeps every triangle well shaped - the reference gets this from gmsh's Delaunay / frontal mesher
    # (mesh_and_materials/mesh.py:129-147)
    cfg = load_cfg(name)
    stack = problem.stack_with_diamond if "p_diam" in cfg["mats"] else problem.stack_no_diamond
    mats, _, _ = stack(cfg)
    m = triangulate_rectangles([x.boundaries for x in mats], [x.mesh_size for x in mats])
    amin, amax, aspect = _angles_and_aspect(m.nodes, m.tris)
    assert amin.min() >= 12.0, amin.min()
    assert amax.max() <= 135.0, amax.max()
    assert aspect.max() <= 6.0, aspect.max()
    assert np.mean(amin >= 30.0) >= 0.97                     # the bulk is (near) right isosceles
    assert 1.0e5 < m.num_nodes < 1.6e5
    # round-1 row mesher on the same input, for the record: slivers wherever fine rows run out to coarse radii
    if name != "geballe_no_diamond":
        rows = triangulate_rectangles([x.boundaries for x in mats], [x.mesh_size for x in mats], method="rows")
        assert _angles_and_aspect(rows.nodes, rows.tris)[0].min() < 1.0


def test_conforming_positive_and_covering(meshed):
    mats, m, _ = meshed
    area = _signed_area(m.nodes, m.tris)
    assert np.all(area > 0)                                           # CCW, non-degenerate
    want = sum((b[1] - b[0]) * (b[3] - b[2]) for b in (x.boundaries for x in mats))
    assert abs(area.sum() - want) <= 1e-12 * want
    # conforming: every edge belongs to exactly 1 (boundary) or 2 (interior) triangles
    _, counts = np.unique(_edges(m.tris), axis=0, return_counts=True)
    assert set(np.unique(counts)) <= {1, 2}
    assert np.unique(m.tris).size == m.num_nodes                      # no orphan nodes
    assert np.unique(m.nodes, axis=0).shape[0] == m.num_nodes          # no duplicate nodes


def test_tags_and_sizes_follow_materials(meshed):
    mats, m, scale = meshed
    cent = m.nodes[m.tris].mean(axis=1)
    for i, mat in enumerate(mats):
        sel = m.cell_tag == i + 1
        assert sel.any()
        z0, z1, r0, r1 = mat.boundaries
        c = cent[sel]
        assert np.all((c[:, 0] > z0) & (c[:, 0] < z1) & (c[:, 1] > r0) & (c[:, 1] < r1))
        # per-area check: cell area never exceeds that of the target-size right triangle by much
        a = _signed_area(m.nodes, m.tris[sel])
        assert a.max() <= 0.5 * (1.35 * mat.mesh_size * scale) ** 2
    # material interfaces are mesh edges: no triangle has vertices strictly on both sides of one
    p_ins = mats[[x.name for x in mats].index("p_ins")]
    zi = p_ins.boundaries[1]
    tz = m.nodes[m.tris][:, :, 0]
    over = cent[:, 1] < p_ins.boundaries[3]                            # the radii over which that interface exists
    assert not np.any(over & (tz.min(axis=1) < zi - 1e-15) & (tz.max(axis=1) > zi + 1e-15))


def test_deterministic_and_row_major_band(meshed):
    mats, m, scale = meshed
    m2 = triangulate_rectangles([x.boundaries for x in mats], [x.mesh_size for x in mats], size_scale=scale, method=m.method)
    assert np.array_equal(m.nodes, m2.nodes) and np.array_equal(m.tris, m2.tris)
    if m.method == "rows":
        band = np.abs(m.tris[:, [0, 1, 2]] - m.tris[:, [1, 2, 0]]).max()
        assert band <= 2 * np.diff(m.row_ptr).max() + 2               # neighbours live in adjacent z-rows
    assert np.all(np.diff(m.nodes[:, 0]) >= 0)                        # z-major node order


def test_refinement_knob_scales_dof_count():
    cfg = load_cfg("geballe_no_diamond")
    mats, _, _ = problem.stack_no_diamond(cfg)
    n = [triangulate_rectangles([x.boundaries for x in mats], [x.mesh_size for x in mats], size_scale=s).num_nodes
         for s in (8.0, 4.0)]
    assert 3.0 < n[1] / n[0] < 5.0


def test_layout_with_hole_and_invalid_input():
    rects = [[0, 1, 0, 1], [1, 2, 0, 1], [0, 1, 1, 2]]                # L-shape: box [1,2]x[1,2] is empty
    for method in ("quadtree", "rows"):
        m = triangulate_rectangles(rects, [0.25, 0.25, 0.5], method=method)
        assert abs(_signed_area(m.nodes, m.tris).sum() - 3.0) < 1e-12
        assert not np.any((m.nodes[:, 0] > 1 + 1e-12) & (m.nodes[:, 1] > 1 + 1e-12))
        assert np.unique(m.tris).size == m.num_nodes
    with pytest.raises(ValueError):
        triangulate_rectangles(rects, [0.25, 0.25, 0.5], method="gmsh")
    with pytest.raises(ValueError):
        triangulate_rectangles([], [])
    with pytest.raises(ValueError):
        triangulate_rectangles([[0, 1, 0, 1]], [0.0])


def test_msh_roundtrip_and_mesh_class(tmp_path):
    cfg = load_cfg("geballe_no_diamond")
    mats, bounds, _ = problem.stack_no_diamond(cfg)
    mesh = Mesh("mesh.msh", bounds, mats)
    mesh.size_scale = 8.0
    with pytest.raises(RuntimeError):
        mesh.write(str(tmp_path / "early.msh"))
    with contextlib.redirect_stdout(io.StringIO()):
        mesh.build_mesh()
    assert [m._tag for m in mats] == [1, 2, 3, 4, 5] and mesh.material_tags["p_sample"] == 3
    path = str(tmp_path / "mesh.msh")
    mesh.write(path)
    nodes, tris, tag, names = read_msh(path)
    assert np.array_equal(nodes, mesh.mesh.nodes) and np.array_equal(tris, mesh.mesh.tris)
    assert np.array_equal(tag, mesh.mesh.cell_tag)
    assert names == {i + 1: m.name for i, m in enumerate(mats)}
    domain, cell_tags, _ = Mesh.msh_to_dolfinx(path)
    assert domain.geometry.x.shape == (len(nodes), 3) and np.all(domain.geometry.x[:, 2] == 0)
    assert np.array_equal(cell_tags.values, tag)
    text = open(path).read()
    assert text.startswith("$MeshFormat\n4.1 0 8\n") and '2 3 "p_sample"' in text


def test_read_msh_gmsh_style_file(tmp_path):
    # hand-written MSH 4.1 as gmsh would emit: point/curve entities, a line element block,
    # sparse node tags, physical tag != surface tag
    p = tmp_path / "g.msh"
    p.write_text("""$MeshFormat
4.1 0 8
$EndMeshFormat
$PhysicalNames
1
2 7 "sample"
$EndPhysicalNames
$Entities
1 1 1 0
1 0 0 0 0
1 0 0 0 1 0 0 0 2 1 -1
3 0 0 0 1 1 0 1 7 1 1
$EndEntities
$Nodes
2 4 1 10
1 1 0 2
1
2
0 0 0
1 0 0
2 3 0 2
5
10
1 1 0
0 1 0
$EndNodes
$Elements
2 3 1 3
1 1 1 1
1 1 2
2 3 2 2
2 1 2 5
3 1 5 10
$EndElements
""")
    nodes, tris, tag, names = read_msh(str(p))
    assert nodes.shape == (4, 2) and tris.shape == (2, 3)
    assert np.array_equal(tag, [7, 7]) and names == {7: "sample"}
    assert np.array_equal(tris, [[0, 1, 2], [0, 2, 3]])
    with pytest.raises(ValueError):
        bad = tmp_path / "b.msh"
        bad.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n")
        read_msh(str(bad))


def test_material_validation():
    with pytest.raises(TypeError):
        Material(3, [0, 1, 0, 1])
    with pytest.raises(ValueError):
        Material("a", [0, 1, 0])
    with pytest.raises(ValueError):
        Material("a", [1, 0, 0, 1])
    with pytest.raises(TypeError):
        Material("a", [0, 1, 0, 1], mesh_size="x")
    m = Material("a", [0, 1, 0, 2], {"k": 1.0}, 0.1)
    assert m.contains(0.5, 2.0) and not m.contains(1.5, 1.0) and getattr(m, "_tag", None) is None
    dup = Mesh("m", [0, 1, 0, 2], [m, Material("b", [0, 1, 0, 2], {}, 0.1)])
    with pytest.raises(RuntimeError):
        with contextlib.redirect_stdout(io.StringIO()):
            dup.build_mesh()


def test_msh_side_cache_is_used_and_invalidated(tmp_path):
    # Mesh.msh_to_dolfinx keeps the parsed arrays in a binary side file next to the mesh folder (the folder itself keeps
    # the reference's two files); a rewritten mesh.msh (other size / mtime) must not be served from the old cache
    import os
    import time
    from heatflow_b200.mesh_and_materials.mesh import Mesh
    from heatflow_b200.mesh_and_materials.msh_io import write_msh
    folder = tmp_path / "meshes" / "w1"
    folder.mkdir(parents=True)
    path = str(folder / "mesh.msh")
    nodes = np.array([[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    tris = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int32)
    write_msh(path, nodes, tris, np.array([1, 2]), {1: "a", 2: "b"})
    d1, t1, _ = Mesh.msh_to_dolfinx(path)
    side = str(tmp_path / "meshes" / "w1") + ".msh_cache.npz"
    assert os.path.isfile(side) and sorted(os.listdir(folder)) == ["mesh.msh"]
    Mesh._msh_cache.clear()                                      # force the side file to be read
    d2, t2, _ = Mesh.msh_to_dolfinx(path)
    assert np.array_equal(d1.geometry.x, d2.geometry.x) and np.array_equal(d1.cells, d2.cells)
    assert np.array_equal(t1.values, t2.values)
    time.sleep(0.01)
    nodes2 = np.vstack([nodes, [[2.0, 0.5]]])
    tris2 = np.vstack([tris, [[1, 4, 2]]]).astype(np.int32)
    write_msh(path, nodes2, tris2, np.array([1, 2, 2]), {1: "a", 2: "b"})
    d3, t3, _ = Mesh.msh_to_dolfinx(path)
    assert d3.geometry.x.shape[0] == 5 and len(t3.values) == 3
