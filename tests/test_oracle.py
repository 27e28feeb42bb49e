"""CPU tests of the oracle: known answers, exact integrals, dolfinx BC conventions, golden fixtures.

The reference ships no golden vectors for this path, so the oracle is pinned
by (a) hand-computed element matrices, (b) an independent quadrature implementation of the same
bilinear forms, (c) structural properties of the assembled system, (d) committed fixtures.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import build_case, make_oracle
from oracle import heat_oracle as ho

HERE = os.path.dirname(os.path.abspath(__file__))


# ---- (a) seed known-answer test from SURVEY.md section 8c --------------------------------------
def test_reference_triangle_known_answer():
    nodes = np.array([[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])          # (z, r)
    tris = np.array([[0, 1, 2]])
    Me, Ke = ho.element_matrices(nodes, tris, axisymmetric=True)
    M = np.array([[2, 1, 2], [1, 2, 2], [2, 2, 6]]) / 120.0
    K = np.array([[2, -1, -1], [-1, 1, 0], [-1, 0, 1]]) / 6.0
    assert np.allclose(Me[0], M, rtol=0, atol=1e-17)
    assert np.allclose(Ke[0], K, rtol=0, atol=1e-17)
    assert abs(Me[0].sum() - 1.0 / 6.0) < 1e-16                      # |T| * rbar
    P = ho.GradientProjector(nodes, tris)
    bz, br = P.load(np.array([0.0, 1.0, 0.0]))                       # u = z: du/dz = 1
    assert np.allclose(bz, [1 / 24, 1 / 24, 2 / 24], atol=1e-17)
    assert np.allclose(br, 0.0, atol=1e-17)


# ---- (b) independent quadrature of the UFL forms (run_with_diamond.py:328-335) ------------------
def _collapsed_gauss(n=4):
    """Tensor Gauss-Legendre on the unit square collapsed to the reference triangle (exact to degree 2n-2)."""
    g, w = np.polynomial.legendre.leggauss(n)
    g, w = 0.5 * (g + 1), 0.5 * w
    a, b = np.meshgrid(g, g, indexing="ij")
    wa, wb = np.meshgrid(w, w, indexing="ij")
    xi, eta = a.ravel(), (b * (1 - a)).ravel()
    return xi, eta, (wa * wb * (1 - a)).ravel()


def _quadrature_matrices(p, axisymmetric):
    xi, eta, w = _collapsed_gauss(5)
    phi = np.stack([1 - xi - eta, xi, eta])                          # [3, nq]
    J = np.array([p[1] - p[0], p[2] - p[0]]).T                       # d(z,r)/d(xi,eta)
    detJ = abs(np.linalg.det(J))
    gref = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])
    grad = gref @ np.linalg.inv(J)                                   # [3, 2] physical gradients
    r = phi.T @ p[:, 1] if axisymmetric else np.ones_like(xi)
    M = np.einsum("iq,jq,q->ij", phi, phi, w * r) * detJ
    K = (grad @ grad.T) * (w * r).sum() * detJ
    return M, K


@pytest.mark.parametrize("axisymmetric", [True, False])
def test_element_matrices_match_quadrature(axisymmetric):
    rng = np.random.default_rng(7)
    for _ in range(25):
        p = rng.random((3, 2)) * np.array([3e-6, 2e-5]) + np.array([-1e-6, 0.0])
        if abs(np.linalg.det(np.array([p[1] - p[0], p[2] - p[0]]))) < 1e-13:
            continue
        for tri in ([0, 1, 2], [0, 2, 1]):                           # both orientations
            Me, Ke = ho.element_matrices(p, np.array([tri]), axisymmetric)
            Mq, Kq = _quadrature_matrices(p[tri], axisymmetric)
            assert np.allclose(Me[0], Mq, rtol=1e-12, atol=1e-13 * np.abs(Mq).max())
            assert np.allclose(Ke[0], Kq, rtol=1e-12, atol=1e-13 * np.abs(Kq).max())


# ---- (c) structure of the assembled system -----------------------------------------------------
@pytest.fixture(scope="module")
def case():
    return build_case("geballe_with_diamond", 8.0)


@pytest.fixture(scope="module")
def oracle(case):
    return make_oracle(case)


def test_pattern_is_node_adjacency_plus_diagonal(case, oracle):
    n = len(case.nodes)
    t = case.tris
    i = np.concatenate([t[:, a] for a in range(3) for _ in range(3)])
    j = np.concatenate([t[:, b] for _ in range(3) for b in range(3)])
    ref = sp.coo_matrix((np.ones(len(i)), (i, j)), shape=(n, n)).tocsr()
    ref.sum_duplicates()
    ref.sort_indices()
    assert np.array_equal(ref.indptr, oracle.rowptr)
    assert np.array_equal(ref.indices, oracle.col)
    assert oracle.rowptr.dtype == np.int32 and oracle.col.dtype == np.int32


def test_mass_and_stiffness_properties(case):
    n = len(case.nodes)
    ones_c = np.ones(len(case.tris))
    M1, A1, rowptr, col = ho.assemble_operators(case.nodes, case.tris, ones_c, ones_c, 1.0)
    K1 = A1 - M1
    # K 1 = 0, M total = int r dA (exact: sum |T| rbar), symmetry
    assert np.abs(K1 @ np.ones(n)).max() <= 1e-12 * np.abs(K1).max()
    area, _, _, r = ho.p1_geometry(case.nodes, case.tris)
    assert abs(M1.sum() - (area * r.mean(axis=1)).sum()) <= 1e-12 * M1.sum()
    assert abs(M1 - M1.T).max() <= 1e-25 and abs(K1 - K1.T).max() <= 1e-22
    # linear fields are in the kernel of the planar stiffness matrix at interior nodes
    Mp, Ap, _, _ = ho.assemble_operators(case.nodes, case.tris, ones_c, ones_c, 1.0, axisymmetric=False)
    Kp = Ap - Mp
    lin = 3.0 * case.nodes[:, 0] - 2.0 * case.nodes[:, 1]
    zmin, zmax = case.nodes[:, 0].min(), case.nodes[:, 0].max()
    rmin, rmax = case.nodes[:, 1].min(), case.nodes[:, 1].max()
    interior = (case.nodes[:, 0] > zmin) & (case.nodes[:, 0] < zmax) & (case.nodes[:, 1] > rmin) & (case.nodes[:, 1] < rmax)
    assert np.abs((Kp @ lin)[interior]).max() <= 1e-9 * np.abs(Kp).max() * np.abs(lin).max()


def test_dirichlet_treatment_matches_dolfinx_convention(case, oracle):
    A = oracle.A
    bc = oracle.bc_dofs
    assert np.all(A.diagonal()[bc] == 1.0)                           # clean_with_ir.ipynb:724
    assert np.all(np.abs(A[bc]).sum(axis=1) == 1.0)                  # rows zeroed
    assert np.all(np.abs(A[:, bc]).sum(axis=0) == 1.0)               # columns zeroed
    assert np.array_equal(A.indptr, oracle.A0.indptr)                # zeroed entries stay in the pattern
    assert abs(A - A.T).max() == 0.0
    free = np.setdiff1d(np.arange(A.shape[0]), bc)
    assert abs(A[free][:, free] - oracle.A0[free][:, free]).max() == 0.0


def test_last_bc_in_list_wins():
    c = build_case("geballe_no_diamond", 8.0)
    # node (z_heat, r = rmax) is in 'top' (const) and in the inner line (gauss): inner is later in the list
    top, inner = c.oracle_bcs[2][0], c.oracle_bcs[3][0]
    shared = np.intersect1d(top, inner)
    assert shared.size == 1
    O = make_oracle(c)
    assert shared[0] in O.gauss_dofs
    g = O.bc_values(2.0e-6)
    r = c.nodes[shared[0], 1]
    amp = ho.heating_amplitude(2.0e-6, c.heat_t, c.heat_T, c.ic)
    assert g[shared[0]] == ho.gaussian_profile(r, amp, c.ic, c.fwhm)


def test_heating_curve_clamps_and_shift():
    c = build_case("geballe_no_diamond", 16.0)
    t, T = c.heat_t, c.heat_T
    assert ho.heating_amplitude(0.0, t, T, c.ic) == c.ic             # before the first sample: ic_temp
    assert ho.heating_amplitude(t[0], t, T, c.ic) == c.ic
    assert ho.heating_amplitude(1.0, t, T, c.ic) == T[-1] - (T[0] - c.ic)
    mid = 0.5 * (t[10] + t[11])
    assert abs(ho.heating_amplitude(mid, t, T, c.ic) - (0.5 * (T[10] + T[11]) - T[0] + c.ic)) < 1e-9


def test_constant_state_is_invariant(case):
    O = make_oracle(case)
    O.ht, O.hT = np.array([0.0, 1.0]), np.array([case.ic, case.ic])  # heating stays at ic
    for s in range(3):
        u = O.step((s + 1) * case.dt)
    assert np.abs(u / case.ic - 1).max() < 1e-11


def test_lu_solution_satisfies_the_system(case):
    O = make_oracle(case)
    u0 = O.u.copy()
    u = O.step(30 * case.dt)
    res = O.A @ u - O.b
    assert np.abs(res).max() <= 1e-9 * np.abs(O.b).max()
    assert np.array_equal(u[O.bc_dofs], O.g[O.bc_dofs]) or np.abs(u[O.bc_dofs] - O.g[O.bc_dofs]).max() < 1e-9
    assert not np.array_equal(u, u0)


# ---- 1-D path ---------------------------------------------------------------------------------
def test_1d_oracle_matches_dense_solve_and_submesh_extraction():
    c = build_case("geballe_no_diamond", 8.0)
    z, cells, tags, verts = ho.extract_axis_submesh(c.nodes, c.tris, c.cell_tag)
    assert np.all(np.diff(z) > 0) and np.array_equal(cells[:, 0] + 1, cells[:, 1])
    assert np.all(np.abs(c.nodes[verts, 1]) <= 1e-10)
    assert set(np.unique(tags)) == {1, 2, 3, 4, 5}
    assert np.all(np.diff(tags) >= 0)                                # layers stacked in z in tag order
    kap, rc = c.kappa_t[tags - 1], c.rhoc_t[tags - 1]
    heat = [int(np.argmin(np.abs(z - c.heating_z)))]
    O = ho.Oracle1D(z, cells, rc, kap, c.dt, [([0], "const"), ([len(z) - 1], "const"), (heat, "heat")], c.ic,
                    c.heat_t, c.heat_T)
    # tridiagonal
    assert np.all(np.abs(O.col - np.repeat(np.arange(len(z)), np.diff(O.rowptr))) <= 1)
    u_prev = O.u.copy()
    src = np.linspace(0.0, 1e12, len(z))
    t = 25 * c.dt
    u = O.step(t, source=src)
    # dense re-derivation of the same step
    h = np.diff(z)
    n = len(z)
    M = np.zeros((n, n)); K = np.zeros((n, n)); M1 = np.zeros((n, n))
    for e in range(n - 1):
        idx = np.ix_([e, e + 1], [e, e + 1])
        M[idx] += rc[e] * h[e] / 6 * np.array([[2, 1], [1, 2]])
        M1[idx] += h[e] / 6 * np.array([[2, 1], [1, 2]])
        K[idx] += kap[e] / h[e] * np.array([[1, -1], [-1, 1]])
    A0 = M + c.dt * K
    g = np.zeros(n); bc = [0, heat[0], n - 1]
    g[[0, n - 1]] = c.ic
    g[heat[0]] = ho.heating_amplitude(t, c.heat_t, c.heat_T, c.ic)
    b = M @ u_prev + c.dt * M1 @ src - A0[:, bc] @ g[bc]
    A = A0.copy(); A[bc, :] = 0; A[:, bc] = 0; A[bc, bc] = 1; b[bc] = g[bc]
    assert np.abs(np.linalg.solve(A, b) / u - 1).max() < 1e-9


# ---- (d) committed fixtures -------------------------------------------------------------------
@pytest.mark.parametrize("name,cfg", [("no_diamond_s16", "geballe_no_diamond"), ("with_diamond_s16", "geballe_with_diamond")])
def test_oracle_reproduces_golden_fixture(name, cfg):
    gold = np.load(os.path.join(HERE, "golden", f"{name}.npz"))
    c = build_case(cfg, float(gold["size_scale"]))
    assert np.array_equal(c.nodes, gold["nodes"]) and np.array_equal(c.tris, gold["tris"])
    assert np.array_equal(c.cell_tag, gold["cell_tag"])
    assert np.array_equal(c.bc_dofs, gold["bc_dofs"]) and np.array_equal(c.gauss_slot, gold["gauss_slot"])
    O = make_oracle(c)
    assert np.array_equal(O.rowptr, gold["rowptr"]) and np.array_equal(O.col, gold["col"])
    hist, fields = O.run(c.num_steps, gold["watch"], keep_fields=True)
    assert np.abs(hist / gold["hist"] - 1).max() < 1e-11
    for k, f in zip(gold["field_steps"], gold["fields"]):
        assert np.abs(fields[int(k)] / f - 1).max() < 1e-11
