"""GPU tests of the notebook-level ``Space`` helper (reference: space/space_and_forms.py) against scipy."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from dirichlet_bc.bc import RowDirichletBC
from helpers import build_case
from oracle import heat_oracle as ho
from space.space_and_forms import Space

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nd():
    return build_case("geballe_no_diamond", 8.0)


def make_space(c):
    from heatflow_b200.mesh_and_materials.mesh import MeshTags
    return Space((c.domain, MeshTags(c.cell_tag), None))


def test_transient_step_matches_oracle_and_reference_api(nd):
    c = nd
    sp_ = make_space(c)
    rho_c = sp_.assign_material_property(c.mats, "rho_cv")
    kappa = sp_.assign_material_property(c.mats, "k")
    assert np.array_equal(rho_c.x.array, c.rhoc_c) and np.array_equal(kappa.x.array, c.kappa_c)
    u_n = sp_.initial_condition(c.ic)
    assert np.all(u_n.x.array == c.ic)
    u_f = sp_.initial_condition(lambda x: 300.0 + 1e6 * x[1])
    u_s = sp_.initial_condition(lambda z, r: 300.0 + 1e6 * r)          # scalar callable goes through vectorize_callable
    assert np.allclose(u_f.x.array, u_s.x.array, rtol=0, atol=0)
    with pytest.raises(ValueError):
        sp_.initial_condition(np.zeros(3))
    a, L = sp_.build_variational_forms(rho_c, kappa, u_n, c.dt, 0.0)
    assert sp_.a_form is a and sp_.L_form is L
    bcs = c.bcs
    O = ho.Oracle2D(c.nodes, c.tris, c.rhoc_c, c.kappa_c, c.dt, c.oracle_bcs, c.ic, c.fwhm, c.heat_t, c.heat_T)
    for k in range(8, 14):
        t = (k + 1) * c.dt
        for bc in bcs:
            bc.update(t)
        if k == 8:
            A = sp_.assemble_matrix(bcs)
            assert abs(A - O.A).max() <= 1e-12 * abs(O.A).max()
            b = sp_.assemble_vector(bcs)
            g = O.bc_values(t)
            assert np.abs(b - O.rhs(np.full(len(c.nodes), c.ic), g)).max() <= 1e-12 * np.abs(b).max()
            O.u = np.full(len(c.nodes), c.ic)
        it, rel = sp_.step([bc.bc for bc in bcs])                       # fem.dirichletbc stand-ins work too
        ou = O.step(t)
        assert it > 0 and np.abs(u_n.x.array / ou - 1).max() <= 1e-10
    # an axis r0 strictly inside the mesh would fold the triangles that straddle it: refused, not silently wrong
    sp_.build_variational_forms(rho_c, kappa, u_n, c.dt, 0.5 * c.nodes[:, 1].max())
    with pytest.raises(ValueError, match="r0"):
        sp_.assemble_matrix(bcs)
    sp_.close()


def test_steady_state_matches_sparse_direct_solve(nd):
    c = nd
    sp_ = make_space(c)
    kappa = sp_.assign_material_property(c.mats, "k")
    f = sp_.initial_condition(lambda x: 1e18 * np.exp(-((x[0] - 0.0) ** 2 + x[1] ** 2) / (2e-6) ** 2))
    sp_.build_steady_state_variational_forms(kappa, f)
    left = RowDirichletBC(sp_.V, "left", value=300.0)
    right = RowDirichletBC(sp_.V, "right", value=350.0)
    for bc in (left, right):
        bc.update(0.0)
    u = sp_.solve_steady_state([left, right])
    # scipy restatement: Cartesian stiffness and mass (space_and_forms.py:141-143), Dirichlet rows/cols eliminated
    Me, Ke = ho.element_matrices(c.nodes, c.tris, axisymmetric=False)
    rowptr, col = ho.csr_pattern(len(c.nodes), c.tris)
    K = ho.assemble_csr(len(c.nodes), c.tris, Ke * c.kappa_c[:, None, None], rowptr, col)
    M = ho.assemble_csr(len(c.nodes), c.tris, Me, rowptr, col)
    n = len(c.nodes)
    g = np.zeros(n)
    g[left.row_dofs], g[right.row_dofs] = 300.0, 350.0
    bc_dofs = np.union1d(left.row_dofs, right.row_dofs)
    free = np.setdiff1d(np.arange(n), bc_dofs)
    rhs = (M @ f.x.array - K @ g)[free]
    ref = g.copy()
    ref[free] = spla.spsolve(sp.csc_matrix(K[free][:, free]), rhs)
    assert np.abs(u.x.array / ref - 1).max() <= 1e-9
    assert u.x.array.max() > 351.0                                      # the source heats the interior
    sp_.close()


def test_rmse_helper():
    import analysis_utils as au
    t = np.linspace(0, 1, 11)
    assert au.calculate_rmse(t, 2 * t, t, 2 * t) == 0.0
    assert abs(au.calculate_rmse([0.25, 0.75], [1.0, 1.0], [0.0, 1.0], [0.0, 2.0]) - 0.5) < 1e-15
