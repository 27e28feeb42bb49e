"""CPU oracle: numpy/scipy restatement of heatflow's per-timestep FEM heat-conduction path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``heatflow_b200/``) may import
this module; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, as the checker / the timed CPU arm.

PARITY UNPINNED.  The reference's arithmetic lives in un-vendored third-party packages
(dolfinx/UFL/FFCx/basix, petsc4py + MUMPS, gmsh; no lock file - versions inferred as
dolfinx ~0.7-0.9, gmsh 4.13.1) that are not installed here, and the reference ships no
meshes, outputs, golden vectors or asserting tests for this path.  The
oracle therefore restates the published algorithms and is anchored on (i) the reference's
call sites cited below, (ii) exact polynomial integrals / known-answer element matrices,
(iii) an independent quadrature implementation of the same forms (tests/test_oracle.py).
tools/make_reference_goldens.py is the committed recipe that pins it on a machine where the reference's stack runs
(tests/test_reference_goldens.py consumes its tests/golden/ref_*.npz files; none exists yet).

What is restated (reference file:line):
  * forms                run_with_diamond.py:321-337, space/space_and_forms.py:98-117
                         a(u,v) = int rho_c u v r + dt int kappa grad u . grad v r ; r = x[1]
                         L(v)   = int rho_c u_n v r            (f == 0, :326)
  * assembly with BCs    run_with_diamond.py:381-382  assemble_matrix(lhs_form, bcs):
                         rows+cols of BC dofs zeroed, unit diagonal (dolfinx convention)
  * BC dof location      dirichlet_bc/bc.py:32-118 (np.isclose incl. default rtol=1e-5, +1e-14)
  * BC list / last wins  run_with_diamond.py:361-374
  * heating curve        run_with_diamond.py:343-359
  * time loop            run_with_diamond.py:469-493 (bc.update, assemble_vector,
                         apply_lifting, set_bc, direct LU solve)
  * LU                   run_with_diamond.py:389-394 PREONLY+LU(MUMPS) -> scipy splu (SuperLU)
  * watchers             run_with_diamond.py:443-449 (cKDTree nearest node)
  * gradient projection  run_no_diamond.py:471-491, 544-566, 494-513, 457-465
  * 1-D path             run_no_diamond_1d.py:30-164, 518-546, 573-607, 658-767
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
from scipy.spatial import cKDTree


# ----------------------------------------------------------------------------------------
# element matrices (closed form; exact for affine r)
# ----------------------------------------------------------------------------------------
def p1_geometry(nodes, tris):
    """Signed-area-free geometry of every triangle: |T|, gradients of the 3 hat functions."""
    p = nodes[tris]                                   # [E,3,2]  (z, r)
    z, r = p[:, :, 0], p[:, :, 1]
    # b_i = r_j - r_k, c_i = z_k - z_j  (i,j,k cyclic); grad phi_i = (b_i, c_i) / (2 * signed area)
    b = np.stack([r[:, 1] - r[:, 2], r[:, 2] - r[:, 0], r[:, 0] - r[:, 1]], axis=1)
    c = np.stack([z[:, 2] - z[:, 1], z[:, 0] - z[:, 2], z[:, 1] - z[:, 0]], axis=1)
    det = b[:, 0] * c[:, 1] - b[:, 1] * c[:, 0]       # = 2 * signed area
    area = 0.5 * np.abs(det)
    gz = b / det[:, None]
    gr = c / det[:, None]
    return area, gz, gr, r


def element_matrices(nodes, tris, axisymmetric=True):
    """Unit-coefficient P1 mass Me[E,3,3] and stiffness Ke[E,3,3].

    Axisymmetric (weight r = x[1], run_with_diamond.py:322):
      M_ii = |T| (r_i/10 + (r_j + r_k)/30),  M_ij = |T| ((r_i + r_j)/30 + r_k/60)
      K_ij = |T| rbar grad phi_i . grad phi_j,  rbar = (r_1+r_2+r_3)/3
    Planar (weight 1): M = |T|/12 (1 + delta_ij), K = |T| grad phi_i . grad phi_j.
    """
    area, gz, gr, r = p1_geometry(nodes, tris)
    E = len(tris)
    G = gz[:, :, None] * gz[:, None, :] + gr[:, :, None] * gr[:, None, :]
    Me = np.empty((E, 3, 3))
    if axisymmetric:
        rsum = r.sum(axis=1)
        for i in range(3):
            for j in range(3):
                if i == j:
                    Me[:, i, j] = area * (r[:, i] / 10.0 + (rsum - r[:, i]) / 30.0)
                else:
                    k = 3 - i - j
                    Me[:, i, j] = area * ((r[:, i] + r[:, j]) / 30.0 + r[:, k] / 60.0)
        Ke = (area * rsum / 3.0)[:, None, None] * G
    else:
        Me[:] = (area / 12.0)[:, None, None] * (np.ones((3, 3)) + np.eye(3))
        Ke = area[:, None, None] * G
    return Me, Ke


# ----------------------------------------------------------------------------------------
# sparsity pattern + assembly
# ----------------------------------------------------------------------------------------
def csr_pattern(num_nodes, cells):
    """Node-node adjacency incl. diagonal, sorted columns (dolfinx create_sparsity_pattern
    for a P1 bilinear form, implicit in assemble_matrix, run_with_diamond.py:381)."""
    nv = cells.shape[1]
    rows = np.repeat(cells, nv, axis=1).ravel().astype(np.int64)
    cols = np.tile(cells, (1, nv)).ravel().astype(np.int64)
    key = np.unique(rows * num_nodes + cols)
    r = key // num_nodes
    c = key % num_nodes
    rowptr = np.zeros(num_nodes + 1, dtype=np.int64)
    np.add.at(rowptr, r + 1, 1)
    # every node of a valid mesh has at least its diagonal; isolated nodes get none (as dolfinx)
    return np.cumsum(rowptr).astype(np.int32), c.astype(np.int32)


def assemble_csr(num_nodes, cells, elem_mats, rowptr, col):
    """Sum element matrices [E,nv,nv] into the given CSR pattern (contributions added in
    ascending cell order, the order the GPU gather-assembly uses)."""
    nv = cells.shape[1]
    rows = np.repeat(cells, nv, axis=1).ravel().astype(np.int64)
    cols = np.tile(cells, (1, nv)).ravel().astype(np.int64)
    key = rows * num_nodes + cols
    order = np.argsort(key, kind="stable")            # stable: ascending cell index per slot
    key_s = key[order]
    vals_s = elem_mats.reshape(-1)[order]
    pat_key = np.repeat(np.arange(num_nodes, dtype=np.int64), np.diff(rowptr)) * num_nodes + col
    slot = np.searchsorted(pat_key, key_s)
    data = np.zeros(len(col))
    # sequential accumulation in sorted order (np.add.at processes indices in order)
    np.add.at(data, slot, vals_s)
    return sp.csr_matrix((data, col.copy(), rowptr.copy()), shape=(num_nodes, num_nodes))


def assemble_operators(nodes, cells, rho_c, kappa, dt, axisymmetric=True):
    """Return (M, A0, rowptr, col): M = mass with rho_c, A0 = M + dt*K(kappa); no BCs yet."""
    n = nodes.shape[0]
    if cells.shape[1] == 3:
        Me, Ke = element_matrices(nodes, cells, axisymmetric)
    else:
        Me, Ke = element_matrices_1d(nodes, cells)
    rowptr, col = csr_pattern(n, cells)
    Mrc = assemble_csr(n, cells, rho_c[:, None, None] * Me, rowptr, col)
    A0 = assemble_csr(n, cells, rho_c[:, None, None] * Me + dt * (kappa[:, None, None] * Ke), rowptr, col)
    return Mrc, A0, rowptr, col


def apply_dirichlet(A0, bc_dofs):
    """dolfinx assemble_matrix(form, bcs): BC rows and columns zeroed (entries stay in the
    pattern), diagonal exactly 1.0."""
    A = A0.copy().tocsr()
    isbc = np.zeros(A.shape[0], dtype=bool)
    isbc[bc_dofs] = True
    row_of = np.repeat(np.arange(A.shape[0]), np.diff(A.indptr))
    kill = isbc[row_of] | isbc[A.indices]
    A.data[kill] = 0.0
    A.data[(row_of == A.indices) & isbc[row_of]] = 1.0
    return A


# ----------------------------------------------------------------------------------------
# Dirichlet dof location (bc.py:32-118) and last-wins resolution
# ----------------------------------------------------------------------------------------
def locate_row_dofs(coords, location, coord=None, length=None, center=None, width=1e-10):
    """coords [N,>=2] (x = z, y = r).  Same predicates as RowDirichletBC.__init__."""
    x, y = coords[:, 0], coords[:, 1]
    xmin, xmax, ymin, ymax = x.min(), x.max(), y.min(), y.max()
    xmid, ymid = 0.5 * (xmin + xmax), 0.5 * (ymin + ymax)
    half = None if length is None else 0.5 * length
    if location in ("x", "y") and center is None:
        center = xmid if location == "x" else ymid

    def centred(vals, c):
        if half is None:
            return np.ones_like(vals, dtype=bool)
        return np.abs(vals - c) <= half + 1e-14

    if location == "left":
        m = np.isclose(x, xmin, atol=width) & centred(y, ymid)
    elif location == "right":
        m = np.isclose(x, xmax, atol=width) & centred(y, ymid)
    elif location == "bottom":
        m = np.isclose(y, ymin, atol=width) & centred(x, xmid)
    elif location == "top":
        m = np.isclose(y, ymax, atol=width) & centred(x, xmid)
    elif location == "x":
        m = np.isclose(x, float(coord), atol=width) & centred(y, center)
    elif location == "y":
        m = np.isclose(y, float(coord), atol=width) & centred(x, center)
    else:
        raise ValueError("Unknown location keyword.")
    dofs = np.flatnonzero(m).astype(np.int32)
    if dofs.size == 0:
        raise RuntimeError("No DOFs found for requested BC location/length.")
    return dofs


def resolve_bcs(num_nodes, bc_dof_lists):
    """Owner BC index per dof (-1 = free); later BCs in the list override earlier ones, as
    dolfinx does for both apply_lifting values and set_bc."""
    owner = np.full(num_nodes, -1, dtype=np.int32)
    for k, dofs in enumerate(bc_dof_lists):
        owner[dofs] = k
    return owner


# ----------------------------------------------------------------------------------------
# heating curve (run_with_diamond.py:254-274, 343-359)
# ----------------------------------------------------------------------------------------
def load_heating(path):
    import pandas as pd
    df = pd.read_csv(path)
    df = (df.sort_values('time')
            .assign(time=pd.to_numeric(df['time'], errors='coerce'),
                    temp=pd.to_numeric(df['temp'], errors='coerce'))
            .dropna(subset=['time', 'temp']).reset_index(drop=True))
    return df['time'].to_numpy(dtype=float), df['temp'].to_numpy(dtype=float)


def heating_amplitude(t, times, temps, ic_temp):
    """heating_offset(t): clamped linear interpolation shifted so the curve starts at ic_temp."""
    return float(np.interp(t, times, temps, left=temps[0], right=temps[-1])) - (temps[0] - ic_temp)


def gaussian_profile(r, amp, ic_temp, fwhm):
    coeff = -4.0 * np.log(2.0) / fwhm ** 2
    return (amp - ic_temp) * np.exp(coeff * (r - 0.0) ** 2) + ic_temp


# ----------------------------------------------------------------------------------------
# 2-D transient solve (run_with_diamond.py:469-493)
# ----------------------------------------------------------------------------------------
class Oracle2D:
    """Backward-Euler P1 solve with a sparse direct LU, factorised once.

    bcs : list of (dofs, kind) in reference list order, kind = 'const' (value ic_temp) or
          'gauss' (Gaussian-in-r profile with time-dependent amplitude).
    """

    def __init__(self, nodes, tris, rho_c_cell, kappa_cell, dt, bcs, ic_temp, fwhm,
                 heat_times, heat_temps, axisymmetric=True):
        self.nodes, self.tris = nodes, tris
        self.n = nodes.shape[0]
        self.dt, self.ic, self.fwhm = float(dt), float(ic_temp), float(fwhm)
        self.ht, self.hT = heat_times, heat_temps
        self.M, self.A0, self.rowptr, self.col = assemble_operators(
            nodes, tris, rho_c_cell, kappa_cell, self.dt, axisymmetric)
        self.owner = resolve_bcs(self.n, [d for d, _ in bcs])
        self.kinds = [k for _, k in bcs]
        self.bc_dofs = np.flatnonzero(self.owner >= 0).astype(np.int32)
        self.gauss_dofs = np.array([d for d in self.bc_dofs if self.kinds[self.owner[d]] == 'gauss'],
                                   dtype=np.int32)
        self.A = apply_dirichlet(self.A0, self.bc_dofs)
        self.lu = None
        self.u = np.full(self.n, self.ic)
        self.g = np.zeros(self.n)                     # bc values, zero on free dofs
        self.A0_bc_cols = self.A0.tocsc()[:, self.bc_dofs].tocsr()

    def factorize(self):
        self.lu = spla.splu(self.A.tocsc())

    def bc_values(self, t):
        g = np.zeros(self.n)
        g[self.bc_dofs] = self.ic
        if self.gauss_dofs.size:
            amp = heating_amplitude(t, self.ht, self.hT, self.ic)
            g[self.gauss_dofs] = gaussian_profile(self.nodes[self.gauss_dofs, 1], amp, self.ic, self.fwhm)
        return g

    def rhs(self, u_n, g):
        b = self.M @ u_n                              # assemble_vector(rhs_form)
        b -= self.A0_bc_cols @ g[self.bc_dofs]        # apply_lifting
        b[self.bc_dofs] = g[self.bc_dofs]             # set_bc
        return b

    def step(self, t):
        if self.lu is None:
            self.factorize()
        self.g = self.bc_values(t)
        self.b = self.rhs(self.u, self.g)
        self.u = self.lu.solve(self.b)
        return self.u

    def run(self, num_steps, watcher_nodes=None, keep_fields=False):
        hist, fields = [], []
        for s in range(num_steps):
            u = self.step((s + 1) * self.dt)
            if watcher_nodes is not None:
                hist.append(u[watcher_nodes].copy())
            if keep_fields:
                fields.append(u.copy())
        return np.array(hist), fields


def nearest_nodes(nodes, points):
    tree = cKDTree(nodes[:, :2])
    return np.array([tree.query(p)[1] for p in points], dtype=np.int32)


# ----------------------------------------------------------------------------------------
# r-weighted L2 projection of grad(u) onto vector P1 (run_no_diamond.py:471-491, 544-550)
# ----------------------------------------------------------------------------------------
class GradientProjector:
    def __init__(self, nodes, tris):
        self.nodes, self.tris = nodes, tris
        n = nodes.shape[0]
        Me, _ = element_matrices(nodes, tris, True)
        rowptr, col = csr_pattern(n, tris)
        self.Mr = assemble_csr(n, tris, Me, rowptr, col)
        self.lu = spla.splu(self.Mr.tocsc())
        self.area, self.gz, self.gr, self.r = p1_geometry(nodes, tris)

    def load(self, u):
        """b_c[i] = sum_T (d_c u)_T |T| (2 r_i + r_j + r_k)/12 for c in (z, r)."""
        ue = u[self.tris]
        dz = (self.gz * ue).sum(axis=1)
        dr = (self.gr * ue).sum(axis=1)
        w = self.area[:, None] * (self.r + self.r.sum(axis=1, keepdims=True)) / 12.0   # [E,3]
        n = self.nodes.shape[0]
        bz = np.zeros(n)
        br = np.zeros(n)
        np.add.at(bz, self.tris.ravel(), (w * dz[:, None]).ravel())
        np.add.at(br, self.tris.ravel(), (w * dr[:, None]).ravel())
        return bz, br

    def project(self, u):
        bz, br = self.load(u)
        return np.column_stack((self.lu.solve(bz), self.lu.solve(br)))


def radial_bins(nodes, dz_bin=0.2e-6, band=0.25e-6):
    """z-bins of nodes with 0 < r <= band (run_no_diamond.py:494-513)."""
    z_min, z_max = nodes[:, 0].min(), nodes[:, 0].max()
    edges = np.arange(z_min, z_max + dz_bin, dz_bin)
    mask = (nodes[:, 1] > 0.0) & (nodes[:, 1] <= band)
    groups = [[] for _ in range(len(edges) - 1)]
    for n in np.flatnonzero(mask):
        k = np.searchsorted(edges, nodes[n, 0]) - 1
        if 0 <= k < len(groups):
            groups[k].append(n)
    centres = [0.5 * (edges[k] + edges[k + 1]) for k, g in enumerate(groups) if g]
    return centres, [np.array(g) for g in groups if g]


def axis_nodes(nodes, tol=1e-12):
    """Nodes on r = 0 sorted by z (run_no_diamond.py:457-465)."""
    idx = np.flatnonzero(np.abs(nodes[:, 1]) <= tol)
    idx = idx[np.argsort(nodes[idx, 0])]
    return idx, nodes[idx, 0]


# ----------------------------------------------------------------------------------------
# 1-D path (run_no_diamond_1d.py)
# ----------------------------------------------------------------------------------------
def extract_axis_submesh(nodes, tris, cell_tag, tol=1e-10):
    """Edges of the 2-D mesh with both vertices on |r| <= tol, as an interval mesh ordered by z;
    tag = tag of the lowest-index 2-D cell containing the edge (run_no_diamond_1d.py:30-164)."""
    on = np.abs(nodes[:, 1]) <= tol
    best = {}
    for e in range(len(tris)):
        t = tris[e]
        for a, b in ((t[0], t[1]), (t[1], t[2]), (t[2], t[0])):
            if on[a] and on[b]:
                key = (min(a, b), max(a, b))
                if key not in best:
                    best[key] = cell_tag[e]
    if not best:
        raise ValueError("No facets found on the r=0 axis. Check tolerance or mesh.")
    verts = np.array(sorted({v for k in best for v in k}, key=lambda v: nodes[v, 0]))
    new_id = {v: i for i, v in enumerate(verts)}
    edges = sorted(best.items(), key=lambda kv: min(nodes[kv[0][0], 0], nodes[kv[0][1], 0]))
    cells, tags = [], []
    for (a, b), tg in edges:
        ia, ib = new_id[a], new_id[b]
        if nodes[a, 0] > nodes[b, 0]:
            ia, ib = ib, ia
        cells.append((ia, ib))
        tags.append(tg)
    return nodes[verts, 0].copy(), np.array(cells, dtype=np.int32), np.array(tags, dtype=np.int32), verts


def element_matrices_1d(z, cells):
    """Un-weighted interval P1 (run_no_diamond_1d.py:537-544): M = h/6 [[2,1],[1,2]], K = 1/h [[1,-1],[-1,1]]."""
    zz = z[:, 0] if z.ndim == 2 else z
    h = np.abs(zz[cells[:, 1]] - zz[cells[:, 0]])
    Me = (h / 6.0)[:, None, None] * np.array([[2.0, 1.0], [1.0, 2.0]])
    Ke = (1.0 / h)[:, None, None] * np.array([[1.0, -1.0], [-1.0, 1.0]])
    return Me, Ke


class Oracle1D:
    """1-D (un-weighted) backward Euler with optional nodal source s:
    A = int rho_c u v + dt int kappa u' v' ; b = M_rc u_n + dt M_1 s  (run_no_diamond_1d.py:537-546).
    BCs: left/right = ic, heating node(s) = spatially uniform heating_offset(t) (:573-591)."""

    def __init__(self, z, cells, rho_c_cell, kappa_cell, dt, bc_lists, ic_temp, heat_times, heat_temps):
        self.z = np.asarray(z, dtype=float)
        self.n = len(self.z)
        self.dt, self.ic = float(dt), float(ic_temp)
        self.ht, self.hT = heat_times, heat_temps
        zz = self.z[:, None]
        self.M, self.A0, self.rowptr, self.col = assemble_operators(zz, cells, rho_c_cell, kappa_cell, dt)
        Me, _ = element_matrices_1d(self.z, cells)
        self.M1 = assemble_csr(self.n, cells, Me, self.rowptr, self.col)
        self.owner = resolve_bcs(self.n, [d for d, _ in bc_lists])
        self.kinds = [k for _, k in bc_lists]
        self.bc_dofs = np.flatnonzero(self.owner >= 0).astype(np.int32)
        self.heat_dofs = np.array([d for d in self.bc_dofs if self.kinds[self.owner[d]] == 'heat'], dtype=np.int32)
        self.A = apply_dirichlet(self.A0, self.bc_dofs)
        self.lu = spla.splu(self.A.tocsc())
        self.A0_bc_cols = self.A0.tocsc()[:, self.bc_dofs].tocsr()
        self.u = np.full(self.n, self.ic)

    def step(self, t, source=None):
        g = np.zeros(self.n)
        g[self.bc_dofs] = self.ic
        g[self.heat_dofs] = heating_amplitude(t, self.ht, self.hT, self.ic)
        b = self.M @ self.u
        if source is not None:
            b += self.dt * (self.M1 @ source)
        b -= self.A0_bc_cols @ g[self.bc_dofs]
        b[self.bc_dofs] = g[self.bc_dofs]
        self.u = self.lu.solve(b)
        return self.u
