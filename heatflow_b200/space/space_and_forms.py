"""``Space``: the notebook-level workspace of the reference (space/space_and_forms.py:7-283) on the GPU.

Same constructor and helpers - ``build_variational_forms`` (transient, r-weighted, :75-117),
``build_steady_state_variational_forms`` (plain Cartesian ``kappa grad u . grad v dx``, :119-149),
``assemble_matrix`` / ``assemble_vector`` (:154-181), ``assign_material_property`` (:186-228),
``initial_condition`` (:233-266), ``vectorize_callable`` (:272-283).  The reference hands the assembled
PETSc objects to a KSP the caller creates; here the matrix comes back as a scipy CSR copy of the device
operator (for inspection) and the two solves the notebooks do with that KSP are methods:
``step(bcs)`` (one backward-Euler step, ``u_n`` updated in place) and ``solve_steady_state(bcs)``.
Both run the library's Jacobi-PCG kernels through the C-ABI; there is no CPU path.

Coefficients may be DG0 ``Function``s (one value per cell, e.g. from ``assign_material_property``) or
scalars; the source ``f`` a P1 ``Function``, a nodal array or a scalar.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

from .. import fem
from ..solver import HeatSolver


def _cell_values(c, num_cells):
    if hasattr(c, "x"):
        arr = np.asarray(c.x.array, dtype=np.float64)
        if arr.size != num_cells:
            raise ValueError("coefficient Function must live on the DG0 space (one value per cell)")
        return arr
    if hasattr(c, "value"):
        c = c.value
    return np.full(num_cells, float(c))


def _nodal_values(f, num_nodes):
    if f is None:
        return None
    if hasattr(f, "x"):
        arr = np.asarray(f.x.array, dtype=np.float64)
    elif np.ndim(f) == 0:
        arr = np.full(num_nodes, float(getattr(f, "value", f)))
    else:
        arr = np.asarray(f, dtype=np.float64)
    if arr.size != num_nodes:
        raise ValueError("source term must have one value per P1 dof")
    return None if not np.any(arr) else arr


def _bc_arrays(bcs, num_dofs):
    """Sorted unique Dirichlet dofs and their values; later list entries win (dolfinx semantics)."""
    owner_val = {}
    for bc in bcs or []:
        dofs = bc.row_dofs if hasattr(bc, "row_dofs") else bc.dofs
        g = bc._g if hasattr(bc, "_g") else bc.g
        for d, v in zip(np.asarray(dofs).tolist(), np.asarray(g.x.array)[dofs].tolist()):
            owner_val[d] = v
    dofs = np.array(sorted(owner_val), dtype=np.int32)
    if dofs.size and (dofs[0] < 0 or dofs[-1] >= num_dofs):
        raise ValueError("Dirichlet dof out of range")
    return dofs, np.array([owner_val[int(d)] for d in dofs], dtype=np.float64)


class _Form:
    """What ``fem.form`` would return: a description of the form the device kernels assemble."""

    def __init__(self, kind, **kw):
        self.kind = kind
        self.__dict__.update(kw)


class Space:
    def __init__(self, mesh_and_tags, V_family="Lagrange", V_degree=1, Q_family="DG", Q_degree=0, device=0):
        if isinstance(mesh_and_tags, tuple) and len(mesh_and_tags) >= 1:
            self.mesh = mesh_and_tags[0]
            self.cell_tags = mesh_and_tags[1] if len(mesh_and_tags) > 1 else None
            self.facet_tags = mesh_and_tags[2] if len(mesh_and_tags) > 2 else None
        else:
            self.mesh, self.cell_tags, self.facet_tags = mesh_and_tags, None, None
        self.V = fem.functionspace(self.mesh, (V_family, V_degree))
        self.Q = fem.functionspace(self.mesh, (Q_family, Q_degree))
        self.a_form = self.L_form = None
        self.a_form_steady = self.L_form_steady = None
        self.device = device
        self._solver = None
        self._built = None            # signature of the operator currently on the device
        self.rtol, self.max_iters = 1e-14, 200000

    # ------------------------------------------------------------------ forms
    def build_variational_forms(self, rho_c, kappa, u_n, dt, r0, f=None):
        self.a_form = _Form("transient_a", rho_c=rho_c, kappa=kappa, dt=float(dt), r0=float(r0))
        self.L_form = _Form("transient_L", rho_c=rho_c, u_n=u_n, dt=float(dt), r0=float(r0), f=f)
        return self.a_form, self.L_form

    def build_steady_state_variational_forms(self, kappa, f=None):
        self.a_form_steady = _Form("steady_a", kappa=kappa)
        self.L_form_steady = _Form("steady_L", f=f)
        return self.a_form_steady, self.L_form_steady

    # ------------------------------------------------------------------ device operator
    def _device_operator(self, steady, bcs):
        E = self.mesh.cells.shape[0]
        N = self.V.num_dofs
        if steady:
            if self.a_form_steady is None:
                raise RuntimeError("call build_steady_state_variational_forms first")
            kap, rc, dt, r0, axisym = _cell_values(self.a_form_steady.kappa, E), np.zeros(E), 1.0, 0.0, False
        else:
            if self.a_form is None:
                raise RuntimeError("call build_variational_forms first")
            a = self.a_form
            kap, rc, dt, r0, axisym = _cell_values(a.kappa, E), _cell_values(a.rho_c, E), a.dt, a.r0, True
        dofs, vals = _bc_arrays(bcs, N)
        sig = (steady, dt, r0, kap.tobytes(), rc.tobytes(), dofs.tobytes())
        if self._solver is None:
            self._solver = HeatSolver(self.device)
        s = self._solver
        if self._built != sig:
            xy = np.array(self.mesh.geometry.x[:, :2], dtype=np.float64)
            # r = sqrt((x[1] - r0)^2), space_and_forms.py:99.  The weight is applied by shifting / mirroring the radial
            # coordinate, which leaves every triangle's shape intact only when r0 is not strictly inside the mesh (a
            # triangle straddling r0 would be folded: wrong area and gradients, not just a wrong weight)
            if axisym and xy[:, 1].min() < r0 < xy[:, 1].max():
                raise ValueError(f"Space: r0 = {r0} lies strictly inside the radial extent of the mesh "
                                 f"[{xy[:, 1].min()}, {xy[:, 1].max()}]; the r-weighted forms need r0 at or outside an edge")
            xy[:, 1] = np.abs(xy[:, 1] - r0)
            pairs, tag = np.unique(np.stack([kap, rc], axis=1), axis=0, return_inverse=True)
            s.set_mesh(xy, self.mesh.cells, (tag.reshape(-1) + 1).astype(np.int32))
            s.set_materials(np.arange(1, len(pairs) + 1, dtype=np.int32), pairs[:, 0], pairs[:, 1])
            s.set_bcs(dofs, vals)
            s.build_operator(dt, axisymmetric=axisym)
            s.set_solver(rtol=self.rtol, max_iters=self.max_iters)
            self._built = sig
        else:
            s.set_bc_values(vals)
        return s

    # ------------------------------------------------------------------ assembly helpers
    def assemble_matrix(self, bcs):
        """The bilinear form with the Dirichlet treatment of ``bcs`` as a scipy CSR matrix (copy of the
        device operator; the reference returns the PETSc ``Mat``)."""
        s = self._device_operator(False, bcs)
        rowptr, col, a, _, _ = s.csr()
        n = self.V.num_dofs
        return sp.csr_matrix((a, col, rowptr), shape=(n, n))

    def assemble_vector(self, bcs):
        """RHS of the transient form after ``apply_lifting`` and ``set_bc`` (numpy vector)."""
        s = self._device_operator(False, bcs)
        L = self.L_form
        u_n = np.asarray(L.u_n.x.array, dtype=np.float64)
        s.set_source(_nodal_values(L.f, self.V.num_dofs))
        s.set_state(u_n)
        s.step(None)                                    # the RHS is a by-product of the fused step kernel
        b = s.get_rhs()
        s.set_state(u_n)
        return b

    def step(self, bcs):
        """One backward-Euler step: solves ``a(u, v) = L(v)`` and overwrites ``u_n`` with the result
        (what the notebooks do with ``KSP.solve(b, u_n.vector)``).  Returns (iterations, relres)."""
        s = self._device_operator(False, bcs)
        L = self.L_form
        s.set_source(_nodal_values(L.f, self.V.num_dofs))
        s.set_state(np.asarray(L.u_n.x.array, dtype=np.float64))
        out = s.step(None)
        L.u_n.x.array[:] = s.get_state()
        return out

    def solve_steady_state(self, bcs, name="u"):
        """Solve ``int kappa grad u . grad v dx = int f v dx`` with ``bcs``; returns a P1 ``Function``."""
        s = self._device_operator(True, bcs)
        n = self.V.num_dofs
        s.set_source(_nodal_values(self.L_form_steady.f, n))
        dofs, vals = _bc_arrays(bcs, n)
        s.set_state(np.full(n, vals.mean() if vals.size else 0.0))
        self.last_solve = s.step(None)
        u = fem.Function(self.V, name=name)
        u.x.array[:] = s.get_state()
        return u

    def close(self):
        if self._solver is not None:
            self._solver.close()
            self._solver = None

    # ------------------------------------------------------------------ material -> DG0 helper
    def assign_material_property(self, materials, property_name):
        if self.cell_tags is None:
            raise RuntimeError("cell_tags not present – cannot map materials.")
        tag_to_val = {}
        for mat in materials:
            if not hasattr(mat, "tag"):
                raise AttributeError("Material object must have a .tag attribute.")
            if property_name not in mat.properties:
                raise KeyError(f"Property '{property_name}' not found in {mat}.")
            tag_to_val[mat.tag] = mat.properties[property_name]
        values = np.zeros(self.Q.num_dofs, dtype=np.float64)
        tags = np.asarray(self.cell_tags.values)
        cells = np.asarray(self.cell_tags.indices)
        for tag, val in tag_to_val.items():
            values[cells[tags == tag]] = val             # unknown tags stay 0, as in the reference
        f_prop = fem.Function(self.Q)
        f_prop.x.array[:] = values
        return f_prop

    # ------------------------------------------------------------------ initial condition helper
    def initial_condition(self, init, *, name="u0"):
        f_ic = fem.Function(self.V, name=name)
        x = self.V.tabulate_dof_coordinates().T
        if isinstance(init, (int, float, np.number)):
            f_ic.x.array[:] = float(init)
            return f_ic
        if callable(init):
            try:
                vals = np.asarray(init(x), dtype=np.float64)
                if vals.shape != (x.shape[1],):
                    raise ValueError
            except Exception:
                vals = self.vectorize_callable(init)(x)
            f_ic.x.array[:] = vals
            return f_ic
        arr = np.asarray(init, dtype=np.float64)
        if arr.size != f_ic.x.array.size:
            raise ValueError("Array length does not match number of DOFs.")
        f_ic.x.array[:] = arr.reshape(-1)
        return f_ic

    @staticmethod
    def vectorize_callable(func):
        def wrapper(points):
            xx, yy = points[0], points[1]
            return np.array([func(xi, yi) for xi, yi in zip(xx, yy)], dtype=np.float64)
        return wrapper
