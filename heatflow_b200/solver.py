"""Host-side handle on the CUDA heat-conduction solver.

``HeatSolver`` is the Python face of the C-ABI (``include/heatflow_b200.h``).  It stands in for
the third-party seam the reference runners use inline - ``fem.form`` / ``assemble_matrix`` /
``create_vector`` / ``assemble_vector`` / ``apply_lifting`` / ``set_bc`` / ``KSP.solve``
(reference: run_with_diamond.py:336-337, :381-394, :474-481).  All numerics run in the
hand-written sm_100a kernels; this class only marshals numpy arrays.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class HeatSolver:
    def __init__(self, device=0):
        self._L = _lib.load()
        self._h = self._L.hf_create(int(device))
        if not self._h:
            raise _lib.HeatflowError(-2, self._L.hf_last_error().decode())
        self.n = 0
        self.n_bc = 0
        self.bc_dofs = np.zeros(0, np.int32)
        self.ens_batch = 0

    # -- lifetime -------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._L.hf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- problem set-up -------------------------------------------------------------
    def set_ordering(self, ordering):
        """Internal node order of the device data: 'auto', 'given' or 'hilbert' (before set_mesh)."""
        code = {"auto": 0, "given": 1, "hilbert": 2}.get(ordering, ordering)
        _lib.check(self._L.hf_set_ordering(self._h, int(code)))

    def set_mesh(self, nodes, cells, cell_tag):
        """nodes [N,2] (z, r) or [N] / [N,1] for interval meshes; cells [E,3] or [E,2]."""
        nodes = _f64(nodes)
        if nodes.ndim == 1 or nodes.shape[1] == 1:
            xy = np.zeros((nodes.shape[0], 2))
            xy[:, 0] = nodes.reshape(-1)
            nodes = xy
        nodes = _f64(nodes[:, :2])
        cells = _i32(cells)
        cell_tag = _i32(cell_tag)
        if cells.ndim != 2 or cells.shape[1] not in (2, 3) or cell_tag.shape[0] != cells.shape[0]:
            raise ValueError("cells must be [E,3] or [E,2] with one tag per cell")
        _lib.check(self._L.hf_set_mesh(self._h, nodes.shape[0], cells.shape[0], cells.shape[1],
                                       _lib.ptr(nodes), _lib.ptr(cells), _lib.ptr(cell_tag)))
        self.n = nodes.shape[0]
        self.nodes = nodes

    def set_materials(self, tags, kappa, rho_c):
        tags, kappa, rho_c = _i32(tags), _f64(kappa), _f64(rho_c)
        _lib.check(self._L.hf_set_materials(self._h, len(tags), _lib.ptr(tags), _lib.ptr(kappa), _lib.ptr(rho_c)))

    def set_bcs(self, bc_dofs, bc_value, gauss_slot=None, gauss_r=None):
        bc_dofs, bc_value = _i32(bc_dofs), _f64(bc_value)
        gauss_slot = _i32(gauss_slot if gauss_slot is not None else [])
        gauss_r = _f64(gauss_r if gauss_r is not None else [])
        _lib.check(self._L.hf_set_bcs(self._h, len(bc_dofs), _lib.ptr(bc_dofs), _lib.ptr(bc_value),
                                      len(gauss_slot), _lib.ptr(gauss_slot), _lib.ptr(gauss_r)))
        self.n_bc = len(bc_dofs)
        self.bc_dofs = bc_dofs

    def set_bc_values(self, bc_value):
        bc_value = _f64(bc_value)
        if bc_value.shape[0] != self.n_bc:
            raise ValueError("one value per Dirichlet dof expected")
        _lib.check(self._L.hf_set_bc_values(self._h, _lib.ptr(bc_value)))

    def build_operator(self, dt, axisymmetric=True):
        _lib.check(self._L.hf_build_operator(self._h, float(dt), 1 if axisymmetric else 0))

    def set_solver(self, rtol=1e-14, max_iters=20000, warm=0.0, mode=0):
        _lib.check(self._L.hf_set_solver(self._h, float(rtol), int(max_iters), float(warm), int(mode)))

    # -- inspection -----------------------------------------------------------------
    def set_recycle(self, max_vectors):
        """Start every solve from the projection of its RHS onto the corrections of up to `max_vectors`
        earlier solves of this simulation (dropped by ``set_state`` / ``build_operator``)."""
        _lib.check(self._L.hf_set_recycle(self._h, int(max_vectors)))

    def solver_path(self):
        """1 = streaming kernel (one launch per iteration), 2 = persistent streaming kernel, 3 = on-chip patch kernel
        (classic CG), 4 = on-chip patch kernel (pipelined CG)."""
        rc = self._L.hf_get_solver_path(self._h)
        if rc < 0:
            _lib.check(rc)
        return rc

    def on_chip(self):
        """True when the mesh fits on chip and the solves run in the on-chip patch kernel."""
        return self.solver_path() >= 3

    def sizes(self):
        n, nnz = C.c_int32(), C.c_int64()
        _lib.check(self._L.hf_get_sizes(self._h, C.byref(n), C.byref(nnz)))
        return n.value, nnz.value

    def csr(self, values=True):
        """(rowptr, col[, A_bc, M, A0]) exactly as stored on the device."""
        n, nnz = self.sizes()
        rowptr = np.empty(n + 1, np.int32)
        col = np.empty(nnz, np.int32)
        if not values:
            _lib.check(self._L.hf_get_csr(self._h, _lib.ptr(rowptr), _lib.ptr(col), None, None, None))
            return rowptr, col
        a, m, a0 = np.empty(nnz), np.empty(nnz), np.empty(nnz)
        _lib.check(self._L.hf_get_csr(self._h, _lib.ptr(rowptr), _lib.ptr(col), _lib.ptr(a), _lib.ptr(m), _lib.ptr(a0)))
        return rowptr, col, a, m, a0

    # -- state ----------------------------------------------------------------------
    def set_state(self, u):
        u = _f64(np.broadcast_to(u, (self.n,)))
        _lib.check(self._L.hf_set_state(self._h, _lib.ptr(u)))

    def get_state(self):
        u = np.empty(self.n)
        _lib.check(self._L.hf_get_state(self._h, _lib.ptr(u)))
        return u

    def get_rhs(self):
        b = np.empty(self.n)
        _lib.check(self._L.hf_get_rhs(self._h, _lib.ptr(b)))
        return b

    def set_source(self, s):
        if s is None:
            _lib.check(self._L.hf_set_source(self._h, None))
        else:
            s = _f64(s)
            _lib.check(self._L.hf_set_source(self._h, _lib.ptr(s)))

    # -- time stepping --------------------------------------------------------------
    def step(self, amp=None, t_ic=0.0, coeff=0.0):
        """One implicit step.  ``amp=None`` keeps the Dirichlet values as last set."""
        it, rel = C.c_int32(), C.c_double()
        use = 0 if amp is None else 1
        _lib.check(self._L.hf_step(self._h, use, float(amp or 0.0), float(t_ic), float(coeff), C.byref(it), C.byref(rel)))
        return it.value, rel.value

    def run(self, amps, t_ic, coeff, watch_nodes=(), keep_fields=False):
        """All steps on the device; returns (hist [S, n_watch], iters [S], fields [S, N] | None)."""
        amps = _f64(amps)
        watch = _i32(watch_nodes)
        S = len(amps)
        hist = np.empty((S, len(watch)))
        iters = np.empty(S, np.int32)
        fields = np.empty((S, self.n)) if keep_fields else None
        _lib.check(self._L.hf_run(self._h, S, _lib.ptr(amps), float(t_ic), float(coeff), len(watch), _lib.ptr(watch),
                                  _lib.ptr(hist), _lib.ptr(fields), _lib.ptr(iters)))
        return hist, iters, fields

    def sample(self, nodes):
        nodes = _i32(nodes)
        out = np.empty(len(nodes))
        _lib.check(self._L.hf_sample(self._h, len(nodes), _lib.ptr(nodes), _lib.ptr(out)))
        return out

    def stats(self):
        """dict: device ms of the last run loop, kernels launched, PCG iterations, last relres."""
        st = np.zeros(5)
        _lib.check(self._L.hf_get_stats(self._h, _lib.ptr(st)))
        return {"run_ms": st[0], "launches": int(st[1]), "iterations": int(st[2]), "relres": st[3], "retries": int(st[4])}

    def set_sharing(self, n_concurrent):
        """Plan the on-chip kernel for ``n_concurrent`` (1 or 2) simulations sharing the GPU (call before
        ``build_operator``); used by the sweep engine, which drives two contexts from two threads."""
        _lib.check(self._L.hf_set_sharing(self._h, int(n_concurrent)))

    def set_profile(self, on=True):
        """CUDA events around every PCG solve of ``run`` (see ``solve_profile``)."""
        _lib.check(self._L.hf_set_profile(self._h, 1 if on else 0))

    def solve_profile(self):
        """(device ms inside the PCG solves of the last ``run``, solver kernels launched there)."""
        ms, n = C.c_double(), C.c_int64()
        _lib.check(self._L.hf_get_solve_profile(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def project_gradient(self):
        g = np.empty((self.n, 2))
        it = C.c_int32()
        _lib.check(self._L.hf_project_gradient(self._h, _lib.ptr(g), C.byref(it)))
        self.last_projection_iters = it.value
        return g

    def spmv(self, x):
        x = _f64(x)
        y = np.empty(self.n)
        _lib.check(self._L.hf_spmv(self._h, _lib.ptr(x), _lib.ptr(y)))
        return y

    def bench_kernels(self, reps=20, flush_l2=True):
        ms = np.zeros(2, np.float32)
        _lib.check(self._L.hf_bench_kernels(self._h, int(reps), 1 if flush_l2 else 0, ms.ctypes.data_as(C.c_void_p)))
        return float(ms[0]), float(ms[1])

    # -- ensemble -------------------------------------------------------------------
    def ens_create(self, k_sample, coeff, sample_tag):
        k_sample, coeff = _f64(k_sample), _f64(coeff)
        if k_sample.shape != coeff.shape:
            raise ValueError("k_sample and coeff must have one entry per variant")
        _lib.check(self._L.hf_ens_create(self._h, len(k_sample), _lib.ptr(k_sample), _lib.ptr(coeff), int(sample_tag)))
        self.ens_batch = len(k_sample)

    def ens_run(self, amps, t_ic, watch_nodes=()):
        amps = _f64(amps)
        watch = _i32(watch_nodes)
        S = len(amps)
        hist = np.empty((self.ens_batch, S, len(watch)))
        iters = np.empty(S, np.int32)
        _lib.check(self._L.hf_ens_run(self._h, S, _lib.ptr(amps), float(t_ic), len(watch), _lib.ptr(watch),
                                      _lib.ptr(hist), _lib.ptr(iters)))
        return hist, iters

    def ens_get_state(self):
        u = np.empty((self.ens_batch, self.n))
        _lib.check(self._L.hf_ens_get_state(self._h, _lib.ptr(u)))
        return u

    def ens_path(self):
        """1: streaming ensemble kernels, 5: batched on-chip kernel (one launch per time step)."""
        p = self._L.hf_ens_get_path(self._h)
        if p < 0:
            _lib.check(p)
        return p

    def ens_destroy(self):
        _lib.check(self._L.hf_ens_destroy(self._h))
        self.ens_batch = 0
