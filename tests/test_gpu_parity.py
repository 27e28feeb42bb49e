"""GPU parity tests: the CUDA path (through the C-ABI) against the scipy oracle on the same inputs.

Bars (BASELINE.json north_star): dof map and sparsity pattern bit-exact; fp64 temperatures
within 1e-10 relative at every output step.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import build_case, make_oracle, make_solver
from heatflow_b200 import _lib
from oracle import heat_oracle as ho

pytestmark = pytest.mark.gpu

RTOL_FIELD = 1e-10   # the tolerance north_star states for temperature histories
# solver modes (hf_set_solver): 1 = streaming kernel, one launch per PCG iteration; 2 = persistent streaming kernel,
# one cooperative launch per solve; 3 = on-chip patch kernel (pipelined CG where the mesh leaves room for it; meant for
# Hilbert-ordered meshes, the default); 4 = on-chip patch kernel, classic CG
ORDERING = {0: "auto", 1: "auto", 2: "auto", 3: "hilbert", 4: "hilbert"}
MODES = pytest.mark.parametrize("mode", [1, 2, 3, 4], ids=["streaming", "stream-persistent", "patch-pipelined", "patch-classic"])


def rel_err(a, b):
    return np.abs(a - b).max() / np.abs(b).max()


@pytest.fixture(scope="module")
def small_nd():
    return build_case("geballe_no_diamond", 8.0)


@pytest.fixture(scope="module")
def small_wd():
    return build_case("geballe_with_diamond", 8.0)


@pytest.mark.parametrize("name", ["geballe_no_diamond", "geballe_with_diamond"])
def test_pattern_bit_exact_and_values(name):
    c = build_case(name, 8.0)
    s = make_solver(c)
    O = make_oracle(c)
    rowptr, col, A, M, A0 = s.csr()
    assert rowptr.dtype == np.int32 and col.dtype == np.int32
    assert np.array_equal(rowptr, O.rowptr)
    assert np.array_equal(col, O.col)
    for got, want in ((M, O.M.data), (A0, O.A0.data), (A, O.A.data)):
        row_of = np.repeat(np.arange(len(c.nodes)), np.diff(rowptr))
        rowmax = np.maximum.reduceat(np.abs(want), rowptr[:-1])[row_of]
        assert np.all(np.abs(got - want) <= 1e-13 * rowmax)
    # Dirichlet rows: exactly identity (dolfinx convention, clean_with_ir.ipynb:724)
    Acsr = sp.csr_matrix((A, col, rowptr), shape=(len(c.nodes),) * 2)
    d = Acsr.diagonal()
    assert np.all(d[c.bc_dofs] == 1.0)
    assert np.all(np.abs(Acsr[c.bc_dofs]).sum(axis=1) == 1.0)
    assert abs(Acsr - Acsr.T).max() == 0.0
    s.close()


def test_spmv_matches_scipy(small_wd):
    c = small_wd
    s = make_solver(c)
    O = make_oracle(c)
    rng = np.random.default_rng(0)
    x = rng.standard_normal(len(c.nodes))
    y = s.spmv(x)
    want = O.A @ x
    scale = np.abs(O.A) @ np.abs(x)
    assert np.all(np.abs(y - want) <= 1e-14 * scale)
    s.close()


def test_rhs_and_single_step(small_nd):
    c = small_nd
    s = make_solver(c)
    O = make_oracle(c)
    # a heated step (amplitude well above ic) from a non-trivial state
    rng = np.random.default_rng(1)
    u0 = c.ic + 50.0 * rng.random(len(c.nodes))
    s.set_state(u0)
    O.u = u0.copy()
    k = 20
    t = (k + 1) * c.dt
    its, rel = s.step(amp=c.amps[k], t_ic=c.ic, coeff=c.coeff)
    uo = O.step(t)
    assert its > 0 and rel <= 1e-14
    b = s.get_rhs()
    scale = abs(O.M) @ np.abs(u0) + abs(O.A0) @ np.abs(O.g)      # row-wise magnitude of the terms summed
    free = np.setdiff1d(np.arange(len(c.nodes)), c.bc_dofs)
    assert np.all(np.abs(b - O.b)[free] <= 1e-14 * scale[free])
    assert np.array_equal(b[c.bc_dofs], O.b[c.bc_dofs]) or rel_err(b[c.bc_dofs], O.b[c.bc_dofs]) < 1e-15
    assert rel_err(s.get_state(), uo) <= RTOL_FIELD
    s.close()


@MODES
@pytest.mark.parametrize("name", ["geballe_no_diamond", "geballe_with_diamond"])
def test_history_every_step(name, mode):
    c = build_case(name, 8.0)
    s = make_solver(c, mode=mode, ordering=ORDERING[mode])
    O = make_oracle(c)
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 0.5 * 6.2e-8, 0.0), (0.9e-6, 0.0)])
    hist, iters, fields = s.run(c.amps, c.ic, c.coeff, watch, keep_fields=True)
    ohist, ofields = O.run(c.num_steps, watch, keep_fields=True)
    worst = max(rel_err(f, of) for f, of in zip(fields, ofields))
    assert worst <= RTOL_FIELD, worst
    # pointwise relative error too (temperatures are >= 300 K everywhere)
    assert max(np.abs(f / of - 1).max() for f, of in zip(fields, ofields)) <= RTOL_FIELD
    assert np.abs(hist / ohist - 1).max() <= RTOL_FIELD
    assert np.array_equal(hist, np.array([f[watch] for f in fields]))
    assert iters.max() > 0
    s.close()


def test_hilbert_internal_ordering_is_invisible(small_wd):
    """Every array crossing the C-ABI stays in the caller's numbering when the library renumbers
    the nodes internally (Hilbert order): pattern, values, SpMV, RHS, histories, fields, gradient."""
    c = small_wd
    s = make_solver(c, ordering="hilbert")
    g = make_solver(c, ordering="given")
    O = make_oracle(c)
    rowptr, col, A, M, A0 = s.csr()
    rowptr_g, col_g, A_g, M_g, A0_g = g.csr()
    assert np.array_equal(rowptr, O.rowptr) and np.array_equal(col, O.col)
    assert np.array_equal(A, A_g) and np.array_equal(M, M_g) and np.array_equal(A0, A0_g)   # same bits, same slots
    x = np.random.default_rng(3).standard_normal(len(c.nodes))
    y, want = s.spmv(x), O.A @ x
    assert np.all(np.abs(y - want) <= 1e-14 * (np.abs(O.A) @ np.abs(x)))
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 0.5 * 6.2e-8, 0.0), (0.9e-6, 0.0)])
    hist, iters, fields = s.run(c.amps[:30], c.ic, c.coeff, watch, keep_fields=True)
    ohist, ofields = O.run(30, watch, keep_fields=True)
    assert max(np.abs(f / of - 1).max() for f, of in zip(fields, ofields)) <= RTOL_FIELD
    assert np.abs(hist / ohist - 1).max() <= RTOL_FIELD
    assert np.array_equal(hist, np.array([f[watch] for f in fields]))
    assert np.array_equal(s.get_state(), fields[-1]) and np.array_equal(s.sample(watch), hist[-1])
    assert np.all(np.abs(s.get_rhs() - O.b) <= 1e-13 * np.abs(O.b).max())
    P = ho.GradientProjector(c.nodes, c.tris)
    want = P.project(s.get_state())
    got = s.project_gradient()
    assert np.all(np.abs(got - want).max(axis=0) <= RTOL_FIELD * np.abs(want).max(axis=0))
    u = c.ic + np.arange(len(c.nodes), dtype=float)
    s.set_state(u)
    assert np.array_equal(s.get_state(), u)
    s.close()
    g.close()


def test_constant_state_invariance(small_wd):
    # all BCs = ic and u0 = ic: the state must stay ic to rounding
    c = small_wd
    s = make_solver(c)
    hist, iters, fields = s.run(np.full(5, c.ic), c.ic, c.coeff, [0], keep_fields=True)
    assert np.abs(fields[-1] / c.ic - 1).max() < 1e-12
    s.close()


@MODES
def test_full_size_no_diamond_vs_oracle(mode):
    # configs[1] at the cfg's own mesh sizes (~1.1e5 dofs, 40 steps)
    c = build_case("geballe_no_diamond", 1.0)
    s = make_solver(c, mode=mode, ordering=ORDERING[mode])
    O = make_oracle(c)
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 0.5 * 6.2e-8, 0.0), (0.951e-6, 0.0), (0.0, 5e-6)])
    hist, iters, _ = s.run(c.amps, c.ic, c.coeff, watch)
    ohist, _ = O.run(c.num_steps, watch)
    assert np.abs(hist / ohist - 1).max() <= RTOL_FIELD
    assert rel_err(s.get_state(), O.u) <= RTOL_FIELD
    s.close()


@MODES
def test_run_is_bit_reproducible(small_nd, mode):
    c = small_nd
    out = []
    for _ in range(2):
        s = make_solver(c, mode=mode, ordering=ORDERING[mode])
        hist, iters, _ = s.run(c.amps[:25], c.ic, c.coeff, [5, 50])
        out.append((hist, iters, s.get_state()))
        s.close()
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2], out[1][2])


@MODES
def test_warm_start_default_of_the_runners_vs_oracle(mode):
    # the runners start every solve from u_n + (u_n - u_{n-1}); same 1e-10 bar against the LU oracle
    c = build_case("geballe_with_diamond", 4.0)
    s = make_solver(c, warm=1.0, mode=mode, ordering=ORDERING[mode])
    O = make_oracle(c)
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 0.5 * 6.2e-8, 0.0), (0.951e-6, 0.0), (0.0, 5e-6)])
    hist, iters, fields = s.run(c.amps, c.ic, c.coeff, watch, keep_fields=True)
    ohist, ofields = O.run(c.num_steps, watch, keep_fields=True)
    assert np.abs(hist / ohist - 1).max() <= RTOL_FIELD
    assert max(np.abs(f / of - 1).max() for f, of in zip(fields, ofields)) <= RTOL_FIELD
    s.close()


def test_warm_start_same_answer(small_wd):
    c = small_wd
    a = make_solver(c, warm=0.0)
    b = make_solver(c, warm=1.0)
    ha, ia, _ = a.run(c.amps[:40], c.ic, c.coeff, [3])
    hb, ib, _ = b.run(c.amps[:40], c.ic, c.coeff, [3])
    assert rel_err(a.get_state(), b.get_state()) <= 1e-11
    a.close()
    b.close()


def test_gradient_projection(small_nd):
    c = small_nd
    s = make_solver(c)
    O = make_oracle(c)
    s.run(c.amps[:25], c.ic, c.coeff, [])
    O.run(25)
    P = ho.GradientProjector(c.nodes, c.tris)
    want = P.project(s.get_state())
    got = s.project_gradient()
    scale = np.abs(want).max(axis=0)
    assert np.all(np.abs(got - want).max(axis=0) <= RTOL_FIELD * scale)      # same input field on both sides: the 1e-10 bar
    s.close()


def test_error_paths(small_nd):
    from heatflow_b200._lib import HeatflowError
    from heatflow_b200.solver import HeatSolver
    c = small_nd
    s = HeatSolver(0)
    with pytest.raises(HeatflowError):
        s.step(amp=300.0, t_ic=300.0, coeff=-1.0)          # no mesh / operator yet
    bad = c.tris.copy()
    bad[0, 0] = len(c.nodes) + 5
    with pytest.raises(HeatflowError):
        s.set_mesh(c.nodes, bad, c.cell_tag)
    s.set_mesh(c.nodes, c.tris, c.cell_tag)
    s.set_materials(c.tags[:-1], c.kappa_t[:-1], c.rhoc_t[:-1])  # one tag without material
    s.set_bcs(c.bc_dofs, c.bc_value, c.gauss_slot, c.gauss_r)
    with pytest.raises(HeatflowError):
        s.build_operator(c.dt, True)
    with pytest.raises(HeatflowError):
        s.set_bcs(c.bc_dofs[::-1].copy(), c.bc_value, c.gauss_slot, c.gauss_r)  # unsorted
    s.close()


# ---- initial guess recycled from the previous solves (hf_set_recycle, the runners' default) ----
@MODES
@pytest.mark.parametrize("cap", [128, 6], ids=["full-history", "frozen-when-full"])
def test_recycled_initial_guess_vs_oracle_every_step(mode, cap):
    c = build_case("geballe_with_diamond", 4.0)
    s = make_solver(c, warm=1.0, mode=mode, recycle=cap, ordering=ORDERING[mode])
    O = make_oracle(c)
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 0.5 * 6.2e-8, 0.0), (0.951e-6, 0.0), (0.0, 5e-6)])
    hist, iters, fields = s.run(c.amps, c.ic, c.coeff, watch, keep_fields=True)
    ohist, ofields = O.run(c.num_steps, watch, keep_fields=True)
    assert np.abs(hist / ohist - 1).max() <= RTOL_FIELD
    assert max(np.abs(f / of - 1).max() for f, of in zip(fields, ofields)) <= RTOL_FIELD
    s.close()


def test_recycle_cuts_iterations_and_is_bit_reproducible(small_wd):
    c = small_wd
    base = make_solver(c, warm=1.0)
    _, i0, _ = base.run(c.amps[:60], c.ic, c.coeff, [3])
    base.close()
    runs = []
    for _ in range(2):
        s = make_solver(c, warm=1.0, recycle=64)
        h, it, _ = s.run(c.amps[:60], c.ic, c.coeff, [3, 50])
        runs.append((h, it, s.get_state()))
        # a new initial state drops the basis: the second run of the same solver repeats the first
        s.set_state(np.full(len(c.nodes), c.ic))
        h2, it2, _ = s.run(c.amps[:60], c.ic, c.coeff, [3, 50])
        assert np.array_equal(h, h2) and np.array_equal(it, it2)
        s.close()
    assert np.array_equal(runs[0][0], runs[1][0]) and np.array_equal(runs[0][1], runs[1][1])
    assert np.array_equal(runs[0][2], runs[1][2])
    assert runs[0][1].sum() < 0.5 * i0.sum()


def test_recycle_single_steps_and_argument_errors(small_nd):
    c = small_nd
    s = make_solver(c, recycle=16)
    for k in range(12):
        s.step(c.amps[k + 10], c.ic, c.coeff)
    s2 = make_solver(c)
    for k in range(12):
        s2.step(c.amps[k + 10], c.ic, c.coeff)
    assert rel_err(s.get_state(), s2.get_state()) <= 1e-11
    with pytest.raises(_lib.HeatflowError):
        s.set_recycle(-1)
    s.set_recycle(0)            # switching it off mid-run is allowed
    s.step(c.amps[30], c.ic, c.coeff)
    s2.step(c.amps[30], c.ic, c.coeff)
    assert rel_err(s.get_state(), s2.get_state()) <= 1e-11
    s.close()
    s2.close()


@pytest.mark.parametrize("scale", [0.8, 0.7], ids=["2.2e5-dofs-1536-row-patches", "2.8e5-dofs-2048-row-patches"])
def test_mid_size_mesh_runs_on_chip_with_the_patch_kernel(scale):
    # 2.2e5 / 2.8e5 dofs (the size of the reference's own gmsh meshes): more than 1024 rows per SM, auto mode picks the
    # pipelined on-chip kernel with 6 / 8 rows per thread; runner defaults (warm start + recycled guess)
    c = build_case("geballe_with_diamond", scale)
    s = make_solver(c, warm=1.0, recycle=32)
    assert s.solver_path() == 4                        # pipelined on-chip kernel
    O = make_oracle(c)
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 0.5 * 6.2e-8, 0.0), (0.951e-6, 0.0), (0.0, 5e-6)])
    S = 16
    hist, iters, _ = s.run(c.amps[:S], c.ic, c.coeff, watch)
    ohist, _ = O.run(S, watch)
    assert np.abs(hist / ohist - 1).max() <= RTOL_FIELD
    assert rel_err(s.get_state(), O.u) <= RTOL_FIELD
    l0 = s.stats()["launches"]
    s.run(c.amps[S:S + 2], c.ic, c.coeff, watch)
    assert s.stats()["launches"] - l0 < 40              # one solver launch per step, not one per iteration
    s.close()


@pytest.fixture(scope="module")
def konopkova_1m():
    return build_case("konopkova", 0.35)


def test_headline_config_full_size_all_steps_runner_defaults():
    # BASELINE config #3 exactly as bench.py and run_with_diamond run it: cfg mesh sizes (1.4e5 dofs), all 100 steps,
    # runner defaults (auto kernel = on-chip patch kernel, warm start, recycled initial guess of 128 vectors);
    # fields at EVERY step against the LU oracle
    c = build_case("geballe_with_diamond", 1.0)
    s = make_solver(c, warm=1.0, recycle=128)
    assert s.on_chip()
    O = make_oracle(c)
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 0.5 * 6.2e-8, 0.0), (0.951e-6, 0.0), (0.0, 5e-6)])
    hist, iters, fields = s.run(c.amps, c.ic, c.coeff, watch, keep_fields=True)
    assert s.stats()["retries"] == 0
    worst = 0.0
    for k in range(c.num_steps):
        uo = O.step((k + 1) * c.dt)
        worst = max(worst, np.abs(fields[k] / uo - 1).max())
        assert np.abs(hist[k] / uo[watch] - 1).max() <= RTOL_FIELD, k
    assert worst <= RTOL_FIELD, worst
    s.close()


def test_konopkova_1m_dofs_vs_lu_oracle_every_step(konopkova_1m):
    # BASELINE config #4 at the size the roofline claim is made on (1.16 M dofs): the persistent streaming kernel with
    # the runner defaults against scipy splu (one factorisation, ~1 min on a host core), fields at every step;
    # then the host-polled streaming kernel on the last step from the same state
    c = konopkova_1m
    n = len(c.nodes)
    assert n > 1_000_000
    O = make_oracle(c)
    O.factorize()
    s = make_solver(c, warm=1.0, recycle=8)
    assert s.solver_path() == 2
    S = 8
    first = int(np.flatnonzero(c.amps != c.ic)[0])          # heated steps: the leading amp == ic ones leave u == ic
    hist, iters, fields = s.run(c.amps[first:first + S], c.ic, c.coeff, [0, n // 2], keep_fields=True)
    assert iters.min() > 50
    u_before_last = fields[S - 2].copy()
    for k in range(S):
        uo = O.step((first + k + 1) * c.dt)
        assert np.abs(fields[k] / uo - 1).max() <= RTOL_FIELD, k
    assert np.abs(uo - c.ic).max() > 10.0                     # a heated state, not the trivial one
    s.set_solver(rtol=1e-14, warm=0.0, mode=1)
    s.set_recycle(0)
    s.set_state(u_before_last)
    s.step(c.amps[first + S - 1], c.ic, c.coeff)
    assert np.abs(s.get_state() / uo - 1).max() <= RTOL_FIELD
    s.close()


def test_large_mesh_size_independent_properties(konopkova_1m):
    # BASELINE config #4 (konopkova refined to 1.16 M dofs, streaming kernels): properties that hold at any size -
    # constant states are fixed points, the update is linear in the heating amplitude, the operator is symmetric,
    # and the recycled initial guess does not change the answer
    c = konopkova_1m
    n = len(c.nodes)
    assert n > 1_000_000
    s = make_solver(c, warm=1.0)
    assert s.solver_path() == 2
    u0 = np.full(n, c.ic)
    # (a) amplitude == initial temperature: nothing may move
    s.set_state(u0)
    for _ in range(2):
        s.step(c.ic, c.ic, c.coeff)
    assert np.abs(s.get_state() / c.ic - 1).max() <= 1e-13
    # (b) linearity: (u(a1) - T0) * (a2 - T0) == (u(a2) - T0) * (a1 - T0) after three steps
    outs = []
    for amp in (c.ic + 400.0, c.ic + 1000.0):
        s.set_state(u0)
        for k in range(3):
            s.step(c.ic + (amp - c.ic) * (k + 1) / 3.0, c.ic, c.coeff)
        outs.append(s.get_state() - c.ic)
    scale = np.abs(outs[1]).max()
    assert scale > 100.0
    assert np.abs(outs[0] * 2.5 - outs[1]).max() <= 1e-9 * scale
    # (c) symmetry of the Dirichlet-treated operator through the production SpMV kernel
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    ax, ay = s.spmv(x), s.spmv(y)
    assert abs(y @ ax - x @ ay) <= 1e-12 * (np.abs(y) @ np.abs(ax))
    # (d) the recycled initial guess (runner default) gives the same temperatures with fewer iterations
    s.set_state(u0)
    h0, it0, _ = s.run(c.amps[:8], c.ic, c.coeff, [0, n // 3, n // 2])
    u_plain = s.get_state()
    s.set_recycle(8)
    s.set_state(u0)
    h1, it1, _ = s.run(c.amps[:8], c.ic, c.coeff, [0, n // 3, n // 2])
    # two solves that each stop at ||r|| <= 1e-14 ||b|| agree to ~1e-11 (measured 1.04e-11 at 1.15 M dofs); the 1e-10
    # bar against the LU oracle is test_konopkova_1m_dofs_vs_lu_oracle_every_step
    assert np.abs(h1 / h0 - 1).max() <= 5e-11 and np.abs(s.get_state() / u_plain - 1).max() <= 5e-11
    assert it1.sum() < it0.sum()
    s.close()


def test_failed_on_chip_solve_is_repeated_with_the_streaming_kernel(small_wd):
    # A partial sum outside the fixed-point range of the on-chip reduction (forced here: the range is shrunk by
    # 2^-60) poisons the solve; the kernel counts the failure and hf_run repeats the run from its initial state with
    # the host-polled streaming kernel - still on the GPU, same answer, and the repeat is visible in the counters
    c = small_wd
    s = make_solver(c, warm=1.0, recycle=16, mode=3, ordering="hilbert")
    O = make_oracle(c)
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 0.5 * 6.2e-8, 0.0), (0.9e-6, 0.0)])
    _lib.check(s._L.hf_debug_fx_shift(s._h, 60))
    hist, iters, fields = s.run(c.amps[:20], c.ic, c.coeff, watch, keep_fields=True)
    assert s.stats()["retries"] == 1
    ohist, ofields = O.run(20, watch, keep_fields=True)
    assert max(np.abs(f / of - 1).max() for f, of in zip(fields, ofields)) <= RTOL_FIELD
    assert np.abs(hist / ohist - 1).max() <= RTOL_FIELD and iters.max() > 0
    # back to the normal range: the next run stays on the on-chip kernel
    _lib.check(s._L.hf_debug_fx_shift(s._h, 0))
    s.run(c.amps[20:30], c.ic, c.coeff, watch)
    for k in range(20, 30):
        O.step((k + 1) * c.dt)
    assert s.stats()["retries"] == 1 and rel_err(s.get_state(), O.u) <= RTOL_FIELD
    s.close()


def test_edge_cases_empty_runs_tiny_mesh_and_iteration_cap(small_nd):
    from heatflow_b200._lib import HeatflowError
    from heatflow_b200.solver import HeatSolver
    c = small_nd
    s = make_solver(c)
    # empty run / no watchers: nothing moves, shapes are right
    u_before = s.get_state()
    hist, iters, _ = s.run(c.amps[:0], c.ic, c.coeff, [0])
    assert hist.shape == (0, 1) and iters.shape == (0,) and np.array_equal(s.get_state(), u_before)
    hist, iters, _ = s.run(c.amps[10:13], c.ic, c.coeff, [])
    assert hist.shape == (3, 0) and np.all(iters > 0)
    # iteration cap: reported as an error, never a silent wrong answer
    s.set_solver(rtol=1e-14, max_iters=5)
    with pytest.raises(HeatflowError, match="iteration|converge"):
        s.run(c.amps[20:22], c.ic, c.coeff, [0])
    s.close()
    # a two-triangle mesh with constant Dirichlet data only (no Gaussian dofs): one patch, one CTA
    nodes = np.array([[0.0, 0.0], [1e-6, 0.0], [1e-6, 1e-6], [0.0, 1e-6]])
    tris = np.array([[0, 1, 2], [0, 2, 3]], dtype=np.int32)
    t = HeatSolver(0)
    t.set_mesh(nodes, tris, np.array([1, 1], dtype=np.int32))
    t.set_materials([1], [10.0], [3.0e6])
    t.set_bcs([0, 3], [300.0, 300.0])
    t.build_operator(1e-7, True)
    t.set_solver(rtol=1e-14)
    t.set_recycle(4)
    t.set_state(np.array([300.0, 500.0, 500.0, 300.0]))
    O = ho.Oracle2D(nodes, tris, np.full(2, 3.0e6), np.full(2, 10.0), 1e-7, [(np.array([0, 3]), "const")], 300.0, 1e-6,
                    np.array([0.0, 1.0]), np.array([300.0, 300.0]))
    O.u = np.array([300.0, 500.0, 500.0, 300.0])
    for k in range(6):
        t.step(None)
        ou = O.step((k + 1) * 1e-7)
        assert np.abs(t.get_state() / ou - 1).max() <= RTOL_FIELD
    assert np.all(t.get_state()[[1, 2]] < 500.0)
    t.close()
