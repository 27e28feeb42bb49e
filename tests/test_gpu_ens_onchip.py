"""GPU parity tests of the batched ON-CHIP ensemble kernel (hf_enspatch.cu: one cooperative launch per time step for a
tile of up to 4 sweep variants) against the scipy oracle, the streaming ensemble kernels and itself."""
import numpy as np
import pytest

from heatflow_b200 import _lib, problem
from helpers import build_case, make_solver
from oracle import heat_oracle as ho
from test_gpu_ensemble import oracle_variant, sample_tag

pytestmark = pytest.mark.gpu
RTOL_FIELD = 1e-10


@pytest.fixture(scope="module")
def wd():
    return build_case("geballe_with_diamond", 8.0)


def run_tile(c, ks, fw, S, watch, recycle=0, warm=0.0, first=0):
    s = make_solver(c, warm=warm, ordering="hilbert", recycle=recycle)
    s.ens_create(ks, [problem.gaussian_coeff(f) for f in fw], sample_tag(c))
    path = s.ens_path()
    hist, iters = s.ens_run(c.amps[first:first + S], c.ic, watch)
    u = s.ens_get_state()
    st = s.stats()
    s.ens_destroy()
    s.close()
    return path, hist, iters, u, st


@pytest.mark.parametrize("nh", ["1", "2"])
@pytest.mark.parametrize("ks,fw", [
    ([1.0, 10.0, 100.0], [1e-6, 1.3e-5, 1e-4]),                # three conductivities: S0 entries from shared memory, padded tile
    ([7.0, 7.0, 7.0, 7.0], [1e-6, 5e-6, 2e-5, 1e-4]),          # one conductivity (a sweep sorted by k): folded operator
    ([2.0, 2.0, 50.0, 50.0], [1e-4, 1e-6, 1e-4, 1e-6]),        # the two half-tiles converge at different iterations
    ([30.0], [1.3e-5]),
])
def test_on_chip_tile_matches_oracle_per_variant(wd, monkeypatch, nh, ks, fw):
    monkeypatch.setenv("HF_ENS_NH", nh)
    c = wd
    S = 30
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 3e-8, 0.0), (0.95e-6, 0.0), (0.0, 5e-6)])
    path, hist, iters, u, st = run_tile(c, ks, fw, S, watch)
    assert path == 5 and st["retries"] == 0
    assert iters.shape == (S,) and np.all(iters[5:] > 0)
    for b, (k, f) in enumerate(zip(ks, fw)):
        O = oracle_variant(c, k, f)
        ohist, _ = O.run(S, watch)
        assert np.abs(hist[b] / ohist - 1).max() <= RTOL_FIELD, (b, k, f)
        assert np.abs(u[b] / O.u - 1).max() <= RTOL_FIELD, (b, k, f)


def test_on_chip_tile_with_runner_defaults_is_reproducible_and_agrees_with_streaming(wd, monkeypatch):
    # warm start + per-variant recycled bases (the sweep's defaults); bit-identical when repeated; the streaming
    # ensemble kernels (HF_ENS_STREAM) give the same temperatures to solver tolerance with one launch per iteration
    c = wd
    S = 40
    ks, fw = [3.0, 3.0, 40.0, 40.0], [2e-6, 3e-5, 2e-6, 3e-5]
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 3e-8, 0.0), (0.95e-6, 0.0)])
    p1, h1, it1, u1, st1 = run_tile(c, ks, fw, S, watch, recycle=64, warm=1.0)
    p2, h2, it2, u2, _ = run_tile(c, ks, fw, S, watch, recycle=64, warm=1.0)
    assert p1 == 5 and p2 == 5
    assert np.array_equal(h1, h2) and np.array_equal(u1, u2) and np.array_equal(it1, it2)
    assert st1["launches"] < 40 * S                                # one solver launch per step, not one per iteration
    monkeypatch.setenv("HF_ENS_STREAM", "1")
    p3, h3, it3, u3, st3 = run_tile(c, ks, fw, S, watch, recycle=64, warm=1.0)
    assert p3 == 1
    assert np.abs(h3 / h1 - 1).max() <= 2e-11 and np.abs(u3 / u1 - 1).max() <= 2e-11
    for b in (0, 3):
        O = oracle_variant(c, ks[b], fw[b])
        ohist, _ = O.run(S, watch)
        assert np.abs(h1[b] / ohist - 1).max() <= RTOL_FIELD
    # without the recycled basis the same tile needs more iterations
    _, _, it0, _, _ = run_tile(c, ks, fw, S, watch, recycle=0, warm=1.0)
    assert it1.sum() < 0.6 * it0.sum()


@pytest.mark.parametrize("recycle,warm", [(0, 0.0), (64, 1.0)])
def test_failed_on_chip_tile_is_repeated_with_the_streaming_kernels(wd, recycle, warm):
    c = wd
    S = 12
    ks, fw = [5.0, 5.0, 5.0, 5.0], [1e-6, 4e-6, 2e-5, 1e-4]
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 3e-8, 0.0)])
    s = make_solver(c, ordering="hilbert", recycle=recycle, warm=warm)
    s.ens_create(ks, [problem.gaussian_coeff(f) for f in fw], sample_tag(c))
    assert s.ens_path() == 5
    _lib.check(s._L.hf_debug_fx_shift(s._h, 60))                    # partial sums leave the fixed-point range
    hist, iters = s.ens_run(c.amps[10:10 + S], c.ic, watch)
    assert s.stats()["retries"] == 1 and np.all(iters > 0)
    _lib.check(s._L.hf_debug_fx_shift(s._h, 0))
    s.ens_destroy()
    s.close()
    # same steps on a clean context (on-chip, no failure): the repeated run gives the same temperatures
    _, h_ok, _, _, st = run_tile(c, ks, fw, S, watch, first=10, recycle=recycle, warm=warm)
    assert st["retries"] == 0
    assert np.abs(hist / h_ok - 1).max() <= 2e-11


def test_config5_tile_at_full_size_matches_oracle():
    # BASELINE config #5 at the cfg's own mesh (1.4e5 dofs, 138 CTAs x 1024 rows): tiles of four sweep variants with
    # one conductivity each (the sweep sorts by k: 64 heating widths per k), sweep defaults (warm start, recycled
    # bases), all 100 steps, watcher histories and final fields of the corner variants against the LU oracle
    c = build_case("geballe_with_diamond", 1.0)
    fw = [1e-6, 1e-5, 3e-5, 1e-4]
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 0.5 * 6.2e-8, 0.0), (0.951e-6, 0.0)])
    for k in (1.0, 100.0):
        path, hist, iters, u, st = run_tile(c, [k] * 4, fw, c.num_steps, watch, recycle=128, warm=1.0)
        assert path == 5 and st["retries"] == 0
        for b in (0, 3):
            O = oracle_variant(c, k, fw[b])
            ohist, _ = O.run(c.num_steps, watch)
            assert np.abs(hist[b] / ohist - 1).max() <= RTOL_FIELD, (k, b)
            assert np.abs(u[b] / O.u - 1).max() <= RTOL_FIELD, (k, b)
