"""GPU: the device random stream — Philox known answers and the documented normal mapping."""

import numpy as np
import pytest

from oracle import philox_oracle as po
from tests import kat

pytestmark = pytest.mark.gpu


def test_device_philox_known_answers(engine):
    np.testing.assert_array_equal(engine.philox_raw(kat.kat_inputs()), kat.kat_outputs())


def test_device_philox_matches_c_oracle_on_random_counters(engine):
    rng = np.random.default_rng(1)
    ck = rng.integers(0, 2**32, size=(4096, 6), dtype=np.uint64).astype(np.uint32)
    got = engine.philox_raw(ck)
    for i in range(0, len(ck), 61):
        np.testing.assert_array_equal(got[i], po.philox4x32_10(ck[i, :4], ck[i, 4:]))


@pytest.mark.parametrize("n_paths,n_steps,stream,begin", [(1000, 8, 0, 0), (333, 7, 5, 12345), (64, 1, 0, 0),
                                                           (17, 365, 2, (1 << 32) - 5), (5, 3, 4095, 1 << 40)])
def test_device_normals_follow_the_documented_stream(engine, n_paths, n_steps, stream, begin):
    """FP32 MUFU normals vs the FP64 libm evaluation of the same mapping (oracle/philox_oracle.c)."""
    seed = 0x1234_5678_9ABC_DEF0
    got = engine.generate_normals(seed, n_paths, n_steps, stream=stream, path_begin=begin)
    want = po.normals(seed, n_paths, n_steps, stream=stream, path_begin=begin)
    # lg2/sqrt/sin/cos approximations: ~1e-6 relative on the radius, ~1e-6 absolute on the angle
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-5)


def test_device_normal_moments(engine):
    z = engine.generate_normals(7, 1_000_000, 8).astype(np.float64).ravel()
    n = z.size
    assert abs(z.mean()) < 4 / np.sqrt(n)
    assert abs(z.var() - 1) < 4 * np.sqrt(2 / n)
    assert abs((z**4).mean() - 3) < 4 * np.sqrt(96 / n)
    assert np.abs(z).max() < 5.7


def test_device_normals_ks_and_pair_independence(engine):
    """8e6 device draws: Kolmogorov-Smirnov against the normal CDF (a fixed 512-angle grid would put an atom of
    mass 1/256 at z = 0 and fail), and the two normals of one Box-Muller pair have uncorrelated squares."""
    from scipy import stats

    zz = engine.generate_normals(2025, 1_000_000, 8).astype(np.float64)
    assert stats.kstest(zz.ravel(), "norm").pvalue > 1e-3
    n = len(zz)
    assert abs(np.corrcoef(zz[:, 0] ** 2, zz[:, 1] ** 2)[0, 1]) < 4 / np.sqrt(n)
    assert abs(np.corrcoef(zz[:, 1], zz[:, 2])[0, 1]) < 4 / np.sqrt(n)
    assert abs(np.corrcoef(zz[:-1, 7], zz[1:, 0])[0, 1]) < 4 / np.sqrt(n)


def test_streams_and_neighbouring_paths_are_uncorrelated(engine):
    """Disjoint Philox counters must give independent draws: the same (path, step) cells of streams s and s+1 (two
    options of a batch), of path blocks one apart (two GPUs of a sharded run) and of seeds differing in one bit are
    uncorrelated, in the values and in their squares; lagged steps within a path likewise (lags 1..9 cross the
    cosine/sine branch of a pair, the 4 pairs of a Philox call and the call boundary)."""
    n_paths, n_steps, seed = 1 << 18, 32, 99
    n = n_paths * n_steps
    lim = 4.5 / np.sqrt(n)
    base = engine.generate_normals(seed, n_paths, n_steps, stream=3).astype(np.float64)
    others = {
        "next stream": engine.generate_normals(seed, n_paths, n_steps, stream=4),
        "next path block": engine.generate_normals(seed, n_paths, n_steps, stream=3, path_begin=n_paths),
        "seed ^ 1": engine.generate_normals(seed ^ 1, n_paths, n_steps, stream=3),
        "seed ^ 2^32": engine.generate_normals(seed ^ (1 << 32), n_paths, n_steps, stream=3),
    }
    a = base.ravel()
    for name, other in others.items():
        b = other.astype(np.float64).ravel()
        assert abs(np.mean(a * b)) < lim, name
        assert abs(np.mean((a * a - 1) * (b * b - 1))) < 2 * lim, name  # var of (z^2-1)(z'^2-1) is 4
        assert np.max(np.abs(a - b)) > 1.0, name
    for lag in range(1, 10):
        x, y = base[:, :-lag].ravel(), base[:, lag:].ravel()
        assert abs(np.mean(x * y)) < 4.5 / np.sqrt(x.size), lag
        assert abs(np.mean((x * x - 1) * (y * y - 1))) < 9.0 / np.sqrt(x.size), lag
    # neighbouring paths of one stream (adjacent threads of a warp)
    x, y = base[:-1].ravel(), base[1:].ravel()
    assert abs(np.mean(x * y)) < 4.5 / np.sqrt(x.size)
