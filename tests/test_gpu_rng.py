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
