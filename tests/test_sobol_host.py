"""CPU: host side of the quasi-Monte Carlo backend (optionslab_b200/sobol.py) — the natural-order direction
table must regenerate scipy's scrambled Sobol points bit for bit, and rank slices must tile the sequence."""

import warnings

import numpy as np
import pytest

from optionslab_b200 import sobol


@pytest.mark.parametrize("d,seed,n", [(1, 0, 257), (7, 42, 5000), (252, 1, 3000), (365, 123456789, 1030)])
def test_natural_order_table_reproduces_scipy_points(d, seed, n):
    from scipy.stats.qmc import Sobol

    table, shift, bits = sobol.sobol_table(d, seed)
    assert table.shape == (d, sobol.WORDS) and table.dtype == np.uint32 and shift.shape == (d,)
    assert not table[:, bits:].any()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = Sobol(d=d, scramble=True, seed=seed).random(n)
    x = sobol.points_from_table(table, shift, 0, n)
    np.testing.assert_array_equal(x.astype(np.float64) * 2.0**-bits, ref)
    # any sub-range of the sequence is addressable directly (what a rank does)
    np.testing.assert_array_equal(sobol.points_from_table(table, shift, 200, n - 200), x[200:])


def test_gray_to_natural_is_the_gray_code_identity():
    rng = np.random.default_rng(0)
    sv = rng.integers(0, 2**30, size=(3, 30), dtype=np.uint64)
    table = sobol.gray_to_natural(sv, 30)
    for i in (0, 1, 2, 3, 12345, 2**29 + 17):
        gray = i ^ (i >> 1)
        want = np.zeros(3, dtype=np.uint64)
        nat = np.zeros(3, dtype=np.uint64)
        for b in range(30):
            if (gray >> b) & 1:
                want ^= sv[:, b]
            if (i >> b) & 1:
                nat ^= table[:, b].astype(np.uint64)
        np.testing.assert_array_equal(nat, want)


@pytest.mark.parametrize("n,world", [(1, 1), (4096, 2), (5000, 4), (100_000, 3), (1 << 20, 8), (12345, 5)])
def test_partition_points_tiles_the_sequence_on_cta_boundaries(n, world):
    parts = [sobol.partition_points(n, r, world) for r in range(world)]
    pos = 0
    for begin, count in parts:
        assert count >= 0
        if count:
            assert begin == pos and begin % sobol.POINT_ALIGNMENT == 0
            pos += count
    assert pos == n
    blocks = [-(-c // sobol.POINT_ALIGNMENT) for _, c in parts]
    assert max(blocks) - min(blocks) <= 1


def test_dimension_limits():
    from optionslab_b200.exceptions import MonteCarloError

    with pytest.raises(MonteCarloError):
        sobol.sobol_table(0, 1)
    with pytest.raises(MonteCarloError):
        sobol.sobol_table(sobol.MAX_DIMS + 1, 1)
