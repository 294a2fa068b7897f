"""Host side of the quasi-Monte Carlo backend: obtain the point set the reference uses and hand
its GF(2)-linear description to the device.

The reference draws ``scipy.stats.qmc.Sobol(d=n_steps, scramble=True, seed=seed).random(N)``
(src/simulation/gbm_qmc.py:32-33).  SciPy generates point i as
``shift ^ XOR_{b in bits(gray(i))} sv[:, b]`` scaled by 2^-bits, with ``sv`` the (linear-matrix-
scrambled) direction numbers and ``shift`` the digital shift.  Because gray(i) = i ^ (i >> 1) is itself
linear over GF(2), the same point is ``shift ^ XOR_{b in bits(i)} (sv[:, b] ^ sv[:, b-1])`` — a table in
the natural order of the point index, which is what ``b200mc_simulate_sobol`` consumes.  The direction
numbers (Joe & Kuo) and the scrambling are SciPy's, read from the constructed sampler; the points the
device generates are bit-identical to ``sampler.random`` (tests/test_gpu_qmc.py).
"""

from __future__ import annotations

from typing import Tuple

import numpy as np

from .exceptions import MonteCarloError

WORDS = 32              # table words per dimension (include/b200mc.h)
POINT_ALIGNMENT = 4096  # ranks split the sequence at multiples of one CTA of points
MAX_DIMS = 21201        # scipy's table of direction numbers (gbm_qmc.py:30)


_TABLE_CACHE: "dict[tuple, tuple]" = {}


def sobol_table(n_dims: int, seed) -> Tuple[np.ndarray, np.ndarray, int]:
    """-> (dirnums [n_dims, 32] uint32 in natural order, shift [n_dims] uint32, bits) for the sampler
    ``Sobol(d=n_dims, scramble=True, seed=seed)``.  Integer seeds are cached (the reference rebuilds the sampler on
    every price call; a bump-and-revalue sweep asks for the same table 8-14 times)."""
    key = (int(n_dims), int(seed)) if isinstance(seed, (int, np.integer)) else None
    if key is not None and key in _TABLE_CACHE:
        return _TABLE_CACHE[key]
    out = _build_table(n_dims, seed)
    if key is not None:
        if len(_TABLE_CACHE) >= 8:
            _TABLE_CACHE.pop(next(iter(_TABLE_CACHE)))
        _TABLE_CACHE[key] = out
    return out


def _build_table(n_dims: int, seed) -> Tuple[np.ndarray, np.ndarray, int]:
    from scipy.stats.qmc import Sobol

    if not (1 <= n_dims <= MAX_DIMS):
        raise MonteCarloError(f"Sobol dimension must be in [1, {MAX_DIMS}]")
    sampler = Sobol(d=int(n_dims), scramble=True, seed=seed)
    try:
        sv = np.asarray(sampler._sv, dtype=np.uint64)
        shift = np.asarray(sampler._shift, dtype=np.uint64)
        bits = int(sampler.bits)
    except AttributeError as exc:  # pragma: no cover - guards against a SciPy that renames its internals
        raise MonteCarloError("this SciPy does not expose the Sobol direction numbers (Sobol._sv/_shift)") from exc
    if sv.shape != (n_dims, bits) or bits > 31:
        raise MonteCarloError(f"unexpected Sobol table shape {sv.shape} / bits {bits}")
    table, shift32 = gray_to_natural(sv, bits), shift.astype(np.uint32)
    _self_test(table, shift32, bits, n_dims, seed)
    return table, shift32, bits


_VERIFIED_SCIPY: "set[str]" = set()


def _self_test(table: np.ndarray, shift: np.ndarray, bits: int, n_dims: int, seed) -> None:
    """``Sobol._sv`` / ``._shift`` are private: the first table built under a given SciPy version is checked against the
    public API - the first 64 points regenerated from the table must equal ``Sobol(...).random(64)`` bit for bit - so
    that a SciPy that changes its internals makes ``MCMethod.QMC`` fail loudly instead of pricing another point set."""
    import scipy

    if scipy.__version__ in _VERIFIED_SCIPY or not isinstance(seed, (int, np.integer)):
        return  # (a Generator / None seed cannot rebuild the same scramble twice: checked on the next integer seed)
    import warnings

    from scipy.stats.qmc import Sobol

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = Sobol(d=int(n_dims), scramble=True, seed=seed).random(64)
    got = points_from_table(table, shift, 0, 64).astype(np.float64) * 2.0 ** -bits
    if not np.array_equal(got, want):
        raise MonteCarloError(f"SciPy {scipy.__version__}: the Sobol points rebuilt from Sobol._sv/_shift differ from Sobol.random(); "
                              "the QMC backend does not know this SciPy's generator layout")
    _VERIFIED_SCIPY.add(scipy.__version__)


def gray_to_natural(sv: np.ndarray, bits: int) -> np.ndarray:
    """Direction numbers for Gray-code enumeration -> table for binary enumeration: c[:, b] = v[:, b] ^ v[:, b-1]."""
    sv = np.asarray(sv, dtype=np.uint64)
    table = np.zeros((sv.shape[0], WORDS), dtype=np.uint32)
    table[:, 0] = sv[:, 0]
    table[:, 1:bits] = sv[:, 1:bits] ^ sv[:, : bits - 1]
    return table


def points_from_table(table: np.ndarray, shift: np.ndarray, point_begin: int, n_points: int) -> np.ndarray:
    """NumPy evaluation of x_j(i) = shift_j ^ XOR_{b in bits(i)} table[j, b] (host check of the contract; the
    device does the same in qmc_european_kernel).  -> [n_points, n_dims] uint32."""
    idx = np.arange(point_begin, point_begin + n_points, dtype=np.uint64)
    x = np.broadcast_to(shift.astype(np.uint32), (n_points, len(shift))).copy()
    for b in range(WORDS):
        sel = ((idx >> np.uint64(b)) & np.uint64(1)).astype(bool)
        if sel.any():
            x[sel] ^= table[:, b]
    return x


def partition_points(n_points: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous split of [0, n_points) whose boundaries are multiples of POINT_ALIGNMENT (the last rank with
    work takes the ragged tail).  Returns (first point, count) for ``rank``."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    blocks = -(-int(n_points) // POINT_ALIGNMENT)
    base, extra = divmod(blocks, world_size)
    first_block = rank * base + min(rank, extra)
    n_blocks = base + (1 if rank < extra else 0)
    begin = min(first_block * POINT_ALIGNMENT, n_points)
    end = min((first_block + n_blocks) * POINT_ALIGNMENT, n_points)
    return begin, end - begin
