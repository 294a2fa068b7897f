"""The simulation layer of the reference on the B200 engine — src/simulation/__init__.py:
``simulate_terminal_prices(S, T, r, sigma, q, n_paths, n_steps, seed) -> ndarray`` in its five flavours
(gbm_numpy.py:15-83, gbm_numba.py:100-129, gbm_qmc.py:14-76).  Same signatures, same array layout: ``[0, N)`` are the
paths on +Z and ``[N, 2N)`` their mirrors on -Z (``np.concatenate([exp(+), exp(-)])``, gbm_numpy.py:51).

The draws come from the engine's Philox stream (stream 0 of ``seed``) or, for the QMC flavours, from scipy's own
scrambled Sobol points regenerated on the device — the very stream ``MonteCarloPricer.price`` uses, so
``payoff(simulate_gbm_numpy(...))`` reproduces the pricer's moments.  Values are FP64 copies of the fused path's FP32
arithmetic.  ``simulate_gbm_paths`` (the full ``(n_paths, n_steps + 1)`` array, unused by the reference itself) is not
offered: not writing path arrays to HBM is the point of this engine.  These calls are per process (no sharding).
"""

from __future__ import annotations

import numpy as np

from . import _ffi, sobol
from .exceptions import MonteCarloError

__all__ = ["simulate_gbm_numpy", "simulate_gbm_numpy_fast", "simulate_gbm_numba", "simulate_gbm_qmc", "simulate_gbm_qmc_antithetic",
           "NUMBA_AVAILABLE"]

NUMBA_AVAILABLE = True  # the name callers test before choosing simulate_gbm_numba (gbm_numba.py:16-20); one engine serves all


def _philox(S, T, r, sigma, q, n_paths, n_steps, seed, antithetic):
    return _ffi.get_engine().terminal_prices(_ffi.make_params(S, 0.0, T, r, sigma, q), int(n_steps), int(seed), int(n_paths),
                                             antithetic=antithetic)


def simulate_gbm_numpy(S: float, T: float, r: float, sigma: float, q: float, n_paths: int, n_steps: int, seed: int,
                       antithetic: bool = True) -> np.ndarray:
    """gbm_numpy.py:15-53."""
    return _philox(S, T, r, sigma, q, n_paths, n_steps, seed, antithetic)


def simulate_gbm_numpy_fast(S: float, T: float, r: float, sigma: float, q: float, n_paths: int, seed: int) -> np.ndarray:
    """gbm_numpy.py:56-83: one exact step, mirrored."""
    return _philox(S, T, r, sigma, q, n_paths, 1, seed, True)


def simulate_gbm_numba(S: float, T: float, r: float, sigma: float, q: float, n_paths: int, n_steps: int, seed: int,
                       parallel: bool = True) -> np.ndarray:
    """gbm_numba.py:100-129 (``parallel`` accepted and ignored): 2 * n_paths terminal prices."""
    return _philox(S, T, r, sigma, q, n_paths, n_steps, seed, True)


def _qmc(S, T, r, sigma, q, n_paths, n_steps, seed, antithetic):
    steps = min(int(n_steps), 21201)  # gbm_qmc.py:30
    table, shift, bits = sobol.sobol_table(steps, seed)
    if n_paths > (1 << bits):
        raise MonteCarloError(f"a {bits}-bit Sobol sequence has 2^{bits} points; asked for {n_paths}")
    return _ffi.get_engine().terminal_prices_sobol(_ffi.make_params(S, 0.0, T, r, sigma, q), steps, table, shift, bits, int(n_paths),
                                                   antithetic=antithetic)


def simulate_gbm_qmc(S: float, T: float, r: float, sigma: float, q: float, n_paths: int, n_steps: int, seed: int) -> np.ndarray:
    """gbm_qmc.py:14-46: n_paths scrambled-Sobol terminal prices."""
    return _qmc(S, T, r, sigma, q, n_paths, n_steps, seed, False)


def simulate_gbm_qmc_antithetic(S: float, T: float, r: float, sigma: float, q: float, n_paths: int, n_steps: int, seed: int) -> np.ndarray:
    """gbm_qmc.py:49-76: the same points and their mirrors, 2 * n_paths terminal prices."""
    return _qmc(S, T, r, sigma, q, n_paths, n_steps, seed, True)
