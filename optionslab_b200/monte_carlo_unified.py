"""``MonteCarloPricerUni`` — drop-in for src/pricing_models/monte_carlo_unified.py:236-689.

Constructor, validation, exception types and method signatures follow the reference; the three
CPU/CuPy backends collapse into the one CUDA engine.  Batches run as a single launch over an
(option x path-tile) grid — option ``i`` draws from Philox stream ``i`` (the reference's Numba
backend seeds option ``i`` with ``seed + i``, monte_carlo_unified.py:190) — and
``delta_gamma[_batch]`` price S-h, S, S+h as three scenarios of the SAME draws in that launch.
"""

from __future__ import annotations

import logging
import math
import threading
from typing import Literal, Optional, Tuple, Union

import numpy as np

from . import _ffi, runtime
from .exceptions import InputValidationError, MonteCarloError
from .greeks import compute_greeks_unified

__all__ = ["MonteCarloPricerUni", "InputValidationError", "MonteCarloError", "NUMBA_AVAILABLE", "GPU_AVAILABLE"]

logger = logging.getLogger(__name__)
NUMBA_AVAILABLE = False
GPU_AVAILABLE = True  # the engine *is* the GPU path; it raises AccelerationError if the device is missing


class MonteCarloPricerUni:
    def __init__(self, num_simulations: int = 100_000, num_steps: int = 100, seed: Optional[int] = None,
                 use_numba: bool = True, use_gpu: bool = False) -> None:
        if num_simulations <= 0 or num_steps <= 0:
            raise InputValidationError("num_simulations and num_steps must be positive integers")
        self.num_simulations = num_simulations
        self.num_steps = num_steps
        self.seed = seed if seed is not None else runtime.fresh_seed()
        self.rng = np.random.default_rng(seed)  # only draws per-call seeds in delta_gamma (:544-545)
        self.use_numba = False                   # accepted and ignored: one engine, no backend switch
        self.use_gpu = True
        self._lock = threading.RLock()
        logger.info("MonteCarloPricerUni initialized: simulations=%d, steps=%d, engine=b200mc", num_simulations, num_steps)

    # -- internals ----------------------------------------------------------------------------
    def _batch_moments(self, S, K, T, r, sigma, q, option_type, seed, bumps=(0.0,)):
        """moments [n_opt, len(bumps)]: scenario j prices S + bumps[j] on the option's shared draws."""
        S = np.asarray(S, dtype=np.float64)
        cols = [_ffi.make_params(S + b, K, T, r, sigma, q) for b in bumps]
        params = np.stack(cols, axis=1)
        spec = _ffi.make_spec(_ffi.EUROPEAN, self.num_steps, is_put=(option_type == "put"), antithetic=True)
        actual_seed = seed if seed is not None else self.seed
        with self._lock:
            return runtime.simulate(spec, params, actual_seed, self.num_simulations, stream_base=0)

    @staticmethod
    def _validate(S, K, T, sigma, option_type):
        if S <= 0 or K <= 0 or T <= 0 or sigma < 0:
            raise InputValidationError("S, K, T must be positive; sigma must be non-negative")
        if option_type not in {"call", "put"}:
            raise InputValidationError("option_type must be 'call' or 'put'")

    # -- reference surface --------------------------------------------------------------------
    def price(self, S: float, K: float, T: float, r: float, sigma: float, option_type: Literal["call", "put"],
              q: float = 0.0, seed: Optional[int] = None) -> float:
        self._validate(S, K, T, sigma, option_type)
        try:
            (total, _, n), = self._scalar_moments([(S, K, T, r, sigma, q)], option_type, seed)
            return runtime.discount(r, T) * total / n
        except Exception as e:  # same wrapping as monte_carlo_unified.py:510-511
            raise MonteCarloError(f"Monte Carlo pricing failed: {e}")

    def delta_gamma(self, S: float, K: float, T: float, r: float, sigma: float, option_type: Literal["call", "put"],
                    q: float = 0.0, h: float = 1e-4, seed: Optional[int] = None) -> Tuple[float, float]:
        if seed is None:
            seed = int(self.rng.integers(0, 2**31))
        self._validate(S + h, K, T, sigma, option_type)
        self._validate(S - h, K, T, sigma, option_type)
        try:  # S+h, S, S-h are three scenarios of ONE launch: same draws (monte_carlo_unified.py:547-549 re-seeds instead)
            m = self._scalar_moments([(S + h, K, T, r, sigma, q), (S, K, T, r, sigma, q), (S - h, K, T, r, sigma, q)], option_type, seed)
        except Exception as e:
            raise MonteCarloError(f"Monte Carlo pricing failed: {e}")
        disc = runtime.discount(r, T)
        up, mid, down = (disc * total / n for total, _, n in m)
        return (up - down) / (2 * h), (up - 2 * mid + down) / (h**2)

    def _scalar_moments(self, scenarios, option_type, seed):
        spec = _ffi.make_spec(_ffi.EUROPEAN, self.num_steps, is_put=(option_type == "put"), antithetic=True)
        with self._lock:
            return runtime.simulate_scalars(spec, scenarios, seed if seed is not None else self.seed, self.num_simulations)

    @staticmethod
    def _arrays(S_vals, K_vals, T_vals, r_vals, sigma_vals, q_vals):
        S_vals = np.asarray(S_vals, dtype=np.float64)
        K_vals = np.asarray(K_vals, dtype=np.float64)
        T_vals = np.asarray(T_vals, dtype=np.float64)
        r_vals = np.asarray(r_vals, dtype=np.float64)
        sigma_vals = np.asarray(sigma_vals, dtype=np.float64)
        if isinstance(q_vals, (int, float)):
            q_vals = np.full_like(S_vals, q_vals)
        else:
            q_vals = np.asarray(q_vals, dtype=np.float64)
        return S_vals, K_vals, T_vals, r_vals, sigma_vals, q_vals

    def price_batch(self, S_vals, K_vals, T_vals, r_vals, sigma_vals, option_type: Literal["call", "put"],
                    q_vals: Union[float, np.ndarray] = 0.0) -> np.ndarray:
        S_vals, K_vals, T_vals, r_vals, sigma_vals, q_vals = self._arrays(S_vals, K_vals, T_vals, r_vals, sigma_vals, q_vals)
        if len(S_vals) == 0:
            return np.empty(0, dtype=np.float64)
        m = self._batch_moments(S_vals, K_vals, T_vals, r_vals, sigma_vals, q_vals, option_type, None)
        return runtime.discounted_price(m[:, 0], r_vals, T_vals)

    def delta_gamma_batch(self, S_vals, K_vals, T_vals, r_vals, sigma_vals, option_type: Literal["call", "put"],
                          q_vals: Union[float, np.ndarray] = 0.0, h: float = 1e-4) -> Tuple[np.ndarray, np.ndarray]:
        S_vals, K_vals, T_vals, r_vals, sigma_vals, q_vals = self._arrays(S_vals, K_vals, T_vals, r_vals, sigma_vals, q_vals)
        if len(S_vals) == 0:
            return np.empty(0), np.empty(0)
        m = self._batch_moments(S_vals, K_vals, T_vals, r_vals, sigma_vals, q_vals, option_type, None, bumps=(-h, 0.0, h))
        prices = runtime.discounted_price(m, r_vals[:, None], T_vals[:, None])
        down, mid, up = prices[:, 0], prices[:, 1], prices[:, 2]
        return (up - down) / (2 * h), (up - 2 * mid + down) / (h**2)

    # -- fused Greeks (PricerProtocol + price_scenarios) ----------------------------------------
    def price_scenarios(self, scenarios, option_type, seed: Optional[int] = None, **_ignored):
        scenarios = [tuple(float(x) for x in sc) for sc in scenarios]
        out = []
        for lo in range(0, len(scenarios), _ffi.MAX_SCENARIOS):
            blk = scenarios[lo:lo + _ffi.MAX_SCENARIOS]
            out += [runtime.discount(sc[3], sc[2]) * total / n for sc, (total, _, n) in zip(blk, self._scalar_moments(blk, option_type, seed))]
        return out

    def greeks(self, S, K, T, r, sigma, option_type="call", q=0.0, include_second_order=True):
        return compute_greeks_unified(self, S, K, T, r, sigma, option_type, q, include_second_order)
