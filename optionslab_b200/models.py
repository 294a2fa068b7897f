"""Heston and jump-diffusion Monte Carlo pricers on the B200 engine — the ``price_monte_carlo`` methods of
src/pricing_models/heston.py:184-255 and src/pricing_models/jump_diffusion.py:160-225, :325-377.

Only the Monte Carlo path of those classes lives here (same constructors, validation and call signatures);
their semi-analytic formulas (Lewis integral, Merton series) are CPU closed forms outside the hot path.
``seed=None`` prices with fresh entropy (the reference leaves NumPy's global state unseeded)."""

from __future__ import annotations

import math
import warnings
from dataclasses import dataclass
from typing import Literal, Optional

import numpy as np

from . import _ffi, distributed, runtime

__all__ = ["HestonPricer", "MertonJumpDiffusion", "KouJumpDiffusion"]


def _sharded(run, n_paths: int, n_out: int = 1) -> np.ndarray:
    """``run(engine, path_begin, count)`` over this process's / device's share of the global paths, moments combined."""
    return distributed.run_sharded(run, n_paths, lambda: np.zeros(n_out, dtype=_ffi.MOMENTS_DTYPE))


@dataclass
class HestonPricer:
    """heston.py:41-82: dS = (r-q)S dt + sqrt(v) S dW1, dv = kappa(theta - v)dt + sigma_v sqrt(v) dW2, corr rho."""

    kappa: float
    theta: float
    sigma_v: float
    rho: float
    v0: float

    def __post_init__(self):
        if self.kappa <= 0:
            raise ValueError("kappa must be positive")
        if self.theta <= 0:
            raise ValueError("theta must be positive")
        if self.sigma_v <= 0:
            raise ValueError("sigma_v must be positive")
        if not -1 <= self.rho <= 1:
            raise ValueError("rho must be in [-1, 1]")
        if self.v0 <= 0:
            raise ValueError("v0 must be positive")
        feller = 2 * self.kappa * self.theta - self.sigma_v**2
        if feller < 0:
            warnings.warn(f"Feller condition not satisfied (2κθ - σᵥ² = {feller:.4f} < 0). Variance may hit zero in simulations.")

    def price_monte_carlo(self, S: float, K: float, T: float, r: float, q: float = 0.0,
                          option_type: Literal["call", "put"] = "call", n_paths: int = 100000, n_steps: int = 252,
                          seed: Optional[int] = None, return_error: bool = False):
        """Full-truncation Euler (heston.py:184-255), one fused launch.  ``return_error=True`` additionally
        returns the standard error of the estimate (not available from the reference)."""
        actual_seed = runtime.entropy_seed() if seed is None else int(seed)
        params = np.zeros(1, dtype=_ffi.HESTON_PARAMS_DTYPE)
        for name, val in (("S", S), ("K", K), ("T", T), ("r", r), ("q", q), ("kappa", self.kappa), ("theta", self.theta),
                          ("sigma_v", self.sigma_v), ("rho", self.rho), ("v0", self.v0)):
            params[name] = val
        is_put = option_type != "call"  # heston.py:249-252
        m = _sharded(lambda eng, b, c: eng.simulate_heston(params, is_put, n_steps, actual_seed, c, path_begin=b), int(n_paths))[0]
        price = float(runtime.discounted_price(m, r, T))
        if return_error:
            return price, float(runtime.discounted_std_error(m, r, T))
        return price

    def price_scenarios(self, scenarios, option_type: str = "call", n_paths: int = 100000, n_steps: int = 252, seed: int = 0):
        """Common-random-number re-pricings in ONE launch: ``scenarios`` = (S, K, T, r, v0, q) tuples, all simulated on the
        same draws (the bumped prices compute_greeks_unified asks a Heston adapter for, unified_greeks.py:295-358)."""
        sc = np.asarray(scenarios, dtype=np.float64).reshape(-1, 6)
        params = np.zeros(len(sc), dtype=_ffi.HESTON_PARAMS_DTYPE)
        params["S"], params["K"], params["T"], params["r"], params["v0"], params["q"] = sc.T
        params["kappa"], params["theta"], params["sigma_v"], params["rho"] = self.kappa, self.theta, self.sigma_v, self.rho
        m = _sharded(lambda eng, b, c: eng.simulate_heston(params, option_type != "call", n_steps, int(seed), c, path_begin=b,
                                                           shared_stream=True), int(n_paths), len(sc))
        return [float(x) for x in runtime.discounted_price(m, sc[:, 3], sc[:, 2])]


def _jump_scenarios(model: int, lambda_j: float, a: float, b: float, c: float, scenarios, option_type, n_paths, n_steps, seed):
    sc = np.asarray(scenarios, dtype=np.float64).reshape(-1, 6)  # (S, K, T, r, sigma, q)
    params = _ffi.make_params(sc[:, 0], sc[:, 1], sc[:, 2], sc[:, 3], sc[:, 4], sc[:, 5])
    jumps = np.zeros(len(sc), dtype=_ffi.JUMP_PARAMS_DTYPE)
    jumps["model"], jumps["lambda_j"], jumps["a"], jumps["b"], jumps["c"] = model, lambda_j, a, b, c
    m = _sharded(lambda eng, b_, c_: eng.simulate_jump_diffusion(params, jumps, option_type != "call", n_steps, int(seed), c_, path_begin=b_,
                                                                 shared_stream=True), int(n_paths), len(sc))
    return [float(x) for x in runtime.discounted_price(m, sc[:, 3], sc[:, 2])]


def _jump_price(model: int, lambda_j: float, a: float, b: float, c: float, S, K, T, r, sigma, option_type, q, n_paths, n_steps,
                seed, return_error):
    actual_seed = runtime.entropy_seed() if seed is None else int(seed)
    params = _ffi.make_params(S, K, T, r, sigma, q).reshape(1)
    jumps = np.zeros(1, dtype=_ffi.JUMP_PARAMS_DTYPE)
    jumps["model"], jumps["lambda_j"], jumps["a"], jumps["b"], jumps["c"] = model, lambda_j, a, b, c
    is_put = option_type != "call"
    m = _sharded(lambda eng, b_, c_: eng.simulate_jump_diffusion(params, jumps, is_put, n_steps, actual_seed, c_, path_begin=b_), int(n_paths))[0]
    price = float(runtime.discounted_price(m, r, T))
    if return_error:
        return price, float(runtime.discounted_std_error(m, r, T))
    return price


@dataclass
class MertonJumpDiffusion:
    """jump_diffusion.py:43-67: Poisson(lambda_j) arrivals of N(mu_j, sigma_j^2) log-jumps."""

    lambda_j: float
    mu_j: float
    sigma_j: float

    def __post_init__(self):
        if self.lambda_j < 0:
            raise ValueError("lambda_j must be non-negative")
        if self.sigma_j < 0:
            raise ValueError("sigma_j must be non-negative")

    @property
    def kappa(self) -> float:
        """Mean jump size E[e^Y - 1] (jump_diffusion.py:65-67)."""
        return math.exp(self.mu_j + 0.5 * self.sigma_j**2) - 1

    def price_monte_carlo(self, S: float, K: float, T: float, r: float, sigma: float,
                          option_type: Literal["call", "put"] = "call", q: float = 0.0, n_paths: int = 100000,
                          n_steps: int = 252, seed: Optional[int] = None, return_error: bool = False):
        return _jump_price(_ffi.JUMP_MERTON, self.lambda_j, self.mu_j, self.sigma_j, 0.0, S, K, T, r, sigma, option_type, q,
                           n_paths, n_steps, seed, return_error)

    def price_scenarios(self, scenarios, option_type: str = "call", n_paths: int = 100000, n_steps: int = 252, seed: int = 0):
        """(S, K, T, r, sigma, q) scenarios on common random numbers, one launch."""
        return _jump_scenarios(_ffi.JUMP_MERTON, self.lambda_j, self.mu_j, self.sigma_j, 0.0, scenarios, option_type, n_paths, n_steps, seed)


@dataclass
class KouJumpDiffusion:
    """jump_diffusion.py:274-308: double-exponential log-jumps, +Exp(eta1) w.p. p, -Exp(eta2) otherwise."""

    lambda_j: float
    p: float
    eta1: float
    eta2: float

    def __post_init__(self):
        if not 0 <= self.p <= 1:
            raise ValueError("p must be in [0, 1]")
        if self.eta1 <= 1:
            raise ValueError("eta1 must be > 1 for finite mean")
        if self.eta2 <= 0:
            raise ValueError("eta2 must be positive")

    @property
    def kappa(self) -> float:
        """Mean jump size E[e^Y - 1] (jump_diffusion.py:302-308)."""
        return self.p * self.eta1 / (self.eta1 - 1) + (1 - self.p) * self.eta2 / (self.eta2 + 1) - 1

    def price_monte_carlo(self, S: float, K: float, T: float, r: float, sigma: float,
                          option_type: Literal["call", "put"] = "call", q: float = 0.0, n_paths: int = 100000,
                          n_steps: int = 252, seed: Optional[int] = None, return_error: bool = False):
        return _jump_price(_ffi.JUMP_KOU, self.lambda_j, self.p, self.eta1, self.eta2, S, K, T, r, sigma, option_type, q,
                           n_paths, n_steps, seed, return_error)

    def price_scenarios(self, scenarios, option_type: str = "call", n_paths: int = 100000, n_steps: int = 252, seed: int = 0):
        """(S, K, T, r, sigma, q) scenarios on common random numbers, one launch."""
        return _jump_scenarios(_ffi.JUMP_KOU, self.lambda_j, self.p, self.eta1, self.eta2, scenarios, option_type, n_paths, n_steps, seed)
