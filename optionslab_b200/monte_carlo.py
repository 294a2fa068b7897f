"""``MonteCarloPricer`` — drop-in for src/pricing_models/monte_carlo.py:46-186 on the B200 engine.

Same constructor (num_simulations / num_steps / seed / method), same ``price`` signature and
return types, antithetic counting (2N samples, monte_carlo.py:150) and T<=0 intrinsic shortcut.
What differs is where the numbers come from: one fused CUDA launch (Philox draws in registers,
log-Euler steps, payoff, FP64 moments) instead of a NumPy ``(N, n_steps)`` normal array.
"""

from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from enum import Enum
from typing import Literal, Optional, Sequence, Union

import numpy as np

from . import _ffi, runtime
from .greeks import compute_greeks_unified

__all__ = ["MonteCarloPricer", "MCMethod", "MCResult", "NUMBA_AVAILABLE"]

NUMBA_AVAILABLE = False  # kept for import compatibility (monte_carlo.py:189); no Numba path exists here


class MCMethod(Enum):
    """monte_carlo.py:28-34.  NUMPY / NUMBA / FAST all run the fused Philox kernels (FAST and
    ``num_steps == 1`` both mean one exact step, as in monte_carlo.py:86-92); QMC runs the Sobol kernel on
    the reference's own scrambled point set (gbm_qmc.py:14-47; N samples, no mirroring)."""

    NUMPY = "numpy"
    NUMBA = "numba"
    QMC = "qmc"
    FAST = "fast"


@dataclass
class MCResult:
    """monte_carlo.py:37-43."""

    price: float
    std_error: float = 0.0
    n_paths: int = 0


class MonteCarloPricer:
    __slots__ = ("num_simulations", "num_steps", "seed", "method", "_use_numba")

    def __init__(self, num_simulations: int = 100000, num_steps: int = 1, seed: Optional[int] = None,
                 method: MCMethod = MCMethod.NUMPY, use_numba: Optional[bool] = None):
        if num_simulations < 1:
            raise ValueError("num_simulations must be >= 1")
        self.num_simulations = num_simulations
        self.num_steps = num_steps
        self.seed = seed if seed is not None else runtime.fresh_seed()
        self.method = method
        self._use_numba = False  # stale callers pass use_numba= (streamlit_app/st_utils.py:309-314); ignored

    # -- internals ----------------------------------------------------------------------------
    def _steps(self) -> int:
        if self.method == MCMethod.FAST:
            return 1
        if self.method == MCMethod.QMC:
            return min(max(int(self.num_steps), 1), 21201)  # gbm_qmc.py:30
        return max(int(self.num_steps), 1)

    def _moments(self, scenarios: Sequence, option_type: str, seed: Optional[int]):
        actual_seed = seed if seed is not None else self.seed  # seed=0 is honoured (monte_carlo.py:84)
        qmc = self.method == MCMethod.QMC
        spec = _ffi.make_spec(_ffi.EUROPEAN, self._steps(), is_put=runtime.validate_option_type(option_type), antithetic=not qmc)
        sc = np.asarray(scenarios, dtype=np.float64).reshape(-1, 6)
        params = _ffi.make_params(sc[:, 0], sc[:, 1], sc[:, 2], sc[:, 3], sc[:, 4], sc[:, 5])[None, :]
        if qmc:  # monte_carlo.py:94-97 -> gbm_qmc.py:14-47
            return runtime.simulate_sobol(spec, params, actual_seed, self.num_simulations)[0], sc
        return runtime.simulate(spec, params, actual_seed, self.num_simulations)[0], sc

    # -- reference surface --------------------------------------------------------------------
    def price(self, S: float, K: float, T: float, r: float, sigma: float, option_type: Literal["call", "put"],
              q: float = 0.0, seed: Optional[int] = None, return_error: bool = False) -> Union[float, MCResult]:
        if T <= 0:
            intrinsic = max(S - K, 0) if option_type == "call" else max(K - S, 0)
            return MCResult(intrinsic, 0.0, 0) if return_error else intrinsic
        m, _ = self._moments([(S, K, T, r, sigma, q)], option_type, seed)
        price = float(runtime.discounted_price(m[0], r, T))
        if return_error:
            return MCResult(price, float(runtime.discounted_std_error(m[0], r, T)), int(m[0]["n"]))
        return price

    def price_with_control_variate(self, S: float, K: float, T: float, r: float, sigma: float,
                                   option_type: Literal["call", "put"], q: float = 0.0, seed: Optional[int] = None) -> float:
        """Terminal spot as control variate, E[S_T] = S*exp((r-q)T) (monte_carlo.py:154-186).  The five
        sums np.cov needs come out of the same fused launch as the payoff."""
        actual_seed = seed if seed is not None else self.seed
        spec = _ffi.make_spec(_ffi.EUROPEAN, self._steps(), is_put=runtime.validate_option_type(option_type), antithetic=True)
        m = runtime.simulate(spec, _ffi.make_params(S, K, T, r, sigma, q).reshape(1, 1), actual_seed, self.num_simulations,
                             control_variate=True)[0, 0]
        return runtime.control_variate_price(m, S, T, r, q)

    # -- fused common-random-number surface ---------------------------------------------------
    def price_scenarios(self, scenarios: Sequence, option_type: str, seed: Optional[int] = None, **_ignored):
        """Prices of up to 16 (S,K,T,r,sigma,q) scenarios sharing one set of draws, ONE launch."""
        out = [None] * len(scenarios)
        live = [i for i, s in enumerate(scenarios) if s[2] > 0]
        for i, s in enumerate(scenarios):
            if s[2] <= 0:
                out[i] = max(s[0] - s[1], 0) if option_type == "call" else max(s[1] - s[0], 0)
        for lo in range(0, len(live), _ffi.MAX_SCENARIOS):
            idx = live[lo:lo + _ffi.MAX_SCENARIOS]
            m, sc = self._moments([scenarios[i] for i in idx], option_type, seed)
            prices = runtime.discounted_price(m, sc[:, 3], sc[:, 2])
            for i, p in zip(idx, prices):
                out[i] = float(p)
        return out

    def greeks(self, S, K, T, r, sigma, option_type="call", q=0.0, include_second_order=True) -> "OrderedDict[str, float]":
        """All Greeks of unified_greeks.py:235-367 from one fused launch."""
        return compute_greeks_unified(self, S, K, T, r, sigma, option_type, q, include_second_order)

    def _greek(self, name, S, K, T, r, sigma, option_type, q):
        return self.greeks(S, K, T, r, sigma, option_type, q, include_second_order=False)[name]

    # README.md:265-269 documents these five calls on the pricer object
    def delta(self, S, K, T, r, sigma, option_type="call", q=0.0):
        return self._greek("delta", S, K, T, r, sigma, option_type, q)

    def gamma(self, S, K, T, r, sigma, option_type="call", q=0.0):
        return self._greek("gamma", S, K, T, r, sigma, option_type, q)

    def vega(self, S, K, T, r, sigma, option_type="call", q=0.0):
        return self._greek("vega", S, K, T, r, sigma, option_type, q)

    def theta(self, S, K, T, r, sigma, option_type="call", q=0.0):
        return self._greek("theta", S, K, T, r, sigma, option_type, q)

    def rho(self, S, K, T, r, sigma, option_type="call", q=0.0):
        return self._greek("rho", S, K, T, r, sigma, option_type, q)

    def delta_gamma(self, S, K, T, r, sigma, option_type="call", q=0.0, h: float = 1e-4, seed: Optional[int] = None):
        """Central differences at S±h over shared draws (the call streamlit_app/st_utils.py:575 makes)."""
        up, mid, down = self.price_scenarios([(S + h, K, T, r, sigma, q), (S, K, T, r, sigma, q), (S - h, K, T, r, sigma, q)],
                                             option_type, seed=seed)
        return (up - down) / (2 * h), (up - 2 * mid + down) / (h**2)
