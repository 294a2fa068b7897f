"""Asian / barrier / lookback options on the B200 engine — drop-ins for
src/pricing_models/exotic_options.py:28-224,347-401 (dataclass fields, ``price`` signatures,
``ValueError("Barrier must be positive")``, ``np.float64`` return, no antithetic mirroring).

The reference materialises an ``(n_paths, n_steps+1)`` path array (46.8 GB for 16M x 365); here
the running average / barrier flag / extremum lives in registers and no path is ever stored.
Monitoring conventions are the reference's: Asian averages exclude t=0 (:119-122), barrier and
lookback extrema include t=0 and use >= / <= (:201-204, :382-383).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Literal, Optional, Sequence

import numpy as np

from . import _ffi, runtime
from .monte_carlo import MCResult

__all__ = ["ExoticOptionBase", "AsianOption", "BarrierOption", "LookbackOption", "AutocallableOption", "CliquetOption",
           "price_asian", "price_barrier", "price_lookback"]


@dataclass
class ExoticOptionBase:
    S: float
    K: float
    T: float
    r: float
    sigma: float
    q: float = 0.0
    seed: Optional[int] = None

    def _seed(self) -> int:
        return self.seed if self.seed is not None else runtime.entropy_seed()

    def _run(self, spec, scenarios: Sequence, n_paths: int, barrier: float = 0.0):
        """-> ([(sum, sum_sq, n)], scenarios) of (S, K, T, r, sigma, q) tuples on common random numbers: one fused launch per
        16 scenarios through the engine's latency path (runtime.simulate_scalars)."""
        sc = [tuple(float(x) for x in s) for s in scenarios]
        seed = self._seed()
        out = []
        for lo in range(0, len(sc), _ffi.MAX_SCENARIOS):
            out += runtime.simulate_scalars(spec, sc[lo:lo + _ffi.MAX_SCENARIOS], seed, n_paths, barrier=barrier)
        return out, sc

    def _finish(self, m, sc, return_error):
        res = []
        for (total, total_sq, n), s in zip(m, sc):
            disc = runtime.discount(s[3], s[2])
            price = disc * total / n  # exp(-rT) * mean(payoffs), exotic_options.py:131
            if return_error:
                mean = total / n
                res.append(MCResult(float(price), float(disc * np.sqrt(max(total_sq / n - mean * mean, 0.0)) / np.sqrt(n)), int(n)))
            else:
                res.append(np.float64(price))
        return res

    def _self_scenario(self):
        return [(self.S, self.K, self.T, self.r, self.sigma, self.q)]


@dataclass
class AsianOption(ExoticOptionBase):
    """exotic_options.py:88-160."""

    def _spec(self, n_steps, avg_type, option_type):
        kind = _ffi.ASIAN_ARITH if avg_type == "arithmetic" else _ffi.ASIAN_GEOM  # any other string = geometric (:121)
        return _ffi.make_spec(kind, n_steps, is_put=(option_type != "call"))

    def price(self, n_paths: int = 100000, n_steps: int = 252,
              avg_type: Literal["arithmetic", "geometric"] = "arithmetic",
              option_type: Literal["call", "put"] = "call", return_error: bool = False):
        m, sc = self._run(self._spec(n_steps, avg_type, option_type), self._self_scenario(), n_paths)
        return self._finish(m, sc, return_error)[0]

    def price_scenarios(self, scenarios, n_paths: int = 100000, n_steps: int = 252, avg_type="arithmetic",
                        option_type="call"):
        m, sc = self._run(self._spec(n_steps, avg_type, option_type), scenarios, n_paths)
        return [float(p) for p in self._finish(m, sc, False)]

    def price_geometric_closed_form(self, option_type: Literal["call", "put"] = "call") -> float:
        """Continuous-averaging closed form (exotic_options.py:133-160) — host arithmetic, no simulation."""
        from math import erf, exp, log, sqrt

        def cdf(x):
            return 0.5 * (1.0 + erf(x / sqrt(2.0)))

        sigma_adj = self.sigma / sqrt(3)
        r_adj = 0.5 * (self.r - self.q - self.sigma**2 / 6)
        d1 = (log(self.S / self.K) + (r_adj + 0.5 * sigma_adj**2) * self.T) / (sigma_adj * sqrt(self.T))
        d2 = d1 - sigma_adj * sqrt(self.T)
        if option_type == "call":
            return self.S * exp((r_adj - self.r) * self.T) * cdf(d1) - self.K * exp(-self.r * self.T) * cdf(d2)
        return self.K * exp(-self.r * self.T) * cdf(-d2) - self.S * exp((r_adj - self.r) * self.T) * cdf(-d1)


@dataclass
class BarrierOption(ExoticOptionBase):
    """exotic_options.py:163-224."""

    barrier: float = 0.0

    def _spec(self, n_steps, barrier_type, option_type):
        if self.barrier <= 0:
            raise ValueError("Barrier must be positive")
        return _ffi.make_spec(_ffi.BARRIER, n_steps, is_put=(option_type != "call"),
                              barrier_down=not barrier_type.startswith("up"), barrier_in=not barrier_type.endswith("out"))

    def price(self, n_paths: int = 100000, n_steps: int = 252,
              barrier_type: Literal["up-and-out", "up-and-in", "down-and-out", "down-and-in"] = "up-and-out",
              option_type: Literal["call", "put"] = "call", return_error: bool = False):
        m, sc = self._run(self._spec(n_steps, barrier_type, option_type), self._self_scenario(), n_paths, self.barrier)
        return self._finish(m, sc, return_error)[0]

    def price_scenarios(self, scenarios, n_paths: int = 100000, n_steps: int = 252, barrier_type="up-and-out",
                        option_type="call"):
        m, sc = self._run(self._spec(n_steps, barrier_type, option_type), scenarios, n_paths, self.barrier)
        return [float(p) for p in self._finish(m, sc, False)]


@dataclass
class LookbackOption(ExoticOptionBase):
    """exotic_options.py:347-401."""

    def _spec(self, n_steps, lookback_type, option_type):
        return _ffi.make_spec(_ffi.LOOKBACK, n_steps, is_put=(option_type != "call"),
                              lookback_fixed=(lookback_type != "floating"))

    def price(self, n_paths: int = 100000, n_steps: int = 252, lookback_type: Literal["floating", "fixed"] = "floating",
              option_type: Literal["call", "put"] = "call", return_error: bool = False):
        m, sc = self._run(self._spec(n_steps, lookback_type, option_type), self._self_scenario(), n_paths)
        return self._finish(m, sc, return_error)[0]

    def price_scenarios(self, scenarios, n_paths: int = 100000, n_steps: int = 252, lookback_type="floating",
                        option_type="call"):
        m, sc = self._run(self._spec(n_steps, lookback_type, option_type), scenarios, n_paths)
        return [float(p) for p in self._finish(m, sc, False)]


def _as_tuples(m):
    return [(float(x["sum"]), float(x["sum_sq"]), float(x["n"])) for x in m]


def _run_structured(opt: ExoticOptionBase, spec, product, scenarios: Sequence, n_paths: int):
    """Moments of every (S, K, T, r, sigma, q) scenario on common random numbers, sharded over ranks like runtime.simulate."""
    from . import distributed

    sc = np.asarray(scenarios, dtype=np.float64).reshape(-1, 6)
    seed = opt._seed()
    out = []
    for lo in range(0, len(sc), _ffi.MAX_SCENARIOS):
        blk = sc[lo:lo + _ffi.MAX_SCENARIOS]
        params = _ffi.make_params(blk[:, 0], blk[:, 1], blk[:, 2], blk[:, 3], blk[:, 4], blk[:, 5])[None, :]
        m = distributed.run_sharded(lambda eng, begin, count: eng.simulate_structured(spec, product, params, seed, count, path_begin=begin),
                                    n_paths, lambda: np.zeros(params.shape, dtype=_ffi.MOMENTS_DTYPE))
        out.append(m[0])
    return np.concatenate(out), sc


@dataclass
class AutocallableOption(ExoticOptionBase):
    """exotic_options.py:404-488.  Barriers are relative to the spot; the price is a fraction of the notional.  The
    redemption / coupon / knock-in logic runs per path in registers; discounting happens per path, as in the reference."""

    autocall_barrier: float = 1.0
    coupon_barrier: float = 0.8
    coupon_rate: float = 0.10
    ki_barrier: float = 0.6

    def _launch(self, scenarios, n_paths, n_steps, observation_freq):
        if observation_freq == 0:
            raise ValueError("range() arg 3 must not be zero")  # what the reference's range(freq, n_steps + 1, freq) raises
        freq = int(observation_freq) if observation_freq > 0 else int(n_steps) + 1  # a negative step gives no observation dates
        product = _ffi.Product(self.autocall_barrier, self.coupon_barrier, self.coupon_rate, self.ki_barrier, min(freq, 0xFFFFFFFF), 0)
        return _run_structured(self, _ffi.make_spec(_ffi.AUTOCALLABLE, n_steps), product, scenarios, n_paths)

    def price(self, n_paths: int = 100000, n_steps: int = 252, observation_freq: int = 21, return_error: bool = False, **kwargs):
        m, _ = self._launch(self._self_scenario(), n_paths, n_steps, observation_freq)
        mean = m["sum"][0] / m["n"][0]  # payoffs are already discounted (exotic_options.py:466,486-488)
        if return_error:
            var = max(m["sum_sq"][0] / m["n"][0] - mean * mean, 0.0)
            return MCResult(float(mean), float(np.sqrt(var / m["n"][0])), int(m["n"][0]))
        return np.float64(mean)

    def price_scenarios(self, scenarios, n_paths: int = 100000, n_steps: int = 252, observation_freq: int = 21, **kwargs):
        m, _ = self._launch(scenarios, n_paths, n_steps, observation_freq)
        return [float(x) for x in m["sum"] / m["n"]]


@dataclass
class CliquetOption(ExoticOptionBase):
    """exotic_options.py:491-552: sum over n_periods reset periods of the locally capped / floored simple returns,
    capped / floored globally, floored at 0, times S, discounted."""

    local_cap: float = 0.05
    local_floor: float = -0.05
    global_cap: float = 0.30
    global_floor: float = 0.0

    def _launch(self, scenarios, n_paths, n_steps, n_periods):
        if n_periods == 0:
            raise ZeroDivisionError("integer division or modulo by zero")  # n_steps // n_periods in the reference
        product = _ffi.Product(self.local_cap, self.local_floor, self.global_cap, self.global_floor, int(n_periods), 0)
        return _run_structured(self, _ffi.make_spec(_ffi.CLIQUET, n_steps), product, scenarios, n_paths)

    def _degenerate(self, scenarios, n_steps, n_periods):
        """More periods than steps: every period starts and ends at column 0, every return is 0 (exotic_options.py:532-546)."""
        total = float(np.clip(sum(float(np.clip(0.0, self.local_floor, self.local_cap)) for _ in range(n_periods)), self.global_floor, self.global_cap))
        return [float(np.exp(-s[3] * s[2]) * max(total, 0.0) * s[0]) for s in scenarios]

    def price(self, n_paths: int = 100000, n_steps: int = 252, n_periods: int = 12, return_error: bool = False, **kwargs):
        if 0 < n_steps < n_periods:
            p = self._degenerate(self._self_scenario(), n_steps, n_periods)[0]
            return MCResult(p, 0.0, int(n_paths)) if return_error else np.float64(p)
        m, sc = self._launch(self._self_scenario(), n_paths, n_steps, n_periods)
        return self._finish(_as_tuples(m), sc.tolist(), return_error)[0]

    def price_scenarios(self, scenarios, n_paths: int = 100000, n_steps: int = 252, n_periods: int = 12, **kwargs):
        if 0 < n_steps < n_periods:
            return self._degenerate(scenarios, n_steps, n_periods)
        m, sc = self._launch(scenarios, n_paths, n_steps, n_periods)
        return [float(p) for p in self._finish(_as_tuples(m), sc.tolist(), False)]


def price_asian(S, K, T, r, sigma, avg_type="arithmetic", option_type="call", n_paths=100000, seed=None) -> float:
    """exotic_options.py:558-572."""
    return AsianOption(S=S, K=K, T=T, r=r, sigma=sigma, seed=seed).price(n_paths=n_paths, avg_type=avg_type, option_type=option_type)


def price_barrier(S, K, T, r, sigma, barrier, barrier_type="up-and-out", option_type="call", n_paths=100000, seed=None) -> float:
    """exotic_options.py:575-590."""
    return BarrierOption(S=S, K=K, T=T, r=r, sigma=sigma, barrier=barrier, seed=seed).price(
        n_paths=n_paths, barrier_type=barrier_type, option_type=option_type)


def price_lookback(S, K, T, r, sigma, lookback_type="floating", option_type="call", n_paths=100000, seed=None) -> float:
    return LookbackOption(S=S, K=K, T=T, r=r, sigma=sigma, seed=seed).price(
        n_paths=n_paths, lookback_type=lookback_type, option_type=option_type)
