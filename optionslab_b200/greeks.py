"""Bump-and-revalue Greeks with common random numbers — the reference's
``compute_greeks_unified`` (src/greeks/unified_greeks.py:235-367) with one difference in HOW the
scenario prices are obtained: a pricer that exposes ``price_scenarios`` (all pricers of this
package do) gets every bumped scenario priced in ONE fused kernel launch over shared Philox draws,
instead of 8–14 separate simulations.  Any other ``PricerProtocol`` object is priced call by
call exactly as the reference does, so this function is a superset of the reference's.

Bump sizes, formulas, key order and error wrapping follow unified_greeks.py:274-367 line by line.
"""

from __future__ import annotations

from collections import OrderedDict
from enum import IntEnum
from typing import Dict, List, Literal, Protocol, Tuple, runtime_checkable

from .exceptions import GreeksError

__all__ = ["PricerProtocol", "ExoticAdapter", "HestonAdapter", "JumpDiffusionAdapter", "compute_greeks_unified",
           "greeks_heston", "greeks_jump_diffusion", "greek_scenarios", "greeks_from_prices", "OptionType", "ExerciseStyle"]


class OptionType(IntEnum):
    """src/greeks/greeks.py:21-25."""

    CALL = 0
    PUT = 1


class ExerciseStyle(IntEnum):
    """src/greeks/greeks.py:28-32."""

    EUROPEAN = 0
    AMERICAN = 1


@runtime_checkable
class PricerProtocol(Protocol):
    """src/greeks/unified_greeks.py:45-66."""

    def price(self, S: float, K: float, T: float, r: float, sigma: float, option_type: Literal["call", "put"],
              q: float = 0.0, **kwargs) -> float: ...


Scenario = Tuple[float, float, float, float, float, float]  # (S, K, T, r, sigma, q)


def greek_bumps(S: float):
    """h_S, h_sigma, h_r, h_T of unified_greeks.py:274-277."""
    return max(1e-4, 0.01 * S), max(1e-4, 0.01), 1e-4, 1 / 365.0


def greek_scenarios(S, K, T, r, sigma, q=0.0, include_second_order=True) -> List[Scenario]:
    """The distinct (S,K,T,r,sigma,q) points unified_greeks.py:295-355 prices, base point first."""
    h_S, h_sigma, h_r, h_T = greek_bumps(S)
    pts = [(S, K, T, r, sigma, q), (S + h_S, K, T, r, sigma, q), (S - h_S, K, T, r, sigma, q),
           (S, K, T, r, sigma + h_sigma, q), (S, K, T, r, sigma - h_sigma, q)]
    if T > h_T:
        pts.append((S, K, T - h_T, r, sigma, q))
    pts += [(S, K, T, r + h_r, sigma, q), (S, K, T, r - h_r, sigma, q)]
    if include_second_order:
        pts += [(S + h_S, K, T, r, sigma + h_sigma, q), (S + h_S, K, T, r, sigma - h_sigma, q),
                (S - h_S, K, T, r, sigma + h_sigma, q), (S - h_S, K, T, r, sigma - h_sigma, q)]
        if T > h_T:
            pts += [(S + h_S, K, T - h_T, r, sigma, q), (S - h_S, K, T - h_T, r, sigma, q)]
    return pts


def greeks_from_prices(P: Dict[Scenario, float], S, K, T, r, sigma, q=0.0, include_second_order=True):
    """Finite-difference formulas of unified_greeks.py:295-362 over a scenario->price map."""
    h_S, h_sigma, h_r, h_T = greek_bumps(S)

    def at(S_=S, T_=T, r_=r, sigma_=sigma):
        return P[(S_, K, T_, r_, sigma_, q)]

    p_mid = at()
    p_S_up, p_S_down = at(S_=S + h_S), at(S_=S - h_S)
    delta = (p_S_up - p_S_down) / (2 * h_S)
    gamma = (p_S_up - 2 * p_mid + p_S_down) / (h_S**2)
    p_sigma_up, p_sigma_down = at(sigma_=sigma + h_sigma), at(sigma_=sigma - h_sigma)
    vega = (p_sigma_up - p_sigma_down) / (2 * h_sigma)
    if T > h_T:
        theta = (at(T_=T - h_T) - p_mid) / h_T
    else:
        theta = -p_mid / max(T, 1e-6)
    rho = (at(r_=r + h_r) - at(r_=r - h_r)) / (2 * h_r)
    greeks = OrderedDict([("price", p_mid), ("delta", delta), ("gamma", gamma), ("vega", vega), ("theta", theta),
                          ("rho", rho)])
    if include_second_order:
        vanna = (at(S_=S + h_S, sigma_=sigma + h_sigma) - at(S_=S + h_S, sigma_=sigma - h_sigma)
                 - at(S_=S - h_S, sigma_=sigma + h_sigma) + at(S_=S - h_S, sigma_=sigma - h_sigma)) / (4 * h_S * h_sigma)
        if T > h_T:
            delta_T_down = (at(S_=S + h_S, T_=T - h_T) - at(S_=S - h_S, T_=T - h_T)) / (2 * h_S)
            charm = (delta_T_down - delta) / h_T
        else:
            charm = 0.0
        vomma = (p_sigma_up - 2 * p_mid + p_sigma_down) / (h_sigma**2)
        greeks["vanna"], greeks["charm"], greeks["vomma"] = vanna, charm, vomma
    return greeks


class ExoticAdapter:
    """src/greeks/unified_greeks.py:177-227: make an exotic option look like a PricerProtocol by
    overwriting its S/K/T/r/sigma/q and forwarding n_paths/n_steps/kwargs to ``exotic.price``."""

    def __init__(self, exotic_option, n_paths: int = 50000, n_steps: int = 252, **exotic_kwargs):
        self.exotic = exotic_option
        self.n_paths = n_paths
        self.n_steps = n_steps
        self.exotic_kwargs = exotic_kwargs

    def _kwargs(self, option_type, kwargs):
        price_kwargs = {**self.exotic_kwargs, **kwargs}
        if "option_type" not in price_kwargs:
            price_kwargs["option_type"] = option_type
        return price_kwargs

    def price(self, S, K, T, r, sigma, option_type, q=0.0, **kwargs) -> float:
        ex = self.exotic
        ex.S, ex.K, ex.T, ex.r, ex.sigma, ex.q = S, K, T, r, sigma, q
        return ex.price(n_paths=self.n_paths, n_steps=self.n_steps, **self._kwargs(option_type, kwargs))

    def price_scenarios(self, scenarios, option_type, **kwargs):
        """Fused CRN path: every scenario in one launch (needs an exotic of this package)."""
        fused = getattr(self.exotic, "price_scenarios", None)
        if fused is None:
            return [self.price(*s[:5], option_type, s[5], **kwargs) for s in scenarios]
        return fused(scenarios, n_paths=self.n_paths, n_steps=self.n_steps, **self._kwargs(option_type, kwargs))


class HestonAdapter:
    """src/greeks/unified_greeks.py:74-104: PricerProtocol face of a HestonPricer, ``sigma`` mapped to v0 = sigma^2.
    The reference's adapter calls the semi-analytic ``price_european``; this package accelerates the Monte Carlo path,
    so the adapter prices with ``price_monte_carlo`` on a FIXED seed (common random numbers across the bumps) and offers
    the fused ``price_scenarios`` route: all 8-14 bumped re-pricings of compute_greeks_unified in one launch."""

    def __init__(self, heston_pricer, n_paths: int = 100000, n_steps: int = 252, seed: int = 0):
        self.heston = heston_pricer
        self.n_paths, self.n_steps, self.seed = n_paths, n_steps, seed
        self._original_v0 = heston_pricer.v0

    def price(self, S, K, T, r, sigma, option_type, q=0.0, **kwargs) -> float:
        self.heston.v0 = sigma**2
        try:
            return self.heston.price_monte_carlo(S, K, T, r, q, option_type, self.n_paths, self.n_steps, seed=self.seed)
        finally:
            self.heston.v0 = self._original_v0

    def price_scenarios(self, scenarios, option_type, **kwargs):
        sc = [(s[0], s[1], s[2], s[3], s[4] ** 2, s[5]) for s in scenarios]
        return self.heston.price_scenarios(sc, option_type, self.n_paths, self.n_steps, self.seed)

    def __repr__(self):
        return f"HestonAdapter({self.heston})"


class JumpDiffusionAdapter:
    """src/greeks/unified_greeks.py:155-174 for MertonJumpDiffusion / KouJumpDiffusion, Monte Carlo backed (fixed seed =
    common random numbers; fused single-launch route as HestonAdapter)."""

    def __init__(self, jd_model, n_paths: int = 100000, n_steps: int = 252, seed: int = 0):
        self.jd = jd_model
        self.n_paths, self.n_steps, self.seed = n_paths, n_steps, seed

    def price(self, S, K, T, r, sigma, option_type, q=0.0, **kwargs) -> float:
        return self.jd.price_monte_carlo(S, K, T, r, sigma, option_type, q, self.n_paths, self.n_steps, seed=self.seed)

    def price_scenarios(self, scenarios, option_type, **kwargs):
        return self.jd.price_scenarios(scenarios, option_type, self.n_paths, self.n_steps, self.seed)


def compute_greeks_unified(pricer: PricerProtocol, S: float, K: float, T: float, r: float, sigma: float,
                           option_type: Literal["call", "put"] = "call", q: float = 0.0,
                           include_second_order: bool = True, **pricer_kwargs) -> "OrderedDict[str, float]":
    """Same signature, keys, order, bumps and error behaviour as unified_greeks.py:235-367."""
    try:
        pts = greek_scenarios(S, K, T, r, sigma, q, include_second_order)
        fused = getattr(pricer, "price_scenarios", None)
        if fused is not None:
            prices = fused(pts, option_type, **pricer_kwargs)
        else:
            prices = [pricer.price(s[0], s[1], s[2], s[3], s[4], option_type, s[5], **pricer_kwargs) for s in pts]
        return greeks_from_prices(dict(zip(pts, prices)), S, K, T, r, sigma, q, include_second_order)
    except Exception as e:
        raise GreeksError(f"Failed to compute unified Greeks: {str(e)}") from e


def greeks_heston(heston_pricer, S, K, T, r, sigma, option_type: str = "call", q: float = 0.0, **adapter_kwargs):
    """unified_greeks.py:375-388 (Monte Carlo backed; ``adapter_kwargs`` = n_paths / n_steps / seed)."""
    return compute_greeks_unified(HestonAdapter(heston_pricer, **adapter_kwargs), S, K, T, r, sigma, option_type, q)


def greeks_jump_diffusion(jd_model, S, K, T, r, sigma, option_type: str = "call", q: float = 0.0, **adapter_kwargs):
    return compute_greeks_unified(JumpDiffusionAdapter(jd_model, **adapter_kwargs), S, K, T, r, sigma, option_type, q)
