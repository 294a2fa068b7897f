"""Host-side dispatch shared by all pricer classes: build the parameter block, run the fused
simulation on this rank's path range, combine moments across ranks, turn moments into prices."""

from __future__ import annotations

import math
import os
from typing import Optional, Sequence

import numpy as np

from . import _ffi, distributed, sobol
from .exceptions import MonteCarloError


def fresh_seed() -> int:
    """Seed for ``seed=None`` constructors: drawn once, so an instance stays deterministic
    (src/pricing_models/monte_carlo.py:68-70, monte_carlo_unified.py:284-286)."""
    return int(np.random.default_rng().integers(0, 2**31))


def entropy_seed() -> int:
    """Unseeded exotic pricing (exotic_options.py:51-52 leaves the global state untouched)."""
    return int.from_bytes(os.urandom(8), "little")


def simulate(spec: _ffi.Spec, params: np.ndarray, seed: int, n_paths: int, *, stream_base: int = 0,
             control_variate: bool = False) -> np.ndarray:
    """Moments [n_opt, n_scen] over ALL ``n_paths`` global paths (sharded + all-reduced when a
    distributed context is active).  Every field of the moment records is a plain sum, so the
    combination across ranks is one element-wise all-reduce."""
    if n_paths < 1:
        raise MonteCarloError("n_paths must be >= 1")
    dtype = _ffi.CV_MOMENTS_DTYPE if control_variate else _ffi.MOMENTS_DTYPE
    fused = None if control_variate else (  # (the control-variate launch exchanges through the engine's collective mode)
        lambda eng, begin, count: eng.simulate(spec, params, seed, count, stream_base=stream_base, path_begin=begin, allreduce=True))
    return distributed.run_sharded(
        lambda eng, begin, count: eng.simulate(spec, params, seed, count, stream_base=stream_base, path_begin=begin,
                                               control_variate=control_variate),
        n_paths, lambda: np.zeros(np.shape(params), dtype=dtype), fused_fn=fused)


def simulate_scalars(spec: _ffi.Spec, scenarios, seed: int, n_paths: int, *, barrier: float = 0.0, stream_base: int = 0):
    """One option, up to 16 common-random-number scenarios given as (S, K, T, r, sigma, q) tuples -> list of
    (sum, sum_sq, n).  The single-device case takes the engine's latency path (no NumPy, one launch, mapped result);
    sharded contexts go through ``simulate``."""
    if n_paths < 1:
        raise MonteCarloError("n_paths must be >= 1")
    ctx = distributed.current()
    if (ctx is None or ctx.world_size == 1) and distributed.local_device_count() <= 1:
        return distributed.default_engine().simulate_scalars(spec, scenarios, seed, n_paths, barrier=barrier, stream_base=stream_base)
    if ctx is not None and ctx.world_size > 1 and ctx.fused:  # one process per GPU: still one launch per rank, nothing else
        begin, count = distributed.partition_paths(n_paths, ctx.rank, ctx.world_size)
        return distributed.default_engine().simulate_scalars(spec, scenarios, seed, count, barrier=barrier, stream_base=stream_base,
                                                             path_begin=begin, allreduce=True)
    sc = np.asarray(scenarios, dtype=np.float64).reshape(-1, 6)
    params = _ffi.make_params(sc[:, 0], sc[:, 1], sc[:, 2], sc[:, 3], sc[:, 4], sc[:, 5], barrier)[None, :]
    m = simulate(spec, params, seed, n_paths, stream_base=stream_base)[0]
    return [(float(x["sum"]), float(x["sum_sq"]), float(x["n"])) for x in m]


def simulate_sobol(spec: _ffi.Spec, params: np.ndarray, seed, n_points: int) -> np.ndarray:
    """Moments [n_opt, n_scen] over the first ``n_points`` points of the reference's scrambled Sobol sequence
    (gbm_qmc.py:32-33: d = n_steps, scramble=True, seed).  Ranks take 4096-aligned slices of the same sequence."""
    if n_points < 1:
        raise MonteCarloError("n_points must be >= 1")
    table, shift, bits = sobol.sobol_table(spec.n_steps, seed)
    if n_points > (1 << bits):
        raise MonteCarloError(f"a {bits}-bit Sobol sequence has 2^{bits} points; asked for {n_points}")
    return distributed.run_sharded(
        lambda eng, begin, count: eng.simulate_sobol(spec, params, table, shift, bits, count, point_begin=begin),
        n_points, lambda: np.zeros(np.shape(params), dtype=_ffi.MOMENTS_DTYPE), partition=sobol.partition_points)


def control_variate_price(m, S, T, r, q) -> float:
    """Terminal-spot control variate from the 5 sums (src/pricing_models/monte_carlo.py:176-186):
    discounted = exp(-rT)*payoff, beta = cov(discounted, S_T)/var(S_T) with ddof = 1 (np.cov),
    beta = 0 when var(S_T) <= 1e-10, result = mean(discounted) - beta*(mean(S_T) - S*exp((r-q)T))."""
    n = float(m["n"])
    disc = math.exp(-r * T)
    mean_d = disc * float(m["sum_payoff"]) / n
    mean_s = float(m["sum_terminal"]) / n
    forward = S * math.exp((r - q) * T)
    if n < 2:
        return mean_d
    cov_ds = disc * (float(m["sum_payoff_terminal"]) - float(m["sum_payoff"]) * float(m["sum_terminal"]) / n) / (n - 1)
    var_s = (float(m["sum_terminal_sq"]) - float(m["sum_terminal"]) ** 2 / n) / (n - 1)
    beta = cov_ds / var_s if var_s > 1e-10 else 0.0
    return float(mean_d - beta * (mean_s - forward))


def discount(r, T) -> float:
    """exp(-rT) through NumPy's exp, so scalar and batched routes (discounted_price) round identically."""
    return float(np.exp(-r * T))


def discounted_price(moments, r, T):
    """exp(-rT) * mean(payoffs)   (monte_carlo.py:145-146)."""
    return np.exp(-np.asarray(r) * np.asarray(T)) * moments["sum"] / moments["n"]


def discounted_std_error(moments, r, T):
    """exp(-rT) * population-std / sqrt(n)   (monte_carlo.py:148-150; ddof = 0 over all samples)."""
    mean = moments["sum"] / moments["n"]
    var = np.maximum(moments["sum_sq"] / moments["n"] - mean * mean, 0.0)
    return np.exp(-np.asarray(r) * np.asarray(T)) * np.sqrt(var) / np.sqrt(moments["n"])


def validate_option_type(option_type: str) -> bool:
    """-> is_put.  The reference treats anything that is not 'call' as a put in MonteCarloPricer
    (monte_carlo.py:140-143) and rejects unknown strings in MonteCarloPricerUni (:490-491)."""
    return option_type != "call"
