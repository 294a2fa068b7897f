"""Host-side dispatch shared by all pricer classes: build the parameter block, run the fused
simulation on this rank's path range, combine moments across ranks, turn moments into prices."""

from __future__ import annotations

import math
import os
from typing import Optional, Sequence

import numpy as np

from . import _ffi, distributed
from .exceptions import MonteCarloError


def fresh_seed() -> int:
    """Seed for ``seed=None`` constructors: drawn once, so an instance stays deterministic
    (src/pricing_models/monte_carlo.py:68-70, monte_carlo_unified.py:284-286)."""
    return int(np.random.default_rng().integers(0, 2**31))


def entropy_seed() -> int:
    """Unseeded exotic pricing (exotic_options.py:51-52 leaves the global state untouched)."""
    return int.from_bytes(os.urandom(8), "little")


def simulate(spec: _ffi.Spec, params: np.ndarray, seed: int, n_paths: int, *, stream_base: int = 0) -> np.ndarray:
    """Moments [n_opt, n_scen] over ALL ``n_paths`` global paths (sharded + all-reduced when a
    distributed context is active)."""
    if n_paths < 1:
        raise MonteCarloError("n_paths must be >= 1")
    eng = _ffi.get_engine()
    ctx = distributed.current()
    if ctx is None or ctx.world_size == 1:
        return eng.simulate(spec, params, seed, n_paths, stream_base=stream_base)
    begin, count = distributed.partition_paths(n_paths, ctx.rank, ctx.world_size)
    if count > 0:
        local = eng.simulate(spec, params, seed, count, stream_base=stream_base, path_begin=begin)
    else:
        local = np.zeros(np.shape(params), dtype=_ffi.MOMENTS_DTYPE)
    return distributed.allreduce_moments(local, ctx)


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
