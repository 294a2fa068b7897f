"""``monte_carlo_convergence_test`` — src/pricing_models/validation.py:202-239 — for the GPU pricers.

Same signature and result dictionary.  The callable is priced 4 x n_trials times exactly as the reference does;
with the B200 engine each call is one fused launch, so the whole study takes milliseconds instead of minutes."""

from __future__ import annotations

from typing import Callable

import numpy as np

__all__ = ["monte_carlo_convergence_test"]


def monte_carlo_convergence_test(price_function: Callable[[int], float], n_trials: int = 10, base_sims: int = 10000) -> dict:
    results = {}
    sim_counts = [base_sims, base_sims * 2, base_sims * 4, base_sims * 10]
    for n_sims in sim_counts:
        prices = [price_function(n_sims) for _ in range(n_trials)]
        results[n_sims] = {"mean": np.mean(prices), "std": np.std(prices), "min": np.min(prices), "max": np.max(prices)}
    stds = [results[n]["std"] for n in sim_counts]
    expected_rate = [stds[0] * np.sqrt(sim_counts[0] / n) for n in sim_counts]
    return {"results": results, "stds": stds, "expected_rate": expected_rate,
            "converging": all(s2 <= s1 * 1.5 for s1, s2 in zip(stds[:-1], stds[1:]))}
