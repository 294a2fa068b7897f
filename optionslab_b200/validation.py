"""``monte_carlo_convergence_test`` for the GPU pricers — the contract of src/pricing_models/validation.py:202-239.

Same signature and the same result dictionary (keys ``results`` / ``stds`` / ``expected_rate`` / ``converging``).  The
study prices ``price_function(n)`` ``n_trials`` times at n = 1x, 2x, 4x and 10x ``base_sims`` and checks that the spread of
the estimates never grows by more than 50 % from one size to the next; ``expected_rate`` is the 1/sqrt(n) law anchored at
the first size.  With the B200 engine every call is one fused launch, so the whole study takes milliseconds."""

from __future__ import annotations

from typing import Callable, Dict, List

import numpy as np

__all__ = ["monte_carlo_convergence_test"]

_SIZE_MULTIPLIERS = (1, 2, 4, 10)  # validation.py:218
_MAX_GROWTH = 1.5                  # validation.py:238


def _spread(estimates: List[float]) -> Dict[str, float]:
    a = np.asarray(estimates, dtype=np.float64)
    return {"mean": a.mean(), "std": a.std(), "min": a.min(), "max": a.max()}  # population std, as np.std defaults to


def monte_carlo_convergence_test(price_function: Callable[[int], float], n_trials: int = 10, base_sims: int = 10000) -> dict:
    sizes = [base_sims * m for m in _SIZE_MULTIPLIERS]
    per_size = {n: _spread([price_function(n) for _ in range(n_trials)]) for n in sizes}
    stds = [per_size[n]["std"] for n in sizes]
    law = [stds[0] * np.sqrt(sizes[0] / n) for n in sizes]
    ok = all(later <= earlier * _MAX_GROWTH for earlier, later in zip(stds, stds[1:]))
    return {"results": per_size, "stds": stds, "expected_rate": law, "converging": ok}
