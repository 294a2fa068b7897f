#!/usr/bin/env python
"""Soak check of the fused European path at BASELINE scale: the C5 grid (4096 options x 1M antithetic pairs x 252 steps)
priced with many seeds; z-scores against Black-Scholes pooled over seeds and options.  Detects a price bias of a few
1e-5 relative (RNG mapping, FP32 accumulation, tail truncation) that no single run can see.  GPU only; prints JSON."""
import json
import os
import sys
from math import erf, exp, log, sqrt

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optionslab_b200 import _ffi, runtime  # noqa: E402


def main(n_seeds=24):
    eng = _ffi.get_engine(0)
    K, T = np.meshgrid(np.linspace(60.0, 140.0, 64), np.linspace(1.0 / 12.0, 2.0, 64), indexing="ij")
    K, T = K.ravel(), T.ravel()
    S, r, sigma, n_paths, n_steps = 100.0, 0.05, 0.2, 1_000_000, 252
    cdf = lambda x: 0.5 * (1.0 + erf(x / sqrt(2.0)))
    d2 = np.array([(log(S / k) + (r - 0.5 * sigma**2) * t) / (sigma * sqrt(t)) for k, t in zip(K, T)])
    bs = np.array([S * cdf(d + sigma * sqrt(t)) - k * exp(-r * t) * cdf(d) for d, k, t in zip(d2, K, T)])
    clt = np.array([cdf(d) for d in d2]) * 2 * n_paths >= 2000
    params = _ffi.make_params(S, K, T, r, sigma).reshape(K.size, 1)
    spec = _ffi.make_spec(_ffi.EUROPEAN, n_steps, antithetic=True)
    err = np.empty((n_seeds, K.size))
    z = np.empty((n_seeds, K.size))
    for i in range(n_seeds):
        m = eng.simulate(spec, params, 1000 + 7 * i, n_paths)[:, 0]
        price, se = runtime.discounted_price(m, r, T), runtime.discounted_std_error(m, r, T)
        err[i] = price - bs
        z[i] = np.where(se > 0, (price - bs) / np.maximum(se, 1e-300), 0.0)
    zc = z[:, clt]
    emp_sd = err[:, clt].std(axis=0, ddof=1)                        # true Monte Carlo noise per option, from the seeds
    t_stat = err[:, clt].mean(axis=0) / (emp_sd / np.sqrt(n_seeds))  # Student t with n_seeds - 1 degrees of freedom per option
    rel_bias = err[:, clt].mean(axis=0) / bs[clt]
    out = {
        "workload": f"{n_seeds} seeds x 4096 options x {n_paths} antithetic pairs x {n_steps} steps = {n_seeds * K.size * n_paths * n_steps:.3e} path-steps",
        "options_in_clt_regime": int(clt.sum()),
        "pooled_z": {"mean": float(zc.mean()), "std": float(zc.std()), "expected_abs_mean_below": 4 * float(zc.std()) / sqrt(zc.size),
                     "max_abs": float(np.abs(zc).max()), "frac_abs_gt_2": float((np.abs(zc) > 2).mean())},
        "per_option_t_over_seeds": {"mean": float(t_stat.mean()), "std": float(t_stat.std()), "max_abs": float(np.abs(t_stat).max()),
                                    "note": f"Student t, {n_seeds - 1} dof: std should be ~{sqrt((n_seeds - 1) / (n_seeds - 3)):.3f}, max |t| over 4036 options ~ 4.5-5.5"},
        "relative_price_bias": {"mean": float(rel_bias.mean()), "median_abs": float(np.median(np.abs(rel_bias))),
                                "weighted_mean_by_inverse_variance": float(np.sum(err[:, clt].mean(axis=0) / emp_sd**2) / np.sum(bs[clt] / emp_sd**2))},
    }
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 24)
