#!/usr/bin/env python
"""delta_gamma at the reference's default bump h = 1e-4 (monte_carlo_unified.py:522) against h = 1.0 and Black-Scholes,
over the spots of VERDICT r01 weak #1.  One JSON line."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import optionslab_b200 as ob  # noqa: E402


def bs_delta(S, K, T, r, sigma, call=True):
    d1 = (math.log(S / K) + (r + 0.5 * sigma * sigma) * T) / (sigma * math.sqrt(T))
    cdf = 0.5 * (1.0 + math.erf(d1 / math.sqrt(2.0)))
    return cdf if call else cdf - 1.0


def main():
    rows = []
    for steps in (1, 100):
        pr = ob.MonteCarloPricerUni(2_000_000, steps, seed=42)
        for S in (90.0, 97.3, 100.0, 103.7, 110.0):
            for typ in ("call", "put"):
                d_small, g_small = pr.delta_gamma(S, 100.0, 1.0, 0.05, 0.2, typ, seed=11)
                d_big, _ = pr.delta_gamma(S, 100.0, 1.0, 0.05, 0.2, typ, h=1.0, seed=11)
                rows.append({"steps": steps, "S": S, "type": typ, "delta_h1e-4": d_small, "gamma_h1e-4": g_small, "delta_h1": d_big,
                             "bs_delta": bs_delta(S, 100.0, 1.0, 0.05, 0.2, typ == "call")})
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
