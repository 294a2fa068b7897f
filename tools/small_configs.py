#!/usr/bin/env python
"""C1 / C2 sized calls (BASELINE.json configs[0], configs[1]) for ncu captures and API latency:
C1 = MonteCarloPricer(100k, 252).price, C2 = the 14-scenario Greeks launch at 1M x 252.
Prints one JSON line with kernel and API times (best of N); under ncu only the launches matter."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import optionslab_b200 as ob  # noqa: E402
from optionslab_b200 import _ffi  # noqa: E402

P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)


def best(fn, reps):
    fn()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    t.sort()
    return t[0], t[len(t) // 2]


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    eng = _ffi.get_engine(0)
    out = {}
    c1 = ob.MonteCarloPricer(100_000, 252, seed=42)
    c1d = ob.MonteCarloPricer(100_000, seed=42)  # the reference's default: one exact step
    c2 = ob.MonteCarloPricer(1_000_000, 252, seed=42)
    uni = ob.MonteCarloPricerUni(100_000, 100, seed=42)
    for name, fn, n in (("C1 price 100k x 252", lambda: c1.price(**P, option_type="call"), reps),
                        ("default price 100k x 1", lambda: c1d.price(**P, option_type="call"), reps),
                        ("Uni default price 100k x 100", lambda: uni.price(**P, option_type="call"), reps),
                        ("Uni delta_gamma 100k x 100", lambda: uni.delta_gamma(**P, option_type="call", seed=7), reps),
                        ("C2 greeks 1M x 252 (14 scenarios)", lambda: c2.greeks(**P, option_type="call"), max(reps // 4, 3))):
        eng.set_kernel_timing(True)
        lo, med = best(fn, n)
        kt = eng.kernel_timing()
        eng.set_kernel_timing(False)
        out[name] = {"api_us_min": lo * 1e6, "api_us_median": med * 1e6, "kernel_us_min": kt["min_ms"] * 1e3, "kernel_us_mean": kt["mean_ms"] * 1e3}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
