#!/usr/bin/env python
"""Where the fixed cost of a small launch goes: kernel time (engine events) and API time (host clock) of European
launches from one CTA upwards, with one step (no work) and 252 steps.  One JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optionslab_b200 import _ffi  # noqa: E402

P = (100.0, 100.0, 1.0, 0.05, 0.2, 0.0)


def main():
    eng = _ffi.get_engine(0)
    rows = []
    for n_paths in (256, 256 * 37, 256 * 148, 256 * 148 * 2, 100_000, 256 * 148 * 6):
        for n_steps in (1, 252):
            spec = _ffi.make_spec(_ffi.EUROPEAN, n_steps, antithetic=True)
            eng.set_plan(0, 1)
            call = lambda: eng.simulate_scalars(spec, [P], 42, n_paths)
            call()
            eng.set_kernel_timing(True)
            t = []
            for _ in range(200):
                t0 = time.perf_counter()
                call()
                t.append(time.perf_counter() - t0)
            kt = eng.kernel_timing()
            eng.set_kernel_timing(False)
            t.sort()
            rows.append({"n_paths": n_paths, "n_steps": n_steps, "ctas": eng.last_plan()["tiles"], "kernel_us_min": round(kt["min_ms"] * 1e3, 2),
                         "kernel_us_mean": round(kt["mean_ms"] * 1e3, 2), "api_us_min": round(t[0] * 1e6, 2), "api_us_median": round(t[100] * 1e6, 2)})
    eng.set_plan()
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
