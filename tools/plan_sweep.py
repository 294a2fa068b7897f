#!/usr/bin/env python
"""Kernel time (CUDA events inside the engine, best of N) of small European launches over every pinned tile shape
(split_shift x paths_per_thread) next to the automatic plan's choice: the data the cost model in plan_tiles (engine.cu)
is calibrated on.  One JSON line per problem size."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optionslab_b200 import _ffi  # noqa: E402

P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)


def kernel_us(eng, spec, params, n_paths, reps=30):
    eng.simulate(spec, params, 42, n_paths)
    eng.set_kernel_timing(True)
    for _ in range(reps):
        eng.simulate(spec, params, 42, n_paths)
    kt = eng.kernel_timing()
    eng.set_kernel_timing(False)
    return kt["min_ms"] * 1e3


def main():
    eng = _ffi.get_engine(0)
    for n_paths, n_steps, n_scen in ((100_000, 252, 1), (10_000, 252, 1), (30_000, 100, 1), (100_000, 100, 3), (100_000, 252, 14),
                                     (1_000_000, 252, 14), (1_000_000, 252, 1), (300_000, 252, 1), (100_000, 1, 1)):
        spec = _ffi.make_spec(_ffi.EUROPEAN, n_steps, antithetic=True)
        params = np.stack([_ffi.make_params(**dict(P, sigma=0.2 + 0.001 * k)) for k in range(n_scen)]).reshape(1, n_scen)
        eng.set_plan()
        auto = kernel_us(eng, spec, params, n_paths)
        row = {"n_paths": n_paths, "n_steps": n_steps, "n_scen": n_scen, "auto_us": auto, "auto_plan": eng.last_plan(), "grid": {}}
        for shift in (0, 1, 2, 3):
            for ppt in (1, 2, 3, 4, 6, 8, 12, 16):
                if n_paths * (1 << shift) / (256 * ppt) < 20:
                    continue
                eng.set_plan(shift, ppt)
                t = kernel_us(eng, spec, params, n_paths, reps=12)
                lp = eng.last_plan()
                if lp["split_shift"] != shift:
                    continue
                row["grid"][f"s{shift}p{ppt}"] = round(t, 2)
        eng.set_plan()
        best = min(row["grid"], key=row["grid"].get)
        row["best"] = [best, row["grid"][best]]
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
