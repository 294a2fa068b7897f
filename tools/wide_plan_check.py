#!/usr/bin/env python
"""The planner's shape for the 8 / 14 scenario European launches against pinned alternatives: kernel us (min of 15) per size."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optionslab_b200 import _ffi
P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)
eng = _ffi.get_engine(0)
N_STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 252
KIND = sys.argv[3] if len(sys.argv) > 3 else "european"
spec = {"european": lambda: _ffi.make_spec(_ffi.EUROPEAN, N_STEPS, antithetic=True), "asian": lambda: _ffi.make_spec(_ffi.ASIAN_ARITH, N_STEPS),
        "barrier": lambda: _ffi.make_spec(_ffi.BARRIER, N_STEPS)}[KIND]()
for n_scen in ([int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else (14, 8)):
    params = np.stack([_ffi.make_params(**dict(P, sigma=0.2 + 0.001 * k), barrier=120.0) for k in range(n_scen)]).reshape(1, n_scen)
    for n_paths in (30_000, 100_000, 200_000, 300_000, 500_000, 1_000_000, 2_000_000, 4_000_000) + ((16_000_000,) if "--big" in sys.argv else ()):
        row = {"kind": KIND, "n_scen": n_scen, "n_paths": n_paths, "n_steps": N_STEPS}
        for ppt in (0, 1, 2, 3, 4, 5, 7, 9, 14, 27, 32):
            if ppt and ppt * 256 > n_paths * 2:
                continue
            eng.set_plan(0, ppt)
            eng.simulate(spec, params, 42, n_paths)
            if ppt == 0:
                row["auto_plan"] = eng.last_plan()
            eng.set_kernel_timing(True)
            for _ in range(15):
                eng.simulate(spec, params, 42, n_paths)
            row[f"p{ppt}"] = round(eng.kernel_timing()["min_ms"] * 1e3, 1)
            eng.set_kernel_timing(False)
        eng.set_plan()
        print(json.dumps(row), flush=True)
