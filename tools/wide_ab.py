#!/usr/bin/env python
"""A/B of the multi-scenario European launches (C2-like) - kernel time over pinned paths-per-thread, for one library build."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optionslab_b200 import _ffi
P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)
eng = _ffi.get_engine(0)
spec = _ffi.make_spec(_ffi.EUROPEAN, 252, antithetic=True)
out = {"lib": os.path.basename(_ffi.LIB_PATH)}
for n_scen, n_paths in ((14, 1_000_000), (8, 1_000_000), (3, 1_000_000), (14, 100_000)):
    params = np.stack([_ffi.make_params(**dict(P, sigma=0.2 + 0.001 * k)) for k in range(n_scen)]).reshape(1, n_scen)
    row = {}
    for ppt in (0, 1, 2, 3, 4, 6, 8, 9, 14):
        eng.set_plan(0, ppt)
        eng.simulate(spec, params, 42, n_paths)
        eng.set_kernel_timing(True)
        for _ in range(15):
            eng.simulate(spec, params, 42, n_paths)
        row[f"p{ppt}"] = round(eng.kernel_timing()["min_ms"] * 1e3, 1)
        eng.set_kernel_timing(False)
    eng.set_plan()
    out[f"{n_scen}x{n_paths}"] = row
print(json.dumps(out))
