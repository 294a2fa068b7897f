import sys; sys.path.insert(0, '.')
import numpy as np
from optionslab_b200 import _ffi, runtime
eng = _ffi.get_engine(0)
for sigma in (0.1, 0.2, 0.3, 0.45):
    p = _ffi.make_params(100.0, 100.0, 1.0, 0.05, sigma).reshape(1, 1)
    a = eng.simulate(_ffi.make_spec(_ffi.ASIAN_ARITH, 252), p, 5, 16_000_000)[0, 0]
    b = eng.simulate(_ffi.make_spec(_ffi.ASIAN_ARITH, 252, exact_ex2=True), p, 5, 16_000_000)[0, 0]
    pa, pb = runtime.discounted_price(a, 0.05, 1.0), runtime.discounted_price(b, 0.05, 1.0)
    se = runtime.discounted_std_error(a, 0.05, 1.0)
    print(f"sigma={sigma}: small-move {pa:.8f}  ex2 {pb:.8f}  rel diff {(pa-pb)/pb:+.2e}  (one standard error = {se/pa:.1e} relative)")
