#!/usr/bin/env python
"""A few small launches of every kernel family (a quick whole-library probe; also the workload for compute-sanitizer where that is allowed):
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py
Sizes are tiny (the tools slow kernels down 10-100x) but cover: one tile / many tiles / more than one fold group, 1-16
scenarios, the lane-split kernel, control variate, path-dependent kinds, QMC with a dimension split, Heston, jumps, structured
products, the FP64 parity kernels and the in-kernel exchange between two engines of one device."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import optionslab_b200 as ob  # noqa: E402
from optionslab_b200 import _ffi  # noqa: E402

P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)


def main():
    eng = _ffi.get_engine(0)
    one = _ffi.make_params(**P).reshape(1, 1)
    many = np.stack([_ffi.make_params(**dict(P, sigma=0.2 + 0.01 * k)) for k in range(14)]).reshape(1, 14)
    grid = _ffi.make_params(100.0, np.linspace(80, 120, 9), 1.0, 0.05, 0.2).reshape(9, 1)
    euro = _ffi.make_spec(_ffi.EUROPEAN, 24, antithetic=True)
    out = []
    out.append(eng.simulate(euro, one, 1, 100)[0, 0]["sum"])                        # one tile
    out.append(eng.simulate(euro, one, 1, 5_000)[0, 0]["sum"])                      # lane split, several tiles
    eng.set_plan(0, 1)
    out.append(eng.simulate(euro, one, 1, 300_000)[0, 0]["sum"])                    # 1172 tiles: two fold groups
    out.append(eng.simulate(euro, many, 1, 270_000)[0, 3]["sum"])                   # 16-scenario kernel, two fold groups
    eng.set_plan()
    out.append(eng.simulate(euro, many, 1, 20_000)[0, 13]["sum"])
    out.append(eng.simulate(euro, grid, 1, 20_000)[8, 0]["sum"])                    # batch: parameters through HBM
    out.append(eng.simulate(euro, one, 1, 20_000, control_variate=True)[0, 0]["sum_payoff_terminal"])
    for kind in (_ffi.ASIAN_ARITH, _ffi.ASIAN_GEOM, _ffi.BARRIER, _ffi.LOOKBACK):
        spec = _ffi.make_spec(kind, 33)
        params = np.stack([_ffi.make_params(**dict(P, S=100.0 + k), barrier=120.0) for k in range(3)]).reshape(1, 3)
        out.append(eng.simulate(spec, params, 2, 30_000)[0, 1]["sum"])
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out.append(ob.MonteCarloPricer(1 << 16, 130, seed=5, method=ob.MCMethod.QMC).price(**P, option_type="call"))  # QMC, dimension split
    out.append(ob.HestonPricer(kappa=2.0, theta=0.04, sigma_v=0.3, rho=-0.7, v0=0.04).price_monte_carlo(100.0, 100.0, 1.0, 0.05, 0.0, "call", 20_000, 16, seed=1))
    out.append(ob.MertonJumpDiffusion(lambda_j=1.0, mu_j=-0.1, sigma_j=0.15).price_monte_carlo(100.0, 100.0, 1.0, 0.05, 0.2, "call", 0.0, 20_000, 16, seed=1))
    out.append(ob.KouJumpDiffusion(2.0, 0.4, 10.0, 5.0).price_monte_carlo(100.0, 100.0, 1.0, 0.05, 0.2, "call", 0.0, 20_000, 16, seed=1))
    out.append(float(ob.AutocallableOption(**P, seed=8).price(20_000, 24, 6)))
    out.append(float(ob.CliquetOption(**P, seed=8).price(20_000, 24, 6)))
    Z = np.random.default_rng(0).standard_normal((3000, 16))
    for kind, anti in ((_ffi.EUROPEAN, True), (_ffi.ASIAN_ARITH, False), (_ffi.BARRIER, False)):
        out.append(eng.payoffs_from_normals(_ffi.make_spec(kind, 16, antithetic=anti), _ffi.make_params(**P, barrier=120.0), Z)[1]["sum"])
    out.append(float(eng.rng_statistics(3, 4096, 16)["moments"][1]))
    a, b = _ffi.Engine(0), _ffi.Engine(0)
    _ffi.connect_local([a, b])
    with ThreadPoolExecutor(max_workers=2) as pool:
        for rep in range(2):
            got = list(pool.map(lambda i: (a, b)[i].simulate(euro, grid, 4, 10_000, path_begin=10_000 * i, allreduce=True), range(2)))
            assert got[0].tobytes() == got[1].tobytes()
            out.append(got[0][0, 0]["sum"])
    a.close(), b.close()
    print("probe ok:", " ".join(f"{float(x):.6g}" for x in out))


if __name__ == "__main__":
    main()
