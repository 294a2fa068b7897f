"""GPU: the structured products of src/pricing_models/exotic_options.py:404-552 (AutocallableOption, CliquetOption).

Test 1 (FP64, the reference's own draws): per-path payoffs within 1e-12 of the oracle restatement and prices within 1e-12
of the values recorded from the real reference.  Test 2 (fused FP32 path, on-device Philox): same draws evaluated in FP64
by the oracle agree to 5e-4 on the sums; prices sit within 3 combined standard errors of the reference's recorded values.
"""

import numpy as np
import pytest

import optionslab_b200 as ob
from optionslab_b200 import _ffi
from oracle import philox_oracle as po
from oracle import reference_mc as orc

pytestmark = pytest.mark.gpu
P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)
TIGHT = dict(autocall_barrier=1.05, coupon_barrier=0.9, coupon_rate=0.08, ki_barrier=0.75)
WIDE = dict(local_cap=0.08, local_floor=-0.03, global_cap=0.5, global_floor=-0.1)
AUTO_DEFAULT = dict(autocall_barrier=1.0, coupon_barrier=0.8, coupon_rate=0.10, ki_barrier=0.6)
CLIQ_DEFAULT = dict(local_cap=0.05, local_floor=-0.05, global_cap=0.30, global_floor=0.0)
CASES = [(100000, 252, 21, 12), (5000, 12, 3, 4), (20000, 100, 7, 9), (4097, 37, 37, 37)]


def _auto_product(terms, freq):
    return _ffi.Product(terms["autocall_barrier"], terms["coupon_barrier"], terms["coupon_rate"], terms["ki_barrier"], freq, 0)


def _cliq_product(terms, nper):
    return _ffi.Product(terms["local_cap"], terms["local_floor"], terms["global_cap"], terms["global_floor"], nper, 0)


@pytest.mark.parametrize("n_paths,n_steps,freq,nper", CASES)
def test_fp64_parity_on_reference_draws(engine, goldens, anchored, n_paths, n_steps, freq, nper):
    Z = orc.normals_legacy(42, (n_paths, n_steps))
    paths = orc.exotic_paths_from_normals(P["S"], P["T"], P["r"], P["sigma"], 0.0, Z)
    sp, tag = goldens["structured"], f"{n_paths}x{n_steps}"
    for name, terms in (("default", AUTO_DEFAULT), ("tight", TIGHT)):
        want = orc.autocallable_payoffs(paths, P["S"], P["T"], P["r"], observation_freq=freq, **terms)
        got, mom = engine.structured_from_normals(_ffi.make_spec(_ffi.AUTOCALLABLE, n_steps), _auto_product(terms, freq), _ffi.make_params(**P), Z)
        assert np.max(np.abs(got - want)) <= 1e-12
        assert mom["n"] == n_paths and mom["sum"] / mom["n"] == pytest.approx(np.mean(want), rel=1e-12)
        if anchored():
            assert mom["sum"] / mom["n"] == pytest.approx(sp[f"autocallable_{name}_{tag}_f{freq}"], rel=1e-12)
    for name, terms in (("default", CLIQ_DEFAULT), ("wide", WIDE)):
        want = orc.cliquet_payoffs(paths, P["S"], n_periods=nper, **terms)
        got, mom = engine.structured_from_normals(_ffi.make_spec(_ffi.CLIQUET, n_steps), _cliq_product(terms, nper), _ffi.make_params(**P), Z)
        assert np.max(np.abs(got - want)) <= 1e-12 * P["S"]
        price = np.exp(-P["r"] * P["T"]) * mom["sum"] / mom["n"]
        assert price == pytest.approx(np.exp(-P["r"] * P["T"]) * np.mean(want), rel=1e-12)
        if anchored():
            assert price == pytest.approx(sp[f"cliquet_{name}_{tag}_p{nper}"], rel=1e-12)


def test_fp64_parity_with_dividend_and_other_spot(engine, goldens):
    if goldens["numpy"] != np.__version__:
        pytest.skip("goldens recorded with another NumPy build")
    Z = orc.normals_legacy(7, (20000, 64))
    p = _ffi.make_params(105.0, 100.0, 1.5, 0.03, 0.35, 0.02)
    _, mom = engine.structured_from_normals(_ffi.make_spec(_ffi.AUTOCALLABLE, 64), _auto_product(AUTO_DEFAULT, 8), p, Z)
    assert mom["sum"] / mom["n"] == pytest.approx(goldens["structured"]["autocallable_q_sigma_20000x64_f8"], rel=1e-12)
    _, mom = engine.structured_from_normals(_ffi.make_spec(_ffi.CLIQUET, 64), _cliq_product(CLIQ_DEFAULT, 8), p, Z)
    assert np.exp(-0.03 * 1.5) * mom["sum"] / mom["n"] == pytest.approx(goldens["structured"]["cliquet_q_sigma_20000x64_p8"], rel=1e-12)


@pytest.mark.parametrize("n_steps,freq,nper", [(37, 5, 6), (252, 21, 12), (12, 12, 12), (64, 100, 3)])
def test_fused_fp32_path_vs_fp64_oracle_same_stream(engine, n_steps, freq, nper):
    n, seed = 20_000, 13
    Z = po.normals(seed, n, n_steps)
    for q, sigma in ((0.0, 0.2), (0.02, 0.35)):
        p = dict(P, sigma=sigma)
        paths = orc.exotic_paths_from_normals(p["S"], p["T"], p["r"], sigma, q, Z)
        params = _ffi.make_params(**p, q=q).reshape(1, 1)
        for terms in (AUTO_DEFAULT, TIGHT):
            want = orc.autocallable_payoffs(paths, p["S"], p["T"], p["r"], observation_freq=freq, **terms)
            m = engine.simulate_structured(_ffi.make_spec(_ffi.AUTOCALLABLE, n_steps), _auto_product(terms, freq), params, seed, n)[0, 0]
            # a path within FP32 rounding of a barrier may flip: allow a few payoffs' worth of slack
            assert m["n"] == n and m["sum"] == pytest.approx(want.sum(), rel=5e-4, abs=3.0)
        for terms in (CLIQ_DEFAULT, WIDE):
            want = orc.cliquet_payoffs(paths, p["S"], n_periods=nper, **terms)
            m = engine.simulate_structured(_ffi.make_spec(_ffi.CLIQUET, n_steps), _cliq_product(terms, nper), params, seed, n)[0, 0]
            assert m["sum"] == pytest.approx(want.sum(), rel=5e-4) and m["sum_sq"] == pytest.approx((want**2).sum(), rel=1e-3)


def test_on_device_prices_within_three_standard_errors_of_the_reference(goldens):
    sp = goldens["structured"]
    n = 2_000_000
    auto = ob.AutocallableOption(**P, seed=42).price(n, 252, 21, return_error=True)
    ref_se = 0.12 / np.sqrt(100_000)  # payoff std ~ 0.12 of notional: the reference's own 100k-path noise
    assert abs(auto.price - sp["autocallable_default_100000x252_f21"]) <= 3 * np.hypot(auto.std_error, ref_se)
    assert auto.n_paths == n and 0.9 < auto.price < 1.1
    tight = ob.AutocallableOption(**P, seed=42, **TIGHT).price(n, 252, 21, return_error=True)
    assert abs(tight.price - sp["autocallable_tight_100000x252_f21"]) <= 3 * np.hypot(tight.std_error, 0.15 / np.sqrt(100_000))
    cl = ob.CliquetOption(**P, seed=42).price(n, 252, 12, return_error=True)
    assert abs(cl.price - sp["cliquet_default_100000x252_p12"]) <= 3 * np.hypot(cl.std_error, 9.0 / np.sqrt(100_000))
    wide = ob.CliquetOption(**P, seed=42, **WIDE).price(n, 252, 12, return_error=True)
    assert abs(wide.price - sp["cliquet_wide_100000x252_p12"]) <= 3 * np.hypot(wide.std_error, 12.0 / np.sqrt(100_000))
    assert isinstance(ob.CliquetOption(**P, seed=1).price(1000, 12, 4), np.float64)


def test_scenarios_share_draws_and_adapter_greeks(goldens):
    """Fused scenarios == separate re-pricings bit for bit; compute_greeks_unified through ExoticAdapter forwards
    n_periods and lands near the reference's own (noisy, 20k-path) adapter Greeks."""
    cl = ob.CliquetOption(**P, seed=42)
    sc = [(100.0, 100.0, 1.0, 0.05, 0.2, 0.0), (101.0, 100.0, 1.0, 0.05, 0.2, 0.0), (100.0, 100.0, 1.0, 0.0501, 0.21, 0.01)]
    assert cl.price_scenarios(sc, n_paths=30_000, n_steps=36, n_periods=6) == [cl.price_scenarios([s], n_paths=30_000, n_steps=36, n_periods=6)[0] for s in sc]
    au = ob.AutocallableOption(**P, seed=42)
    assert au.price_scenarios(sc, n_paths=30_000, n_steps=36, observation_freq=6) == [au.price_scenarios([s], n_paths=30_000, n_steps=36, observation_freq=6)[0] for s in sc]
    g = ob.compute_greeks_unified(ob.ExoticAdapter(ob.CliquetOption(**P, seed=42), n_paths=400_000, n_steps=36), **P, option_type="call", n_periods=6)
    ref = goldens["structured"]["cliquet_adapter_greeks_20000x36"]
    assert g["price"] == pytest.approx(ref["price"], rel=0.02)       # the reference value carries ~0.6% Monte Carlo noise
    assert g["delta"] == pytest.approx(g["price"] / 100.0, rel=1e-6)  # the payoff is homogeneous of degree 1 in S
    assert abs(g["gamma"]) < 1e-6
    assert g["vega"] == pytest.approx(ref["vega"], rel=0.15) and g["rho"] == pytest.approx(ref["rho"], rel=0.1)


def test_structured_edge_cases(engine):
    # more cliquet periods than steps: every return is 0 (reference: start and end index are both 0)
    assert ob.CliquetOption(**P, seed=1).price(1000, 4, 12) == 0.0
    assert ob.CliquetOption(**P, seed=1, global_floor=0.02).price(1000, 4, 12) == pytest.approx(np.exp(-0.05) * 0.02 * 100.0)
    # no observation date inside the horizon: only the maturity leg (coupon above 0.8, knock-in loss below 0.6)
    a = ob.AutocallableOption(**P, seed=3).price(200_000, 50, 100)
    b = ob.AutocallableOption(**P, seed=3).price(200_000, 50, -5)
    assert a == b and 0.9 < a < np.exp(-0.05) * 1.1
    with pytest.raises(ValueError):
        ob.AutocallableOption(**P, seed=3).price(1000, 50, 0)
    with pytest.raises(ZeroDivisionError):
        ob.CliquetOption(**P, seed=3).price(1000, 50, 0)
    # always redeemed at the first observation: deterministic price
    first = ob.AutocallableOption(**P, seed=3, autocall_barrier=0.0).price(10_000, 12, 3, return_error=True)
    assert first.price == pytest.approx((1 + 0.10 * 0.25 * 1.0) * np.exp(-0.05 * 0.25), rel=1e-6) and first.std_error < 1e-6
    with pytest.raises(ob.MonteCarloError):
        engine.simulate(_ffi.make_spec(_ffi.CLIQUET, 12), _ffi.make_params(**P).reshape(1, 1), 1, 100)
    with pytest.raises(ob.MonteCarloError):
        engine.simulate_structured(_ffi.make_spec(_ffi.BARRIER, 12), _cliq_product(CLIQ_DEFAULT, 4), _ffi.make_params(**P).reshape(1, 1), 1, 100)
