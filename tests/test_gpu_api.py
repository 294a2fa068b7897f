"""GPU: the reference's own behavioural tests (tests/test_monte_carlo.py, tests/test_exotic_options.py),
re-run against the drop-in classes with the reference's fixtures and tolerances."""

import numpy as np
import pytest

import optionslab_b200 as ob
from optionslab_b200 import _ffi
from oracle import reference_mc as orc

pytestmark = pytest.mark.gpu
P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)


@pytest.fixture
def basic_pricer():
    return ob.MonteCarloPricer(num_simulations=10000, num_steps=50, seed=42)


@pytest.fixture
def unified_pricer():
    return ob.MonteCarloPricerUni(num_simulations=10000, num_steps=50, seed=42, use_numba=False, use_gpu=False)


# ---- tests/test_monte_carlo.py:119-168 ----
def test_price_call_and_put_close_to_black_scholes(basic_pricer):
    for ot in ("call", "put"):
        price = basic_pricer.price(**P, option_type=ot)
        assert isinstance(price, float) and price > 0
        assert abs(price - orc.black_scholes(**P, option_type=ot)) < 1.0


def test_price_with_dividend_and_reproducibility(basic_pricer):
    assert basic_pricer.price(**P, option_type="call", q=0.02) < basic_pricer.price(**P, option_type="call", q=0.0)
    assert basic_pricer.price(100, 100, 1.0, 0.05, 0.2, "call") == basic_pricer.price(100, 100, 1.0, 0.05, 0.2, "call")
    res = basic_pricer.price(**P, option_type="call", return_error=True)
    assert isinstance(res, ob.MCResult) and 0 < res.std_error < res.price and res.n_paths == 20000


# ---- tests/test_monte_carlo.py:392-463 ----
def test_unified_price_delta_gamma_and_batch(unified_pricer):
    for ot in ("call", "put"):
        price = unified_pricer.price(**P, option_type=ot)
        assert price > 0 and abs(price - orc.black_scholes(**P, option_type=ot)) < 1.5
    delta, gamma = unified_pricer.delta_gamma(**P, option_type="call", h=1.0)
    assert 0 < delta < 1 and gamma > 0
    d2, _ = unified_pricer.delta_gamma(**P, option_type="call")  # default h = 1e-4 (gamma is noise there)
    assert 0 < d2 < 1
    S = np.array([100.0, 110.0, 90.0, 100.0, 100.0]); K = np.array([100.0, 100.0, 100.0, 95.0, 105.0])
    T = np.array([1.0, 1.0, 1.0, 0.5, 0.5]); r = np.full(5, 0.05); s = np.array([0.2, 0.2, 0.2, 0.3, 0.15])
    q = np.array([0.0, 0.0, 0.0, 0.02, 0.01])
    prices = unified_pricer.price_batch(S, K, T, r, s, "call", q)
    assert prices.shape == (5,) and np.all(prices > 0)
    for i in range(5):
        assert abs(prices[i] - orc.black_scholes(S[i], K[i], T[i], r[i], s[i], "call", q[i])) < 1.5
    assert unified_pricer.price_batch(S[:1], K[:1], T[:1], r[:1], s[:1], "call")[0] == unified_pricer.price(**P, option_type="call")
    deltas, gammas = unified_pricer.delta_gamma_batch(S, K, T, r, s, "call", q, h=1.0)
    assert deltas.shape == gammas.shape == (5,) and np.all((deltas > 0) & (deltas < 1)) and np.all(gammas > 0)
    assert unified_pricer.price_batch([], [], [], [], [], "call").shape == (0,)
    assert unified_pricer.price_batch(S, K, T, r, s, "put", 0.01).shape == (5,)  # scalar q broadcast


# ---- tests/test_monte_carlo.py:506-546 ----
def test_put_call_parity_and_monotonicities(basic_pricer):
    S, K, T, r, sigma, q = 100, 100, 1.0, 0.05, 0.2, 0.02
    call = basic_pricer.price(S, K, T, r, sigma, "call", q)
    put = basic_pricer.price(S, K, T, r, sigma, "put", q)
    assert abs((call - put) - (S * np.exp(-q * T) - K * np.exp(-r * T))) < 2.0
    itm, atm, otm = (basic_pricer.price(S_, 100, 1.0, 0.05, 0.2, "call") for S_ in (110, 100, 90))
    assert itm > atm > otm
    assert basic_pricer.price(100, 100, 1.0, 0.05, 0.4, "call") > basic_pricer.price(100, 100, 1.0, 0.05, 0.1, "call")
    assert basic_pricer.price(100, 100, 2.0, 0.05, 0.2, "call") > basic_pricer.price(100, 100, 0.25, 0.05, 0.2, "call")


# ---- tests/test_exotic_options.py:53-193 ----
def test_asian_reference_assertions():
    a = ob.AsianOption(**P, seed=42)
    price = a.price(n_paths=50000, n_steps=100)
    assert isinstance(price, np.float64) and 0 < price < orc.black_scholes(**P)
    geo_mc = a.price(n_paths=100000, n_steps=252, avg_type="geometric")
    assert abs(geo_mc - a.price_geometric_closed_form()) / a.price_geometric_closed_form() < 0.05
    assert a.price(50000, 100, "arithmetic", "put") > 0
    assert ob.price_asian(100, 100, 1.0, 0.05, 0.2, seed=42, n_paths=50000) == ob.price_asian(100, 100, 1.0, 0.05, 0.2, seed=42, n_paths=50000)


def test_barrier_reference_assertions():
    b = ob.BarrierOption(**P, seed=42, barrier=120.0)
    euro = orc.black_scholes(**P)
    out = b.price(n_paths=100000, n_steps=252, barrier_type="up-and-out")
    inn = b.price(n_paths=100000, n_steps=252, barrier_type="up-and-in")
    assert 0 <= out < euro and inn >= 0
    assert abs((out + inn) - euro) / euro < 0.10
    for bt in ("down-and-out", "down-and-in"):
        assert ob.BarrierOption(**P, seed=42, barrier=85.0).price(50000, 100, bt, "put") >= 0
    assert ob.price_barrier(100, 100, 1.0, 0.05, 0.2, barrier=120, seed=1, n_paths=20000) >= 0
    knocked = ob.BarrierOption(**P, seed=1, barrier=99.0).price(10000, 10, "up-and-out")  # S0 >= B: dead at t=0
    assert knocked == 0.0


def test_lookback_dominates_vanilla_and_adapter_greeks():
    lb = ob.LookbackOption(**P, seed=42)
    assert lb.price(100000, 252, "floating", "call") > orc.black_scholes(**P)
    assert lb.price(100000, 252, "fixed", "call") > orc.black_scholes(**P)
    assert lb.price(100000, 252, "floating", "put") > 0 and lb.price(100000, 252, "fixed", "put") > 0
    ad = ob.ExoticAdapter(ob.AsianOption(**P, seed=42), n_paths=200000, n_steps=64, avg_type="arithmetic")
    g = ob.compute_greeks_unified(ad, **P, option_type="call")
    assert 0.4 < g["delta"] < 0.7 and g["gamma"] > 0 and g["vega"] > 0
    # the adapter's plain price() route (what the REFERENCE's compute_greeks_unified would call) prices the same draws:
    # equal to FP32 summation order, and bit for bit once both launches cut the paths into the same tiles
    assert ad.price(**P, option_type="call") == pytest.approx(g["price"], rel=1e-6)
    eng = _ffi.get_engine(0)
    eng.set_plan(0, 4)
    try:
        assert ad.price(**P, option_type="call") == ob.compute_greeks_unified(ad, **P, option_type="call")["price"]
    finally:
        eng.set_plan()
    gb = ob.compute_greeks_unified(ob.ExoticAdapter(ob.BarrierOption(**P, seed=42, barrier=130.0), 200000, 64,
                                                    barrier_type="up-and-out"), **P, option_type="call")
    assert gb["price"] > 0 and np.isfinite(list(gb.values())).all()


def test_engine_error_paths(engine):
    from optionslab_b200 import _ffi
    with pytest.raises(ob.MonteCarloError):
        engine.simulate(_ffi.make_spec(_ffi.EUROPEAN, 0), _ffi.make_params(**P).reshape(1, 1), 1, 10)
    with pytest.raises(ob.MonteCarloError):
        engine.simulate(_ffi.make_spec(_ffi.EUROPEAN, 4), _ffi.make_params(**P).reshape(1, 1), 1, 0)
    with pytest.raises(ob.MonteCarloError):
        engine.simulate(_ffi.make_spec(9, 4), _ffi.make_params(**P).reshape(1, 1), 1, 10)
    with pytest.raises(ob.MonteCarloError):
        engine.simulate(_ffi.make_spec(_ffi.EUROPEAN, 4), np.zeros((1, 17), dtype=_ffi.PARAMS_DTYPE), 1, 10)
    info = engine.info()
    assert info["cc_major"] >= 10 and info["sm_count"] > 0


def test_monte_carlo_convergence_test_with_fresh_seeds():
    """src/pricing_models/validation.py:202-239 driven by the GPU pricer: the spread of repeated estimates shrinks
    like 1/sqrt(N) (each trial uses a fresh seed, as a seed=None reference pricer would)."""
    import optionslab_b200 as ob

    def price(n_sims):
        return ob.MonteCarloPricer(n_sims, 16).price(100.0, 100.0, 1.0, 0.05, 0.2, "call")

    out = ob.monte_carlo_convergence_test(price, n_trials=12, base_sims=20000)
    assert out["converging"]
    assert list(out["results"]) == [20000, 40000, 80000, 200000]
    assert out["stds"][-1] < out["stds"][0]
    assert 0.3 < out["stds"][-1] / out["expected_rate"][-1] < 3.0
    assert all(abs(v["mean"] - 10.4506) < 0.15 for v in out["results"].values())


def test_pricers_are_thread_safe_on_one_engine():
    """The engine serialises calls per handle (one mutex), ctypes releases the GIL: pricers used from several Python threads at
    once must return exactly what they return one after the other (the reference's objects are single-threaded; callers such
    as a Streamlit server are not)."""
    from concurrent.futures import ThreadPoolExecutor

    import optionslab_b200 as ob

    P_ = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)

    def job(i):
        if i % 4 == 0:
            return ob.MonteCarloPricer(50_000 + i, 16, seed=i).price(**P_, option_type="call")
        if i % 4 == 1:
            return float(ob.AsianOption(**P_, seed=i).price(30_000 + i, 12))
        if i % 4 == 2:
            return ob.MonteCarloPricerUni(20_000 + i, 8, seed=i).price_batch([100.0, 95.0], [100.0, 100.0], [1.0, 0.5], [0.05, 0.05], [0.2, 0.3], "put").tolist()
        return float(ob.CliquetOption(**P_, seed=i).price(20_000 + i, 12, 4))

    sequential = [job(i) for i in range(24)]
    with ThreadPoolExecutor(max_workers=8) as pool:
        concurrent = list(pool.map(job, range(24)))
    assert concurrent == sequential
