"""GPU, correctness test 1 of BASELINE.json: fed the reference's own NumPy normal draws, the engine
reproduces the reference's per-path payoffs and prices in FP64 within 1e-12 relative.

"Relative" is applied as SURVEY.md §7 derives it: |d payoff| <= 1e-12 * max(S_T, K)-scale per path
(a call payoff near the money is a cancelled difference), and |d price| <= 1e-12 * price.
The oracle (oracle/reference_mc.py) is pinned bit-for-bit to the real reference by
tests/test_oracle_golden.py; the goldens from the real reference are also compared directly.
"""

import numpy as np
import pytest

from optionslab_b200 import _ffi
from oracle import reference_mc as orc

pytestmark = pytest.mark.gpu
TOL = 1e-12
P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)


def _check(got, want, scale):
    assert got.shape == want.shape
    err = np.max(np.abs(got - want) / scale)
    assert err <= TOL, f"max per-path error {err:.3e}"


def _moments_ok(mom, pay):
    assert mom["n"] == len(pay)
    assert mom["sum"] == pytest.approx(float(np.sum(pay)), rel=TOL)
    assert mom["sum_sq"] == pytest.approx(float(np.sum(pay * pay)), rel=TOL)


@pytest.mark.parametrize("n_paths,n_steps", [(100_000, 252), (10_000, 50), (4097, 7), (1, 1), (129, 33), (5000, 365)])
@pytest.mark.parametrize("ot", ["call", "put"])
def test_european_from_reference_draws(engine, n_paths, n_steps, ot):
    q = 0.01
    Z = orc.normals_generator(42, (n_paths, n_steps))
    terminal = orc.gbm_terminal_from_normals(P["S"], P["T"], P["r"], P["sigma"], q, Z)
    want = orc.vanilla_payoffs(terminal, P["K"], ot)
    spec = _ffi.make_spec(_ffi.EUROPEAN, n_steps, is_put=(ot == "put"), antithetic=True)
    got, mom = engine.payoffs_from_normals(spec, _ffi.make_params(**P, q=q), Z)
    _check(got, want, np.maximum(terminal, P["K"]))
    _moments_ok(mom, want)
    price = float(np.exp(-P["r"] * P["T"]) * mom["sum"] / mom["n"])
    assert price == pytest.approx(orc.discounted_mean(want, P["r"], P["T"]), rel=TOL)


def test_european_price_and_error_match_real_reference_goldens(engine, goldens):
    if goldens["numpy"] != np.__version__:
        pytest.skip("goldens recorded with another NumPy build")
    from optionslab_b200 import runtime
    for tag, n_sims, n_steps in [("10000x50", 10000, 50), ("100000x252", 100000, 252), ("4096x7", 4096, 7)]:
        for ot in ("call", "put"):
            g = goldens["european"][f"{tag}_{ot}"]
            Z = orc.normals_generator(42, (n_sims, n_steps))
            spec = _ffi.make_spec(_ffi.EUROPEAN, n_steps, is_put=(ot == "put"), antithetic=True)
            pay, mom = engine.payoffs_from_normals(spec, _ffi.make_params(**P), Z)
            assert float(runtime.discounted_price(mom, P["r"], P["T"])) == pytest.approx(g["price"], rel=TOL)
            assert float(runtime.discounted_std_error(mom, P["r"], P["T"])) == pytest.approx(g["std_error"], rel=1e-9)
            assert int(mom["n"]) == g["n_paths"]
            np.testing.assert_allclose(pay[:8], g["payoff_head"], rtol=0, atol=TOL * 200)
            np.testing.assert_allclose(pay[n_sims:n_sims + 8], g["payoff_mirror_head"], rtol=0, atol=TOL * 200)


def test_single_step_european_matches_reference_fast_path(engine, goldens, anchored):
    Z = orc.normals_generator(42, 100_000)
    terminal = orc.gbm_terminal_single_step(P["S"], P["T"], P["r"], P["sigma"], 0.0, Z)
    want = orc.vanilla_payoffs(terminal, P["K"], "call")
    spec = _ffi.make_spec(_ffi.EUROPEAN, 1, antithetic=True)
    got, mom = engine.payoffs_from_normals(spec, _ffi.make_params(**P), Z.reshape(-1, 1))
    _check(got, want, np.maximum(terminal, P["K"]))
    if anchored():
        price = float(np.exp(-P["r"] * P["T"]) * mom["sum"] / mom["n"])
        assert price == pytest.approx(goldens["european"]["100000x1_call"]["price"], rel=TOL)


def test_uni_cumsum_form_from_reference_draws(engine, goldens, anchored):
    """monte_carlo_unified.py:333-343 accumulates increments; one option per call in parity mode."""
    b = goldens["uni"]["batch_inputs"]
    n_opt, N, n = len(b["S"]), b["num_simulations"], b["num_steps"]
    Z = orc.normals_generator(42, (n_opt, N, n))
    S, K, T, r, s, q = (np.array(b[k]) for k in ("S", "K", "T", "r", "sigma", "q"))
    terminal = orc.uni_terminal_from_normals(S, T, r, s, q, Z)
    prices = []
    for i in range(n_opt):
        want = np.maximum(terminal[i] - K[i], 0.0)
        spec = _ffi.make_spec(_ffi.EUROPEAN, n, antithetic=True)
        got, mom = engine.payoffs_from_normals(spec, _ffi.make_params(S[i], K[i], T[i], r[i], s[i], q[i]), Z[i], accumulate=True)
        _check(got, want, np.maximum(terminal[i], K[i]))
        prices.append(np.exp(-r[i] * T[i]) * mom["sum"] / mom["n"])
    if anchored():
        np.testing.assert_allclose(prices, goldens["uni"]["price_batch_call"], rtol=TOL)


@pytest.mark.parametrize("n_paths,n_steps", [(100_000, 252), (5000, 12), (1031, 5)])
def test_asian_from_reference_draws(engine, goldens, anchored, n_paths, n_steps):
    Z = orc.normals_legacy(42, (n_paths, n_steps))
    paths = orc.exotic_paths_from_normals(P["S"], P["T"], P["r"], P["sigma"], 0.0, Z)
    for kind, avg in [(_ffi.ASIAN_ARITH, "arithmetic"), (_ffi.ASIAN_GEOM, "geometric")]:
        for ot in ("call", "put"):
            want = orc.asian_payoffs(paths, P["K"], avg, ot)
            spec = _ffi.make_spec(kind, n_steps, is_put=(ot == "put"))
            got, mom = engine.payoffs_from_normals(spec, _ffi.make_params(**P), Z)
            _check(got, want, np.maximum(paths.mean(axis=1), P["K"]))
            _moments_ok(mom, want)
            key = f"asian_{'arith' if avg == 'arithmetic' else 'geom'}_{ot}_{n_paths}x{n_steps}"
            if key in goldens["exotics"] and anchored():
                price = float(np.exp(-P["r"] * P["T"]) * mom["sum"] / mom["n"])
                assert price == pytest.approx(goldens["exotics"][key], rel=TOL)


@pytest.mark.parametrize("n_paths,n_steps", [(100_000, 365), (5000, 12)])
def test_barrier_from_reference_draws(engine, goldens, anchored, n_paths, n_steps):
    Z = orc.normals_legacy(42, (n_paths, n_steps))
    paths = orc.exotic_paths_from_normals(P["S"], P["T"], P["r"], P["sigma"], 0.0, Z)
    for B, kinds in [(120.0, ("up-and-out", "up-and-in")), (85.0, ("down-and-out", "down-and-in")), (99.0, ("up-and-out", "up-and-in")),
                     (101.0, ("down-and-out", "down-and-in"))]:  # last two: knocked at t=0 (column 0 is monitored)
        for bt in kinds:
            for ot in ("call", "put"):
                want = orc.barrier_payoffs(paths, P["K"], B, bt, ot)
                spec = _ffi.make_spec(_ffi.BARRIER, n_steps, is_put=(ot == "put"), barrier_down=bt.startswith("down"),
                                      barrier_in=bt.endswith("in"))
                got, mom = engine.payoffs_from_normals(spec, _ffi.make_params(**P, barrier=B), Z)
                # a knock decision can only differ for a path within rounding of the barrier: none expected
                assert np.count_nonzero((got == 0) != (want == 0)) == 0
                _check(got, want, np.maximum(paths[:, -1], P["K"]))
                _moments_ok(mom, want)
                key = f"barrier_{bt}_{ot}_B{int(B)}_{n_paths}x{n_steps}"
                if key in goldens["exotics"] and anchored():
                    price = float(np.exp(-P["r"] * P["T"]) * mom["sum"] / mom["n"])
                    assert price == pytest.approx(goldens["exotics"][key], rel=TOL, abs=1e-15)


def test_lookback_from_reference_draws(engine, goldens, anchored):
    n_paths, n_steps = 5000, 12
    Z = orc.normals_legacy(42, (n_paths, n_steps))
    paths = orc.exotic_paths_from_normals(P["S"], P["T"], P["r"], P["sigma"], 0.0, Z)
    for lt in ("floating", "fixed"):
        for ot in ("call", "put"):
            want = orc.lookback_payoffs(paths, P["K"], lt, ot)
            spec = _ffi.make_spec(_ffi.LOOKBACK, n_steps, is_put=(ot == "put"), lookback_fixed=(lt == "fixed"))
            got, mom = engine.payoffs_from_normals(spec, _ffi.make_params(**P), Z)
            _check(got, want, np.maximum(paths.max(axis=1), P["K"]))
            if anchored():
                price = float(np.exp(-P["r"] * P["T"]) * mom["sum"] / mom["n"])
                assert price == pytest.approx(goldens["exotics"][f"lookback_{lt}_{ot}_{n_paths}x{n_steps}"], rel=TOL)


def test_greeks_from_reference_draws_match_real_reference(engine, goldens):
    """compute_greeks_unified over parity-mode prices == the real reference's Greeks (CRN via one Z)."""
    if goldens["numpy"] != np.__version__:
        pytest.skip("goldens recorded with another NumPy build")
    from optionslab_b200 import compute_greeks_unified, runtime
    Z = orc.normals_generator(42, (100_000, 252))

    class ParityPricer:
        def price(self, S, K, T, r, sigma, option_type, q=0.0, **kw):
            spec = _ffi.make_spec(_ffi.EUROPEAN, 252, is_put=(option_type == "put"), antithetic=True)
            _, mom = engine.payoffs_from_normals(spec, _ffi.make_params(S, K, T, r, sigma, q), Z, want_payoffs=False)
            return float(runtime.discounted_price(mom, r, T))

    for ot in ("call", "put"):
        got = compute_greeks_unified(ParityPricer(), **P, option_type=ot)
        for k, v in goldens["greeks"][f"100000x252_{ot}"].items():
            # second differences divide 1e-12-level price agreement by h^2 = 1e-4 (vomma) .. 1 (gamma)
            assert got[k] == pytest.approx(v, rel=1e-7, abs=1e-7), k


def test_parity_mode_argument_errors(engine):
    from optionslab_b200 import MonteCarloError
    spec = _ffi.make_spec(_ffi.ASIAN_ARITH, 4, antithetic=True)
    with pytest.raises(MonteCarloError, match="antithetic"):
        engine.payoffs_from_normals(spec, _ffi.make_params(**P), np.zeros((8, 4)))
    with pytest.raises(MonteCarloError):
        engine.payoffs_from_normals(_ffi.make_spec(_ffi.EUROPEAN, 4), _ffi.make_params(**P), np.zeros((8, 5)))


@pytest.mark.parametrize("n_steps", [2, 30, 32, 34, 64, 96, 250])
@pytest.mark.parametrize("n_paths", [1, 31, 129, 4097, 50_000])
def test_bulk_async_staged_kernel_equals_plain_load_kernel_bitwise(engine, n_paths, n_steps):
    """european_from_normals_tma_kernel (cp.async.bulk + mbarrier ring, taken for even n_steps) and from_normals_kernel
    (flag NO_BULK_COPY) run the same FP64 statement sequence: identical payoffs and moments, bit for bit — partial
    warps, partial last chunks (n_steps % 32 != 0), fewer chunks than ring stages, both summation forms."""
    Z = orc.normals_generator(7, (n_paths, n_steps))
    p = _ffi.make_params(**P, q=0.02)
    for accumulate in (False, True):
        for anti in (True, False):
            a_pay, a_mom = engine.payoffs_from_normals(_ffi.make_spec(_ffi.EUROPEAN, n_steps, antithetic=anti), p, Z, accumulate=accumulate)
            b_pay, b_mom = engine.payoffs_from_normals(_ffi.make_spec(_ffi.EUROPEAN, n_steps, antithetic=anti, no_bulk_copy=True), p, Z,
                                                       accumulate=accumulate)
            assert np.array_equal(a_pay, b_pay)
            assert a_mom["sum"] == b_mom["sum"] and a_mom["sum_sq"] == b_mom["sum_sq"] and a_mom["n"] == b_mom["n"] == n_paths * (2 if anti else 1)
    terminal = orc.gbm_terminal_from_normals(P["S"], P["T"], P["r"], P["sigma"], 0.02, Z)
    got, _ = engine.payoffs_from_normals(_ffi.make_spec(_ffi.EUROPEAN, n_steps, antithetic=True), p, Z)
    _check(got, orc.vanilla_payoffs(terminal, P["K"], "call"), np.maximum(terminal, P["K"]))
