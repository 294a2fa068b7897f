"""CPU: the NumPy oracle restatement must reproduce the real reference's outputs.

The goldens were produced by ``tests/golden/make_goldens.py`` importing the
unmodified reference.  Same NumPy build => draws are identical and the oracle
must agree to the last bit (we allow 2 ulp-ish 1e-15 rel for BLAS/SIMD paths);
another NumPy build => only the RNG-free anchors are compared.
"""

import numpy as np
import pytest

from oracle import reference_mc as orc

P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)
REL = 1e-14


def _same_numpy(goldens):
    return goldens["numpy"] == np.__version__


def test_black_scholes_anchor(goldens):
    assert orc.black_scholes(**P, option_type="call") == pytest.approx(goldens["black_scholes"]["call"], rel=1e-15)
    assert orc.black_scholes(**P, option_type="put") == pytest.approx(goldens["black_scholes"]["put"], rel=1e-15)
    assert orc.black_scholes(**P, option_type="call", q=0.02) == pytest.approx(goldens["black_scholes"]["call_q2"], rel=1e-15)
    # reference tests/test_black_scholes.py:9,14
    assert orc.black_scholes(**P, option_type="call") == pytest.approx(10.45, rel=1e-2)
    assert orc.black_scholes(**P, option_type="put") == pytest.approx(5.57, rel=1e-2)


@pytest.mark.parametrize("n_sims,n_steps", [(10000, 50), (100000, 1), (4096, 7), (100000, 252)])
@pytest.mark.parametrize("ot", ["call", "put"])
def test_european_matches_reference(goldens, n_sims, n_steps, ot):
    if not _same_numpy(goldens):
        pytest.skip("goldens recorded with another NumPy build")
    g = goldens["european"][f"{n_sims}x{n_steps}_{ot}"]
    res = orc.european_price(**P, option_type=ot, num_simulations=n_sims, num_steps=n_steps, seed=42)
    assert res.price == pytest.approx(g["price"], rel=REL)
    assert res.std_error == pytest.approx(g["std_error"], rel=1e-12)
    assert res.n_paths == g["n_paths"]
    np.testing.assert_allclose(res.payoffs[:8], g["payoff_head"], rtol=REL, atol=0)
    np.testing.assert_allclose(res.payoffs[n_sims:n_sims + 8], g["payoff_mirror_head"], rtol=REL, atol=0)
    assert float(np.sum(res.payoffs)) == pytest.approx(g["payoff_sum"], rel=REL)


def test_european_with_dividend(goldens):
    if not _same_numpy(goldens):
        pytest.skip("goldens recorded with another NumPy build")
    res = orc.european_price(105.0, 95.0, 0.75, 0.03, 0.35, "call", q=0.02, num_simulations=20000, num_steps=64, seed=7)
    assert res.price == pytest.approx(goldens["european"]["20000x64_call_q"]["price"], rel=REL)


def test_control_variate_matches_reference(goldens):
    if not _same_numpy(goldens):
        pytest.skip("goldens recorded with another NumPy build")
    g = goldens["control_variate"]
    kw = dict(num_simulations=10000, num_steps=50, seed=42)
    assert orc.european_price_control_variate(**P, option_type="call", **kw) == pytest.approx(g["10000x50_call"], rel=REL)
    assert orc.european_price_control_variate(**P, option_type="put", **kw) == pytest.approx(g["10000x50_put"], rel=REL)
    assert orc.european_price_control_variate(**P, option_type="call", num_simulations=100000, num_steps=1, seed=42) == pytest.approx(g["100000x1_call"], rel=REL)
    assert orc.european_price_control_variate(105.0, 95.0, 0.75, 0.03, 0.35, "call", q=0.02, num_simulations=20000, num_steps=64, seed=7) == pytest.approx(g["20000x64_call_q"], rel=REL)


def test_uni_matches_reference(goldens):
    if not _same_numpy(goldens):
        pytest.skip("goldens recorded with another NumPy build")
    u = goldens["uni"]
    assert orc.uni_price(**P, option_type="call", num_simulations=10000, num_steps=50, seed=42) == pytest.approx(u["price_call"], rel=REL)
    assert orc.uni_price(**P, option_type="put", num_simulations=10000, num_steps=50, seed=42) == pytest.approx(u["price_put"], rel=REL)
    b = u["batch_inputs"]
    for ot in ("call", "put"):
        got = orc.uni_price_batch(b["S"], b["K"], b["T"], b["r"], b["sigma"], ot, np.array(b["q"]),
                                  num_simulations=b["num_simulations"], num_steps=b["num_steps"], seed=42)
        np.testing.assert_allclose(got, u[f"price_batch_{ot}"], rtol=REL)
    S = np.array(b["S"])
    kw = dict(num_simulations=b["num_simulations"], num_steps=b["num_steps"], seed=42)
    args = (b["K"], b["T"], b["r"], b["sigma"], "call", np.array(b["q"]))
    d, g = orc.central_delta_gamma(orc.uni_price_batch(S + 1.0, *args, **kw), orc.uni_price_batch(S, *args, **kw),
                                   orc.uni_price_batch(S - 1.0, *args, **kw), 1.0)
    np.testing.assert_allclose(d, u["delta_gamma_batch_h1"]["delta"], rtol=1e-12)
    np.testing.assert_allclose(g, u["delta_gamma_batch_h1"]["gamma"], rtol=1e-9)


@pytest.mark.parametrize("ot", ["call", "put"])
def test_greeks_match_reference(goldens, ot):
    if not _same_numpy(goldens):
        pytest.skip("goldens recorded with another NumPy build")
    g = goldens["greeks"][f"100000x252_{ot}"]

    def price_fn(S, K, T, r, sigma, q):
        return orc.european_price(S, K, T, r, sigma, ot, q, num_simulations=100000, num_steps=252, seed=42).price

    out = orc.greeks_bump_and_revalue(price_fn, **P)
    assert list(out.keys()) == ["price", "delta", "gamma", "vega", "theta", "rho", "vanna", "charm", "vomma"]
    for k, v in g.items():
        assert out[k] == pytest.approx(v, rel=1e-9, abs=1e-9), k


def test_greeks_edge_cases(goldens):
    if not _same_numpy(goldens):
        pytest.skip("goldens recorded with another NumPy build")

    def fn(ot):
        return lambda S, K, T, r, sigma, q: orc.european_price(
            S, K, T, r, sigma, ot, q, num_simulations=20000, num_steps=16, seed=3).price

    out = orc.greeks_bump_and_revalue(fn("put"), 105.0, 95.0, 0.75, 0.03, 0.35, q=0.02)
    for k, v in goldens["greeks"]["20000x16_put_q"].items():
        assert out[k] == pytest.approx(v, rel=1e-9, abs=1e-9), k
    out = orc.greeks_bump_and_revalue(fn("call"), 100.0, 100.0, 0.002, 0.05, 0.2)   # T < 1/365 branch
    for k, v in goldens["greeks"]["20000x16_call_shortT"].items():
        assert out[k] == pytest.approx(v, rel=1e-9, abs=1e-9), k


@pytest.mark.parametrize("tag,n_paths,n_steps", [("5000x12", 5000, 12), ("100000x252", 100000, 252)])
def test_asian_lookback_match_reference(goldens, tag, n_paths, n_steps):
    if not _same_numpy(goldens):
        pytest.skip("goldens recorded with another NumPy build")
    ex = goldens["exotics"]
    kw = dict(seed=42, n_paths=n_paths, n_steps=n_steps)
    assert orc.exotic_price("asian", **P, **kw, avg_type="arithmetic", option_type="call") == pytest.approx(ex[f"asian_arith_call_{tag}"], rel=REL)
    assert orc.exotic_price("asian", **P, **kw, avg_type="arithmetic", option_type="put") == pytest.approx(ex[f"asian_arith_put_{tag}"], rel=REL)
    assert orc.exotic_price("asian", **P, **kw, avg_type="geometric", option_type="call") == pytest.approx(ex[f"asian_geom_call_{tag}"], rel=REL)
    for lt in ("floating", "fixed"):
        for ot in ("call", "put"):
            got = orc.exotic_price("lookback", **P, **kw, lookback_type=lt, option_type=ot)
            assert got == pytest.approx(ex[f"lookback_{lt}_{ot}_{tag}"], rel=REL)


@pytest.mark.parametrize("tag,n_paths,n_steps", [("5000x12", 5000, 12), ("100000x365", 100000, 365)])
def test_barrier_matches_reference(goldens, tag, n_paths, n_steps):
    if not _same_numpy(goldens):
        pytest.skip("goldens recorded with another NumPy build")
    ex = goldens["exotics"]
    for B, kinds in [(120.0, ("up-and-out", "up-and-in")), (85.0, ("down-and-out", "down-and-in"))]:
        for bt in kinds:
            for ot in ("call", "put"):
                got = orc.exotic_price("barrier", **P, seed=42, n_paths=n_paths, n_steps=n_steps,
                                       barrier=B, barrier_type=bt, option_type=ot)
                assert got == pytest.approx(ex[f"barrier_{bt}_{ot}_B{int(B)}_{tag}"], rel=REL, abs=1e-15)


def test_asian_closed_form_and_adapter_greeks(goldens):
    ex = goldens["exotics"]
    assert orc.asian_geometric_closed_form(**P) == pytest.approx(ex["asian_geom_closed_form_call"], rel=1e-14)
    if not _same_numpy(goldens):
        pytest.skip("goldens recorded with another NumPy build")

    def fn(S, K, T, r, sigma, q):
        return orc.exotic_price("asian", S, K, T, r, sigma, q, seed=42, n_paths=20000, n_steps=32)

    out = orc.greeks_bump_and_revalue(fn, **P)
    for k, v in ex["asian_adapter_greeks_20000x32"].items():
        assert out[k] == pytest.approx(v, rel=1e-9, abs=1e-9), k


def test_barrier_rejects_nonpositive_barrier():
    # reference tests/test_exotic_options.py:187-193
    with pytest.raises(ValueError, match="positive"):
        orc.exotic_price("barrier", **P, seed=1, n_paths=10, n_steps=2, barrier=0.0)


@pytest.mark.parametrize("n_sims,n_steps", [(4096, 7), (16384, 64), (65536, 252), (10000, 50)])
@pytest.mark.parametrize("ot", ["call", "put"])
def test_qmc_backend_matches_reference(goldens, n_sims, n_steps, ot):
    """MCMethod.QMC (monte_carlo.py:94-97 -> gbm_qmc.py:14-47): scipy's scrambled Sobol + norm.ppf."""
    import scipy
    import warnings

    if not _same_numpy(goldens) or goldens["scipy"] != scipy.__version__:
        pytest.skip("goldens recorded with another NumPy / SciPy build")
    g = goldens["qmc"][f"{n_sims}x{n_steps}_{ot}"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # scipy: N not a power of two
        res = orc.european_price_qmc(**P, option_type=ot, num_simulations=n_sims, num_steps=n_steps, seed=42)
    assert res.price == pytest.approx(g["price"], rel=REL)
    assert res.std_error == pytest.approx(g["std_error"], rel=1e-12)
    assert res.n_paths == g["n_paths"] == n_sims
    np.testing.assert_allclose(res.payoffs[:8], g["payoff_head"], rtol=REL, atol=0)
    assert float(np.sum(res.payoffs)) == pytest.approx(g["payoff_sum"], rel=REL)


HES = dict(kappa=2.0, theta=0.04, sigma_v=0.3, rho=-0.7, v0=0.04)
MER = dict(lambda_j=1.0, mu_j=-0.1, sigma_j=0.15)
KOU = dict(lambda_j=2.0, p=0.4, eta1=10.0, eta2=5.0)


@pytest.mark.parametrize("n_paths,n_steps", [(20000, 50), (4097, 7), (100000, 252)])
@pytest.mark.parametrize("ot", ["call", "put"])
def test_heston_mc_matches_reference(goldens, n_paths, n_steps, ot):
    """HestonPricer.price_monte_carlo (heston.py:184-255), full-truncation Euler on the legacy global generator."""
    if not _same_numpy(goldens):
        pytest.skip("goldens recorded with another NumPy build")
    got = orc.heston_price_mc(100.0, 100.0, 1.0, 0.05, 0.01, ot, **HES, n_paths=n_paths, n_steps=n_steps, seed=42)
    assert got == pytest.approx(goldens["models"][f"heston_{ot}_{n_paths}x{n_steps}"], rel=REL)


def test_heston_mc_with_truncation_active_matches_reference(goldens):
    if not _same_numpy(goldens):
        pytest.skip("goldens recorded with another NumPy build")
    got = orc.heston_price_mc(100.0, 110.0, 0.5, 0.03, 0.0, "call", kappa=1.0, theta=0.09, sigma_v=0.8, rho=-0.3, v0=0.02,
                              n_paths=20000, n_steps=50, seed=7)
    assert got == pytest.approx(goldens["models"]["heston_feller_violated_call_20000x50"], rel=REL)


@pytest.mark.parametrize("n_paths,n_steps", [(5000, 20), (20000, 50)])
@pytest.mark.parametrize("ot", ["call", "put"])
def test_jump_diffusion_mc_matches_reference(goldens, n_paths, n_steps, ot):
    """Merton / Kou price_monte_carlo (jump_diffusion.py:160-225, :325-377): the oracle replays the reference's
    interleaved standard_normal / poisson / jump-size draws."""
    if not _same_numpy(goldens):
        pytest.skip("goldens recorded with another NumPy build")
    kw = dict(n_paths=n_paths, n_steps=n_steps, seed=42)
    got = orc.merton_price_mc(100.0, 100.0, 1.0, 0.05, 0.2, ot, 0.01, **MER, **kw)
    assert got == pytest.approx(goldens["models"][f"merton_{ot}_{n_paths}x{n_steps}"], rel=REL)
    got = orc.kou_price_mc(100.0, 100.0, 1.0, 0.05, 0.2, ot, 0.01, **KOU, **kw)
    assert got == pytest.approx(goldens["models"][f"kou_{ot}_{n_paths}x{n_steps}"], rel=REL)


TIGHT_AUTOCALL = dict(autocall_barrier=1.05, coupon_barrier=0.9, coupon_rate=0.08, ki_barrier=0.75)
WIDE_CLIQUET = dict(local_cap=0.08, local_floor=-0.03, global_cap=0.5, global_floor=-0.1)
STRUCTURED_CASES = [(100000, 252, 21, 12), (5000, 12, 3, 4), (20000, 100, 7, 9), (4097, 37, 37, 37)]


@pytest.mark.parametrize("n_paths,n_steps,freq,nper", STRUCTURED_CASES)
def test_autocallable_and_cliquet_match_reference(goldens, n_paths, n_steps, freq, nper):
    """exotic_options.py:404-552 restated (autocallable_payoffs / cliquet_payoffs) vs the real classes."""
    if not _same_numpy(goldens):
        pytest.skip("goldens recorded with another NumPy build")
    sp, tag = goldens["structured"], f"{n_paths}x{n_steps}"
    kw = dict(seed=42, n_paths=n_paths, n_steps=n_steps)
    assert orc.exotic_price("autocallable", **P, **kw, observation_freq=freq) == pytest.approx(sp[f"autocallable_default_{tag}_f{freq}"], rel=REL)
    assert orc.exotic_price("autocallable", **P, **kw, observation_freq=freq, **TIGHT_AUTOCALL) == pytest.approx(
        sp[f"autocallable_tight_{tag}_f{freq}"], rel=REL)
    assert orc.exotic_price("cliquet", **P, **kw, n_periods=nper) == pytest.approx(sp[f"cliquet_default_{tag}_p{nper}"], rel=REL)
    assert orc.exotic_price("cliquet", **P, **kw, n_periods=nper, **WIDE_CLIQUET) == pytest.approx(sp[f"cliquet_wide_{tag}_p{nper}"], rel=REL)


def test_structured_products_with_dividend_and_adapter_greeks(goldens):
    if not _same_numpy(goldens):
        pytest.skip("goldens recorded with another NumPy build")
    sp = goldens["structured"]
    kw = dict(seed=7, n_paths=20000, n_steps=64)
    assert orc.exotic_price("autocallable", 105.0, 100.0, 1.5, 0.03, 0.35, 0.02, **kw, observation_freq=8) == pytest.approx(
        sp["autocallable_q_sigma_20000x64_f8"], rel=REL)
    assert orc.exotic_price("cliquet", 105.0, 100.0, 1.5, 0.03, 0.35, 0.02, **kw, n_periods=8) == pytest.approx(sp["cliquet_q_sigma_20000x64_p8"], rel=REL)

    def fn(S, K, T, r, sigma, q):
        return orc.exotic_price("cliquet", S, K, T, r, sigma, q, seed=42, n_paths=20000, n_steps=36, n_periods=6)

    out = orc.greeks_bump_and_revalue(fn, **P)
    for k, v in sp["cliquet_adapter_greeks_20000x36"].items():
        assert out[k] == pytest.approx(v, rel=1e-8, abs=1e-8), k


def test_numba_backend_restatement_has_the_reference_layout_and_law():
    """gbm_numba.py:74-97 restated for the CPU timing arm: 2N values, [i] and [i+N] mirrored (their product is
    S^2 exp(2 (r-q-sigma^2/2) T) exactly), seeded per path, mean = forward within 4 standard errors."""
    pytest.importorskip("numba")
    n, steps = 20000, 16
    out = orc.numba_backend_terminal(100.0, 1.0, 0.05, 0.2, 0.01, n, steps, 42)
    assert out.shape == (2 * n,)
    np.testing.assert_allclose(out[:n] * out[n:], 100.0**2 * np.exp(2 * (0.05 - 0.01 - 0.02) * 1.0), rtol=1e-12)
    assert abs(out.mean() - 100.0 * np.exp(0.04)) < 4 * out.std() / np.sqrt(2 * n)
    again = orc.numba_backend_terminal(100.0, 1.0, 0.05, 0.2, 0.01, n, steps, 42)
    assert np.array_equal(out, again)  # per-path seeding: deterministic for any thread count
