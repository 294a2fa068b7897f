"""GPU: Heston and jump-diffusion Monte Carlo.

Test 1 (FP64, reference draws): the oracle replays the generator calls of the reference's price_monte_carlo
methods (pinned to the real reference by tests/test_oracle_golden.py); fed those draws the FP64 kernels
reproduce the per-path payoffs within 1e-12 and the recorded reference prices within 1e-12 relative.
Test 2 (on-device Philox): prices within 3 combined standard errors of the reference's recorded Monte Carlo
values, of the Merton series, and of an exact-law NumPy sampler; plus FP32-vs-FP64 on identical draws (Heston)."""

import numpy as np
import pytest

import optionslab_b200 as ob
from optionslab_b200 import _ffi
from oracle import philox_oracle
from oracle import reference_mc as orc

pytestmark = pytest.mark.gpu
TOL = 1e-12
HES = dict(kappa=2.0, theta=0.04, sigma_v=0.3, rho=-0.7, v0=0.04)
MER = dict(lambda_j=1.0, mu_j=-0.1, sigma_j=0.15)
KOU = dict(lambda_j=2.0, p=0.4, eta1=10.0, eta2=5.0)
PT = dict(S=100.0, K=100.0, T=1.0, r=0.05)


def _heston_params(S, K, T, r, q, **h):
    p = np.zeros(1, dtype=_ffi.HESTON_PARAMS_DTYPE)
    for k, v in dict(S=S, K=K, T=T, r=r, q=q, **h).items():
        p[k] = v
    return p


# ------------------------------------------------------------------ test 1: FP64 on the reference's draws
@pytest.mark.parametrize("n_paths,n_steps", [(20000, 50), (4097, 7), (1, 3), (129, 33)])
@pytest.mark.parametrize("ot", ["call", "put"])
def test_heston_fp64_from_reference_draws(engine, goldens, anchored, n_paths, n_steps, ot):
    Z = orc.heston_draws(42, n_paths, n_steps)
    want = orc.heston_payoffs_from_normals(**PT, q=0.01, **HES, Z=Z, option_type=ot)
    got, mom = engine.heston_from_normals(_heston_params(**PT, q=0.01, **HES), ot == "put", Z)
    assert np.max(np.abs(got - want) / np.maximum(want + PT["K"], PT["K"])) <= TOL
    assert mom["n"] == n_paths and mom["sum"] == pytest.approx(float(np.sum(want)), rel=TOL)
    key = f"heston_{ot}_{n_paths}x{n_steps}"
    if key in goldens["models"] and anchored():
        price = float(np.exp(-PT["r"] * PT["T"]) * mom["sum"] / mom["n"])
        assert price == pytest.approx(goldens["models"][key], rel=TOL)


def test_heston_fp64_with_truncation_active(engine, goldens, anchored):
    h = dict(kappa=1.0, theta=0.09, sigma_v=0.8, rho=-0.3, v0=0.02)  # Feller violated: v hits the floor
    Z = orc.heston_draws(7, 20000, 50)
    want = orc.heston_payoffs_from_normals(100.0, 110.0, 0.5, 0.03, 0.0, **h, Z=Z, option_type="call")
    got, mom = engine.heston_from_normals(_heston_params(100.0, 110.0, 0.5, 0.03, 0.0, **h), False, Z)
    assert np.max(np.abs(got - want)) <= TOL * 110.0
    if anchored():
        assert float(np.exp(-0.03 * 0.5) * mom["sum"] / mom["n"]) == pytest.approx(goldens["models"]["heston_feller_violated_call_20000x50"], rel=TOL)


@pytest.mark.parametrize("model", ["merton", "kou"])
@pytest.mark.parametrize("n_paths,n_steps", [(5000, 20), (20000, 50)])
@pytest.mark.parametrize("ot", ["call", "put"])
def test_jump_diffusion_fp64_from_reference_draws(engine, goldens, anchored, model, n_paths, n_steps, ot):
    if model == "merton":
        dW, J = orc.merton_draws(42, **MER, T=PT["T"], n_paths=n_paths, n_steps=n_steps)
        lk = MER["lambda_j"] * orc.merton_kappa(MER["mu_j"], MER["sigma_j"])
    else:
        dW, J = orc.kou_draws(42, **KOU, T=PT["T"], n_paths=n_paths, n_steps=n_steps)
        lk = KOU["lambda_j"] * orc.kou_kappa(KOU["p"], KOU["eta1"], KOU["eta2"])
    assert np.count_nonzero(J) > 0
    want = orc.jump_payoffs_from_draws(**PT, sigma=0.2, q=0.01, lambda_kappa=lk, dW=dW, J=J, option_type=ot)
    got, mom = engine.jump_diffusion_from_draws(_ffi.make_params(**PT, sigma=0.2, q=0.01), lk, ot == "put", dW, J)
    assert np.max(np.abs(got - want) / np.maximum(want + PT["K"], PT["K"])) <= TOL
    if anchored():
        price = float(np.exp(-PT["r"] * PT["T"]) * mom["sum"] / mom["n"])
        assert price == pytest.approx(goldens["models"][f"{model}_{ot}_{n_paths}x{n_steps}"], rel=TOL)


# ------------------------------------------------------------------ test 2: on-device Philox
def test_heston_fused_fp32_vs_fp64_on_identical_draws():
    """The C oracle regenerates the engine's own stream in FP64 (pair n of a path = step n: cos -> Z1, sin -> Z);
    the FP32 fused kernel must agree with the FP64 restatement on those draws to 2e-4 relative on the sum."""
    n_paths, n_steps, seed = 1 << 15, 64, 5
    Zs = philox_oracle.normals(seed, n_paths, 2 * n_steps)           # [path, 2*step + {0: cos, 1: sin}]
    Z = np.ascontiguousarray(Zs.reshape(n_paths, n_steps, 2).transpose(1, 2, 0))
    for ot in ("call", "put"):
        want = orc.heston_payoffs_from_normals(**PT, q=0.01, **HES, Z=Z, option_type=ot)
        price, se = ob.HestonPricer(**HES).price_monte_carlo(**PT, q=0.01, option_type=ot, n_paths=n_paths, n_steps=n_steps, seed=seed,
                                                            return_error=True)
        ref_price = orc.discounted_mean(want, PT["r"], PT["T"])
        assert price == pytest.approx(ref_price, rel=2e-4)
        assert abs(price - ref_price) < 0.05 * se


def test_heston_mc_within_three_standard_errors_of_the_reference(goldens):
    n = 1_000_000
    for ot in ("call", "put"):
        price, se = ob.HestonPricer(**HES).price_monte_carlo(**PT, q=0.01, option_type=ot, n_paths=n, n_steps=252, seed=42, return_error=True)
        ref = goldens["models"][f"heston_{ot}_100000x252"]
        se_ref = se * np.sqrt(n / 100_000)
        assert abs(price - ref) <= 3 * np.hypot(se, se_ref), (price, ref, se)
    # put-call parity on common draws: C - P = S e^{-qT} - K e^{-rT}
    c = ob.HestonPricer(**HES).price_monte_carlo(**PT, q=0.01, option_type="call", n_paths=n, n_steps=64, seed=9)
    p = ob.HestonPricer(**HES).price_monte_carlo(**PT, q=0.01, option_type="put", n_paths=n, n_steps=64, seed=9)
    assert c - p == pytest.approx(100 * np.exp(-0.01) - 100 * np.exp(-0.05), abs=4 * 0.2 * 100 / np.sqrt(n))


def test_heston_is_deterministic_and_additive_over_path_ranges(engine):
    p = _heston_params(**PT, q=0.0, **HES)
    a = engine.simulate_heston(p, False, 32, 11, 100_000)[0]
    b = engine.simulate_heston(p, False, 32, 11, 100_000)[0]
    assert a == b
    lo = engine.simulate_heston(p, False, 32, 11, 37_000)[0]
    hi = engine.simulate_heston(p, False, 32, 11, 63_000, path_begin=37_000)[0]
    assert lo["sum"] + hi["sum"] == pytest.approx(a["sum"], rel=1e-6)
    assert lo["n"] + hi["n"] == a["n"]


def test_merton_mc_matches_series_and_reference(goldens):
    n = 4_000_000
    mer = ob.MertonJumpDiffusion(**MER)
    for ot in ("call", "put"):
        price, se = mer.price_monte_carlo(**PT, sigma=0.2, option_type=ot, q=0.01, n_paths=n, n_steps=50, seed=42, return_error=True)
        series = orc.merton_series_price(**PT, sigma=0.2, option_type=ot, q=0.01, **MER)
        assert series == pytest.approx(goldens["models"][f"merton_analytic_{ot}"], rel=1e-12)
        assert abs(price - series) <= 3 * se, (price, series, se)
        ref = goldens["models"][f"merton_{ot}_20000x50"]
        assert abs(price - ref) <= 3 * se * np.sqrt(n / 20_000)


@pytest.mark.parametrize("model,cls,jp", [("merton", "MertonJumpDiffusion", MER), ("kou", "KouJumpDiffusion", KOU)])
def test_jump_diffusion_mc_matches_exact_law_sampler_and_reference(goldens, model, cls, jp):
    n = 4_000_000
    pricer = getattr(ob, cls)(**jp)
    terminal = orc.jump_terminal_exact_law(model, PT["S"], PT["T"], PT["r"], 0.2, 0.01, jp, n, seed=123)
    for ot in ("call", "put"):
        price, se = pricer.price_monte_carlo(**PT, sigma=0.2, option_type=ot, q=0.01, n_paths=n, n_steps=50, seed=42, return_error=True)
        pay = orc.vanilla_payoffs(terminal, PT["K"], ot)
        ref_price, ref_se = orc.discounted_mean(pay, PT["r"], PT["T"]), orc.discounted_std_error(pay, PT["r"], PT["T"])
        assert abs(price - ref_price) <= 3 * np.hypot(se, ref_se), (price, ref_price, se, ref_se)
        ref = goldens["models"][f"{model}_{ot}_20000x50"]
        assert abs(price - ref) <= 3 * se * np.sqrt(n / 20_000)
    # martingale: a call struck at ~0 prices the discounted forward S e^{-qT} (tests the lambda*kappa compensator)
    fwd, se = pricer.price_monte_carlo(100.0, 1e-9, 1.0, 0.05, 0.2, "call", 0.01, n_paths=n, n_steps=20, seed=1, return_error=True)
    assert abs(fwd - 100 * np.exp(-0.01)) <= 3.5 * se


def test_jump_kernel_without_jumps_is_the_european_kernel(engine):
    """lambda_j = 0: same draws, same terminal sum as the (non-antithetic) European kernel."""
    params = _ffi.make_params(**PT, sigma=0.2, q=0.01).reshape(1)
    jumps = np.zeros(1, dtype=_ffi.JUMP_PARAMS_DTYPE)
    jumps["model"], jumps["lambda_j"], jumps["a"], jumps["b"] = _ffi.JUMP_MERTON, 0.0, -0.1, 0.15
    a = engine.simulate_jump_diffusion(params, jumps, False, 40, 3, 200_000)[0]
    b = engine.simulate(_ffi.make_spec(_ffi.EUROPEAN, 40, antithetic=False), params.reshape(1, 1), 3, 200_000)[0, 0]
    assert a["n"] == b["n"]
    assert a["sum"] == pytest.approx(b["sum"], rel=1e-6) and a["sum_sq"] == pytest.approx(b["sum_sq"], rel=1e-6)


def test_model_constructors_validate_like_the_reference():
    with pytest.raises(ValueError):
        ob.HestonPricer(kappa=0.0, theta=0.04, sigma_v=0.3, rho=-0.7, v0=0.04)
    with pytest.raises(ValueError):
        ob.HestonPricer(kappa=2.0, theta=0.04, sigma_v=0.3, rho=-1.7, v0=0.04)
    with pytest.warns(UserWarning):
        ob.HestonPricer(kappa=1.0, theta=0.09, sigma_v=0.8, rho=-0.3, v0=0.02)
    with pytest.raises(ValueError):
        ob.MertonJumpDiffusion(-1.0, 0.0, 0.1)
    with pytest.raises(ValueError):
        ob.KouJumpDiffusion(1.0, 0.4, 1.0, 5.0)


# ---------------- bump-and-revalue Greeks of the model pricers: adapters + single-launch CRN scenarios ----------

def test_model_scenarios_share_draws_and_match_separate_repricings():
    """price_scenarios (one launch, option axis = CRN scenario axis) == the same re-pricings issued call by call
    with the same seed; only the tile plan (hence the FP64 summation order) may differ."""
    hes = ob.HestonPricer(kappa=2.0, theta=0.04, sigma_v=0.3, rho=-0.7, v0=0.04)
    sc = [(100.0, 100.0, 1.0, 0.05, 0.04, 0.01), (101.0, 100.0, 1.0, 0.05, 0.04, 0.01), (100.0, 95.0, 0.5, 0.03, 0.0441, 0.0)]
    fused = hes.price_scenarios(sc, "put", n_paths=200_000, n_steps=50, seed=9)
    for (S, K, T, r, v0, q), got in zip(sc, fused):
        one = ob.HestonPricer(kappa=2.0, theta=0.04, sigma_v=0.3, rho=-0.7, v0=v0).price_monte_carlo(S, K, T, r, q, "put", 200_000, 50, seed=9)
        assert got == pytest.approx(one, rel=1e-8)  # tile plans differ: FP32 thread sums associate differently
    for pricer in (ob.MertonJumpDiffusion(**MER), ob.KouJumpDiffusion(**KOU)):
        sj = [(100.0, 100.0, 1.0, 0.05, 0.2, 0.01), (99.0, 100.0, 1.0, 0.05, 0.21, 0.01)]
        fused = pricer.price_scenarios(sj, "call", n_paths=200_000, n_steps=40, seed=5)
        for (S, K, T, r, sig, q), got in zip(sj, fused):
            assert got == pytest.approx(pricer.price_monte_carlo(S, K, T, r, sig, "call", q, 200_000, 40, seed=5), rel=1e-8)


def test_adapter_greeks_fused_equal_call_by_call_and_approach_black_scholes():
    """compute_greeks_unified through HestonAdapter / JumpDiffusionAdapter: the fused single-launch route equals the
    reference's call-by-call route (same seed => same draws), and in the Black-Scholes limits of the models
    (vol-of-vol -> 0 with v0 = theta; lambda_j = 0) delta / gamma / rho approach the closed form."""
    bs = orc.black_scholes  # noqa: F841  (closed-form anchors below are its derivatives at S=K=100, T=1, r=5%, sigma=20%)
    BS_DELTA, BS_GAMMA, BS_VEGA, BS_RHO = 0.6368306511756191, 0.018762017345846895, 37.52403469169379, 53.232481545376345

    class CallByCall:
        def __init__(self, inner):
            self.inner = inner

        def price(self, *a, **k):
            return self.inner.price(*a, **k)

    hes = ob.HestonPricer(kappa=1.0, theta=0.04, sigma_v=1e-3, rho=0.0, v0=0.04)
    ad = ob.HestonAdapter(hes, n_paths=400_000, n_steps=64, seed=21)
    g = ob.compute_greeks_unified(ad, 100.0, 100.0, 1.0, 0.05, 0.2, "call")
    g2 = ob.compute_greeks_unified(CallByCall(ad), 100.0, 100.0, 1.0, 0.05, 0.2, "call")
    assert list(g) == list(g2) == ["price", "delta", "gamma", "vega", "theta", "rho", "vanna", "charm", "vomma"]
    assert g["price"] == pytest.approx(g2["price"], rel=1e-9) and g["delta"] == pytest.approx(g2["delta"], rel=1e-6)
    assert hes.v0 == 0.04  # the adapter restores the pricer's state
    assert g["delta"] == pytest.approx(BS_DELTA, abs=0.01) and g["gamma"] == pytest.approx(BS_GAMMA, rel=0.1)
    assert g["rho"] == pytest.approx(BS_RHO, rel=0.03)
    assert 0 < g["vega"] < BS_VEGA  # sigma only moves v0, which mean-reverts to theta: less than the flat-vol vega

    jd = ob.JumpDiffusionAdapter(ob.MertonJumpDiffusion(0.0, -0.1, 0.15), n_paths=400_000, n_steps=32, seed=4)
    gj = ob.greeks_jump_diffusion(ob.MertonJumpDiffusion(0.0, -0.1, 0.15), 100.0, 100.0, 1.0, 0.05, 0.2, "call", n_paths=400_000, n_steps=32, seed=4)
    assert dict(gj) == dict(ob.compute_greeks_unified(jd, 100.0, 100.0, 1.0, 0.05, 0.2, "call"))
    assert gj["delta"] == pytest.approx(BS_DELTA, abs=0.01) and gj["vega"] == pytest.approx(BS_VEGA, rel=0.03)
    assert gj["rho"] == pytest.approx(BS_RHO, rel=0.03)
    # jumps add convexity: with lambda_j > 0 the at-the-money call is worth more, its delta stays in (0, 1)
    gm = ob.greeks_jump_diffusion(ob.MertonJumpDiffusion(**MER), 100.0, 100.0, 1.0, 0.05, 0.2, "call", n_paths=400_000, n_steps=32, seed=4)
    assert gm["price"] > gj["price"] and 0 < gm["delta"] < 1
