"""GPU: Greeks at the reference's DEFAULT spot bump h = 1e-4 (monte_carlo_unified.py:522, :642; the call the Streamlit
page makes, streamlit_app/pages/3_MonteCarlo_Unified.py:167,279).

Scenarios S-h, S, S+h share every draw and every normalised terminal value S_T/S_0; the bump reaches the payoff only
through the strike ratio K/S, which moves by 2e-6.  The kernels price against an FP32 K/S and the fold moves the strike
to its FP64 value through the count of paid samples (mc_kernels.cuh, "Strike precision").  The checker is the FP64
oracle evaluated on the SAME Philox stream (oracle/philox_oracle.normals), so the comparison carries no Monte Carlo
noise: |delta_gpu - delta_oracle| <= 2e-4 (VERDICT r01 asked for 1e-3; round 1 was off by 1e-2 .. 2e-2 here).
"""

import numpy as np
import pytest

import optionslab_b200 as ob
from optionslab_b200 import _ffi
from oracle import philox_oracle as po
from oracle import reference_mc as orc

pytestmark = pytest.mark.gpu
K, T, R, SIGMA = 100.0, 1.0, 0.05, 0.2
SPOTS = (90.0, 97.3, 100.0, 103.7, 110.0)
H = 1e-4
TOL = 2e-4


def _oracle_delta(Z, S, ot, q=0.0, h=H, K_=K, T_=T, r=R, sigma=SIGMA):
    """Central difference of the FP64 antithetic estimator on the draws Z (gbm_numpy.py:46-51, monte_carlo.py:140-146)."""
    def price(s):
        terminal = orc.gbm_terminal_from_normals(s, T_, r, sigma, q, Z)
        return orc.discounted_mean(orc.vanilla_payoffs(terminal, K_, ot), r, T_)
    up, mid, down = price(S + h), price(S), price(S - h)
    return (up - down) / (2 * h), (up - 2 * mid + down) / h**2


@pytest.mark.parametrize("n_paths,n_steps", [(400_000, 1), (200_000, 100), (30_000, 252)])
def test_delta_gamma_at_default_h_matches_fp64_oracle_on_the_same_stream(n_paths, n_steps):
    seed = 11
    Z = po.normals(seed, n_paths, n_steps)
    pr = ob.MonteCarloPricerUni(n_paths, n_steps, seed=42)
    for S in SPOTS:
        for ot in ("call", "put"):
            delta, gamma = pr.delta_gamma(S, K, T, R, SIGMA, ot, seed=seed)  # default h
            want, _ = _oracle_delta(Z, S, ot)
            assert abs(delta - want) <= TOL, (S, ot, delta, want)
            assert np.isfinite(gamma)
            bs = orc.black_scholes_greeks(S, K, T, R, SIGMA)[0] - (0.0 if ot == "call" else 1.0)
            assert abs(delta - bs) <= 6.0 / np.sqrt(n_paths)  # pathwise delta: per-sample std < 1


def test_emulated_case_of_the_round_1_verdict():
    """S = 97.3, single step, 2M pairs: FP32 strike ratio gave 0.5642, FP64 0.5843."""
    pr = ob.MonteCarloPricerUni(2_000_000, 1, seed=42)
    d_small, _ = pr.delta_gamma(97.3, K, T, R, SIGMA, "call", seed=11)
    d_big, _ = pr.delta_gamma(97.3, K, T, R, SIGMA, "call", h=1.0, seed=11)
    bs = orc.black_scholes_greeks(97.3, K, T, R, SIGMA)[0]
    assert abs(d_small - bs) < 1.5e-3 and abs(d_small - d_big) < 1.5e-3


def test_delta_gamma_batch_at_default_h():
    """delta_gamma_batch (monte_carlo_unified.py:633-689): option i draws from Philox stream i."""
    n_paths, n_steps = 100_000, 50
    S = np.array(SPOTS)
    Kv = np.array([100.0, 100.0, 100.0, 95.0, 105.0])
    Tv = np.array([1.0, 1.0, 0.5, 0.5, 2.0])
    rv = np.full(5, R)
    sv = np.array([0.2, 0.2, 0.3, 0.15, 0.25])
    qv = np.array([0.0, 0.0, 0.02, 0.01, 0.0])
    pr = ob.MonteCarloPricerUni(n_paths, n_steps, seed=42)
    for ot in ("call", "put"):
        deltas, gammas = pr.delta_gamma_batch(S, Kv, Tv, rv, sv, ot, qv)  # default h
        assert np.all(np.isfinite(gammas))
        for i in range(5):
            Z = po.normals(42, n_paths, n_steps, stream=i)
            want, _ = _oracle_delta(Z, S[i], ot, q=qv[i], K_=Kv[i], T_=Tv[i], r=rv[i], sigma=sv[i])
            assert abs(deltas[i] - want) <= TOL, (i, ot, deltas[i], want)


def test_monte_carlo_pricer_delta_gamma_default_h():
    """MonteCarloPricer.delta_gamma - the call streamlit_app/st_utils.py:575 makes - has the same default bump."""
    n_paths, n_steps, seed = 250_000, 1, 5
    Z = po.normals(seed, n_paths, n_steps)
    pr = ob.MonteCarloPricer(n_paths, n_steps, seed=seed)
    for S in (97.3, 110.0):
        delta, _ = pr.delta_gamma(S, K, T, R, SIGMA, "call")
        assert abs(delta - _oracle_delta(Z, S, "call")[0]) <= TOL


def test_small_bumps_on_the_path_dependent_and_model_kernels(engine):
    """The same strike refinement in pathdep_kernel (Asian / fixed-strike lookback / barrier payoffs) and in the Heston
    and jump kernels: a 1e-4 spot bump against the FP64 oracle on the same draws (Asian), and against the h = 1 bump."""
    n_paths, n_steps, seed = 100_000, 64, 3
    Z = po.normals(seed, n_paths, n_steps)
    for S in (97.3, 103.7):
        params = np.stack([_ffi.make_params(S + b, K, T, R, SIGMA) for b in (H, 0.0, -H)]).reshape(1, 3)
        m = engine.simulate(_ffi.make_spec(_ffi.ASIAN_ARITH, n_steps), params, seed, n_paths)[0]
        prices = np.exp(-R * T) * m["sum"] / m["n"]
        delta = (prices[0] - prices[2]) / (2 * H)

        def asian(s):
            paths = orc.exotic_paths_from_normals(s, T, R, SIGMA, 0.0, Z)
            return orc.discounted_mean(orc.asian_payoffs(paths, K, "arithmetic", "call"), R, T)

        want = (asian(S + H) - asian(S - H)) / (2 * H)
        assert abs(delta - want) <= TOL, (S, delta, want)
    hes = ob.HestonPricer(kappa=2.0, theta=0.04, sigma_v=0.3, rho=-0.7, v0=0.04)
    mer = ob.MertonJumpDiffusion(lambda_j=1.0, mu_j=-0.1, sigma_j=0.15)
    for price in (lambda s: hes.price_monte_carlo(s, K, T, R, 0.0, "call", 400_000, 50, seed=9),
                  lambda s: mer.price_monte_carlo(s, K, T, R, SIGMA, "call", 0.0, 400_000, 50, seed=9)):
        d_small = (price(97.3 + H) - price(97.3 - H)) / (2 * H)
        d_big = (price(98.3) - price(96.3)) / 2.0
        assert abs(d_small - d_big) < 4e-3  # both estimate the same delta on the same draws (MC + curvature)


def test_control_variate_sums_follow_the_fp64_strike(engine):
    """The five control-variate sums (monte_carlo.py:176-186) get the same refinement: payoff sums of the CV launch equal
    those of the plain launch, and sum(payoff * S_T) matches the FP64 oracle on the same draws."""
    n_paths, n_steps, seed = 60_000, 8, 3
    Z = po.normals(seed, n_paths, n_steps)
    for S, ot in ((97.3, "call"), (103.7, "put")):
        spec = _ffi.make_spec(_ffi.EUROPEAN, n_steps, is_put=ot == "put", antithetic=True)
        m = engine.simulate(spec, _ffi.make_params(S, K, T, R, SIGMA).reshape(1, 1), seed, n_paths, control_variate=True)[0, 0]
        terminal = orc.gbm_terminal_from_normals(S, T, R, SIGMA, 0.0, Z)
        pay = orc.vanilla_payoffs(terminal, K, ot)
        assert m["sum_payoff"] == pytest.approx(pay.sum(), rel=2e-6)
        assert m["sum_payoff_terminal"] == pytest.approx((pay * terminal).sum(), rel=2e-6)
        assert m["sum_terminal"] == pytest.approx(terminal.sum(), rel=2e-6)
    # a batch (parameters through HBM, two scenarios per option): the control-variate launch and the plain launch agree on the payoff sums
    strikes = np.array([95.0, 100.0, 105.0])
    params = np.stack([_ffi.make_params(100.0 + b, strikes, T, R, SIGMA) for b in (H, -H)], axis=1)
    spec = _ffi.make_spec(_ffi.EUROPEAN, n_steps, antithetic=True)
    cv = engine.simulate(spec, params, seed, n_paths, control_variate=True)
    plain = engine.simulate(spec, params, seed, n_paths)
    assert cv.shape == (3, 2) and np.array_equal(cv["sum_payoff"], plain["sum"]) and np.array_equal(cv["sum_payoff_sq"], plain["sum_sq"])
    assert np.all(cv["sum_terminal"][:, 0] > cv["sum_terminal"][:, 1])  # S + h against S - h on the same draws
