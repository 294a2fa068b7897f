"""GPU: the simulation layer (src/simulation/__init__.py) — terminal price arrays in the reference's layout."""

import warnings

import numpy as np
import pytest

import optionslab_b200 as ob
from optionslab_b200 import simulation as sim
from oracle import philox_oracle as po
from oracle import reference_mc as orc

pytestmark = pytest.mark.gpu
P = dict(S=100.0, T=1.0, r=0.05, sigma=0.2, q=0.01)


@pytest.mark.parametrize("n_paths,n_steps", [(100_000, 252), (4097, 7), (1, 1), (129, 33), (50_000, 8)])
def test_simulate_gbm_numpy_layout_mirror_identity_and_stream(n_paths, n_steps):
    out = sim.simulate_gbm_numpy(**P, n_paths=n_paths, n_steps=n_steps, seed=42)
    assert out.dtype == np.float64 and out.shape == (2 * n_paths,)
    # S_T(+Z) * S_T(-Z) = S^2 exp(2 (r - q - sigma^2/2) T) path by path (gbm_numpy.py:46-50)
    want = P["S"] ** 2 * np.exp(2 * (P["r"] - P["q"] - 0.5 * P["sigma"] ** 2) * P["T"])
    np.testing.assert_allclose(out[:n_paths] * out[n_paths:], want, rtol=3e-6)
    # the documented stream: FP64 oracle evaluation of the same Philox normals
    Z = po.normals(42, n_paths, n_steps)
    np.testing.assert_allclose(out, orc.gbm_terminal_from_normals(P["S"], P["T"], P["r"], P["sigma"], P["q"], Z), rtol=2e-5)
    single = sim.simulate_gbm_numpy(**P, n_paths=n_paths, n_steps=n_steps, seed=42, antithetic=False)
    assert single.shape == (n_paths,) and np.array_equal(single, out[:n_paths])
    assert np.array_equal(sim.simulate_gbm_numba(**P, n_paths=n_paths, n_steps=n_steps, seed=42, parallel=False), out)


def test_terminal_array_reproduces_the_pricers_moments_and_law():
    n, steps = 400_000, 32
    pr = ob.MonteCarloPricer(n, steps, seed=9)
    term = pr._simulate(**P)
    assert term.shape == (2 * n,)
    for ot, K in (("call", 100.0), ("put", 105.0)):
        pay = orc.vanilla_payoffs(term, K, ot)
        res = pr.price(P["S"], K, P["T"], P["r"], P["sigma"], ot, q=P["q"], return_error=True)
        assert res.price == pytest.approx(orc.discounted_mean(pay, P["r"], P["T"]), rel=1e-5)
        assert res.std_error == pytest.approx(orc.discounted_std_error(pay, P["r"], P["T"]), rel=1e-4)
    x = np.log(term[:n] / P["S"])
    mu, var = (P["r"] - P["q"] - 0.5 * P["sigma"] ** 2) * P["T"], P["sigma"] ** 2 * P["T"]
    assert abs(x.mean() - mu) < 4 * np.sqrt(var / n) and abs(x.var() - var) < 4 * var * np.sqrt(2 / n)
    fast = sim.simulate_gbm_numpy_fast(**P, n_paths=n, seed=9)
    assert np.array_equal(fast, ob.MonteCarloPricer(n, 1, seed=9)._simulate(**P))
    y = np.log(fast[:n] / P["S"])
    assert abs(y.mean() - mu) < 4 * np.sqrt(var / n) and abs(y.var() - var) < 4 * var * np.sqrt(2 / n)


@pytest.mark.parametrize("n_points,n_steps", [(4096, 7), (65536, 64), (10_000, 50), (1 << 18, 252)])
def test_simulate_gbm_qmc_matches_scipy_points(n_points, n_steps):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # scipy: N not a power of two
        normals = orc.qmc_normals_from_uniforms(orc.qmc_uniforms(42, n_points, n_steps))
        want = orc.qmc_terminal_from_normals(P["S"], P["T"], P["r"], P["sigma"], P["q"], normals)
        got = sim.simulate_gbm_qmc(**P, n_paths=n_points, n_steps=n_steps, seed=42)
        both = sim.simulate_gbm_qmc_antithetic(**P, n_paths=n_points, n_steps=n_steps, seed=42)
        via_pricer = ob.MonteCarloPricer(n_points, n_steps, seed=42, method=ob.MCMethod.QMC)._simulate(**P)
    assert got.shape == (n_points,) and both.shape == (2 * n_points,)
    np.testing.assert_allclose(got, want, rtol=3e-5)  # device Phi^-1 is within 1.3e-6 of norm.ppf per dimension
    assert np.array_equal(both[:n_points], got) and np.array_equal(via_pricer, got)
    prod = P["S"] ** 2 * np.exp(2 * (P["r"] - P["q"] - 0.5 * P["sigma"] ** 2) * P["T"])
    np.testing.assert_allclose(both[:n_points] * both[n_points:], prod, rtol=3e-6)


@pytest.mark.parametrize("n_steps", [2, 5, 8, 9, 252, 365])
def test_pair_sum_terminal_equals_the_step_by_step_device_normals(n_steps):
    """The terminal kernels take a Box-Muller pair as one term, sqrt(2) r sin(theta + pi/4) (normal.cuh, box_muller_pair_sum);
    b200mc_normals evaluates both branches of the same words.  The two describe the same draws: the terminal prices equal the
    FP64 terminal of the DEVICE's own step-by-step normals to FP32 rounding, odd tails included."""
    from optionslab_b200 import _ffi

    n, seed = 20_000, 77
    eng = _ffi.get_engine(0)
    Z = eng.generate_normals(seed, n, n_steps).astype(np.float64)
    want = orc.gbm_terminal_from_normals(P["S"], P["T"], P["r"], P["sigma"], P["q"], Z)
    got = sim.simulate_gbm_numpy(**P, n_paths=n, n_steps=n_steps, seed=seed)
    # MUFU.SIN on theta + pi/4 vs MUFU.COS + MUFU.SIN on theta: ~5e-7 absolute per pair on the unit circle
    np.testing.assert_allclose(got, want, rtol=3e-6)
