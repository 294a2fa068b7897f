"""GPU: the quasi-Monte Carlo backend (MCMethod.QMC; src/simulation/gbm_qmc.py:14-47).

Integer work is bit-exact: the Sobol integers generated on the device equal scipy's.  The FP64 parity
kernel fed the reference's own norm.ppf values reproduces its payoffs within 1e-12.  The fused FP32 kernel
(inverse normal in registers) is compared with the reference's recorded prices with the tolerance stated
at each assert."""

import warnings

import numpy as np
import pytest

import optionslab_b200 as ob
from optionslab_b200 import _ffi, sobol
from oracle import reference_mc as orc

pytestmark = pytest.mark.gpu
P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)


def _ref_uniforms(seed, n, d):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return orc.qmc_uniforms(seed, n, d)


@pytest.mark.parametrize("d,seed,n,begin", [(7, 42, 5000, 0), (252, 1, 3000, 0), (33, 5, 4100, 8192), (1, 9, 17, 12345)])
def test_device_sobol_integers_equal_scipy(engine, d, seed, n, begin):
    table, shift, bits = sobol.sobol_table(d, seed)
    got = engine.sobol_points(table, shift, bits, n, point_begin=begin)
    want = _ref_uniforms(seed, begin + n, d)[begin:]
    np.testing.assert_array_equal(got.astype(np.float64) * 2.0**-bits, want)


def test_device_inverse_normal_matches_norm_ppf(engine):
    """FP32 branch-free polynomial (tools/fit_inverse_normal.py) vs scipy norm.ppf(clip(u)): abs 3e-6 everywhere
    (fit error 1.3e-6 + MUFU lg2/sqrt), odd symmetry exact."""
    bits = 30
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.integers(0, 1 << bits, size=2_000_000, dtype=np.uint64).astype(np.uint32),
                        np.arange(0, 70_000, dtype=np.uint32), (1 << bits) - 1 - np.arange(0, 70_000, dtype=np.uint32),
                        (1 << (bits - 1)) + np.arange(-1000, 1000, dtype=np.int64).astype(np.uint32)])
    got = engine.sobol_normals(x, bits).astype(np.float64)
    want = orc.qmc_normals_from_uniforms(x.astype(np.float64) * 2.0**-bits)
    assert np.max(np.abs(got - want)) < 3e-6
    assert got[x == 0][0] == pytest.approx(want[x == 0][0], abs=3e-6)  # u = 0 is clipped to 1e-10 (gbm_qmc.py:36)
    lo, hi = np.uint32(12345), np.uint32((1 << bits) - 12345)               # u and 1 - u
    pair = engine.sobol_normals(np.array([lo, hi], dtype=np.uint32), bits)
    assert pair[0] == -pair[1]


@pytest.mark.parametrize("n,d", [(4096, 7), (10000, 50), (16384, 64)])
@pytest.mark.parametrize("ot", ["call", "put"])
def test_fp64_parity_on_the_reference_qmc_draws(engine, goldens, anchored, n, d, ot):
    """Correctness test 1 for the QMC backend: the reference's own normals -> payoffs within 1e-12."""
    normals = orc.qmc_normals_from_uniforms(_ref_uniforms(42, n, d))
    terminal = orc.qmc_terminal_from_normals(P["S"], P["T"], P["r"], P["sigma"], 0.0, normals)
    want = orc.vanilla_payoffs(terminal, P["K"], ot)
    spec = _ffi.make_spec(_ffi.EUROPEAN, d, is_put=(ot == "put"), antithetic=False)
    got, mom = engine.payoffs_from_normals(spec, _ffi.make_params(**P), normals)
    assert np.max(np.abs(got - want) / np.maximum(terminal, P["K"])) <= 1e-12
    price = float(np.exp(-P["r"] * P["T"]) * mom["sum"] / mom["n"])
    if anchored(scipy=True):
        assert price == pytest.approx(goldens["qmc"][f"{n}x{d}_{ot}"]["price"], rel=1e-12)


@pytest.mark.parametrize("n,d", [(4096, 7), (10000, 50), (16384, 64), (65536, 252)])
@pytest.mark.parametrize("ot", ["call", "put"])
def test_fused_qmc_price_matches_the_real_reference(goldens, n, d, ot):
    """Fused FP32 kernel vs MonteCarloPricer(method=QMC) of the reference on the SAME point set.  FP32 inverse
    normal (6e-7 relative) + FP32 sum of d normals + MUFU.EX2: 2e-5 relative on the price, 1e-4 on the std error."""
    import scipy
    if goldens["numpy"] != np.__version__ or goldens["scipy"] != scipy.__version__:
        pytest.skip("goldens recorded with another NumPy / SciPy build")
    g = goldens["qmc"][f"{n}x{d}_{ot}"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        res = ob.MonteCarloPricer(n, d, seed=42, method=ob.MCMethod.QMC).price(**P, option_type=ot, return_error=True)
    assert res.n_paths == g["n_paths"] == n
    assert res.price == pytest.approx(g["price"], rel=2e-5)
    assert res.std_error == pytest.approx(g["std_error"], rel=1e-4)


def test_fused_qmc_with_dividend_and_other_point(goldens):
    import scipy
    if goldens["numpy"] != np.__version__ or goldens["scipy"] != scipy.__version__:
        pytest.skip("goldens recorded with another NumPy / SciPy build")
    got = ob.MonteCarloPricer(16384, 32, seed=7, method=ob.MCMethod.QMC).price(105.0, 95.0, 0.75, 0.03, 0.35, "call", q=0.02)
    assert got == pytest.approx(goldens["qmc"]["16384x32_call_q"]["price"], rel=2e-5)


def test_qmc_point_ranges_are_additive(engine):
    """Rank slices (multiples of 4096 points) of the same sequence add up to the whole: the multi-GPU contract."""
    d, n = 40, 20000
    table, shift, bits = sobol.sobol_table(d, 11)
    spec = _ffi.make_spec(_ffi.EUROPEAN, d, antithetic=False)
    params = _ffi.make_params(**P).reshape(1, 1)
    whole = engine.simulate_sobol(spec, params, table, shift, bits, n)[0, 0]
    parts = [engine.simulate_sobol(spec, params, table, shift, bits, c, point_begin=b)[0, 0]
             for b, c in (sobol.partition_points(n, r, 3) for r in range(3)) if c]
    assert sum(p["n"] for p in parts) == whole["n"] == n
    assert sum(p["sum"] for p in parts) == pytest.approx(whole["sum"], rel=1e-12)
    assert sum(p["sum_sq"] for p in parts) == pytest.approx(whole["sum_sq"], rel=1e-12)


def test_qmc_converges_faster_than_mc_and_agrees_with_black_scholes():
    bs = orc.black_scholes(**P, option_type="call")
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        q = ob.MonteCarloPricer(1 << 20, 64, seed=3, method=ob.MCMethod.QMC).price(**P, option_type="call", return_error=True)
    # 1M scrambled Sobol points: the error is far inside one MC standard error (0.014)
    assert abs(q.price - bs) < 0.25 * q.std_error
    g = ob.MonteCarloPricer(1 << 16, 32, seed=3, method=ob.MCMethod.QMC).greeks(**P, option_type="call")
    assert g["delta"] == pytest.approx(0.6368306511756191, abs=2e-3)
    assert g["vega"] == pytest.approx(37.52403469169379, rel=1e-2)


def test_qmc_rejects_misaligned_ranges_and_wrong_specs(engine):
    from optionslab_b200.exceptions import MonteCarloError

    table, shift, bits = sobol.sobol_table(8, 1)
    params = _ffi.make_params(**P).reshape(1, 1)
    with pytest.raises(MonteCarloError):
        engine.simulate_sobol(_ffi.make_spec(_ffi.EUROPEAN, 8), params, table, shift, bits, 100, point_begin=100)
    with pytest.raises(MonteCarloError):
        engine.simulate_sobol(_ffi.make_spec(_ffi.EUROPEAN, 8, antithetic=True), params, table, shift, bits, 100)
    with pytest.raises(MonteCarloError):
        engine.simulate_sobol(_ffi.make_spec(_ffi.ASIAN_ARITH, 8), params, table, shift, bits, 100)
