"""CPU: host-side mirror of the reference interface (no kernel is launched)."""

from collections import OrderedDict

import numpy as np
import pytest

import optionslab_b200 as ob
from optionslab_b200 import distributed, greeks, runtime
from oracle import reference_mc as orc

P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)


class BSPricer:
    """A PricerProtocol object with no price_scenarios: exercises the call-by-call route."""

    def __init__(self):
        self.calls = 0

    def price(self, S, K, T, r, sigma, option_type, q=0.0, **kw):
        self.calls += 1
        return orc.black_scholes(S, K, T, r, sigma, option_type, q)


class FusedBS(BSPricer):
    def price_scenarios(self, scenarios, option_type, **kw):
        self.fused_sizes = getattr(self, "fused_sizes", []) + [len(scenarios)]
        return [orc.black_scholes(s[0], s[1], s[2], s[3], s[4], option_type, s[5]) for s in scenarios]


def test_constructor_contracts():
    # tests/test_monte_carlo.py:95-112 and :375-390 of the reference
    p = ob.MonteCarloPricer(num_simulations=5000, num_steps=100, seed=123)
    assert (p.num_simulations, p.num_steps, p.seed) == (5000, 100, 123)
    for bad in (0, -100):
        with pytest.raises(ValueError):
            ob.MonteCarloPricer(num_simulations=bad)
    assert ob.MonteCarloPricer().num_steps == 1 and ob.MonteCarloPricerUni().num_steps == 100
    assert isinstance(ob.MonteCarloPricer().seed, int) and 0 <= ob.MonteCarloPricer().seed < 2**31
    ob.MonteCarloPricer(10, 1, seed=0, method=ob.MCMethod.NUMBA, use_numba=True)  # stale kwarg accepted
    u = ob.MonteCarloPricerUni(num_simulations=5000, num_steps=50, seed=1, use_numba=False, use_gpu=False)
    assert (u.num_simulations, u.num_steps, u.seed) == (5000, 50, 1)
    with pytest.raises(ob.InputValidationError):
        ob.MonteCarloPricerUni(num_simulations=0)
    with pytest.raises(ob.InputValidationError):
        ob.MonteCarloPricerUni(num_steps=-1)
    assert issubclass(ob.InputValidationError, ob.MonteCarloError)


def test_validation_happens_before_any_launch():
    u = ob.MonteCarloPricerUni(100, 5, seed=1)
    for args in [(0, 100, 1.0, 0.05, 0.2, "call"), (100, 100, 1.0, 0.05, -0.2, "call"),
                 (100, 100, 1.0, 0.05, 0.2, "invalid"), (100, 0, 1.0, 0.05, 0.2, "put"), (100, 100, 0.0, 0.05, 0.2, "put")]:
        with pytest.raises(ob.InputValidationError):
            u.price(*args)
    with pytest.raises(ValueError, match="positive"):
        ob.BarrierOption(**P, barrier=0.0).price(10, 2)
    with pytest.raises(ValueError, match="positive"):
        ob.BarrierOption(**P, barrier=-5.0).price(10, 2)


def test_expired_option_is_intrinsic_without_gpu():
    p = ob.MonteCarloPricer(10, 1, seed=1)
    assert p.price(110, 100, 0.0, 0.05, 0.2, "call") == 10
    assert p.price(90, 100, -1.0, 0.05, 0.2, "put") == 10
    r = p.price(110, 100, 0.0, 0.05, 0.2, "put", return_error=True)
    assert (r.price, r.std_error, r.n_paths) == (0, 0.0, 0)


def test_greek_scenarios_and_formulas_match_reference_restatement():
    for (S, K, T, r, sigma, q) in [(100, 100, 1.0, 0.05, 0.2, 0.0), (105, 95, 0.75, 0.03, 0.35, 0.02), (100, 100, 0.002, 0.05, 0.2, 0.0)]:
        for ot in ("call", "put"):
            want = orc.greeks_bump_and_revalue(lambda S_, K_, T_, r_, s_, q_: orc.black_scholes(S_, K_, T_, r_, s_, ot, q_), S, K, T, r, sigma, q)
            plain, fused = BSPricer(), FusedBS()
            got_plain = ob.compute_greeks_unified(plain, S, K, T, r, sigma, ot, q)
            got_fused = ob.compute_greeks_unified(fused, S, K, T, r, sigma, ot, q)
            assert isinstance(got_plain, OrderedDict)
            assert list(got_plain) == list(want) == ["price", "delta", "gamma", "vega", "theta", "rho", "vanna", "charm", "vomma"]
            for k in want:
                assert got_plain[k] == want[k] and got_fused[k] == want[k], k
            n_expected = 14 if T > 1 / 365 else 11
            assert plain.calls == n_expected and fused.fused_sizes == [n_expected] and fused.calls == 0
    first = ob.compute_greeks_unified(BSPricer(), **P, include_second_order=False)
    assert list(first) == ["price", "delta", "gamma", "vega", "theta", "rho"]
    assert len(greeks.greek_scenarios(**P, include_second_order=False)) == 8


def test_greeks_error_wrapping():
    class Broken:
        def price(self, *a, **k):
            raise RuntimeError("boom")

    with pytest.raises(ob.GreeksError, match="Failed to compute unified Greeks: boom"):
        ob.compute_greeks_unified(Broken(), **P)


def test_exotic_adapter_mutates_and_forwards():
    class Fake:
        S = K = T = r = sigma = q = None

        def price(self, n_paths, n_steps, **kw):
            self.seen = (n_paths, n_steps, kw)
            return self.S + self.sigma

    ex = Fake()
    ad = ob.ExoticAdapter(ex, n_paths=123, n_steps=7, avg_type="geometric")
    assert ad.price(101.0, 99.0, 0.5, 0.01, 0.3, "put", q=0.02) == 101.3
    assert (ex.S, ex.K, ex.T, ex.r, ex.sigma, ex.q) == (101.0, 99.0, 0.5, 0.01, 0.3, 0.02)
    assert ex.seen == (123, 7, {"avg_type": "geometric", "option_type": "put"})
    assert isinstance(ad, ob.PricerProtocol) and isinstance(ob.MonteCarloPricer(10), ob.PricerProtocol)
    assert ad.price_scenarios([(100, 100, 1, 0.05, 0.2, 0.0), (101, 100, 1, 0.05, 0.2, 0.0)], "call") == [100.2, 101.2]


def test_moment_arithmetic_matches_reference_formulas():
    rng = np.random.default_rng(0)
    pay = np.maximum(rng.normal(5, 10, 20000), 0)
    m = np.zeros((), dtype=ob._ffi.MOMENTS_DTYPE)
    m["sum"], m["sum_sq"], m["n"] = pay.sum(), (pay**2).sum(), len(pay)
    assert float(runtime.discounted_price(m, 0.05, 1.0)) == pytest.approx(orc.discounted_mean(pay, 0.05, 1.0), rel=1e-14)
    assert float(runtime.discounted_std_error(m, 0.05, 1.0)) == pytest.approx(orc.discounted_std_error(pay, 0.05, 1.0), rel=1e-10)


def test_control_variate_formula_matches_reference_restatement():
    """runtime.control_variate_price over the five sums == np.cov-based reference formula."""
    rng = np.random.default_rng(5)
    S, K, T, r, q = 100.0, 105.0, 0.75, 0.03, 0.01
    terminal = S * np.exp(rng.normal(-0.02, 0.2, 50_000))
    for ot in ("call", "put"):
        pay = orc.vanilla_payoffs(terminal, K, ot)
        m = np.zeros((), dtype=ob._ffi.CV_MOMENTS_DTYPE)
        m["sum_payoff"], m["sum_payoff_sq"], m["sum_terminal"] = pay.sum(), (pay**2).sum(), terminal.sum()
        m["sum_terminal_sq"], m["sum_payoff_terminal"], m["n"] = (terminal**2).sum(), (pay * terminal).sum(), len(pay)
        assert runtime.control_variate_price(m, S, T, r, q) == pytest.approx(
            orc.control_variate_from_terminal(terminal, S, K, T, r, q, ot), rel=1e-9)
    m["sum_terminal"], m["sum_terminal_sq"] = 100.0 * 50_000, 100.0**2 * 50_000  # degenerate control: beta = 0
    assert runtime.control_variate_price(m, S, T, r, q) == pytest.approx(np.exp(-r * T) * m["sum_payoff"] / m["n"], rel=1e-12)


def test_partition_paths_is_an_exact_cover():
    for n in (1, 7, 1000, 16_000_000, 2**40 + 5):
        for w in (1, 2, 3, 4, 8):
            spans = [distributed.partition_paths(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        distributed.partition_paths(10, 2, 2)


def test_closed_form_geometric_asian_matches_reference():
    assert ob.AsianOption(**P).price_geometric_closed_form("call") == pytest.approx(orc.asian_geometric_closed_form(**P), rel=1e-12)
    assert ob.AsianOption(**P, q=0.01).price_geometric_closed_form("put") == pytest.approx(
        orc.asian_geometric_closed_form(**P, q=0.01, option_type="put"), rel=1e-12)


def test_model_adapters_map_arguments_like_the_reference():
    """HestonAdapter: sigma -> v0 = sigma^2 for the call and restored afterwards (unified_greeks.py:96-101);
    JumpDiffusionAdapter forwards (S, K, T, r, sigma, option_type, q) (unified_greeks.py:163-174).  No GPU: fake pricers."""
    calls = []

    class FakeHeston:
        v0 = 0.09

        def price_monte_carlo(self, S, K, T, r, q, option_type, n_paths, n_steps, seed=None):
            calls.append(("mc", S, K, T, r, q, option_type, n_paths, n_steps, seed, self.v0))
            return 1.0

        def price_scenarios(self, sc, option_type, n_paths, n_steps, seed):
            calls.append(("sc", sc, option_type, n_paths, n_steps, seed))
            return [float(i) for i in range(len(sc))]

    h = FakeHeston()
    ad = ob.HestonAdapter(h, n_paths=123, n_steps=7, seed=5)
    assert ad.price(100.0, 90.0, 0.5, 0.01, 0.2, "put", q=0.03) == 1.0
    assert calls[-1] == ("mc", 100.0, 90.0, 0.5, 0.01, 0.03, "put", 123, 7, 5, pytest.approx(0.04)) and h.v0 == 0.09
    out = ad.price_scenarios([(100.0, 90.0, 0.5, 0.01, 0.2, 0.03), (101.0, 90.0, 0.5, 0.01, 0.3, 0.03)], "call")
    assert out == [0.0, 1.0] and calls[-1][1][1] == (101.0, 90.0, 0.5, 0.01, pytest.approx(0.09), 0.03)
    g = ob.compute_greeks_unified(ad, 100.0, 100.0, 1.0, 0.05, 0.2, "call")  # takes the fused route: 14 scenarios, one call
    assert calls[-1][0] == "sc" and len(calls[-1][1]) == 14 and list(g)[:3] == ["price", "delta", "gamma"]

    class FakeJD:
        def price_monte_carlo(self, S, K, T, r, sigma, option_type, q, n_paths, n_steps, seed=None):
            return S - K + sigma + q + n_paths + n_steps + seed

    assert ob.JumpDiffusionAdapter(FakeJD(), 10, 2, 1).price(100.0, 90.0, 1.0, 0.0, 0.25, "call", q=0.5) == 10 + 0.25 + 0.5 + 10 + 2 + 1


def test_structured_products_host_side_contracts():
    """No GPU: the cases CliquetOption / AutocallableOption settle on the host, mirroring exotic_options.py:432-552."""
    P_ = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)
    # more periods than steps: start and end index of every period are column 0, every return is 0 (:532-546)
    assert ob.CliquetOption(**P_, seed=1).price(1000, 4, 12) == 0.0
    assert ob.CliquetOption(**P_, seed=1, global_floor=0.02).price(1000, 4, 12) == pytest.approx(np.exp(-0.05) * 0.02 * 100.0, rel=1e-15)
    res = ob.CliquetOption(**P_, seed=1, global_floor=0.02).price(1000, 4, 12, return_error=True)
    assert (res.std_error, res.n_paths) == (0.0, 1000)
    assert ob.CliquetOption(**P_, seed=1, local_floor=0.01, global_cap=0.05).price_scenarios(
        [(100.0, 100.0, 1.0, 0.05, 0.2, 0.0), (50.0, 100.0, 2.0, 0.0, 0.2, 0.0)], n_steps=4, n_periods=12) == pytest.approx(
        [np.exp(-0.05) * 0.05 * 100.0, 0.05 * 50.0])
    with pytest.raises(ZeroDivisionError):          # n_steps // n_periods
        ob.CliquetOption(**P_, seed=1).price(1000, 12, 0)
    with pytest.raises(ValueError, match="range"):  # range(freq, n_steps + 1, freq)
        ob.AutocallableOption(**P_, seed=1).price(1000, 12, 0)
    a = ob.AutocallableOption(**P_, seed=1)
    assert (a.autocall_barrier, a.coupon_barrier, a.coupon_rate, a.ki_barrier) == (1.0, 0.8, 0.10, 0.6)
    c = ob.CliquetOption(**P_)
    assert (c.local_cap, c.local_floor, c.global_cap, c.global_floor, c.seed) == (0.05, -0.05, 0.30, 0.0, None)


def test_monte_carlo_convergence_study_contract():
    """validation.py:202-239: sizes 1x/2x/4x/10x, population std per size, 1/sqrt(n) law anchored at the first size,
    `converging` = no size's spread exceeds 1.5x the previous one."""
    rng = np.random.default_rng(3)
    calls = []

    def noisy(n):
        calls.append(n)
        return 10.0 + rng.standard_normal() / np.sqrt(n)

    out = ob.monte_carlo_convergence_test(noisy, n_trials=9, base_sims=100)
    assert list(out["results"]) == [100, 200, 400, 1000] and calls == [100] * 9 + [200] * 9 + [400] * 9 + [1000] * 9
    assert set(out["results"][100]) == {"mean", "std", "min", "max"} and len(out["stds"]) == 4
    assert out["expected_rate"] == pytest.approx([out["stds"][0] * np.sqrt(100 / n) for n in (100, 200, 400, 1000)])
    grow = iter([0.0, 1.0] * 2 + [0.0, 10.0] * 6)
    assert ob.monte_carlo_convergence_test(lambda n: next(grow), n_trials=4, base_sims=10)["converging"] is False
    assert ob.monte_carlo_convergence_test(lambda n: 1.0, n_trials=3, base_sims=10)["converging"] is True


# ---- the launch planner (a pure host function of the library: b200mc_plan_tiles) --------------------------------------
def test_tile_plans_cover_the_paths_and_follow_the_measured_rules():
    from optionslab_b200 import _ffi

    euro = _ffi.make_spec(_ffi.EUROPEAN, 252, antithetic=True)
    asian = _ffi.make_spec(_ffi.ASIAN_ARITH, 252)
    for spec, n_opt, n_scen, n_paths in [(euro, 1, 1, 1), (euro, 1, 1, 257), (euro, 1, 1, 10_000), (euro, 1, 1, 100_000), (euro, 1, 14, 1_000_000),
                                         (euro, 4096, 1, 1_000_000), (euro, 4096, 3, 100_000), (asian, 1, 1, 4_000_000), (asian, 1, 14, 200_000),
                                         (_ffi.make_spec(_ffi.BARRIER, 365), 1, 1, 16_000_000), (euro, 1, 1, 2**33)]:
        plan = _ffi.plan_tiles(spec, n_opt, n_scen, n_paths)
        per_tile = (256 >> plan["split_shift"]) * plan["paths_per_thread"]
        assert 1 <= plan["paths_per_thread"] <= 32 and 0 <= plan["split_shift"] <= 3
        assert plan["tiles"] * per_tile >= n_paths > (plan["tiles"] - 1) * per_tile          # every path in exactly one tile, no empty tile
        assert plan == _ffi.plan_tiles(spec, n_opt, n_scen, n_paths)                           # a pure function: same call, same shape
        if spec.kind != _ffi.EUROPEAN:
            assert plan["split_shift"] == 0                                                    # path-dependent state: one thread per path
    # lane split only while one thread per path leaves most SMs without a CTA (profiles/r02_plan_sweep.jsonl)
    assert _ffi.plan_tiles(euro, 1, 1, 10_000)["split_shift"] >= 1
    assert _ffi.plan_tiles(euro, 1, 1, 1_000)["split_shift"] == 3
    assert _ffi.plan_tiles(euro, 1, 1, 30_000)["split_shift"] == 0 and _ffi.plan_tiles(euro, 1, 1, 100_000)["split_shift"] == 0
    assert _ffi.plan_tiles(euro, 64, 1, 1_000)["split_shift"] == 0                             # 64 options fill the chip by themselves
    assert _ffi.plan_tiles(euro, 1, 1, 10_000, control_variate=True)["split_shift"] == 0       # the control-variate kernel has no split form
    assert _ffi.plan_tiles(_ffi.make_spec(_ffi.EUROPEAN, 8, antithetic=True), 1, 1, 1_000)["split_shift"] == 0  # one Philox call: nothing to split
    # 8-16 scenario European launches (two 8-warp CTAs per SM): one wave of the 2 x 148 CTA slots when the paths allow it, one CTA
    # per SM for the reference's own sizes (profiles/r02_wide_plan_check.jsonl)
    assert _ffi.plan_tiles(euro, 1, 14, 1_000_000)["tiles"] <= 296 and _ffi.plan_tiles(euro, 1, 8, 2_000_000)["tiles"] <= 296
    assert _ffi.plan_tiles(euro, 1, 14, 100_000)["tiles"] <= 148 and _ffi.plan_tiles(euro, 1, 14, 100_000)["split_shift"] == 0
    # large grids amortise the per-CTA work over many paths per thread; the shape scales with the SM count
    assert _ffi.plan_tiles(euro, 4096, 1, 1_000_000)["paths_per_thread"] >= 16
    small, big = _ffi.plan_tiles(euro, 1, 1, 4_000, sm_count=16), _ffi.plan_tiles(euro, 1, 1, 4_000, sm_count=148)
    assert small["split_shift"] <= big["split_shift"]
    with pytest.raises(Exception):
        _ffi.plan_tiles(euro, 1, 17, 1000)


# ---- bench.py helpers that the driver's bench run depends on (no GPU) ---------------------------------------------------------
def test_bench_reads_roofline_traffic_from_the_committed_capture_and_its_budgets_match_the_shipped_sass(tmp_path):
    import importlib.util
    import os
    import re
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    traffic, source = bench.traffic_from_capture()
    assert source == os.path.join("profiles", "r02_ncu_european.txt")
    assert 360_448 <= traffic < 8 * 360_448            # at least the algorithmic bytes (parameters in, moments out), same order
    with pytest.raises(SystemExit):
        bench.traffic_from_capture(str(tmp_path / "missing.txt"))
    (tmp_path / "empty.txt").write_text("== void other_kernel()\ndram__bytes_read.sum 1 byte\n")
    with pytest.raises(SystemExit):
        bench.traffic_from_capture(str(tmp_path / "empty.txt"))
    # the per-path-step budget bench.py multiplies the kernel rate by is the one of the SHIPPED binary (tools/sass_loop.py)
    import shutil

    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH: SASS budget not re-counted")
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "sass_loop.py"), "european_kernelILi1ELb1ELi6ELb0ELi1ELb0"],
                         capture_output=True, text=True).stdout
    m = re.search(r"loop 0x[0-9a-f]+\.\.0x[0-9a-f]+: (\d+) instructions, (\d+) MUFU", out)
    assert m, out
    assert int(m.group(1)) / 8 == pytest.approx(bench.INSTR_PER_STEP) and int(m.group(2)) / 8 == pytest.approx(bench.MUFU_PER_STEP)
