"""GPU, correctness test 2 of BASELINE.json: with the on-device Philox stream, prices and Greeks
fall within 3 standard errors of the reference and of the Black-Scholes closed form — plus the
exact properties the fused FP32 path must have (determinism, shard additivity, fused == separate).

The FP32 path is also compared, draw for draw, against the FP64 oracle evaluation of the SAME
Philox stream (oracle/philox_oracle.c normals through oracle/reference_mc.py): that isolates
arithmetic error (tolerance 2e-4 relative on a price, i.e. far below one standard error) from
Monte Carlo noise.
"""

import numpy as np
import pytest

import optionslab_b200 as ob
from optionslab_b200 import _ffi, runtime
from oracle import philox_oracle as po
from oracle import reference_mc as orc

pytestmark = pytest.mark.gpu
P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)
BS_CALL, BS_PUT = 10.450583572185565, 5.573526022256971
# reference values at 1M paths (recorded from the real reference, NumPy 2.3.5)
REF_1M = dict(call=(10.457915429444164, 0.010426610119053227), delta=0.6370571458618244, gamma=0.019004492473989387,
              vega=37.564849270217415, asian=5.768271672258132, up_and_out=1.2995964895711083, up_and_in=9.13740139898588)


def _price_se(m, r=P["r"], T=P["T"]):
    return float(runtime.discounted_price(m, r, T)), float(runtime.discounted_std_error(m, r, T))


# ---------------- arithmetic accuracy against the FP64 oracle on the same draws -------------------

@pytest.mark.parametrize("n_steps", [1, 3, 16, 50, 253])
def test_european_fp32_path_vs_fp64_oracle_same_stream(engine, n_steps):
    n, seed = 20_000, 31
    Z = po.normals(seed, n, n_steps)
    for ot, q in (("call", 0.0), ("put", 0.02)):
        pay = orc.vanilla_payoffs(orc.gbm_terminal_from_normals(P["S"], P["T"], P["r"], P["sigma"], q, Z), P["K"], ot)
        spec = _ffi.make_spec(_ffi.EUROPEAN, n_steps, is_put=(ot == "put"), antithetic=True)
        m = engine.simulate(spec, _ffi.make_params(**P, q=q).reshape(1, 1), seed, n)[0, 0]
        assert m["n"] == 2 * n
        assert m["sum"] == pytest.approx(pay.sum(), rel=2e-4)
        assert m["sum_sq"] == pytest.approx((pay**2).sum(), rel=4e-4)


def test_exotics_fp32_path_vs_fp64_oracle_same_stream(engine):
    n, n_steps, seed = 20_000, 37, 5  # 37 = 9 Philox blocks + 1 leftover draw
    Z = po.normals(seed, n, n_steps)
    paths = orc.exotic_paths_from_normals(P["S"], P["T"], P["r"], P["sigma"], 0.0, Z)
    cases = []
    for ot in ("call", "put"):
        cases += [(_ffi.make_spec(_ffi.ASIAN_ARITH, n_steps, is_put=ot == "put"), 0.0, orc.asian_payoffs(paths, P["K"], "arithmetic", ot)),
                  (_ffi.make_spec(_ffi.ASIAN_GEOM, n_steps, is_put=ot == "put"), 0.0, orc.asian_payoffs(paths, P["K"], "geometric", ot))]
        for B, bts in ((115.0, ("up-and-out", "up-and-in")), (90.0, ("down-and-out", "down-and-in"))):
            for bt in bts:
                cases.append((_ffi.make_spec(_ffi.BARRIER, n_steps, is_put=ot == "put", barrier_down=bt.startswith("down"),
                                             barrier_in=bt.endswith("in")), B, orc.barrier_payoffs(paths, P["K"], B, bt, ot)))
        for lt in ("floating", "fixed"):
            cases.append((_ffi.make_spec(_ffi.LOOKBACK, n_steps, is_put=ot == "put", lookback_fixed=lt == "fixed"), 0.0,
                          orc.lookback_payoffs(paths, P["K"], lt, ot)))
    for spec, B, pay in cases:
        m = engine.simulate(spec, _ffi.make_params(**P, barrier=B).reshape(1, 1), seed, n)[0, 0]
        assert m["n"] == n
        # barrier: a path within FP32 rounding of B may flip; allow a few payoffs' worth of slack
        assert m["sum"] == pytest.approx(pay.sum(), rel=5e-4, abs=60.0 if spec.kind == _ffi.BARRIER else 0.0), (spec.kind, spec.is_put)


# ---------------- statistical agreement with the reference and Black-Scholes ----------------------

def test_config1_european_100k_x_252_vs_reference_and_black_scholes():
    res = ob.MonteCarloPricer(100_000, 252, seed=42).price(**P, option_type="call", return_error=True)
    assert res.n_paths == 200_000 and 0 < res.std_error < res.price
    ref_price, ref_se = 10.46001409391075, 0.03303353316736617  # real reference, same config (goldens)
    assert res.std_error == pytest.approx(ref_se, rel=0.02)
    assert abs(res.price - BS_CALL) <= 3 * res.std_error
    assert abs(res.price - ref_price) <= 3 * np.hypot(res.std_error, ref_se)
    put = ob.MonteCarloPricer(100_000, 252, seed=42).price(**P, option_type="put", return_error=True)
    assert abs(put.price - BS_PUT) <= 3 * put.std_error


def test_single_step_default_matches_black_scholes():
    res = ob.MonteCarloPricer(1_000_000, seed=1).price(**P, option_type="call", return_error=True)
    assert abs(res.price - BS_CALL) <= 3 * res.std_error
    fast = ob.MonteCarloPricer(1_000_000, 50, seed=1, method=ob.MCMethod.FAST).price(**P, option_type="call")
    assert fast == res.price  # FAST forces one step (monte_carlo.py:86-92)


def test_config2_greeks_1m_x_252_crn_single_launch(engine):
    pr = ob.MonteCarloPricer(1_000_000, 252, seed=42)
    before = engine.kernel_launches()
    g = pr.greeks(**P, option_type="call")
    assert engine.kernel_launches() - before == 1  # ONE launch: 14 scenarios simulated and folded by the same kernel
    d_bs, g_bs, v_bs = orc.black_scholes_greeks(**P)
    assert g["price"] == pytest.approx(BS_CALL, abs=3 * 0.0105)
    assert g["delta"] == pytest.approx(d_bs, abs=1.5e-3) and g["delta"] == pytest.approx(REF_1M["delta"], abs=2e-3)
    assert g["gamma"] == pytest.approx(g_bs, abs=1e-3) and g["gamma"] == pytest.approx(REF_1M["gamma"], abs=1e-3)
    assert g["vega"] == pytest.approx(v_bs, abs=0.3) and g["vega"] == pytest.approx(REF_1M["vega"], abs=0.4)
    assert g["theta"] == pytest.approx(-6.414, abs=0.15) and g["rho"] == pytest.approx(53.232, abs=0.5)
    assert list(g) == ["price", "delta", "gamma", "vega", "theta", "rho", "vanna", "charm", "vomma"]
    p = pr.greeks(**P, option_type="put")
    assert p["delta"] == pytest.approx(d_bs - 1.0, abs=1.5e-3) and p["gamma"] == pytest.approx(g["gamma"], rel=2e-3)  # equal in exact arithmetic; FP32 noise / h_S^2
    assert pr.delta(**P) == g["delta"] and pr.vega(**P) == g["vega"]


def test_control_variate_next_row(engine):
    """price_with_control_variate (monte_carlo.py:154-186): same draws => FP32 path vs FP64 oracle on the engine's
    own stream; statistically => closer to Black-Scholes than the plain estimator's noise."""
    n, steps, seed = 50_000, 20, 13
    pr = ob.MonteCarloPricer(n, steps, seed=seed)
    for ot, bs in (("call", BS_CALL), ("put", BS_PUT)):
        got = pr.price_with_control_variate(**P, option_type=ot)
        terminal = orc.gbm_terminal_from_normals(P["S"], P["T"], P["r"], P["sigma"], 0.0, po.normals(seed, n, steps))
        want = orc.control_variate_from_terminal(terminal, P["S"], P["K"], P["T"], P["r"], 0.0, ot)
        assert got == pytest.approx(want, rel=3e-4)
        plain = pr.price(**P, option_type=ot, return_error=True)
        assert abs(got - bs) < 4 * plain.std_error
    big = ob.MonteCarloPricer(2_000_000, 50, seed=1)
    cv = big.price_with_control_variate(**P, option_type="call")
    assert abs(cv - BS_CALL) < 0.015  # plain SE at this size is 0.0074; the control removes most of it
    m = engine.simulate(_ffi.make_spec(_ffi.EUROPEAN, 8, antithetic=True), _ffi.make_params(**P).reshape(1, 1), 3, 10_000, control_variate=True)[0, 0]
    m2 = engine.simulate(_ffi.make_spec(_ffi.EUROPEAN, 8, antithetic=True), _ffi.make_params(**P).reshape(1, 1), 3, 10_000)[0, 0]
    assert m["n"] == 20_000 and m["sum_payoff"] == m2["sum"] and m["sum_payoff_sq"] == m2["sum_sq"]
    assert m["sum_terminal"] / m["n"] == pytest.approx(100 * np.exp(0.05), rel=5e-3)
    with pytest.raises(ob.MonteCarloError, match="European"):
        engine.simulate(_ffi.make_spec(_ffi.ASIAN_ARITH, 8), _ffi.make_params(**P).reshape(1, 1), 3, 100, control_variate=True)


def test_fused_greeks_equal_separate_repricings_bitwise(engine):
    """One launch with 14 scenarios == 14 launches of 1 scenario (same draws, same per-scenario expression shape, same
    reduction order).  Bit-for-bit equality needs both launches to cut the paths into the same tiles - the planner sizes
    tiles from the scenario count too - so the tile shape is pinned for that half of the test; with the automatic plan
    the two routes agree to FP32 summation order."""
    pr = ob.MonteCarloPricer(50_000, 32, seed=9)

    class CallByCall:  # hides price_scenarios: forces the reference's route
        def price(self, *a, **k):
            return pr.price(*a, **k)

    for shift, ppt in ((0, 1), (1, 3), (2, 7)):
        engine.set_plan(shift, ppt)
        try:
            fused = ob.compute_greeks_unified(pr, **P, option_type="put", q=0.01)
            separate = ob.compute_greeks_unified(CallByCall(), **P, option_type="put", q=0.01)
        finally:
            engine.set_plan()
        assert dict(fused) == dict(separate)
    fused = ob.compute_greeks_unified(pr, **P, option_type="put", q=0.01)
    separate = ob.compute_greeks_unified(CallByCall(), **P, option_type="put", q=0.01)
    assert fused["price"] == pytest.approx(separate["price"], rel=1e-6)
    assert fused["delta"] == pytest.approx(separate["delta"], abs=1e-5) and fused["vega"] == pytest.approx(separate["vega"], abs=2e-3)


def test_lane_split_plans_price_the_same_draws(engine):
    """european_kernel<SPLIT>: 2, 4 or 8 adjacent lanes share one path's Philox calls and complete W = sum(z) with a
    shuffle butterfly.  Same stream contract, so every tile shape agrees with the one-thread-per-path launch to FP32
    summation order and with the FP64 oracle on the same draws - ragged path / step counts included."""
    for n_paths, n_steps, seed in ((10_000, 252, 3), (1_237, 100, 4), (100_000, 33, 5), (257, 1000, 6), (5_000, 31, 7)):
        Z = po.normals(seed, n_paths, n_steps)
        for ot in ("call", "put"):
            want = orc.vanilla_payoffs(orc.gbm_terminal_from_normals(P["S"], P["T"], P["r"], P["sigma"], 0.0, Z), P["K"], ot)
            spec = _ffi.make_spec(_ffi.EUROPEAN, n_steps, is_put=ot == "put", antithetic=True)
            params = np.stack([_ffi.make_params(**P), _ffi.make_params(**dict(P, sigma=0.21))]).reshape(1, 2)
            base = None
            for shift in (0, 1, 2, 3):
                for ppt in (0, 1, 5):
                    engine.set_plan(shift, ppt)
                    try:
                        m = engine.simulate(spec, params, seed, n_paths)[0]
                    finally:
                        engine.set_plan()
                    assert m["n"][0] == 2 * n_paths
                    assert m["sum"][0] == pytest.approx(want.sum(), rel=3e-4)
                    if base is None:
                        base = m
                    assert m["sum"] == pytest.approx(base["sum"], rel=2e-6) and m["sum_sq"] == pytest.approx(base["sum_sq"], rel=4e-6)
    # more than 1024 tiles per option: the fold takes its second level (groups of 1024 tiles, then the group totals)
    n_paths, n_steps, seed = 300_000, 16, 8
    Z = po.normals(seed, n_paths, n_steps)
    want = orc.vanilla_payoffs(orc.gbm_terminal_from_normals(P["S"], P["T"], P["r"], P["sigma"], 0.0, Z), P["K"], "call")
    spec = _ffi.make_spec(_ffi.EUROPEAN, n_steps, antithetic=True)
    many = np.stack([_ffi.make_params(**dict(P, sigma=0.2 + 0.0 * k)) for k in range(14)]).reshape(1, 14)
    for params in (_ffi.make_params(**P).reshape(1, 1), many):
        engine.set_plan(0, 1)
        try:
            fine = engine.simulate(spec, params, seed, n_paths)[0]
            assert engine.last_plan()["tiles"] == 1172
        finally:
            engine.set_plan()
        auto = engine.simulate(spec, params, seed, n_paths)[0]
        assert np.all(fine["n"] == 2 * n_paths)
        np.testing.assert_allclose(fine["sum"], want.sum(), rtol=3e-4)
        np.testing.assert_allclose(fine["sum"], auto["sum"], rtol=2e-6)
        np.testing.assert_allclose(fine["sum_sq"], auto["sum_sq"], rtol=4e-6)
        assert np.all(fine["sum"] == fine["sum"][0])  # identical scenarios of one launch: identical bits
    # the automatic plan of an under-filled launch does split, and repeats bit for bit
    spec = _ffi.make_spec(_ffi.EUROPEAN, 252, antithetic=True)
    a = engine.simulate(spec, _ffi.make_params(**P).reshape(1, 1), 1, 10_000)
    b = engine.simulate(spec, _ffi.make_params(**P).reshape(1, 1), 1, 10_000)
    assert a.tobytes() == b.tobytes()


def test_launches_on_different_streams_do_not_share_scratch_unordered(engine):
    """b200mc_simulate_device on two caller streams: consecutive enqueues share the tile partials and the
    ticket counters, so the engine orders each enqueue after the previous one; interleaved results equal the
    one-stream results bit for bit."""
    import torch

    dev = torch.device("cuda", 0)
    specs = [_ffi.make_spec(_ffi.EUROPEAN, 64, antithetic=True), _ffi.make_spec(_ffi.ASIAN_ARITH, 48)]
    n_opt = 96
    K = np.linspace(80.0, 120.0, n_opt)
    params = torch.from_numpy(_ffi.make_params(100.0, K, 1.0, 0.05, 0.2).view(np.float64).reshape(n_opt, 8).copy()).to(dev)
    main = torch.cuda.current_stream(dev)

    def run(streams):
        outs = [torch.zeros((n_opt, 3), dtype=torch.float64, device=dev) for _ in range(6)]
        for i, out in enumerate(outs):
            st = streams[i % len(streams)]
            engine.simulate_device(specs[i % 2], params.data_ptr(), n_opt, 1, 7 + i, 40_000 + 1000 * i, out.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize(dev)
        return [o.cpu().numpy() for o in outs]

    want = run([main])
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for _ in range(3):
        got = run([s1, s2])
        for g, w in zip(got, want):
            assert g.tobytes() == w.tobytes()


def test_config3_asian_4m_x_252():
    res = ob.AsianOption(**P, seed=42).price(n_paths=4_000_000, n_steps=252, return_error=True)
    ref_se_1m = 8.0 / np.sqrt(1e6)  # payoff std ~ 8 => reference's own 1M-path noise
    assert abs(res.price - REF_1M["asian"]) <= 3 * np.hypot(res.std_error, ref_se_1m)
    assert res.n_paths == 4_000_000 and res.price < BS_CALL
    geo = ob.AsianOption(**P, seed=42).price(n_paths=1_000_000, n_steps=252, avg_type="geometric", return_error=True)
    assert abs(geo.price - 5.552463811592335) <= 3 * np.hypot(geo.std_error, 0.008)
    assert geo.price == pytest.approx(ob.AsianOption(**P).price_geometric_closed_form(), rel=0.05)
    assert geo.price < res.price  # AM-GM


def test_config4_barrier_16m_x_365_and_in_out_parity(engine):
    opt = ob.BarrierOption(**P, seed=42, barrier=120.0)
    out = opt.price(16_000_000, 365, "up-and-out", "call", return_error=True)
    assert abs(out.price - REF_1M["up_and_out"]) <= 3 * np.hypot(out.std_error, 0.0035)
    inn = opt.price(16_000_000, 365, "up-and-in", "call", return_error=True)
    assert abs(inn.price - REF_1M["up_and_in"]) <= 3 * np.hypot(inn.std_error, 0.016)
    # same draws: knock-in + knock-out is the vanilla payoff path by path
    never = ob.BarrierOption(**P, seed=42, barrier=1e9).price(16_000_000, 365, "up-and-out", "call", return_error=True)
    assert out.price + inn.price == pytest.approx(never.price, rel=1e-5)
    assert abs(never.price - BS_CALL) <= 3 * never.std_error


def test_same_seed_is_bit_identical_and_seeds_differ():
    a = ob.MonteCarloPricer(100_000, 50, seed=42).price(**P, option_type="call")
    b = ob.MonteCarloPricer(100_000, 50, seed=42).price(**P, option_type="call")
    c = ob.MonteCarloPricer(100_000, 50, seed=43).price(**P, option_type="call")
    assert a == b and a != c
    pr = ob.MonteCarloPricer(100_000, 50, seed=1)
    assert pr.price(**P, option_type="call", seed=42) == a  # per-call override, monte_carlo.py:84
    assert ob.MonteCarloPricer(100_000, 50, seed=0).price(**P, option_type="call") != a  # seed=0 honoured
    x = ob.AsianOption(**P, seed=7).price(50_000, 64)
    assert x == ob.AsianOption(**P, seed=7).price(50_000, 64)
    assert ob.AsianOption(**P).price(50_000, 64) != ob.AsianOption(**P).price(50_000, 64)  # unseeded


def test_path_ranges_are_additive_the_multi_gpu_contract(engine):
    """moments([0,N)) == sum of moments over any partition of [0,N) (up to FP64 summation order)."""
    spec = _ffi.make_spec(_ffi.EUROPEAN, 20, antithetic=True)
    params = np.stack([_ffi.make_params(100.0, K, 1.0, 0.05, 0.2) for K in (90.0, 100.0, 110.0)]).reshape(3, 1)
    N = 300_001
    whole = engine.simulate(spec, params, 5, N)
    for world in (2, 3, 8):
        acc = np.zeros((3, 1, 3))
        for rank in range(world):
            b, c = ob.distributed.partition_paths(N, rank, world)
            part = engine.simulate(spec, params, 5, c, path_begin=b)
            acc += np.stack([part["sum"], part["sum_sq"], part["n"]], axis=-1)
        np.testing.assert_allclose(acc[..., 0], whole["sum"], rtol=1e-6)  # per-thread FP32 partial sums regroup
        np.testing.assert_allclose(acc[..., 1], whole["sum_sq"], rtol=1e-6)
        np.testing.assert_array_equal(acc[..., 2], whole["n"])


def test_scaling_homogeneity_is_exact(engine):
    """price(aS, aK) = a * price(S, K): the kernels work on S_t/S_0, so this holds to FP64 rounding."""
    spec = _ffi.make_spec(_ffi.ASIAN_ARITH, 16)
    base = engine.simulate(spec, _ffi.make_params(100.0, 95.0, 1.0, 0.05, 0.3).reshape(1, 1), 3, 100_000)[0, 0]
    scaled = engine.simulate(spec, _ffi.make_params(400.0, 380.0, 1.0, 0.05, 0.3).reshape(1, 1), 3, 100_000)[0, 0]
    assert scaled["sum"] == pytest.approx(4 * base["sum"], rel=1e-14)


def test_ragged_and_tiny_sizes(engine):
    for n_paths in (1, 2, 31, 255, 257, 1025):
        for n_steps in (1, 2, 3, 4, 5):
            spec = _ffi.make_spec(_ffi.EUROPEAN, n_steps, antithetic=True)
            m = engine.simulate(spec, _ffi.make_params(**P).reshape(1, 1), 1, n_paths)[0, 0]
            Z = po.normals(1, n_paths, n_steps)
            pay = orc.vanilla_payoffs(orc.gbm_terminal_from_normals(P["S"], P["T"], P["r"], P["sigma"], 0.0, Z), P["K"], "call")
            assert m["n"] == 2 * n_paths
            assert m["sum"] == pytest.approx(pay.sum(), rel=3e-4, abs=1e-3)


# ---------------- arithmetic Asian: multiplicative small-move update vs the MUFU.EX2 form --------

def test_asian_small_move_update_agrees_with_exact_ex2_form_and_fp64_oracle(engine):
    """mc_kernels.cuh picks, per option, s_t = s_{t-1} + s_{t-1}*(2^x - 1) with a degree-5 polynomial when every
    possible per-step log2 move is <= 0.25; B200MC_FLAG_EXACT_EX2 forces l_t += x, S_t = 2^l_t.  Same draws:
    the two agree to FP32 rounding, and both sit within 2e-4 of the FP64 oracle evaluation of the stream."""
    n, n_steps, seed = 50_000, 252, 17
    Z = po.normals(seed, n, n_steps)
    for sigma in (0.2, 0.45):  # 0.45: a 5.65-sigma daily draw moves log2 S by 0.23 (just inside the bound)
        p = dict(P, sigma=sigma)
        paths = orc.exotic_paths_from_normals(p["S"], p["T"], p["r"], sigma, 0.0, Z)
        for ot in ("call", "put"):
            want = orc.asian_payoffs(paths, p["K"], "arithmetic", ot)
            fast = engine.simulate(_ffi.make_spec(_ffi.ASIAN_ARITH, n_steps, is_put=ot == "put"), _ffi.make_params(**p).reshape(1, 1), seed, n)[0, 0]
            exact = engine.simulate(_ffi.make_spec(_ffi.ASIAN_ARITH, n_steps, is_put=ot == "put", exact_ex2=True),
                                    _ffi.make_params(**p).reshape(1, 1), seed, n)[0, 0]
            assert fast["sum"] == pytest.approx(exact["sum"], rel=2e-5)
            assert fast["sum_sq"] == pytest.approx(exact["sum_sq"], rel=4e-5)
            assert fast["sum"] == pytest.approx(want.sum(), rel=2e-4)
            assert exact["sum"] == pytest.approx(want.sum(), rel=2e-4)


def test_asian_large_moves_take_the_ex2_form_bit_for_bit(engine):
    """Above the bound (here sigma = 1.5 on 12 monthly steps) the kernel must not use the polynomial:
    default and EXACT_EX2 launches are the same code path, hence identical bits; a launch whose scenarios
    straddle the bound takes the EX2 form for all of them."""
    n, n_steps, seed = 30_000, 12, 4
    big = dict(P, sigma=1.5)
    a = engine.simulate(_ffi.make_spec(_ffi.ASIAN_ARITH, n_steps), _ffi.make_params(**big).reshape(1, 1), seed, n)[0, 0]
    b = engine.simulate(_ffi.make_spec(_ffi.ASIAN_ARITH, n_steps, exact_ex2=True), _ffi.make_params(**big).reshape(1, 1), seed, n)[0, 0]
    assert a["sum"] == b["sum"] and a["sum_sq"] == b["sum_sq"]
    Z = po.normals(seed, n, n_steps)
    want = orc.asian_payoffs(orc.exotic_paths_from_normals(big["S"], big["T"], big["r"], big["sigma"], 0.0, Z), big["K"], "arithmetic", "call")
    assert a["sum"] == pytest.approx(want.sum(), rel=2e-4)
    # scenarios straddling the bound: sigma = 0.2 (small moves) next to sigma = 1.5 in one launch
    both = _ffi.make_params([P["S"]] * 2, P["K"], P["T"], P["r"], [0.2, 1.5]).reshape(1, 2)
    m = engine.simulate(_ffi.make_spec(_ffi.ASIAN_ARITH, n_steps), both, seed, n)[0]
    solo = engine.simulate(_ffi.make_spec(_ffi.ASIAN_ARITH, n_steps, exact_ex2=True), _ffi.make_params(**P).reshape(1, 1), seed, n)[0, 0]
    assert m[0]["sum"] == solo["sum"] and m[1]["sum"] == a["sum"]


def test_asian_fused_scenarios_equal_separate_repricings_bitwise():
    opt = ob.AsianOption(**P, seed=11)
    sc = [(100.0, 100.0, 1.0, 0.05, 0.2, 0.0), (101.0, 100.0, 1.0, 0.05, 0.2, 0.0), (100.0, 100.0, 1.0, 0.05, 0.21, 0.0),
          (100.0, 100.0, 1.0 - 1 / 365, 0.0501, 0.2, 0.01)]
    fused = opt.price_scenarios(sc, n_paths=40_000, n_steps=64)
    separate = [opt.price_scenarios([s], n_paths=40_000, n_steps=64)[0] for s in sc]
    assert fused == separate


# ---------------- C5 grid at full size: z-scores against Black-Scholes over all 4096 options -----

def test_config5_grid_z_scores_are_standard_normal_where_the_clt_applies(engine):
    """BASELINE.json configs[4] at full size (4096 options x 1M antithetic pairs x 252 steps, 0.5 s on a B200).
    Options with >= 2000 expected in-the-money samples are in the Gaussian regime: their z-scores against the
    closed form must look like N(0, <=1) (the reference's std error treats mirrored pairs as independent, which
    over-states it in the money, so the spread may be below 1 but not above), with no option beyond 4.75 sigma
    (two-sided tail probability 2e-6 per option).  Everywhere else |price - BS| <= 5 se + 1e-5."""
    from math import erf, exp, log, sqrt

    K, T = np.meshgrid(np.linspace(60.0, 140.0, 64), np.linspace(1.0 / 12.0, 2.0, 64), indexing="ij")
    K, T = K.ravel(), T.ravel()
    n_opt, n_paths, n_steps = K.size, 1_000_000, 252
    S, r, sigma = 100.0, 0.05, 0.2
    params = _ffi.make_params(S, K, T, r, sigma).reshape(n_opt, 1)
    m = engine.simulate(_ffi.make_spec(_ffi.EUROPEAN, n_steps, antithetic=True), params, 42, n_paths)[:, 0]
    price, se = runtime.discounted_price(m, r, T), runtime.discounted_std_error(m, r, T)
    cdf = lambda x: 0.5 * (1.0 + erf(x / sqrt(2.0)))
    d2 = np.array([(log(S / k) + (r - 0.5 * sigma**2) * t) / (sigma * sqrt(t)) for k, t in zip(K, T)])
    bs = np.array([S * cdf(d + sigma * sqrt(t)) - k * exp(-r * t) * cdf(d) for d, k, t in zip(d2, K, T)])
    assert np.all(np.abs(price - bs) <= 5.0 * se + 1e-5)
    gaussian = np.array([cdf(d) for d in d2]) * 2 * n_paths >= 2000
    assert gaussian.sum() > 3500
    z = (price - bs)[gaussian] / se[gaussian]
    assert abs(z.mean()) < 4.0 / np.sqrt(z.size), z.mean()
    assert 0.5 < z.std() < 1.08, z.std()
    assert np.max(np.abs(z)) < 4.75, np.max(np.abs(z))


def test_martingale_and_variance_of_the_terminal_spot_over_maturities_and_step_counts(engine):
    """E[S_T] = S e^{(r-q)T} and Var[ln S_T] = sigma^2 T for every maturity and for step counts that are not multiples of the
    8-step Philox block: a call struck at ~0 prices the discounted forward (first moment), and the control-variate
    moments give Var[S_T] = F^2 (e^{sigma^2 T} - 1) (second moment) - both within 4 standard errors, 40 cases."""
    n = 2_000_000
    for n_steps in (1, 7, 12, 252, 365):
        T = np.linspace(0.25, 2.0, 8)
        S, r, q, sigma = 100.0, 0.05, 0.015, 0.3
        params = _ffi.make_params(S, 1e-9, T, r, sigma, q).reshape(len(T), 1)
        m = engine.simulate(_ffi.make_spec(_ffi.EUROPEAN, n_steps, antithetic=False), params, 77, n, control_variate=True)[:, 0]
        fwd = S * np.exp((r - q) * T)
        mean = m["sum_terminal"] / m["n"]
        var = m["sum_terminal_sq"] / m["n"] - mean**2
        se = np.sqrt(var / m["n"])
        assert np.all(np.abs(mean - fwd) < 4 * se), (n_steps, (mean - fwd) / se)
        want_var = fwd**2 * (np.exp(sigma**2 * T) - 1)
        # std error of a sample variance of a lognormal: sqrt((mu4 - var^2)/n); bound mu4 by the lognormal's fourth central moment
        w = np.exp(sigma**2 * T)
        mu4 = fwd**4 * (w - 1) ** 2 * (w**4 + 2 * w**3 + 3 * w**2 - 3)
        assert np.all(np.abs(var - want_var) < 4.5 * np.sqrt((mu4 - want_var**2) / m["n"])), (n_steps, var / want_var)
        assert np.allclose(m["sum_payoff"], m["sum_terminal"], rtol=1e-6)  # K ~ 0: the payoff is the terminal spot


def test_discrete_geometric_asian_matches_its_exact_lognormal_law(engine):
    """The discrete geometric average of a GBM is exactly lognormal: ln G ~ N(ln S + (r-q-sigma^2/2) T (n+1)/(2n),
    sigma^2 T (n+1)(2n+1)/(6 n^2)).  The path-dependent kernel (ASIAN_GEOM, running sum of l_t in registers) must price
    within 4 standard errors of that closed form for calls and puts over strikes, maturities and step counts."""
    from math import erf, exp, log, sqrt

    cdf = lambda x: 0.5 * (1.0 + erf(x / sqrt(2.0)))
    n_paths = 4_000_000
    for n_steps, T, sigma, q in ((252, 1.0, 0.2, 0.0), (12, 0.5, 0.35, 0.02), (37, 2.0, 0.15, 0.01)):
        S, r = 100.0, 0.05
        mu = log(S) + (r - q - 0.5 * sigma**2) * T * (n_steps + 1) / (2 * n_steps)
        var = sigma**2 * T * (n_steps + 1) * (2 * n_steps + 1) / (6 * n_steps**2)
        for K in (90.0, 100.0, 115.0):
            d2 = (mu - log(K)) / sqrt(var)
            d1 = d2 + sqrt(var)
            call = exp(-r * T) * (exp(mu + 0.5 * var) * cdf(d1) - K * cdf(d2))
            put = exp(-r * T) * (K * cdf(-d2) - exp(mu + 0.5 * var) * cdf(-d1))
            for ot, want in (("call", call), ("put", put)):
                m = engine.simulate(_ffi.make_spec(_ffi.ASIAN_GEOM, n_steps, is_put=ot == "put"), _ffi.make_params(S, K, T, r, sigma, q).reshape(1, 1),
                                    2024 + n_steps, n_paths)[0, 0]
                price, se = _price_se(m, r, T)
                assert abs(price - want) <= 4 * se, (n_steps, K, ot, price, want, se)


def test_randomised_cases_against_black_scholes_and_the_fp64_oracle(engine):
    """60 random (S, K, T, r, sigma, q, n_paths, n_steps) draws over wide ranges: ragged path / step counts exercise the
    tile planner and the trailing-step code; every price must sit within 4.5 standard errors of the closed form, and on the
    small cases the FP32 sums must match the FP64 oracle evaluation of the same Philox stream to 3e-4."""
    rng = np.random.default_rng(20260)
    worst = 0.0
    for case in range(60):
        S = float(rng.uniform(5.0, 500.0))
        K = S * float(rng.uniform(0.7, 1.4))
        T = float(rng.uniform(0.02, 3.0))
        r, q = float(rng.uniform(-0.01, 0.08)), float(rng.uniform(0.0, 0.05))
        sigma = float(rng.uniform(0.05, 0.8))
        n_steps = int(rng.integers(1, 400))
        n_paths = int(rng.integers(1_000, 300_000))
        ot = "call" if case % 2 == 0 else "put"
        res = ob.MonteCarloPricer(n_paths, n_steps, seed=case).price(S, K, T, r, sigma, ot, q=q, return_error=True)
        bs = orc.black_scholes(S, K, T, r, sigma, ot, q)
        assert res.n_paths == 2 * n_paths
        z = abs(res.price - bs) / max(res.std_error, 1e-12 * S)
        worst = max(worst, z)
        assert z < 4.5 or abs(res.price - bs) < 1e-6 * S, (case, S, K, T, r, sigma, q, n_paths, n_steps, res.price, bs, res.std_error)
        if n_paths * n_steps < 4_000_000:
            Z = po.normals(case, n_paths, n_steps)
            pay = orc.vanilla_payoffs(orc.gbm_terminal_from_normals(S, T, r, sigma, q, Z), K, ot)
            m = engine.simulate(_ffi.make_spec(_ffi.EUROPEAN, n_steps, is_put=ot == "put", antithetic=True),
                                _ffi.make_params(S, K, T, r, sigma, q).reshape(1, 1), case, n_paths)[0, 0]
            assert m["sum"] == pytest.approx(pay.sum(), rel=3e-4, abs=1e-6 * S * n_paths), case
    assert worst > 0.5  # the test is alive


def test_randomised_path_dependent_cases_against_the_fp64_oracle(engine):
    """40 random path-dependent cases (kind, flags, S, K, T, r, sigma, q, barrier, n_paths, n_steps): the fused FP32 kernels
    against the FP64 oracle evaluation of the same Philox stream.  Barrier-type payoffs may flip for a path within FP32
    rounding of the barrier, hence an absolute slack of a few payoffs."""
    rng = np.random.default_rng(777)
    for case in range(40):
        S = float(rng.uniform(20.0, 300.0))
        K = S * float(rng.uniform(0.8, 1.25))
        T = float(rng.uniform(0.05, 2.0))
        r, q = float(rng.uniform(0.0, 0.07)), float(rng.uniform(0.0, 0.04))
        sigma = float(rng.uniform(0.08, 0.9))  # beyond ~0.5 the arithmetic Asian leaves the small-move form
        n_steps = int(rng.integers(2, 120))
        n_paths = int(rng.integers(500, 30_000))
        is_put = bool(rng.integers(0, 2))
        ot = "put" if is_put else "call"
        Z = po.normals(1000 + case, n_paths, n_steps)
        paths = orc.exotic_paths_from_normals(S, T, r, sigma, q, Z)
        kind = case % 5
        B, slack = 0.0, 0.0
        if kind == 0:
            spec, want = _ffi.make_spec(_ffi.ASIAN_ARITH, n_steps, is_put=is_put), orc.asian_payoffs(paths, K, "arithmetic", ot)
        elif kind == 1:
            spec, want = _ffi.make_spec(_ffi.ASIAN_GEOM, n_steps, is_put=is_put), orc.asian_payoffs(paths, K, "geometric", ot)
        elif kind == 2:
            down, knock_in = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
            B = S * (float(rng.uniform(0.7, 0.97)) if down else float(rng.uniform(1.03, 1.4)))
            bt = ("down" if down else "up") + "-and-" + ("in" if knock_in else "out")
            spec, want = _ffi.make_spec(_ffi.BARRIER, n_steps, is_put=is_put, barrier_down=down, barrier_in=knock_in), orc.barrier_payoffs(paths, K, B, bt, ot)
            slack = 3.0 * S
        elif kind == 3:
            fixed = bool(rng.integers(0, 2))
            spec, want = (_ffi.make_spec(_ffi.LOOKBACK, n_steps, is_put=is_put, lookback_fixed=fixed),
                          orc.lookback_payoffs(paths, K, "fixed" if fixed else "floating", ot))
        else:
            nper = int(rng.integers(1, n_steps + 1))
            terms = dict(local_cap=float(rng.uniform(0.01, 0.1)), local_floor=-float(rng.uniform(0.0, 0.1)),
                         global_cap=float(rng.uniform(0.1, 0.6)), global_floor=float(rng.uniform(-0.1, 0.05)))
            want = orc.cliquet_payoffs(paths, S, n_periods=nper, **terms)
            m = engine.simulate_structured(_ffi.make_spec(_ffi.CLIQUET, n_steps), _ffi.Product(terms["local_cap"], terms["local_floor"], terms["global_cap"],
                                                                                             terms["global_floor"], nper, 0),
                                           _ffi.make_params(S, K, T, r, sigma, q).reshape(1, 1), 1000 + case, n_paths)[0, 0]
            assert m["sum"] == pytest.approx(want.sum(), rel=5e-4, abs=1e-4 * S * n_paths), (case, nper, terms)
            continue
        m = engine.simulate(spec, _ffi.make_params(S, K, T, r, sigma, q, B).reshape(1, 1), 1000 + case, n_paths)[0, 0]
        assert m["n"] == n_paths
        assert m["sum"] == pytest.approx(want.sum(), rel=5e-4, abs=slack + 1e-6 * S * n_paths), (case, kind, S, K, T, sigma, n_paths, n_steps)
