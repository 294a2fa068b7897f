"""GPU: the sharded host path end to end with the real kernels — two ranks (gloo for the moment
all-reduce, both on cuda:0 because the test box has one GPU) must reproduce the one-process price.
The NCCL flavour of the same code is exercised by bench.py under torchrun (profiles/r01_bench_n2.json)."""

import json
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)


def _prices():
    import optionslab_b200 as ob

    out = {}
    res = ob.MonteCarloPricer(300_001, 24, seed=11).price(**P, option_type="call", return_error=True)
    out["euro"] = [res.price, res.std_error, res.n_paths]
    out["asian"] = float(ob.AsianOption(**P, seed=5).price(200_003, 20))
    out["barrier"] = float(ob.BarrierOption(**P, seed=5, barrier=115.0).price(200_003, 20, "up-and-in", "put"))
    uni = ob.MonteCarloPricerUni(100_001, 16, seed=3)
    out["batch"] = uni.price_batch([100.0, 90.0, 110.0], [100.0, 95.0, 105.0], [1.0, 0.5, 2.0], [0.05] * 3, [0.2, 0.3, 0.1], "put").tolist()
    out["greeks"] = dict(ob.MonteCarloPricer(100_001, 12, seed=2).greeks(**P, option_type="call"))
    out["autocall"] = float(ob.AutocallableOption(**P, seed=8).price(200_003, 24, 6))
    out["cliquet"] = ob.CliquetOption(**P, seed=8).price_scenarios([(100.0, 100.0, 1.0, 0.05, 0.2, 0.0), (101.0, 100.0, 1.0, 0.05, 0.21, 0.01)],
                                                                  n_paths=200_003, n_steps=24, n_periods=6)
    hes = ob.HestonPricer(kappa=2.0, theta=0.04, sigma_v=0.3, rho=-0.7, v0=0.04)
    out["heston"] = [hes.price_monte_carlo(100.0, 100.0, 1.0, 0.05, 0.01, "call", 200_003, 20, seed=4)] + \
        hes.price_scenarios([(100.0, 100.0, 1.0, 0.05, 0.04, 0.01), (101.0, 100.0, 1.0, 0.05, 0.0441, 0.01)], "put", 200_003, 20, seed=4)
    out["kou"] = ob.KouJumpDiffusion(2.0, 0.4, 10.0, 5.0).price_monte_carlo(100.0, 100.0, 1.0, 0.05, 0.2, "call", 0.01, 200_003, 20, seed=4)
    out["qmc"] = ob.MonteCarloPricer(1 << 16, 16, seed=42, method=ob.MCMethod.QMC).price(**P, option_type="call")
    return out


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0", B200MC_DEVICE="0",
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from optionslab_b200 import distributed

    distributed.init(backend="gloo")
    res = _prices()
    with open(os.path.join(out_dir, f"rank{rank}.json"), "w") as f:
        json.dump(res, f)
    distributed.shutdown()


def test_two_ranks_reproduce_single_process_prices(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from optionslab_b200 import distributed

    distributed.shutdown()
    whole = _prices()
    for rank in range(2):
        got = json.load(open(os.path.join(tmp_path, f"rank{rank}.json")))
        assert got["euro"][2] == whole["euro"][2] == 600_002
        # per-thread FP32 partial sums regroup when the path range is split: agreement is ~1e-7, far below 1 SE
        assert got["euro"][0] == pytest.approx(whole["euro"][0], rel=1e-6)
        assert got["euro"][1] == pytest.approx(whole["euro"][1], rel=1e-5)
        assert got["asian"] == pytest.approx(whole["asian"], rel=1e-6)
        assert got["barrier"] == pytest.approx(whole["barrier"], rel=1e-6)
        np.testing.assert_allclose(got["batch"], whole["batch"], rtol=1e-6)
        for k, v in whole["greeks"].items():
            assert got["greeks"][k] == pytest.approx(v, rel=2e-3, abs=2e-3), k
        assert got["autocall"] == pytest.approx(whole["autocall"], rel=1e-6)
        np.testing.assert_allclose(got["cliquet"], whole["cliquet"], rtol=1e-6)
        np.testing.assert_allclose(got["heston"], whole["heston"], rtol=1e-6)
        assert got["kou"] == pytest.approx(whole["kou"], rel=1e-6)
        assert got["qmc"] == pytest.approx(whole["qmc"], rel=1e-6)
    a = json.load(open(os.path.join(tmp_path, "rank0.json")))
    b = json.load(open(os.path.join(tmp_path, "rank1.json")))
    assert a == b  # every rank holds the identical all-reduced result


def _fused_pair(engines, spec, params, seed, n_paths, split_at):
    """Both connected engines price their path range with the all-reduce fused into the kernel tail (threads: the two
    launches wait for each other's records)."""
    from concurrent.futures import ThreadPoolExecutor

    ranges = [(0, split_at), (split_at, n_paths - split_at)]
    with ThreadPoolExecutor(max_workers=2) as pool:
        return list(pool.map(lambda i: engines[i].simulate(spec, params, seed, ranges[i][1], path_begin=ranges[i][0], allreduce=True), range(2)))


def test_fused_allreduce_between_two_engines_of_one_device():
    """The in-kernel exchange (b200mc_simulate_allreduce) with two engine handles on cuda:0: each rank's finishing CTA
    publishes its records, waits for the peer's flag and adds the peer's records in rank order.  Both ranks end with
    identical bits, equal to the sum of two plain launches over the same path ranges; empty shards, many options and the
    latency path (one option, mapped result) included.  The same kernel code runs between GPUs over NVLink."""
    from optionslab_b200 import _ffi

    a, b = _ffi.Engine(0), _ffi.Engine(0)
    plain = _ffi.get_engine(0)
    try:
        _ffi.connect_local([a, b])
        assert a.comm_world() == b.comm_world() == 2
        K = np.linspace(80.0, 120.0, 37)
        cases = [(_ffi.make_spec(_ffi.EUROPEAN, 32, antithetic=True), _ffi.make_params(100.0, K, 1.0, 0.05, 0.2).reshape(37, 1), 50_001, 20_000),
                 (_ffi.make_spec(_ffi.ASIAN_ARITH, 24), np.stack([_ffi.make_params(100.0 + d, 100.0, 1.0, 0.05, 0.2) for d in (1.0, 0.0, -1.0)]).reshape(1, 3), 70_003, 1),
                 (_ffi.make_spec(_ffi.BARRIER, 40, barrier_in=False), _ffi.make_params(100.0, 100.0, 1.0, 0.05, 0.2, barrier=120.0).reshape(1, 1), 300_000, 299_999),
                 (_ffi.make_spec(_ffi.EUROPEAN, 8, is_put=True, antithetic=True), _ffi.make_params(100.0, 105.0, 0.5, 0.03, 0.3).reshape(1, 1), 4_096, 0)]  # rank 0 empty
        for rep in range(3):  # consecutive epochs alternate the exchange slot
            for spec, params, n_paths, split_at in cases:
                got = _fused_pair([a, b], spec, params, 11 + rep, n_paths, split_at)
                assert got[0].tobytes() == got[1].tobytes()
                want = plain.simulate(spec, params, 11 + rep, n_paths - split_at, path_begin=split_at).copy()
                if split_at:
                    first = plain.simulate(spec, params, 11 + rep, split_at)
                    for f in ("sum", "sum_sq", "n"):
                        want[f] = first[f] + want[f]
                assert np.array_equal(got[0]["n"], want["n"])
                np.testing.assert_allclose(got[0]["sum"], want["sum"], rtol=1e-12)
                np.testing.assert_allclose(got[0]["sum_sq"], want["sum_sq"], rtol=1e-12)
        # unconnected again: allreduce degrades to the plain launch
        a.comm_disconnect(), b.comm_disconnect()
        spec, params, n_paths, _ = cases[0]
        assert a.simulate(spec, params, 5, n_paths, allreduce=True).tobytes() == plain.simulate(spec, params, 5, n_paths).tobytes()
    finally:
        a.close(), b.close()


def test_collective_mode_covers_every_fused_family():
    """b200mc_comm_set_collective: with the engine in collective mode the ORDINARY entry points of the other kernel families -
    control variate, structured products, Heston, jump diffusions, Sobol QMC - add up the connected ranks' records in their
    kernel tail too (what distributed.run_sharded switches on around each call).  Two engine handles on cuda:0: both ranks end
    with identical bits, equal to the field-wise sum of two plain launches over the same ranges; an empty share included."""
    from concurrent.futures import ThreadPoolExecutor

    from optionslab_b200 import _ffi, sobol

    a, b = _ffi.Engine(0), _ffi.Engine(0)
    plain = _ffi.get_engine(0)
    engines = [a, b]

    def both(call, ranges):
        def one(i):
            engines[i].comm_set_collective(True)
            try:
                return call(engines[i], *ranges[i])
            finally:
                engines[i].comm_set_collective(False)

        with ThreadPoolExecutor(max_workers=2) as pool:
            return list(pool.map(one, range(2)))

    hes = np.zeros(2, dtype=_ffi.HESTON_PARAMS_DTYPE)
    for i, K in enumerate((100.0, 105.0)):
        hes[i] = (100.0, K, 1.0, 0.05, 0.01, 2.0, 0.04, 0.3, -0.7, 0.04, (0.0, 0.0))
    gbm = _ffi.make_params(100.0, np.array([95.0, 100.0]), 1.0, 0.05, 0.2, 0.01)
    jumps = np.zeros(2, dtype=_ffi.JUMP_PARAMS_DTYPE)
    jumps["model"], jumps["lambda_j"], jumps["a"], jumps["b"], jumps["c"] = _ffi.JUMP_KOU, 2.0, 0.4, 10.0, 5.0
    table, shift, bits = sobol.sobol_table(16, 5)
    euro = _ffi.make_spec(_ffi.EUROPEAN, 16, antithetic=True)
    qmc = _ffi.make_spec(_ffi.EUROPEAN, 16)
    product = _ffi.Product(0.05, -0.05, 0.30, 0.0, 4, 0)
    grid2 = gbm.reshape(2, 1)
    cases = {
        "control variate": (lambda e, lo, n: e.simulate(euro, grid2, 9, n, path_begin=lo, control_variate=True), [(0, 30_000), (30_000, 20_001)]),
        "cliquet": (lambda e, lo, n: e.simulate_structured(_ffi.make_spec(_ffi.CLIQUET, 16), product, grid2, 9, n, path_begin=lo), [(0, 25_000), (25_000, 25_000)]),
        "heston": (lambda e, lo, n: e.simulate_heston(hes, False, 16, 9, n, path_begin=lo), [(0, 1), (1, 49_999)]),
        "kou": (lambda e, lo, n: e.simulate_jump_diffusion(gbm, jumps, True, 16, 9, n, path_begin=lo), [(0, 40_000), (40_000, 0)]),  # rank 1 empty
        "sobol": (lambda e, lo, n: e.simulate_sobol(qmc, grid2, table, shift, bits, n, point_begin=lo), [(0, 8192), (8192, 8192 + 77)]),
    }
    try:
        _ffi.connect_local(engines)
        for name, (call, ranges) in cases.items():
            for rep in range(2):
                got = both(call, ranges)
                assert got[0].tobytes() == got[1].tobytes(), name
                parts = [call(plain, lo, n) for lo, n in ranges if n > 0]
                for field in got[0].dtype.names:
                    np.testing.assert_allclose(got[0][field], sum(p[field] for p in parts), rtol=1e-12, err_msg=f"{name}.{field}")
        # outside collective mode the same entry points stay local even on connected engines
        assert a.simulate_heston(hes, False, 16, 9, 1000).tobytes() == plain.simulate_heston(hes, False, 16, 9, 1000).tobytes()
    finally:
        a.close(), b.close()


def test_a_missing_peer_is_an_error_not_a_hang():
    """Failure detection of the in-kernel exchange: if a connected rank never launches, the waiting kernel gives up after the
    configured bound and the call fails with AccelerationError(backend="nvlink") (B200MC_ERR_COMM); the engine refuses further
    exchanges until the communicator is rebuilt, after which everything works again."""
    import time

    import optionslab_b200 as ob
    from optionslab_b200 import _ffi

    a, b = _ffi.Engine(0), _ffi.Engine(0)
    try:
        _ffi.connect_local([a, b])
        a.comm_set_timeout_ms(150)
        spec = _ffi.make_spec(_ffi.EUROPEAN, 16, antithetic=True)
        params = _ffi.make_params(**P).reshape(1, 1)
        t0 = time.perf_counter()
        with pytest.raises(ob.AccelerationError, match="timed out"):
            a.simulate(spec, params, 3, 10_000, allreduce=True)  # rank 1 (engine b) never shows up
        assert 0.1 < time.perf_counter() - t0 < 5.0
        with pytest.raises(ob.AccelerationError, match="reconnect"):
            a.simulate(spec, params, 3, 10_000, allreduce=True)
        _ffi.connect_local([a, b])  # rebuild: flags, epochs and the time-out word start over
        got = _fused_pair([a, b], spec, params, 3, 20_000, 10_000)
        assert got[0].tobytes() == got[1].tobytes() and got[0]["n"][0, 0] == 40_000
    finally:
        a.close(), b.close()


def test_local_devices_mode_on_two_gpus():
    """One process driving two GPUs from threads (distributed.local_devices): same global paths; the GBM launches add
    up their records in the kernel tail over peer memory, the other model families on the host."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in this process")
    import optionslab_b200 as ob
    from optionslab_b200 import distributed

    distributed.shutdown()
    whole = _prices()
    with distributed.local_devices(2):
        assert distributed.fused_exchange()
        split = _prices()
    assert split["euro"][2] == whole["euro"][2]
    assert split["euro"][0] == pytest.approx(whole["euro"][0], rel=1e-6)
    for k in ("asian", "barrier", "autocall", "kou", "qmc"):
        assert split[k] == pytest.approx(whole[k], rel=1e-6), k
    np.testing.assert_allclose(split["batch"], whole["batch"], rtol=1e-6)
    np.testing.assert_allclose(split["cliquet"], whole["cliquet"], rtol=1e-6)
    np.testing.assert_allclose(split["heston"], whole["heston"], rtol=1e-6)
    # and it scales: 2 devices on a grid large enough to amortise the threads
    import time

    g = dict(S=np.full(512, 100.0), K=np.linspace(60, 140, 512), T=np.full(512, 1.0), r=np.full(512, 0.05), sigma=np.full(512, 0.2))
    uni = ob.MonteCarloPricerUni(1_000_000, 252, seed=1)
    uni.price_batch(g["S"], g["K"], g["T"], g["r"], g["sigma"], "call")
    t0 = time.perf_counter()
    one = uni.price_batch(g["S"], g["K"], g["T"], g["r"], g["sigma"], "call")
    t1 = time.perf_counter()
    with distributed.local_devices(2):
        uni.price_batch(g["S"], g["K"], g["T"], g["r"], g["sigma"], "call")
        t2 = time.perf_counter()
        two = uni.price_batch(g["S"], g["K"], g["T"], g["r"], g["sigma"], "call")
        t3 = time.perf_counter()
    np.testing.assert_allclose(two, one, rtol=1e-6)
    assert (t3 - t2) < 0.65 * (t1 - t0), (t1 - t0, t3 - t2)
