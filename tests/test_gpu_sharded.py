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


def test_local_devices_mode_on_two_gpus():
    """One process driving two GPUs from threads (distributed.local_devices): same global paths, host-side sum."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in this process")
    import optionslab_b200 as ob
    from optionslab_b200 import distributed

    distributed.shutdown()
    whole = _prices()
    with distributed.local_devices(2):
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
