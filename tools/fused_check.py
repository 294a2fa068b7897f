#!/usr/bin/env python
"""Run under torchrun (one process per GPU): the pricer classes with the all-reduce fused into the simulation kernel
(distributed.init -> b200mc_comm_connect over CUDA IPC) against the same calls through NCCL all_reduce
(B200MC_FUSED_ALLREDUCE=0 semantics, toggled in-process), plus per-call latency of the single-option configs C3 / C4.
Rank 0 prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import optionslab_b200 as ob  # noqa: E402
from optionslab_b200 import distributed  # noqa: E402

P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)


def prices():
    out = {}
    res = ob.MonteCarloPricer(300_001, 24, seed=11).price(**P, option_type="call", return_error=True)
    out["euro"] = [res.price, res.std_error, res.n_paths]
    out["asian"] = float(ob.AsianOption(**P, seed=5).price(200_003, 20))
    out["barrier"] = float(ob.BarrierOption(**P, seed=5, barrier=115.0).price(200_003, 20, "up-and-in", "put"))
    uni = ob.MonteCarloPricerUni(100_001, 16, seed=3)
    out["batch"] = uni.price_batch([100.0, 90.0, 110.0], [100.0, 95.0, 105.0], [1.0, 0.5, 2.0], [0.05] * 3, [0.2, 0.3, 0.1], "put").tolist()
    out["delta_gamma"] = list(uni.delta_gamma(**P, option_type="call", seed=4))
    out["greeks"] = dict(ob.MonteCarloPricer(100_001, 12, seed=2).greeks(**P, option_type="call"))
    out["tiny"] = ob.MonteCarloPricer(3, 4, seed=1).price(**P, option_type="put")  # fewer paths than ranks: empty shards
    # the other fused families exchange through the engine's collective mode
    out["control_variate"] = ob.MonteCarloPricer(200_003, 16, seed=6).price_with_control_variate(**P, option_type="call")
    out["autocall"] = float(ob.AutocallableOption(**P, seed=8).price(200_003, 24, 6))
    out["cliquet"] = float(ob.CliquetOption(**P, seed=8).price(200_003, 24, 6))
    hes = ob.HestonPricer(kappa=2.0, theta=0.04, sigma_v=0.3, rho=-0.7, v0=0.04)
    out["heston"] = hes.price_monte_carlo(100.0, 100.0, 1.0, 0.05, 0.01, "call", 200_003, 20, seed=4)
    out["kou"] = ob.KouJumpDiffusion(2.0, 0.4, 10.0, 5.0).price_monte_carlo(100.0, 100.0, 1.0, 0.05, 0.2, "call", 0.01, 200_003, 20, seed=4)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out["qmc"] = ob.MonteCarloPricer((1 << 16) + 5, 16, seed=42, method=ob.MCMethod.QMC).price(**P, option_type="call")
    return out


def timed(fn, reps):
    import torch.distributed as dist

    fn()
    dist.barrier()
    t = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        t.append(time.perf_counter() - t0)
    t.sort()
    return t[0] * 1e6, t[len(t) // 2] * 1e6


def main():
    import torch
    import torch.distributed as dist

    sys.stdout.flush()
    real_stdout = os.dup(1)  # NCCL prints its version banner on fd 1: park stdout on stderr until the JSON line
    os.dup2(2, 1)
    ctx = distributed.init(backend="nccl")
    fused_ok = ctx.fused
    a = prices()
    lat = {}
    asian, barrier = ob.AsianOption(**P, seed=42), ob.BarrierOption(**P, seed=42, barrier=120.0)
    c1 = ob.MonteCarloPricer(100_000, 252, seed=42)
    for name, fn, reps in (("C1 100k x 252", lambda: c1.price(**P, option_type="call"), 200),
                           ("C3 Asian 4M x 252", lambda: asian.price(4_000_000, 252), 50),
                           ("C4 barrier 16M x 365", lambda: barrier.price(16_000_000, 365, "up-and-out"), 20)):
        lat[name] = {"fused_us_min_median": timed(fn, reps)}
    ctx.fused = False  # same process group, moments through NCCL all_reduce instead
    b = prices()
    for name, fn, reps in (("C1 100k x 252", lambda: c1.price(**P, option_type="call"), 200),
                           ("C3 Asian 4M x 252", lambda: asian.price(4_000_000, 252), 50),
                           ("C4 barrier 16M x 365", lambda: barrier.price(16_000_000, 365, "up-and-out"), 20)):
        lat[name]["nccl_us_min_median"] = timed(fn, reps)
    ctx.fused = fused_ok
    gathered = [None] * ctx.world_size
    dist.all_gather_object(gathered, a)
    if ctx.rank == 0:
        same_on_all_ranks = all(g == gathered[0] for g in gathered)

        def close(x, y):
            return bool(np.allclose(np.asarray(list(x.values()) if isinstance(x, dict) else x, dtype=float),
                                    np.asarray(list(y.values()) if isinstance(y, dict) else y, dtype=float), rtol=1e-9, atol=1e-12))

        # delta_gamma's second entry is the h = 1e-4 second difference: last-bit differences of the two summation orders
        # (rank order in the kernel, NCCL's tree) are amplified by 1/h^2 = 1e8 there - compare the delta only
        a["delta_gamma"], b["delta_gamma"] = a["delta_gamma"][:1], b["delta_gamma"][:1]
        agree = {k: close(a[k], b[k]) for k in a}
        os.write(real_stdout, (json.dumps({"world": ctx.world_size, "fused_connected": fused_ok, "identical_on_all_ranks": same_on_all_ranks,
                          "fused_equals_nccl": agree, "latency": lat, "prices": a}) + "\n").encode())
    distributed.shutdown()


if __name__ == "__main__":
    main()
