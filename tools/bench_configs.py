#!/usr/bin/env python
"""Per-config measurements for BASELINE.json configs C1-C4 (C5 is bench.py): kernel time (CUDA events
inside the engine), end-to-end API time, achieved fraction of the measured XU/issue peaks, and the
NumPy oracle timed on a bounded sample next to it.  Writes one JSON line per config."""

import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import optionslab_b200 as ob  # noqa: E402
from optionslab_b200 import _ffi  # noqa: E402
from oracle import reference_mc as orc  # noqa: E402  (CPU baseline / checker only)

P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)
# per path-step (instructions, MUFU) of each kernel family, from the shipped SASS (profiles/r01_sass_*.txt)
BUDGET = {"european": (78 / 8, 1.5), "asian": (117 / 8, 2.0), "asian_ex2": (111 / 8, 3.0), "barrier": (99 / 8, 2.0), "qmc": (330 / 16, 2.0),
          "heston": (112 / 4, 4.0), "jump": (80 / 8, 1.5), "structured": (130 / 8, 2.0)}  # tools/sass_loop.py (Heston: one Box-Muller pair + sqrt(v) per step)


# --profile: the run under `ncu --set full` (tools/gpu/r02_full_pass.sh).  One warm + one measured call per config and no CPU
# legs, so that a bounded capture count reaches every kernel family; the timings of such a run are not measurements (ncu
# replays each captured kernel ~40 times) and no file is written.
PROFILE = "--profile" in sys.argv[1:]


def timed(fn, reps=5):
    fn()
    best = 1e30
    for _ in range(1 if PROFILE else reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def main():
    eng = _ffi.get_engine(0)
    peaks = eng.measure_peaks()
    rows = []

    def record(name, family, path_steps, traj_steps, api_fn, cpu_fn, cpu_steps, note=""):
        eng.set_kernel_timing(True)
        api_s, result = timed(api_fn)
        kt = eng.kernel_timing()
        eng.set_kernel_timing(False)
        t0 = time.perf_counter()
        cpu_result = None if PROFILE else cpu_fn()
        cpu_s = max(time.perf_counter() - t0, 1e-9)
        instr, mufu = BUDGET[family]
        rate = traj_steps / (kt["min_ms"] * 1e-3)
        rows.append({"config": name, "kernel_ms": kt["min_ms"], "api_ms": api_s * 1e3,
                     "path_steps_per_s_kernel": path_steps / (kt["min_ms"] * 1e-3), "path_steps_per_s_api": path_steps / api_s,
                     "trajectory_steps_per_s_kernel": rate,
                     "xu_frac": rate * mufu / peaks["mufu_per_s"], "issue_frac": rate * instr / peaks["issue_per_s"],
                     "result": result, "cpu_oracle": {"seconds": cpu_s, "path_steps_per_s": cpu_steps / cpu_s, "result": cpu_result,
                                                      "cores": 1, "kind": "port"}, "note": note})
        print(json.dumps(rows[-1]), flush=True)

    # C1: European call, 100k paths x 252 steps
    pr = ob.MonteCarloPricer(100_000, 252, seed=42)
    record("C1 European call 100k x 252", "european", 100_000 * 252, 100_000 * 252,
           lambda: pr.price(**P, option_type="call"),
           lambda: orc.european_price(**P, option_type="call", num_simulations=100_000, num_steps=252, seed=42).price, 100_000 * 252)
    # C2: Greeks, 1M x 252, 14 CRN scenarios in one launch (reference: 14 separate simulations)
    pr2 = ob.MonteCarloPricer(1_000_000, 252, seed=42)
    cpu_pr = lambda S, K, T, r, s, q: orc.european_price(S, K, T, r, s, "call", q, num_simulations=100_000, num_steps=252, seed=42).price
    record("C2 Greeks (14 CRN scenarios) 1M x 252", "european", 1_000_000 * 252, 1_000_000 * 252,
           lambda: dict(pr2.greeks(**P, option_type="call")),
           lambda: dict(orc.greeks_bump_and_revalue(cpu_pr, **P)), 14 * 100_000 * 252,
           note="CPU oracle at 100k paths, 14 re-simulations; path-steps counted once per re-simulation on the CPU side")
    # C3: arithmetic Asian call, 4M x 252
    asian = ob.AsianOption(**P, seed=42)
    record("C3 Asian arithmetic call 4M x 252", "asian", 4_000_000 * 252, 4_000_000 * 252,
           lambda: float(asian.price(4_000_000, 252)),
           lambda: float(orc.exotic_price("asian", **P, seed=42, n_paths=100_000, n_steps=252)), 100_000 * 252,
           note="CPU oracle at 100k paths (the full path array of 4M x 253 doubles is 8 GB per temporary)")
    # the same C3 launch with B200MC_FLAG_EXACT_EX2 (additive log2 state + MUFU.EX2 per step): A/B of the small-move update
    a_params = _ffi.make_params(**P).reshape(1, 1)
    a_price = lambda spec: float(ob.runtime.discounted_price(eng.simulate(spec, a_params, 42, 4_000_000)[0, 0], P["r"], P["T"]))
    record("C3 Asian arithmetic call 4M x 252, MUFU.EX2 form (flag EXACT_EX2)", "asian_ex2", 4_000_000 * 252, 4_000_000 * 252,
           lambda: a_price(_ffi.make_spec(_ffi.ASIAN_ARITH, 252, exact_ex2=True)), lambda: None, 1,
           note="A/B row: the default C3 row above takes the multiplicative small-move update; no CPU leg")
    # C4: up-and-out barrier call, 16M x 365
    bar = ob.BarrierOption(**P, seed=42, barrier=120.0)
    record("C4 up-and-out barrier call 16M x 365", "barrier", 16_000_000 * 365, 16_000_000 * 365,
           lambda: float(bar.price(16_000_000, 365, "up-and-out")),
           lambda: float(orc.exotic_price("barrier", **P, seed=42, n_paths=100_000, n_steps=365, barrier=120.0)), 100_000 * 365,
           note="CPU oracle at 100k paths (16M x 366 doubles = 46.8 GB per array does not fit)")
    # structured products of the same file (exotic_options.py:404-552): autocallable (monthly observations) and cliquet (12 periods)
    auto = ob.AutocallableOption(**P, seed=42)
    record("Autocallable 4M x 252, 12 observation dates", "structured", 4_000_000 * 252, 4_000_000 * 252,
           lambda: float(auto.price(4_000_000, 252, 21)),
           lambda: float(orc.exotic_price("autocallable", **P, seed=42, n_paths=100_000, n_steps=252, observation_freq=21)), 100_000 * 252,
           note="CPU oracle at 100k paths")
    cliq = ob.CliquetOption(**P, seed=42)
    record("Cliquet 4M x 252, 12 reset periods", "structured", 4_000_000 * 252, 4_000_000 * 252,
           lambda: float(cliq.price(4_000_000, 252, 12)),
           lambda: float(orc.exotic_price("cliquet", **P, seed=42, n_paths=100_000, n_steps=252, n_periods=12)), 100_000 * 252,
           note="CPU oracle at 100k paths")
    # QMC: scrambled Sobol, 2^20 points x 252 dimensions (MCMethod.QMC; N samples, no mirroring)
    import warnings
    warnings.simplefilter("ignore")  # scipy: N not a power of two (CPU sample below)
    prq = ob.MonteCarloPricer(1 << 20, 252, seed=42, method=ob.MCMethod.QMC)
    record("QMC European call 2^20 Sobol points x 252", "qmc", (1 << 20) * 252, (1 << 20) * 252,
           lambda: prq.price(**P, option_type="call"),
           lambda: orc.european_price_qmc(**P, option_type="call", num_simulations=1 << 16, num_steps=252, seed=42).price, (1 << 16) * 252,
           note="CPU oracle (scipy Sobol + norm.ppf) at 2^16 points; api_ms includes building scipy's direction table on the host")
    # the same at 2^24 points: enough CTAs (4096) to fill the machine; the reference's (N, d) arrays would take 2 x 34 GB
    prq24 = ob.MonteCarloPricer(1 << 24, 252, seed=42, method=ob.MCMethod.QMC)
    record("QMC European call 2^24 Sobol points x 252", "qmc", (1 << 24) * 252, (1 << 24) * 252,
           lambda: prq24.price(**P, option_type="call"), lambda: None, 1, note="no CPU leg at this size")
    # Heston full-truncation Euler and Merton jump diffusion, 4M paths x 252 steps
    hes = ob.HestonPricer(kappa=2.0, theta=0.04, sigma_v=0.3, rho=-0.7, v0=0.04)
    record("Heston European call 4M x 252", "heston", 4_000_000 * 252, 4_000_000 * 252,
           lambda: hes.price_monte_carlo(100.0, 100.0, 1.0, 0.05, 0.01, "call", 4_000_000, 252, seed=42),
           lambda: orc.heston_price_mc(100.0, 100.0, 1.0, 0.05, 0.01, "call", kappa=2.0, theta=0.04, sigma_v=0.3, rho=-0.7, v0=0.04,
                                       n_paths=100_000, n_steps=252, seed=42), 100_000 * 252, note="CPU oracle at 100k paths")
    mer = ob.MertonJumpDiffusion(lambda_j=1.0, mu_j=-0.1, sigma_j=0.15)
    record("Merton jump-diffusion call 4M x 252", "jump", 4_000_000 * 252, 4_000_000 * 252,
           lambda: mer.price_monte_carlo(100.0, 100.0, 1.0, 0.05, 0.2, "call", 0.01, 4_000_000, 252, seed=42),
           lambda: orc.merton_price_mc(100.0, 100.0, 1.0, 0.05, 0.2, "call", 0.01, lambda_j=1.0, mu_j=-0.1, sigma_j=0.15,
                                       n_paths=20_000, n_steps=252, seed=42), 20_000 * 252,
           note="CPU oracle at 20k paths (vectorised replay of the reference's draws; the reference itself loops over paths in Python)")
    # FP64 parity mode (HBM-bound by design: 8 bytes of Z per path-step, read once)
    try:
        import torch

        dev = torch.device("cuda", 0)
        n_paths, n_steps = 4_000_000, 252
        Z = torch.randn((n_paths, n_steps), dtype=torch.float64, device=dev)
        pay = torch.empty(2 * n_paths, dtype=torch.float64, device=dev)
        mom = torch.empty(3, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        for kind, name, anti, plain in ((_ffi.EUROPEAN, "european (bulk-async staged)", True, False),
                                        (_ffi.EUROPEAN, "european (plain loads, flag NO_BULK_COPY)", True, True),
                                        (_ffi.ASIAN_ARITH, "asian", False, False), (_ffi.BARRIER, "barrier", False, False)):
            spec = _ffi.make_spec(kind, n_steps, antithetic=anti, no_bulk_copy=plain)
            eng.set_kernel_timing(True)
            for _ in range(2 if PROFILE else 4):
                eng.payoffs_from_normals_device(spec, _ffi.make_params(**P, barrier=120.0), Z.data_ptr(), n_paths, pay.data_ptr(), mom.data_ptr(), stream)
            torch.cuda.synchronize(dev)
            kt = eng.kernel_timing()
            eng.set_kernel_timing(False)
            gbs = n_paths * n_steps * 8 / (kt["min_ms"] * 1e-3) / 1e9
            rows.append({"config": f"FP64 parity mode {name} 4M x 252 (Z resident in HBM)", "kernel_ms": kt["min_ms"],
                         "path_steps_per_s_kernel": n_paths * n_steps / (kt["min_ms"] * 1e-3), "hbm_gbs": gbs, "hbm_peak_gbs": hbm_peak,
                         "hbm_frac": gbs / hbm_peak})
            print(json.dumps(rows[-1]), flush=True)
        del Z, pay
    except Exception as exc:  # measurement extra; never fatal
        print(json.dumps({"parity_mode_timing_failed": repr(exc)}), flush=True)
    out = {"peaks": peaks, "device": eng.info(), "rows": rows}
    if PROFILE:  # must not overwrite the file of the plain run
        return
    path = os.path.join(ROOT, "gpurun_out", "configs_r02.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
