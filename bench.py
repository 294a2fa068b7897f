#!/usr/bin/env python
"""Benchmark of the Monte Carlo hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

Workload (BASELINE.json configs[4], SURVEY.md §8d "C5"): a grid of 4096 European calls
(64 strikes x 64 maturities, S=100, r=5%, sigma=20%), each simulated INDEPENDENTLY with
1,000,000 antithetic path pairs x 252 log-Euler steps from its own Philox stream.
One "step" = one pass over the whole grid = 4096 x 1e6 x 252 = 1.032e12 GBM path-steps.
Metric: GBM path-steps / second (independent normal streams x steps; antithetic mirrors and
CRN scenarios are NOT counted), whole job over all N GPUs.  N > 1: paths are partitioned across
ranks (strong scaling: total work fixed) and the (sum, sum^2, n) moments are added up in the tail
of the simulation kernel over NVLink peer memory (b200mc_simulate_allreduce_device; NCCL all-reduce
only if the engines could not be connected).

The JSON line also carries: `e2e` (MonteCarloPricerUni.price_batch with host arrays in / prices out, same step count),
`roofline` (XU, issue and FMA-pipe fractions of the peaks measured live by b200mc_measure_peaks next to the paper peak, HBM
figure, DRAM traffic read from the committed ncu capture), `cpu_baseline` (the UNMODIFIED reference installed in
oracle/_ref - its Numba batch backend on all host threads - on a bounded sample; the NumPy restatement if that install
is absent), `configs` (BASELINE.json configs C1-C4, the Greeks launch and the Asian grid through the public API: kernel
and API time, path-steps/s, XU fraction, price, standard error - at N > 1 these are strong-scaling numbers),
`z_vs_black_scholes` (z-scores of the timed prices) and `clocks` (nvidia-smi during the run).

  python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA engine
  python bench.py --impl reference [...]                        # the reference's own CPU implementation on host cores
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC, UNIT = "gbm_path_steps_per_sec", "path-steps/s"
N_STRIKES, N_MATURITIES = 64, 64
N_OPT, N_PATHS, N_STEPS, SEED = N_STRIKES * N_MATURITIES, 1_000_000, 252, 42
WORKLOAD = (f"C5: {N_OPT}-option European call grid ({N_STRIKES} strikes 60..140 x {N_MATURITIES} maturities 1/12..2y, "
            f"S=100 r=0.05 sigma=0.2) x {N_PATHS} paths x {N_STEPS} steps, each option simulated independently "
            f"(antithetic, own Philox stream)")
# Instruction budget of the dominant kernel (european_kernel<1,true> inner loop, counted from the shipped SASS
# with cuobjdump - tools/sass_loop.py, profiles/r02_sass_loops.txt): 78 issued instructions per Philox call = 8 path-steps,
# of which 12 MUFU (LG2, SQRT and ONE SIN per Box-Muller pair: the terminal price adds a pair's two normals, and
# rad (cos + sin) = sqrt(2) rad sin(theta + pi/4)), 16 IMAD.WIDE (fmaheavy pipe), 18 LOP3 + 8 LEA.HI + 4 PRMT (ALU pipe), 16 FP32.
INSTR_PER_STEP, MUFU_PER_STEP, IMAD_PER_STEP, LOP_PER_STEP = 78 / 8, 1.5, 2.0, 30 / 8
IMAD_WIDE_PIPE_CYCLES, FP32_PER_STEP = 4.35, 16 / 8  # scratch/variants15.cu: 16 IMAD.WIDE (+ XORs, loop) per warp take 69.7 SMSP cycles
ASIAN_INSTR_PER_STEP = 117 / 8  # pathdep_kernel<ASIAN_ARITH,1> small-move loop (tools/sass_loop.py --all): 20 FFMA2 + 28 FP32, 16 MUFU per 8 steps


def grid_params():
    K = np.linspace(60.0, 140.0, N_STRIKES)
    T = np.linspace(1.0 / 12.0, 2.0, N_MATURITIES)
    KK, TT = np.meshgrid(K, T, indexing="ij")
    n = KK.size
    return dict(S=np.full(n, 100.0), K=KK.ravel().copy(), T=TT.ravel().copy(), r=np.full(n, 0.05),
                sigma=np.full(n, 0.2), q=np.zeros(n))


def black_scholes_call(S, K, T, r, sigma):
    """Closed-form sanity anchor for the timed run's prices (no oracle import in the GPU arm)."""
    from math import erf, exp, log, sqrt

    d1 = (log(S / K) + (r + 0.5 * sigma * sigma) * T) / (sigma * sqrt(T))
    d2 = d1 - sigma * sqrt(T)
    cdf = lambda x: 0.5 * (1.0 + erf(x / sqrt(2.0)))
    return S * cdf(d1) - K * exp(-r * T) * cdf(d2)


# ------------------------------------------------------------------------------------------------
# nvidia-smi clock sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index: int):
        self.device_index = device_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.device_index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, power, reasons = [], [], [], set()
        for t, line in self.lines:
            if not (t0 <= t <= t1 + 0.2):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])), mx.append(float(f[2])), power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in timed region"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power), "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle port) on host cores
# ------------------------------------------------------------------------------------------------
def _cpu_price_option(args):
    from oracle import reference_mc as orc

    i, n_paths = args
    return orc.cpu_grid_sample([i], n_paths, N_STEPS, SEED)[0]


def cpu_sample(option_indices, n_paths, processes):
    """Price a slice of the grid with the NumPy restatement of the reference; returns (seconds, prices)."""
    t0 = time.perf_counter()
    if processes <= 1:
        prices = [_cpu_price_option((i, n_paths)) for i in option_indices]
    else:
        import multiprocessing as mp

        with mp.get_context("fork").Pool(processes) as pool:
            prices = pool.map(_cpu_price_option, [(i, n_paths) for i in option_indices])
    return time.perf_counter() - t0, np.array(prices)


def _load_reference():
    """The unmodified reference installed by oracle/build_ref.py, or None (then the NumPy restatement stands in)."""
    try:
        from oracle import build_ref

        if not build_ref.available():
            return None
        import logging
        import warnings

        warnings.simplefilter("ignore")
        ref = build_ref.load()
        logging.getLogger("src").setLevel(logging.WARNING)
        return ref
    except Exception as exc:  # a broken install must not take the bench down: fall back to the port and say so
        sys.stderr.write(f"bench.py: oracle/_ref unusable ({exc!r}); timing the NumPy restatement instead\n")
        return None


def reference_grid_sample(ref, option_indices, n_paths, warm=False):
    """One pass of the REFERENCE over a slice of the C5 grid: MonteCarloPricerUni.price_batch on its Numba batch backend
    (monte_carlo_unified.py:145-204, prange over options - the reference's own multi-core path for this workload).
    -> (seconds, prices)."""
    g = grid_params()
    idx = np.asarray(option_indices)
    pricer = ref.MonteCarloPricerUni(num_simulations=n_paths, num_steps=N_STEPS, seed=SEED, use_numba=True)
    if warm:  # JIT (cache=True: compiled once per box)
        pricer.price_batch(g["S"][idx[:2]], g["K"][idx[:2]], g["T"][idx[:2]], g["r"][idx[:2]], g["sigma"][idx[:2]], "call", g["q"][idx[:2]])
    t0 = time.perf_counter()
    prices = pricer.price_batch(g["S"][idx], g["K"][idx], g["T"][idx], g["r"][idx], g["sigma"][idx], "call", g["q"][idx])
    return time.perf_counter() - t0, np.asarray(prices)


def reference_extras(ref):
    """The reference's other backends on the single-option configs ("CPU baseline beside it"),
    bounded sizes: NumPy (one thread - NumPy's generator is serial) and MCMethod.NUMBA (prange over paths) for the
    European call, AsianOption / BarrierOption NumPy, compute_greeks_unified.  Best of 2 after a warm-up."""
    P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)
    out = {}

    def best(fn, reps=2):
        fn()
        t = []
        for _ in range(reps):
            t0 = time.perf_counter()
            res = fn()
            t.append(time.perf_counter() - t0)
        return min(t), res

    try:
        pr = ref.MonteCarloPricer(100_000, N_STEPS, seed=SEED)
        t, res = best(lambda: pr.price(**P, option_type="call"))
        out["C1 MonteCarloPricer NUMPY 100k x 252"] = {"seconds": t, "path_steps_per_s": 100_000 * N_STEPS / t, "price": float(res), "threads": 1}
        prn = ref.MonteCarloPricer(100_000, N_STEPS, seed=SEED, method=ref.MCMethod.NUMBA)
        t, res = best(lambda: prn.price(**P, option_type="call"))
        import numba

        out["C1 MonteCarloPricer NUMBA 100k x 252"] = {"seconds": t, "path_steps_per_s": 100_000 * N_STEPS / t, "price": float(res),
                                                        "threads": int(numba.get_num_threads())}
        t, res = best(lambda: ref.AsianOption(**P, seed=SEED).price(n_paths=100_000, n_steps=N_STEPS), reps=1)
        out["C3 AsianOption 100k x 252"] = {"seconds": t, "path_steps_per_s": 100_000 * N_STEPS / t, "price": float(res), "threads": 1}
        t, res = best(lambda: ref.BarrierOption(**P, seed=SEED, barrier=120.0).price(n_paths=100_000, n_steps=365), reps=1)
        out["C4 BarrierOption 100k x 365"] = {"seconds": t, "path_steps_per_s": 100_000 * 365 / t, "price": float(res), "threads": 1}
        t, res = best(lambda: ref.compute_greeks_unified(pr, **P, option_type="call"), reps=1)
        out["C2 compute_greeks_unified 100k x 252 (14 re-simulations)"] = {"seconds": t, "path_steps_per_s": 14 * 100_000 * N_STEPS / t,
                                                                           "delta": float(res["delta"]), "threads": 1}
    except Exception as exc:
        out["error"] = repr(exc)
    return out


def _host_versions():
    out = {"cpu_count": os.cpu_count(), "numpy": np.__version__}
    try:
        import numba

        out["numba"], out["numba_threads"] = numba.__version__, int(numba.get_num_threads())
    except Exception:
        out["numba"] = None
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    n_paths = 100_000
    # stdout carries ONE JSON line: whatever the reference or Numba print goes to stderr meanwhile
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    ref = _load_reference()
    extras = None
    if ref is not None:
        import numba

        threads = int(numba.get_num_threads())
        per_step = max(2 * threads, 8)  # options per timed step: two per Numba thread (prange over options)
        idx = list(np.linspace(0, N_OPT - 1, per_step).astype(int))
        reference_grid_sample(ref, idx[:max(threads, 2)], 2_000, warm=True)
        for _ in range(max(args.warmup - 1, 0)):
            reference_grid_sample(ref, idx[:max(threads, 2)], 2_000)
        total_t = 0.0
        for _ in range(args.steps):
            t, _ = reference_grid_sample(ref, idx, n_paths)
            total_t += t
        kind, used = "reference", threads
        sample = (f"{per_step} grid options x {n_paths} paths x {N_STEPS} steps per step through the UNMODIFIED reference "
                  f"(oracle/_ref: MonteCarloPricerUni.price_batch, Numba batch backend, prange over options, {threads} threads; host has {cores} cores)")
        extras = reference_extras(ref)
    else:
        procs = max(1, min(cores, 64))
        per_step = procs * 2  # options per timed step: two per worker
        idx = list(np.linspace(0, N_OPT - 1, per_step).astype(int))
        for _ in range(args.warmup):
            cpu_sample(idx[:procs], 2_000, procs)
        total_t = 0.0
        for _ in range(args.steps):
            t, _ = cpu_sample(idx, n_paths, procs)
            total_t += t
        kind, used = "port", procs
        sample = (f"{per_step} grid options x {n_paths} paths x {N_STEPS} steps per step, NumPy restatement of "
                  f"simulate_gbm_numpy+payoff (oracle/reference_mc.py; oracle/_ref not installed), one process per core over options")
    work = per_step * n_paths * N_STEPS * args.steps
    value = work / total_t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": used, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "reference_other_backends": extras, "host": _host_versions()}
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
# MUFU per path-step of each kernel family (profiles/r02_sass_loops.txt): the XU pipe is the roof of all of them
CONFIG_MUFU = {"european": 1.5, "asian": 2.0, "barrier": 2.0}
NCU_CAPTURE = os.path.join(ROOT, "profiles", "r02_ncu_european.txt")


def traffic_from_capture(path=NCU_CAPTURE):
    """dram__bytes_read.sum + dram__bytes_write.sum of the headline kernel, per launch, from the COMMITTED ncu --set full
    capture of this bench command (tools/ncu_summary.py output).  Fails loudly when the capture is missing."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    if not os.path.exists(path):
        raise SystemExit(f"bench.py: {path} is missing - roofline.traffic is read from the committed ncu capture, not typed in "
                         "(regenerate with tools/gpu/r02_full_pass.sh)")
    total, seen, in_kernel = 0.0, 0, False
    with open(path) as f:
        for line in f:
            if line.startswith("=="):
                if seen:
                    break
                in_kernel = "european_kernel<1, 1" in line
            elif in_kernel and (line.startswith("dram__bytes_read.sum") or line.startswith("dram__bytes_write.sum")):
                parts = line.split()
                total += float(parts[1]) * scale[parts[2]]
                seen += 1
    if seen != 2:
        raise SystemExit(f"bench.py: {path} holds no european_kernel<1,1,...> capture with DRAM byte counters")
    return total, os.path.relpath(path, ROOT)


def measure_configs(eng, peaks, rank, world, barrier):
    """BASELINE.json configs[0..3] (SURVEY.md section 8d inputs) through the PUBLIC API, on every rank collectively: the
    pricer classes shard the global path range and the kernel tail adds up the ranks' moments.  Best of `reps` after one
    warm-up; kernel time from the engine's event ring (rank 0's share of the paths), API time by the host clock."""
    import optionslab_b200 as ob

    P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)
    rows = {}

    def record(name, family, path_steps, fn, reps, describe):
        fn()
        barrier()
        eng.set_kernel_timing(True)
        best, result = 1e30, None
        for _ in range(reps):
            t0 = time.perf_counter()
            result = fn()
            best = min(best, time.perf_counter() - t0)
        kt = eng.kernel_timing()
        eng.set_kernel_timing(False)
        barrier()
        if rank == 0:
            kernel_s = kt["min_ms"] * 1e-3
            row = {"kernel_ms": kt["min_ms"], "api_ms": best * 1e3, "path_steps": path_steps,
                   "path_steps_per_s_api": path_steps / best, "path_steps_per_s_kernel": path_steps / world / kernel_s * world,
                   "xu_frac_kernel": path_steps / world * CONFIG_MUFU[family] / kernel_s / peaks["mufu_per_s"],
                   "plan": eng.last_plan()}
            row.update(describe(result))
            rows[name] = row

    c1 = ob.MonteCarloPricer(100_000, 252, seed=42)
    record("C1 European call 100k x 252", "european", 100_000 * 252, lambda: c1.price(**P, option_type="call", return_error=True), 50,
           lambda r: {"price": r.price, "std_error": r.std_error})
    c1d = ob.MonteCarloPricer(100_000, seed=42)
    record("MonteCarloPricer default (100k x 1 step)", "european", 100_000, lambda: c1d.price(**P, option_type="call", return_error=True), 50,
           lambda r: {"price": r.price, "std_error": r.std_error})
    uni = ob.MonteCarloPricerUni(100_000, 100, seed=42)
    record("MonteCarloPricerUni default delta_gamma (100k x 100, h=1e-4, 3 scenarios)", "european", 100_000 * 100,
           lambda: uni.delta_gamma(**P, option_type="call", seed=7), 50, lambda r: {"delta": r[0], "gamma": r[1]})
    c2 = ob.MonteCarloPricer(1_000_000, 252, seed=42)
    for ot in ("call", "put"):
        record(f"C2 Greeks {ot} 1M x 252 (14 CRN scenarios, one launch)", "european", 1_000_000 * 252,
               lambda ot=ot: c2.greeks(**P, option_type=ot), 10, lambda g: {k: g[k] for k in ("price", "delta", "gamma", "vega")})
    asian = ob.AsianOption(**P, seed=42)
    record("C3 arithmetic Asian call 4M x 252", "asian", 4_000_000 * 252, lambda: asian.price(4_000_000, 252, return_error=True), 10,
           lambda r: {"price": r.price, "std_error": r.std_error})
    bar = ob.BarrierOption(**P, seed=42, barrier=120.0)
    record("C4 up-and-out barrier call 16M x 365", "barrier", 16_000_000 * 365, lambda: bar.price(16_000_000, 365, "up-and-out", return_error=True), 5,
           lambda r: {"price": r.price, "std_error": r.std_error})
    return rows


def run_engine_arm(args):
    import torch

    from optionslab_b200 import MonteCarloPricerUni, _ffi, distributed

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run --nproc-per-node N")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    import torch.distributed as dist

    # rank 0 prints ONE JSON line on stdout: while the job runs, file descriptor 1 points at stderr, so that library
    # chatter written straight to fd 1 (NCCL's version banner) cannot land in front of the JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    ctx = distributed.init(backend="nccl") if world > 1 else None
    eng = _ffi.get_engine(local)
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream(dev)

    g = grid_params()
    params_np = _ffi.make_params(g["S"], g["K"], g["T"], g["r"], g["sigma"], g["q"]).reshape(N_OPT, 1)
    params_dev = torch.from_numpy(params_np.view(np.float64).reshape(N_OPT, 8).copy()).to(dev)
    out_dev = torch.zeros((N_OPT, 3), dtype=torch.float64, device=dev)
    spec = _ffi.make_spec(_ffi.EUROPEAN, N_STEPS, antithetic=True)
    begin, count = distributed.partition_paths(N_PATHS, rank, world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    fused = ctx is not None and ctx.fused  # the ranks' moments are added up in the tail of the simulation kernel (NVLink peer reads)

    def device_step():
        flush.zero_()
        eng.simulate_device(spec, params_dev.data_ptr(), N_OPT, 1, SEED, count, out_dev.data_ptr(), stream.cuda_stream,
                            path_begin=begin, allreduce=fused)
        if world > 1 and not fused:
            dist.all_reduce(out_dev)

    peaks = eng.measure_peaks() if rank == 0 else None

    # ---- value: inputs resident in HBM, device-timed ------------------------------------------------
    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    eng.set_kernel_timing(True)
    launches0 = eng.kernel_launches()
    barrier()
    t_wall0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        device_step()
    ev1.record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    elapsed = torch.tensor([ev0.elapsed_time(ev1) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(elapsed, op=dist.ReduceOp.MAX)
    elapsed_s = float(elapsed.item())
    launches = eng.kernel_launches() - launches0 + args.steps  # + the L2-flush memset kernel per step
    ktime = eng.kernel_timing()
    eng.set_kernel_timing(False)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    moments = out_dev.cpu().numpy()
    work_per_step = float(N_OPT) * N_PATHS * N_STEPS
    value = work_per_step * args.steps / elapsed_s

    # ---- e2e: the public API, host buffers in, host prices out ---------------------------------------
    pricer = MonteCarloPricerUni(num_simulations=N_PATHS, num_steps=N_STEPS, seed=SEED)
    e2e_steps = args.steps  # the same K steps as `value`
    prices = pricer.price_batch(g["S"], g["K"], g["T"], g["r"], g["sigma"], "call", g["q"])  # warm-up (pinned buffers)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        prices = pricer.price_batch(g["S"], g["K"], g["T"], g["r"], g["sigma"], "call", g["q"])
    barrier()
    e2e_t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = work_per_step * e2e_steps / float(e2e_t.item())

    # ---- the same grid as arithmetic-average Asian calls (north_star: "batched European/Asian grid") -------------
    # Not the headline: 2 device-timed passes after 1 warm-up, reported under "asian_grid".
    asian = None
    if not args.no_asian_grid:
        aspec = _ffi.make_spec(_ffi.ASIAN_ARITH, N_STEPS)
        aout = torch.zeros((N_OPT, 3), dtype=torch.float64, device=dev)

        def asian_step():
            flush.zero_()
            eng.simulate_device(aspec, params_dev.data_ptr(), N_OPT, 1, SEED, count, aout.data_ptr(), stream.cuda_stream, path_begin=begin,
                                allreduce=fused)
            if world > 1 and not fused:
                dist.all_reduce(aout)

        asian_step()
        barrier()
        eng.set_kernel_timing(True)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(2):
            asian_step()
        a1.record(stream)
        barrier()
        at = torch.tensor([a0.elapsed_time(a1) * 1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(at, op=dist.ReduceOp.MAX)
        akt = eng.kernel_timing()
        eng.set_kernel_timing(False)
        if rank == 0:
            am = aout.cpu().numpy().view(_ffi.MOMENTS_DTYPE).reshape(N_OPT)
            arate = work_per_step / world / (akt["mean_ms"] * 1e-3)
            asian = {"workload": f"{N_OPT} arithmetic-average Asian calls (same strikes/maturities) x {N_PATHS} paths x {N_STEPS} steps, no mirroring",
                     "value": work_per_step * 2 / float(at.item()), "unit": UNIT, "steps": 2, "ms_per_step": 1e3 * float(at.item()) / 2,
                     "kernel": "pathdep_kernel<ASIAN_ARITH,NS=1> (small-move multiplicative update, packed FFMA2)", "kernel_ms": akt["mean_ms"],
                     "per_path_step": {"instructions": ASIAN_INSTR_PER_STEP, "mufu": 2.0},
                     "xu_frac": arate * 2.0 / peaks["mufu_per_s"], "issue_frac": arate * ASIAN_INSTR_PER_STEP / peaks["issue_per_s"],
                     "all_prices_below_european": None, "moments": am}

    # ---- BASELINE configs C1-C4 + the Greeks launch through the public API (collective under torchrun) -------------
    configs = None if args.no_configs else measure_configs(eng, peaks, rank, world, barrier)

    if rank == 0:
        # sanity: the timed run produced prices (checked against Black-Scholes within 4 standard errors)
        from optionslab_b200 import runtime

        m = moments.view(_ffi.MOMENTS_DTYPE).reshape(N_OPT)
        dev_prices = runtime.discounted_price(m, g["r"], g["T"])
        se = runtime.discounted_std_error(m, g["r"], g["T"])
        bs = np.array([black_scholes_call(g["S"][i], g["K"][i], g["T"][i], g["r"][i], g["sigma"][i]) for i in range(N_OPT)])
        z = np.abs(dev_prices - bs) / np.maximum(se, 1e-300)
        z = np.where(se > 0, z, 0.0)  # deep out-of-the-money short maturities: every payoff is 0 and Black-Scholes is < 1e-5
        # z is only a z-score where the CLT applies: far out of the money a handful of tiny payoffs gives a tiny se and a
        # large ratio although |price - BS| < 1e-5.  Report the Gaussian regime (>= 2000 expected in-the-money samples) too.
        from math import erf, log, sqrt
        p_itm = np.array([0.5 * (1.0 + erf((log(g["S"][i] / g["K"][i]) + (g["r"][i] - 0.5 * g["sigma"][i] ** 2) * g["T"][i])
                                           / (g["sigma"][i] * sqrt(g["T"][i])) / sqrt(2.0))) for i in range(N_OPT)])
        clt = p_itm * 2 * N_PATHS >= 2000
        i_max = int(np.argmax(z))
        ok = bool(np.all(np.abs(dev_prices - bs) <= 5.0 * se + 1e-5) and np.allclose(prices, dev_prices, rtol=1e-9))

        if asian is not None:  # an arithmetic-average call is worth less than the European call on the same strike/maturity
            ap = runtime.discounted_price(asian.pop("moments"), g["r"], g["T"])
            asian["all_prices_below_european"] = bool(np.all(ap <= dev_prices + 5.0 * se + 1e-5))

        if world > 1:
            traffic, traffic_src = None, "single-GPU figure only (the capture is of the N=1 launch)"
        elif args.traffic_capture == "none":
            traffic, traffic_src = None, "not read (--traffic-capture none)"
        else:
            traffic, traffic_src = traffic_from_capture(args.traffic_capture)
        kernel_s = ktime["mean_ms"] * 1e-3
        per_gpu_steps = work_per_step / world
        kernel_rate = per_gpu_steps / kernel_s  # path-steps/s of ONE GPU inside the kernel
        mufu_frac = kernel_rate * MUFU_PER_STEP / peaks["mufu_per_s"]
        issue_frac = kernel_rate * INSTR_PER_STEP / peaks["issue_per_s"]
        hbm_bytes = N_OPT * (64 + 24)  # algorithmic HBM traffic per launch: parameter block in, moments out
        try:
            hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
            hbm_src = "measured (MEASURED_PEAKS.json)"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        bound = "xu" if mufu_frac >= issue_frac else "issue"
        roofline = {
            "bound": bound,
            "kernel": "european_kernel<NS=1,ANTI>",
            "kernel_note": "every path-step's normal is drawn (one 32-bit Philox word per Box-Muller pair = two steps); the terminal "
                           "log-price adds a pair's two normals as sqrt(2) r sin(theta + pi/4) - the same value as r cos(theta) + "
                           "r sin(theta), 3 MUFU per pair instead of 4 (DESIGN.md, 'pair sum')",
            "kernel_ms": ktime["mean_ms"], "kernel_ms_min": ktime["min_ms"], "kernels_timed": ktime["count"],
            "achieved": kernel_rate * (MUFU_PER_STEP if bound == "xu" else INSTR_PER_STEP),
            "peak": peaks["mufu_per_s"] if bound == "xu" else peaks["issue_per_s"],
            "unit": "MUFU op/s" if bound == "xu" else "thread-instr/s",
            "frac": max(mufu_frac, issue_frac),
            "peak_source": "measured live by b200mc_measure_peaks on this GPU (pipe microbenchmarks)",
            # 16 MUFU lanes per SM per clock at the SM clock the run sustained (B200: 148 SMs)
            "peak_paper": eng.info()["sm_count"] * 16 * (clocks["sm_mhz"] or 1965.0) * 1e6,
            "frac_of_paper": kernel_rate * MUFU_PER_STEP / (eng.info()["sm_count"] * 16 * (clocks["sm_mhz"] or 1965.0) * 1e6),
            "per_path_step": {"instructions": INSTR_PER_STEP, "mufu": MUFU_PER_STEP, "imad_wide": IMAD_PER_STEP, "alu": LOP_PER_STEP},
            "xu_frac": mufu_frac, "issue_frac": issue_frac,
            "imad_frac": kernel_rate * IMAD_PER_STEP / peaks["imad_wide_per_s"],
            "alu_frac": kernel_rate * LOP_PER_STEP / peaks["lop3_per_s"],
            "vs_rng_only_probe": kernel_rate / peaks["normals_per_s"],
            # FMA-pipe view (profiles/r01_variants15_fma_pipe_model.txt): per SMSP a warp-wide IMAD.WIDE occupies the FMA pipe for
            # ~4.35 cycles, an FP32 instruction for 1.04, and they add up; fraction of the pipe's cycles (= the measured issue peak)
            "fma_pipe_frac": kernel_rate * (IMAD_PER_STEP * IMAD_WIDE_PIPE_CYCLES + FP32_PER_STEP * 1.04) / peaks["issue_per_s"],
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at this size (N=1), read from the committed ncu --set
            # full capture of this bench command (the tile partials stay in the 126 MB L2 and are folded there by the same
            # kernel); algorithmic: 262 KB of parameters in + 98 KB of moments out
            "traffic": traffic if world == 1 else None, "traffic_source": traffic_src,
            "hbm": {"achieved_gbs": hbm_bytes / kernel_s / 1e9, "peak_gbs": hbm_peak, "peak_source": hbm_src,
                    "frac": hbm_bytes / kernel_s / 1e9 / hbm_peak, "algorithmic_bytes_per_launch": hbm_bytes},
            "pipe_peaks": peaks,
        }
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            ref = _load_reference()
            if ref is not None:  # the unmodified reference, its own multi-core path, ~10-30 core-seconds
                import numba

                threads = int(numba.get_num_threads())
                idx = list(np.linspace(0, N_OPT - 1, max(4 * threads, 16)).astype(int))
                reference_grid_sample(ref, idx[:max(threads, 2)], 2_000, warm=True)
                t_cpu, cpu_prices = reference_grid_sample(ref, idx, 100_000)
                cpu = {"value": len(idx) * 100_000 * N_STEPS / t_cpu, "unit": UNIT, "cores": threads, "kind": "reference",
                       "sample": f"{len(idx)} grid options x 100000 paths x {N_STEPS} steps through the UNMODIFIED reference (oracle/_ref: "
                                 f"MonteCarloPricerUni.price_batch, Numba batch backend, prange over options, {threads} threads; host has {cores} cores)",
                       "seconds": t_cpu, "max_abs_price_diff_vs_gpu": float(np.max(np.abs(cpu_prices - dev_prices[idx])))}
            else:
                idx = list(np.linspace(0, N_OPT - 1, 40).astype(int))  # ~12 s of single-core NumPy work
                cpu_sample(idx[:1], 2_000, 1)
                t_cpu, cpu_prices = cpu_sample(idx, 100_000, 1)
                cpu = {"value": len(idx) * 100_000 * N_STEPS / t_cpu, "unit": UNIT, "cores": 1, "kind": "port",
                       "sample": f"{len(idx)} grid options x 100000 paths x {N_STEPS} steps, NumPy restatement of the reference "
                                 f"(oracle/reference_mc.py; oracle/_ref not installed), host has {cores} cores",
                       "seconds": t_cpu, "max_abs_price_diff_vs_gpu": float(np.max(np.abs(cpu_prices - dev_prices[idx])))}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * elapsed_s / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_options": N_OPT, "paths_per_option": N_PATHS, "steps_per_path": N_STEPS,
                       "parallelism": f"paths partitioned over {world} rank(s); the {N_OPT}x3 doubles of moments are summed over the ranks "
                                      + ("in the tail of the simulation kernel over NVLink peer memory (b200mc_simulate_allreduce_device)" if fused
                                         else "by one NCCL all-reduce per step" if world > 1 else "(single rank: no exchange)"),
                       "l2": "256 MiB memset between steps (inside the timed region); the kernel's HBM input is 262 KB"},
            "options_per_sec": N_OPT * args.steps / elapsed_s,
            # path-steps count independent normal streams; the antithetic mirrors ride on the same draws
            "trajectory_steps_per_sec": 2.0 * work_per_step * args.steps / elapsed_s,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(params_np.nbytes),
                    "d2h_bytes_per_step": int(N_OPT * 24), "steps": e2e_steps, "api": "MonteCarloPricerUni.price_batch (numpy in/out)"},
            "gpu_launches": int(launches),
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "asian_grid": asian, "configs": configs,
            "prices_ok": ok, "max_abs_z_vs_black_scholes": float(np.max(np.abs(z))),
            "z_vs_black_scholes": {"options_in_clt_regime": int(clt.sum()), "max_abs_clt": float(np.max(z[clt])),
                                   "mean_clt": float(np.mean(((dev_prices - bs) / np.maximum(se, 1e-300))[clt])),
                                   "std_clt": float(np.std(((dev_prices - bs) / np.maximum(se, 1e-300))[clt])),
                                   "argmax_all": {"K": float(g["K"][i_max]), "T": float(g["T"][i_max]), "price": float(dev_prices[i_max]),
                                                  "bs": float(bs[i_max]), "se": float(se[i_max])}},
        }
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if ctx is not None:
        distributed.shutdown()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    os.close(real_stdout)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["engine", "reference"], default="engine")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-asian-grid", action="store_true", help="skip the extra Asian-grid measurement (ncu runs)")
    ap.add_argument("--no-configs", action="store_true", help="skip the C1-C4 block (ncu runs)")
    ap.add_argument("--traffic-capture", default=NCU_CAPTURE, help="ncu summary roofline.traffic is read from; 'none' only for the run that produces it")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_engine_arm(args)


if __name__ == "__main__":
    sys.exit(main())
