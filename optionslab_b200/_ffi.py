"""ctypes binding of libb200mc.so (include/b200mc.h) — the only gateway to the GPU.

There is deliberately no fallback: if the shared library is missing, cannot be loaded, or no
sm_100 device is present, ``AccelerationError`` is raised.  Nothing here imports ``oracle``.
"""

from __future__ import annotations

import ctypes as C
import functools
import os
import threading
from typing import Optional

import numpy as np

from .exceptions import AccelerationError, MonteCarloError

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200MC_LIB") or os.path.join(_HERE, "libb200mc.so")  # B200MC_LIB: another build of the SAME library (A/B timing)
ABI_VERSION = 10
MAX_SCENARIOS = 16

EUROPEAN, ASIAN_ARITH, ASIAN_GEOM, BARRIER, LOOKBACK, CLIQUET, AUTOCALLABLE = range(7)
FLAG_EXACT_EX2 = 1
FLAG_NO_BULK_COPY = 2


class Spec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("is_put", C.c_int32), ("antithetic", C.c_int32),
                ("barrier_down", C.c_int32), ("barrier_in", C.c_int32), ("lookback_fixed", C.c_int32),
                ("n_steps", C.c_uint32), ("flags", C.c_uint32)]


class Product(C.Structure):
    """b200mc_product_t: CLIQUET (local_cap, local_floor, global_cap, global_floor, n_periods) /
    AUTOCALLABLE (autocall_barrier, coupon_barrier, coupon_rate, ki_barrier, observation_freq)."""
    _fields_ = [("a", C.c_double), ("b", C.c_double), ("c", C.c_double), ("d", C.c_double), ("period", C.c_uint32), ("reserved", C.c_uint32)]


class Info(C.Structure):
    _fields_ = [("device", C.c_int32), ("sm_count", C.c_int32), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
                ("sm_clock_khz", C.c_int32), ("mem_clock_khz", C.c_int32), ("total_mem_bytes", C.c_int64),
                ("l2_bytes", C.c_int32), ("reserved", C.c_int32), ("name", C.c_char * 64)]


class Peaks(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("ffma_per_s", "imad_wide_per_s", "lop3_per_s", "mufu_per_s", "mufu_ex2_per_s",
                                          "issue_per_s", "philox_per_s", "normals_per_s", "sm_clock_mhz_seen")] + \
               [("reserved", C.c_double * 3)]


PARAMS_DTYPE = np.dtype([("S", "f8"), ("K", "f8"), ("T", "f8"), ("r", "f8"), ("sigma", "f8"), ("q", "f8"),
                         ("barrier", "f8"), ("reserved", "f8")])
MOMENTS_DTYPE = np.dtype([("sum", "f8"), ("sum_sq", "f8"), ("n", "f8")])
CV_MOMENTS_DTYPE = np.dtype([("sum_payoff", "f8"), ("sum_payoff_sq", "f8"), ("sum_terminal", "f8"), ("sum_terminal_sq", "f8"),
                             ("sum_payoff_terminal", "f8"), ("n", "f8")])

RNG_STATS_DTYPE = np.dtype([("hist_z", "u8", (256,)), ("hist_joint", "u8", (4096,)), ("tails", "u8", (4,)), ("moments", "f8", (16,))])
HESTON_PARAMS_DTYPE = np.dtype([(n, "f8") for n in ("S", "K", "T", "r", "q", "kappa", "theta", "sigma_v", "rho", "v0")] + [("reserved", "f8", (2,))])
JUMP_PARAMS_DTYPE = np.dtype([("model", "i4"), ("reserved0", "i4"), ("lambda_j", "f8"), ("a", "f8"), ("b", "f8"), ("c", "f8"),
                              ("reserved", "f8", (3,))])
JUMP_MERTON, JUMP_KOU = 0, 1
MODEL_PUT, MODEL_SHARED_STREAM = 1, 2

# name -> (restype, argtypes); also the list the CPU tests check the .so exports against the header.
_P = C.c_void_p
SIGNATURES = {
    "b200mc_abi_version": (C.c_int, []),
    "b200mc_create": (C.c_int, [C.POINTER(_P), C.c_int]),
    "b200mc_destroy": (None, [_P]),
    "b200mc_last_error": (C.c_char_p, [_P]),
    "b200mc_device_info": (C.c_int, [_P, C.POINTER(Info)]),
    "b200mc_simulate": (C.c_int, [_P, C.POINTER(Spec), _P, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32,
                                  C.c_uint64, C.c_uint64, _P]),
    "b200mc_simulate_device": (C.c_int, [_P, C.POINTER(Spec), _P, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32,
                                         C.c_uint64, C.c_uint64, _P, _P]),
    "b200mc_simulate_allreduce": (C.c_int, [_P, C.POINTER(Spec), _P, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32,
                                            C.c_uint64, C.c_uint64, _P]),
    "b200mc_simulate_allreduce_device": (C.c_int, [_P, C.POINTER(Spec), _P, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32,
                                                   C.c_uint64, C.c_uint64, _P, _P]),
    "b200mc_comm_export": (C.c_int, [_P, _P]),
    "b200mc_comm_connect": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "b200mc_comm_connect_local": (C.c_int, [C.POINTER(_P), C.c_int]),
    "b200mc_comm_disconnect": (C.c_int, [_P]),
    "b200mc_comm_world": (C.c_int, [_P]),
    "b200mc_comm_set_timeout_ms": (C.c_int, [_P, C.c_uint32]),
    "b200mc_comm_set_collective": (C.c_int, [_P, C.c_int]),
    "b200mc_simulate_control_variate": (C.c_int, [_P, C.POINTER(Spec), _P, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32,
                                                  C.c_uint64, C.c_uint64, _P]),
    "b200mc_simulate_structured": (C.c_int, [_P, C.POINTER(Spec), C.POINTER(Product), _P, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32,
                                             C.c_uint64, C.c_uint64, _P]),
    "b200mc_structured_from_normals": (C.c_int, [_P, C.POINTER(Spec), C.POINTER(Product), _P, _P, C.c_uint64, _P, _P]),
    "b200mc_terminal_prices": (C.c_int, [_P, _P, C.c_uint32, C.c_int, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, _P]),
    "b200mc_terminal_prices_sobol": (C.c_int, [_P, _P, C.c_uint32, C.c_int, _P, _P, C.c_uint32, C.c_uint64, C.c_uint64, _P]),
    "b200mc_simulate_sobol": (C.c_int, [_P, C.POINTER(Spec), _P, C.c_uint32, C.c_uint32, _P, _P, C.c_uint32, C.c_uint64, C.c_uint64, _P]),
    "b200mc_sobol_points": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, _P]),
    "b200mc_sobol_normals": (C.c_int, [_P, _P, C.c_uint64, C.c_uint32, _P]),
    "b200mc_simulate_heston": (C.c_int, [_P, _P, C.c_uint32, C.c_int, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, _P]),
    "b200mc_simulate_jump_diffusion": (C.c_int, [_P, _P, _P, C.c_uint32, C.c_int, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64,
                                                 C.c_uint64, _P]),
    "b200mc_heston_from_normals": (C.c_int, [_P, _P, C.c_int, C.c_uint32, _P, C.c_uint64, _P, _P]),
    "b200mc_jump_diffusion_from_draws": (C.c_int, [_P, _P, C.c_double, C.c_int, C.c_uint32, _P, _P, C.c_uint64, _P, _P]),
    "b200mc_payoffs_from_normals": (C.c_int, [_P, C.POINTER(Spec), _P, C.c_int, _P, C.c_uint64, _P, _P]),
    "b200mc_payoffs_from_normals_device": (C.c_int, [_P, C.POINTER(Spec), _P, C.c_int, _P, C.c_uint64, _P, _P, _P]),
    "b200mc_generate_normals": (C.c_int, [_P, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, _P]),
    "b200mc_rng_statistics": (C.c_int, [_P, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, _P]),
    "b200mc_philox_raw": (C.c_int, [_P, _P, C.c_uint32, _P]),
    "b200mc_measure_peaks": (C.c_int, [_P, C.POINTER(Peaks)]),
    "b200mc_kernel_launches": (C.c_uint64, [_P]),
    "b200mc_set_plan": (C.c_int, [_P, C.c_int, C.c_uint32]),
    "b200mc_plan_tiles": (C.c_int, [C.c_int, C.POINTER(Spec), C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(C.c_uint32),
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "b200mc_last_plan": (C.c_int, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "b200mc_set_kernel_timing": (C.c_int, [_P, C.c_int]),
    "b200mc_kernel_timing": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int32)]),
}

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """dlopen libb200mc.so and type its entry points.  Raises AccelerationError if it is absent."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise AccelerationError(
                f"{LIB_PATH} is missing — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)", backend="cuda")
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as exc:
            raise AccelerationError(f"cannot load {LIB_PATH}: {exc}", backend="cuda") from exc
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if lib.b200mc_abi_version() != ABI_VERSION:
            raise AccelerationError(f"libb200mc ABI {lib.b200mc_abi_version()} != expected {ABI_VERSION}", backend="cuda")
        _lib = lib
        return lib


@functools.lru_cache(maxsize=512)
def make_spec(kind: int, n_steps: int, *, is_put=False, antithetic=False, barrier_down=False, barrier_in=False,
              lookback_fixed=False, exact_ex2=False, no_bulk_copy=False) -> Spec:
    """The launch description.  Cached: callers treat the returned struct as read-only."""
    return Spec(int(kind), int(bool(is_put)), int(bool(antithetic)), int(bool(barrier_down)), int(bool(barrier_in)),
                int(bool(lookback_fixed)), int(n_steps), (FLAG_EXACT_EX2 if exact_ex2 else 0) | (FLAG_NO_BULK_COPY if no_bulk_copy else 0))


def make_params(S, K, T, r, sigma, q=0.0, barrier=0.0) -> np.ndarray:
    """Broadcast the arguments into a PARAMS_DTYPE array (any common shape)."""
    arrs = np.broadcast_arrays(*(np.asarray(x, dtype=np.float64) for x in (S, K, T, r, sigma, q, barrier)))
    out = np.zeros(arrs[0].shape, dtype=PARAMS_DTYPE)
    for name, a in zip(("S", "K", "T", "r", "sigma", "q", "barrier"), arrs):
        out[name] = a
    return out


class Engine:
    """One device-side engine handle.  Thread-safe (the C side serialises calls per handle)."""

    def __init__(self, device: Optional[int] = None):
        self._lib = load_library()
        if device is None:
            device = int(os.environ.get("B200MC_DEVICE", os.environ.get("LOCAL_RANK", "0")))
        self.device = int(device)
        h = _P()
        rc = self._lib.b200mc_create(C.byref(h), self.device)
        if rc != 0:
            msg = self._lib.b200mc_last_error(None).decode()
            raise AccelerationError(f"b200mc_create({self.device}) failed: {msg}", backend="cuda")
        self._h = h
        self._tls = threading.local()  # per-thread staging buffers of simulate_scalars

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200mc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            msg = self._lib.b200mc_last_error(self._h).decode()
            if rc == -1:
                raise MonteCarloError(f"{what}: {msg}")
            raise AccelerationError(f"{what}: {msg}", backend="nvlink" if rc == -3 else "cuda")

    # -- queries ----------------------------------------------------------------------------
    def info(self) -> dict:
        out = Info()
        self._check(self._lib.b200mc_device_info(self._h, C.byref(out)), "b200mc_device_info")
        d = {f: getattr(out, f) for f, _ in Info._fields_ if f not in ("reserved", "name")}
        d["name"] = out.name.decode()
        return d

    def kernel_launches(self) -> int:
        return int(self._lib.b200mc_kernel_launches(self._h))

    def set_plan(self, split_shift: int = -1, paths_per_thread: int = 0):
        """Pin the tile shape (tuning / tests); the defaults restore the automatic plan."""
        self._check(self._lib.b200mc_set_plan(self._h, int(split_shift), int(paths_per_thread)), "b200mc_set_plan")

    def last_plan(self) -> dict:
        t, p, sft = C.c_uint32(), C.c_uint32(), C.c_uint32()
        self._check(self._lib.b200mc_last_plan(self._h, C.byref(t), C.byref(p), C.byref(sft)), "b200mc_last_plan")
        return {"tiles": int(t.value), "paths_per_thread": int(p.value), "split_shift": int(sft.value)}

    def set_kernel_timing(self, enabled: bool):
        self._check(self._lib.b200mc_set_kernel_timing(self._h, int(enabled)), "b200mc_set_kernel_timing")

    def kernel_timing(self) -> dict:
        """Mean / min duration (ms) and count of the simulation kernels timed since set_kernel_timing(True)."""
        mean, best, n = C.c_float(), C.c_float(), C.c_int32()
        self._check(self._lib.b200mc_kernel_timing(self._h, C.byref(mean), C.byref(best), C.byref(n)), "b200mc_kernel_timing")
        return {"mean_ms": float(mean.value), "min_ms": float(best.value), "count": int(n.value)}

    def measure_peaks(self) -> dict:
        out = Peaks()
        self._check(self._lib.b200mc_measure_peaks(self._h, C.byref(out)), "b200mc_measure_peaks")
        return {f: getattr(out, f) for f, _ in Peaks._fields_ if f != "reserved"}

    # -- hot path ---------------------------------------------------------------------------
    def simulate(self, spec: Spec, params: np.ndarray, seed: int, n_paths: int, *, stream_base: int = 0,
                 path_begin: int = 0, control_variate: bool = False, allreduce: bool = False) -> np.ndarray:
        """params: PARAMS_DTYPE array [n_opt, n_scen] -> MOMENTS_DTYPE (or CV_MOMENTS_DTYPE) array [n_opt, n_scen].
        ``allreduce``: the moments of ALL connected ranks' path ranges (fused into the kernel's tail; collective call)."""
        params = np.ascontiguousarray(params, dtype=PARAMS_DTYPE)
        if params.ndim != 2:
            raise MonteCarloError("params must have shape [n_opt, n_scen]")
        out = np.empty(params.shape, dtype=CV_MOMENTS_DTYPE if control_variate else MOMENTS_DTYPE)
        if allreduce and control_variate:
            raise MonteCarloError("the fused all-reduce carries (sum, sum_sq, n) records")
        fn = self._lib.b200mc_simulate_control_variate if control_variate else \
            self._lib.b200mc_simulate_allreduce if allreduce else self._lib.b200mc_simulate
        rc = fn(self._h, C.byref(spec), params.ctypes.data, params.shape[0], params.shape[1],
                int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_base) & 0xFFFFFFFF, int(path_begin), int(n_paths), out.ctypes.data)
        self._check(rc, "b200mc_simulate_control_variate" if control_variate else "b200mc_simulate")
        return out

    def simulate_scalars(self, spec: Spec, scenarios, seed: int, n_paths: int, *, barrier: float = 0.0, stream_base: int = 0,
                         path_begin: int = 0, allreduce: bool = False):
        """Latency path for ONE option: ``scenarios`` = up to 16 (S, K, T, r, sigma, q) tuples on common random numbers ->
        list of (sum, sum_sq, n) tuples.  No NumPy on the way: the parameter block is filled into a preallocated ctypes
        buffer and the C side launches one kernel whose arguments carry the coefficients and whose finishing CTA writes
        the moments into mapped pinned memory (b200mc_simulate, n_opt == 1)."""
        n = len(scenarios)
        if not 1 <= n <= MAX_SCENARIOS:
            raise MonteCarloError(f"between 1 and {MAX_SCENARIOS} scenarios per launch")
        tls = self._tls
        try:
            buf_in, buf_out = tls.buf_in, tls.buf_out
        except AttributeError:
            buf_in = tls.buf_in = (C.c_double * (8 * MAX_SCENARIOS))()
            buf_out = tls.buf_out = (C.c_double * (3 * MAX_SCENARIOS))()
        b = float(barrier)
        for k, sc in enumerate(scenarios):
            buf_in[8 * k:8 * k + 8] = (sc[0], sc[1], sc[2], sc[3], sc[4], sc[5], b, 0.0)
        fn = self._lib.b200mc_simulate_allreduce if allreduce else self._lib.b200mc_simulate
        rc = fn(self._h, C.byref(spec), buf_in, 1, n, int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_base) & 0xFFFFFFFF,
                int(path_begin), int(n_paths), buf_out)
        if rc != 0:
            self._check(rc, "b200mc_simulate_allreduce" if allreduce else "b200mc_simulate")
        out = buf_out[0:3 * n]
        return [(out[3 * k], out[3 * k + 1], out[3 * k + 2]) for k in range(n)]

    def simulate_device(self, spec: Spec, params_ptr: int, n_opt: int, n_scen: int, seed: int, n_paths: int, out_ptr: int,
                        cuda_stream: int, *, stream_base: int = 0, path_begin: int = 0, allreduce: bool = False):
        """Asynchronous variant on raw device pointers (params / moments already in HBM)."""
        fn = self._lib.b200mc_simulate_allreduce_device if allreduce else self._lib.b200mc_simulate_device
        rc = fn(self._h, C.byref(spec), params_ptr, n_opt, n_scen, int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_base) & 0xFFFFFFFF,
                int(path_begin), int(n_paths), out_ptr, cuda_stream)
        self._check(rc, "b200mc_simulate_allreduce_device" if allreduce else "b200mc_simulate_device")

    # -- communicator of the fused all-reduce ---------------------------------------------------
    def comm_export(self) -> bytes:
        """64-byte CUDA IPC handle of this engine's exchange block (step 1 of connecting ranks)."""
        buf = C.create_string_buffer(64)
        self._check(self._lib.b200mc_comm_export(self._h, buf), "b200mc_comm_export")
        return buf.raw

    def comm_connect(self, rank: int, world: int, handles: bytes):
        if len(handles) != 64 * world:
            raise MonteCarloError("handles must hold 64 bytes per rank")
        self._check(self._lib.b200mc_comm_connect(self._h, int(rank), int(world), handles), "b200mc_comm_connect")

    def comm_disconnect(self):
        self._check(self._lib.b200mc_comm_disconnect(self._h), "b200mc_comm_disconnect")

    def comm_set_collective(self, on: bool):
        """While on, every fused launch of this engine's host entry points adds up the connected ranks' moments in its tail."""
        self._check(self._lib.b200mc_comm_set_collective(self._h, int(bool(on))), "b200mc_comm_set_collective")

    def comm_set_timeout_ms(self, milliseconds: int):
        self._check(self._lib.b200mc_comm_set_timeout_ms(self._h, int(milliseconds)), "b200mc_comm_set_timeout_ms")

    def comm_world(self) -> int:
        return int(self._lib.b200mc_comm_world(self._h))

    # -- structured products (cliquet / autocallable) ------------------------------------------
    def simulate_structured(self, spec: Spec, product: Product, params: np.ndarray, seed: int, n_paths: int, *, stream_base: int = 0,
                            path_begin: int = 0) -> np.ndarray:
        """params [n_opt, n_scen] -> MOMENTS_DTYPE [n_opt, n_scen] (CLIQUET: undiscounted currency payoff;
        AUTOCALLABLE: discounted payoff per unit notional)."""
        params = np.ascontiguousarray(params, dtype=PARAMS_DTYPE)
        if params.ndim != 2:
            raise MonteCarloError("params must have shape [n_opt, n_scen]")
        out = np.empty(params.shape, dtype=MOMENTS_DTYPE)
        rc = self._lib.b200mc_simulate_structured(self._h, C.byref(spec), C.byref(product), params.ctypes.data, params.shape[0], params.shape[1],
                                                  int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_base) & 0xFFFFFFFF, int(path_begin), int(n_paths),
                                                  out.ctypes.data)
        self._check(rc, "b200mc_simulate_structured")
        return out

    def structured_from_normals(self, spec: Spec, product: Product, params: np.ndarray, Z: np.ndarray):
        """FP64 on the caller's draws Z [n_paths, n_steps] -> (payoffs [n_paths], moments)."""
        Z = np.ascontiguousarray(Z, dtype=np.float64)
        if Z.ndim != 2 or Z.shape[1] != spec.n_steps:
            raise MonteCarloError("Z must have shape [n_paths, n_steps]")
        p = np.ascontiguousarray(params, dtype=PARAMS_DTYPE).reshape(-1)
        if p.size != 1:
            raise MonteCarloError("parity mode prices one parameter set per call")
        pay = np.empty(Z.shape[0], dtype=np.float64)
        mom = np.empty(1, dtype=MOMENTS_DTYPE)
        rc = self._lib.b200mc_structured_from_normals(self._h, C.byref(spec), C.byref(product), p.ctypes.data, Z.ctypes.data, Z.shape[0],
                                                      pay.ctypes.data, mom.ctypes.data)
        self._check(rc, "b200mc_structured_from_normals")
        return pay, mom[0]

    # -- simulation layer: terminal price arrays ------------------------------------------------
    def terminal_prices(self, params: np.ndarray, n_steps: int, seed: int, n_paths: int, *, antithetic: bool = True, stream: int = 0,
                        path_begin: int = 0) -> np.ndarray:
        """-> float64 [n_paths] (or [2 n_paths]: +Z paths then their -Z mirrors, gbm_numpy.py:51)."""
        p = np.ascontiguousarray(params, dtype=PARAMS_DTYPE).reshape(-1)
        if p.size != 1:
            raise MonteCarloError("terminal_prices takes one parameter set")
        out = np.empty(int(n_paths) * (2 if antithetic else 1), dtype=np.float64)
        rc = self._lib.b200mc_terminal_prices(self._h, p.ctypes.data, int(n_steps), int(bool(antithetic)), int(seed) & 0xFFFFFFFFFFFFFFFF,
                                              int(stream) & 0xFFFFFFFF, int(path_begin), int(n_paths), out.ctypes.data)
        self._check(rc, "b200mc_terminal_prices")
        return out

    def terminal_prices_sobol(self, params: np.ndarray, n_steps: int, dirnums: np.ndarray, shift: np.ndarray, bits: int, n_points: int, *,
                              antithetic: bool = False, point_begin: int = 0) -> np.ndarray:
        p = np.ascontiguousarray(params, dtype=PARAMS_DTYPE).reshape(-1)
        if p.size != 1:
            raise MonteCarloError("terminal_prices_sobol takes one parameter set")
        dirnums = np.ascontiguousarray(dirnums, dtype=np.uint32)
        shift = np.ascontiguousarray(shift, dtype=np.uint32)
        if dirnums.shape != (n_steps, 32) or shift.shape != (n_steps,):
            raise MonteCarloError("Sobol table must have shape [n_steps, 32] and shift [n_steps]")
        out = np.empty(int(n_points) * (2 if antithetic else 1), dtype=np.float64)
        rc = self._lib.b200mc_terminal_prices_sobol(self._h, p.ctypes.data, int(n_steps), int(bool(antithetic)), dirnums.ctypes.data,
                                                    shift.ctypes.data, int(bits), int(point_begin), int(n_points), out.ctypes.data)
        self._check(rc, "b200mc_terminal_prices_sobol")
        return out

    # -- quasi-Monte Carlo ------------------------------------------------------------------
    def simulate_sobol(self, spec: Spec, params: np.ndarray, dirnums: np.ndarray, shift: np.ndarray, bits: int, n_points: int,
                       *, point_begin: int = 0) -> np.ndarray:
        """params [n_opt, n_scen] -> MOMENTS_DTYPE [n_opt, n_scen] over Sobol points [point_begin, point_begin+n_points)."""
        params = np.ascontiguousarray(params, dtype=PARAMS_DTYPE)
        if params.ndim != 2:
            raise MonteCarloError("params must have shape [n_opt, n_scen]")
        dirnums = np.ascontiguousarray(dirnums, dtype=np.uint32)
        shift = np.ascontiguousarray(shift, dtype=np.uint32)
        if dirnums.shape != (spec.n_steps, 32) or shift.shape != (spec.n_steps,):
            raise MonteCarloError("Sobol table must have shape [n_steps, 32] and shift [n_steps]")
        out = np.empty(params.shape, dtype=MOMENTS_DTYPE)
        rc = self._lib.b200mc_simulate_sobol(self._h, C.byref(spec), params.ctypes.data, params.shape[0], params.shape[1],
                                             dirnums.ctypes.data, shift.ctypes.data, int(bits), int(point_begin), int(n_points),
                                             out.ctypes.data)
        self._check(rc, "b200mc_simulate_sobol")
        return out

    def sobol_points(self, dirnums: np.ndarray, shift: np.ndarray, bits: int, n_points: int, *, point_begin: int = 0) -> np.ndarray:
        dirnums = np.ascontiguousarray(dirnums, dtype=np.uint32)
        shift = np.ascontiguousarray(shift, dtype=np.uint32)
        out = np.empty((n_points, dirnums.shape[0]), dtype=np.uint32)
        rc = self._lib.b200mc_sobol_points(self._h, dirnums.ctypes.data, shift.ctypes.data, dirnums.shape[0], int(bits),
                                           int(point_begin), int(n_points), out.ctypes.data)
        self._check(rc, "b200mc_sobol_points")
        return out

    def sobol_normals(self, x: np.ndarray, bits: int) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.uint32)
        out = np.empty(x.shape, dtype=np.float32)
        self._check(self._lib.b200mc_sobol_normals(self._h, x.ctypes.data, x.size, int(bits), out.ctypes.data), "b200mc_sobol_normals")
        return out

    # -- Heston / jump-diffusion models -----------------------------------------------------
    def simulate_heston(self, params: np.ndarray, is_put: bool, n_steps: int, seed: int, n_paths: int, *, stream_base: int = 0,
                        path_begin: int = 0, shared_stream: bool = False) -> np.ndarray:
        """params: HESTON_PARAMS_DTYPE [n_opt] -> MOMENTS_DTYPE [n_opt].  shared_stream: all entries price the same draws (CRN)."""
        params = np.ascontiguousarray(params, dtype=HESTON_PARAMS_DTYPE).reshape(-1)
        out = np.empty(params.shape, dtype=MOMENTS_DTYPE)
        rc = self._lib.b200mc_simulate_heston(self._h, params.ctypes.data, params.size,
                                              (MODEL_PUT if is_put else 0) | (MODEL_SHARED_STREAM if shared_stream else 0), int(n_steps),
                                              int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_base) & 0xFFFFFFFF, int(path_begin), int(n_paths),
                                              out.ctypes.data)
        self._check(rc, "b200mc_simulate_heston")
        return out

    def simulate_jump_diffusion(self, params: np.ndarray, jumps: np.ndarray, is_put: bool, n_steps: int, seed: int, n_paths: int, *,
                                stream_base: int = 0, path_begin: int = 0, shared_stream: bool = False) -> np.ndarray:
        """params: PARAMS_DTYPE [n_opt], jumps: JUMP_PARAMS_DTYPE [n_opt] -> MOMENTS_DTYPE [n_opt]."""
        params = np.ascontiguousarray(params, dtype=PARAMS_DTYPE).reshape(-1)
        jumps = np.ascontiguousarray(jumps, dtype=JUMP_PARAMS_DTYPE).reshape(-1)
        if jumps.size != params.size:
            raise MonteCarloError("one jump parameter set per option")
        out = np.empty(params.shape, dtype=MOMENTS_DTYPE)
        rc = self._lib.b200mc_simulate_jump_diffusion(self._h, params.ctypes.data, jumps.ctypes.data, params.size,
                                                      (MODEL_PUT if is_put else 0) | (MODEL_SHARED_STREAM if shared_stream else 0),
                                                      int(n_steps), int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream_base) & 0xFFFFFFFF,
                                                      int(path_begin), int(n_paths), out.ctypes.data)
        self._check(rc, "b200mc_simulate_jump_diffusion")
        return out

    def heston_from_normals(self, params: np.ndarray, is_put: bool, Z: np.ndarray):
        """Z: [n_steps, 2, n_paths] FP64 (step-major, as heston.py:228-229 draws) -> (payoffs [n_paths], moments)."""
        Z = np.ascontiguousarray(Z, dtype=np.float64)
        if Z.ndim != 3 or Z.shape[1] != 2:
            raise MonteCarloError("Z must have shape [n_steps, 2, n_paths]")
        p = np.ascontiguousarray(params, dtype=HESTON_PARAMS_DTYPE).reshape(-1)
        pay = np.empty(Z.shape[2], dtype=np.float64)
        mom = np.empty(1, dtype=MOMENTS_DTYPE)
        rc = self._lib.b200mc_heston_from_normals(self._h, p.ctypes.data, int(bool(is_put)), Z.shape[0], Z.ctypes.data, Z.shape[2],
                                                  pay.ctypes.data, mom.ctypes.data)
        self._check(rc, "b200mc_heston_from_normals")
        return pay, mom[0]

    def jump_diffusion_from_draws(self, params: np.ndarray, lambda_kappa: float, is_put: bool, dW: np.ndarray, J: Optional[np.ndarray]):
        """dW, J: [n_steps, n_paths] FP64 (step-major) -> (payoffs [n_paths], moments)."""
        dW = np.ascontiguousarray(dW, dtype=np.float64)
        if dW.ndim != 2:
            raise MonteCarloError("dW must have shape [n_steps, n_paths]")
        if J is not None:
            J = np.ascontiguousarray(J, dtype=np.float64)
            if J.shape != dW.shape:
                raise MonteCarloError("J must have the shape of dW")
        p = np.ascontiguousarray(params, dtype=PARAMS_DTYPE).reshape(-1)
        pay = np.empty(dW.shape[1], dtype=np.float64)
        mom = np.empty(1, dtype=MOMENTS_DTYPE)
        rc = self._lib.b200mc_jump_diffusion_from_draws(self._h, p.ctypes.data, float(lambda_kappa), int(bool(is_put)), dW.shape[0],
                                                        dW.ctypes.data, J.ctypes.data if J is not None else None, dW.shape[1],
                                                        pay.ctypes.data, mom.ctypes.data)
        self._check(rc, "b200mc_jump_diffusion_from_draws")
        return pay, mom[0]

    # -- FP64 parity mode -------------------------------------------------------------------
    def payoffs_from_normals(self, spec: Spec, params: np.ndarray, Z: np.ndarray, *, accumulate: bool = False,
                             want_payoffs: bool = True):
        Z = np.ascontiguousarray(Z, dtype=np.float64)
        if Z.ndim != 2 or Z.shape[1] != spec.n_steps:
            raise MonteCarloError("Z must have shape [n_paths, n_steps]")
        p = np.ascontiguousarray(params, dtype=PARAMS_DTYPE).reshape(-1)
        if p.size != 1:
            raise MonteCarloError("parity mode prices one parameter set per call")
        n_paths = Z.shape[0]
        pay = np.empty(n_paths * (2 if spec.antithetic else 1), dtype=np.float64) if want_payoffs else None
        mom = np.empty(1, dtype=MOMENTS_DTYPE)
        rc = self._lib.b200mc_payoffs_from_normals(self._h, C.byref(spec), p.ctypes.data, int(accumulate), Z.ctypes.data,
                                                   n_paths, pay.ctypes.data if want_payoffs else None, mom.ctypes.data)
        self._check(rc, "b200mc_payoffs_from_normals")
        return pay, mom[0]

    def payoffs_from_normals_device(self, spec: Spec, params: np.ndarray, Z_ptr: int, n_paths: int, payoffs_ptr: int,
                                    out_ptr: int, cuda_stream: int, *, accumulate: bool = False):
        p = np.ascontiguousarray(params, dtype=PARAMS_DTYPE).reshape(-1)
        rc = self._lib.b200mc_payoffs_from_normals_device(self._h, C.byref(spec), p.ctypes.data, int(accumulate), Z_ptr,
                                                          n_paths, payoffs_ptr, out_ptr, cuda_stream)
        self._check(rc, "b200mc_payoffs_from_normals_device")

    # -- stream inspection ------------------------------------------------------------------
    def generate_normals(self, seed: int, n_paths: int, n_steps: int, *, stream: int = 0, path_begin: int = 0) -> np.ndarray:
        out = np.empty((n_paths, n_steps), dtype=np.float32)
        rc = self._lib.b200mc_generate_normals(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF, stream, path_begin, n_paths, n_steps,
                                               out.ctypes.data)
        self._check(rc, "b200mc_generate_normals")
        return out

    def rng_statistics(self, seed: int, n_paths: int, n_steps: int, *, stream: int = 0, path_begin: int = 0) -> dict:
        """Device-side histograms / cross moments / tail counts of the normal stream (b200mc_rng_stats_t)."""
        out = np.zeros(1, dtype=RNG_STATS_DTYPE)
        rc = self._lib.b200mc_rng_statistics(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF, int(stream) & 0xFFFFFFFF, int(path_begin), int(n_paths),
                                             int(n_steps), out.ctypes.data)
        self._check(rc, "b200mc_rng_statistics")
        r = out[0]
        return {"hist_z": r["hist_z"].copy(), "hist_joint": r["hist_joint"].reshape(64, 64).copy(), "tails": r["tails"].copy(),
                "moments": r["moments"].copy()}

    def philox_raw(self, ctr_key: np.ndarray) -> np.ndarray:
        ck = np.ascontiguousarray(ctr_key, dtype=np.uint32).reshape(-1, 6)
        out = np.empty((ck.shape[0], 4), dtype=np.uint32)
        self._check(self._lib.b200mc_philox_raw(self._h, ck.ctypes.data, ck.shape[0], out.ctypes.data), "b200mc_philox_raw")
        return out


def plan_tiles(spec: Spec, n_opt: int, n_scen: int, n_paths: int, *, sm_count: int = 148, control_variate: bool = False) -> dict:
    """The planner's tile shape for a fused launch (pure host function of the library; needs no device)."""
    t, p, sft = C.c_uint32(), C.c_uint32(), C.c_uint32()
    rc = load_library().b200mc_plan_tiles(int(sm_count), C.byref(spec), int(n_opt), int(n_scen), int(n_paths), int(bool(control_variate)),
                                          C.byref(t), C.byref(p), C.byref(sft))
    if rc != 0:
        raise MonteCarloError("b200mc_plan_tiles: bad argument")
    return {"tiles": int(t.value), "paths_per_thread": int(p.value), "split_shift": int(sft.value)}


def connect_local(engines) -> None:
    """Connect engines of THIS process (one per device) for the fused all-reduce: rank = position in the list."""
    arr = (_P * len(engines))(*[e._h for e in engines])
    rc = load_library().b200mc_comm_connect_local(arr, len(engines))
    if rc != 0:
        msgs = "; ".join(e._lib.b200mc_last_error(e._h).decode() for e in engines)
        raise AccelerationError(f"b200mc_comm_connect_local failed ({rc}): {msgs}", backend="nvlink")


_engines = {}
_engines_lock = threading.Lock()


_default_device: Optional[int] = None


def default_device() -> int:
    """This process's device: B200MC_DEVICE, else LOCAL_RANK (torchrun), else 0 - read once."""
    global _default_device
    if _default_device is None:
        _default_device = int(os.environ.get("B200MC_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    return _default_device


def get_engine(device: Optional[int] = None) -> Engine:
    """Process-wide engine per device (created on first use)."""
    if device is None:
        device = default_device()
        eng = _engines.get(device)  # the latency path comes through here on every call: no lock when the engine exists
        if eng is not None and eng._h is not None:
            return eng
    with _engines_lock:
        eng = _engines.get(device)
        if eng is None or eng._h is None:
            eng = _engines[device] = Engine(device)
        return eng
