"""Multi-GPU sharding of the Monte Carlo path: one process per GPU, paths partitioned by global
index, ONE all-reduce of the per-(option, scenario) moments.

Two ways to use several GPUs: one PROCESS per GPU (``init()`` under torchrun; what bench.py measures), or one process
driving several devices from threads (``with local_devices(n):`` — each device has its own engine handle and stream,
ctypes releases the GIL, the partial moments are summed on the host).  Both split the same global path range.

Every (option, path) is independent: rank g of G simulates global paths
[g*N/G, (g+1)*N/G) of every option — the Philox counter carries the global path index, so the
draws are disjoint by construction and independent of G — and the only exchanged data are the
FP64 (sum, sum^2, n) triples.

The exchange itself: on NVLink-connected GPUs (the NCCL backend under torchrun, or ``local_devices``) every fused launch -
European / Asian / barrier / lookback (``b200mc_simulate_allreduce``), control variate, structured products, Heston, jump
diffusions, Sobol QMC (their ordinary entry points with the engine in collective mode) - adds up the ranks' records in the
TAIL OF THE SIMULATION KERNEL over peer memory: the finishing CTA of every rank reads its peers' exchange blocks in rank
order, so there is no collective launch and no host hop between the kernel and the result (include/b200mc.h,
"multi-GPU").  ``torch.distributed`` then only carries the 64-byte IPC handles at ``init()`` time.  Gloo (the CPU tests),
unconnected devices and ``B200MC_FUSED_ALLREDUCE=0`` sum the moment records with one ``all_reduce`` / on the host.
"""

from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

__all__ = ["ShardContext", "init", "shutdown", "current", "partition_paths", "allreduce_moments", "run_sharded", "local_devices", "fused_exchange"]


@dataclass
class ShardContext:
    rank: int
    world_size: int
    backend: str
    device: Optional[int]  # CUDA device index for NCCL, None for gloo/CPU
    owns_group: bool = False
    fused: bool = False    # the engines of all ranks are connected for the in-kernel all-reduce


_ctx: Optional[ShardContext] = None


def partition_paths(n_paths: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced split of [0, n_paths): returns (first global path, count) for ``rank``."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, extra = divmod(int(n_paths), world_size)
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


def init(backend: Optional[str] = None) -> ShardContext:
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun)."""
    global _ctx
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    device = None
    if backend == "nccl":
        device = local
        torch.cuda.set_device(device)
    owns = False
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {"device_id": torch.device("cuda", device)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
        owns = True
    _ctx = ShardContext(rank=rank, world_size=world, backend=backend, device=device, owns_group=owns)
    if world > 1 and backend == "nccl" and os.environ.get("B200MC_FUSED_ALLREDUCE", "1") != "0":
        _connect_fused(_ctx)
    return _ctx


def _connect_fused(ctx: ShardContext) -> None:
    """Exchange the engines' IPC handles and map every peer's exchange block (include/b200mc.h).  Ranks that cannot
    (more than 8 ranks, no peer access) all fall back to the all_reduce route together."""
    import torch.distributed as dist

    from . import _ffi
    from .exceptions import MonteCarloError

    eng = _ffi.get_engine(ctx.device)
    ok = True
    try:
        mine = eng.comm_export() if ctx.world_size <= 8 else b""
    except MonteCarloError:
        mine = b""
    handles = [None] * ctx.world_size
    dist.all_gather_object(handles, mine)
    if all(len(h) == 64 for h in handles):
        try:
            eng.comm_connect(ctx.rank, ctx.world_size, b"".join(handles))
        except MonteCarloError:
            ok = False
    else:
        ok = False
    flags = [None] * ctx.world_size
    dist.all_gather_object(flags, ok)  # also the barrier between connecting and the first exchange
    ctx.fused = all(flags)
    if not ctx.fused and ok:
        eng.comm_disconnect()


def shutdown():
    global _ctx
    if _ctx is not None and _ctx.fused:
        from . import _ffi

        try:
            _ffi.get_engine(_ctx.device).comm_disconnect()
        except Exception:
            pass
    if _ctx is not None and _ctx.owns_group:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()
    _ctx = None


def current() -> Optional[ShardContext]:
    return _ctx


def allreduce_moments(moments: np.ndarray, ctx: Optional[ShardContext] = None) -> np.ndarray:
    """Sum a MOMENTS_DTYPE array (sum, sum_sq, n) over all ranks; identity when not sharded."""
    ctx = ctx or _ctx
    if ctx is None or ctx.world_size == 1:
        return moments
    import torch
    import torch.distributed as dist

    flat = np.ascontiguousarray(moments).view(np.float64).reshape(-1).copy()
    t = torch.from_numpy(flat)
    if ctx.backend == "nccl":
        t = t.cuda(ctx.device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out = t.cpu().numpy().view(moments.dtype).reshape(moments.shape)
    return out


# ---- one process, several devices --------------------------------------------------------------------------------
_local: Optional[Tuple[int, ...]] = None
_local_fused = False


def _connect_local(devices) -> None:
    global _local_fused
    _local_fused = False
    if len(devices) < 2 or len(set(devices)) != len(devices) or os.environ.get("B200MC_FUSED_ALLREDUCE", "1") == "0":
        return
    from . import _ffi
    from .exceptions import MonteCarloError

    try:
        engines = [_ffi.get_engine(d) for d in devices]
        if not all(isinstance(e, _ffi.Engine) for e in engines):
            return  # stand-in engines (CPU tests of the host logic): host-side sum
        _ffi.connect_local(engines)
        _local_fused = True
    except MonteCarloError:
        _local_fused = False  # no peer access between these devices: host-side sum


class local_devices:
    """``with local_devices(4): pricer.price(...)`` — split every simulation over CUDA devices 0..3 of THIS process (or over
    an explicit list of device indices).  Not combinable with the process-group mode."""

    def __init__(self, devices):
        self.devices = tuple(range(devices)) if isinstance(devices, int) else tuple(int(d) for d in devices)
        if not self.devices:
            raise ValueError("at least one device")

    def __enter__(self):
        global _local
        if _ctx is not None and _ctx.world_size > 1:
            raise RuntimeError("local_devices cannot be nested inside a multi-process shard context")
        self._previous, self._previous_fused = _local, _local_fused
        _local = self.devices
        _connect_local(self.devices)
        return self

    def __exit__(self, *exc):
        global _local, _local_fused
        if _local_fused:
            from . import _ffi

            for d in self.devices:
                _ffi.get_engine(d).comm_disconnect()
        _local, _local_fused = self._previous, self._previous_fused
        return False


def local_device_count() -> int:
    """Devices of the active ``local_devices`` block (0 outside one)."""
    return len(_local) if _local is not None else 0


def default_engine():
    """The engine an unsharded call runs on: the first device of a ``local_devices`` block, else this process's device."""
    from . import _ffi

    return _ffi.get_engine(_local[0]) if _local else _ffi.get_engine()


def fused_exchange() -> bool:
    """True when the active sharding (process group or local_devices) adds up moment records inside the kernel."""
    if _local is not None and len(_local) > 1:
        return _local_fused
    return _ctx is not None and _ctx.world_size > 1 and _ctx.fused


def _collective(eng, fn, begin, count, fused_fn):
    """One rank's share of a collective call on connected engines: the all-reduce happens in the kernel's tail.  ``fused_fn``
    asks for it explicitly (b200mc_simulate_allreduce); every other family runs its ordinary entry point with the engine in
    collective mode for the duration of the call (b200mc_comm_set_collective)."""
    if fused_fn is not None:
        return fused_fn(eng, begin, count)
    eng.comm_set_collective(True)
    try:
        return fn(eng, begin, count)
    finally:
        eng.comm_set_collective(False)


def run_sharded(fn, n_units: int, empty, partition=partition_paths, fused_fn=None) -> np.ndarray:
    """Run ``fn(engine, begin, count) -> moments`` over this process's share of ``n_units`` global paths (or Sobol points)
    and return the moments of ALL units: plain call when unsharded; on connected engines (NVLink: torchrun with the NCCL
    backend, or ``local_devices``) every rank / device runs its share - empty shares included - and the kernel's tail adds up
    the moments over peer memory; otherwise (gloo, unconnected devices) partition + ``all_reduce`` / host-side sum, where
    ``empty()`` builds the zero moments of a shard that received no unit."""
    from . import _ffi

    ctx = _ctx
    fused = fused_exchange()
    if _local is not None and len(_local) > 1:
        from concurrent.futures import ThreadPoolExecutor

        def one(i):
            begin, count = partition(n_units, i, len(_local))
            if fused:
                return _collective(_ffi.get_engine(_local[i]), fn, begin, count, fused_fn)
            return fn(_ffi.get_engine(_local[i]), begin, count) if count > 0 else empty()

        with ThreadPoolExecutor(max_workers=len(_local)) as pool:
            parts = list(pool.map(one, range(len(_local))))
        if fused:
            return parts[0]  # every device holds the identical total
        total = np.ascontiguousarray(parts[0]).copy()
        flat = total.view(np.float64)
        for part in parts[1:]:
            flat += np.ascontiguousarray(part).view(np.float64)
        return total
    eng = _ffi.get_engine(_local[0]) if _local else _ffi.get_engine()
    if ctx is None or ctx.world_size == 1:
        return fn(eng, 0, n_units)
    begin, count = partition(n_units, ctx.rank, ctx.world_size)
    if fused:
        return _collective(eng, fn, begin, count, fused_fn)
    local = fn(eng, begin, count) if count > 0 else empty()
    return allreduce_moments(local, ctx)
