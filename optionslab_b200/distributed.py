"""Multi-GPU sharding of the Monte Carlo path: one process per GPU, paths partitioned by global
index, ONE all-reduce of the per-(option, scenario) moments.

Every (option, path) is independent: rank g of G simulates global paths
[g*N/G, (g+1)*N/G) of every option — the Philox counter carries the global path index, so the
draws are disjoint by construction and independent of G — and the only exchanged data are the
FP64 (sum, sum^2, n) triples.  ``torch.distributed`` is plumbing only (NCCL over NVLink on the
GPU box, gloo in the CPU tests); the simulation never touches torch.
"""

from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

__all__ = ["ShardContext", "init", "shutdown", "current", "partition_paths", "allreduce_moments"]


@dataclass
class ShardContext:
    rank: int
    world_size: int
    backend: str
    device: Optional[int]  # CUDA device index for NCCL, None for gloo/CPU
    owns_group: bool = False


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
    return _ctx


def shutdown():
    global _ctx
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
