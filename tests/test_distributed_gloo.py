"""CPU, world_size 2, gloo: the sharding contract of the multi-GPU path.

Each rank evaluates only its partition of the global path range (here with the CPU oracle of the
engine's Philox stream, since there is no GPU in this container) and the single all-reduce of
(sum, sum^2, n) must reproduce the one-process moments over the whole range.
"""

import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from optionslab_b200 import _ffi, distributed
from oracle import philox_oracle, reference_mc as orc

N_PATHS, N_STEPS, SEED = 10_001, 12, 99  # odd path count: ranks get unequal shares
P = dict(S=100.0, K=105.0, T=0.5, r=0.03, sigma=0.25)


def _moments_for_range(begin, count):
    m = np.zeros((3, 1), dtype=_ffi.MOMENTS_DTYPE)
    for opt in range(3):  # option i draws from stream i, like price_batch
        if count == 0:
            continue
        Z = philox_oracle.normals(SEED, count, N_STEPS, stream=opt, path_begin=begin)
        pay = orc.vanilla_payoffs(orc.gbm_terminal_from_normals(P["S"], P["T"], P["r"], P["sigma"], 0.0, Z), P["K"] + opt, "call")
        m[opt, 0] = (pay.sum(), (pay**2).sum(), len(pay))
    return m


def _worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    ctx = distributed.init(backend="gloo")
    assert (ctx.rank, ctx.world_size) == (rank, world)
    begin, count = distributed.partition_paths(N_PATHS, rank, world)
    total = distributed.allreduce_moments(_moments_for_range(begin, count))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), total)
    distributed.shutdown()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_moments_equal_single_process(tmp_path, world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    whole = _moments_for_range(0, N_PATHS)
    for rank in range(world):
        got = np.load(os.path.join(tmp_path, f"rank{rank}.npy"))
        assert got["n"].tolist() == whole["n"].tolist()
        np.testing.assert_allclose(got["sum"], whole["sum"], rtol=1e-13)
        np.testing.assert_allclose(got["sum_sq"], whole["sum_sq"], rtol=1e-13)


def test_unsharded_allreduce_is_identity():
    distributed.shutdown()
    m = _moments_for_range(0, 100)
    assert distributed.allreduce_moments(m) is m
