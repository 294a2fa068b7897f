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


# ---- the pricer classes themselves under world_size 2 (host logic of the N > 1 path, no GPU) ---------------------
class _OracleEngine:
    """Stands in for _ffi.Engine on the CPU: same simulate() contract (global path range in, MOMENTS out), computed
    by the FP64 oracle evaluation of the engine's own Philox stream.  Test infrastructure only."""

    def simulate(self, spec, params, seed, n_paths, *, stream_base=0, path_begin=0, control_variate=False):
        params = np.asarray(params)
        out = np.zeros(params.shape, dtype=_ffi.MOMENTS_DTYPE)
        for opt in range(params.shape[0]):
            Z = philox_oracle.normals(seed, n_paths, spec.n_steps, stream=stream_base + opt, path_begin=path_begin)
            for k in range(params.shape[1]):
                p = params[opt, k]
                ot = "put" if spec.is_put else "call"
                if spec.kind == _ffi.EUROPEAN:
                    st = orc.gbm_terminal_from_normals(p["S"], p["T"], p["r"], p["sigma"], p["q"], Z)  # antithetic: 2N terminals
                    pay = orc.vanilla_payoffs(st if spec.antithetic else st[: len(Z)], p["K"], ot)
                else:
                    paths = orc.exotic_paths_from_normals(p["S"], p["T"], p["r"], p["sigma"], p["q"], Z)
                    pay = orc.asian_payoffs(paths, p["K"], "arithmetic", ot)
                out[opt, k] = (pay.sum(), (pay**2).sum(), len(pay))
        return out

    def simulate_scalars(self, spec, scenarios, seed, n_paths, *, barrier=0.0, stream_base=0, path_begin=0):
        """The unsharded latency path of _ffi.Engine: one option, (S, K, T, r, sigma, q) tuples -> (sum, sum_sq, n) tuples."""
        sc = np.asarray(scenarios, dtype=np.float64).reshape(-1, 6)
        params = _ffi.make_params(sc[:, 0], sc[:, 1], sc[:, 2], sc[:, 3], sc[:, 4], sc[:, 5], barrier)[None, :]
        m = self.simulate(spec, params, seed, n_paths, stream_base=stream_base, path_begin=path_begin)[0]
        return [(float(x["sum"]), float(x["sum_sq"]), float(x["n"])) for x in m]

    def simulate_structured(self, spec, product, params, seed, n_paths, *, stream_base=0, path_begin=0):
        params = np.asarray(params)
        out = np.zeros(params.shape, dtype=_ffi.MOMENTS_DTYPE)
        for opt in range(params.shape[0]):
            Z = philox_oracle.normals(seed, n_paths, spec.n_steps, stream=stream_base + opt, path_begin=path_begin)
            for k in range(params.shape[1]):
                p = params[opt, k]
                paths = orc.exotic_paths_from_normals(p["S"], p["T"], p["r"], p["sigma"], p["q"], Z)
                if spec.kind == _ffi.CLIQUET:
                    pay = orc.cliquet_payoffs(paths, p["S"], product.a, product.b, product.c, product.d, product.period)
                else:
                    pay = orc.autocallable_payoffs(paths, p["S"], p["T"], p["r"], product.a, product.b, product.c, product.d, product.period)
                out[opt, k] = (pay.sum(), (pay**2).sum(), len(pay))
        return out


def _price_everything():
    import optionslab_b200 as ob

    pr = ob.MonteCarloPricer(N_PATHS, N_STEPS, seed=SEED)
    res = pr.price(**P, option_type="put", return_error=True)
    greeks = pr.greeks(**P, option_type="call", include_second_order=False)
    grid = ob.MonteCarloPricerUni(N_PATHS, N_STEPS, seed=SEED).price_batch([100.0, 101.0], [105.0, 95.0], 0.5, 0.03, 0.25, "call")
    asian = ob.AsianOption(**P, seed=SEED).price(n_paths=N_PATHS, n_steps=N_STEPS, return_error=True)
    auto = ob.AutocallableOption(**P, seed=SEED).price(N_PATHS, N_STEPS, 4, return_error=True)
    cliq = ob.CliquetOption(**P, seed=SEED).price_scenarios([(100.0, 105.0, 0.5, 0.03, 0.25, 0.0), (101.0, 105.0, 0.5, 0.03, 0.26, 0.0)],
                                                            n_paths=N_PATHS, n_steps=N_STEPS, n_periods=3)
    return np.array([res.price, res.std_error, res.n_paths, greeks["delta"], greeks["vega"], grid[0], grid[1], asian.price, asian.n_paths,
                     auto.price, auto.std_error, auto.n_paths, cliq[0], cliq[1]])


def _pricer_worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    _ffi.get_engine = lambda device=None: _OracleEngine()
    distributed.init(backend="gloo")
    np.save(os.path.join(out_dir, f"pricers{rank}.npy"), _price_everything())
    distributed.shutdown()


def test_pricer_classes_shard_paths_and_allreduce(tmp_path, monkeypatch):
    """MonteCarloPricer / MonteCarloPricerUni.price_batch / AsianOption / AutocallableOption / CliquetOption on 2 ranks == 1 process: every rank simulates
    its slice of the global path range (runtime.simulate) and the all-reduced moments give identical prices, Greeks,
    standard errors and sample counts on all ranks."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_pricer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    distributed.shutdown()
    monkeypatch.setattr(_ffi, "get_engine", lambda device=None: _OracleEngine())
    whole = _price_everything()
    for rank in range(2):
        np.testing.assert_allclose(np.load(os.path.join(tmp_path, f"pricers{rank}.npy")), whole, rtol=1e-12)
    assert whole[2] == 2 * N_PATHS and whole[8] == N_PATHS  # antithetic European counts 2N samples, the Asian N


def test_local_devices_mode_splits_paths_over_engines_and_sums(monkeypatch):
    """One process, several devices: every device's engine gets its slice of the global path range (threads), the partial
    moments are summed on the host; the prices equal the single-engine ones.  Engines are oracle stand-ins here."""
    import optionslab_b200 as ob

    seen = []

    class Recording(_OracleEngine):
        def __init__(self, device):
            self.device = device

        def simulate(self, spec, params, seed, n_paths, **kw):
            seen.append((self.device, kw.get("path_begin", 0), n_paths))
            return super().simulate(spec, params, seed, n_paths, **kw)

    engines = {}
    monkeypatch.setattr(_ffi, "get_engine", lambda device=None: engines.setdefault(device or 0, Recording(device or 0)))
    distributed.shutdown()
    pr = ob.MonteCarloPricer(N_PATHS, N_STEPS, seed=SEED)
    whole = pr.price(**P, option_type="call", return_error=True)
    seen.clear()
    with distributed.local_devices(3):
        split = pr.price(**P, option_type="call", return_error=True)
        asian = ob.AsianOption(**P, seed=SEED).price(n_paths=N_PATHS, n_steps=N_STEPS)
    assert sorted(seen[:3]) == [(0, 0, 3334), (1, 3334, 3334), (2, 6668, 3333)]
    assert split.n_paths == whole.n_paths and split.price == pytest.approx(whole.price, rel=1e-12) and split.std_error == pytest.approx(whole.std_error, rel=1e-12)
    assert asian == pytest.approx(ob.AsianOption(**P, seed=SEED).price(n_paths=N_PATHS, n_steps=N_STEPS), rel=1e-12)
    with distributed.local_devices([0]):  # a single device is the plain path
        assert pr.price(**P, option_type="call") == pytest.approx(whole.price, rel=1e-15)
