"""CPU: the UNMODIFIED reference installed by oracle/build_ref.py (oracle/_ref, git-ignored, shipped to the GPU box for
``bench.py --impl reference`` / ``cpu_baseline``) against the NumPy restatement the parity tests use
(oracle/reference_mc.py): bit for bit where both exist.  Skipped when the install is absent (a checkout without
/root/reference); the restatement is then still pinned by tests/test_oracle_golden.py."""

import json
import os

import numpy as np
import pytest

from oracle import build_ref
from oracle import reference_mc as orc

pytestmark = pytest.mark.skipif(not build_ref.available(), reason="oracle/_ref not installed (python oracle/build_ref.py needs /root/reference)")
P = dict(S=100.0, K=100.0, T=1.0, r=0.05, sigma=0.2)


@pytest.fixture(scope="module")
def ref():
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return build_ref.load()


def test_install_is_byte_identical_to_the_manifest():
    import hashlib

    manifest = json.load(open(build_ref.MANIFEST))
    assert len(manifest["files"]) > 50
    for rel, digest in manifest["files"].items():
        with open(os.path.join(build_ref.REF_DIR, rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == digest, rel
    if os.path.isdir("/root/reference/src"):  # in the build container: the manifest is the reference itself
        for rel, digest in manifest["files"].items():
            with open(os.path.join("/root/reference", rel), "rb") as f:
                assert hashlib.sha256(f.read()).hexdigest() == digest, rel


@pytest.mark.parametrize("n_sims,n_steps", [(10_000, 50), (20_000, 1), (4096, 7)])
@pytest.mark.parametrize("ot", ["call", "put"])
def test_european_restatement_equals_the_installed_reference(ref, n_sims, n_steps, ot):
    want = ref.MonteCarloPricer(n_sims, n_steps, seed=42).price(**P, option_type=ot, q=0.01, return_error=True)
    got = orc.european_price(**P, option_type=ot, q=0.01, num_simulations=n_sims, num_steps=n_steps, seed=42)
    assert (got.price, got.std_error, got.n_paths) == (want.price, want.std_error, want.n_paths)
    uni = ref.MonteCarloPricerUni(n_sims, max(n_steps, 1), seed=42, use_numba=False)
    assert orc.uni_price(**P, option_type=ot, num_simulations=n_sims, num_steps=max(n_steps, 1), seed=42) == uni.price(**P, option_type=ot)


def test_exotics_and_greeks_restatement_equal_the_installed_reference(ref):
    kw = dict(seed=42, n_paths=5000, n_steps=12)
    for avg in ("arithmetic", "geometric"):
        for ot in ("call", "put"):
            want = ref.AsianOption(**P, seed=42).price(5000, 12, avg, ot)
            assert orc.exotic_price("asian", **P, **kw, avg_type=avg, option_type=ot) == want
    for bt in ("up-and-out", "up-and-in", "down-and-out", "down-and-in"):
        B = 120.0 if bt.startswith("up") else 85.0
        assert orc.exotic_price("barrier", **P, **kw, barrier=B, barrier_type=bt) == ref.BarrierOption(**P, seed=42, barrier=B).price(5000, 12, bt)
    pr = ref.MonteCarloPricer(5000, 10, seed=3)
    want = ref.compute_greeks_unified(pr, **P, option_type="call")
    got = orc.greeks_bump_and_revalue(lambda S, K, T, r, s, q: orc.european_price(S, K, T, r, s, "call", q, num_simulations=5000, num_steps=10, seed=3).price, **P)
    assert dict(got) == dict(want)
