"""pytest configuration: markers and shared fixtures."""

import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: long-running statistical case")


@pytest.fixture(scope="session")
def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "reference_goldens.json")) as f:
        return json.load(f)


@pytest.fixture
def anchored(goldens):
    """``anchored()`` -> True when the goldens were recorded with this NumPy (``scipy=True``: and this SciPy) build, i.e.
    when the values recorded from the REAL reference can be compared draw for draw.  A test that had to leave an anchor
    uncompared still runs its other checks and is then reported as one SKIP, so a drifted NumPy shows up in the count
    instead of silently dropping the real-reference comparison."""
    import numpy as np

    missed = []

    def same(scipy=False):
        ok = goldens["numpy"] == np.__version__
        if scipy:
            import scipy as sp

            ok = ok and goldens["scipy"] == sp.__version__
        if not ok:
            missed.append(1)
        return ok

    yield same
    if missed:
        pytest.skip(f"{len(missed)} real-reference golden comparison(s) not made: goldens were recorded with NumPy {goldens['numpy']} / "
                    f"SciPy {goldens['scipy']}")


def _gpu_present() -> bool:
    return os.path.exists("/dev/nvidiactl") or os.path.exists("/dev/nvidia0")


def pytest_collection_modifyitems(config, items):
    """Without a device the gpu-marked tests are skipped (plain ``pytest tests`` on a CPU box then reports skips, not
    errors from b200mc_create); with ``-m gpu`` on the GPU box nothing is skipped here."""
    if _gpu_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this machine (gpu-marked tests run on the B200 box)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def engine():
    """One CUDA engine handle for the whole GPU test session."""
    from optionslab_b200 import _ffi

    yield _ffi.get_engine(0)  # the same process-wide engine the pricer classes use
