"""TEST / BASELINE INFRASTRUCTURE ONLY — install the UNMODIFIED reference where the GPU box can run it.

The reference (Diegotistical/OptionsLab) is pure Python: there is nothing to compile, and ``/root/reference`` does not
exist on the GPU box.  This recipe is the Python counterpart of "compile the reference's own sources into
``oracle/_ref``": it installs the reference's ``src`` package - byte for byte, nothing edited - from where it lies
under ``/root/reference`` into ``oracle/_ref/`` (git-ignored: never part of this repository's history; NOT
gpurun-ignored: it travels to the GPU box like a built ``.so``), next to a pass-through ``streamlit`` stub, because
``src/utils/decorators/caching.py:3`` imports streamlit for two decorators and the image has none.

Used by ``bench.py --impl reference`` / ``cpu_baseline`` (kind "reference": the reference's own NumPy and Numba
backends timed on the box's host cores) and by ``tests/test_reference_install.py`` (the NumPy restatement in
``oracle/reference_mc.py`` equals the installed reference bit for bit).  Nothing under ``optionslab_b200/`` touches it.

    python oracle/build_ref.py          # (re)install when /root/reference is present
"""

from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = os.environ.get("OPTIONSLAB_REFERENCE", "/root/reference")
MANIFEST = os.path.join(REF_DIR, "MANIFEST.json")

STREAMLIT_STUB = '''"""Pass-through stand-in for streamlit (oracle/build_ref.py): the reference only uses these two decorators."""


def _passthrough(*dargs, **dkwargs):
    if len(dargs) == 1 and callable(dargs[0]) and not dkwargs:
        return dargs[0]
    return lambda fn: fn


cache_data = _passthrough
cache_resource = _passthrough
'''


def _sha(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def build(force: bool = False) -> str | None:
    """Install ``<reference>/src`` into oracle/_ref/src.  Returns the directory, or None when neither the reference nor a
    previous install is present (the GPU box only ever uses the prebuilt directory)."""
    src = os.path.join(REFERENCE_ROOT, "src")
    if not os.path.isdir(src):
        return REF_DIR if os.path.exists(MANIFEST) else None
    if os.path.exists(MANIFEST) and not force:
        return REF_DIR
    if os.path.isdir(REF_DIR):
        shutil.rmtree(REF_DIR)
    os.makedirs(REF_DIR)
    shutil.copytree(src, os.path.join(REF_DIR, "src"), ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    os.makedirs(os.path.join(REF_DIR, "_stubs"))
    with open(os.path.join(REF_DIR, "_stubs", "streamlit.py"), "w") as f:
        f.write(STREAMLIT_STUB)
    files = {}
    for root, _, names in os.walk(os.path.join(REF_DIR, "src")):
        for n in sorted(names):
            p = os.path.join(root, n)
            rel = os.path.relpath(p, REF_DIR)
            files[rel] = _sha(p)
            assert files[rel] == _sha(os.path.join(REFERENCE_ROOT, rel)), f"{rel} differs from the reference"
    with open(MANIFEST, "w") as f:
        json.dump({"source": REFERENCE_ROOT, "files": files}, f, indent=1, sort_keys=True)
    return REF_DIR


def available() -> bool:
    return os.path.exists(MANIFEST)


_loaded = None


def load() -> types.SimpleNamespace:
    """Import the installed reference (its own modules, unmodified) and return the classes of the Monte Carlo path."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise ImportError("oracle/_ref is not installed (python oracle/build_ref.py needs /root/reference)")
    try:
        import streamlit  # noqa: F401
    except ImportError:
        sys.path.insert(0, os.path.join(REF_DIR, "_stubs"))
    if "src" in sys.modules and not getattr(sys.modules["src"], "__file__", "").startswith(REF_DIR):
        raise ImportError("another top-level package named 'src' is already imported")
    sys.path.insert(0, REF_DIR)
    from src.greeks.unified_greeks import ExoticAdapter, compute_greeks_unified
    from src.pricing_models.black_scholes import black_scholes
    from src.pricing_models.exotic_options import AsianOption, BarrierOption, LookbackOption
    from src.pricing_models.monte_carlo import MCMethod, MonteCarloPricer
    from src.pricing_models.monte_carlo_unified import MonteCarloPricerUni
    from src.simulation import gbm_numba, gbm_numpy

    _loaded = types.SimpleNamespace(MonteCarloPricer=MonteCarloPricer, MCMethod=MCMethod, MonteCarloPricerUni=MonteCarloPricerUni,
                                    AsianOption=AsianOption, BarrierOption=BarrierOption, LookbackOption=LookbackOption,
                                    compute_greeks_unified=compute_greeks_unified, ExoticAdapter=ExoticAdapter,
                                    black_scholes=black_scholes, gbm_numpy=gbm_numpy, gbm_numba=gbm_numba, root=REF_DIR)
    return _loaded


if __name__ == "__main__":
    print(build(force=True))
