"""CPU: libb200mc.so loads without a GPU and exports exactly what include/b200mc.h declares."""

import ctypes as C
import os
import re

import numpy as np
import pytest

from optionslab_b200 import _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "b200mc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200mc_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _header_functions() == sorted(_ffi.SIGNATURES)


def test_library_exports_every_declared_symbol():
    lib = _ffi.load_library()
    for name in _header_functions():
        assert hasattr(lib, name), name
    assert lib.b200mc_abi_version() == _ffi.ABI_VERSION


def test_struct_layouts_match_header():
    assert C.sizeof(_ffi.Spec) == 32
    assert _ffi.PARAMS_DTYPE.itemsize == 64 and _ffi.MOMENTS_DTYPE.itemsize == 24
    assert C.sizeof(_ffi.Info) == 4 * 6 + 8 + 4 * 2 + 64
    assert C.sizeof(_ffi.Peaks) == 12 * 8
    assert C.sizeof(_ffi.Product) == 4 * 8 + 2 * 4 and _ffi.Product.period.offset == 32
    p = _ffi.make_params(100.0, [90.0, 110.0], 1.0, 0.05, 0.2, 0.0, 120.0)
    assert p.shape == (2,) and p["K"].tolist() == [90.0, 110.0] and p["barrier"].tolist() == [120.0, 120.0]


def test_null_engine_calls_are_rejected_not_crashing():
    lib = _ffi.load_library()
    assert lib.b200mc_kernel_launches(None) == 0
    assert lib.b200mc_set_kernel_timing(None, 1) == -1
    assert lib.b200mc_simulate(None, None, None, 1, 1, 0, 0, 0, 1, None) == -1
    lib.b200mc_destroy(None)


def test_no_cpu_fallback_without_a_device():
    """On a box with no GPU the product path must fail loudly, never compute on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    from optionslab_b200 import AccelerationError, AsianOption, MonteCarloPricer, MonteCarloPricerUni
    _ffi._engines.clear()
    with pytest.raises(AccelerationError):
        MonteCarloPricer(1000, 4, seed=1).price(100, 100, 1.0, 0.05, 0.2, "call")
    with pytest.raises(AccelerationError):
        AsianOption(100, 100, 1.0, 0.05, 0.2, seed=1).price(1000, 4)
    from optionslab_b200 import MonteCarloError
    with pytest.raises(MonteCarloError):  # Uni wraps engine failures like monte_carlo_unified.py:510-511
        MonteCarloPricerUni(1000, 4, seed=1).price(100, 100, 1.0, 0.05, 0.2, "call")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "optionslab_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "oracle/" not in text or f.endswith((".cuh", ".cu")), f


def test_header_is_plain_c_and_a_c_client_links():
    """include/b200mc.h must be consumable from C (the boundary is extern "C", plain pointers and sizes): compile it as
    strict C99 and link a tiny C client against libb200mc.so that only calls b200mc_abi_version() (no device needed)."""
    import shutil
    import subprocess
    import tempfile

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    header = os.path.join(ROOT, "include", "b200mc.h")
    subprocess.run([gcc, "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", header], check=True)
    with tempfile.TemporaryDirectory() as d:
        src = os.path.join(d, "client.c")
        with open(src, "w") as f:
            f.write('#include "b200mc.h"\n#include <stdio.h>\nint main(void) { b200mc_spec_t s = {0}; s.flags = B200MC_FLAG_EXACT_EX2;\n'
                    '  printf("%d %u\\n", b200mc_abi_version(), (unsigned)sizeof(b200mc_params_t) + s.flags); return 0; }\n')
        exe = os.path.join(d, "client")
        libdir = os.path.dirname(_ffi.LIB_PATH)
        subprocess.run([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", libdir, "-l:libb200mc.so",
                        f"-Wl,-rpath,{libdir}"], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split()
        assert out == [str(_ffi.ABI_VERSION), "65"]


@pytest.mark.gpu
def test_c_client_prices_through_the_c_abi_bit_for_bit(tmp_path):
    """a call THROUGH the boundary from C.  tests/c_client/price_client.c does b200mc_create ->
    b200mc_simulate (one option / three CRN scenarios, then a five-option barrier batch) -> b200mc_destroy and prints the
    moments as exact hexadecimal doubles; the Python pricer path (ctypes onto the same entry points) must return the
    same bits, and the delta the C client's moments imply equals MonteCarloPricerUni.delta_gamma at the default bump."""
    import shutil
    import subprocess

    import numpy as np

    import optionslab_b200 as ob

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    exe = str(tmp_path / "price_client")
    libdir = os.path.dirname(_ffi.LIB_PATH)
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_client", "price_client.c"), "-o", exe, "-L", libdir, "-l:libb200mc.so", f"-Wl,-rpath,{libdir}"],
                   check=True)
    lines = subprocess.run([exe, "0"], check=True, capture_output=True, text=True).stdout.splitlines()
    rows = {(f[0], int(f[1])): tuple(float.fromhex(x) for x in f[2:5]) for f in (ln.split() for ln in lines) if f[0] in ("european", "barrier")}
    assert len(rows) == 8
    eng = _ffi.get_engine(0)
    h = 1e-4
    params = np.stack([_ffi.make_params(100.0 + b, 100.0, 1.0, 0.05, 0.2, 0.01) for b in (h, 0.0, -h)]).reshape(1, 3)
    m = eng.simulate(_ffi.make_spec(_ffi.EUROPEAN, 50, antithetic=True), params, 42, 100_000)[0]
    for i in range(3):
        assert rows[("european", i)] == (m["sum"][i], m["sum_sq"][i], m["n"][i])
    K = 90.0 + 5.0 * np.arange(5)
    mb = eng.simulate(_ffi.make_spec(_ffi.BARRIER, 64), _ffi.make_params(100.0, K, 0.5, 0.03, 0.25, 0.0, 125.0).reshape(5, 1), 7, 200_001,
                      stream_base=3, path_begin=1000)[:, 0]
    for i in range(5):
        assert rows[("barrier", i)] == (mb["sum"][i], mb["sum_sq"][i], mb["n"][i])
    disc = float(np.exp(-0.05))
    up, down = (disc * rows[("european", i)][0] / rows[("european", i)][2] for i in (0, 2))
    delta, _ = ob.MonteCarloPricerUni(100_000, 50, seed=1).delta_gamma(100.0, 100.0, 1.0, 0.05, 0.2, "call", q=0.01, seed=42)
    assert (up - down) / (2 * h) == delta
    invalid = [ln for ln in lines if ln.startswith("invalid")][0]
    assert invalid.startswith("invalid -1 ") and "n_steps" in invalid
    assert [ln for ln in lines if ln.startswith("launches")] == ["launches 2"]
