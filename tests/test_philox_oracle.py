"""CPU: pin the Philox restatements (C oracle and the host build of the device header) to the
Random123 known answers, and sanity-check the documented uniform->normal mapping."""

import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import philox_oracle as po
from tests import kat

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_header_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("hostphilox") / "libhost_philox.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(out), os.path.join(ROOT, "tests", "host_philox.cpp")], check=True)
    lib = C.CDLL(str(out))
    lib.host_philox4x32_10.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    lib.host_philox4x32_10_keyed.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
    return lib


def test_c_oracle_known_answers():
    for ctr, key, want in kat.PHILOX4X32_10_KAT:
        assert tuple(int(x) for x in po.philox4x32_10(ctr, key)) == want


def test_device_header_known_answers_and_matches_oracle(host_header_lib):
    ck = kat.kat_inputs()
    out = np.empty((len(ck), 4), dtype=np.uint32)
    host_header_lib.host_philox4x32_10(ck.ctypes.data, len(ck), out.ctypes.data)
    np.testing.assert_array_equal(out, kat.kat_outputs())
    rng = np.random.default_rng(0)
    ck = rng.integers(0, 2**32, size=(2000, 6), dtype=np.uint64).astype(np.uint32)
    out = np.empty((len(ck), 4), dtype=np.uint32)
    host_header_lib.host_philox4x32_10(ck.ctypes.data, len(ck), out.ctypes.data)
    for i in range(0, len(ck), 97):
        np.testing.assert_array_equal(out[i], po.philox4x32_10(ck[i, :4], ck[i, 4:]))


def test_keyed_form_used_by_the_kernels_equals_the_reference_form(host_header_lib):
    """philox_expand_key + philox4x32_10 (round keys as launch constants, what every simulation kernel calls) against the
    known answers and, on random counters / keys, against the textbook loop that bumps the key every round."""
    ck = kat.kat_inputs()
    out = np.empty((len(ck), 4), dtype=np.uint32)
    host_header_lib.host_philox4x32_10_keyed(ck.ctypes.data, len(ck), out.ctypes.data)
    np.testing.assert_array_equal(out, kat.kat_outputs())
    rng = np.random.default_rng(5)
    ck = rng.integers(0, 2**32, size=(5000, 6), dtype=np.uint64).astype(np.uint32)
    a, b = np.empty((len(ck), 4), dtype=np.uint32), np.empty((len(ck), 4), dtype=np.uint32)
    host_header_lib.host_philox4x32_10(ck.ctypes.data, len(ck), a.ctypes.data)
    host_header_lib.host_philox4x32_10_keyed(ck.ctypes.data, len(ck), b.ctypes.data)
    np.testing.assert_array_equal(a, b)


def test_stream_layout_is_counter_based():
    """Path p / step s depends only on (seed, stream, p, s): sub-ranges and prefixes agree."""
    full = po.normals(11, 64, 10, stream=3)
    np.testing.assert_array_equal(po.normals(11, 16, 10, stream=3, path_begin=40), full[40:56])
    np.testing.assert_array_equal(po.normals(11, 64, 7, stream=3), full[:, :7])
    assert not np.array_equal(po.normals(11, 64, 10, stream=4), full)
    assert not np.array_equal(po.normals(12, 64, 10, stream=3), full)
    big = po.normals(5, 2, 4, path_begin=(1 << 32) - 1)  # crosses the 32-bit path-word boundary
    assert np.all(np.isfinite(big)) and not np.array_equal(big[0], big[1])


def test_normal_mapping_moments():
    z = po.normals(2024, 250_000, 8).ravel()  # 2e6 draws
    n = z.size
    assert abs(z.mean()) < 4 / np.sqrt(n)
    assert abs(z.var() - 1) < 4 * np.sqrt(2 / n)
    assert abs((z**3).mean()) < 4 * np.sqrt(15 / n)
    assert abs((z**4).mean() - 3) < 4 * np.sqrt(96 / n)
    assert np.abs(z).max() < 5.7  # radius is capped at sqrt(2*23*ln2) = 5.647 by the 23-bit uniform
    # Kolmogorov-Smirnov against the normal CDF
    from scipy import stats
    assert stats.kstest(z[:200_000], "norm").pvalue > 1e-3
    # successive draws of a path and neighbouring paths are uncorrelated
    zz = z.reshape(-1, 8)
    assert abs(np.corrcoef(zz[:, 0], zz[:, 1])[0, 1]) < 4 / np.sqrt(len(zz))
    assert abs(np.corrcoef(zz[:-1, 0], zz[1:, 0])[0, 1]) < 4 / np.sqrt(len(zz))
    # the two normals of one Box-Muller pair share a radius: their squares must still be uncorrelated
    assert abs(np.corrcoef(zz[:, 0] ** 2, zz[:, 1] ** 2)[0, 1]) < 4 / np.sqrt(len(zz))
    assert abs(np.corrcoef(zz[:, 1] ** 2, zz[:, 2] ** 2)[0, 1]) < 4 / np.sqrt(len(zz))


def test_normal_mapping_ks_at_3e7_draws_and_symmetric_tails():
    """The (radius, angle) bit layout matters at this sample size: a fixed 512-angle grid, or fine angle bits
    taken from the radius' leading bits, both fail here."""
    from scipy import stats
    z = po.normals(2025, 4_000_000, 8).ravel()
    assert stats.kstest(z, "norm").pvalue > 1e-3
    expect = z.size * stats.norm.sf(4.0)
    for tail in ((z > 4.0).sum(), (z < -4.0).sum()):
        assert abs(tail - expect) < 5 * np.sqrt(expect)


def test_single_step_paths_draw_one_64_bit_normal():
    """n_steps == 1 (the reference's default): the path's one normal takes the whole first word for the radius
    (u = (w0 + 1) 2^-32, FP32-converted as on the device) and the top 23 bits of the second word for the angle
    (normal.cuh, box_muller_single).  Contract restated from the raw Philox words, then moments / KS / tails."""
    from scipy import stats

    seed, n = 31, 4_000_000
    z = po.normals(seed, n, 1).ravel()
    for path in (0, 1, 12345, n - 1):
        w = po.philox4x32_10([path, 0, 0, 0], [seed, 0])
        u = np.float32(np.float32(w[0]) * np.float32(2.0**-32) + np.float32(2.0**-32))  # fmaf: exact product, one rounding
        turns = np.frombuffer(np.uint32((int(w[1]) >> 9) | 0x3F800000).tobytes(), dtype=np.float32)[0]
        theta = float(turns) * float(np.float32(6.28318530717958647692)) + float(np.float32(-9.42477796076937971538))
        assert z[path] == pytest.approx(np.sqrt(-2.0 * np.log(float(u))) * np.cos(theta), rel=1e-12, abs=1e-15)
    assert abs(z.mean()) < 4 / np.sqrt(n) and abs(z.var() - 1) < 4 * np.sqrt(2 / n) and abs((z**4).mean() - 3) < 4 * np.sqrt(96 / n)
    assert stats.kstest(z, "norm").pvalue > 1e-3
    expect = n * stats.norm.sf(3.5)
    for tail in ((z > 3.5).sum(), (z < -3.5).sum()):
        assert abs(tail - expect) < 5 * np.sqrt(expect)
    # a two-step path is NOT the single-step draw followed by another one: the layouts differ (32 vs 64 bits)
    assert not np.allclose(po.normals(seed, 16, 2)[:, 0], z[:16])


def test_word_to_pair_mapping_matches_the_documented_contract():
    """normal.cuh: word n = 4j+i of a path -> steps 2n, 2n+1; radius mantissa = top 23 bits, angle mantissa =
    top 23 bits of the byte-reversed word (turns), radius normalised so that the 2^23-point grid has E[r^2] = 2."""
    seed, stream, path = 99, 5, 123456789012
    z = po.normals(seed, 1, 16, stream=stream, path_begin=path)[0]
    rad_norm = 1.000000529893528569531
    two_pi, m3pi = float(np.float32(6.28318530717958647692)), float(np.float32(-9.42477796076937971538))
    for j in range(2):
        w = po.philox4x32_10([path & 0xFFFFFFFF, j, path >> 32, stream], [seed, 0])
        for i in range(4):
            word = int(w[i])
            u = 2.0 - (1.0 + (word >> 9) / 2.0**23)
            rev = int.from_bytes(word.to_bytes(4, "little"), "big")
            turns = 1.0 + (rev >> 9) / 2.0**23
            theta = turns * two_pi + m3pi
            r = rad_norm * np.sqrt(-2.0 * np.log(u))
            n = 4 * j + i
            assert z[2 * n] == pytest.approx(r * np.cos(theta), abs=1e-14)
            assert z[2 * n + 1] == pytest.approx(r * np.sin(theta), abs=1e-14)


def test_radius_grid_second_moment_is_normalised():
    """E[-2 ln u] over u = j/2^23 (j = 1..2^23) times kRadNorm^2 equals 2 (so E[z^2] = 1 on the grid)."""
    from math import lgamma, log
    M = 2.0**23
    e_r2 = 2.0 * (log(M) - lgamma(M + 1.0) / M)
    assert e_r2 * 1.000000529893528569531**2 == pytest.approx(2.0, rel=1e-9)
