#!/usr/bin/env python
"""RNG evidence at GPU scale: statistics of ~1e10 device normals (b200mc_rng_statistics) and the
single-step strike sweep at 2^32 samples.  Importable (tests/test_gpu_rng.py uses the same functions); as a script it
writes one JSON document (profiles/r02_rng_evidence.json is a committed run)."""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# Multi-step streams: the radius takes 2^23 equally spaced u = j / 2^23 (normal.cuh), so beyond ~4.5 sigma the law thins out:
# P(z > c) relative to the normal law, by summing arccos(c / r_j) / pi over the grid (the angle is continuous to 2^-23).
GRID_TAIL_RATIO = {4.0: 1 - 4.9e-4, 4.5: 1 - 4.0e-3, 5.0: 1 - 3.74e-2}


def _phi(x):
    return 0.5 * math.erfc(-x / math.sqrt(2.0))


def chi_square(counts, probs, n, min_expected=50.0):
    """Pearson chi-square of observed counts against cell probabilities; neighbouring cells (in array order) are merged
    until every merged cell expects >= min_expected.  -> (chi2, degrees of freedom)."""
    counts = np.asarray(counts, dtype=np.float64).ravel()
    expected = np.asarray(probs, dtype=np.float64).ravel() * n
    order = np.argsort(expected)  # merge the smallest cells with each other first
    counts, expected = counts[order], expected[order]
    chi2, cells, c_acc, e_acc = 0.0, 0, 0.0, 0.0
    for c, e in zip(counts, expected):
        c_acc += c
        e_acc += e
        if e_acc >= min_expected:
            chi2 += (c_acc - e_acc) ** 2 / e_acc
            cells += 1
            c_acc = e_acc = 0.0
    if e_acc > 0:
        chi2 += (c_acc - e_acc) ** 2 / max(e_acc, 1e-300) if e_acc >= 5 else 0.0
        cells += 1 if e_acc >= 5 else 0
    return chi2, cells - 1


def stream_statistics(eng, seed=2026, n_paths=1 << 25, n_steps=320):
    """Statistics of n_paths x n_steps normals.  The joint (same-word) chi-square is only meaningful while the sample is
    small against the 2^32 distinct (z1, z2) points a 32-bit word can produce: tests assert it at 2.7e8 pairs
    (n_paths = 2^22, n_steps = 128); at 5e9 pairs the sample exceeds the lattice and the chi-square resolves it."""
    st = eng.rng_statistics(seed, n_paths, n_steps)
    m = st["moments"]
    n, n_pairs, n_lag = m[11], m[12], m[13]
    edges = np.linspace(-6.0, 6.0, 257)
    cdf = np.array([_phi(x) for x in edges])
    pz = np.diff(cdf)
    pz[0] += cdf[0]
    pz[-1] += 1.0 - cdf[-1]
    chi_all, df_all = chi_square(st["hist_z"], pz, n)  # against the NORMAL law, all 256 bins: resolves the grid's thin tail at 1e10 draws
    # the body |z| <= 4.5 (192 bins, where the 2^23-point radius grid follows the normal law to < 0.4%) plus one cell per tail,
    # the tails against the grid law
    hz = np.asarray(st["hist_z"], dtype=np.float64)
    lo, hi = 32, 224  # bin edges -4.5 and +4.5
    body_counts = np.concatenate([[hz[:lo].sum()], hz[lo:hi], [hz[hi:].sum()]])
    tail_p = (1.0 - _phi(4.5)) * GRID_TAIL_RATIO[4.5]
    body_p = np.concatenate([[tail_p], pz[lo:hi], [tail_p]])
    body_p[1:-1] *= (1.0 - 2.0 * tail_p) / body_p[1:-1].sum()
    chi_z, df_z = chi_square(body_counts, body_p, n)
    e2 = np.linspace(-4.0, 4.0, 65)
    c2 = np.array([_phi(x) for x in e2])
    p1 = np.diff(c2)
    p1[0] += c2[0]
    p1[-1] += 1.0 - c2[-1]
    chi_j, df_j = chi_square(st["hist_joint"], np.outer(p1, p1), n_pairs)
    sf = lambda c: 0.5 * math.erfc(c / math.sqrt(2.0))
    tails = {}
    for name, count, c in (("z>4", st["tails"][0], 4.0), ("z>5", st["tails"][1], 5.0), ("z<-4", st["tails"][2], 4.0), ("z<-5", st["tails"][3], 5.0)):
        tails[name] = {"count": int(count), "normal_law": n * sf(c), "grid_law": n * sf(c) * GRID_TAIL_RATIO[c]}
    return {
        "draws": n, "same_word_pairs": n_pairs, "lag1_pairs": n_lag, "seed": seed, "n_paths": n_paths, "n_steps": n_steps,
        "mean": m[0] / n, "second_moment": m[1] / n, "third_moment": m[2] / n, "fourth_moment": m[3] / n,
        "same_word": {"E[z1 z2]": m[4] / n_pairs, "E[z1^2 z2^2]": m[5] / n_pairs, "E[z1 z2^3]": m[6] / n_pairs, "E[z1^3 z2]": m[7] / n_pairs},
        "lag1": {"E[a b]": m[8] / n_lag, "E[a^2 b^2]": m[9] / n_lag, "E[a b^3]": m[10] / n_lag},
        "chi2_z": {"chi2": chi_z, "df": df_z, "z_score": (chi_z - df_z) / math.sqrt(2 * df_z),
                   "cells": "192 bins of width 3/64 over |z| <= 4.5 (normal law) + one cell per tail (law of the 2^23-point radius grid)"},
        "chi2_z_all_256_bins_vs_normal_law": {"chi2": chi_all, "df": df_all, "z_score": (chi_all - df_all) / math.sqrt(2 * df_all),
                                              "note": "includes the bins beyond 4.5 sigma, where the radius grid thins out (-0.4% at 4.5, -3.7% at 5, "
                                                      "nothing beyond 5.65 sigma): at 1e10 draws that deficit is resolved"},
        "hist_z_tail_bins": {f"{edges[i]:+.4f}": int(st["hist_z"][i]) for i in list(range(0, 32)) + list(range(224, 256))},
        "chi2_joint_64x64": {"chi2": chi_j, "df": df_j, "z_score": (chi_j - df_j) / math.sqrt(2 * df_j)},
        "tails": tails,
    }


def single_step_statistics(eng, seed=5, n_paths=1 << 32):
    """The 64-bit single-step draw (n_steps == 1, normal.cuh box_muller_single) at 2^32 draws: moments, body chi-square and
    tail counts against the NORMAL law - this draw has no thin tail (cap 6.66 sigma)."""
    st = eng.rng_statistics(seed, n_paths, 1)
    m = st["moments"]
    n = m[11]
    edges = np.linspace(-6.0, 6.0, 257)
    cdf = np.array([_phi(x) for x in edges])
    pz = np.diff(cdf)
    pz[0] += cdf[0]
    pz[-1] += 1.0 - cdf[-1]
    chi, df = chi_square(st["hist_z"], pz, n)
    sf = lambda c: 0.5 * math.erfc(c / math.sqrt(2.0))
    beyond = {f"|z|>{c}": {"count": int(sum(int(st["hist_z"][i]) for i in range(256) if edges[i] >= c or edges[i + 1] <= -c)),
                           "normal_law": 2 * n * sf(c)} for c in (5.25, 5.625)}
    return {"draws": n, "mean": m[0] / n, "second_moment": m[1] / n, "fourth_moment": m[3] / n,
            "chi2_z_all_256_bins_vs_normal_law": {"chi2": chi, "df": df, "z_score": (chi - df) / math.sqrt(2 * df)},
            "tails": {name: {"count": int(c), "normal_law": n * sf(x)} for name, c, x in
                      (("z>4", st["tails"][0], 4.0), ("z>5", st["tails"][1], 5.0), ("z<-4", st["tails"][2], 4.0), ("z<-5", st["tails"][3], 5.0))},
            "beyond_the_23_bit_cap": beyond}


def bs_price(S, K, T, r, sigma, call=True):
    d1 = (math.log(S / K) + (r + 0.5 * sigma * sigma) * T) / (sigma * math.sqrt(T))
    d2 = d1 - sigma * math.sqrt(T)
    if call:
        return S * _phi(d1) - K * math.exp(-r * T) * _phi(d2)
    return K * math.exp(-r * T) * _phi(-d2) - S * _phi(-d1)


def strike_sweep(eng, n_pairs=1 << 31, seed=7):
    """The reference's DEFAULT path - one exact step (monte_carlo.py:59) - priced at strikes S*exp(k sigma sqrt(T)), k = -5..5:
    out-of-the-money options whose value is a direct functional of ONE draw's tail.  2^31 antithetic pairs = 2^32 samples per
    strike, 14 strikes per launch on common random numbers."""
    from optionslab_b200 import _ffi

    S, T, r, sigma = 100.0, 1.0, 0.05, 0.2
    ks = [round(-5.0 + 0.25 * i, 2) for i in range(41)]
    rows = []
    for lo in range(0, len(ks), 14):
        blk = ks[lo:lo + 14]
        for is_put in (False, True):
            use = [k for k in blk if (k <= 0) == is_put or k == 0]  # OTM side: puts below the spot, calls above
            if not use:
                continue
            params = np.stack([_ffi.make_params(S, S * math.exp(k * sigma * math.sqrt(T)), T, r, sigma) for k in use]).reshape(1, len(use))
            m = eng.simulate(_ffi.make_spec(_ffi.EUROPEAN, 1, is_put=is_put, antithetic=True), params, seed, n_pairs, stream_base=lo)[0]
            disc = math.exp(-r * T)
            for k, rec in zip(use, m):
                n = rec["n"]
                mean = rec["sum"] / n
                se = disc * math.sqrt(max(rec["sum_sq"] / n - mean * mean, 0.0) / n)
                bs = bs_price(S, S * math.exp(k * sigma * math.sqrt(T)), T, r, sigma, call=not is_put)
                rows.append({"k_sigma": k, "type": "put" if is_put else "call", "price": disc * mean, "bs": bs, "std_error": se,
                             "z": (disc * mean - bs) / se if se > 0 else 0.0, "rel": disc * mean / bs - 1.0, "samples": n})
    return rows


def main():
    from optionslab_b200 import _ffi

    eng = _ffi.get_engine(0)
    out = {"stream_statistics_1.07e10_draws": stream_statistics(eng),
           "stream_statistics_5.4e8_draws": stream_statistics(eng, seed=77, n_paths=1 << 22, n_steps=128),
           "note_joint": "the 64x64 chi-square of the two normals of one word is asserted on the 5.4e8-draw run (2.7e8 pairs, 1/16 of the "
                         "2^32 points one 32-bit word can produce); the 1.07e10-draw run holds 5.4e9 pairs - more than there are distinct "
                         "points - so its joint chi-square measures the lattice of a 32-bit-per-pair generator, mostly in the corners beyond "
                         "r = 4.4 where one radius atom meets 512 equally spaced angles (arc spacing 0.11)",
           "single_step_draw_2^32_draws": single_step_statistics(eng),
           "strike_sweep_single_step_2^32_samples": strike_sweep(eng)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
