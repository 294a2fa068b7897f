#!/usr/bin/env python
"""Derive the FP32 polynomial of the device inverse normal CDF (csrc/sobol.cuh, inverse_normal_from_sobol).

Form: with t = min(u, 1-u) in [1e-10, 1/2] and y = sqrt(-2 ln t) in [1.1774, 6.7861],
    Phi^-1(u) = sign(u - 1/2) * P((y - c) / h),      P of degree 14,
ONE branch-free polynomial over the whole range (2 MUFU: lg2, sqrt; 15 FFMA).  A piecewise form after M. Giles
("Approximating the erfinv function", GPU Computing Gems 2011: polynomial in w = -ln(4t(1-t)) for w < 5, in sqrt(w)
beyond) is 6e-7 relative but its tail branch diverges in 19% of the warps; on B200 the branch-free form is 25% faster
(profiles/r01_variants13_inverse_normal.txt).  The coefficients are a Chebyshev-node least-squares fit computed here
in FP64 against scipy.special.ndtri; the script evaluates the FP32 Horner form and prints its worst error.
Run:  python tools/fit_inverse_normal.py   (prints the C initialisers)."""
import numpy as np
from scipy import special

T_MIN, DEG = 1e-10, 14


def main():
    ymin, ymax = np.sqrt(-2 * np.log(0.5)), np.sqrt(-2 * np.log(T_MIN))
    c, h = 0.5 * (ymin + ymax), 0.5 * (ymax - ymin)
    n = 8000
    y = c + h * np.cos(np.pi * (np.arange(n) + 0.5) / n)
    z = -special.ndtri(np.exp(-y * y / 2))
    cheb, *_ = np.linalg.lstsq(np.polynomial.chebyshev.chebvander((y - c) / h, DEG), z, rcond=None)
    power = np.polynomial.chebyshev.cheb2poly(cheb)[::-1]  # highest degree first (Horner order)
    print(f"// y -> v = y * {1 / h:.9e}f + {-c / h:.9e}f")
    print("// degree %d in v, highest degree first" % DEG)
    print("{" + ", ".join(f"{a:.9e}f" for a in power) + "}")
    # FP32 evaluation error over a dense grid of t
    t = np.concatenate([np.logspace(np.log10(T_MIN), np.log10(0.5), 400001), np.linspace(1e-3, 0.5, 400001)]).astype(np.float32)
    yy = np.sqrt((np.log2(t).astype(np.float32) * np.float32(-2 * np.log(2))).astype(np.float32)).astype(np.float32)
    v = (yy * np.float32(1 / h) + np.float32(-c / h)).astype(np.float32)
    p = np.full_like(v, np.float32(power[0]))
    for a in power[1:]:
        p = (p * v + np.float32(a)).astype(np.float32)
    exact = -special.ndtri(t.astype(np.float64))
    err = np.abs(p.astype(np.float64) - exact)
    print(f"// FP32 evaluation: max abs error {err.max():.3e} (at t = {t[err.argmax()]:.3e})")


if __name__ == "__main__":
    main()
