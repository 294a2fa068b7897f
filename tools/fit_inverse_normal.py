#!/usr/bin/env python
"""Derive the FP32 polynomial coefficients of the device inverse normal CDF (csrc/sobol.cuh).

Form (after M. Giles, "Approximating the erfinv function", GPU Computing Gems 2011): with t = min(u, 1-u),
w = -ln(4 t (1-t)) = -ln(1 - x^2) for x = 1 - 2t, the ratio  sqrt(2) erfinv(x) / x  is a smooth function of w:
    central branch  w <  5 :  polynomial of degree 8 in (w - 2.5)
    tail branch     w >= 5 :  polynomial of degree 8 in (sqrt(w) - 3)      (w <= 21.7 for t >= 1e-10)
so that  Phi^-1(u) = sign(u - 1/2) * x * poly.  The coefficients below are a Chebyshev-node least-squares fit
computed here in FP64 against scipy.special.erfinv / ndtri; the script then evaluates the FP32 Horner form and
prints its worst error.  Run:  python tools/fit_inverse_normal.py  (prints the C initialisers)."""
import numpy as np
from scipy import special


def ratio_of_w(w):
    """sqrt(2)*erfinv(x)/x as a function of w = -ln(1-x^2), stable for small and large w."""
    w = np.asarray(w, dtype=np.float64)
    one_minus_x2 = np.exp(-w)
    x = np.sqrt(-np.expm1(-w))
    # ndtri((1+x)/2) loses digits as x -> 1; use t = (1-x)/2 = (1-x^2)/(2(1+x)) and ndtri(t) = -z
    t = one_minus_x2 / (2.0 * (1.0 + x))
    z = -special.ndtri(t)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = z / x
    small = w < 1e-8
    r[small] = np.sqrt(np.pi / 2.0)  # limit: erfinv(x) ~ sqrt(pi)/2 x
    return r


def cheb_fit(fun, lo, hi, centre, deg, n=4000):
    k = np.arange(n)
    nodes = 0.5 * (lo + hi) + 0.5 * (hi - lo) * np.cos(np.pi * (k + 0.5) / n)
    A = np.vander(nodes - centre, deg + 1, increasing=False)
    y = fun(nodes)
    # minimise the RELATIVE error
    coef, *_ = np.linalg.lstsq(A / y[:, None], np.ones_like(y), rcond=None)
    return coef  # highest degree first (Horner order)


def horner32(coef, v):
    c = coef.astype(np.float32)
    p = np.full_like(v, c[0], dtype=np.float32)
    for a in c[1:]:
        p = (p * v + a).astype(np.float32)
    return p


def main():
    central = cheb_fit(ratio_of_w, 0.0, 5.0, 2.5, 8)
    tail = cheb_fit(lambda s: ratio_of_w(s * s), np.sqrt(5.0), 4.7, 3.0, 8)
    print("// central: degree 8 in (w - 2.5), highest degree first")
    print("{" + ", ".join(f"{c:.9e}f" for c in central) + "}")
    print("// tail: degree 8 in (sqrt(w) - 3), highest degree first")
    print("{" + ", ".join(f"{c:.9e}f" for c in tail) + "}")
    # FP32 evaluation error over a dense grid of t in [1e-10, 0.5]
    t = np.concatenate([np.logspace(-10, np.log10(0.5), 400001), np.linspace(1e-3, 0.5, 400001)])
    t32 = t.astype(np.float32)
    w = (-np.log((4.0 * t32 * (1.0 - t32)).astype(np.float32))).astype(np.float32)
    x = (1.0 - 2.0 * t32).astype(np.float32)
    p = np.where(w < 5.0, horner32(central, (w - np.float32(2.5)).astype(np.float32)),
                 horner32(tail, (np.sqrt(w) - np.float32(3.0)).astype(np.float32)))
    z = (p * x).astype(np.float64)
    exact = -special.ndtri(t32.astype(np.float64))
    err = np.abs(z - exact)
    print(f"// FP32 evaluation: max abs error {err.max():.3e} (at t={t32[err.argmax()]:.3e}), "
          f"max rel error for |z|>0.1 {np.max(err[exact > 0.1] / exact[exact > 0.1]):.3e}")


if __name__ == "__main__":
    main()
