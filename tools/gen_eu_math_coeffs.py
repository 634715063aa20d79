#!/usr/bin/env python3
"""Derive the polynomial coefficients used in include/eu_math.h.

Weighted least squares on Chebyshev nodes in float64 (near-minimax, far below float32
resolution), coefficients then rounded to float32 and printed as hex-float literals.
Run: python tools/gen_eu_math_coeffs.py
"""
import numpy as np


def cheb_nodes(a, b, n):
    k = np.arange(n)
    return 0.5 * (a + b) + 0.5 * (b - a) * np.cos(np.pi * (2 * k + 1) / (2 * n))


def fit(g, w, a, b, ncoef, n=4000):
    u = cheb_nodes(a, b, n)
    A = np.vander(u, ncoef, increasing=True) * w(u)[:, None]
    y = g(u) * w(u)
    c, *_ = np.linalg.lstsq(A, y, rcond=None)
    return c


def hexf(c):
    return float(np.float32(c)).hex() + "f"


def main():
    q = (np.pi / 4) ** 2 * 1.02
    eps = 1e-12

    # sin(r) = r + r^3 * S(u), u = r^2
    def g_sin(u):
        r = np.sqrt(u)
        # series near 0 to avoid cancellation
        return np.where(u < 1e-4, -1 / 6 + u / 120 - u * u / 5040, (np.sin(r) - r) / np.maximum(r, eps) ** 3)
    S = fit(g_sin, lambda u: u, 0.0, q, 4)

    # cos(r) = 1 - u/2 + u^2 * C(u)
    def g_cos(u):
        r = np.sqrt(u)
        return np.where(u < 1e-3, 1 / 24 - u / 720 + u * u / 40320, (np.cos(r) - 1 + u / 2) / np.maximum(u, eps) ** 2)
    C = fit(g_cos, lambda u: u * u, 0.0, q, 4)

    # atan(t) = t + t^3 * A(u), u = t^2 in [0,1]
    def g_atan(u):
        t = np.sqrt(u)
        return np.where(u < 1e-4, -1 / 3 + u / 5 - u * u / 7, (np.arctan(t) - t) / np.maximum(t, eps) ** 3)
    A = fit(g_atan, lambda u: u, 0.0, 1.0, 9)

    for name, c in (("S", S), ("C", C), ("A", A)):
        print(name, ", ".join(hexf(x) for x in c))
        print("   ", ", ".join(repr(float(np.float32(x))) for x in c))

    # pi/2 split in three float32 parts, 2/pi, pi, pi/2 hi/lo for atan
    import mpmath
    mpmath.mp.prec = 200
    pio2 = mpmath.pi / 2
    p1 = np.float32(float(pio2))
    p2 = np.float32(float(pio2 - mpmath.mpf(float(p1))))
    p3 = np.float32(float(pio2 - mpmath.mpf(float(p1)) - mpmath.mpf(float(p2))))
    print("PIO2_1..3", hexf(p1), hexf(p2), hexf(p3))
    print("TWO_OVER_PI", hexf(float(2 / mpmath.pi)))
    pi1 = np.float32(float(mpmath.pi))
    pi2 = np.float32(float(mpmath.pi - mpmath.mpf(float(pi1))))
    print("PI hi/lo", hexf(pi1), hexf(pi2))


if __name__ == "__main__":
    main()
