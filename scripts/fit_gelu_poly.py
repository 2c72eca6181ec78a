"""Coefficients of the odd polynomials the fused GELU epilogues use (bf_gemm_act.cu: gelu_poly2, gelu_grad_poly2):
least-maximum fits of  Phi(z) - 1/2 = z Q(z^2)  and  gelu'(z) - 1/2 = z R(z^2)  on |z| <= 4 (Lawson-reweighted least
squares in a Chebyshev basis of t = z^2), checked in float32 Horner arithmetic with the clamp over [-8, 8]."""
from math import erf, pi, sqrt

import numpy as np
import numpy.polynomial.chebyshev as C

verf = np.vectorize(erf)
f_cdf = lambda z: 0.5 * verf(z / sqrt(2))
f_grad = lambda z: 0.5 * verf(z / sqrt(2)) + z * np.exp(-z * z / 2) / sqrt(2 * pi)


def fit(func, L, deg_t, n=8001, iters=300):
    z = np.linspace(1e-6, L, n)
    fz = func(z)
    V = C.chebvander(2 * z * z / (L * L) - 1, deg_t)
    w, best = np.ones_like(z), None
    for _ in range(iters):
        c, *_ = np.linalg.lstsq(V * (z * w)[:, None], fz * w, rcond=None)
        err = np.abs((V @ c) * z - fz)
        if best is None or err.max() < best[1]:
            best = (c.copy(), err.max())
        w = w * (err / err.max() + 0.05)
        w /= w.mean()
    poly = C.cheb2poly(best[0])
    return np.polynomial.Polynomial(poly)(np.polynomial.Polynomial([-1, 2 / (L * L)])).coef


def check(coef, func, L):
    zz = np.linspace(-8, 8, 800001).astype(np.float32)
    zc = np.clip(zz, -L, L)
    tt = zc * zc
    acc = np.full_like(tt, np.float32(coef[-1]))
    for k in range(len(coef) - 2, -1, -1):
        acc = acc * tt + np.float32(coef[k])
    got = (np.float32(0.5) + zc * acc).astype(np.float64)
    want = 0.5 + func(zz.astype(np.float64))
    inside = np.abs(zz) <= L
    return np.abs(got - want)[inside].max(), np.abs(got - want).max()


for name, func, deg in (("Phi(z) - 1/2", f_cdf, 7), ("gelu'(z) - 1/2", f_grad, 8)):
    coef = fit(func, 4.0, deg)
    e_in, e_all = check(coef, func, 4.0)
    print(f"{name}: degree {2 * deg + 1}, max abs err {e_in:.2e} on |z| <= 4, {e_all:.2e} on |z| <= 8 (clamped)")
    print("   ", ", ".join(f"{x:.10e}f" for x in coef))
