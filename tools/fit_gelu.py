#!/usr/bin/env python
"""Fit and validate the coefficients of fd::gelu_erf_fast (csrc/common.cuh).

gelu(x) = (h + |h|) - |h| * erfc(|x|/sqrt2), h = x/2, erfc(|x|/sqrt2) = exp2(a * q(a)), a = min(|x|, 4 sqrt2):
q is a weighted minimax-ish (Lawson iteration) polynomial fit of -log2(erfc(a/sqrt2)) / a, weight = the
sensitivity of erf to q.  Prints the coefficients and the error of an fp32 emulation against fp64.
"""
import numpy as np
from numpy.polynomial import chebyshev as C
from scipy.special import erf, erfc

ZMAX, D = 4.0, 9
n = 6000
zk = 0.5 * ZMAX * (1 - np.cos(np.pi * (np.arange(n) + 0.5) / n))
y = -np.log(erfc(zk)) / zk
w = erfc(zk) * zk + 1e-12
V = C.chebvander(2 * zk / ZMAX - 1, D - 1)
coef = np.linalg.lstsq(V * w[:, None], y * w, rcond=None)[0]
for _ in range(60):
    r = (V @ coef - y) * erfc(zk) * zk
    w = w * (1 + np.abs(r) / np.abs(r).max())
    coef = np.linalg.lstsq(V * w[:, None], y * w, rcond=None)[0]
base, poly = np.poly1d([2 / ZMAX, -1]), np.poly1d([0.0])
for k, c in enumerate(C.cheb2poly(coef)):
    poly = poly + c * (base ** k)
q = poly.coeffs[::-1]
q2 = np.array([-np.log2(np.e) * q[k] / np.sqrt(2) ** (k + 1) for k in range(len(q))]).astype(np.float32)
print("q (ascending):", ", ".join("%.9ef" % c for c in q2))

x = np.linspace(-9, 9, 1800001).astype(np.float32)
a = np.minimum(np.abs(x), np.float32(ZMAX * np.sqrt(2)))
acc = np.full_like(a, q2[-1])
for c in q2[-2::-1]:
    acc = np.float32(np.float64(acc) * np.float64(a) + np.float64(c))
e = np.exp2((acc * a).astype(np.float32).astype(np.float64)).astype(np.float32)
h = (np.float32(0.5) * x).astype(np.float32)
g = np.float32(np.float64(-np.abs(h)) * np.float64(e) + np.float64((h + np.abs(h)).astype(np.float32)))
ref = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
print("max abs error vs fp64: %.3e" % np.abs(g - ref).max())
