"""Fit and verify the exponent polynomial of gelu_fast_f (lavie_b200/csrc/common.cuh).

Phi(-t) = erfc(t/sqrt2)/2 = 2^-q(t), q(t) = 1 + t*Q(t); Q is fitted by reweighted least squares on the error of
2^-q, then the fp32 Horner evaluation is checked against the erf GELU on a dense grid."""
import numpy as np
from scipy.special import erf, erfc

DEG = 4
t = np.linspace(1e-6, 6.0, 60001)
e = 0.5 * erfc(t / np.sqrt(2))
y = (-np.log2(e) - 1) / t
w = e * t * np.log(2)
V = np.vander(t, DEG + 1, increasing=True)
c = np.linalg.lstsq(V * w[:, None], y * w, rcond=None)[0]
for _ in range(60):
    err = 2.0 ** (-(1 + t * (V @ c))) - e
    w2 = w * (1 + 3 * (np.abs(err) / np.abs(err).max()) ** 2)
    c = np.linalg.lstsq(V * w2[:, None], y * w2, rcond=None)[0]
c32 = c.astype(np.float32)
print("coefficients (t^0..t^4 of Q):", [float(x) for x in c32])

g = np.linspace(-40, 40, 800001).astype(np.float32)
tt = np.abs(g)
q = np.zeros_like(tt) + c32[DEG]
for k in range(DEG - 1, -1, -1):
    q = q * tt + c32[k]
q = q * tt + np.float32(1)
fast = np.maximum(g, 0).astype(np.float64) - tt.astype(np.float64) * np.exp2(-q.astype(np.float64))
ref = 0.5 * g.astype(np.float64) * (1 + erf(g.astype(np.float64) / np.sqrt(2)))
print("max |gelu_fast - gelu_erf| on [-40, 40]:", np.abs(fast - ref).max())
