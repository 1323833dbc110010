"""Coefficients of the GELU used by the fc1 epilogue (csrc/gemm_tc.cu: gelu_fast2).

    GELU(x) = max(x, 0) - 0.5 * t * erfc(t / sqrt 2),  t = min(|x|, T)
    erfc(t / sqrt 2) = 2 ** (t * P(t)),                P = degree-6 polynomial (weighted minimax fit, Lawson iteration)

One ex2 per element and no reciprocal (the Abramowitz-Stegun 7.1.26 form it replaces needs exp AND 1/(1+pz)); the fit is
weighted with the factor that turns an error of the exponent into an error of GELU, 0.5 t erfc(t/sqrt 2) ln 2.  Prints the
fp32 coefficients and the maximum absolute error of the fp32 evaluation against the fp64 erf GELU."""
import numpy as np
from scipy import special

T, DEG = 5.7, 7
t = np.linspace(0, T, 40001)[1:]
q = np.log2(special.erfc(t / np.sqrt(2))) / t
W = 0.5 * t * special.erfc(t / np.sqrt(2)) * np.log(2) * t
A = np.vander(t, DEG, increasing=True)
lw = np.ones_like(W)
for _ in range(200):
    coef, *_ = np.linalg.lstsq(A * (W * lw)[:, None], q * W * lw, rcond=None)
    err = np.abs((A @ coef - q) * W)
    lw *= 0.5 + err / err.max()
    lw /= lw.mean()
c32 = coef.astype(np.float32)
print("coefficients (t^0 .. t^%d):" % (DEG - 1))
for c in c32:
    print("    %.9ef," % c)


def gelu_f32(x):
    x = x.astype(np.float32)
    tc = np.minimum(np.abs(x), np.float32(T))
    p = np.full_like(tc, c32[-1])
    for c in c32[-2::-1]:
        p = p * tc + c
    e = np.exp2(p * tc)
    return np.maximum(x, np.float32(0)) + (np.float32(-0.5) * tc) * e


xs = np.concatenate([np.linspace(-12, 12, 2000001), np.array([-1e4, 1e4, 0.0, -0.0])])
ref = 0.5 * xs * (1 + special.erf(xs / np.sqrt(2)))
d = np.abs(gelu_f32(xs) - ref)
print("max |err| %.3e at x = %.4f; max |err| for |x| < 3: %.3e" % (d.max(), xs[d.argmax()], d[np.abs(xs) < 3].max()))
