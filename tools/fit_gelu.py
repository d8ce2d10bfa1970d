"""Fits the polynomial used by gelu_erf() in csrc/fcn_conv.cu and reports its error against the exact-erf GELU.

    GELU(x) = x * Phi(x) ~= x / (1 + exp(-2 x (a + b s + c s^2))),  s = min(x^2, 36)

Iteratively re-weighted least squares towards the minimax absolute error on [-12, 12]; then the fp32 evaluation the
kernel performs (ex2 with the -2 log2(e) folded into the coefficients) is emulated and compared with float64 erf."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf

SMAX = 36.0


def main():
    x = np.linspace(-12, 12, 400001)
    g = 0.5 * x * (1 + erf(x / np.sqrt(2)))

    def err(p):
        s = np.minimum(x * x, SMAX)
        return x / (1 + np.exp(-2 * x * ((p[2] * s + p[1]) * s + p[0]))) - g

    p, w, best = np.array([0.7975, 0.0370, -0.00035]), np.ones_like(x), None
    for _ in range(60):
        p = least_squares(lambda q: err(q) * w, p, method="lm").x
        e = np.abs(err(p))
        if best is None or e.max() < best[0]:
            best = (e.max(), p.copy())
        w = w * (1 + 3 * e / e.max())
        w /= w.mean()
    k = -2 * np.log2(np.e)
    q = [np.float32(c * k) for c in best[1]]
    print("a, b, c =", list(best[1]), " max |err| (float64) =", best[0])
    print("ex2 coefficients:", [float(c) for c in q])
    xf = np.linspace(-12, 12, 2000001).astype(np.float32)
    s = np.minimum(xf * xf, np.float32(SMAX))
    t = (xf * ((q[2] * s + q[1]) * s + q[0])).astype(np.float32)
    y = (xf / (np.float32(1) + np.exp2(t).astype(np.float32))).astype(np.float32)
    gd = 0.5 * xf.astype(np.float64) * (1 + erf(xf.astype(np.float64) / np.sqrt(2)))
    print("fp32 evaluation: max |err| = %.3g at x = %.3f" % (np.abs(y - gd).max(), xf[np.argmax(np.abs(y - gd))]))


if __name__ == "__main__":
    main()
