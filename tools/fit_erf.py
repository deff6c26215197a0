"""Fit the FMA-pipe polynomial of the exact-erf variant (csrc/vrt_common.cuh: erf_exact).

    erf(x) ~= 1 - 2^(-x * P(x)),  x in [0, XMAX],  P of degree DEG   (x is clamped to XMAX on the device)

Minimises the maximum ABSOLUTE error of erf (the quantity that enters ln T additively) with an
iteratively re-weighted least-squares loop, then reports the error of an fp32 Horner evaluation.
Run:  python tools/fit_erf.py
"""
import numpy as np
from scipy.special import erf, erfc

XMAX, DEG = 4.0, 6


def model(c, x):
    p = np.zeros_like(x)
    for k in range(DEG, -1, -1):
        p = p * x + c[k]
    return 1.0 - np.exp2(-x * p)


def fit():
    x = np.linspace(1e-4, XMAX, 40001)
    target_g = -np.log2(erfc(x)) / x
    V = np.vander(x, DEG + 1, increasing=True)
    # d erf / d P = erfc(x) ln2 x  -> weight of the linearised problem
    sens = erfc(x) * np.log(2.0) * x
    w = np.ones_like(x)
    best = None
    for it in range(200):
        W = (sens * w)[:, None]
        c, *_ = np.linalg.lstsq(V * W, target_g * W[:, 0], rcond=None)
        err = model(c, x) - erf(x)
        m = np.abs(err).max()
        if best is None or m < best[0]:
            best = (m, c.copy())
        w *= (1.0 + 4.0 * np.abs(err) / m) / 3.0  # push weight towards the worst points
        w /= w.mean()
    return best


def eval_f32(c, x):
    c32 = [np.float32(v) for v in c]
    x = x.astype(np.float32)
    p = np.full_like(x, c32[DEG])
    for k in range(DEG - 1, -1, -1):
        p = p * x + c32[k]
    return np.float32(1.0) - np.exp2(-(p * x)).astype(np.float32)


if __name__ == "__main__":
    m, c = fit()
    print("max abs err (double eval): %.3e" % m)
    xs = np.linspace(0, XMAX, 400001)
    e32 = np.abs(eval_f32(c, xs).astype(np.float64) - erf(xs)).max()
    print("max abs err (fp32 Horner): %.3e" % e32)
    print("value at XMAX: %.9f (1 - %.2e)" % (model(c, np.array([XMAX]))[0], 1 - model(c, np.array([XMAX]))[0]))
    print("constexpr float " + ", ".join("EX_C%d = %.10ef" % (k, v) for k, v in enumerate(c)) + ";")
