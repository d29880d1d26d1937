"""Fits the polynomial of the GELU epilogue (vitpose_b200/csrc/gemm.cuh: gelu_erf2):
    gelu(x) = max(x, 0) - 0.5 |x| erfc(|x| / sqrt 2),   erfc(a / sqrt 2) ~= exp2(q(a)),  q(a) = c1 a + ... + c5 a^5
minimising the maximum ABSOLUTE error of gelu over a in [0, 9] (Lawson-weighted least squares), then checks the fp32
evaluation over every finite bf16 input. CPU only (numpy / scipy)."""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import erf, erfc

DEG, A = 5, 9.0


def model(c, a):
    q = np.zeros_like(a)
    for ci in c[::-1]:
        q = (q + ci) * a
    return np.exp2(q)


def fit():
    a = np.concatenate([np.linspace(0, 2, 4001), np.linspace(2, A, 6001)])
    target = erfc(a / np.sqrt(2))
    mask = target > 1e-12
    V = np.vander(a[mask], DEG + 1, increasing=True)[:, 1:]
    w = np.sqrt(target[mask])
    c, *_ = np.linalg.lstsq(V * w[:, None], np.log2(target[mask]) * w, rcond=None)
    lw = np.ones_like(a)
    for _ in range(40):
        res = least_squares(lambda c: (model(c, a) - target) * lw * np.maximum(0.5 * a, 0.05), c, method='lm',
                            xtol=1e-15, ftol=1e-15)
        c = res.x
        err = np.abs((model(c, a) - target) * np.maximum(0.5 * a, 0.05))
        lw = lw * (0.5 + err / err.max())
        lw /= lw.mean()
    return c


def gelu_f32(x, c):
    """fp32 emulation of the device code (fma rounding aside)."""
    x = x.astype(np.float32)
    a = np.abs(x)
    c = [np.float32(v) for v in c]
    q = c[4] * np.ones_like(a)
    for ci in c[3::-1]:
        q = (q * a + ci).astype(np.float32)
    q = (q * a).astype(np.float32)
    e = np.exp2(q).astype(np.float32)
    t = ((a * np.float32(-0.5)).astype(np.float32) * e).astype(np.float32)
    return ((x + a).astype(np.float32) * np.float32(0.5) + t).astype(np.float32)


if __name__ == '__main__':
    c = fit()
    print('coefficients c1..c5:', [repr(float(np.float32(v))) for v in c])
    bits = np.arange(65536, dtype=np.uint32) << 16
    x = bits.view(np.float32)
    x = x[np.isfinite(x) & (np.abs(x) < 1e30)]
    xd = x.astype(np.float64)
    ref = 0.5 * xd * (1 + erf(xd / np.sqrt(2)))
    err = np.abs(gelu_f32(x, c) - ref)
    print(f'max |gelu - exact| over all bf16 inputs: {err.max():.3e} at x = {x[err.argmax()]}')
