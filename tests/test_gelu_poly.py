"""The polynomial of the GELU epilogue (vitpose_b200/csrc/gemm.cuh: gelu_erf2), evaluated in fp32 on the CPU with the
constants read from the CUDA source, against the exact erf form (torch.nn.GELU default, mmcv/vit.py Mlp act_layer)."""
import os
import re
import sys

import numpy as np
from scipy.special import erf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tools'))


def _constants():
    src = open(os.path.join(ROOT, 'vitpose_b200', 'csrc', 'gemm.cuh')).read()
    vals = [float(re.search(r'GELU_Q%d = (-?[0-9.e-]+)f' % i, src).group(1)) for i in range(1, 6)]
    assert len(vals) == 5
    return vals


def test_gelu_polynomial_is_float_exact_over_all_bf16_inputs():
    from fit_gelu import gelu_f32
    c = _constants()
    bits = np.arange(65536, dtype=np.uint32) << 16
    x = bits.view(np.float32)
    x = x[np.isfinite(x) & (np.abs(x) < 1e30)]
    xd = x.astype(np.float64)
    ref = 0.5 * xd * (1 + erf(xd / np.sqrt(2)))
    with np.errstate(over='ignore'):
        got = gelu_f32(x, c)
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() < 1e-6
    big = np.abs(x) > 16
    assert (got[big] == np.maximum(x[big], 0)).all()      # erfc underflows to exactly 0


def test_gelu_polynomial_is_decreasing():
    c = _constants()
    a = np.linspace(0, 1000, 2000001)
    dq = c[0] + 2 * c[1] * a + 3 * c[2] * a ** 2 + 4 * c[3] * a ** 3 + 5 * c[4] * a ** 4
    assert dq.max() < 0
