"""Small driver for ncu captures of single kernels (GEMM fc1/qkv/proj shapes of ViTPose-B at 128 crops)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitpose_b200 import _lib, ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else 'fc1'
dev = torch.device('cuda:0')
M, D = 128 * 2 * 192, 768
BF16 = torch.bfloat16
shapes = {'fc1': (4 * D, D, _lib.EPI_GELU_BF16), 'qkv': (3 * D, D, _lib.EPI_BIAS_BF16),
          'proj': (D, D, _lib.EPI_RESID_F32), 'fc2': (D, 4 * D, _lib.EPI_RESID_F32)}
if which == 'attn':
    qkv = torch.randn(256, 192, 3 * D, device=dev).to(BF16)
    for _ in range(3):
        ops.attention(qkv, 12)
elif which == 'attn_bwd':
    qkv = torch.randn(64, 192, 3 * D, device=dev).to(BF16)
    o, lse_ = ops.attention_with_lse(qkv, 12)
    do = torch.randn(64, 192, D, device=dev).to(BF16)
    for _ in range(3):
        ops.attention_bwd(qkv, o, lse_, do, 12)
elif which == 'wgrad':
    dy = torch.randn(64 * 192, 4 * D, device=dev).to(BF16)
    xx = torch.randn(64 * 192, D, device=dev).to(BF16)
    dw = torch.zeros(4 * D, D, device=dev)
    for _ in range(3):
        ops.gemm_atb_accum(dy, xx, dw)
elif which == 'decode':
    hm = torch.rand(512, 17, 64, 48, device=dev)
    fi = torch.arange(17, device=dev, dtype=torch.int32)
    c, s_ = torch.rand(256, 2, device=dev), torch.rand(256, 2, device=dev)
    for _ in range(3):
        ops.decode(hm[:256], hm[256:], fi, False, _lib.DECODE_UDP_DARK, 11, True, c, s_)
elif which in ('proj_ln', 'fc2_ln'):
    K = D if which == 'proj_ln' else 4 * D
    A = torch.randn(M, K, device=dev).to(BF16)
    B = (torch.randn(D, K, device=dev) / math.sqrt(K)).to(BF16)
    bias, gm, bt = torch.randn(D, device=dev), torch.ones(D, device=dev), torch.zeros(D, device=dev)
    out = torch.randn(M, D, device=dev)
    for _ in range(3):
        ops.gemm_layernorm(A, B, _lib.EPI_RESID_F32, bias, out, gm, bt, out=out)
else:
    N, K, epi = shapes[which]
    A = torch.randn(M, K, device=dev).to(BF16)
    B = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(BF16)
    bias = torch.randn(N, device=dev)
    out = torch.randn(M, N, device=dev) if epi == _lib.EPI_RESID_F32 else torch.empty(M, N, device=dev, dtype=BF16)
    for _ in range(3):
        ops.gemm(A, B, epi, bias=bias, out=out, aux=out if epi == _lib.EPI_RESID_F32 else None)
torch.cuda.synchronize()
print('done', which)
