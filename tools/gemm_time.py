"""Times the residual GEMM + fused LayerNorm kernels alone (ViTPose-B shapes, L2-cold: operands rotate over
buffer sets larger than L2) — used for A/B runs with VPB_GEMM_FLAGS / VPB_GEMM_CG / VPB_LN_FUSED.

  python tools/gemm_time.py [crops=256] [model=base]
"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitpose_b200 import _lib, ops  # noqa: E402

crops = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = sys.argv[2] if len(sys.argv) > 2 else 'base'
D = {'small': 384, 'base': 768, 'large': 1024, 'huge': 1280}[model]
M = crops * 2 * 192
dev = torch.device('cuda:0')
BF16 = torch.bfloat16
NBUF = 3


def timeit(fn, iters=12, warm=3):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


tag = ' '.join(f'{k}={os.environ[k]}' for k in ('VPB_GEMM_FLAGS', 'VPB_GEMM_CG', 'VPB_LN_FUSED') if k in os.environ)
for name, K in (('proj_ln', D), ('fc2_ln', 4 * D)):
    As = [torch.randn(M, K, device=dev).to(BF16) for _ in range(NBUF)]
    B = (torch.randn(D, K, device=dev) / math.sqrt(K)).to(BF16)
    bias, gm, bt = torch.randn(D, device=dev), torch.ones(D, device=dev), torch.zeros(D, device=dev)
    xs = [torch.randn(M, D, device=dev) for _ in range(NBUF)]
    ms = timeit(lambda i: ops.gemm_layernorm(As[i % NBUF], B, _lib.EPI_RESID_F32, bias, xs[i % NBUF], gm, bt,
                                             out=xs[i % NBUF]))
    byts = M * K * 2 + M * D * (4 + 4 + 2)
    print(f'{name} M={M} N={D} K={K}: {ms * 1e3:.1f} us  {2.0 * M * D * K / ms / 1e9:.0f} TFLOP/s  '
          f'{byts / ms / 1e6:.0f} GB/s  [{tag}]')
    del As, xs
for name, N, K, epi in (('qkv', 3 * D, D, _lib.EPI_BIAS_BF16), ('fc1', 4 * D, D, _lib.EPI_GELU_BF16)):
    As = [torch.randn(M, K, device=dev).to(BF16) for _ in range(NBUF)]
    B = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(BF16)
    bias = torch.randn(N, device=dev)
    outs = [torch.empty(M, N, device=dev, dtype=BF16) for _ in range(2)]
    ms = timeit(lambda i: ops.gemm(As[i % NBUF], B, epi, bias=bias, out=outs[i % 2]))
    print(f'{name} M={M} N={N} K={K}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s  [{tag}]')
    del As, outs

# folded LayerNorm: producers (plain bf16 rows + per-tile statistics) and consumers (normalise in the epilogue)
for name, K in (('proj_fold', D), ('fc2_fold', 4 * D)):
    As = [torch.randn(M, K, device=dev).to(BF16) for _ in range(NBUF)]
    B = (torch.randn(D, K, device=dev) / math.sqrt(K)).to(BF16)
    bias = torch.randn(D, device=dev)
    xs = [torch.randn(M, D, device=dev) for _ in range(NBUF)]
    xb = torch.empty(M, D, device=dev, dtype=BF16)
    _, _, stats = ops.gemm_resid_stats(As[0], B, _lib.EPI_RESID_F32, bias, xs[0], out=xs[0], xb=xb)
    ms = timeit(lambda i: ops.gemm_resid_stats(As[i % NBUF], B, _lib.EPI_RESID_F32, bias, xs[i % NBUF], out=xs[i % NBUF],
                                               xb=xb, stats=stats))
    byts = M * K * 2 + M * D * (4 + 4 + 2)
    print(f'{name} M={M} N={D} K={K}: {ms * 1e3:.1f} us  {2.0 * M * D * K / ms / 1e9:.0f} TFLOP/s  '
          f'{byts / ms / 1e6:.0f} GB/s  [{tag}]')
    del As, xs
for name, N, K, epi in (('qkv_fold', 3 * D, D, _lib.EPI_BIAS_BF16), ('fc1_fold', 4 * D, D, _lib.EPI_GELU_BF16)):
    As = [torch.randn(M, K, device=dev).to(BF16) for _ in range(NBUF)]
    wf, s, c = ops.fold_layernorm_linear(torch.randn(N, K, device=dev) / math.sqrt(K), torch.randn(N, device=dev),
                                         torch.ones(K, device=dev), torch.zeros(K, device=dev))
    outs = [torch.empty(M, N, device=dev, dtype=BF16) for _ in range(2)]
    ms = timeit(lambda i: ops.gemm_lnfold(As[i % NBUF], wf, s, c, stats, epilogue=epi, out=outs[i % 2]))
    print(f'{name} M={M} N={N} K={K}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s  [{tag}]')
    del As, outs
