"""Attention backward alone: us per launch (CUDA events), optional per-phase cycle stamps of CTA 0 (VPB_ATTBWD_DEBUG=1).
   python tools/attbwd_time.py [crops=64] [heads=12] [head_dim=64]"""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitpose_b200 import ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
heads = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hd = int(sys.argv[3]) if len(sys.argv) > 3 else 64
T, D = 192, heads * hd
g = torch.Generator(device='cuda').manual_seed(0)
sets = []
for _ in range(3):
    qkv = torch.randn(n, T, 3 * D, device='cuda', generator=g).bfloat16()
    dout = torch.randn(n, T, D, device='cuda', generator=g).bfloat16()
    out, lse = ops.attention_with_lse(qkv, heads)
    sets.append((qkv, out, lse, dout))
dbias = torch.zeros(3 * D, device='cuda')
for i in range(6):
    ops.attention_bwd(*sets[i % 3], heads, dbias=dbias)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for with_bias in (True, False):
    e0.record()
    for i in range(30):
        ops.attention_bwd(*sets[i % 3], heads, dbias=dbias if with_bias else None)
    e1.record()
    torch.cuda.synchronize()
    print(f'attention_bwd n={n} heads={heads} hd={hd} dbias={with_bias}: {e0.elapsed_time(e1) / 30 * 1e3:.1f} us per launch')
