"""Prints where the persistent attention kernel's roles wait (cycle counters of CTA 0). Run with VPB_ATT_DEBUG=96."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitpose_b200 import ops
qkv = torch.randn(256, 192, 2304, device='cuda').to(torch.bfloat16)
for _ in range(4):
    ops.attention(qkv, 12)
torch.cuda.synchronize()
