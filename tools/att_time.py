"""Times ops.attention alone (ViTPose-B shape, 256 image passes) — used with VPB_ATT_DEBUG A/B flags."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitpose_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
hd = int(sys.argv[2]) if len(sys.argv) > 2 else 64
heads = 12 if hd == 64 else 16
qkvs = [torch.randn(n, 192, 3 * heads * hd, device='cuda').to(torch.bfloat16) for _ in range(3)]
for q in qkvs:
    ops.attention(q, heads)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(30):
    ops.attention(qkvs[i % 3], heads)
b.record()
torch.cuda.synchronize()
print(f'attention n={n} hd={hd}: {a.elapsed_time(b) / 30 * 1e3:.1f} us  (VPB_ATT_DEBUG={os.environ.get("VPB_ATT_DEBUG", "0")})')
