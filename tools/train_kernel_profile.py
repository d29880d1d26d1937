"""Kernel-time table of the training step (torch.profiler / CUPTI, warm caches, kernels running back to back):
   python tools/train_kernel_profile.py [crops=64] [steps=5]   -> per-kernel us / step, GPU busy time vs step time"""
import collections
import os
import re
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.argv = [sys.argv[0]] + sys.argv[1:]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.argv = [sys.argv[0], str(n)]
import train_cpu_probe as probe  # noqa: E402  (builds the model, runs warm-up steps)

torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        probe.step()
    e1.record()
    torch.cuda.synchronize()
span = e0.elapsed_time(e1) / steps
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type is not None and 'cuda' in str(ev.device_type).lower():
        name = re.sub(r'\(.*', '', ev.name)[:80]
        agg[name][0] += 1
        agg[name][1] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
busy = sum(v[1] for v in agg.values()) / steps
print(f'crops {n}: step {span * 1e3:.0f} us (under the profiler), kernels busy {busy:.0f} us / step, '
      f'{sum(v[0] for v in agg.values()) / steps:.0f} launches / step  [VPB_TRAIN_FUSE={os.environ.get("VPB_TRAIN_FUSE", "1")}]')
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f'{t / steps:9.1f} us  {c / steps:6.1f} x  {name}')
