"""Is the training step bound by the GPU or by the Python thread that enqueues it? Per step: host time until the last
launch has been enqueued (no synchronisation), and device time between CUDA events.   python tools/train_cpu_probe.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitpose_b200 as V  # noqa: E402
from vitpose_b200 import configs, synthetic  # noqa: E402
from vitpose_b200.optim import LayerDecayOptimizerConstructor  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = configs.baseline_model_cfg('B-classic-17')
K = cfg['keypoint_head']['out_channels']
model = V.build_posenet(cfg)
model.load_state_dict(synthetic.scaled_init_state_dict(cfg, 0), strict=True)
model = model.cuda().train()
opt = LayerDecayOptimizerConstructor(dict(type='AdamW', lr=5e-4, betas=(0.9, 0.999), weight_decay=0.1),
                                     dict(num_layers=12, layer_decay_rate=0.75))(model)
img = synthetic.synthetic_crops(n, seed=1).cuda()
tgt = torch.rand(n, K, 64, 48).cuda()
tw = torch.ones(n, K, 1).cuda()


def step():
    out = model.train_step(dict(img=img, target=tgt, target_weight=tw, img_metas=None), opt)
    opt.zero_grad(set_to_none=True)
    out['loss'].backward()
    opt.step(max_norm=1.0)


for _ in range(3):
    step()
torch.cuda.synchronize()
host, dev = [], []
for _ in range(10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    step()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    host.append((t1 - t0) * 1e3)
    dev.append(e0.elapsed_time(e1))
print(f'crops {n}: host enqueue {sorted(host)[len(host) // 2]:.2f} ms / step, device {sorted(dev)[len(dev) // 2]:.2f} ms / step')
