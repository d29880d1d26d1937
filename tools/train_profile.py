"""Where a training step spends its time: host wall-clock (launch-bound?) vs device time per phase."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitpose_b200 as V
from vitpose_b200 import configs, synthetic, _lib
from vitpose_b200.optim import LayerDecayOptimizerConstructor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = configs.baseline_model_cfg('B-classic-17'); cfg['backbone']['drop_path_rate'] = 0.0
K = 17
model = V.build_posenet(cfg); model.load_state_dict(synthetic.scaled_init_state_dict(cfg, 0)); model = model.cuda().train()
opt = LayerDecayOptimizerConstructor(dict(type='AdamW', lr=5e-4, betas=(0.9, 0.999), weight_decay=0.1),
                                     dict(num_layers=12, layer_decay_rate=0.75))(model)
img = synthetic.synthetic_crops(n, 0).cuda()
tgt = torch.rand(n, K, 64, 48, device='cuda'); tw = torch.ones(n, K, 1, device='cuda')
batch = dict(img=img, target=tgt, target_weight=tw, img_metas=None)

def phase(fn):
    torch.cuda.synchronize(); c0 = _lib.ABI_CALLS[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(); r = fn(); e1.record(); host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return r, host * 1e3, e0.elapsed_time(e1), _lib.ABI_CALLS[0] - c0

for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    out, h1, d1, c1 = phase(lambda: model.train_step(batch, opt))
    opt.zero_grad(set_to_none=True)
    _, h2, d2, c2 = phase(lambda: out['loss'].backward())
    _, h3, d3, c3 = phase(lambda: opt.step(max_norm=1.0))
    print(f'iter {it}: forward host {h1:.1f} ms dev {d1:.1f} ms calls {c1} | backward host {h2:.1f} dev {d2:.1f} calls {c2} '
          f'| optimizer host {h3:.1f} dev {d3:.1f} calls {c3}')
