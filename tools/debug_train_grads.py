"""TEST INFRASTRUCTURE (diagnostic, uses the oracle as checker like tests/ do; never imported by the product path):
per-parameter gradient error of the training step vs autograd over the fp32 oracle.
    python tools/debug_train_grads.py B-classic-17 3 2"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from oracle import vitpose_torch as VT
from vitpose_b200 import configs, synthetic
import vitpose_b200 as V
from test_gpu_train_step import _targets, _rel, _cos

name, n, depth = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = configs.tiny_model_cfg(5) if name == 'tiny' else configs.baseline_model_cfg(name)
cfg['backbone'].update(depth=depth, drop_path_rate=0.0)
K = cfg['keypoint_head']['out_channels']
sd = synthetic.scaled_init_state_dict(cfg, 7)
img = synthetic.synthetic_crops(n, 7)
target, tw = _targets(n, K, 7)
ref_sd = {k: v.clone() for k, v in sd.items()}
loss_ref, hm_ref, g_ref = VT.train_loss_and_grads(ref_sd, img, target, tw, cfg)
model = V.build_posenet(cfg); model.load_state_dict(sd, strict=True); model = model.cuda().train()
losses = model(img=img.cuda(), target=target.cuda(), target_weight=tw.cuda(), img_metas=None, return_loss=True)
losses['heatmap_loss'].backward()
print('loss', losses['heatmap_loss'].item(), loss_ref.item())
for nm, p in model.named_parameters():
    print(f'{nm:55s} rel {_rel(p.grad.cpu(), g_ref[nm]):.4f} cos {_cos(p.grad.cpu(), g_ref[nm]):.5f} norm {g_ref[nm].norm():.3e}')
