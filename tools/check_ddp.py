"""2-GPU check (torchrun): the gradient all-reduce overlapped inside backward() == local backward + explicit
parallel.allreduce_gradients, and all ranks end up with identical parameters after an optimizer step."""
import os, sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vitpose_b200 as V
from vitpose_b200 import configs, synthetic, parallel
from vitpose_b200.optim import LayerDecayOptimizerConstructor

rank, lr = int(os.environ['RANK']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr)
dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
cfg = configs.baseline_model_cfg('B-classic-17'); cfg['backbone'].update(depth=4, drop_path_rate=0.0)
sd = synthetic.scaled_init_state_dict(cfg, 0)
n, K = 8, 17
img = synthetic.synthetic_crops(n, 100 + rank).cuda()
tgt = torch.rand(n, K, 64, 48, generator=torch.Generator().manual_seed(rank)).cuda()
tw = torch.ones(n, K, 1).cuda()
grads = {}
for mode in ('in_backward', 'in_backward2', 'explicit'):
    model = V.build_posenet(cfg); model.load_state_dict(sd); model = model.cuda().train()
    model.allreduce_in_backward = mode.startswith('in_backward')
    out = model.train_step(dict(img=img, target=tgt, target_weight=tw, img_metas=None), None)
    out['loss'].backward()
    if mode == 'explicit':
        parallel.allreduce_gradients(list(model.parameters()))
    grads[mode] = {k: p.grad.clone() for k, p in model.named_parameters()}
def _diff(a, b):
    d = {k: float((grads[a][k] - grads[b][k]).norm() / (grads[b][k].norm() + 1e-30)) for k in grads[b]}
    return sorted(d.items(), key=lambda kv: -kv[1])[:4]
if rank == 0:
    print('in_backward vs in_backward2 (run-to-run):', _diff('in_backward', 'in_backward2'))
    print('in_backward vs explicit:', _diff('in_backward', 'explicit'))
worst = max(float((grads['in_backward'][k] - grads['explicit'][k]).norm() / (grads['explicit'][k].norm() + 1e-30))
            for k in grads['explicit'])
opt = LayerDecayOptimizerConstructor(dict(type='AdamW', lr=5e-4, betas=(0.9, 0.999), weight_decay=0.1),
                                     dict(num_layers=4, layer_decay_rate=0.75))(model)
opt.step(max_norm=1.0)
flat = torch.cat([p.detach().flatten() for p in model.parameters()])
other = [torch.empty_like(flat) for _ in range(dist.get_world_size())]
dist.all_gather(other, flat)
same = all(torch.equal(o, flat) for o in other)
if rank == 0:
    print(f'overlapped vs explicit all-reduce: worst relative difference {worst:.3e}; replicas identical after step: {same}')
    assert worst < 1e-5 and same
dist.destroy_process_group()
