"""Run-to-run determinism of the operators of the training step (same inputs, two launches)."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitpose_b200 import ops, _lib, configs, synthetic
import vitpose_b200 as V
BF16 = torch.bfloat16
dev = torch.device('cuda:0')
g = torch.Generator(device='cuda').manual_seed(0)
def rnd(*s, dtype=BF16, scale=1.0): return (torch.randn(*s, device=dev, generator=g) * scale).to(dtype)
def rep(name, fn, n=3):
    outs = [fn() for _ in range(n)]
    torch.cuda.synchronize()
    def flat(o): return torch.cat([t.float().flatten() for t in (o if isinstance(o, (tuple, list)) else [o])])
    d = max(float((flat(outs[0]) - flat(o)).abs().max()) for o in outs[1:])
    print(f'{name:28s} max |run0 - runk| = {d:.3e}')
M, D = 192 * 64, 768
x, w = rnd(M, D), rnd(3 * D, D, scale=1 / math.sqrt(D))
rep('gemm bias bf16', lambda: ops.gemm(x, w, _lib.EPI_BIAS_BF16))
res = rnd(M, D, dtype=torch.float32); wp = rnd(D, D, scale=1 / math.sqrt(D)); gm = torch.ones(D, device=dev); bt = torch.zeros(D, device=dev)
rep('gemm resid+ln (short K)', lambda: ops.gemm_layernorm(x, wp, _lib.EPI_RESID_F32, bt, res, gm, bt))
h = rnd(M, 4 * D); w2 = rnd(D, 4 * D, scale=1 / math.sqrt(4 * D))
rep('gemm resid+ln (long K)', lambda: ops.gemm_layernorm(h, w2, _lib.EPI_RESID_F32, bt, res, gm, bt))
qkv = rnd(64, 192, 3 * D)
rep('attention fwd', lambda: ops.attention(qkv, 12))
o, lse_ = ops.attention_with_lse(qkv, 12); do = rnd(64, 192, D)
rep('attention bwd', lambda: ops.attention_bwd(qkv, o, lse_, do, 12))
dy = rnd(M, D)
def lnb():
    dx = torch.zeros(M, D, device=dev); dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev)
    ops.layernorm_bwd(res, gm, dy, dx, dg, db); return dx
rep('layernorm bwd dx', lnb)
rep('gelu bwd', lambda: ops.gelu_bwd(h, h))
rep('transpose', lambda: ops.transpose(h))
from vitpose_b200.engine import pack_deconv_weight, pack_deconv_weight_dgrad
f = rnd(64, 16, 12, D); wt = pack_deconv_weight(rnd(D, 256, 4, 4, scale=0.02).float())
rep('deconv raw', lambda: ops.deconv4x4s2_raw(f, wt))
raw = ops.deconv4x4s2_raw(f, wt)
rep('bn stats (atomics)', lambda: ops.bn_train_stats(raw))
mean, rstd = ops.bn_train_stats(raw); g1 = torch.ones(256, device=dev); b1 = torch.zeros(256, device=dev)
rep('bn relu fwd', lambda: ops.bn_relu_fwd(raw, mean, rstd, g1, b1))
dact = rnd(*raw.shape)
rep('bn relu bwd (draw)', lambda: ops.bn_relu_bwd(raw, dact, mean, rstd, g1, b1, torch.zeros(256, device=dev), torch.zeros(256, device=dev)))
rep('deconv dgrad', lambda: ops.gemm(ops.deconv_gather_dy(dact), pack_deconv_weight_dgrad(wt), _lib.EPI_BIAS_BF16))
# whole forward / backward
cfg = configs.baseline_model_cfg('B-classic-17'); cfg['backbone'].update(depth=4, drop_path_rate=0.0)
model = V.build_posenet(cfg); model.load_state_dict(synthetic.scaled_init_state_dict(cfg, 0)); model = model.cuda().train()
img = synthetic.synthetic_crops(8, 1).cuda(); tgt = torch.rand(8, 17, 64, 48, device=dev); tw = torch.ones(8, 17, 1, device=dev)
from vitpose_b200.training import network_heatmaps_train
rep('train forward heatmaps', lambda: network_heatmaps_train(model, img).detach())
def grads():
    model.zero_grad(set_to_none=True)
    model(img=img, target=tgt, target_weight=tw, img_metas=None, return_loss=True)['heatmap_loss'].backward()
    return [p.grad.clone() for p in model.parameters()]
gs = [grads() for _ in range(2)]
names = [n for n, _ in model.named_parameters()]
d = sorted(((float((a - b).norm() / (b.norm() + 1e-30)), n) for n, a, b in zip(names, gs[0], gs[1])), reverse=True)[:5]
print('param grads run-to-run (rel):', d)
