"""Training step (-m gpu): TopDown.forward_train -> loss.backward() on the B200 path against
torch.autograd over the fp32 oracle (functional restatement of the reference modules, BatchNorm in training mode),
on the same weights and batch; then the layer-decay AdamW update.  Tolerances are bf16-level: GEMM operands and the
activation gradients between operators are bf16 on the GPU path, everything else fp32."""
import numpy as np
import pytest
import torch

from oracle import vitpose_torch as VT
from vitpose_b200 import configs, synthetic

pytestmark = pytest.mark.gpu


def _targets(n, K, seed):
    g = torch.Generator().manual_seed(seed)
    ys, xs = torch.meshgrid(torch.arange(64.), torch.arange(48.), indexing='ij')
    cx = torch.rand(n, K, generator=g) * 47
    cy = torch.rand(n, K, generator=g) * 63
    t = torch.exp(-((xs - cx[..., None, None]) ** 2 + (ys - cy[..., None, None]) ** 2) / (2 * 2.0 ** 2))
    w = (torch.rand(n, K, 1, generator=g) > 0.2).float()
    return t.contiguous(), w


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize('name,n,depth,drop,fused', [('tiny', 4, 2, 0.0, True), ('B-classic-17', 3, 2, 0.0, True),
                                                     ('B-classic-17', 2, 12, 0.0, True), ('B-classic-17', 6, 3, 0.3, True),
                                                     ('B-classic-17', 6, 3, 0.3, False),
                                                     ('L-classic-17', 2, 2, 0.0, True),
                                                     ('S-classic-17', 3, 2, 0.0, True),
                                                     ('H-classic-17', 2, 2, 0.0, True),
                                                     ('B-simple-17', 3, 2, 0.0, True)])
def test_forward_train_backward_vs_oracle(name, n, depth, drop, fused, monkeypatch):
    """drop > 0: stochastic depth with the SAME per-crop masks injected into both implementations. fused = the MLP /
    bias-gradient fusions of the training step (training.FUSE_MLP); the un-fused kernels stay covered by one case."""
    import vitpose_b200 as V
    from vitpose_b200 import training
    monkeypatch.setattr(training, 'FUSE_MLP', fused)
    if name == 'tiny':
        cfg = configs.tiny_model_cfg(5)
    elif name == 'L-classic-17':      # ViTPose-L with the classic decoder (the reference trains it: logs/vitpose-l.log.json)
        cfg = configs.baseline_model_cfg('B-classic-17')
        cfg['backbone'].update(embed_dim=1024, num_heads=16)
        cfg['keypoint_head'].update(in_channels=1024)
    elif name == 'S-classic-17':      # ViTPose-S: D = 384, head_dim 32
        cfg = configs.baseline_model_cfg('B-classic-17')
        cfg['backbone'].update(embed_dim=384, num_heads=12)
        cfg['keypoint_head'].update(in_channels=384)
    elif name == 'H-classic-17':      # ViTPose-H: D = 1280, head_dim 80 (logs/vitpose-h.log.json)
        cfg = configs.baseline_model_cfg('B-classic-17')
        cfg['backbone'].update(embed_dim=1280, num_heads=16)
        cfg['keypoint_head'].update(in_channels=1280)
    elif name == 'B-simple-17':       # the simple decoder (ReLU -> bilinear x4 -> 3x3 conv; logs/vitpose-*-simple.log.json)
        cfg = configs.baseline_model_cfg('L-simple-17')
        cfg['backbone'].update(embed_dim=768, num_heads=12)
        cfg['keypoint_head'].update(in_channels=768)
    else:
        cfg = configs.baseline_model_cfg(name)
    simple = cfg['keypoint_head'].get('num_deconv_layers', 3) == 0
    cfg['backbone'].update(depth=depth, drop_path_rate=0.0)
    K = cfg['keypoint_head']['out_channels']
    sd = synthetic.scaled_init_state_dict(cfg, 7)
    img = synthetic.synthetic_crops(n, 7)
    target, tw = _targets(n, K, 7)
    ref_sd = {k: v.clone() for k, v in sd.items()}
    scales = None
    if drop > 0:
        gen = torch.Generator().manual_seed(11)
        scales = []
        for p in torch.linspace(0, drop, depth).tolist():
            keep = 1.0 - p
            scales.append(tuple(torch.floor(keep + torch.rand(n, generator=gen)) / keep for _ in range(2)))
        assert any((s == 0).any() for pair in scales for s in pair), 'the masks must drop something'
    loss_ref, hm_ref, g_ref = VT.train_loss_and_grads(ref_sd, img, target, tw, cfg, drop_scales=scales)

    cfg['backbone']['drop_path_rate'] = drop
    model = V.build_posenet(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    if scales is not None:
        model.backbone._drop_path_scales = [tuple(s.float().cuda().contiguous() for s in pair) for pair in scales]
    losses = model(img=img.cuda(), target=target.cuda(), target_weight=tw.cuda(), img_metas=None, return_loss=True)
    assert set(losses) >= {'heatmap_loss'}
    loss = losses['heatmap_loss']
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - loss_ref.item()) <= 2e-2 * abs(loss_ref.item()), (loss.item(), loss_ref.item())
    # BatchNorm running statistics were updated as nn.BatchNorm2d(train) does
    for i in (() if simple else (1, 4)):
        for b in ('running_mean', 'running_var'):
            k = f'keypoint_head.deconv_layers.{i}.{b}'
            assert _rel(model.state_dict()[k].cpu(), ref_sd[k]) < 2e-2, k
    worst = {}
    for nm, p in model.named_parameters():
        assert p.grad is not None, f'{nm}: no gradient'
        gr, rf = p.grad.cpu(), g_ref[nm]
        assert gr.shape == rf.shape and torch.isfinite(gr).all(), nm
        if rf.norm() < 1e-12:
            continue
        worst[nm] = (_rel(gr, rf), _cos(gr, rf))
    # Upstream of the first ReLU everything is a linear function of bf16-rounded operands: tight.
    tight = ('keypoint_head.final_layer.weight', 'keypoint_head.final_layer.bias')
    if not simple:
        tight += ('keypoint_head.deconv_layers.4.weight', 'keypoint_head.deconv_layers.4.bias')
    for nm in tight:
        assert worst[nm][0] < 0.02, (nm, worst[nm])
    # Below a ReLU the comparison is statistical: the GPU path keeps the conv outputs in bf16, so ~0.4 % of the
    # BatchNorm+ReLU units that sit within rounding distance of zero take the other branch than in the fp32 oracle,
    # and a gradient field with a fraction f of its entries toggled is off by ~sqrt(f) (~6 %) in L2 -- on
    # random-init weights that noise is not averaged out by the weight-gradient sums. Direction and norm must agree.
    bad = {k: v for k, v in worst.items() if v[0] > 0.15 or v[1] < 0.99}
    top = sorted(worst.items(), key=lambda kv: -kv[1][0])[:5]
    assert not bad, f'gradient mismatch (rel err, cosine): {bad}; worst five: {top}'
    # whole-model gradient: direction and norm
    flat = torch.cat([p.grad.flatten().cpu() for _, p in model.named_parameters()])
    flat_ref = torch.cat([g_ref[nm].flatten() for nm, _ in model.named_parameters()])
    assert _cos(flat, flat_ref) > 0.998 and abs(float(flat.norm() / flat_ref.norm()) - 1) < 0.02


def test_train_step_with_layer_decay_adamw_decreases_loss():
    import vitpose_b200 as V
    from vitpose_b200.optim import LayerDecayOptimizerConstructor
    cfg = configs.tiny_model_cfg(5)
    cfg['backbone'].update(drop_path_rate=0.0)
    sd = synthetic.scaled_init_state_dict(cfg, 3)
    model = V.build_posenet(cfg)
    model.load_state_dict(sd, strict=True)
    model = model.cuda().train()
    opt = LayerDecayOptimizerConstructor(dict(type='AdamW', lr=2e-3, betas=(0.9, 0.999), weight_decay=0.1),
                                         dict(num_layers=2, layer_decay_rate=0.75))(model)
    n = 4
    img = synthetic.synthetic_crops(n, 3).cuda()
    target, tw = _targets(n, 5, 3)
    batch = dict(img=img, target=target.cuda(), target_weight=tw.cuda(), img_metas=None)
    hist = []
    for _ in range(6):
        out = model.train_step(batch, opt)
        assert set(out) == {'loss', 'log_vars', 'num_samples'} and out['num_samples'] == n
        opt.zero_grad()
        out['loss'].backward()
        norm = opt.step(max_norm=1.0)
        assert torch.isfinite(norm)
        hist.append(out['loss'].item())
    assert hist[-1] < hist[0], hist


def test_drop_path_random_masks_follow_the_schedule():
    """Without injected masks the factors are 0 or 1/keep_prob with keep_prob = 1 - linspace(0, rate, depth)[i]."""
    from types import SimpleNamespace
    from vitpose_b200.training import drop_path_scales
    bb = SimpleNamespace(drop_path_rate=0.3, depth=4, training=True)
    sc = drop_path_scales(bb, 4096, torch.device('cuda:0'))
    assert sc[0] == (None, None)
    for i, p in enumerate(torch.linspace(0, 0.3, 4).tolist()):
        if i == 0:
            continue
        for s in sc[i]:
            vals = torch.unique(s).cpu()
            assert torch.allclose(vals, torch.tensor([0.0, 1.0 / (1.0 - p)]), atol=1e-6)
            assert abs(float((s == 0).float().mean()) - p) < 0.03
    bb.training = False
    assert all(pair == (None, None) for pair in drop_path_scales(bb, 8, torch.device('cuda:0')))
    # frozen blocks are in eval mode (vit.py:249-259): their DropPath is the identity
    import vitpose_b200 as V
    vit = V.ViT(img_size=(256, 192), embed_dim=64, depth=4, num_heads=2, qkv_bias=True, drop_path_rate=0.3,
                frozen_stages=2).train()
    sc = drop_path_scales(vit, 16, torch.device('cuda:0'))
    assert sc[1] == (None, None) and sc[2] == (None, None) and sc[3][0] is not None
