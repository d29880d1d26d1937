"""Training-config rows of SURVEY.md §8 (a17 JointsMSELoss, a18 layer-decay AdamW): host grouping logic on CPU,
kernels on the GPU against torch / golden KATs."""
import os

import numpy as np
import pytest
import torch

import vitpose_b200 as V
from vitpose_b200 import configs, optim


def test_layer_decay_groups_base_config():
    """ViTPose-B: 14 layer ids x {decay, no_decay}; lr = 5e-4 * 0.75 ** (13 - layer_id)
    (mmcv_custom/layer_decay_optimizer_constructor.py:36-60, ViTPose_base_coco_256x192.py:16-28)."""
    cfg = configs.baseline_model_cfg('B-classic-17')
    cfg['backbone']['depth'] = 12
    model = V.build_posenet(cfg)
    groups = optim.layer_decay_param_groups(model, 5e-4, 0.1, 12, 0.75)
    by_name = {g['group_name']: g for g in groups}
    assert set(by_name) == {f'layer_{i}_{k}' for i in range(14) for k in ('decay', 'no_decay')}
    assert by_name['layer_0_no_decay']['param_names'][:1] == ['backbone.pos_embed']
    assert 'backbone.patch_embed.proj.weight' in by_name['layer_0_decay']['param_names']
    assert 'backbone.blocks.3.attn.qkv.weight' in by_name['layer_4_decay']['param_names']
    assert 'backbone.blocks.3.attn.qkv.bias' in by_name['layer_4_no_decay']['param_names']
    assert 'keypoint_head.final_layer.weight' in by_name['layer_13_decay']['param_names']
    assert 'backbone.last_norm.weight' in by_name['layer_13_no_decay']['param_names']
    for i in range(14):
        g = by_name[f'layer_{i}_decay']
        assert abs(g['lr'] - 5e-4 * 0.75 ** (13 - i)) < 1e-12 and g['weight_decay'] == 0.1
        assert by_name[f'layer_{i}_no_decay']['weight_decay'] == 0.0
    n_params = sum(len(g['params']) for g in groups)
    assert n_params == len([p for p in model.parameters() if p.requires_grad])


@pytest.mark.reference
def test_layer_decay_groups_match_reference_constructor():
    """Run the reference's own constructor file (loaded by path, mmcv.runner shimmed) on the same module."""
    import importlib.util
    import sys
    import types
    from oracle import ref_loader
    ref_loader.load_reference()

    class _Default:
        def __init__(self, optimizer_cfg, paramwise_cfg=None):
            self.optimizer_cfg, self.paramwise_cfg = optimizer_cfg, paramwise_cfg
            self.base_lr, self.base_wd = optimizer_cfg.get('lr'), optimizer_cfg.get('weight_decay')

    class _Reg:
        def register_module(self):
            return lambda c: c

    runner = sys.modules['mmcv.runner']
    runner.OPTIMIZER_BUILDERS, runner.DefaultOptimizerConstructor = _Reg(), _Default
    runner.get_dist_info = lambda: (1, 1)
    spec = importlib.util.spec_from_file_location(
        'ref_layer_decay', os.path.join(ref_loader.REF_ROOT, 'mmcv_custom/layer_decay_optimizer_constructor.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    model = V.build_posenet(configs.tiny_model_cfg(5, depth=3))
    ctor = mod.LayerDecayOptimizerConstructor(dict(type='AdamW', lr=5e-4, weight_decay=0.1),
                                              dict(num_layers=3, layer_decay_rate=0.75))
    ref_groups = []
    ctor.add_params(ref_groups, model)
    ours = optim.layer_decay_param_groups(model, 5e-4, 0.1, 3, 0.75)
    assert [g['group_name'] for g in ref_groups] == [g['group_name'] for g in ours]
    for a, b in zip(ref_groups, ours):
        assert a['param_names'] == b['param_names'] and a['lr'] == b['lr'] and a['weight_decay'] == b['weight_decay']


@pytest.mark.gpu
def test_joints_mse_loss_kats_and_grad(golden_dir):
    loss = V.build_loss(dict(type='JointsMSELoss', use_target_weight=True))
    dev = torch.device('cuda:0')
    z, one = torch.zeros(1, 3, 64, 64, device=dev), torch.ones(1, 3, 64, 64, device=dev)
    w1 = torch.ones(1, 3, 1, device=dev)
    assert torch.allclose(loss(z, z, w1), torch.tensor(0., device=dev))      # test_top_down_losses.py:27-41
    assert torch.allclose(loss(one, z, w1), torch.tensor(1., device=dev))
    plain = V.build_loss(dict(type='JointsMSELoss'))
    p = torch.zeros(1, 2, 64, 64, device=dev)
    p[0, 0] += 1
    assert torch.allclose(plain(p, torch.zeros_like(p), None), torch.tensor(0.5, device=dev))
    g = np.load(os.path.join(golden_dir, 'loss_kat.npz'))
    o, t, w = (torch.from_numpy(g[k]).to(dev) for k in ('output', 'target', 'weight'))
    np.testing.assert_allclose(loss(o, t, w).item(), g['loss_weighted'], rtol=1e-5)
    np.testing.assert_allclose(plain(o, t, None).item(), g['loss_unweighted'], rtol=1e-5)
    # gradient vs autograd of the torch expression
    o1 = o.clone().requires_grad_(True)
    loss(o1, t, w).backward()
    o2 = o.clone().requires_grad_(True)
    n, k = o.shape[:2]
    (((o2.reshape(n, k, -1) - t.reshape(n, k, -1)) * w) ** 2).mean(dim=(0, 2)).sum().div(k).backward()
    np.testing.assert_allclose(o1.grad.cpu().numpy(), o2.grad.cpu().numpy(), rtol=1e-5, atol=1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize('multi', [True, False])
def test_layer_decay_adamw_matches_torch(multi):
    """multi: one vpb_adamw_multi launch for all tensors; else the per-tensor vpb_adamw_step path."""
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    cfg = configs.tiny_model_cfg(5, depth=2)
    m1 = V.build_posenet(cfg).to(dev)
    m2 = V.build_posenet(cfg).to(dev)
    m2.load_state_dict(m1.state_dict())
    ctor = optim.LayerDecayOptimizerConstructor(dict(type='AdamW', lr=5e-4, betas=(0.9, 0.999), weight_decay=0.1),
                                                dict(num_layers=2, layer_decay_rate=0.75))
    opt1 = ctor(m1)
    groups2 = optim.layer_decay_param_groups(m2, 5e-4, 0.1, 2, 0.75)
    opt2 = torch.optim.AdamW([dict(params=g['params'], lr=g['lr'], weight_decay=g['weight_decay']) for g in groups2],
                             lr=5e-4, betas=(0.9, 0.999), weight_decay=0.1)
    for step in range(3):
        gen = torch.Generator(device='cuda').manual_seed(step)
        for p1, p2 in zip(m1.parameters(), m2.parameters()):
            gr = torch.randn(p1.shape, device=dev, generator=gen) * 0.05
            p1.grad, p2.grad = gr.clone(), gr.clone()
        total = opt1.step(max_norm=1.0) if multi else opt1.step_per_tensor(max_norm=1.0)
        ref_total = torch.nn.utils.clip_grad_norm_(m2.parameters(), 1.0)
        opt2.step()
        assert abs(total.item() - ref_total.item()) / ref_total.item() < 1e-4
    for (n1, p1), (_, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert torch.allclose(p1, p2, rtol=2e-5, atol=2e-7), n1
