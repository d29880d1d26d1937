"""the ViTPose+ variant (ViTMoE backbone + TopDownMoE detector). Host logic on CPU; the B200
path against the fp32 oracle (MoEMlp restated functionally) on a batch that MIXES dataset indices (-m gpu)."""
import numpy as np
import pytest
import torch

from oracle import vitpose_torch as VT
from vitpose_b200 import configs, synthetic


def _moe_cfg(num_expert=3, part=32, depth=2):
    cfg = configs.tiny_model_cfg(5, depth=depth)
    cfg['type'] = 'TopDownMoE'
    cfg['backbone'].update(type='ViTMoE', num_expert=num_expert, part_features=part, drop_path_rate=0.0)
    cfg['associate_keypoint_head'] = [dict(cfg['keypoint_head'], out_channels=4),
                                      dict(cfg['keypoint_head'], out_channels=6)]
    return cfg


def _randomise(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.dim() > 1:
                p.copy_(torch.randn(p.shape, generator=g) * (0.5 / np.sqrt(p[0].numel())))
            elif n.endswith('.bias'):
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
        for n, b in model.named_buffers():
            if n.endswith('running_var'):
                b.copy_(torch.rand(b.shape, generator=g) + 0.5)
            elif n.endswith('running_mean'):
                b.copy_(torch.randn(b.shape, generator=g) * 0.1)


def test_moe_state_dict_layout_and_split():
    import vitpose_b200 as V
    from vitpose_b200.checkpoint import split_moe_state_dict
    cfg = _moe_cfg()
    model = V.build_posenet(cfg)
    sd = model.state_dict()
    D, part = cfg['backbone']['embed_dim'], 32
    assert sd['backbone.blocks.0.mlp.fc2.weight'].shape == (D - part, 4 * D)
    assert sd['backbone.blocks.1.mlp.experts.2.weight'].shape == (part, 4 * D)
    assert sd['associate_keypoint_heads.1.final_layer.weight'].shape[0] == 6
    _randomise(model, 0)
    sd = model.state_dict()
    # tools/model_split.py: per-dataset plain TopDown checkpoints
    s0 = split_moe_state_dict(sd, 0)
    assert not any('experts' in k or k.startswith('associate') for k in s0)
    w = s0['backbone.blocks.1.mlp.fc2.weight']
    assert w.shape == (D, 4 * D)
    assert torch.equal(w[:D - part], sd['backbone.blocks.1.mlp.fc2.weight'])
    assert torch.equal(w[D - part:], sd['backbone.blocks.1.mlp.experts.0.weight'])
    s2 = split_moe_state_dict(sd, 2, num_keypoints=6)
    assert torch.equal(s2['backbone.blocks.0.mlp.fc2.bias'][D - part:], sd['backbone.blocks.0.mlp.experts.2.bias'])
    assert torch.equal(s2['keypoint_head.final_layer.weight'], sd['associate_keypoint_heads.1.final_layer.weight'][:6])
    plain = configs.tiny_model_cfg(6, depth=2)
    plain['backbone']['drop_path_rate'] = 0.0
    V.build_posenet(plain).load_state_dict(s2, strict=True)     # loads into an ordinary ViT + head
    # the backbone's per-dataset view is the same tensor
    eff = model.backbone.effective_state_dict(2)
    assert torch.equal(eff['blocks.0.mlp.fc2.weight'], s2['backbone.blocks.0.mlp.fc2.weight'])
    with pytest.raises(IndexError):
        model.backbone.effective_state_dict(3)


@pytest.mark.reference
def test_moe_state_dict_keys_equal_reference():
    import importlib.util
    import os
    import sys
    from oracle import ref_loader
    ref_loader.load_reference()
    path = '/root/reference/mmpose/models/backbones/vit_moe.py'
    if not os.path.exists(path):
        pytest.skip('reference tree not mounted')
    spec = importlib.util.spec_from_file_location('mmpose.models.backbones.vit_moe', path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    import vitpose_b200 as V
    kw = dict(img_size=(256, 192), patch_size=16, embed_dim=128, depth=2, num_heads=2, ratio=1, mlp_ratio=4,
              qkv_bias=True, drop_path_rate=0.0, num_expert=3, part_features=32)
    ref = mod.ViTMoE(**kw)
    ours = V.build_backbone(dict(type='ViTMoE', **kw))
    rs, os_ = ref.state_dict(), ours.state_dict()
    assert list(rs) == list(os_)
    for k in rs:
        assert rs[k].shape == os_[k].shape, k
    # and the reference module agrees with the oracle restatement on a mixed batch
    ours.load_state_dict(rs)
    img = synthetic.synthetic_crops(3, 1)
    src = torch.tensor([2, 0, 1])
    ref.eval()
    with torch.no_grad():
        f_ref = ref(img, src)
        f_or = VT.vit_features({'backbone.' + k: v for k, v in rs.items()}, img, 2, 2, dataset_source=src)
    assert torch.allclose(f_ref, f_or, atol=1e-5)


@pytest.mark.gpu
def test_topdown_moe_mixed_batch_vs_oracle():
    import vitpose_b200 as V
    cfg = _moe_cfg()
    model = V.build_posenet(cfg)
    _randomise(model, 1)
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    n = 5
    img = synthetic.synthetic_crops(n, 4)
    metas = synthetic.synthetic_metas(n, 5, 4)
    for m, d in zip(metas, [1, 0, 2, 1, 0]):
        m['dataset_idx'] = d
    ocfg = dict(cfg, backbone=dict(cfg['backbone']))
    ref = VT.forward_test(sd, img, metas, ocfg, return_heatmap=True)
    model = model.cuda().eval()
    r = model(img=img.cuda(), img_metas=metas, return_loss=False, return_heatmap=True)
    err = np.abs(r['output_heatmap'] - ref['output_heatmap']).max()
    std = ref['output_heatmap'].std()
    assert err < 1e-2 and err < 0.1 * std, f'heatmap err {err:.4g}, std {std:.4g}'
    assert r['bbox_ids'] == ref['bbox_ids'] and r['image_paths'] == ref['image_paths']
    np.testing.assert_allclose(r['boxes'], ref['boxes'], rtol=1e-6)
    # the mixed batch ran in ONE forward pass: the engine of dataset 0 with every dataset's mlp.fc2 attached (vpb_moe_runs)
    eng = model.backbone.moe_engine(model.keypoint_head)
    assert sorted(eng.experts) == [0, 1, 2] and getattr(eng, '_moe_runs', None) is None
    order, runs = model.backbone.dataset_runs([1, 0, 2, 1, 0])
    assert order == [1, 4, 0, 3, 2] and runs == [(0, 2), (1, 2), (2, 1)]
    from vitpose_b200 import _lib
    L = _lib.lib()
    c0 = L.vpb_launch_count()
    model(img=img.cuda(), img_metas=metas, return_loss=False)
    per_mixed = L.vpb_launch_count() - c0
    for m in metas:
        m['dataset_idx'] = 1
    c0 = L.vpb_launch_count()
    model(img=img.cuda(), img_metas=metas, return_loss=False)
    per_single = L.vpb_launch_count() - c0
    assert per_mixed == per_single, (per_mixed, per_single)     # same kernel sequence: one pass, not one per dataset
    for m, d in zip(metas, [1, 0, 2, 1, 0]):
        m['dataset_idx'] = d
    # a homogeneous batch goes through the single-engine path; the experts must actually matter
    for m in metas:
        m['dataset_idx'] = 2
    r2 = model(img=img.cuda(), img_metas=metas, return_loss=False, return_heatmap=True)
    ref2 = VT.forward_test(sd, img, metas, ocfg, return_heatmap=True)
    assert np.abs(r2['output_heatmap'] - ref2['output_heatmap']).max() < 1e-2
    assert np.abs(ref2['output_heatmap'] - ref['output_heatmap']).max() > 10 * err
    # the three per-dataset engines share every backbone tensor except mlp.fc2 (one backbone + 3 x fc2 in HBM)
    engines = [model.backbone.engine(model.keypoint_head, d) for d in range(3)]
    base = engines[0].weights if engines[0].weights.shared_bytes == 0 else engines[1].weights
    others = [e.weights for e in engines if e.weights is not base]
    total = sum(t.numel() * t.element_size() for t in base.by_id.values())
    for w in others:
        assert w.shared_bytes > 0.5 * total, (w.shared_bytes, total)
        n_private = sum(1 for i in w.by_id if w.by_id[i].data_ptr() != base.by_id[i].data_ptr())
        assert n_private <= 4 * len(model.backbone.blocks), n_private      # fc2 weight + bias (+ unkeyed small ones)
    feats = model.backbone(img.cuda(), torch.tensor([1, 0, 2, 1, 0]))
    f_ref = VT.vit_features(sd, img, 2, 2, dataset_source=torch.tensor([1, 0, 2, 1, 0]))
    assert (feats.cpu() - f_ref).abs().max() < 0.05 * max(1.0, float(f_ref.abs().max()))


def test_moe_loads_plain_vit_checkpoint(tmp_path):
    """``TopDownMoE(pretrained=<MAE / plain ViT checkpoint>)`` — vit_moe.py:336 passes ``part_features`` to
    load_checkpoint, which splits every fc2 [D, 4D] into the shared fc2 [D - part, 4D] and one copy of the last
    ``part`` rows per expert (mmcv_custom/checkpoint.py:396-405)."""
    import vitpose_b200 as V
    plain_cfg = configs.tiny_model_cfg(5, depth=2)
    plain_cfg['backbone']['img_size'] = (224, 224)        # MAE pretrain: square 14 x 14 position grid, resized on load
    plain = V.build_backbone(plain_cfg['backbone'])
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for p in plain.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.1)
    path = str(tmp_path / 'vit.pth')
    torch.save(dict(state_dict={'backbone.' + k: v for k, v in plain.state_dict().items()}), path)
    cfg = _moe_cfg(num_expert=3, part=32)
    moe = V.build_backbone(cfg['backbone'])
    moe.init_weights(pretrained=path)
    D, part = cfg['backbone']['embed_dim'], 32
    for i in range(2):
        w, b = plain.state_dict()[f'blocks.{i}.mlp.fc2.weight'], plain.state_dict()[f'blocks.{i}.mlp.fc2.bias']
        sd = moe.state_dict()
        assert torch.equal(sd[f'blocks.{i}.mlp.fc2.weight'], w[:D - part])
        assert torch.equal(sd[f'blocks.{i}.mlp.fc2.bias'], b[:D - part])
        for e in range(3):
            assert torch.equal(sd[f'blocks.{i}.mlp.experts.{e}.weight'], w[D - part:])
            assert torch.equal(sd[f'blocks.{i}.mlp.experts.{e}.bias'], b[D - part:])
        assert torch.equal(sd[f'blocks.{i}.attn.qkv.weight'], plain.state_dict()[f'blocks.{i}.attn.qkv.weight'])
    # every dataset's effective FFN is then the plain ViT's
    assert torch.equal(moe.effective_state_dict(1)['blocks.0.mlp.fc2.weight'], plain.state_dict()['blocks.0.mlp.fc2.weight'])
    # freeze_ffn also freezes the expert FFNs that replace the plain ones
    frozen = V.build_backbone(dict(cfg['backbone'], freeze_ffn=True))
    assert not any(p.requires_grad for n, p in frozen.named_parameters() if '.mlp.' in n)
