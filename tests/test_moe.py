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


def _train_batch(n, K, seed, src):
    g = torch.Generator().manual_seed(seed)
    ys, xs = torch.meshgrid(torch.arange(64.), torch.arange(48.), indexing='ij')
    cx = torch.rand(n, K, generator=g) * 47
    cy = torch.rand(n, K, generator=g) * 63
    t = torch.exp(-((xs - cx[..., None, None]) ** 2 + (ys - cy[..., None, None]) ** 2) / 8.0).contiguous()
    w = (torch.rand(n, K, 1, generator=g) > 0.2).float()
    return synthetic.synthetic_crops(n, seed), t, w, [dict(dataset_idx=int(d)) for d in src]


def _moe_train_cfg(depth=2):
    """every head with the keypoint count of the batch's targets (the reference's get_loss needs that, mse_loss.py:30)"""
    cfg = _moe_cfg(depth=depth)
    cfg['associate_keypoint_head'] = [dict(cfg['keypoint_head']), dict(cfg['keypoint_head'])]
    return cfg


@pytest.mark.reference
def test_moe_training_oracle_equals_live_reference():
    """Pins oracle.train_loss_and_grads_moe: the unmodified reference TopDownMoE.forward_train + _parse_losses'
    loss + backward() on the same weights and a batch that mixes datasets 0 and 2 (dataset 1 absent)."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip('reference tree not mounted')
    cfg = _moe_train_cfg()
    ref = ref_loader.build_reference_topdown_moe(cfg)
    _randomise(ref, 3)
    ref.train()
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    src = [2, 0, 0, 2, 2]
    img, target, tw, metas = _train_batch(5, 5, 3, src)
    losses = ref.forward_train(img, target, tw, metas)
    loss, log_vars = ref._parse_losses(losses)
    loss.backward()
    o_losses, _, g = VT.train_loss_and_grads_moe(sd, img, target, tw, cfg, torch.tensor(src))
    assert set(o_losses) == {k for k in losses if 'loss' in k} == {'main_stream_loss', '1_loss', '2_loss'}
    for k, v in o_losses.items():
        assert abs(float(v) - float(losses[k])) <= 1e-6 * max(1.0, abs(float(v))), k
    assert float(o_losses['1_loss']) == 0.0
    for nm, p in ref.named_parameters():
        assert nm in g, nm
        if p.grad is None:
            assert g[nm] is None or float(g[nm].abs().max()) == 0.0, nm
            continue
        assert torch.allclose(p.grad, g[nm], rtol=1e-4, atol=1e-7), nm
    # the absent dataset's expert still gets a (zero) gradient: the dense masked form of vit_moe.py:107-111
    assert float(g['backbone.blocks.0.mlp.experts.1.weight'].abs().max()) == 0.0
    for k, v in sd.items():      # BatchNorm running statistics of every head were updated identically
        if 'running_' in k:
            assert torch.allclose(v, ref.state_dict()[k], rtol=1e-5, atol=1e-7), k


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.gpu
@pytest.mark.parametrize('src,fused,drop', [([2, 0, 0, 2, 1, 2], True, 0.0), ([2, 0, 0, 2, 2], True, 0.0),
                                            ([1, 1, 1], True, 0.0), ([0, 2, 1, 0], False, 0.0),
                                            ([2, 0, 1, 2, 1, 0, 0, 2], True, 0.4)])
def test_topdown_moe_forward_train_backward_vs_oracle(src, fused, drop, monkeypatch):
    """TopDownMoE.forward_train -> sum of the losses -> backward() on the B200 path against torch.autograd over the fp32
    oracle (pinned to the live reference above): mixed batches in arbitrary order, an absent dataset (its expert and
    head get zero / no gradient), a homogeneous batch. Tolerances as in tests/test_gpu_train_step.py."""
    import vitpose_b200 as V
    from vitpose_b200 import training
    monkeypatch.setattr(training, 'FUSE_MLP', fused)
    cfg = _moe_train_cfg()
    model = V.build_posenet(cfg)
    _randomise(model, 5)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    n = len(src)
    img, target, tw, metas = _train_batch(n, 5, 5, src)
    ref_sd = {k: v.clone() for k, v in sd.items()}
    scales = None
    if drop > 0:            # stochastic depth: the SAME per-crop factors in both implementations (batch order)
        gen = torch.Generator().manual_seed(13)
        scales = []
        for p_drop in torch.linspace(0, drop, 2).tolist():
            keep = 1.0 - p_drop
            scales.append(tuple(torch.floor(keep + torch.rand(n, generator=gen)) / keep for _ in range(2)))
        assert any((s_ == 0).any() for pair in scales for s_ in pair), 'the masks must drop something'
    l_ref, hm_ref, g_ref = VT.train_loss_and_grads_moe(ref_sd, img, target, tw, cfg, torch.tensor(src),
                                                       drop_scales=scales)
    model = model.cuda().train()
    if scales is not None:
        # forward_train sorts the crops by dataset: the injected factors follow the crops
        order, _ = model.backbone.dataset_runs(src)
        perm = torch.tensor(order)
        model.backbone.drop_path_rate = drop
        model.backbone._drop_path_scales = [tuple(s_[perm].float().cuda().contiguous() for s_ in pair)
                                            for pair in scales]
    out = model.train_step(dict(img=img.cuda(), target=target.cuda(), target_weight=tw.cuda(), img_metas=metas))
    assert set(out['log_vars']) == {'main_stream_loss', 'main_stream_acc', '1_loss', '1_acc', '2_loss', '2_acc', 'loss'}
    out['loss'].backward()
    torch.cuda.synchronize()
    for k, v in l_ref.items():
        assert abs(out['log_vars'][k] - float(v)) <= 2e-2 * abs(float(v)) + 1e-12, (k, out['log_vars'][k], float(v))
    assert abs(out['loss'].item() - float(sum(l_ref.values()))) <= 2e-2 * float(sum(l_ref.values()))
    present = set(src)
    worst = {}
    for nm, p in model.named_parameters():
        rf = g_ref[nm]
        if p.grad is None:          # only a head whose loss is identically zero may lack a gradient
            assert rf is None or float(rf.abs().max()) == 0.0, nm
            continue
        gr = p.grad.cpu()
        assert gr.shape == rf.shape and torch.isfinite(gr).all(), nm
        if rf.norm() < 1e-12:
            assert float(gr.abs().max()) <= 1e-12, (nm, float(gr.abs().max()))
            continue
        worst[nm] = (_rel(gr, rf), _cos(gr, rf))
    for e in range(3):              # every expert has a gradient tensor; absent datasets' are exactly zero
        for l in range(2):
            gw = dict(model.named_parameters())[f'backbone.blocks.{l}.mlp.experts.{e}.weight'].grad
            assert gw is not None
            live = float(g_ref[f'backbone.blocks.{l}.mlp.experts.{e}.weight'].abs().max()) > 0
            assert (float(gw.abs().max()) > 0) == live, (e, l)
            if drop == 0:
                assert live == (e in present), (e, l)
    bad = {k: v for k, v in worst.items() if v[0] > 0.15 or v[1] < 0.99}
    top = sorted(worst.items(), key=lambda kv: -kv[1][0])[:5]
    assert not bad, f'gradient mismatch (rel err, cosine): {bad}; worst five: {top}'
    names = [nm for nm, p in model.named_parameters() if p.grad is not None]
    flat = torch.cat([dict(model.named_parameters())[nm].grad.flatten().cpu() for nm in names])
    flat_ref = torch.cat([g_ref[nm].flatten() for nm in names])
    assert _cos(flat, flat_ref) > 0.998 and abs(float(flat.norm() / flat_ref.norm()) - 1) < 0.02
    # BatchNorm running statistics of every head that ran
    for k in ref_sd:
        if 'running_' in k:
            assert _rel(model.state_dict()[k].cpu(), ref_sd[k]) < 2e-2, k


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


def test_moe_forward_train_has_no_cpu_path_and_host_chunk_follows_width():
    """The product path fails loudly without CUDA tensors (no CPU fallback); the first H2D chunk of forward_test is
    chosen from the backbone width (64 crops below D = 768, 32 otherwise)."""
    import vitpose_b200 as V
    from vitpose_b200 import _lib
    cfg = _moe_train_cfg()
    model = V.build_posenet(cfg).train()
    img, target, tw, metas = _train_batch(3, 5, 1, [1, 0, 2])
    with pytest.raises(_lib.VitposeLibError):
        model(img=img, target=target, target_weight=tw, img_metas=metas, return_loss=True)
    assert model._vpb_dataset_runs is None                       # reset even when the step fails
    assert model.host_chunk == 64                                # tiny config: D = 128
    b = V.build_posenet(configs.baseline_model_cfg('B-classic-17'))
    assert b.host_chunk == 32 and b.host_chunks == []
