"""CPU-side tests (-m "not gpu"): C-ABI library loads and exports what include/vitpose_b200.h declares,
registry / constructor contract of the host mirror, checkpoint layout, weight repacking, loud failure
without CUDA.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import vitpose_b200 as V
from vitpose_b200 import _lib, configs, synthetic
from vitpose_b200.engine import (decode_mode_from_cfg, fold_bn, model_desc_from_cfg, pack_deconv_weight,
                                 resolve_decode_mode)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from vitpose_b200 import build
    build.build()
    header = open(os.path.join(ROOT, 'include', 'vitpose_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(vpb_[a-z0-9_]+)\s*\(', header)))
    assert len(declared) >= 15
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f'{name} declared in the header but not exported'
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS)
    assert _lib.lib().vpb_abi_version() == _lib.ABI_VERSION
    assert f'#define VPB_ABI_VERSION {_lib.ABI_VERSION}' in header


def test_struct_layout_matches_header():
    # field order/size sanity: 13 int32/float fields + int32[3]
    assert ctypes.sizeof(_lib.ModelDesc) == 4 * 15
    assert ctypes.sizeof(_lib.BlockWeights) == 8 * 12
    assert ctypes.sizeof(_lib.BlockFold) == 8 * 6
    assert ctypes.sizeof(_lib.MoeRuns) == 8 * 4
    assert ctypes.sizeof(_lib.Weights) == 8 * (6 + 9 + 2 + 1 + 1)


def test_workspace_bytes_no_gpu_needed():
    cfg = configs.baseline_model_cfg('B-classic-17')
    d = model_desc_from_cfg(cfg['backbone'], cfg['keypoint_head'])
    b1 = _lib.lib().vpb_workspace_bytes(ctypes.byref(d), 64)
    b2 = _lib.lib().vpb_workspace_bytes(ctypes.byref(d), 128)
    assert 0 < b1 < b2 and abs(b2 / b1 - 2) < 0.01
    rows = 64 * 192
    assert b1 >= rows * (3072 * 2 + 768 * 4 + 768 * 2 + 2304 * 2 + 768 * 2)


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    hm = np.zeros((1, 1, 64, 48), dtype=np.float32)
    with pytest.raises(_lib.VitposeLibError):
        V.keypoints_from_heatmaps(hm, np.zeros((1, 2), np.float32), np.ones((1, 2), np.float32))
    with pytest.raises(_lib.VitposeLibError):
        V.flip_back(hm, [])
    model = V.build_posenet(configs.tiny_model_cfg(5))
    with pytest.raises(RuntimeError):
        model.forward_test(torch.zeros(1, 3, 256, 192), [dict(center=[1, 1], scale=[1, 1], image_file='')])
    from vitpose_b200 import ops
    with pytest.raises(_lib.VitposeLibError):
        ops.layernorm(torch.zeros(4, 128), torch.ones(128), torch.zeros(128))


def test_registry_builds_reference_config_blocks():
    for name in configs.BASELINE_CONFIGS:
        cfg = configs.baseline_model_cfg(name)
        if name.startswith(('L', 'H')):
            cfg['backbone'].update(depth=2)          # keep the CPU test light; structure is depth-independent
        m = V.build_posenet(cfg)
        assert type(m).__name__ == 'TopDown' and type(m.backbone).__name__ == 'ViT'
        assert type(m.keypoint_head).__name__ == 'TopdownHeatmapSimpleHead'
        assert set(synthetic.state_dict_shapes(cfg)) <= set(m.state_dict())
        for k, shape in synthetic.state_dict_shapes(cfg).items():
            assert tuple(m.state_dict()[k].shape) == tuple(shape), k
    with pytest.raises(KeyError):
        V.build_backbone(dict(type='ResNet'))


@pytest.mark.reference
def test_state_dict_keys_equal_reference():
    from oracle import ref_loader
    for decoder in ('classic', 'simple'):
        cfg = configs.tiny_model_cfg(5, decoder)
        ref = ref_loader.build_reference_topdown(cfg)
        ours = V.build_posenet(cfg)
        rs, os_ = ref.state_dict(), ours.state_dict()
        assert list(rs) == list(os_)
        for k in rs:
            assert rs[k].shape == os_[k].shape, k
        ours.load_state_dict(rs, strict=True)


def test_head_constructor_errors():
    """tests/test_models/test_top_down_head.py:139-190 error cases."""
    loss = dict(type='JointsMSELoss', use_target_weight=True)
    with pytest.raises(TypeError):
        V.TopdownHeatmapSimpleHead(out_channels=3, in_channels=512, extra=[], loss_keypoint=loss)
    with pytest.raises(ValueError):
        V.TopdownHeatmapSimpleHead(out_channels=3, in_channels=512, num_deconv_layers=3,
                                   num_deconv_filters=(256, 256), num_deconv_kernels=(4, 4), loss_keypoint=loss)
    with pytest.raises(ValueError):
        V.TopdownHeatmapSimpleHead(out_channels=3, in_channels=512, num_deconv_layers=3,
                                   num_deconv_filters=(256, 256, 256), num_deconv_kernels=(4, 4), loss_keypoint=loss)
    with pytest.raises(ValueError):
        V.TopdownHeatmapSimpleHead(out_channels=3, in_channels=512, num_deconv_layers=-1, loss_keypoint=loss)
    with pytest.raises(ValueError):
        V.TopdownHeatmapSimpleHead(out_channels=3, in_channels=512, num_deconv_layers=3,
                                   num_deconv_filters=(256, 256, 256), num_deconv_kernels=(3, 2, 0),
                                   loss_keypoint=loss)
    h = V.TopdownHeatmapSimpleHead(out_channels=3, in_channels=512, extra={'final_conv_kernel': 3},
                                   loss_keypoint=loss)
    assert h.final_layer.padding == (1, 1)
    h = V.TopdownHeatmapSimpleHead(out_channels=3, in_channels=512, extra={'final_conv_kernel': 1},
                                   loss_keypoint=loss)
    assert h.final_layer.padding == (0, 0)
    h = V.TopdownHeatmapSimpleHead(out_channels=3, in_channels=512, extra={'final_conv_kernel': 0},
                                   loss_keypoint=loss)
    assert isinstance(h.final_layer, torch.nn.Identity)


def test_keypoints_from_heatmaps_argument_checks():
    hm = np.ones((1, 1, 64, 64), dtype=np.float32)
    c, s = np.array([[127, 127]]), np.array([[0.32, 0.32]])
    with pytest.raises(AssertionError):
        V.keypoints_from_heatmaps(hm, c, s, post_process='unbiased', kernel=0)
    with pytest.raises(AssertionError):
        V.keypoints_from_heatmaps(hm, c, s, unbiased=True, post_process=None)
    with pytest.raises(AssertionError):
        V.keypoints_from_heatmaps(hm, c, s, use_udp=True, post_process='megvii')


def test_decode_mode_resolution():
    assert decode_mode_from_cfg(configs.TEST_CFG_UDP) == _lib.DECODE_UDP_DARK
    assert decode_mode_from_cfg(configs.TEST_CFG_SHIFT) == _lib.DECODE_DEFAULT
    assert resolve_decode_mode('unbiased', False, False) == _lib.DECODE_UNBIASED
    assert resolve_decode_mode('default', True, False) == _lib.DECODE_UNBIASED
    assert resolve_decode_mode(None, False, False) == _lib.DECODE_NONE
    assert resolve_decode_mode('none', False, False) == _lib.DECODE_DEFAULT   # any non-None string -> quarter offset
    assert resolve_decode_mode(None, False, True) == _lib.DECODE_UDP_DARK


def test_pack_deconv_weight_matches_conv_transpose():
    """The 4-phase 2x2 decomposition (host repack) reproduces ConvTranspose2d(k4,s2,p1) exactly in fp32."""
    g = torch.Generator().manual_seed(0)
    cin, cout, h, w = 8, 6, 5, 4
    x = torch.randn(2, cin, h, w, generator=g)
    wt = torch.randn(cin, cout, 4, 4, generator=g)
    ref = F.conv_transpose2d(x, wt, None, stride=2, padding=1)
    # evaluate the packed weights the way the kernel does (fp32 here; bf16 rounding switched off by reusing values)
    wp = pack_deconv_weight(wt.to(torch.bfloat16).float()).float()
    wt_r = wt.to(torch.bfloat16).float()
    ref = F.conv_transpose2d(x, wt_r, None, stride=2, padding=1)
    xp = F.pad(x, (1, 1, 1, 1))
    out = torch.zeros(2, cout, 2 * h, 2 * w)
    for py in range(2):
        for px in range(2):
            acc = torch.zeros(2, cout, h, w)
            for ty in range(2):
                for tx in range(2):
                    dy = 0 if ty == 0 else (-1 if py == 0 else 1)
                    dx = 0 if tx == 0 else (-1 if px == 0 else 1)
                    t = ty * 2 + tx
                    wk = wp[py * 2 + px][:, t * cin:(t + 1) * cin]           # [cout, cin]
                    xs = xp[:, :, 1 + dy:1 + dy + h, 1 + dx:1 + dx + w]
                    acc += torch.einsum('oc,nchw->nohw', wk, xs)
            out[:, :, py::2, px::2] = acc
    assert torch.allclose(out, ref, atol=1e-5)
    s, t = fold_bn(torch.tensor([2.0]), torch.tensor([0.5]), torch.tensor([1.0]), torch.tensor([3.0]), 1e-5)
    assert abs(s.item() - 2 / np.sqrt(3 + 1e-5)) < 1e-6 and abs(t.item() - (0.5 - s.item())) < 1e-6


def test_flip_index_matches_pairwise_swap():
    from vitpose_b200.core.post_processing import flip_index_from_pairs
    perm = flip_index_from_pairs(133, configs.WHOLEBODY133_FLIP_PAIRS)
    assert sorted(perm.tolist()) == list(range(133))
    assert (perm[perm] == np.arange(133)).all()
    assert perm[1] == 2 and perm[2] == 1 and perm[0] == 0 and perm[91] == 112


def test_joints_mse_loss_has_no_cpu_path(golden_dir):
    g = np.load(os.path.join(golden_dir, 'loss_kat.npz'))
    loss = V.build_loss(dict(type='JointsMSELoss', use_target_weight=True))
    o, t, w = (torch.from_numpy(g[k]) for k in ('output', 'target', 'weight'))
    with pytest.raises(_lib.VitposeLibError):
        loss(o, t, w)          # value parity vs the golden KATs is checked on the GPU (tests/test_training_ops.py)


def test_mae_checkpoint_adaptation(tmp_path):
    """mmcv_custom/checkpoint.py:361-395: a 224x224 MAE ViT (14x14 patch kernel, 14x14 position grid + cls token,
    `model` key, `module.` prefix) loaded into the 256x192 pose backbone: patch kernel zero-padded to 16x16, position
    tokens bicubically resized to 16x12, cls slot kept."""
    from vitpose_b200.checkpoint import adapt_state_dict, extract_state_dict
    cfg = configs.tiny_model_cfg(5)
    D = cfg['backbone']['embed_dim']
    g = torch.Generator().manual_seed(0)
    src = {'module.patch_embed.proj.weight': torch.randn(D, 3, 14, 14, generator=g),
           'module.patch_embed.proj.bias': torch.randn(D, generator=g),
           'module.pos_embed': torch.randn(1, 1 + 14 * 14, D, generator=g),
           'module.blocks.0.norm1.weight': torch.randn(D, generator=g)}
    path = os.path.join(str(tmp_path), 'mae.pth')
    torch.save({'model': src}, path)
    bb = V.build_backbone(dict(cfg['backbone'], patch_padding='pad'))
    bb.init_weights(pretrained=path)
    w = bb.patch_embed.proj.weight.detach()
    assert torch.equal(w[:, :, 1:15, 1:15], src['module.patch_embed.proj.weight'])
    assert float(w[:, :, 0].abs().max()) == 0 and float(w[:, :, 15].abs().max()) == 0
    pe = src['module.pos_embed']
    ref = F.interpolate(pe[:, 1:].reshape(1, 14, 14, D).permute(0, 3, 1, 2), size=(16, 12), mode='bicubic',
                        align_corners=False).permute(0, 2, 3, 1).flatten(1, 2)
    assert torch.equal(bb.pos_embed.detach()[:, :1], pe[:, :1])
    assert torch.allclose(bb.pos_embed.detach()[:, 1:], ref)
    assert torch.equal(bb.blocks[0].norm1.weight.detach(), src['module.blocks.0.norm1.weight'])
    # resize modes and the ViTPose+ expert split
    sd = extract_state_dict({'state_dict': {k[7:]: v for k, v in src.items()}})
    bic = adapt_state_dict(sd, bb, 'bicubic')['patch_embed.proj.weight']
    assert bic.shape[2:] == (16, 16) and torch.allclose(
        bic, F.interpolate(sd['patch_embed.proj.weight'], size=(16, 16), mode='bicubic', align_corners=False))


def _frozen_names(backbone):
    return sorted(n for n, p in backbone.named_parameters() if not p.requires_grad)


def test_freeze_stages_selection():
    """vit.py:249-284: frozen_stages / freeze_attn / freeze_ffn set requires_grad=False on the same tensors as the
    reference (including its loop from blocks[1]), survive .train(), and put frozen blocks in eval mode."""
    kw = dict(img_size=(256, 192), embed_dim=64, depth=4, num_heads=2, qkv_bias=True)
    bb = V.ViT(frozen_stages=2, **kw)
    fr = _frozen_names(bb)
    assert 'patch_embed.proj.weight' in fr and 'patch_embed.proj.bias' in fr
    assert all(n.startswith(('patch_embed.', 'blocks.1.', 'blocks.2.')) for n in fr)
    assert any(n.startswith('blocks.1.') for n in fr) and any(n.startswith('blocks.2.') for n in fr)
    assert not any(n.startswith('blocks.0.') for n in fr)          # the reference's range(1, frozen_stages + 1)
    bb.train()
    assert _frozen_names(bb) == fr
    assert not bb.blocks[1].training and not bb.blocks[2].training and bb.blocks[0].training and bb.blocks[3].training
    bb = V.ViT(freeze_attn=True, **kw).train()
    fr = _frozen_names(bb)
    assert fr and all('.attn.' in n or '.norm1.' in n for n in fr) and len(fr) == 4 * 6
    bb = V.ViT(freeze_ffn=True, **kw).train()
    fr = _frozen_names(bb)
    assert 'pos_embed' in fr and 'patch_embed.proj.weight' in fr
    assert all(n == 'pos_embed' or n.startswith('patch_embed.') or '.mlp.' in n or '.norm2.' in n for n in fr)
    assert _frozen_names(V.ViT(**kw)) == []
    # the layer-decay constructor leaves frozen tensors out, as the reference does (constructor :35-36)
    from vitpose_b200.optim import layer_decay_param_groups
    cfg = configs.tiny_model_cfg(5)
    cfg['backbone'].update(frozen_stages=1)
    model = V.build_posenet(cfg)
    names = [n for g in layer_decay_param_groups(model, 5e-4, 0.1, 2, 0.75) for n in g['param_names']]
    assert not any(n.startswith(('backbone.patch_embed', 'backbone.blocks.1.')) for n in names)
    assert any(n.startswith('backbone.blocks.0.') for n in names)


@pytest.mark.reference
def test_freeze_stages_equal_reference():
    from oracle import ref_loader
    ref = ref_loader.load_reference()
    for kw in (dict(frozen_stages=2), dict(frozen_stages=0), dict(freeze_attn=True), dict(freeze_ffn=True),
               dict(frozen_stages=1, freeze_ffn=True)):
        args = dict(img_size=(256, 192), embed_dim=64, depth=4, num_heads=2, qkv_bias=True, **kw)
        r, o = ref.ViT(**args), V.ViT(**args)
        r.train()          # (the reference's override returns None)
        o.train()
        assert _frozen_names(r) == _frozen_names(o), kw
        assert [b.training for b in r.blocks] == [b.training for b in o.blocks], kw
