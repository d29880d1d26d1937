"""Synthetic inputs and weights for parity tests and the benchmark.

No datasets or checkpoints are available offline, so crops are ``randn`` (≈ ImageNet-normalised
statistics), boxes follow ``_box2cs`` aspect 0.75 (mmpose/apis/inference.py:99-112), and weights
are random.  ``scaled_init_state_dict`` draws a full reference-layout state dict whose head
produces heatmaps of realistic amplitude (std ≈ 0.1), because the reference ``init_weights()``
(head std 0.001) gives ~1e-5-magnitude heatmaps on which absolute tolerances are vacuous.
"""
import math

import numpy as np
import torch

from .configs import flip_pairs_for


def synthetic_crops(n, seed=0, height=256, width=192):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 3, height, width, generator=g)


def synthetic_metas(n, num_keypoints, seed=0):
    rng = np.random.RandomState(seed + 1)
    pairs = flip_pairs_for(num_keypoints)
    metas = []
    for i in range(n):
        center = np.array([96.0, 128.0], dtype=np.float32) + rng.uniform(-20, 20, 2).astype(np.float32)
        scale = (np.array([192.0, 256.0], dtype=np.float32) / 200.0 * 1.25 *
                 np.float32(rng.uniform(0.8, 1.2)))
        metas.append(dict(center=center, scale=scale.astype(np.float32), rotation=0,
                          image_file='', bbox_score=1.0, bbox_id=i, flip_pairs=pairs))
    return metas


def gaussian_peak_heatmaps(n, k, seed=0, height=64, width=48, sigma=2.0, noise=0.01,
                           border_peaks=True):
    """Decode-only parity set: one Gaussian peak per map (centres include borders/corners)
    plus N(0, noise) noise, float32 [n, k, H, W]."""
    rng = np.random.RandomState(seed)
    lo = 0.0 if border_peaks else 4.0
    cx = rng.uniform(lo, width - 1 - lo, size=(n, k, 1, 1))
    cy = rng.uniform(lo, height - 1 - lo, size=(n, k, 1, 1))
    amp = rng.uniform(0.2, 1.0, size=(n, k, 1, 1))
    ys = np.arange(height).reshape(1, 1, height, 1)
    xs = np.arange(width).reshape(1, 1, 1, width)
    hm = amp * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * sigma * sigma))
    hm = hm + rng.normal(0, noise, size=hm.shape)
    return hm.astype(np.float32)


def state_dict_shapes(model_cfg):
    """Reference-layout parameter/buffer shapes (SURVEY.md §8b 'Checkpoint compatibility')."""
    bb, hd = model_cfg['backbone'], model_cfg['keypoint_head']
    D, L = bb['embed_dim'], bb['depth']
    Hp, Wp = bb['img_size'][0] // 16, bb['img_size'][1] // 16
    hidden = int(D * bb.get('mlp_ratio', 4))
    s = {'backbone.pos_embed': (1, Hp * Wp + 1, D),
         'backbone.patch_embed.proj.weight': (D, 3, 16, 16),
         'backbone.patch_embed.proj.bias': (D,)}
    for i in range(L):
        b = f'backbone.blocks.{i}.'
        s.update({b + 'norm1.weight': (D,), b + 'norm1.bias': (D,),
                  b + 'attn.qkv.weight': (3 * D, D), b + 'attn.qkv.bias': (3 * D,),
                  b + 'attn.proj.weight': (D, D), b + 'attn.proj.bias': (D,),
                  b + 'norm2.weight': (D,), b + 'norm2.bias': (D,),
                  b + 'mlp.fc1.weight': (hidden, D), b + 'mlp.fc1.bias': (hidden,),
                  b + 'mlp.fc2.weight': (D, hidden), b + 'mlp.fc2.bias': (D,)})
    s.update({'backbone.last_norm.weight': (D,), 'backbone.last_norm.bias': (D,)})
    c_in = hd['in_channels']
    for i in range(hd.get('num_deconv_layers', 3)):
        c_out = hd['num_deconv_filters'][i]
        k = hd['num_deconv_kernels'][i]
        p = f'keypoint_head.deconv_layers.{3 * i}.'
        q = f'keypoint_head.deconv_layers.{3 * i + 1}.'
        s.update({p + 'weight': (c_in, c_out, k, k), q + 'weight': (c_out,), q + 'bias': (c_out,),
                  q + 'running_mean': (c_out,), q + 'running_var': (c_out,)})
        c_in = c_out
    fk = (hd.get('extra') or {}).get('final_conv_kernel', 1)
    if fk:
        s.update({'keypoint_head.final_layer.weight': (hd['out_channels'], c_in, fk, fk),
                  'keypoint_head.final_layer.bias': (hd['out_channels'],)})
    return s


def scaled_init_state_dict(model_cfg, seed=0, heatmap_std=0.1):
    """Random weights with trained-network-like statistics: transformer Linear ~ N(0, 0.02²)
    scaled up so activations stay O(1) through depth, LayerNorm/BN affine ≈ 1 ± 0.1, non-zero
    biases, BN running stats randomised, head fan-in scaled so heatmap std ≈ ``heatmap_std``."""
    g = torch.Generator().manual_seed(seed)
    shapes = state_dict_shapes(model_cfg)
    sd = {}

    def randn(shape, std):
        return torch.randn(*shape, generator=g) * std

    for name, shape in shapes.items():
        leaf = name.rsplit('.', 1)[-1]
        if name.endswith('pos_embed'):
            t = randn(shape, 0.02)
        elif 'norm' in name and leaf == 'weight':
            t = 1.0 + randn(shape, 0.1)
        elif 'norm' in name and leaf == 'bias':
            t = randn(shape, 0.05)
        elif leaf == 'running_mean':
            t = randn(shape, 0.1)
        elif leaf == 'running_var':
            t = 0.5 + torch.rand(*shape, generator=g)
        elif 'deconv_layers' in name and len(shape) == 1:          # BN affine
            t = (1.0 + randn(shape, 0.1)) if leaf == 'weight' else randn(shape, 0.1)
        elif 'deconv_layers' in name:                               # ConvTranspose [Cin,Cout,4,4]
            fan_in = shape[0] * 4                                   # 2x2 taps reach each output
            t = randn(shape, math.sqrt(2.0 / fan_in))
        elif name.endswith('final_layer.weight'):
            fan_in = shape[1] * shape[2] * shape[3]
            t = randn(shape, heatmap_std * 1.6 / math.sqrt(fan_in))
        elif name.endswith('final_layer.bias'):
            t = randn(shape, 0.01)
        elif leaf == 'bias':
            t = randn(shape, 0.02)
        elif name.endswith('patch_embed.proj.weight'):
            t = randn(shape, 1.0 / math.sqrt(768))
        else:                                                       # transformer Linear weights
            t = randn(shape, 1.0 / math.sqrt(shape[1]))
        sd[name] = t.float().contiguous()
    for name in list(sd):
        if name.endswith('running_mean'):
            sd[name.replace('running_mean', 'num_batches_tracked')] = torch.tensor(0)
    return sd
