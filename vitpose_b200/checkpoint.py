"""MAE-pretrain / cross-resolution checkpoint adaptation for the ViT backbone: what the
reference's ``load_checkpoint`` does before ``load_state_dict`` (mmcv_custom/checkpoint.py:312-409, called from
``BaseBackbone.init_weights`` with the backbone's ``patch_padding``, vit.py:292):

* pick ``state_dict`` / ``model`` / ``module`` out of the file, strip a ``module.`` prefix, unwrap MoBY's
  ``encoder.`` branch (:341-357);
* patch embedding trained with a smaller kernel (MAE ViT: 14x14) -> this model's kernel: zero-pad symmetrically
  ('pad'), or resize ('bilinear' / 'bicubic') (:361-375);
* positional embedding trained on a square grid -> this model's (H, W) patch grid by bicubic interpolation of the
  position tokens, extra (cls) tokens kept (:377-395);
* ViTPose+ ``part_features``: the last ``part_features`` output rows of every ``mlp.fc2`` become the experts'
  weights, the rest stays in ``fc2`` (:397-405).

One-time host-side work on CPU tensors (exactly the torch ops the reference runs); nothing here is on the hot path.
"""
import re

import torch
import torch.nn.functional as F


def extract_state_dict(checkpoint):
    if not isinstance(checkpoint, dict):
        raise RuntimeError('No state_dict found in checkpoint')
    for key in ('state_dict', 'model', 'module'):
        if key in checkpoint:
            sd = checkpoint[key]
            break
    else:
        sd = checkpoint
    if list(sd.keys())[0].startswith('module.'):
        sd = {k[7:]: v for k, v in sd.items()}
    if sorted(list(sd.keys()))[0].startswith('encoder'):
        sd = {k.replace('encoder.', ''): v for k, v in sd.items() if k.startswith('encoder.')}
    return sd


def adapt_state_dict(state_dict, model, patch_padding='pad', part_features=None):
    """Returns a new state dict for ``model`` (a ViT backbone: ``patch_embed.proj``, ``patch_embed.patch_shape``,
    ``patch_embed.num_patches``, ``pos_embed``)."""
    sd = dict(state_dict)
    if 'patch_embed.proj.weight' in sd:
        w = sd['patch_embed.proj.weight']
        orig, cur = tuple(w.shape[2:]), tuple(model.patch_embed.proj.weight.shape[2:])
        if orig != cur:
            pad = cur[0] - orig[0]
            left = pad // 2
            right = pad - left
            if 'pad' in patch_padding:
                w = F.pad(w, (left, right, left, right))
            elif 'bilinear' in patch_padding:
                w = F.interpolate(w, size=cur, mode='bilinear', align_corners=False)
            elif 'bicubic' in patch_padding:
                w = F.interpolate(w, size=cur, mode='bicubic', align_corners=False)
            sd['patch_embed.proj.weight'] = w
    if 'pos_embed' in sd:
        pe = sd['pos_embed']
        D = pe.shape[-1]
        H, W = model.patch_embed.patch_shape
        extra = model.pos_embed.shape[-2] - model.patch_embed.num_patches
        side = int((pe.shape[-2] - extra) ** 0.5)
        tokens = pe[:, extra:].reshape(-1, side, side, D).permute(0, 3, 1, 2)
        tokens = F.interpolate(tokens, size=(H, W), mode='bicubic', align_corners=False)
        sd['pos_embed'] = torch.cat((pe[:, :extra], tokens.permute(0, 2, 3, 1).flatten(1, 2)), dim=1)
    if part_features is not None:
        out = dict(sd)
        for key in model.state_dict().keys():
            if 'mlp.experts' in key:
                out[key] = sd[re.sub(r'experts.\d+.', 'fc2.', key)][-part_features:]
            elif 'fc2' in key:
                out[key] = sd[key][:-part_features]
        sd = out
    return sd


def load_checkpoint(model, filename, map_location='cpu', strict=False, patch_padding='pad', part_features=None):
    """Reference call convention (mmcv_custom/checkpoint.py:312-319) for local files; returns the raw checkpoint."""
    checkpoint = torch.load(filename, map_location=map_location)
    sd = adapt_state_dict(extract_state_dict(checkpoint), model, patch_padding, part_features)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    if strict and (missing or unexpected):
        raise RuntimeError(f'missing keys {missing}, unexpected keys {unexpected}')
    return checkpoint


COCO_PLUS_DATASETS = ('coco', 'aic', 'mpii', 'ap10k', 'apt36k', 'wholebody')
COCO_PLUS_KEYPOINTS = (17, 14, 16, 17, 17, 133)


def split_moe_state_dict(state_dict, dataset_idx, num_keypoints=None):
    """ViTPose+ multi-dataset state dict -> the plain ``TopDown`` / ``ViT`` state dict of one dataset, as the
    reference's tools/model_split.py writes it: every ``mlp.fc2`` becomes ``cat(fc2, experts[dataset_idx])`` on the
    output dimension (:36-40, :70-79); for ``dataset_idx`` > 0 the tensors of ``associate_keypoint_heads[idx-1]``
    replace the main head's (:83-84) and the final layer is cut to the dataset's keypoint count (:86-87)."""
    sd = dict(state_dict)
    out = {}
    for k, v in sd.items():
        if 'mlp.experts' in k or k.startswith('associate_keypoint_heads.'):
            continue
        if 'mlp.fc2' in k:
            v = torch.cat([v, sd[k.replace('fc2.', f'experts.{dataset_idx}.')]], dim=0)
        out[k] = v
    if dataset_idx > 0:
        pre = f'associate_keypoint_heads.{dataset_idx - 1}.'
        for k, v in sd.items():
            if k.startswith(pre):
                out['keypoint_head.' + k[len(pre):]] = v
        nk = COCO_PLUS_KEYPOINTS[dataset_idx] if num_keypoints is None else num_keypoints
        for k in ('keypoint_head.final_layer.weight', 'keypoint_head.final_layer.bias'):
            out[k] = out[k][:nk]
    return out
