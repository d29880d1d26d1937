"""``ViT`` backbone with the reference's registry name, constructor signature and state-dict layout
(mmpose/models/backbones/vit.py:200-341), executed by the sm_100a kernels behind vpb_vitpose_forward.

The nn.Module tree below only *holds parameters* under the reference's names
(``patch_embed.proj``, ``pos_embed``, ``blocks.{i}.norm1/attn.qkv/attn.proj/norm2/mlp.fc1/mlp.fc2``,
``last_norm``) so reference checkpoints load with ``load_state_dict``; none of these modules is ever called.
There is no eager fallback: ``forward`` needs a CUDA device and the built library.
"""
import math
from functools import partial

import torch
import torch.nn as nn

from ..builder import BACKBONES
from ..engine import VitPoseEngine


def _to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


class _ParamHolder(nn.Module):
    def forward(self, *a, **k):   # pragma: no cover - guards against accidental eager use
        raise RuntimeError('vitpose_b200 modules hold parameters only; compute runs in libvitpose_b200.so')


class _Attention(_ParamHolder):
    def __init__(self, dim, num_heads, qkv_bias):
        super().__init__()
        self.num_heads = num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class _Mlp(_ParamHolder):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(_ParamHolder):
    def __init__(self, dim, num_heads, mlp_ratio, qkv_bias, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _Attention(dim, num_heads, qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class _PatchEmbed(_ParamHolder):
    def __init__(self, img_size, patch_size, in_chans, embed_dim, ratio):
        super().__init__()
        img_size, patch_size = _to_2tuple(img_size), _to_2tuple(patch_size)
        self.patch_shape = (int(img_size[0] // patch_size[0] * ratio), int(img_size[1] // patch_size[1] * ratio))
        self.origin_patch_shape = (int(img_size[0] // patch_size[0]), int(img_size[1] // patch_size[1]))
        self.img_size, self.patch_size = img_size, patch_size
        self.num_patches = (img_size[1] // patch_size[1]) * (img_size[0] // patch_size[0]) * (ratio ** 2)
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=(patch_size[0] // ratio),
                              padding=4 + 2 * (ratio // 2 - 1))


@BACKBONES.register_module()
class ViT(nn.Module):

    def __init__(self,
                 img_size=224, patch_size=16, in_chans=3, num_classes=80, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0.,
                 drop_path_rate=0., hybrid_backbone=None, norm_layer=None, use_checkpoint=False,
                 frozen_stages=-1, ratio=1, last_norm=True,
                 patch_padding='pad', freeze_attn=False, freeze_ffn=False,
                 ):
        super().__init__()
        if hybrid_backbone is not None:
            raise NotImplementedError('hybrid_backbone is not used by any ViTPose config')
        if qk_scale is not None:
            raise NotImplementedError('qk_scale override is not used by any ViTPose config')
        if in_chans != 3:
            raise NotImplementedError('the patch-embed kernel takes RGB crops (in_chans=3)')
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.frozen_stages = frozen_stages
        self.use_checkpoint = use_checkpoint
        self.patch_padding = patch_padding
        self.freeze_attn = freeze_attn
        self.freeze_ffn = freeze_ffn
        self.depth = depth
        self.num_heads = num_heads
        self.drop_path_rate = drop_path_rate   # inference: DropPath is the identity (vit.py:55-56)
        self._cfg = dict(img_size=_to_2tuple(img_size), patch_size=patch_size, embed_dim=embed_dim, depth=depth,
                         num_heads=num_heads, mlp_ratio=mlp_ratio, ratio=ratio, last_norm=last_norm)

        self.patch_embed = _PatchEmbed(img_size, patch_size, in_chans, embed_dim, ratio)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + 1, embed_dim))
        self.blocks = nn.ModuleList([_Block(embed_dim, num_heads, mlp_ratio, qkv_bias, norm_layer)
                                     for _ in range(depth)])
        self.last_norm = norm_layer(embed_dim) if last_norm else nn.Identity()
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        self._engine = None
        self._engine_key = None
        if use_checkpoint:
            # activation checkpointing (vit.py:323-324) only trades memory for recompute; the training step here
            # always keeps the activations it needs, so the flag changes nothing — say so instead of ignoring it
            import warnings
            warnings.warn('vitpose_b200: use_checkpoint=True has no effect (activations are kept; results identical)')
        self._freeze_stages()

    def _freeze_stages(self):
        """vit.py:249-284, same selection rules (including the reference's loop ``range(1, frozen_stages + 1)``, which
        leaves blocks[0] trainable): frozen tensors get ``requires_grad = False`` — the backward pass then skips their
        weight gradients and the layer-decay optimizer leaves them out — and frozen blocks are put in eval mode, which
        switches their stochastic depth off (``DropPath`` is a child of ``Block``)."""
        def freeze(module):
            module.eval()
            for param in module.parameters():
                param.requires_grad = False

        if self.frozen_stages >= 0:
            freeze(self.patch_embed)
        for i in range(1, self.frozen_stages + 1):
            freeze(self.blocks[i])
        if self.freeze_attn:
            for blk in self.blocks:
                freeze(blk.attn)
                freeze(blk.norm1)
        if self.freeze_ffn:
            self.pos_embed.requires_grad = False
            freeze(self.patch_embed)
            for blk in self.blocks:
                freeze(blk.mlp)
                freeze(blk.norm2)

    # ---- reference API ------------------------------------------------------------------------------
    def init_weights(self, pretrained=None):
        """vit.py:286-304: trunc-normal(0.02) Linear weights, zero biases, LayerNorm 1/0."""
        if pretrained is not None:
            # base_backbone.py:init_weights -> mmcv_custom load_checkpoint(self, pretrained, strict=False,
            # patch_padding=self.patch_padding): MAE-pretrain adaptation of patch / position embeddings
            from ..checkpoint import adapt_state_dict, extract_state_dict
            sd = extract_state_dict(torch.load(pretrained, map_location='cpu'))
            sd = {k[len('backbone.'):] if k.startswith('backbone.') else k: v for k, v in sd.items()}
            # ViTPose+ (vit_moe.py:336): part_features splits every MAE fc2 [D, 4D] into the shared fc2 and the experts
            self.load_state_dict(adapt_state_dict(sd, self, self.patch_padding, getattr(self, 'part_features', None)),
                                 strict=False)
            self._engine = None
            return
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)
        self._engine = None

    def get_num_layers(self):
        return len(self.blocks)

    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    # ---- execution -----------------------------------------------------------------------------------
    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self, head=None):
        """Packed-weight engine for this backbone (+ optional head); rebuilt when parameters change."""
        key = (self._weights_version(), None if head is None else head._weights_version())
        if self._engine is None or self._engine_key != key:
            sd = {'backbone.' + k: v for k, v in self.state_dict().items()}
            head_cfg = None
            if head is not None:
                sd.update({'keypoint_head.' + k: v for k, v in head.state_dict().items()})
                head_cfg = head.cfg_dict()
            dev = next(self.parameters()).device
            if dev.type != 'cuda':
                raise RuntimeError('vitpose_b200 has no CPU path: move the model to a CUDA device (model.cuda())')
            self._engine = VitPoseEngine(self._cfg, head_cfg, sd, device=dev)
            self._engine_key = key
        return self._engine

    def forward_features(self, x):
        eng = self.engine()
        _, tokens = eng.forward_heatmaps(x.float(), flip=False, want_features=True, want_heatmaps=False)
        from .. import ops
        hp, wp = eng.tokens_hw
        return ops.tokens_to_nchw(tokens, hp, wp)        # [N, D, Hp, Wp] fp32, as vit.py:330

    def forward(self, x):
        return self.forward_features(x)

    def train(self, mode=True):
        """vit.py:338-341: frozen parts stay frozen (and in eval mode) whenever the mode is switched."""
        super().train(mode)
        self._freeze_stages()
        return self
