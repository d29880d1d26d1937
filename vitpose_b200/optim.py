"""Layer-decay AdamW of the ViTPose training configs.

* :func:`layer_decay_param_groups` — the grouping rule of ``LayerDecayOptimizerConstructor.add_params``
  (mmcv_custom/layer_decay_optimizer_constructor.py:6-78): ``layer_id`` 0 for pos_embed / patch_embed,
  ``i + 1`` for ``backbone.blocks.i``, ``num_layers + 1`` otherwise; 1-D tensors, ``.bias`` and ``pos_embed`` get
  weight decay 0; ``lr = base_lr * rate ** (num_layers + 2 - layer_id - 1)``.
* :class:`LayerDecayAdamW` — AdamW whose update (and the global-norm gradient clip of
  ``optimizer_config = dict(grad_clip=dict(max_norm=1.))``) runs in the vpb_adamw_step / vpb_grad_sq_norm kernels.
"""
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


def get_num_layer_for_vit(var_name, num_max_layer):
    if var_name in ('backbone.cls_token', 'backbone.mask_token', 'backbone.pos_embed'):
        return 0
    elif var_name.startswith('backbone.patch_embed'):
        return 0
    elif var_name.startswith('backbone.blocks'):
        return int(var_name.split('.')[2]) + 1
    return num_max_layer - 1


def layer_decay_param_groups(module, base_lr, weight_decay, num_layers, layer_decay_rate):
    """Returns the list of param-group dicts the reference constructor would hand to AdamW."""
    groups = {}
    num_layers = num_layers + 2
    for name, param in module.named_parameters():
        if not param.requires_grad:
            continue
        if len(param.shape) == 1 or name.endswith('.bias') or 'pos_embed' in name:
            kind, wd = 'no_decay', 0.
        else:
            kind, wd = 'decay', weight_decay
        layer_id = get_num_layer_for_vit(name, num_layers)
        key = 'layer_%d_%s' % (layer_id, kind)
        if key not in groups:
            scale = layer_decay_rate ** (num_layers - layer_id - 1)
            groups[key] = dict(weight_decay=wd, params=[], param_names=[], lr_scale=scale, group_name=key,
                               lr=scale * base_lr)
        groups[key]['params'].append(param)
        groups[key]['param_names'].append(name)
    return list(groups.values())


class LayerDecayOptimizerConstructor:
    """Same call convention as the mmcv constructor: ``LayerDecayOptimizerConstructor(optimizer_cfg,
    paramwise_cfg)(model)`` with ``optimizer_cfg = dict(type='AdamW', lr=..., betas=..., weight_decay=...)`` and
    ``paramwise_cfg = dict(num_layers=12, layer_decay_rate=0.75, ...)`` (ViTPose_base_coco_256x192.py:16-28)."""

    def __init__(self, optimizer_cfg, paramwise_cfg=None):
        if not isinstance(optimizer_cfg, dict):
            raise TypeError('optimizer_cfg should be a dict')
        self.optimizer_cfg = dict(optimizer_cfg)
        self.paramwise_cfg = {} if paramwise_cfg is None else paramwise_cfg
        self.base_lr = optimizer_cfg.get('lr', None)
        self.base_wd = optimizer_cfg.get('weight_decay', None)

    def __call__(self, model):
        if hasattr(model, 'module'):
            model = model.module
        cfg = dict(self.optimizer_cfg)
        typ = cfg.pop('type', 'AdamW')
        if typ != 'AdamW':
            raise NotImplementedError('ViTPose configs train with AdamW')
        groups = layer_decay_param_groups(model, self.base_lr, self.base_wd, self.paramwise_cfg.get('num_layers'),
                                          self.paramwise_cfg.get('layer_decay_rate'))
        cfg.pop('lr', None)
        cfg.pop('weight_decay', None)
        return LayerDecayAdamW(groups, lr=self.base_lr, weight_decay=self.base_wd, **cfg)


class LayerDecayAdamW(torch.optim.Optimizer):
    """AdamW with per-group lr / weight decay; ``step(max_norm=...)`` also applies clip_grad_norm_ semantics."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._sq_norm = None

    _ENTRY = [('param', '<u8'), ('grad', '<u8'), ('exp_avg', '<u8'), ('exp_avg_sq', '<u8'), ('n', '<i8'),
              ('lr', '<f4'), ('weight_decay', '<f4'), ('step', '<i4'), ('pad', '<i4')]     # vpb_tensor_entry
    _CHUNK = 4096

    @torch.no_grad()
    def step(self, max_norm=None):
        """One multi-tensor launch for the clip norm and one for the update (vpb_adamw_multi): the descriptor table
        (pointers, sizes, per-group lr / weight decay, step counts) is rebuilt on the host and uploaded each call
        because autograd hands out fresh gradient tensors every step."""
        import numpy as np
        L = lib()
        items = [(g, p) for g in self.param_groups for p in g['params'] if p.grad is not None]
        if not items:
            return None
        dev = items[0][1].device
        if dev.type != 'cuda':
            raise _lib.VitposeLibError('LayerDecayAdamW runs on CUDA parameters only (no CPU fallback)')
        b1, b2 = self.param_groups[0]['betas']
        eps = self.param_groups[0]['eps']
        if any(g['betas'] != (b1, b2) or g['eps'] != eps for g in self.param_groups):
            raise NotImplementedError('per-group betas / eps')
        table = np.zeros(len(items), dtype=self._ENTRY)
        starts = np.zeros(len(items) + 1, dtype=np.int32)
        keep = []
        for i, (group, p) in enumerate(items):
            st = self.state[p]
            if not st:
                st['step'] = 0
                st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st['step'] += 1
            g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
            if g.dtype != torch.float32 or p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.VitposeLibError('LayerDecayAdamW expects contiguous fp32 parameters and gradients')
            keep.append(g)
            table[i] = (p.data_ptr(), g.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(), p.numel(),
                        group['lr'], group['weight_decay'], st['step'], 0)
            starts[i + 1] = starts[i] + (p.numel() + self._CHUNK - 1) // self._CHUNK
        d_table = torch.from_numpy(table.view(np.uint8)).to(dev)
        d_starts = torch.from_numpy(starts).to(dev)
        sq = None
        if max_norm is not None:
            need = 1 + int(starts[-1])
            if self._sq_norm is None or self._sq_norm.device != dev or self._sq_norm.numel() < need:
                self._sq_norm = torch.zeros(need, device=dev, dtype=torch.float32)
            sq = self._sq_norm
        check(L.vpb_adamw_multi(ptr(d_table), ptr(d_starts), len(items), int(starts[-1]), float(b1), float(b2),
                                float(eps), ptr(sq), float(max_norm) if max_norm is not None else 0.0, stream_ptr()),
              'vpb_adamw_multi')
        return torch.sqrt(sq[0]) if sq is not None else None

    @torch.no_grad()
    def step_per_tensor(self, max_norm=None):
        """The same update with one vpb_grad_sq_norm_accumulate + vpb_adamw_step call per tensor."""
        L = lib()
        params = [p for g in self.param_groups for p in g['params'] if p.grad is not None]
        if not params:
            return None
        dev = params[0].device
        if dev.type != 'cuda':
            raise _lib.VitposeLibError('LayerDecayAdamW runs on CUDA parameters only (no CPU fallback)')
        sq = None
        if max_norm is not None:
            if self._sq_norm is None or self._sq_norm.device != dev:
                self._sq_norm = torch.zeros(1, device=dev, dtype=torch.float32)
            self._sq_norm.zero_()
            for p in params:
                g = p.grad.contiguous()
                check(L.vpb_grad_sq_norm_accumulate(ptr(g), g.numel(), ptr(self._sq_norm), stream_ptr()),
                      'vpb_grad_sq_norm_accumulate')
            sq = self._sq_norm
        for group in self.param_groups:
            b1, b2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st['step'] += 1
                g = p.grad.contiguous()
                check(L.vpb_adamw_step(ptr(p.data), ptr(g), ptr(st['exp_avg']), ptr(st['exp_avg_sq']), p.numel(),
                                       float(group['lr']), float(b1), float(b2), float(group['eps']),
                                       float(group['weight_decay']), int(st['step']), ptr(sq),
                                       float(max_norm) if max_norm is not None else 0.0, stream_ptr()),
                      'vpb_adamw_step')
        return torch.sqrt(sq[0]) if sq is not None else None
