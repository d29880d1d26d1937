"""TEST INFRASTRUCTURE — loads the UNMODIFIED reference (MiraPurkrabek/ViTPose) by file path.

Only usable in the authoring container, where ``/root/reference`` is mounted.  It is
used by ``oracle/make_golden.py`` (to generate ``tests/golden/*.npz``) and by the
``not gpu`` tests that pin the oracle restatement against the real reference code.
Nothing on the product path, in ``-m gpu`` tests, ``smoke()`` or ``bench.py`` may import it.

The reference cannot be ``import mmpose``-ed (``mmpose/__init__.py:2`` imports mmcv, which
is not installable offline), so the ten files on the hot path are executed by path under
stub parent packages; ``mmcv``/``timm`` are replaced by ~40 lines of shims that return the
stock torch modules the real builders return.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get('VITPOSE_REFERENCE_ROOT', '/root/reference')


# the reference files on the hot path (everything load_reference() executes), relative to the reference root
REFERENCE_FILES = (
    'mmpose/core/post_processing/post_transforms.py',
    'mmpose/core/evaluation/top_down_eval.py',
    'mmpose/models/builder.py',
    'mmpose/models/utils/ops.py',
    'mmpose/models/losses/mse_loss.py',
    'mmpose/models/backbones/base_backbone.py',
    'mmpose/models/backbones/vit.py',
    'mmpose/models/heads/topdown_heatmap_base_head.py',
    'mmpose/models/heads/topdown_heatmap_simple_head.py',
    'mmpose/models/detectors/base.py',
    'mmpose/models/detectors/top_down.py',
)


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'mmpose/models/backbones/vit.py'))


def stage_reference(dst, src='/root/reference'):
    """Copies the UNMODIFIED reference files of the hot path into ``dst`` (the git-ignored ``baseline/_ref``), so
    that ``bench.py --impl reference`` can run the real reference on a GPU box, where /root/reference does not
    exist. Called by ``__graft_entry__.build()`` in the authoring container. Returns the number of files copied."""
    import shutil
    n = 0
    for rel in REFERENCE_FILES:
        s = os.path.join(src, rel)
        if not os.path.isfile(s):
            return 0
        d = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if not os.path.isfile(d) or os.path.getmtime(d) < os.path.getmtime(s):
            shutil.copy2(s, d)
        n += 1
    return n


class _Registry:
    """Minimal stand-in for mmcv.utils.Registry (dict + build(cfg))."""

    def __init__(self, name, build_func=None, parent=None, scope=None):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, force=False, module=None):
        def _reg(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return _reg(module)
        return _reg

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg, default_args=None):
        cfg = dict(cfg)
        if default_args:
            for k, v in default_args.items():
                cfg.setdefault(k, v)
        typ = cfg.pop('type')
        cls = self.module_dict[typ] if isinstance(typ, str) else typ
        return cls(**cfg)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _load(dotted, relpath):
    path = os.path.join(REF_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(dotted, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[dotted] = mod
    spec.loader.exec_module(mod)
    return mod


_LOADED = None


def load_reference():
    """Returns a namespace with the reference's own classes/functions."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError(f'reference tree not found at {REF_ROOT}')
    import torch
    import torch.nn as nn

    def build_model_from_cfg(cfg, registry, default_args=None):
        return registry.build(cfg, default_args)

    def build_conv_layer(cfg, *args, **kwargs):
        return nn.Conv2d(*args, **kwargs)

    def build_norm_layer(cfg, num_features, postfix=''):
        return 'bn', nn.BatchNorm2d(num_features)

    def build_upsample_layer(cfg, *args, **kwargs):
        return nn.ConvTranspose2d(*args, **kwargs)

    def constant_init(module, val, bias=0):
        if getattr(module, 'weight', None) is not None:
            nn.init.constant_(module.weight, val)
        if getattr(module, 'bias', None) is not None:
            nn.init.constant_(module.bias, bias)

    def normal_init(module, mean=0, std=1, bias=0):
        if getattr(module, 'weight', None) is not None:
            nn.init.normal_(module.weight, mean, std)
        if getattr(module, 'bias', None) is not None:
            nn.init.constant_(module.bias, bias)

    def _identity_deco(*a, **k):
        def deco(f):
            return f
        return deco

    def drop_path(x, drop_prob=0., training=False):
        if drop_prob == 0. or not training:
            return x
        keep = 1 - drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
        mask.floor_()
        return x.div(keep) * mask

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    mmcv_models = _Registry('model')
    _stub('mmcv')
    _stub('mmcv.cnn', MODELS=mmcv_models, build_model_from_cfg=build_model_from_cfg,
          build_conv_layer=build_conv_layer, build_norm_layer=build_norm_layer,
          build_upsample_layer=build_upsample_layer, constant_init=constant_init,
          normal_init=normal_init)
    _stub('mmcv.utils', Registry=_Registry)
    _stub('mmcv.utils.misc', deprecated_api_warning=_identity_deco)
    _stub('mmcv.runner', auto_fp16=_identity_deco)
    _stub('mmcv.image', imwrite=None)
    _stub('mmcv.visualization')
    _stub('mmcv.visualization.image', imshow=None)
    _stub('timm')
    _stub('timm.models')
    _stub('timm.models.layers', drop_path=drop_path, to_2tuple=to_2tuple,
          trunc_normal_=nn.init.trunc_normal_)
    _stub('mmcv_custom')
    _stub('mmcv_custom.checkpoint', load_checkpoint=None)
    _stub('mmpose')
    core = _stub('mmpose.core', imshow_bboxes=None, imshow_keypoints=None)
    models = _stub('mmpose.models')
    for sub in ('backbones', 'heads', 'detectors', 'losses', 'utils'):
        _stub('mmpose.models.' + sub)

    post = _load('mmpose.core.post_processing', 'mmpose/core/post_processing/post_transforms.py')
    _stub('mmpose.core.evaluation')
    tde = _load('mmpose.core.evaluation.top_down_eval', 'mmpose/core/evaluation/top_down_eval.py')
    ev = sys.modules['mmpose.core.evaluation']
    ev.top_down_eval = tde
    ev.pose_pck_accuracy = tde.pose_pck_accuracy
    ev.keypoints_from_heatmaps = tde.keypoints_from_heatmaps
    core.post_processing = post
    core.evaluation = ev

    builder = _load('mmpose.models.builder', 'mmpose/models/builder.py')
    models.builder = builder
    _load('mmpose.models.utils.ops', 'mmpose/models/utils/ops.py')
    mse = _load('mmpose.models.losses.mse_loss', 'mmpose/models/losses/mse_loss.py')
    _load('mmpose.models.backbones.base_backbone', 'mmpose/models/backbones/base_backbone.py')
    vit = _load('mmpose.models.backbones.vit', 'mmpose/models/backbones/vit.py')
    _load('mmpose.models.heads.topdown_heatmap_base_head',
          'mmpose/models/heads/topdown_heatmap_base_head.py')
    head = _load('mmpose.models.heads.topdown_heatmap_simple_head',
                 'mmpose/models/heads/topdown_heatmap_simple_head.py')
    _load('mmpose.models.detectors.base', 'mmpose/models/detectors/base.py')
    td = _load('mmpose.models.detectors.top_down', 'mmpose/models/detectors/top_down.py')

    ns = types.SimpleNamespace(
        builder=builder, ViT=vit.ViT, TopdownHeatmapSimpleHead=head.TopdownHeatmapSimpleHead,
        TopDown=td.TopDown, JointsMSELoss=mse.JointsMSELoss,
        keypoints_from_heatmaps=tde.keypoints_from_heatmaps, _get_max_preds=tde._get_max_preds,
        post_dark_udp=tde.post_dark_udp, _gaussian_blur=tde._gaussian_blur, _taylor=tde._taylor,
        flip_back=post.flip_back, transform_preds=post.transform_preds,
        pose_pck_accuracy=tde.pose_pck_accuracy)
    _LOADED = ns
    return ns


def build_reference_topdown(model_cfg):
    """model_cfg: dict with backbone / keypoint_head / test_cfg (same fields as the config's
    ``model=`` block, e.g. configs/body/.../ViTPose_base_coco_256x192.py:52-84)."""
    import copy
    ref = load_reference()
    cfg = copy.deepcopy(model_cfg)
    cfg.pop('type', None)
    cfg.pop('pretrained', None)
    cfg.setdefault('train_cfg', dict())
    cfg['backbone'] = dict(cfg['backbone'])
    cfg['keypoint_head'] = dict(cfg['keypoint_head'])
    m = ref.TopDown(**cfg)
    m.eval()
    return m


def build_reference_topdown_moe(model_cfg):
    """The reference's ViTPose+ detector (mmpose/models/detectors/top_down_moe.py:15-92) over its ViTMoE backbone
    (mmpose/models/backbones/vit_moe.py), from a ``model=`` block with ``associate_keypoint_head``."""
    import copy
    load_reference()
    if 'mmpose.models.backbones.vit_moe' not in sys.modules:
        _load('mmpose.models.backbones.vit_moe', 'mmpose/models/backbones/vit_moe.py')
    if 'mmpose.models.detectors.top_down_moe' not in sys.modules:
        _load('mmpose.models.detectors.top_down_moe', 'mmpose/models/detectors/top_down_moe.py')
    cfg = copy.deepcopy(model_cfg)
    cfg.pop('type', None)
    cfg.pop('pretrained', None)
    cfg.setdefault('train_cfg', dict())
    return sys.modules['mmpose.models.detectors.top_down_moe'].TopDownMoE(**cfg)
