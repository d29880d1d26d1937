"""Registry boundary, mirroring mmpose/models/builder.py:85-123: config dicts with a ``type`` key are
resolved to classes registered under the reference's names (``ViT``, ``TopdownHeatmapSimpleHead``,
``TopDown``, ``JointsMSELoss``), so the ``model = dict(...)`` block of a ViTPose config builds unchanged.

If a real mmpose/mmcv is importable, :func:`register_into_mmpose` drops these classes into mmpose's own
``MODELS`` registry with ``force=True`` — the drop-in a maintainer would use.
"""
import copy


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key)

    def __contains__(self, key):
        return key in self._module_dict

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            key = name or cls.__name__
            if not force and key in self._module_dict:
                raise KeyError(f'{key} is already registered in {self._name}')
            self._module_dict[key] = cls
            return cls
        if module is not None:
            return _register(module)
        return _register

    def build(self, cfg, default_args=None):
        if not isinstance(cfg, dict):
            raise TypeError(f'cfg must be a dict, but got {type(cfg)}')
        if 'type' not in cfg and not (default_args and 'type' in default_args):
            raise KeyError(f'`cfg` or `default_args` must contain the key "type", but got {cfg}')
        args = copy.copy(cfg)
        if default_args is not None:
            for k, v in default_args.items():
                args.setdefault(k, v)
        obj_type = args.pop('type')
        if isinstance(obj_type, str):
            cls = self.get(obj_type)
            if cls is None:
                raise KeyError(f'{obj_type} is not in the {self._name} registry')
        elif isinstance(obj_type, type):
            cls = obj_type
        else:
            raise TypeError(f'type must be a str or valid type, but got {type(obj_type)}')
        return cls(**args)


MODELS = Registry('models')
BACKBONES = MODELS
NECKS = MODELS
HEADS = MODELS
LOSSES = MODELS
POSENETS = MODELS


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_neck(cfg):
    return NECKS.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_posenet(cfg):
    return POSENETS.build(cfg)


def register_into_mmpose():
    """Overrides the reference's registered classes with the B200 ones (needs a working mmpose install)."""
    from mmpose.models.builder import MODELS as MM_MODELS   # noqa: raises ImportError without mmpose/mmcv
    for name, cls in MODELS.module_dict.items():
        MM_MODELS.register_module(name=name, force=True, module=cls)
    return sorted(MODELS.module_dict)
