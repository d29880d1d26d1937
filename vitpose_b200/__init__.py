"""vitpose_b200 — B200-native (sm_100a) ViTPose top-down inference hot path behind the reference's own
registry / function boundary.  Importing the package registers ``ViT``, ``TopdownHeatmapSimpleHead``,
``TopDown`` and ``JointsMSELoss`` in :mod:`vitpose_b200.builder`; all compute is in libvitpose_b200.so."""
from . import builder
from .backbones import ViT, ViTMoE
from .builder import build_backbone, build_head, build_loss, build_posenet
from .core import flip_back, keypoints_from_heatmaps, transform_preds
from .detectors import TopDown, TopDownMoE
from .heads import TopdownHeatmapSimpleHead
from .losses import JointsMSELoss

__version__ = '0.1.0'
__all__ = ['builder', 'ViT', 'ViTMoE', 'TopDownMoE', 'TopdownHeatmapSimpleHead', 'TopDown', 'JointsMSELoss', 'build_backbone',
           'build_head', 'build_loss', 'build_posenet', 'keypoints_from_heatmaps', 'flip_back', 'transform_preds']
