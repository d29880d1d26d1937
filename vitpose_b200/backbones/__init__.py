from .vit import ViT
from .vit_moe import ViTMoE

__all__ = ['ViT', 'ViTMoE']
