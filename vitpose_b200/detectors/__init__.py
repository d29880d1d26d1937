from .top_down import TopDown
from .top_down_moe import TopDownMoE

__all__ = ['TopDown', 'TopDownMoE']
