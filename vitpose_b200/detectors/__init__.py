from .top_down import TopDown

__all__ = ['TopDown']
