from .topdown_heatmap_simple_head import TopdownHeatmapBaseHead, TopdownHeatmapSimpleHead

__all__ = ['TopdownHeatmapBaseHead', 'TopdownHeatmapSimpleHead']
