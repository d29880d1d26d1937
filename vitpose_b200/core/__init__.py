from .evaluation import keypoints_from_heatmaps
from .post_processing import flip_back, transform_preds

__all__ = ['keypoints_from_heatmaps', 'flip_back', 'transform_preds']
