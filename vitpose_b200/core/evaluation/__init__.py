from .top_down_eval import _get_max_preds, keypoints_from_heatmaps

__all__ = ['keypoints_from_heatmaps', '_get_max_preds']
