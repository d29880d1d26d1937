from .evaluation import keypoints_from_heatmaps
from .post_processing import flip_back, oks_nms, oks_nms_batched, soft_oks_nms, transform_preds
from .results import write_result_keypoints

__all__ = ['keypoints_from_heatmaps', 'flip_back', 'transform_preds', 'oks_nms', 'soft_oks_nms', 'oks_nms_batched',
           'write_result_keypoints']
