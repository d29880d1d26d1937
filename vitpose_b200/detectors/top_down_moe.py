"""``TopDownMoE`` — the ViTPose+ detector (mmpose/models/detectors/top_down_moe.py:15-272): a ``ViTMoE`` backbone
whose FFN experts are selected per crop by ``img_metas[i]['dataset_idx']``, the main keypoint head and the
``associate_keypoint_heads`` of the other datasets (used by the reference for training and by
tools/model_split.py to cut per-dataset models; ``forward_test`` decodes with the main head, :205-244)."""
import numpy as np
import torch
import torch.nn as nn

from .. import builder
from ..builder import POSENETS
from .top_down import TopDown


@POSENETS.register_module()
class TopDownMoE(TopDown):

    def __init__(self, backbone, neck=None, keypoint_head=None, associate_keypoint_head=None, train_cfg=None,
                 test_cfg=None, pretrained=None, loss_pose=None):
        super().__init__(backbone, neck, keypoint_head, train_cfg, test_cfg, pretrained, loss_pose)
        heads = []
        if associate_keypoint_head is not None:
            if not isinstance(associate_keypoint_head, list):
                associate_keypoint_head = [associate_keypoint_head]
            for cfg in associate_keypoint_head:
                cfg = dict(cfg)
                cfg['train_cfg'] = train_cfg
                cfg['test_cfg'] = test_cfg
                heads.append(builder.build_head(cfg))
        self.associate_keypoint_heads = nn.ModuleList(heads)
        self.keypoint_heads_cnt = len(heads) + 1
        for h in self.associate_keypoint_heads:
            h.init_weights()

    def forward_train(self, img, target, target_weight, img_metas=None, **kwargs):
        raise NotImplementedError('multi-dataset ViTPose+ training is outside the B200 path (inference only)')

    @torch.no_grad()
    def forward_test(self, img, img_metas, return_heatmap=False, **kwargs):
        """top_down_moe.py:205-244. Crops are grouped by dataset index; each group runs through the engine packed
        with that dataset's effective FFN weights, and the groups' results are put back in batch order."""
        assert img.size(0) == len(img_metas)
        if img.size(0) > 1:
            assert 'bbox_id' in img_metas[0]
        src = np.array([int(m['dataset_idx']) for m in img_metas])
        groups = sorted(set(src.tolist()))
        if len(groups) == 1:
            self._dataset_idx = groups[0]
            return TopDown.forward_test(self, img, img_metas, return_heatmap=return_heatmap, **kwargs)
        merged = None
        n = len(img_metas)
        for d in groups:
            idx = np.nonzero(src == d)[0]
            self._dataset_idx = d
            sel = torch.from_numpy(idx).to(img.device)
            # the reference flips the WHOLE batch back with img_metas[0]['flip_pairs'] (top_down_moe.py:229-232)
            metas_d = [img_metas[i] for i in idx]
            if 'flip_pairs' in img_metas[0]:
                metas_d[0] = dict(metas_d[0], flip_pairs=img_metas[0]['flip_pairs'])
            r = TopDown.forward_test(self, img.index_select(0, sel), metas_d, return_heatmap=return_heatmap, **kwargs)
            if merged is None:
                merged = dict(preds=np.zeros((n,) + r['preds'].shape[1:], r['preds'].dtype),
                              boxes=np.zeros((n,) + r['boxes'].shape[1:], r['boxes'].dtype),
                              image_paths=[None] * n, bbox_ids=[None] * n if r['bbox_ids'] is not None else None,
                              output_heatmap=None if r['output_heatmap'] is None else
                              np.zeros((n,) + r['output_heatmap'].shape[1:], r['output_heatmap'].dtype))
            merged['preds'][idx], merged['boxes'][idx] = r['preds'], r['boxes']
            for j, i in enumerate(idx):
                merged['image_paths'][i] = r['image_paths'][j]
                if merged['bbox_ids'] is not None:
                    merged['bbox_ids'][i] = r['bbox_ids'][j]
            if merged['output_heatmap'] is not None:
                merged['output_heatmap'][idx] = r['output_heatmap']
        return merged

    def _engine(self):
        return self.backbone.engine(self.keypoint_head if self.with_keypoint else None,
                                    getattr(self, '_dataset_idx', 0))
