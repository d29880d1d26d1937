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
        """top_down_moe.py:166-203: the backbone with each crop's expert, then EVERY head on the whole batch; head i's
        loss / accuracy see only the crops of dataset i (targets and target weights of the others multiplied by 0).
        Here the crops are first sorted by dataset (losses and accuracies are means over the batch: the order does
        not matter), so that mlp.fc2 and its gradients run once per run of crops; backbone + all heads are one autograd
        node (vitpose_b200/training.py) and ``loss.backward()`` fills every parameter's ``.grad`` — every expert
        gets one (zero when its dataset is absent from the batch), as the reference's dense masked form does."""
        from ..training import network_heatmaps_train
        src = [int(m['dataset_idx']) for m in img_metas]
        assert len(src) == img.size(0)
        order, runs = self.backbone.dataset_runs(src)
        if order != list(range(len(src))):
            perm = torch.tensor(order, device=img.device)
            img = img.index_select(0, perm)
            target = target.index_select(0, perm.to(target.device))
            target_weight = target_weight.index_select(0, perm.to(target_weight.device))
        img_sources = torch.tensor([src[i] for i in order], device=target.device)
        self._vpb_dataset_runs = runs
        try:
            outputs = network_heatmaps_train(self, img)
        finally:
            self._vpb_dataset_runs = None
        if not isinstance(outputs, tuple):
            outputs = (outputs,)
        losses = dict()
        heads = [self.keypoint_head, *self.associate_keypoint_heads]
        for idx, (head, output) in enumerate(zip(heads, outputs)):
            select = (img_sources == idx)
            target_select = target * select.view(-1, 1, 1, 1)
            target_weight_select = target_weight * select.view(-1, 1, 1)
            loss = head.get_loss(output, target_select, target_weight_select)['heatmap_loss']
            acc = head.get_accuracy(output, target_select, target_weight_select)['acc_pose']
            losses['main_stream_loss' if idx == 0 else f'{idx}_loss'] = loss
            losses['main_stream_acc' if idx == 0 else f'{idx}_acc'] = acc
        return losses

    @torch.no_grad()
    def forward_test(self, img, img_metas, return_heatmap=False, **kwargs):
        """top_down_moe.py:205-244. A batch of one dataset runs through the engine packed with that dataset's effective
        FFN weights; a mixed batch runs in one pass with the crops sorted by dataset (see below)."""
        assert img.size(0) == len(img_metas)
        if img.size(0) > 1:
            assert 'bbox_id' in img_metas[0]
        src = np.array([int(m['dataset_idx']) for m in img_metas])
        groups = sorted(set(src.tolist()))
        if len(groups) == 1:
            self._dataset_idx = groups[0]
            return TopDown.forward_test(self, img, img_metas, return_heatmap=return_heatmap, **kwargs)
        # Mixed batch: ONE forward pass. Crops are sorted by dataset; every kernel runs on the whole batch except
        # mlp.fc2 (+ residual + LayerNorm), launched once per run of crops with that dataset's expert columns
        # (vpb_moe_runs); the results go back in batch order.
        order, runs = self.backbone.dataset_runs(src)
        eng = self.backbone.moe_engine(self.keypoint_head if self.with_keypoint else None)
        dev = eng.device
        perm = torch.tensor(order)
        metas_p = [img_metas[i] for i in order]
        # the reference flips the WHOLE batch back with img_metas[0]['flip_pairs'] (top_down_moe.py:229-232)
        if 'flip_pairs' in img_metas[0]:
            metas_p[0] = dict(metas_p[0], flip_pairs=img_metas[0]['flip_pairs'])
        img_p = img.index_select(0, perm.to(img.device)).to(dev, non_blocking=True)     # (one call: no host chunking)
        eng.set_moe_runs(runs)
        self._use_moe_engine = True
        try:
            r = TopDown.forward_test(self, img_p, metas_p, return_heatmap=return_heatmap, **kwargs)
        finally:
            self._use_moe_engine = False
            eng.set_moe_runs(None)
        n = len(img_metas)
        inv = np.empty(n, dtype=np.int64)
        inv[np.asarray(order)] = np.arange(n)
        out = dict(r)
        out['preds'], out['boxes'] = r['preds'][inv], r['boxes'][inv]
        out['image_paths'] = [r['image_paths'][j] for j in inv]
        out['bbox_ids'] = None if r['bbox_ids'] is None else [r['bbox_ids'][j] for j in inv]
        if r.get('output_heatmap') is not None:
            out['output_heatmap'] = r['output_heatmap'][inv]
        return out

    def _engine(self):
        head = self.keypoint_head if self.with_keypoint else None
        if getattr(self, '_use_moe_engine', False):
            return self.backbone.moe_engine(head)
        return self.backbone.engine(head, getattr(self, '_dataset_idx', 0))
