"""``TopDown`` pose detector with the reference's registry name, constructor and ``forward`` /
``forward_test`` contract (mmpose/models/detectors/top_down.py:23-218).

``forward_test`` is the fused B200 path: one launch sequence runs the crops AND their horizontal flips
through backbone + head (the flip is generated inside the im2col kernel), and one decode kernel does
flip_back + shift + average + argmax + refinement + transform_preds on the device.  The only device->host
copy is the [N,K,3] result (plus the averaged heatmap when ``return_heatmap=True``); the reference copies
every heatmap twice and decodes in Python loops (simple_head.py:219,226; top_down_eval.py:598-617).
"""
import os
import warnings

import numpy as np
import torch
import torch.nn as nn

from .. import builder
from ..builder import POSENETS
from ..core.post_processing import flip_index_from_pairs
from ..engine import decode_mode_from_cfg
from ..heads.topdown_heatmap_simple_head import pack_results


@POSENETS.register_module()
class TopDown(nn.Module):

    def __init__(self, backbone, neck=None, keypoint_head=None, train_cfg=None, test_cfg=None, pretrained=None,
                 loss_pose=None):
        super().__init__()
        self.fp16_enabled = False
        # crops in the first H2D chunk when forward_test is fed host tensors (VPB_HOST_CHUNK: A/B switch)
        self.host_chunk = int(os.environ.get('VPB_HOST_CHUNK', '0'))      # 0: chosen from the backbone width below
        # optional schedule of leading chunk sizes, the rest goes as one batch (VPB_HOST_CHUNKS=16,48: A/B switch)
        self.host_chunks = [int(v) for v in os.environ.get('VPB_HOST_CHUNKS', '').split(',') if v]
        self._copy_stream = None
        self.backbone = builder.build_backbone(backbone)
        if self.host_chunk <= 0:
            # the first chunk must compute for as long as the copy of the rest takes (~10.7 us per crop over PCIe):
            # ViTPose-S needs ~37 us per crop (with flip) -> 64 crops of 256, ViTPose-B 84 us -> 32 (measured), L / H less
            self.host_chunk = 64 if int(getattr(self.backbone, 'embed_dim', 768)) < 768 else 32
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg if test_cfg is not None else {}
        if neck is not None:
            raise NotImplementedError('necks are not used by any ViTPose config')
        if keypoint_head is not None:
            keypoint_head = dict(keypoint_head)
            keypoint_head['train_cfg'] = train_cfg
            keypoint_head['test_cfg'] = test_cfg
            if 'loss_keypoint' not in keypoint_head and loss_pose is not None:
                warnings.warn('`loss_pose` for TopDown is deprecated, use `loss_keypoint` for heads instead. See '
                              'https://github.com/open-mmlab/mmpose/pull/382 for more information.',
                              DeprecationWarning)
                keypoint_head['loss_keypoint'] = loss_pose
            self.keypoint_head = builder.build_head(keypoint_head)
        self.init_weights(pretrained=pretrained)

    @property
    def with_neck(self):
        return hasattr(self, 'neck')

    @property
    def with_keypoint(self):
        return hasattr(self, 'keypoint_head')

    def init_weights(self, pretrained=None):
        self.backbone.init_weights(pretrained)
        if self.with_keypoint:
            self.keypoint_head.init_weights()

    def forward(self, img, target=None, target_weight=None, img_metas=None, return_loss=True,
                return_heatmap=False, **kwargs):
        """Dispatch of top_down.py:138-141."""
        if return_loss:
            return self.forward_train(img, target, target_weight, img_metas, **kwargs)
        return self.forward_test(img, img_metas, return_heatmap=return_heatmap, **kwargs)

    def forward_train(self, img, target, target_weight, img_metas=None, **kwargs):
        """top_down.py:143-161: heatmaps of the batch, then the head's loss. The network is one autograd node whose
        backward is the hand-written CUDA backward pass (vitpose_b200/training.py), so ``loss.backward()`` fills
        ``.grad`` of every parameter exactly as in the reference's training loop."""
        from ..training import network_heatmaps_train
        output = network_heatmaps_train(self, img)
        losses = dict()
        if self.with_keypoint:
            losses.update(self.keypoint_head.get_loss(output, target, target_weight))
            losses.update(self.keypoint_head.get_accuracy(output, target, target_weight))
        return losses

    # mmpose/models/detectors/base.py:66-74 turns every logged value into a Python float with one ``.item()`` (a
    # device->host sync) and, when distributed, one all_reduce per value. Here the values are stacked, averaged over
    # the ranks with ONE all_reduce (NCCL) and read back with ONE copy. ``log_vars_on_device = True`` skips the
    # read-back and leaves 0-dim device tensors (no host sync in the step; not what an mmcv log buffer expects).
    log_vars_on_device = os.environ.get('VPB_LOG_ON_DEVICE', '0') == '1'

    def _parse_losses(self, losses):
        """mmpose/models/detectors/base.py:37-76: ``loss`` = sum of the entries whose name contains 'loss';
        ``log_vars`` = every entry (and the total), averaged over the ranks when torch.distributed is initialised."""
        import torch.distributed as dist
        log_vars = {}
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                log_vars[name] = value.mean()
            elif isinstance(value, float):
                log_vars[name] = value
            elif isinstance(value, list):
                log_vars[name] = sum(v.mean() for v in value)
            else:
                raise TypeError(f'{name} is not a tensor or list of tensors or float')
        loss = sum(v for k, v in log_vars.items() if 'loss' in k)
        log_vars['loss'] = loss
        names = [k for k, v in log_vars.items() if not isinstance(v, float)]
        if names:
            dev = next(v.device for v in (log_vars[k] for k in names))
            packed = torch.stack([log_vars[k].detach().to(device=dev, dtype=torch.float32) for k in names])
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                packed = packed / dist.get_world_size()
                dist.all_reduce(packed)
            if self.log_vars_on_device:
                for i, k in enumerate(names):
                    log_vars[k] = packed[i]
            else:
                for k, v in zip(names, packed.tolist()):
                    log_vars[k] = v
        return loss, log_vars

    def train_step(self, data_batch, optimizer=None, **kwargs):
        """base.py:88-119: forward + loss parsing; the caller (runner hook) does backward() and optimizer.step()."""
        losses = self.forward(**data_batch)
        loss, log_vars = self._parse_losses(losses)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(next(iter(data_batch.values()))))

    def _engine(self):
        return self.backbone.engine(self.keypoint_head if self.with_keypoint else None)

    @torch.no_grad()
    def forward_test(self, img, img_metas, return_heatmap=False, **kwargs):
        """Same checks, same result dict as top_down.py:163-200."""
        assert img.size(0) == len(img_metas)
        batch_size, _, img_height, img_width = img.shape
        if batch_size > 1:
            assert 'bbox_id' in img_metas[0]
        test_cfg = self.test_cfg
        eng = self._engine()
        dev = eng.device
        flip = bool(test_cfg.get('flip_test', True))
        result = {}
        if not self.with_keypoint:
            return result
        n = batch_size
        K = eng.desc.num_keypoints
        H4, W4 = eng.heatmap_size
        chunk = self.host_chunk
        if (not img.is_cuda) and img.dtype == torch.float32 and n > chunk:
            # Host crops: copy chunk i+1 on a side stream while chunk i runs through the network, writing every
            # chunk's maps into contiguous [N,...] main / flipped buffers so ONE decode sees the whole batch.
            hm_main = torch.empty(n, K, H4, W4, device=dev, dtype=torch.float32)
            hm_flip = torch.empty(n, K, H4, W4, device=dev, dtype=torch.float32) if flip else None
            main_stream = torch.cuda.current_stream(dev)
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(dev)
            staged = []
            # two chunks: a first one of `host_chunk` crops starts the GPU after a ~0.35 ms copy, and the copy of all
            # the remaining crops (PCIe: ~2.4 ms per 224 crops) hides behind its ~2.7 ms of compute; the rest then runs
            # as one large batch (few, full-size GEMM launches). Measured (256 crops + flips, end-to-end minus
            # device-resident ms per step, same box): first chunk 64: 0.82 / 1.14, 48: 0.71, 32: 0.47-0.63, 24: 0.71;
            # three chunks 16,48,rest: 0.75-0.92; 8,24,64,rest: 2.0 (profiles/r02_summary.md section 15).
            sizes = list(self.host_chunks) if self.host_chunks else [max(1, chunk)]
            lo = 0
            while lo < n:
                size = min(sizes.pop(0) if sizes else n - lo, n - lo)
                with torch.cuda.stream(self._copy_stream):
                    d = img[lo:lo + size].to(dev, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(self._copy_stream)
                staged.append((lo, d, ev))
                lo += size
            for lo, d, ev in staged:
                main_stream.wait_event(ev)
                d.record_stream(main_stream)
                eng.forward_into(d, flip, hm_main[lo:lo + d.shape[0]],
                                 hm_flip[lo:lo + d.shape[0]] if flip else None)
            hm = (hm_main, hm_flip)
        else:
            img = img.to(device=dev, dtype=torch.float32, non_blocking=True)
            hm_all, _ = eng.forward_heatmaps(img, flip=flip)
            hm = (hm_all[:n], hm_all[n:2 * n] if flip else None)
        c = np.zeros((n, 2), dtype=np.float32)
        s = np.zeros((n, 2), dtype=np.float32)
        score = np.ones(n)
        image_paths = []
        bbox_ids = [] if 'bbox_id' in img_metas[0] else None
        for i in range(n):
            c[i, :] = img_metas[i]['center']
            s[i, :] = img_metas[i]['scale']
            image_paths.append(img_metas[i]['image_file'])
            if 'bbox_score' in img_metas[i]:
                score[i] = np.array(img_metas[i]['bbox_score']).reshape(-1)[0]
            if bbox_ids is not None:
                bbox_ids.append(img_metas[i]['bbox_id'])
        cs = torch.from_numpy(np.concatenate([c, s], axis=1)).to(dev, non_blocking=True)
        flip_index = None
        if flip:
            flip_index = torch.from_numpy(flip_index_from_pairs(K, img_metas[0]['flip_pairs'])).to(dev)
        mode = decode_mode_from_cfg(test_cfg)
        from .. import ops
        r = ops.decode(hm[0], hm[1], flip_index, bool(test_cfg.get('shift_heatmap', False)), mode,
                       test_cfg.get('modulate_kernel', 11), bool(test_cfg.get('use_udp', False)),
                       cs[:, 0:2].contiguous(), cs[:, 2:4].contiguous(), want_merged=return_heatmap and flip)
        packed = torch.cat([r['preds'], r['maxvals']], dim=2)          # [N,K,3] -> one D2H copy
        packed = packed.cpu().numpy()
        result.update(pack_results(packed[:, :, 0:2], packed[:, :, 2:3], c, s, score, image_paths, bbox_ids))
        if return_heatmap:
            output_heatmap = r['merged'].cpu().numpy() if flip else hm[0].cpu().numpy()
        else:
            output_heatmap = None
        result['output_heatmap'] = output_heatmap
        return result

    @torch.no_grad()
    def forward_dummy(self, img):
        """Heatmaps for FLOP counting tools (top_down.py:202-218)."""
        hm, _ = self._engine().forward_heatmaps(img.float(), flip=False)
        return hm
