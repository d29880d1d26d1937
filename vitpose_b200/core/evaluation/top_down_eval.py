"""``keypoints_from_heatmaps`` with the reference's signature, argument meaning, assertions and
deprecation warnings (mmpose/core/evaluation/top_down_eval.py:474-622) — computed by the fused CUDA decode
kernel (vpb_decode_heatmaps).  NumPy in, NumPy out; there is no CPU implementation here."""
import warnings

import numpy as np
import torch

from ... import _lib, ops
from ...engine import resolve_decode_mode


def _get_max_preds(heatmaps):
    """Reference: top_down_eval.py:63-95.  Argmax + max on the GPU, first index wins on ties."""
    assert isinstance(heatmaps, np.ndarray), ('heatmaps should be numpy.ndarray')
    assert heatmaps.ndim == 4, 'batch_images should be 4-ndim'
    _lib.require_cuda()
    hm = torch.from_numpy(np.ascontiguousarray(heatmaps, dtype=np.float32)).cuda()
    r = ops.decode(hm, mode=_lib.DECODE_NONE)
    return r['preds'].cpu().numpy(), r['maxvals'].cpu().numpy()


def keypoints_from_heatmaps(heatmaps,
                            center,
                            scale,
                            unbiased=False,
                            post_process='default',
                            kernel=11,
                            valid_radius_factor=0.0546875,
                            use_udp=False,
                            target_type='GaussianHeatmap'):
    """Get final keypoint predictions from heatmaps and transform them back to the image.

    Args and returns as the reference: heatmaps np.ndarray[N,K,H,W]; center, scale np.ndarray[N,2];
    returns (preds np.ndarray[N,K,2], maxvals np.ndarray[N,K,1]).
    Only the GaussianHeatmap branches exist ('megvii' and CombinedTarget are not used by any ViTPose config)."""
    # detect conflicts (same assertions as the reference, :529-534)
    if unbiased:
        assert post_process not in [False, None, 'megvii']
    if post_process in ['megvii', 'unbiased']:
        assert kernel > 0
    if use_udp:
        assert not post_process == 'megvii'

    # normalize configs (:537-560)
    if post_process is False:
        warnings.warn('post_process=False is deprecated, please use post_process=None instead',
                      DeprecationWarning)
        post_process = None
    elif post_process is True:
        if unbiased is True:
            warnings.warn("post_process=True, unbiased=True is deprecated, please use "
                          "post_process='unbiased' instead", DeprecationWarning)
            post_process = 'unbiased'
        else:
            warnings.warn("post_process=True, unbiased=False is deprecated, please use "
                          "post_process='default' instead", DeprecationWarning)
            post_process = 'default'
    elif post_process == 'default':
        if unbiased is True:
            warnings.warn("unbiased=True is deprecated, please use post_process='unbiased' instead",
                          DeprecationWarning)
            post_process = 'unbiased'

    if post_process == 'megvii':
        raise NotImplementedError("post_process='megvii' is outside the ViTPose hot path")
    if use_udp:
        tt = target_type.lower()
        if tt == 'CombinedTarget'.lower():
            raise NotImplementedError("target_type='CombinedTarget' is outside the ViTPose hot path")
        if tt != 'GaussianHeatMap'.lower():
            raise ValueError("target_type should be either 'GaussianHeatmap' or 'CombinedTarget'")

    assert isinstance(heatmaps, np.ndarray), ('heatmaps should be numpy.ndarray')
    assert heatmaps.ndim == 4, 'batch_images should be 4-ndim'
    _lib.require_cuda()
    N, K, H, W = heatmaps.shape
    mode = resolve_decode_mode(post_process, False, use_udp)
    center = np.asarray(center)
    scale = np.asarray(scale)
    dev = torch.device('cuda')
    hm = torch.from_numpy(np.ascontiguousarray(heatmaps, dtype=np.float32)).to(dev)
    # transform_preds is fused into the kernel in float32 — the dtype TopdownHeatmapBaseHead.decode passes
    # (topdown_heatmap_base_head.py:61-62). float64/int boxes are rounded to float32 first (<= 1 ulp of the
    # float64-promoted reference expression after its own float32 store).
    c = torch.from_numpy(np.ascontiguousarray(center.reshape(N, 2), dtype=np.float32)).to(dev)
    s = torch.from_numpy(np.ascontiguousarray(scale.reshape(N, 2), dtype=np.float32)).to(dev)
    r = ops.decode(hm, None, None, False, mode, kernel, use_udp, c, s)
    preds = r['preds'].cpu().numpy()
    maxvals = r['maxvals'].cpu().numpy()
    return preds, maxvals


def pose_pck_accuracy(output, target, mask, thr=0.05, normalize=None):
    """PCK accuracy from heatmaps with the reference's signature (top_down_eval.py:133-178): NumPy in, NumPy out;
    arg-max and the distance / threshold arithmetic run on the GPU (vpb_decode_heatmaps + vpb_pose_pck_accuracy).
    ``normalize`` (np.ndarray[N,2]) must be the same for every sample (the reference's default is [[H, W]] * N)."""
    N, K, H, W = output.shape
    if K == 0:
        return None, 0, 0
    _lib.require_cuda()
    norm = None
    if normalize is not None:
        normalize = np.asarray(normalize)
        assert (normalize == normalize[:1]).all(), 'per-sample normalisation factors are not implemented'
        norm = (normalize[0, 0], normalize[0, 1])
    o = torch.from_numpy(np.ascontiguousarray(output, dtype=np.float32)).cuda()
    t = torch.from_numpy(np.ascontiguousarray(target, dtype=np.float32)).cuda()
    w = torch.from_numpy(np.ascontiguousarray(mask).astype(np.float32)).cuda()
    acc, avg, cnt = ops.pose_pck_accuracy(o, t, w, thr, norm)
    return acc.cpu().numpy(), float(avg.item()), int(cnt.item())
