"""``flip_back`` / ``transform_preds`` with the reference's signatures
(mmpose/core/post_processing/post_transforms.py:110-147, :150-194), executed by the CUDA kernels behind
vpb_flip_back / vpb_transform_preds.  On the fused forward_test path neither is called: both are folded
into the decode kernel."""
import numpy as np
import torch

from .. import _lib, ops


def flip_index_from_pairs(num_keypoints, flip_pairs):
    """Channel permutation equivalent to the pairwise swaps of flip_back (post_transforms.py:138-141)."""
    perm = np.arange(num_keypoints, dtype=np.int32)
    for left, right in flip_pairs:
        perm[left], perm[right] = right, left
    return perm


def flip_back(output_flipped, flip_pairs, target_type='GaussianHeatmap'):
    """Flip the flipped heatmaps back to the original form.  np.ndarray[N,K,H,W] -> np.ndarray[N,K,H,W]."""
    assert output_flipped.ndim == 4, \
        'output_flipped should be [batch_size, num_keypoints, height, width]'
    if target_type.lower() != 'GaussianHeatmap'.lower():
        raise NotImplementedError("target_type='CombinedTarget' is outside the ViTPose hot path")
    _lib.require_cuda()
    dtype = output_flipped.dtype
    dev = torch.device('cuda')
    x = torch.from_numpy(np.ascontiguousarray(output_flipped, dtype=np.float32)).to(dev)
    perm = torch.from_numpy(flip_index_from_pairs(output_flipped.shape[1], flip_pairs)).to(dev)
    out = ops.flip_back(x, perm, False).cpu().numpy()
    return out if dtype == np.float32 else out.astype(dtype)


def transform_preds(coords, center, scale, output_size, use_udp=False):
    """coords np.ndarray[K, 2|4|5]; center, scale (2,); output_size (W, H) -> coordinates in the image."""
    assert coords.shape[1] in (2, 4, 5)
    assert len(center) == 2
    assert len(scale) == 2
    assert len(output_size) == 2
    _lib.require_cuda()
    dev = torch.device('cuda')
    xy = torch.from_numpy(np.ascontiguousarray(coords[None, :, :2], dtype=np.float32)).to(dev)
    c = torch.tensor(np.asarray(center, dtype=np.float32).reshape(1, 2), device=dev)
    s = torch.tensor(np.asarray(scale, dtype=np.float32).reshape(1, 2), device=dev)
    den = (float(output_size[0]), float(output_size[1]))
    if float(den[0]).is_integer() and float(den[1]).is_integer():
        out = ops.transform_preds(xy, c, s, (int(den[0]), int(den[1])), use_udp)[0].cpu().numpy()
    else:
        raise ValueError('output_size must be integral')
    target = np.ones_like(coords)
    target[:, :2] = out
    return target


# ---- post-decode evaluation step: OKS NMS (mmpose/core/post_processing/nms.py:51-207) ------------------------------
COCO_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89]) / 10.0


def oks_nms_batched(kpts, areas, scores, group_start, thr, sigmas=None, vis_thr=None, soft=False, max_dets=20,
                    rescore=False, rescore_vis_thr=None, areas_float32=False):
    """All images of an evaluation at once: ``kpts`` [P,K,3], ``areas`` [P], ``scores`` [P] (box scores when
    ``rescore``), ``group_start`` [G+1] (poses of image g are rows group_start[g]:group_start[g+1]).
    Returns (list of G index arrays in selection order — global row indices —, the [P] scores used).
    One CTA per image runs rescoring (topdown_coco_dataset.py:476-490) and oks_nms / soft_oks_nms on the GPU.
    ``areas_float32``: the areas are float32 values (the dataset path), whose pair sums NumPy rounds to float32."""
    _lib.require_cuda()
    dev = torch.device('cuda')
    kpts = np.ascontiguousarray(kpts, dtype=np.float32)
    P, K, _ = kpts.shape
    gs = np.ascontiguousarray(group_start, dtype=np.int32)
    G = len(gs) - 1
    if P == 0 or G <= 0:
        return [np.zeros(0, dtype=np.intp) for _ in range(max(G, 0))], np.zeros(0)
    sig = COCO_SIGMAS if sigmas is None else np.asarray(sigmas, dtype=np.float64)
    assert len(sig) == K, 'one sigma per keypoint'
    var = torch.from_numpy((sig * 2) ** 2).to(dev)
    d_k = torch.from_numpy(kpts).to(dev)
    d_a = torch.from_numpy(np.ascontiguousarray(areas, dtype=np.float64)).to(dev)
    d_s = torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float64)).to(dev)
    d_g = torch.from_numpy(gs).to(dev)
    out_s = torch.empty(P, device=dev, dtype=torch.float64)
    keep = torch.empty(P, device=dev, dtype=torch.int32)
    cnt = torch.empty(G, device=dev, dtype=torch.int32)
    use_vis = vis_thr is not None
    vt = float(vis_thr) if use_vis else (float(rescore_vis_thr) if rescore_vis_thr is not None else 0.0)
    if rescore and rescore_vis_thr is not None and use_vis:
        assert float(rescore_vis_thr) == float(vis_thr), 'one visibility threshold per call'
    _lib.check(_lib.lib().vpb_oks_nms(_lib.ptr(d_k), _lib.ptr(d_a), _lib.ptr(d_s), _lib.ptr(d_g), G, K,
                                      int(np.diff(gs).max()), _lib.ptr(var), float(thr), int(use_vis), vt,
                                      int(bool(rescore)) | (2 if areas_float32 else 0), int(bool(soft)),
                                      int(max_dets), _lib.ptr(out_s),
                                      _lib.ptr(keep), _lib.ptr(cnt), _lib.stream_ptr()), 'vpb_oks_nms')
    keep, cnt = keep.cpu().numpy(), cnt.cpu().numpy()
    return [keep[gs[g]:gs[g] + cnt[g]].astype(np.intp) for g in range(G)], out_s.cpu().numpy()


def _db_arrays(kpts_db, score_per_joint):
    if score_per_joint:
        scores = np.array([k['score'].mean() for k in kpts_db])
    else:
        scores = np.array([k['score'] for k in kpts_db])
    kpts = np.array([np.asarray(k['keypoints']).reshape(-1, 3) for k in kpts_db])
    areas = np.array([k['area'] for k in kpts_db])
    return kpts, areas, scores


def oks_nms(kpts_db, thr, sigmas=None, vis_thr=None, score_per_joint=False):
    """Reference signature (nms.py:89-128): list of dicts with 'keypoints', 'score', 'area' -> indices to keep."""
    if len(kpts_db) == 0:
        return []
    kpts, areas, scores = _db_arrays(kpts_db, score_per_joint)
    keep, _ = oks_nms_batched(kpts, areas, scores, [0, len(kpts_db)], thr, sigmas, vis_thr)
    return keep[0]


def soft_oks_nms(kpts_db, thr, max_dets=20, sigmas=None, vis_thr=None, score_per_joint=False):
    """Reference signature (nms.py:154-207)."""
    if len(kpts_db) == 0:
        return []
    kpts, areas, scores = _db_arrays(kpts_db, score_per_joint)
    keep, _ = oks_nms_batched(kpts, areas, scores, [0, len(kpts_db)], thr, sigmas, vis_thr, soft=True,
                              max_dets=max_dets)
    return keep[0]
