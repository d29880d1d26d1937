"""TEST INFRASTRUCTURE — CPU (NumPy) restatement of the post-decode evaluation step of the top-down COCO datasets
(SURVEY.md §8f rank 2). Only ``tests/`` may import it; never the product path.

* ``rescore``       mmpose/datasets/datasets/top_down/topdown_coco_dataset.py:476-490: pose score = mean of the joint
                    scores above ``vis_thr`` times the box score (float32 accumulation in joint order).
* ``oks_iou``       mmpose/core/post_processing/nms.py:51-86 (COCO sigmas default; squared distances in the keypoints'
                    dtype, the rest in float64; float32 result; with ``vis_thr`` only the DETECTION's visibility selects
                    joints — ``list(vg > t) and list(vd > t)`` evaluates to the second list).
* ``oks_nms``       nms.py:89-128: greedy, keeps poses whose OKS with every kept pose is <= thr.
* ``soft_oks_nms``  nms.py:131-207 with the Gaussian ``_rescore`` (:131-151): scores *= exp(-oks^2 / thr), at most
                    ``max_dets`` poses, in selection order.

Parity pin: tests/test_nms.py runs these against the reference functions themselves (nms.py needs only NumPy and is
loaded by path) when /root/reference is mounted, and against tests/golden/nms.npz everywhere.
"""
import numpy as np

COCO_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89]) / 10.0


def rescore(keypoints, box_scores, vis_thr):
    """keypoints [P,K,3] (x, y, score), box_scores [P] -> pose scores [P] (topdown_coco_dataset.py:476-490)."""
    out = []
    for kp, bs in zip(keypoints, box_scores):
        acc, cnt = 0, 0
        for t in kp[:, 2]:
            if t > vis_thr:
                acc = acc + t
                cnt += 1
        if cnt != 0:
            acc = acc / cnt
        out.append(acc * bs)
    return np.array(out)


def oks_iou(g, d, a_g, a_d, sigmas=None, vis_thr=None):
    """g [3K] flat keypoints of the reference pose, d [n,3K] the others; a_g, a_d[n] areas."""
    sigmas = COCO_SIGMAS if sigmas is None else sigmas
    var = (sigmas * 2) ** 2
    xg, yg = g[0::3], g[1::3]
    out = np.zeros(len(d), dtype=np.float32)
    for n in range(len(d)):
        dx = d[n, 0::3] - xg
        dy = d[n, 1::3] - yg
        e = (dx ** 2 + dy ** 2) / var / ((a_g + a_d[n]) / 2 + np.spacing(1)) / 2
        if vis_thr is not None:
            e = e[d[n, 2::3] > vis_thr]
        out[n] = np.sum(np.exp(-e)) / len(e) if len(e) != 0 else 0.0
    return out


def oks_nms(kpts, scores, areas, thr, sigmas=None, vis_thr=None):
    """kpts [P,3K], scores [P], areas [P] -> kept indices in selection order (nms.py:89-128)."""
    if len(kpts) == 0:
        return np.zeros(0, dtype=np.intp)
    order = scores.argsort()[::-1]
    keep = []
    while len(order) > 0:
        i = order[0]
        keep.append(i)
        ovr = oks_iou(kpts[i], kpts[order[1:]], areas[i], areas[order[1:]], sigmas, vis_thr)
        order = order[np.where(ovr <= thr)[0] + 1]
    return np.array(keep)


def soft_oks_nms(kpts, scores, areas, thr, max_dets=20, sigmas=None, vis_thr=None):
    if len(kpts) == 0:
        return np.zeros(0, dtype=np.intp)
    order = scores.argsort()[::-1]
    scores = scores[order]
    keep = []
    while len(order) > 0 and len(keep) < max_dets:
        i = order[0]
        ovr = oks_iou(kpts[i], kpts[order[1:]], areas[i], areas[order[1:]], sigmas, vis_thr)
        order = order[1:]
        scores = scores[1:] * np.exp(-ovr ** 2 / thr)
        tmp = scores.argsort()[::-1]
        order, scores = order[tmp], scores[tmp]
        keep.append(i)
    return np.array(keep, dtype=np.intp)
