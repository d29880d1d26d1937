"""rescoring + OKS NMS. Oracle (oracle/nms_np.py) pinned against golden outputs of the
reference functions and, when mounted, the reference itself; the CUDA kernel (-m gpu) against the oracle."""
import os

import numpy as np
import pytest

from oracle import nms_np as O
from oracle.make_golden_nms import REF, cases, pose_set


def _sig(K):
    return None if K == 17 else np.full(K, 0.05)


def test_oracle_matches_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'nms.npz'))
    for seed, P, K, thr, vis in cases():
        kp, areas, scores = g[f'c{seed}_kpts'], g[f'c{seed}_areas'], g[f'c{seed}_scores']
        flat = kp.reshape(P, -1)
        np.testing.assert_array_equal(O.oks_iou(flat[0], flat, areas[0], areas, _sig(K), vis), g[f'c{seed}_iou'])
        np.testing.assert_array_equal(O.oks_nms(flat, scores, areas, thr, _sig(K), vis), g[f'c{seed}_hard'])
        np.testing.assert_array_equal(O.soft_oks_nms(flat, scores, areas, thr, 20, _sig(K), vis), g[f'c{seed}_soft'])


@pytest.mark.reference
def test_oracle_matches_live_reference():
    if not os.path.exists(REF):
        pytest.skip('reference tree not mounted')
    from oracle.make_golden_nms import load_ref
    ref = load_ref()
    for seed in range(20, 26):
        P, K = 30 + seed, 17
        kp, areas, scores = pose_set(seed, P, K)
        db = [dict(keypoints=kp[i], score=scores[i], area=areas[i]) for i in range(P)]
        flat = kp.reshape(P, -1)
        for vis in (None, 0.2):
            np.testing.assert_array_equal(ref.oks_nms(db, 0.9, vis_thr=vis), O.oks_nms(flat, scores, areas, 0.9, None, vis))
            np.testing.assert_array_equal(ref.soft_oks_nms(db, 0.9, vis_thr=vis),
                                          O.soft_oks_nms(flat, scores, areas, 0.9, 20, None, vis))


def test_rescore_oracle():
    kp = np.zeros((2, 3, 3), dtype=np.float32)
    kp[0, :, 2] = [0.9, 0.1, 0.5]
    kp[1, :, 2] = [0.1, 0.1, 0.1]
    s = O.rescore(kp, np.array([0.8, 0.5], dtype=np.float32), 0.2)
    assert abs(s[0] - np.float32(0.7) * np.float32(0.8)) < 1e-7 and s[1] == 0


@pytest.mark.gpu
@pytest.mark.parametrize('soft', [False, True])
def test_gpu_oks_nms_matches_oracle(soft, golden_dir):
    from vitpose_b200.core.post_processing import oks_nms, oks_nms_batched, soft_oks_nms
    g = np.load(os.path.join(golden_dir, 'nms.npz'))
    # the reference's own outputs, through the reference-signature functions (one image per call)
    for seed, P, K, thr, vis in cases():
        kp, areas, scores = g[f'c{seed}_kpts'], g[f'c{seed}_areas'], g[f'c{seed}_scores']
        db = [dict(keypoints=kp[i], score=scores[i], area=areas[i]) for i in range(P)]
        got = (soft_oks_nms(db, thr, 20, _sig(K), vis) if soft else oks_nms(db, thr, _sig(K), vis))
        np.testing.assert_array_equal(got, g[f'c{seed}_soft' if soft else f'c{seed}_hard'])
    # a whole evaluation at once: 37 images with 0..60 poses each, with rescoring, vs the oracle image by image
    rng = np.random.RandomState(9)
    sizes = [0, 1] + list(rng.randint(0, 61, size=35))
    K, vis, thr = 17, 0.2, 0.9
    kps, areas, box = [], [], []
    for i, P in enumerate(sizes):
        kp, a, _ = pose_set(100 + i, max(P, 1), K)
        kps.append(kp[:P]); areas.append(a[:P]); box.append(rng.rand(P).astype(np.float32))
    starts = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    kp_all, a_all, b_all = np.concatenate(kps), np.concatenate(areas), np.concatenate(box)
    keep, used = oks_nms_batched(kp_all, a_all, b_all, starts, thr, None, vis, soft=soft, max_dets=20, rescore=True,
                                 rescore_vis_thr=vis)
    for i, P in enumerate(sizes):
        lo = starts[i]
        if P == 0:
            assert len(keep[i]) == 0
            continue
        sc = O.rescore(kps[i], box[i], vis).astype(np.float64)
        np.testing.assert_allclose(used[lo:lo + P], sc, rtol=0, atol=1e-7)
        flat = kps[i].reshape(P, -1)
        exp = (O.soft_oks_nms(flat, sc, areas[i], thr, 20, None, vis) if soft
               else O.oks_nms(flat, sc, areas[i], thr, None, vis))
        np.testing.assert_array_equal(keep[i] - lo, exp)
