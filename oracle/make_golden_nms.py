"""Generates tests/golden/nms.npz by running the UNMODIFIED reference functions (mmpose/core/post_processing/nms.py,
loaded by path — the module needs only NumPy) on seeded pose sets. Run in the authoring container:
    python -m oracle.make_golden_nms
"""
import importlib.util
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference/mmpose/core/post_processing/nms.py'


def load_ref():
    spec = importlib.util.spec_from_file_location('ref_nms', REF)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def pose_set(seed, P, K):
    """P poses around a few people: clusters of near-duplicates (what NMS is for) with distinct scores."""
    rng = np.random.RandomState(seed)
    people = max(1, P // 4)
    base = rng.rand(people, K, 2).astype(np.float32) * np.float32(200) + np.float32(50)
    who = rng.randint(people, size=P)
    kp = np.zeros((P, K, 3), dtype=np.float32)
    kp[:, :, :2] = base[who] + rng.randn(P, K, 2).astype(np.float32) * rng.choice([1.0, 4.0, 15.0], size=(P, 1, 1)).astype(np.float32)
    kp[:, :, 2] = rng.rand(P, K).astype(np.float32)
    areas = (rng.rand(P) * 3e4 + 5e3)
    scores = rng.permutation(P).astype(np.float64) / P + rng.rand(P) * 1e-3       # distinct
    return kp, areas, scores


def cases():
    return [(0, 12, 17, 0.9, None), (1, 40, 17, 0.9, 0.2), (2, 25, 133, 0.9, None), (3, 1, 17, 0.9, 0.2),
            (4, 64, 17, 0.5, None), (5, 30, 133, 0.8, 0.3)]


def main():
    ref = load_ref()
    out = {}
    for seed, P, K, thr, vis in cases():
        kp, areas, scores = pose_set(seed, P, K)
        sig = None if K == 17 else np.full(K, 0.05)
        db = [dict(keypoints=kp[i], score=scores[i], area=areas[i]) for i in range(P)]
        out[f'c{seed}_kpts'], out[f'c{seed}_areas'], out[f'c{seed}_scores'] = kp, areas, scores
        out[f'c{seed}_hard'] = np.asarray(ref.oks_nms(db, thr, sigmas=sig, vis_thr=vis), dtype=np.int64)
        out[f'c{seed}_soft'] = np.asarray(ref.soft_oks_nms(db, thr, max_dets=20, sigmas=sig, vis_thr=vis), dtype=np.int64)
        out[f'c{seed}_iou'] = ref.oks_iou(kp[0].flatten(), kp.reshape(P, -1), areas[0], areas, sigmas=sig, vis_thr=vis)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'nms.npz'), **out)
    print('wrote nms.npz:', {k: v.shape for k, v in out.items() if k.endswith('hard')})


if __name__ == '__main__':
    main()
