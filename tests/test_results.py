"""SURVEY.md §8f rank 2, second half: ``forward_test`` results -> ``result_keypoints.json``.
The oracle (oracle/results_np.py) is pinned byte for byte against tests/golden/result_keypoints.json, the text the
unmodified reference ``TopDownCocoDataset.evaluate`` wrote for the seeded cases of oracle/make_golden_results.py (and
against the live reference when mounted); the product path (-m gpu: rescoring + OKS NMS of all images in one CUDA
launch) must write the same bytes."""
import json
import os

import numpy as np
import pytest

from oracle import make_golden_results as G
from oracle import results_np as R

SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89]) / 10.0


def _golden(golden_dir):
    return (np.load(os.path.join(golden_dir, 'results_case.npz')),
            json.load(open(os.path.join(golden_dir, 'result_keypoints.json'))))


def _oracle_text(results, name2id, case):
    k = R.sort_and_unique_bboxes(R.collect(results, name2id, G.IMG_PREFIX))
    v = R.rescore_and_nms(k, 0.2, 0.9, SIGMAS, case['use_nms'], case['soft_nms'], case.get('rle_score', False))
    return R.dump(R.result_entries(v, 17))


def test_oracle_matches_golden_json(golden_dir):
    g, texts = _golden(golden_dir)
    for case in G.CASES:
        results, name2id = G.unpack_results(g, case['tag'])
        assert _oracle_text(results, name2id, case) == texts[case['tag']], case['tag']
    # wire format: a flat list, one entry per kept pose, K * 3 floats
    entries = json.loads(texts['hard'])
    assert isinstance(entries, list) and set(entries[0]) == {'image_id', 'category_id', 'keypoints', 'score', 'center',
                                                             'scale'}
    assert len(entries[0]['keypoints']) == 17 * 3 and entries[0]['category_id'] == 1


def test_sort_and_unique_edge_cases():
    mk = lambda b: dict(bbox_id=b, keypoints=np.zeros((17, 3), np.float32), score=np.float32(1), area=np.float32(1))
    k = R.sort_and_unique_bboxes({7: [mk(3), mk(1), mk(3), mk(2), mk(1)], 9: [mk(5)], 11: []})
    assert [p['bbox_id'] for p in k[7]] == [1, 2, 3] and len(k[9]) == 1 and k[11] == []
    from vitpose_b200.core import results as P
    k2 = P.sort_and_unique_bboxes({7: [mk(3), mk(1), mk(3), mk(2), mk(1)], 9: [mk(5)], 11: []})
    assert [p['bbox_id'] for p in k2[7]] == [1, 2, 3] and len(k2[9]) == 1 and k2[11] == []
    assert P.coco_keypoint_results([[], []], 17) == []


@pytest.mark.reference
def test_oracle_matches_live_reference(tmp_path):
    for seed, case in ((11, G.CASES[0]), (12, G.CASES[1]), (13, G.CASES[2])):
        results, name2id = G.synthetic_results(seed, images=5)
        path = G.run_reference(results, name2id, str(tmp_path), use_nms=case['use_nms'], soft_nms=case['soft_nms'],
                               sigmas=SIGMAS)
        assert open(path).read() == _oracle_text(results, name2id, case)


@pytest.mark.gpu
def test_gpu_path_writes_the_reference_bytes(golden_dir, tmp_path):
    from vitpose_b200.core import results as P
    g, texts = _golden(golden_dir)
    for case in G.CASES:
        results, name2id = G.unpack_results(g, case['tag'])
        path = P.write_result_keypoints(results, str(tmp_path), name2id, G.IMG_PREFIX, 17, vis_thr=0.2, oks_thr=0.9,
                                        sigmas=SIGMAS, use_nms=case['use_nms'], soft_nms=case['soft_nms'],
                                        rle_score=case.get('rle_score', False))
        assert os.path.basename(path) == 'result_keypoints.json'
        assert open(path).read() == texts[case['tag']], case['tag']


@pytest.mark.gpu
def test_gpu_path_larger_evaluation_vs_oracle(tmp_path):
    from vitpose_b200.core import results as P
    results, name2id = G.synthetic_results(21, images=200)
    path = P.write_result_keypoints(results, str(tmp_path), name2id, G.IMG_PREFIX, 17, sigmas=SIGMAS)
    assert open(path).read() == _oracle_text(results, name2id, G.CASES[0])
