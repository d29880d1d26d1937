"""TEST INFRASTRUCTURE — CPU restatement of the post-decode evaluation step of the top-down COCO datasets up to the
``result_keypoints.json`` wire format. Only ``tests/`` may import it; never the product path.

* ``collect``                 topdown_coco_dataset.py:447-467: ``forward_test`` result dicts -> {image_id: [person dicts]}
                              (image id from ``name2id[path[len(img_prefix):]]``, insertion order of the images kept)
* ``sort_and_unique_bboxes``  :667-676: per image, stable sort by ``bbox_id``, later duplicates dropped
* ``rescore_and_nms``         :471-503: pose score = mean joint score above ``vis_thr`` x box score (or the RLE score),
                              then ``oks_nms`` / ``soft_oks_nms`` per image (oracle/nms_np.py)
* ``result_entries`` / ``dump`` :529-571: one dict per kept pose — image_id, category_id, keypoints (K*3 floats),
                              score, center, scale — written with ``json.dump(..., sort_keys=True, indent=4)``
                              (``json_tricks.dump`` in the reference: the same stdlib encoder for these plain types).

Parity pin: tests/test_results.py compares ``dump`` with tests/golden/result_keypoints.json, produced by the unmodified
reference ``TopDownCocoDataset.evaluate`` (oracle/make_golden_results.py), byte for byte.
"""
import json
from collections import defaultdict

import numpy as np

from . import nms_np


def collect(results, name2id, img_prefix):
    kpts = defaultdict(list)
    for result in results:
        preds, boxes = result['preds'], result['boxes']
        for i, path in enumerate(result['image_paths']):
            image_id = name2id[path[len(img_prefix):]]
            kpts[image_id].append(dict(keypoints=preds[i], center=boxes[i][0:2], scale=boxes[i][2:4],
                                       area=boxes[i][4], score=boxes[i][5], image_id=image_id,
                                       bbox_id=result['bbox_ids'][i]))
    return kpts


def sort_and_unique_bboxes(kpts, key='bbox_id'):
    for img_id in kpts:
        people = sorted(kpts[img_id], key=lambda x: x[key])
        out = []
        for p in people:
            if not out or out[-1][key] != p[key]:
                out.append(p)
        kpts[img_id] = out
    return kpts


def rescore_and_nms(kpts, vis_thr, oks_thr, sigmas=None, use_nms=True, soft_nms=False, rle_score=False):
    valid = []
    for image_id in kpts.keys():
        people = kpts[image_id]
        kp = np.stack([p['keypoints'] for p in people])
        box = np.array([p['score'] for p in people])
        if rle_score:
            scores = [float(b + np.mean(k[:, 2]) + np.max(k[:, 2])) for k, b in zip(kp, box)]
        else:
            scores = list(nms_np.rescore(kp, box, vis_thr))
        for p, s in zip(people, scores):
            p['score'] = s
        if use_nms:
            flat = kp.reshape(len(people), -1)
            sc = np.array([p['score'] for p in people])
            areas = np.array([p['area'] for p in people])
            fn = nms_np.soft_oks_nms if soft_nms else nms_np.oks_nms
            keep = fn(flat, sc, areas, oks_thr, sigmas=sigmas) if not soft_nms else fn(flat, sc, areas, oks_thr, 20, sigmas)
            valid.append([people[int(k)] for k in keep])
        else:
            valid.append(people)
    return valid


def result_entries(valid_kpts, num_joints, cat_id=1):
    out = []
    for people in valid_kpts:
        if len(people) == 0:
            continue
        kp = np.array([p['keypoints'] for p in people]).reshape(-1, num_joints * 3)
        for p, row in zip(people, kp):
            out.append(dict(image_id=p['image_id'], category_id=cat_id, keypoints=row.tolist(), score=float(p['score']),
                            center=p['center'].tolist(), scale=p['scale'].tolist()))
    return out


def dump(entries):
    return json.dumps(entries, sort_keys=True, indent=4)
