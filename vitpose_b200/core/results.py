"""Post-decode evaluation step of the top-down COCO datasets, from ``forward_test`` result dicts to the
``result_keypoints.json`` wire format (reference
mmpose/datasets/datasets/top_down/topdown_coco_dataset.py:405-571).

``collect_results`` -> ``sort_and_unique_bboxes`` are host bookkeeping (dict / sort, as in the reference);
``rescore_and_nms`` runs the per-pose rescoring (:476-490) and ``oks_nms`` / ``soft_oks_nms`` (:492-497,
mmpose/core/post_processing/nms.py:89-207) of EVERY image of the evaluation in ONE launch of the CUDA kernel
(``vpb_oks_nms``: one CTA per image, NumPy's fp32 / fp64 arithmetic reproduced) instead of O(P^2) NumPy work per image;
``write_coco_keypoint_results`` writes the same JSON text as ``_write_coco_keypoint_results`` (:529-571).
COCOeval itself (xtcocotools) stays outside the path.
"""
import json
import os

import numpy as np

from .post_processing import oks_nms_batched


def collect_results(results, name2id, img_prefix):
    """[{preds [N,K,3], boxes [N,6], image_paths, bbox_ids}, ...] -> {image_id: [person dict, ...]} in the order the
    images first appear (topdown_coco_dataset.py:447-467)."""
    kpts = {}
    for result in results:
        preds, boxes, paths, bbox_ids = result['preds'], result['boxes'], result['image_paths'], result['bbox_ids']
        for i in range(len(paths)):
            image_id = name2id[paths[i][len(img_prefix):]]
            kpts.setdefault(image_id, []).append({
                'keypoints': preds[i], 'center': boxes[i][0:2], 'scale': boxes[i][2:4], 'area': boxes[i][4],
                'score': boxes[i][5], 'image_id': image_id, 'bbox_id': bbox_ids[i]})
    return kpts


def sort_and_unique_bboxes(kpts, key='bbox_id'):
    """Per image: sort by ``bbox_id`` (stable) and keep the first of every run of equal ids (:667-676) — the
    distributed sampler pads the last batch with repeated samples."""
    for img_id, persons in kpts.items():
        persons = sorted(persons, key=lambda x: x[key])
        kpts[img_id] = [p for j, p in enumerate(persons) if j == 0 or persons[j - 1][key] != p[key]]
    return kpts


def rescore_and_nms(kpts, vis_thr, oks_thr, sigmas=None, use_nms=True, soft_nms=False, rle_score=False, max_dets=20):
    """Sets ``score`` of every pose (mean of the joint scores above ``vis_thr`` times the box score; with
    ``rle_score`` box + mean + max of the joint scores) and returns, per image in dict order, the poses that survive
    OKS NMS in selection order. All images go through one ``vpb_oks_nms`` launch."""
    image_ids = list(kpts.keys())
    people = [p for i in image_ids for p in kpts[i]]
    if not people:
        return [[] for _ in image_ids]
    starts = np.zeros(len(image_ids) + 1, dtype=np.int32)
    starts[1:] = np.cumsum([len(kpts[i]) for i in image_ids])
    kp = np.stack([np.asarray(p['keypoints'], dtype=np.float32).reshape(-1, 3) for p in people])
    f32_areas = all(isinstance(p['area'], np.float32) for p in people)      # boxes[:, 4] of forward_test
    areas = np.array([p['area'] for p in people], dtype=np.float64)
    box = np.array([p['score'] for p in people])
    if rle_score:
        scores = np.array([float(b + np.mean(k[:, 2]) + np.max(k[:, 2])) for k, b in zip(kp, box)])
        keep, _ = oks_nms_batched(kp, areas, scores, starts, oks_thr, sigmas, None, soft=soft_nms, max_dets=max_dets,
                                  areas_float32=f32_areas)
        used = scores
        for p, s in zip(people, used):
            p['score'] = float(s)
    else:
        keep, used = oks_nms_batched(kp, areas, box.astype(np.float64), starts, oks_thr, sigmas, None, soft=soft_nms,
                                     max_dets=max_dets, rescore=True, rescore_vis_thr=vis_thr,
                                     areas_float32=f32_areas)
        for p, s in zip(people, used):
            p['score'] = np.float32(s)              # the reference's float32 product kpt_score * box_score
    if not use_nms:
        return [kpts[i] for i in image_ids]
    return [[people[int(j)] for j in keep[g]] for g in range(len(image_ids))]


def coco_keypoint_results(valid_kpts, num_joints, cat_id=1):
    """``_coco_keypoint_results_one_category_kernel`` (:547-571): one entry per kept pose."""
    out = []
    for img_kpts in valid_kpts:
        if len(img_kpts) == 0:
            continue
        key_points = np.array([p['keypoints'] for p in img_kpts]).reshape(-1, num_joints * 3)
        out.extend({'image_id': p['image_id'], 'category_id': cat_id, 'keypoints': kp.tolist(),
                    'score': float(p['score']), 'center': np.asarray(p['center']).tolist(),
                    'scale': np.asarray(p['scale']).tolist()} for p, kp in zip(img_kpts, key_points))
    return out


def write_coco_keypoint_results(valid_kpts, res_file, num_joints, cat_id=1):
    with open(res_file, 'w') as f:
        json.dump(coco_keypoint_results(valid_kpts, num_joints, cat_id), f, sort_keys=True, indent=4)


def write_result_keypoints(results, res_folder, name2id, img_prefix, num_joints, vis_thr=0.2, oks_thr=0.9, sigmas=None,
                           use_nms=True, soft_nms=False, rle_score=False):
    """The part of ``TopDownCocoDataset.evaluate`` (:405-505) before COCOeval: writes
    ``<res_folder>/result_keypoints.json`` and returns its path."""
    kpts = sort_and_unique_bboxes(collect_results(results, name2id, img_prefix))
    valid = rescore_and_nms(kpts, vis_thr, oks_thr, sigmas, use_nms, soft_nms, rle_score)
    res_file = os.path.join(res_folder, 'result_keypoints.json')
    write_coco_keypoint_results(valid, res_file, num_joints)
    return res_file
