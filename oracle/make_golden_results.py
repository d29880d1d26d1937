"""Generates tests/golden/results_case.npz + tests/golden/result_keypoints.json by running the UNMODIFIED reference
``TopDownCocoDataset.evaluate`` (mmpose/datasets/datasets/top_down/topdown_coco_dataset.py:405-571: collect ->
``_sort_and_unique_bboxes`` -> rescoring -> ``oks_nms`` -> ``_write_coco_keypoint_results``) on a seeded set of
``forward_test`` results. The module is loaded by path under stub parent packages (mmcv / json_tricks / COCOeval are not
installed: ``json_tricks.dump`` is replaced by the stdlib ``json.dump`` it wraps, ``deprecated_api_warning`` by the
identity), the dataset object is created without ``__init__`` and given only the attributes ``evaluate`` reads. Run in
the authoring container:
    python -m oracle.make_golden_results
"""
import importlib.util
import json
import os
import sys
import types
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = '/root/reference'
IMG_PREFIX = 'data/coco/val2017/'


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def load_reference_dataset_class():
    def ident(*a, **k):
        def deco(f):
            return f
        return deco

    class _Reg:
        def register_module(self, *a, **k):
            return lambda cls: cls

    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.split('.')[0] in ('mmcv', 'mmpose', 'json_tricks')}
    for k in saved:
        sys.modules.pop(k, None)
    try:
        _stub('mmcv', Config=object, deprecated_api_warning=ident)
        _stub('json_tricks', dump=json.dump, dumps=json.dumps, load=json.load)
        spec = importlib.util.spec_from_file_location('ref_nms_for_results',
                                                      os.path.join(REF_ROOT, 'mmpose/core/post_processing/nms.py'))
        nms = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(nms)
        _stub('mmpose')
        _stub('mmpose.core')
        _stub('mmpose.core.post_processing', oks_nms=nms.oks_nms, soft_oks_nms=nms.soft_oks_nms)
        _stub('mmpose.datasets')
        _stub('mmpose.datasets.builder', DATASETS=_Reg())
        _stub('mmpose.datasets.datasets')
        _stub('mmpose.datasets.datasets.base', Kpt2dSviewRgbImgTopDownDataset=object)
        _stub('mmpose.datasets.datasets.top_down')
        _stub('mmpose.datasets.datasets.top_down._cocoeval', COCOeval=object)
        name = 'mmpose.datasets.datasets.top_down.topdown_coco_dataset'
        spec = importlib.util.spec_from_file_location(
            name, os.path.join(REF_ROOT, 'mmpose/datasets/datasets/top_down/topdown_coco_dataset.py'))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
        return mod.TopDownCocoDataset
    finally:
        for k in [k for k in sys.modules if k.split('.')[0] in ('mmcv', 'mmpose', 'json_tricks')]:
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v


def synthetic_results(seed=0, images=6, K=17, batch=8):
    """``forward_test`` result dicts of an evaluation: several detections per image (clusters of near-duplicate poses,
    what NMS is for), a few bbox_ids repeated across batches (the sampler pads the last batch, top-down datasets dedupe
    by bbox_id), boxes [cx, cy, sx, sy, area, score] and image paths under IMG_PREFIX."""
    rng = np.random.RandomState(seed)
    per_img = rng.randint(1, 9, size=images)
    rows = []
    bbox_id = 0
    for im, n in enumerate(per_img):
        people = max(1, n // 3)
        base = rng.rand(people, K, 2).astype(np.float32) * np.float32(180) + np.float32(40)
        for j in range(n):
            kp = np.zeros((K, 3), dtype=np.float32)
            kp[:, :2] = base[rng.randint(people)] + rng.randn(K, 2).astype(np.float32) * np.float32(rng.choice([1.0, 3.0, 12.0]))
            kp[:, 2] = rng.rand(K).astype(np.float32)
            scale = (np.array([0.9, 1.2], dtype=np.float32) * np.float32(rng.uniform(0.8, 1.5)))
            box = np.array([120 + 5 * j, 140 + 3 * j, scale[0], scale[1], np.prod(scale * 200.0), rng.uniform(0.3, 1.0)],
                           dtype=np.float32)
            rows.append((im, bbox_id, kp, box))
            bbox_id += 1
    rows += [rows[i] for i in (1, len(rows) // 2)]            # duplicates from sampler padding
    order = rng.permutation(len(rows))
    rows = [rows[i] for i in order]
    results = []
    for lo in range(0, len(rows), batch):
        part = rows[lo:lo + batch]
        results.append(dict(preds=np.stack([r[2] for r in part]), boxes=np.stack([r[3] for r in part]),
                            image_paths=[f'{IMG_PREFIX}{r[0]:012d}.jpg' for r in part],
                            bbox_ids=[r[1] for r in part], output_heatmap=None))
    name2id = {f'{im:012d}.jpg': 1000 + im for im in range(images)}
    return results, name2id


def run_reference(results, name2id, res_folder, use_nms=True, soft_nms=False, vis_thr=0.2, oks_thr=0.9, K=17,
                  sigmas=None, rle_score=False):
    cls = load_reference_dataset_class()
    ds = cls.__new__(cls)
    ds.name2id, ds.img_prefix = name2id, IMG_PREFIX
    ds.ann_info = dict(num_joints=K)
    ds.vis_thr, ds.oks_thr, ds.use_nms, ds.soft_nms = vis_thr, oks_thr, use_nms, soft_nms
    ds.sigmas = sigmas
    ds.classes = ['__background__', 'person']
    ds._class_to_coco_ind = dict(person=1)
    ds.coco = types.SimpleNamespace(dataset={})             # no annotations: evaluate() writes the file and returns {}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        out = ds.evaluate(results, res_folder=res_folder, metric='mAP', rle_score=rle_score)
    assert out == {}
    return os.path.join(res_folder, 'result_keypoints.json')


CASES = [dict(tag='hard', seed=0, use_nms=True, soft_nms=False), dict(tag='soft', seed=1, use_nms=True, soft_nms=True),
         dict(tag='nonms', seed=2, use_nms=False, soft_nms=False), dict(tag='rle', seed=3, use_nms=True, soft_nms=False,
                                                                       rle_score=True)]


def pack_results(results):
    return dict(preds=np.concatenate([r['preds'] for r in results]), boxes=np.concatenate([r['boxes'] for r in results]),
                image_ids=np.array([int(os.path.basename(p)[:-4]) for r in results for p in r['image_paths']]),
                bbox_ids=np.array([b for r in results for b in r['bbox_ids']]),
                batch=np.array([len(r['bbox_ids']) for r in results]))


def unpack_results(g, tag):
    out, lo = [], 0
    for n in g[f'{tag}_batch']:
        sl = slice(lo, lo + int(n))
        out.append(dict(preds=g[f'{tag}_preds'][sl], boxes=g[f'{tag}_boxes'][sl],
                        image_paths=[f'{IMG_PREFIX}{int(i):012d}.jpg' for i in g[f'{tag}_image_ids'][sl]],
                        bbox_ids=[int(b) for b in g[f'{tag}_bbox_ids'][sl]], output_heatmap=None))
        lo += int(n)
    images = int(g[f'{tag}_image_ids'].max()) + 1
    return out, {f'{im:012d}.jpg': 1000 + im for im in range(images)}


def main():
    import tempfile
    sigmas = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89]) / 10.0
    out, texts = {}, {}
    for c in CASES:
        results, name2id = synthetic_results(c['seed'])
        with tempfile.TemporaryDirectory() as d:
            path = run_reference(results, name2id, d, use_nms=c['use_nms'], soft_nms=c['soft_nms'], sigmas=sigmas,
                                 rle_score=c.get('rle_score', False))
            texts[c['tag']] = open(path).read()
        for k, v in pack_results(results).items():
            out[f"{c['tag']}_{k}"] = v
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'results_case.npz'), **out)
    with open(os.path.join(ROOT, 'tests', 'golden', 'result_keypoints.json'), 'w') as f:
        json.dump(texts, f)
    print({k: len(v) for k, v in texts.items()})


if __name__ == '__main__':
    main()
