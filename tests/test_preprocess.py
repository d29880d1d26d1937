"""Preprocessing widening: oracle pinned against cv2 / the reference / golden vectors on
CPU; the fused CUDA warp + normalise kernel bit-exact against the oracle on the GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import preprocess_np as P

MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def test_warp_restatement_bit_exact_vs_cv2():
    cv2 = pytest.importorskip('cv2')
    rng = np.random.RandomState(0)
    for t in range(12):
        h, w = rng.randint(150, 500), rng.randint(150, 600)
        img = rng.randint(0, 256, size=(h, w, 3)).astype(np.uint8)
        box = [rng.uniform(-20, 0.6 * w), rng.uniform(-20, 0.6 * h), rng.uniform(20, w), rng.uniform(20, h)]
        c, s = P.box2cs(box)
        size = np.array([192, 256])
        m = P.get_warp_matrix(rng.uniform(-40, 40) if t % 3 == 0 else 0., c * 2.0, size - 1.0, s * 200.0)
        np.testing.assert_array_equal(P.warp_affine_linear_u8(img, m, (192, 256)),
                                      cv2.warpAffine(img, m, (192, 256), flags=cv2.INTER_LINEAR))
        m2 = P.get_affine_transform(c, s, 0., size)
        np.testing.assert_array_equal(P.warp_affine_linear_u8(img, m2, (192, 256)),
                                      cv2.warpAffine(img, m2, (192, 256), flags=cv2.INTER_LINEAR))


def test_golden_preprocess(golden_dir):
    g = np.load(os.path.join(golden_dir, 'preprocess.npz'))
    for tag, udp in (('udp', True), ('affine', False)):
        for i, box in enumerate(g['boxes']):
            crop, c, s, trans = P.preprocess_crop(g['img'], box, use_udp=udp)
            np.testing.assert_array_equal(np.concatenate([c, s]), g[f'{tag}_cs'][i])
            np.testing.assert_allclose(trans, g[f'{tag}_mats'][i], rtol=1e-12 if udp else 1e-9, atol=1e-9)
            np.testing.assert_array_equal(P.warp_affine_linear_u8(g['img'], g[f'{tag}_mats'][i], (192, 256)),
                                          g[f'{tag}_warps_u8'][i])
            if i == 0:
                np.testing.assert_array_equal(P.to_tensor_normalize(g[f'{tag}_warps_u8'][0], MEAN, STD),
                                              g[f'{tag}_crop0_f32'])


@pytest.mark.reference
def test_matrices_vs_live_reference():
    import sys
    from oracle import ref_loader
    ref_loader.load_reference()
    post = sys.modules['mmpose.core.post_processing']
    rng = np.random.RandomState(1)
    for _ in range(20):
        c = rng.uniform(0, 500, 2).astype(np.float32)
        s = rng.uniform(0.2, 3, 2).astype(np.float32)
        size = np.array([192, 256])
        np.testing.assert_array_equal(P.get_warp_matrix(0, c * 2.0, size - 1.0, s * 200.0),
                                      post.get_warp_matrix(0, c * 2.0, size - 1.0, s * 200.0))
        np.testing.assert_allclose(P.get_affine_transform(c, s, 0, size), post.get_affine_transform(c, s, 0, size),
                                   rtol=1e-9, atol=1e-9)


def test_host_mirror_matrices_and_boxes():
    from vitpose_b200 import pipelines as PL
    rng = np.random.RandomState(2)
    for _ in range(10):
        box = [rng.uniform(0, 300), rng.uniform(0, 300), rng.uniform(10, 200), rng.uniform(10, 300)]
        c1, s1 = P.box2cs(box)
        c2, s2 = PL.box2cs(box, (192, 256))
        np.testing.assert_array_equal(c1, c2)
        np.testing.assert_array_equal(s1, s2)
        size = np.array([192, 256])
        np.testing.assert_array_equal(PL.get_warp_matrix(0, c1 * 2.0, size - 1.0, s1 * 200.0),
                                      P.get_warp_matrix(0, c1 * 2.0, size - 1.0, s1 * 200.0))
        np.testing.assert_allclose(PL.get_affine_transform(c1, s1, 0, size), P.get_affine_transform(c1, s1, 0, size),
                                   rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize('use_udp', [True, False])
def test_gpu_preprocess_bit_exact(golden_dir, use_udp):
    from vitpose_b200 import pipelines as PL
    g = np.load(os.path.join(golden_dir, 'preprocess.npz'))
    rng = np.random.RandomState(5)
    big = rng.randint(0, 256, size=(333, 517, 3)).astype(np.uint8)
    images = [g['img'], big]
    boxes = [(0, b) for b in g['boxes']] + [(1, [40., 30., 200., 280.]), (1, [-30., -20., 90., 120.]),
                                            (1, [400., 250., 150., 100.])]
    dev = torch.device('cuda:0')
    imgs_dev = [torch.from_numpy(i).to(dev) for i in images]
    crops, metas = PL.preprocess_crops(imgs_dev, boxes, image_size=(192, 256), use_udp=use_udp, mean=MEAN, std=STD)
    torch.cuda.synchronize()
    assert crops.shape == (len(boxes), 3, 256, 192) and crops.dtype == torch.float32
    for i, (idx, box) in enumerate(boxes):
        ref, c, s, _ = P.preprocess_crop(images[idx], box, use_udp=use_udp)
        np.testing.assert_array_equal(crops[i].cpu().numpy(), ref)            # bit-exact
        np.testing.assert_array_equal(metas[i]['center'], c)
        np.testing.assert_array_equal(metas[i]['scale'], s)


@pytest.mark.parametrize('dtype', [np.float32, np.float64, 'pyfloat'])
def test_vectorised_host_maths_equals_scalar(dtype):
    """The batched box -> (center, scale, inverse map) maths is bit-identical to the per-box scalar path."""
    from vitpose_b200 import pipelines as PL
    rng = np.random.RandomState(3)
    raw = np.stack([rng.uniform(-50, 800, 300), rng.uniform(-50, 500, 300), rng.uniform(5, 600, 300),
                    rng.uniform(5, 700, 300)], axis=1)
    if dtype == 'pyfloat':
        boxes = [(0, [float(v) for v in r]) for r in raw]
    else:
        boxes = [(0, r.astype(dtype)) for r in raw]
    c1, s1, i1 = PL.box_transforms(boxes, (192, 256), True, 0., vectorize=True)
    c2, s2, i2 = PL.box_transforms(boxes, (192, 256), True, 0., vectorize=False)
    np.testing.assert_array_equal(np.stack(c1), np.stack(c2))
    np.testing.assert_array_equal(np.stack(s1), np.stack(s2))
    np.testing.assert_array_equal(i1, i2)


@pytest.mark.gpu
def test_preprocess_feeds_forward_test():
    """Boxes on a full image -> fused GPU preprocessing -> TopDown.forward_test, everything on the device."""
    import vitpose_b200 as V
    from vitpose_b200 import configs, pipelines as PL, synthetic
    dev = torch.device('cuda:0')
    cfg = configs.tiny_model_cfg(5)
    model = V.build_posenet(cfg)
    model.load_state_dict(synthetic.scaled_init_state_dict(cfg, 0))
    model = model.cuda().eval()
    img = torch.randint(0, 256, (480, 640, 3), device=dev, dtype=torch.uint8)
    boxes = [(0, [50., 40., 120., 300., 0.9]), (0, [300., 100., 200., 250., 0.8]), (0, [10., 10., 600., 460., 0.7])]
    crops, metas = PL.preprocess_crops([img], boxes, flip_pairs=configs.flip_pairs_for(5))
    r = model(img=crops, img_metas=metas, return_loss=False)
    assert r['preds'].shape == (3, 5, 3) and np.isfinite(r['preds']).all()
    np.testing.assert_allclose(r['boxes'][:, 5], [0.9, 0.8, 0.7], rtol=1e-6)
    # keypoints land inside (a padded version of) their boxes
    for i, (_, b) in enumerate(boxes):
        assert (r['preds'][i, :, 0] > b[0] - b[2]).all() and (r['preds'][i, :, 0] < b[0] + 2 * b[2]).all()
