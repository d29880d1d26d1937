"""Pins the oracle (oracle/decode_np.py, oracle/vitpose_torch.py) before anything trusts it:
(1) the reference's own known-answer tests, (2) golden vectors produced by the unmodified
reference (oracle/make_golden.py), (3) the live reference code when /root/reference is mounted."""
import os

import numpy as np
import pytest
import torch

from oracle import decode_np as O
from oracle import vitpose_torch as VT
from oracle.make_golden import DECODE_MODES
from vitpose_b200 import configs, synthetic

COORD_TOL = 2e-4   # px in image space (coords ~1e2; float32 ulp there is 1.5e-5)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


# ---- (1) reference KATs: tests/test_evaluation/test_top_down_eval.py:29-89 -------------------
def test_kat_keypoints_from_heatmaps():
    heatmaps = np.ones((1, 1, 64, 64), dtype=np.float32)
    heatmaps[0, 0, 31, 31] = 2
    center = np.array([[127, 127]])
    scale = np.array([[64 / 200.0, 64 / 200.0]])
    preds, maxvals = O.keypoints_from_heatmaps(heatmaps, center, scale)
    np.testing.assert_array_almost_equal(preds, np.array([[[126, 126]]]), decimal=4)
    np.testing.assert_array_almost_equal(maxvals, np.array([[[2]]]), decimal=4)
    assert isinstance(preds, np.ndarray) and isinstance(maxvals, np.ndarray)
    with pytest.raises(AssertionError):
        O.keypoints_from_heatmaps(heatmaps, center, scale, post_process='unbiased', kernel=0)
    preds, maxvals = O.keypoints_from_heatmaps(heatmaps, center, scale, post_process='unbiased')
    np.testing.assert_array_almost_equal(preds, np.array([[[126, 126]]]), decimal=4)
    np.testing.assert_array_almost_equal(maxvals, np.array([[[2]]]), decimal=4)


def test_kat_udp():
    heatmaps = np.ones((32, 17, 64, 64), dtype=np.float32)
    heatmaps[:, :, 31, 31] = 2
    center = np.tile([127, 127], (32, 1))
    scale = np.tile([32, 32], (32, 1))
    preds, maxvals = O.keypoints_from_heatmaps(heatmaps, center, scale, post_process='default',
                                               target_type='GaussianHeatMap', use_udp=True)
    np.testing.assert_array_almost_equal(preds, np.tile([76, 76], [32, 17, 1]), decimal=0)
    np.testing.assert_array_almost_equal(maxvals, np.tile([2], [32, 17, 1]), decimal=4)
    for tt in ('GaussianHeatMap', 'gaussianheatmap'):   # case-insensitive target_type (:72-87)
        O.keypoints_from_heatmaps(heatmaps, center, scale, use_udp=True, target_type=tt)
    with pytest.raises(ValueError):
        O.keypoints_from_heatmaps(heatmaps, center, scale, use_udp=True, target_type='nope')


# ---- reference KATs: tests/test_post_processing.py:34-65 --------------------------------------
def test_kat_flip_back():
    heatmaps = np.random.random([1, 2, 32, 32])
    flipped = O.flip_back(heatmaps, [[0, 1]])
    np.testing.assert_array_almost_equal(heatmaps, O.flip_back(flipped, [[0, 1]]))
    np.testing.assert_array_almost_equal(heatmaps[:, 0], flipped[:, 1, :, ::-1])


def test_kat_transform_preds():
    coords = np.random.random([2, 2])
    center = np.array([50, 50])
    scale = np.array([100 / 200.0, 100 / 200.0])
    size = np.array([100, 100])
    np.testing.assert_array_almost_equal(coords, O.transform_preds(coords, center, scale, size))
    coords = np.random.random([2, 2])
    center = np.array([50, 50])
    scale = np.array([100 / 200.0, 100 / 200.0])
    size = np.array([101, 101])
    np.testing.assert_array_almost_equal(
        coords, O.transform_preds(coords, center, scale, size, use_udp=True))


# ---- reference KATs: tests/test_losses/test_top_down_losses.py:27-41 --------------------------
def test_kat_joints_mse():
    z, one = torch.zeros(1, 3, 64, 64), torch.ones(1, 3, 64, 64)
    w = torch.ones(1, 3, 1)
    assert torch.allclose(VT.joints_mse_loss(z, z, w), torch.tensor(0.))
    assert torch.allclose(VT.joints_mse_loss(one, z, w), torch.tensor(1.))
    p = torch.zeros(1, 2, 64, 64)
    p[0, 0] += 1
    assert torch.allclose(VT.joints_mse_loss(p, torch.zeros(1, 2, 64, 64), None, False),
                          torch.tensor(0.5))


# ---- (2) golden vectors from the unmodified reference ----------------------------------------
def test_golden_kat_file(golden_dir):
    g = _load(golden_dir, 'kat.npz')
    p, m = O.keypoints_from_heatmaps(g['heatmaps'], g['center'], g['scale'])
    np.testing.assert_allclose(p, g['preds_default'], atol=COORD_TOL)
    np.testing.assert_array_equal(m, g['maxvals_default'])
    p, _ = O.keypoints_from_heatmaps(g['heatmaps'], g['center'], g['scale'], post_process='unbiased')
    np.testing.assert_allclose(p, g['preds_unbiased'], atol=COORD_TOL)


@pytest.mark.parametrize('mode', sorted(DECODE_MODES))
def test_golden_decode_modes(golden_dir, mode):
    g = _load(golden_dir, 'decode_cases.npz')
    with np.errstate(all='ignore'):
        p, m = O.keypoints_from_heatmaps(g['heatmaps'], g['center'], g['scale'], **DECODE_MODES[mode])
    np.testing.assert_array_equal(m, g[f'maxvals_{mode}'])
    ref = g[f'preds_{mode}']
    assert np.array_equal(np.isfinite(p), np.isfinite(ref))
    ok = np.isfinite(ref)
    np.testing.assert_allclose(p[ok], ref[ok], atol=COORD_TOL)


@pytest.mark.parametrize('shift', [0, 1])
def test_golden_flip_merge(golden_dir, shift):
    g = _load(golden_dir, 'decode_cases.npz')
    merged = O.merge_flip(g['heatmaps'], g['heatmaps_flipped_raw'], g['flip_pairs'].tolist(), bool(shift))
    np.testing.assert_array_equal(merged, g[f'merged_shift{shift}'])          # bit-exact


@pytest.mark.parametrize('name,decoder', [('tiny_classic', 'classic'), ('tiny_simple', 'simple')])
def test_golden_model(golden_dir, name, decoder):
    g = _load(golden_dir, f'model_{name}.npz')
    sd = {k[2:]: torch.from_numpy(g[k].astype(np.float32) if g[k].dtype == np.float16 else g[k])
          for k in g.files if k.startswith('w:')}
    cfg = configs.tiny_model_cfg(5, decoder, depth=2 if decoder == 'classic' else 1)
    img = torch.from_numpy(g['img'].astype(np.float32))
    K = 5
    metas = [dict(center=g['center'][i], scale=g['scale'][i], image_file='', bbox_id=i,
                  bbox_score=1.0, flip_pairs=g['flip_pairs'].tolist()) for i in range(img.shape[0])]
    with torch.no_grad():
        feat = VT.vit_features(sd, img, cfg['backbone']['depth'], cfg['backbone']['num_heads'])
        np.testing.assert_allclose(feat.numpy(), g['features'], atol=2e-5)
        np.testing.assert_allclose(VT.model_heatmaps(sd, img, cfg).numpy(), g['heatmaps_noflip'], atol=2e-6)
    for tag, tc in (('udp', configs.TEST_CFG_UDP), ('shift', configs.TEST_CFG_SHIFT),
                    ('unbiased', dict(flip_test=True, post_process='unbiased', shift_heatmap=False,
                                      modulate_kernel=11))):
        c = dict(cfg, test_cfg=dict(tc))
        r = VT.forward_test(sd, img, metas, c, return_heatmap=True)
        np.testing.assert_allclose(r['output_heatmap'], g[f'{tag}_heatmap'], atol=2e-6)
        np.testing.assert_allclose(r['boxes'], g[f'{tag}_boxes'], rtol=1e-6)
        np.testing.assert_array_equal(r['preds'][..., 2], g[f'{tag}_preds'][..., 2])
        # coordinates on these noise-dominated random-weight heatmaps: refinement amplifies the
        # 1e-7 blur differences, so allow 0.02 px here (tight tolerances are on the peaked set)
        np.testing.assert_allclose(r['preds'][..., :2], g[f'{tag}_preds'][..., :2], atol=2e-2)


def test_golden_loss(golden_dir):
    g = _load(golden_dir, 'loss_kat.npz')
    o, t, w = (torch.from_numpy(g[k]) for k in ('output', 'target', 'weight'))
    np.testing.assert_allclose(VT.joints_mse_loss(o, t, w).numpy(), g['loss_weighted'], rtol=1e-6)
    np.testing.assert_allclose(VT.joints_mse_loss(o, t, None, False).numpy(), g['loss_unweighted'], rtol=1e-6)


def test_kat_pose_pck_accuracy():
    """tests/test_evaluation/test_top_down_eval.py:11-26."""
    output = np.zeros((1, 5, 64, 64), dtype=np.float32)
    target = np.zeros((1, 5, 64, 64), dtype=np.float32)
    mask = np.array([[True, True, False, False, False]])
    output[0, 0, 20, 20] = 1
    target[0, 0, 10, 10] = 1
    output[0, 1, 30, 30] = 1
    target[0, 1, 30, 30] = 1
    acc, avg_acc, cnt = O.pose_pck_accuracy(output, target, mask)
    np.testing.assert_array_almost_equal(acc, np.array([0, 1, -1, -1, -1]), decimal=4)
    assert abs(avg_acc - 0.5) < 1e-4 and abs(cnt - 2) < 1e-4


@pytest.mark.reference
def test_live_reference_pose_pck_accuracy():
    from oracle import ref_loader
    ref = ref_loader.load_reference()
    rng = np.random.RandomState(3)
    N, K = 9, 17
    target = np.zeros((N, K, 64, 48), dtype=np.float32)
    output = rng.randn(N, K, 64, 48).astype(np.float32) * 0.05
    for n in range(N):
        for k in range(K):
            y, x = rng.randint(64), rng.randint(48)
            target[n, k, y, x] = 1
            oy, ox = np.clip(y + rng.randint(-4, 5), 0, 63), np.clip(x + rng.randint(-4, 5), 0, 47)
            output[n, k, oy, ox] += 1
    target[0, 3] = 0                      # a map whose maximum is <= 0 decodes to (-1, -1)
    mask = rng.rand(N, K) > 0.25
    mask[:, 5] = False                    # a keypoint that is never visible
    a0, v0, c0 = ref.pose_pck_accuracy(output, target, mask.copy())
    a1, v1, c1 = O.pose_pck_accuracy(output, target, mask.copy())
    np.testing.assert_array_equal(a0, a1)
    assert v0 == v1 and c0 == c1


def test_gaussian_taps_match_cv2():
    cv2 = pytest.importorskip('cv2')
    for k in (1, 3, 5, 7, 9, 11, 17):
        np.testing.assert_array_equal(O.gaussian_taps(k), cv2.getGaussianKernel(k, 0, cv2.CV_32F)[:, 0])
    x = np.random.RandomState(0).rand(64, 48).astype(np.float32)
    np.testing.assert_allclose(O.blur_maps(x, 11, 'reflect101'), cv2.GaussianBlur(x, (11, 11), 0), atol=1e-6)


# ---- (3) the live reference ---------------------------------------------------------------
@pytest.mark.reference
@pytest.mark.parametrize('seed', [0, 1])
def test_live_reference_decode(seed):
    from oracle import ref_loader
    ref = ref_loader.load_reference()
    hm = synthetic.gaussian_peak_heatmaps(5, 17, seed)
    hm[0, 0] = 0
    hm[1, 3] = -np.abs(hm[1, 3])
    metas = synthetic.synthetic_metas(5, 17, seed)
    c = np.stack([m['center'] for m in metas])
    s = np.stack([m['scale'] for m in metas])
    for kw in DECODE_MODES.values():
        with np.errstate(all='ignore'):
            p1, m1 = ref.keypoints_from_heatmaps(hm, c, s, **kw)
            p2, m2 = O.keypoints_from_heatmaps(hm, c, s, **kw)
        np.testing.assert_array_equal(m1, m2)
        ok = np.isfinite(p1)
        np.testing.assert_allclose(p1[ok], p2[ok], atol=COORD_TOL)


@pytest.mark.reference
def test_live_reference_model_small_classic():
    """ViTPose-S classic (BASELINE configs[0] architecture) on 2 crops: oracle == reference."""
    from oracle import ref_loader
    cfg = configs.baseline_model_cfg('S-classic-17')
    model = ref_loader.build_reference_topdown(cfg)
    sd = synthetic.scaled_init_state_dict(cfg, 1)
    model.load_state_dict(sd, strict=True)
    img = synthetic.synthetic_crops(2, 1)
    metas = synthetic.synthetic_metas(2, 17, 1)
    with torch.no_grad():
        r1 = model(img=img, img_metas=metas, return_loss=False, return_heatmap=True)
    r2 = VT.forward_test(sd, img, metas, cfg, return_heatmap=True)
    np.testing.assert_allclose(r1['output_heatmap'], r2['output_heatmap'], atol=1e-5)
    np.testing.assert_allclose(r1['preds'], r2['preds'], atol=1e-3)
