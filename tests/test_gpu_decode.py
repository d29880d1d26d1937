"""Decode parity (-m gpu): the fused CUDA decode, called through the C ABI, against the oracle
(oracle/decode_np.py) and against golden vectors produced by the unmodified reference.
Bar: argmax indices and maxvals bit-exact; coordinates within 1e-3 px (float32 blur/log ulps)."""
import os

import numpy as np
import pytest
import torch

from oracle import decode_np as O
from oracle.make_golden import DECODE_MODES, adversarial_heatmaps
from vitpose_b200 import synthetic

pytestmark = pytest.mark.gpu

COORD_TOL = 1e-3


def _gpu_decode(hm, center, scale, mode_kw, hm_f=None, pairs=None, shift=False, want_merged=False):
    from vitpose_b200 import ops
    from vitpose_b200.engine import resolve_decode_mode
    dev = torch.device('cuda:0')
    mode = resolve_decode_mode(mode_kw.get('post_process', 'default'), False, mode_kw.get('use_udp', False))
    K = hm.shape[1]
    fi = None
    if hm_f is not None:
        fi = torch.from_numpy(O.flip_index(K, pairs).astype(np.int32)).to(dev)
    r = ops.decode(torch.from_numpy(hm).to(dev), None if hm_f is None else torch.from_numpy(hm_f).to(dev), fi, shift,
                   mode, mode_kw.get('kernel', 11), mode_kw.get('use_udp', False),
                   torch.from_numpy(center.astype(np.float32)).to(dev),
                   torch.from_numpy(scale.astype(np.float32)).to(dev), want_merged=want_merged, want_argmax=True)
    torch.cuda.synchronize()
    return {k: (v.cpu().numpy() if v is not None else None) for k, v in r.items()}


def _compare(r, preds_ref, maxvals_ref, hm):
    N, K, H, W = hm.shape
    np.testing.assert_array_equal(r['argmax'], hm.reshape(N, K, -1).argmax(2).astype(np.int32))   # bit-exact
    np.testing.assert_array_equal(r['maxvals'], maxvals_ref.astype(np.float32))                  # bit-exact
    ok = np.isfinite(preds_ref)
    assert np.array_equal(np.isfinite(r['preds']), ok)
    np.testing.assert_allclose(r['preds'][ok], preds_ref[ok], atol=COORD_TOL, rtol=0)


@pytest.mark.parametrize('mode', sorted(DECODE_MODES))
def test_decode_vs_golden(golden_dir, mode):
    g = np.load(os.path.join(golden_dir, 'decode_cases.npz'))
    r = _gpu_decode(g['heatmaps'], g['center'], g['scale'], DECODE_MODES[mode])
    _compare(r, g[f'preds_{mode}'], g[f'maxvals_{mode}'], g['heatmaps'])


@pytest.mark.parametrize('shift', [0, 1])
def test_flip_merge_vs_golden(golden_dir, shift):
    g = np.load(os.path.join(golden_dir, 'decode_cases.npz'))
    for mode in ('udp_dark', 'default'):
        r = _gpu_decode(g['heatmaps'], g['center'], g['scale'], DECODE_MODES[mode], g['heatmaps_flipped_raw'],
                        g['flip_pairs'].tolist(), bool(shift), want_merged=True)
        np.testing.assert_array_equal(r['merged'], g[f'merged_shift{shift}'])                   # bit-exact average
        ref = g[f'merged_shift{shift}_preds_{mode}']
        ok = np.isfinite(ref)
        np.testing.assert_allclose(r['preds'][ok], ref[ok], atol=COORD_TOL, rtol=0)


def test_kat_through_cabi(golden_dir):
    """The reference's own KATs (tests/test_evaluation/test_top_down_eval.py:29-89) through the CUDA path."""
    g = np.load(os.path.join(golden_dir, 'kat.npz'))
    r = _gpu_decode(g['heatmaps'], g['center'], g['scale'], dict(post_process='default'))
    np.testing.assert_array_almost_equal(r['preds'], np.array([[[126, 126]]]), decimal=4)
    np.testing.assert_array_almost_equal(r['maxvals'], np.array([[[2]]]), decimal=4)
    r = _gpu_decode(g['heatmaps'], g['center'], g['scale'], dict(post_process='unbiased'))
    np.testing.assert_array_almost_equal(r['preds'], np.array([[[126, 126]]]), decimal=4)
    hm = np.ones((32, 17, 64, 64), dtype=np.float32)
    hm[:, :, 31, 31] = 2
    r = _gpu_decode(hm, g['udp_center'], g['udp_scale'], dict(use_udp=True))
    np.testing.assert_array_almost_equal(r['preds'], np.tile([76, 76], (32, 17, 1)), decimal=0)
    np.testing.assert_allclose(r['preds'], g['preds_udp'], atol=COORD_TOL)
    np.testing.assert_array_almost_equal(r['maxvals'], np.tile([2], (32, 17, 1)), decimal=4)


@pytest.mark.parametrize('seed,K', [(0, 17), (1, 133), (2, 5)])
@pytest.mark.parametrize('mode', ['udp_dark', 'default', 'unbiased', 'none'])
def test_decode_vs_oracle_random(seed, K, mode):
    n = 6
    hm = synthetic.gaussian_peak_heatmaps(n, K, seed)
    hm[0, 0] = 0
    hm[n - 1, K - 1] = -1.0
    metas = synthetic.synthetic_metas(n, K, seed)
    c = np.stack([m['center'] for m in metas])
    s = np.stack([m['scale'] for m in metas])
    with np.errstate(all='ignore'):
        p, m = O.keypoints_from_heatmaps(hm, c, s, **DECODE_MODES[mode])
    r = _gpu_decode(hm, c, s, DECODE_MODES[mode])
    _compare(r, p, m, hm)


@pytest.mark.parametrize('kernel,width,height', [(11, 48, 64), (11, 46, 61), (7, 48, 64), (11, 36, 44), (11, 64, 64)])
def test_unbiased_blur_paths_vs_oracle(kernel, width, height):
    """post_process='unbiased' runs a register-blocked 11-tap blur when W % 4 == 0 and a generic loop otherwise;
    both must reproduce the oracle (cv2.GaussianBlur on the zero-bordered map) for ragged sizes and other kernels."""
    n, K = 4, 9
    hm = synthetic.gaussian_peak_heatmaps(n, K, 11, height=height, width=width)
    metas = synthetic.synthetic_metas(n, K, 11)
    c = np.stack([m['center'] for m in metas])
    s = np.stack([m['scale'] for m in metas])
    kw = dict(post_process='unbiased', kernel=kernel)
    with np.errstate(all='ignore'):
        p, m = O.keypoints_from_heatmaps(hm, c, s, **kw)
    r = _gpu_decode(hm, c, s, kw)
    _compare(r, p, m, hm)


@pytest.mark.parametrize('mode', ['udp_dark', 'default'])
def test_decode_with_flip_vs_oracle(mode):
    n, K = 5, 17
    from vitpose_b200.configs import COCO17_FLIP_PAIRS
    hm = synthetic.gaussian_peak_heatmaps(n, K, 3)
    hm_f = synthetic.gaussian_peak_heatmaps(n, K, 4)
    metas = synthetic.synthetic_metas(n, K, 3)
    c = np.stack([m['center'] for m in metas])
    s = np.stack([m['scale'] for m in metas])
    for shift in (False, True):
        merged = O.merge_flip(hm, hm_f, COCO17_FLIP_PAIRS, shift)
        p, m = O.keypoints_from_heatmaps(merged, c, s, **DECODE_MODES[mode])
        r = _gpu_decode(hm, c, s, DECODE_MODES[mode], hm_f, COCO17_FLIP_PAIRS, shift, want_merged=True)
        np.testing.assert_array_equal(r['merged'], merged)
        _compare(r, p, m, merged)


def test_decode_full_size_properties():
    """BASELINE size (256 crops x 17, and 64 x 133): properties that need no oracle run —
    argmax == torch.argmax on the device-merged map, flip-merge involution, idempotent outputs."""
    dev = torch.device('cuda:0')
    from vitpose_b200 import ops, _lib
    for n, K in ((256, 17), (64, 133)):
        g = torch.Generator(device='cuda').manual_seed(n)
        hm = torch.rand(n, K, 64, 48, device=dev, generator=g)
        hm_f = torch.rand(n, K, 64, 48, device=dev, generator=g)
        fi = torch.arange(K, device=dev, dtype=torch.int32)
        c = torch.rand(n, 2, device=dev) * 100
        s = torch.rand(n, 2, device=dev) + 0.5
        r = ops.decode(hm, hm_f, fi, False, _lib.DECODE_UDP_DARK, 11, True, c, s, want_merged=True, want_argmax=True)
        merged = (hm + hm_f.flip(3)) * 0.5
        assert torch.equal(r['merged'], merged)
        assert torch.equal(r['argmax'].long(), merged.reshape(n, K, -1).argmax(2))
        assert torch.equal(r['maxvals'][..., 0], merged.reshape(n, K, -1).amax(2))
        r2 = ops.decode(merged, None, None, False, _lib.DECODE_UDP_DARK, 11, True, c, s)
        assert torch.equal(r['preds'], r2['preds'])           # fused merge == decode of the merged map
        assert torch.isfinite(r['preds']).all()


def test_decode_config4_full_size_vs_oracle_subset():
    """BASELINE configs[3] decode size: 2048 crops x 133 keypoints (272 384 maps), shift_heatmap + quarter offset as
    shipped. Whole batch: batch-split invariance (bitwise); a random subset of crops: against the oracle."""
    from vitpose_b200 import ops, _lib
    from vitpose_b200.configs import WHOLEBODY133_FLIP_PAIRS
    from vitpose_b200.core.post_processing import flip_index_from_pairs
    dev = torch.device('cuda:0')
    n, K = 2048, 133
    g = torch.Generator(device='cuda').manual_seed(4)
    hm = torch.rand(n, K, 64, 48, device=dev, generator=g)
    hm_f = torch.rand(n, K, 64, 48, device=dev, generator=g)
    # a sharp peak per map so that the decode is well-posed
    idx = torch.randint(0, 64 * 48, (n, K), device=dev, generator=g)
    hm.view(n, K, -1).scatter_(2, idx[..., None], 3.0)
    perm = flip_index_from_pairs(K, WHOLEBODY133_FLIP_PAIRS)
    fi = torch.from_numpy(perm).to(dev)
    c = torch.rand(n, 2, device=dev, generator=g) * 100 + 50
    s = torch.rand(n, 2, device=dev, generator=g) + 0.5
    full = ops.decode(hm, hm_f, fi, True, _lib.DECODE_DEFAULT, 11, False, c, s)
    for lo, hi in ((0, 1024), (1024, 2048), (777, 779)):
        part = ops.decode(hm[lo:hi].contiguous(), hm_f[lo:hi].contiguous(), fi, True, _lib.DECODE_DEFAULT, 11, False,
                          c[lo:hi].contiguous(), s[lo:hi].contiguous())
        assert torch.equal(part['preds'], full['preds'][lo:hi]) and torch.equal(part['maxvals'], full['maxvals'][lo:hi])
    pick = np.random.RandomState(0).choice(n, 24, replace=False)
    merged = O.merge_flip(hm[pick].cpu().numpy(), hm_f[pick].cpu().numpy(), WHOLEBODY133_FLIP_PAIRS, True)
    p_ref, m_ref = O.keypoints_from_heatmaps(merged, c[pick].cpu().numpy(), s[pick].cpu().numpy(),
                                             post_process='default', use_udp=False)
    np.testing.assert_array_equal(full['maxvals'][pick].cpu().numpy(), m_ref)
    np.testing.assert_allclose(full['preds'][pick].cpu().numpy(), p_ref, atol=1e-3)


def test_decode_empty_and_errors():
    from vitpose_b200 import ops, _lib
    dev = torch.device('cuda:0')
    r = ops.decode(torch.zeros(0, 17, 64, 48, device=dev), mode=_lib.DECODE_DEFAULT)
    assert r['preds'].shape == (0, 17, 2)
    with pytest.raises(_lib.VitposeLibError):
        ops.decode(torch.zeros(1, 1, 64, 48, device=dev), mode=_lib.DECODE_UNBIASED, kernel=0)
    with pytest.raises(_lib.VitposeLibError):
        ops.decode(torch.zeros(1, 1, 64, 48), mode=_lib.DECODE_DEFAULT)          # CPU tensor: no fallback


def test_pose_pck_accuracy_matches_oracle():
    """acc_pose of forward_train (simple_head.py:170-195): device arg-max + PCK arithmetic vs the oracle, incl. the
    reference KAT (tests/test_evaluation/test_top_down_eval.py:11-26), invisible keypoints and empty maps."""
    import vitpose_b200 as V
    from vitpose_b200.core.evaluation.top_down_eval import pose_pck_accuracy
    output = np.zeros((1, 5, 64, 64), dtype=np.float32)
    target = np.zeros((1, 5, 64, 64), dtype=np.float32)
    mask = np.array([[True, True, False, False, False]])
    output[0, 0, 20, 20] = 1
    target[0, 0, 10, 10] = 1
    output[0, 1, 30, 30] = 1
    target[0, 1, 30, 30] = 1
    acc, avg_acc, cnt = pose_pck_accuracy(output, target, mask)
    np.testing.assert_array_almost_equal(acc, np.array([0, 1, -1, -1, -1]), decimal=4)
    assert abs(avg_acc - 0.5) < 1e-4 and cnt == 2
    rng = np.random.RandomState(5)
    N, K = 33, 17
    target = np.zeros((N, K, 64, 48), dtype=np.float32)
    output = rng.randn(N, K, 64, 48).astype(np.float32) * 0.05
    for n in range(N):
        for k in range(K):
            y, x = rng.randint(64), rng.randint(48)
            target[n, k, y, x] = 1
            output[n, k, np.clip(y + rng.randint(-4, 5), 0, 63), np.clip(x + rng.randint(-4, 5), 0, 47)] += 1
    target[0, 3] = 0
    mask = rng.rand(N, K) > 0.25
    mask[:, 5] = False
    a0, v0, c0 = O.pose_pck_accuracy(output, target, mask)
    a1, v1, c1 = pose_pck_accuracy(output, target, mask)
    np.testing.assert_allclose(a1, a0, rtol=0, atol=1e-7)
    assert abs(v1 - v0) < 1e-6 and c1 == c0
