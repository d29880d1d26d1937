"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):  ``python -m oracle.make_golden``

The reference's own source files (loaded by path through oracle/ref_loader.py) produce every
output stored here; the oracle restatement and the CUDA path are then checked against these files
on machines that do not have the reference tree (the GPU box).

Fixtures
* decode_cases.npz   fp32 heatmaps (Gaussian peaks + noise + adversarial maps), centre/scale, and
                     the reference ``keypoints_from_heatmaps`` outputs for every mode on the path;
                     flip pairs + raw flipped maps and the reference's merged maps (with/without
                     shift_heatmap).
* kat.npz            the reference's known-answer tests restated as data
                     (tests/test_evaluation/test_top_down_eval.py:29-89).
* model_tiny_classic.npz / model_tiny_simple.npz
                     weights (fp16-representable), input crops, metas, and the reference
                     ``TopDown.forward_test`` heatmaps / preds / boxes for a tiny ViT + classic
                     (deconv) and simple (upsample + 3x3) decoder.
* loss_kat.npz       JointsMSELoss reference outputs (tests/test_losses/test_top_down_losses.py:27-41
                     plus a random case).
"""
import os

import numpy as np
import torch

from oracle import ref_loader
from vitpose_b200 import configs, synthetic

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')

DECODE_MODES = {
    'udp_dark': dict(use_udp=True, post_process='default', kernel=11),
    'default': dict(use_udp=False, post_process='default', kernel=11),
    'unbiased': dict(use_udp=False, post_process='unbiased', kernel=11),
    'none': dict(use_udp=False, post_process=None, kernel=11),
    'udp_dark_k5': dict(use_udp=True, post_process='default', kernel=5),
    'unbiased_k7': dict(use_udp=False, post_process='unbiased', kernel=7),
}


def adversarial_heatmaps(seed=0, n=4, k=9):
    hm = synthetic.gaussian_peak_heatmaps(n, k, seed)
    H, W = hm.shape[2:]
    hm[0, 0] = 0.0                                   # all-zero map: coords -1
    hm[0, 1] = -np.abs(hm[0, 1]) - 0.01              # all-negative map
    hm[1, 0, 10, 10] = 2.0
    hm[1, 0, 40, 30] = 2.0                           # exact tie: first index wins
    corners = [(0, 0), (0, W - 1), (H - 1, 0), (H - 1, W - 1), (1, 1), (H - 2, W - 2),
               (0, 20), (30, 0), (H - 1, 7), (12, W - 1), (2, 2), (H - 3, W - 3)]
    for i, (y, x) in enumerate(corners):
        hm[2 + i // k, i % k, y, x] = 3.0            # peaks on / next to every border
    hm[1, 1] = 1.0                                   # constant map (argmax 0, zero gradients)
    return hm


def make_decode_cases(ref):
    hm = adversarial_heatmaps()
    n, k = hm.shape[:2]
    metas = synthetic.synthetic_metas(n, k, seed=3)
    center = np.stack([m['center'] for m in metas]).astype(np.float32)
    scale = np.stack([m['scale'] for m in metas]).astype(np.float32)
    out = dict(heatmaps=hm, center=center, scale=scale)
    for name, kw in DECODE_MODES.items():
        with np.errstate(all='ignore'):
            p, m = ref.keypoints_from_heatmaps(hm, center, scale, **kw)
        out[f'preds_{name}'] = p.astype(np.float32)
        out[f'maxvals_{name}'] = m.astype(np.float32)
    # flip merge: raw flipped-pass maps -> flip_back (-> shift) -> average
    pairs = [[1, 2], [3, 4], [5, 8]]
    raw_f = synthetic.gaussian_peak_heatmaps(n, k, seed=11)
    out['flip_pairs'] = np.asarray(pairs, dtype=np.int64)
    out['heatmaps_flipped_raw'] = raw_f
    for shift in (False, True):
        fb = ref.flip_back(raw_f.copy(), pairs)
        if shift:
            fb[:, :, :, 1:] = fb[:, :, :, :-1]
        merged = (hm + fb) * 0.5
        out[f'merged_shift{int(shift)}'] = merged.astype(np.float32)
        with np.errstate(all='ignore'):
            p, m = ref.keypoints_from_heatmaps(merged, center, scale, **DECODE_MODES['udp_dark'])
        out[f'merged_shift{int(shift)}_preds_udp_dark'] = p.astype(np.float32)
        p, m = ref.keypoints_from_heatmaps(merged, center, scale, **DECODE_MODES['default'])
        out[f'merged_shift{int(shift)}_preds_default'] = p.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, 'decode_cases.npz'), **out)


def make_kat(ref):
    """The reference's KATs as data: a 64x64 all-ones map with a 2.0 peak at (31,31)."""
    hm = np.ones((1, 1, 64, 64), dtype=np.float32)
    hm[0, 0, 31, 31] = 2
    c = np.array([[127, 127]], dtype=np.float32)
    s = np.array([[64 / 200.0, 64 / 200.0]], dtype=np.float32)
    out = dict(heatmaps=hm, center=c, scale=s)
    out['preds_default'], out['maxvals_default'] = ref.keypoints_from_heatmaps(hm, c, s)
    out['preds_unbiased'], _ = ref.keypoints_from_heatmaps(hm, c, s, post_process='unbiased')
    hm2 = np.ones((32, 17, 64, 64), dtype=np.float32)
    hm2[:, :, 31, 31] = 2
    c2 = np.tile([127, 127], (32, 1)).astype(np.float32)
    s2 = np.tile([32, 32], (32, 1)).astype(np.float32)
    out['udp_center'], out['udp_scale'] = c2, s2
    p, m = ref.keypoints_from_heatmaps(hm2, c2, s2, use_udp=True)
    out['preds_udp'], out['maxvals_udp'] = p.astype(np.float32), m.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, 'kat.npz'), **out)


def _fp16_round(sd):
    out = {}
    for k, v in sd.items():
        out[k] = v.half().float() if v.is_floating_point() else v
    # keep BN variance safely positive after rounding
    return out


def make_model_fixture(ref, name, cfg, n=2, seed=0):
    sd = _fp16_round(synthetic.scaled_init_state_dict(cfg, seed))
    model = ref_loader.build_reference_topdown(cfg)
    model.load_state_dict(sd, strict=True)
    K = cfg['keypoint_head']['out_channels']
    img = synthetic.synthetic_crops(n, seed).half().float()
    metas = synthetic.synthetic_metas(n, K, seed)
    out = {'w:' + k: (v.numpy().astype(np.float16) if v.is_floating_point() else v.numpy())
           for k, v in sd.items()}
    out['img'] = img.numpy().astype(np.float16)
    out['center'] = np.stack([m['center'] for m in metas])
    out['scale'] = np.stack([m['scale'] for m in metas])
    out['flip_pairs'] = np.asarray(metas[0]['flip_pairs'], dtype=np.int64)
    with torch.no_grad():
        feat = model.backbone(img)
        raw = model.keypoint_head(feat)
        out['features'] = feat.numpy()
        out['heatmaps_noflip'] = raw.numpy()
        for tag, tc in (('udp', configs.TEST_CFG_UDP), ('shift', configs.TEST_CFG_SHIFT),
                        ('unbiased', dict(flip_test=True, post_process='unbiased',
                                          shift_heatmap=False, modulate_kernel=11))):
            model.test_cfg = dict(tc)
            model.keypoint_head.test_cfg = dict(tc)
            r = model(img=img, img_metas=metas, return_loss=False, return_heatmap=True)
            out[f'{tag}_heatmap'] = r['output_heatmap']
            out[f'{tag}_preds'] = r['preds']
            out[f'{tag}_boxes'] = r['boxes']
    np.savez_compressed(os.path.join(OUT, f'model_{name}.npz'), **out)
    print(name, 'heatmap std', float(out['udp_heatmap'].std()),
          'params', sum(v.numel() for v in sd.values()))


def make_loss_kat(ref):
    out = {}
    loss = ref.JointsMSELoss(use_target_weight=True)
    g = torch.Generator().manual_seed(5)
    o = torch.rand(3, 4, 8, 6, generator=g)
    t = torch.rand(3, 4, 8, 6, generator=g)
    w = torch.rand(3, 4, 1, generator=g)
    out['output'], out['target'], out['weight'] = o.numpy(), t.numpy(), w.numpy()
    out['loss_weighted'] = loss(o, t, w).numpy()
    out['loss_unweighted'] = ref.JointsMSELoss(use_target_weight=False)(o, t, None).numpy()
    # reference KATs: zeros vs zeros -> 0; ones vs zeros -> 1; ones*w(0.5)... (test_top_down_losses.py)
    z = torch.zeros(1, 3, 64, 64)
    one = torch.ones(1, 3, 64, 64)
    out['kat_zero'] = loss(z, z, torch.ones(1, 3, 1)).numpy()
    out['kat_one'] = loss(one, z, torch.ones(1, 3, 1)).numpy()
    out['kat_w0'] = loss(one, z, torch.zeros(1, 3, 1)).numpy()
    np.savez_compressed(os.path.join(OUT, 'loss_kat.npz'), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = ref_loader.load_reference()
    make_decode_cases(ref)
    make_kat(ref)
    make_loss_kat(ref)
    make_model_fixture(ref, 'tiny_classic', configs.tiny_model_cfg(5, 'classic'))
    make_model_fixture(ref, 'tiny_simple', configs.tiny_model_cfg(5, 'simple', depth=1))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == '__main__':
    main()
