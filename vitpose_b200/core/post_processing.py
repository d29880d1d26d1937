"""``flip_back`` / ``transform_preds`` with the reference's signatures
(mmpose/core/post_processing/post_transforms.py:110-147, :150-194), executed by the CUDA kernels behind
vpb_flip_back / vpb_transform_preds.  On the fused forward_test path neither is called: both are folded
into the decode kernel."""
import numpy as np
import torch

from .. import _lib, ops


def flip_index_from_pairs(num_keypoints, flip_pairs):
    """Channel permutation equivalent to the pairwise swaps of flip_back (post_transforms.py:138-141)."""
    perm = np.arange(num_keypoints, dtype=np.int32)
    for left, right in flip_pairs:
        perm[left], perm[right] = right, left
    return perm


def flip_back(output_flipped, flip_pairs, target_type='GaussianHeatmap'):
    """Flip the flipped heatmaps back to the original form.  np.ndarray[N,K,H,W] -> np.ndarray[N,K,H,W]."""
    assert output_flipped.ndim == 4, \
        'output_flipped should be [batch_size, num_keypoints, height, width]'
    if target_type.lower() != 'GaussianHeatmap'.lower():
        raise NotImplementedError("target_type='CombinedTarget' is outside the ViTPose hot path")
    _lib.require_cuda()
    dtype = output_flipped.dtype
    dev = torch.device('cuda')
    x = torch.from_numpy(np.ascontiguousarray(output_flipped, dtype=np.float32)).to(dev)
    perm = torch.from_numpy(flip_index_from_pairs(output_flipped.shape[1], flip_pairs)).to(dev)
    out = ops.flip_back(x, perm, False).cpu().numpy()
    return out if dtype == np.float32 else out.astype(dtype)


def transform_preds(coords, center, scale, output_size, use_udp=False):
    """coords np.ndarray[K, 2|4|5]; center, scale (2,); output_size (W, H) -> coordinates in the image."""
    assert coords.shape[1] in (2, 4, 5)
    assert len(center) == 2
    assert len(scale) == 2
    assert len(output_size) == 2
    _lib.require_cuda()
    dev = torch.device('cuda')
    xy = torch.from_numpy(np.ascontiguousarray(coords[None, :, :2], dtype=np.float32)).to(dev)
    c = torch.tensor(np.asarray(center, dtype=np.float32).reshape(1, 2), device=dev)
    s = torch.tensor(np.asarray(scale, dtype=np.float32).reshape(1, 2), device=dev)
    den = (float(output_size[0]), float(output_size[1]))
    if float(den[0]).is_integer() and float(den[1]).is_integer():
        out = ops.transform_preds(xy, c, s, (int(den[0]), int(den[1])), use_udp)[0].cpu().numpy()
    else:
        raise ValueError('output_size must be integral')
    target = np.ones_like(coords)
    target[:, :2] = out
    return target
