"""Test-time preprocessing in front of the hot path, batched on the GPU.

Host side (a few floats per box, same expressions as the reference): ``box2cs`` (mmpose/apis/inference.py:85-112),
``get_warp_matrix`` (mmpose/core/post_processing/post_transforms.py:312-340, UDP configs) and
``get_affine_transform`` (:197-252, non-UDP configs; the 3-point solve of cv2.getAffineTransform is a float64
linear solve). Device side: ``preprocess_crops`` runs TopDownAffine + ToTensor + NormalizeTensor for all boxes in
one launch of the fused warp/normalise kernel (vpb_warp_affine_normalize), bit-identical to
cv2.warpAffine(uint8, INTER_LINEAR) + torchvision's to_tensor / normalize.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr


def box2cs(box, image_size):
    """bbox (x, y, w, h) -> (center float32[2], scale float32[2]) with the model's aspect ratio and 1.25 padding."""
    x, y, w, h = box[:4]
    aspect_ratio = image_size[0] / image_size[1]
    center = np.array([x + w * 0.5, y + h * 0.5], dtype=np.float32)
    if w > aspect_ratio * h:
        h = w * 1.0 / aspect_ratio
    elif w < aspect_ratio * h:
        w = h * aspect_ratio
    scale = np.array([w / 200.0, h / 200.0], dtype=np.float32)
    scale = scale * 1.25
    return center, scale


def get_warp_matrix(theta, size_input, size_dst, size_target):
    """UDP transform matrix (float32 [2,3])."""
    theta = np.deg2rad(theta)
    matrix = np.zeros((2, 3), dtype=np.float32)
    scale_x = size_dst[0] / size_target[0]
    scale_y = size_dst[1] / size_target[1]
    matrix[0, 0] = math.cos(theta) * scale_x
    matrix[0, 1] = -math.sin(theta) * scale_x
    matrix[0, 2] = scale_x * (-0.5 * size_input[0] * math.cos(theta) + 0.5 * size_input[1] * math.sin(theta) +
                              0.5 * size_target[0])
    matrix[1, 0] = math.sin(theta) * scale_y
    matrix[1, 1] = math.cos(theta) * scale_y
    matrix[1, 2] = scale_y * (-0.5 * size_input[0] * math.sin(theta) - 0.5 * size_input[1] * math.cos(theta) +
                              0.5 * size_target[1])
    return matrix


def get_affine_transform(center, scale, rot, output_size, shift=(0., 0.), inv=False):
    """Non-UDP transform matrix (float64 [2,3]) from center / scale / rotation."""
    assert len(center) == 2 and len(scale) == 2 and len(output_size) == 2 and len(shift) == 2
    scale_tmp = scale * 200.0
    shift = np.array(shift)
    src_w, dst_w, dst_h = scale_tmp[0], output_size[0], output_size[1]
    rot_rad = np.pi * rot / 180
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    src_dir = [-(src_w * -0.5) * sn, (src_w * -0.5) * cs]           # rotate_point([0, -src_w/2], rot_rad)
    dst_dir = np.array([0., dst_w * -0.5])
    src = np.zeros((3, 2), dtype=np.float32)
    src[0, :] = center + scale_tmp * shift
    src[1, :] = center + src_dir + scale_tmp * shift
    d = src[0, :] - src[1, :]
    src[2, :] = src[1, :] + np.array([-d[1], d[0]], dtype=np.float32)
    dst = np.zeros((3, 2), dtype=np.float32)
    dst[0, :] = [dst_w * 0.5, dst_h * 0.5]
    dst[1, :] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    d = dst[0, :] - dst[1, :]
    dst[2, :] = dst[1, :] + np.array([-d[1], d[0]], dtype=np.float32)
    a, b = (dst, src) if inv else (src, dst)
    A = np.concatenate([a.astype(np.float64), np.ones((3, 1))], axis=1)
    return np.linalg.solve(A, b.astype(np.float64)).T.copy()


def _invert_affine(M):
    """cv::invertAffineTransform in double, as warpAffine applies it before mapping dst -> src."""
    M = np.array(M, dtype=np.float64).reshape(2, 3).copy()
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    M[0, 0] = A11
    M[0, 1] *= -D
    M[1, 0] *= -D
    M[1, 1] = A22
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2], M[1, 2] = b1, b2
    return M


def box_transforms(boxes, image_size, use_udp=True, rotation=0., vectorize=True):
    """Host maths per box: (centers float32 list, scales float32 list, inverse maps float64 [n,6]).
    boxes: list of (image_index, (x, y, w, h[, score]))."""
    n = len(boxes)
    W, H = int(image_size[0]), int(image_size[1])
    size = np.array([W, H], dtype=np.float64)
    def _dtype_of(bx):
        return bx.dtype if isinstance(bx, np.ndarray) else np.asarray(bx[:4]).dtype
    uniform = vectorize and n > 0 and len({_dtype_of(bx) for _, bx in boxes}) == 1     # mixed dtypes: per-box scalar maths
    if use_udp and rotation == 0 and uniform:
        # vectorised over boxes, same dtype sequence as the scalar code (float64 box maths -> float32 center/scale ->
        # float32 scale*200 -> float64 quotients -> float32 matrix -> float64 inverse)
        b = np.array([list(bx[:4]) for _, bx in boxes])          # keeps the boxes' own dtype, like the scalar code
        if b.dtype.kind != 'f':
            b = b.astype(np.float64)
        x, y, w, h = b[:, 0], b[:, 1], b[:, 2].copy(), b[:, 3].copy()
        ar = W / H
        centers = np.stack([x + w * 0.5, y + h * 0.5], axis=1).astype(np.float32)
        wide, tall = w > ar * h, w < ar * h
        h = np.where(wide, w * 1.0 / ar, h)
        w = np.where(tall, h * ar, w)
        scales = np.stack([w / 200.0, h / 200.0], axis=1).astype(np.float32) * np.float32(1.25)
        target = scales * np.float32(200.0)                      # float32
        sxy = (size - 1.0)[None, :] / target                     # float64
        m = np.zeros((n, 2, 3), dtype=np.float32)
        m[:, 0, 0] = sxy[:, 0]
        m[:, 1, 1] = sxy[:, 1]
        m[:, 0, 2] = sxy[:, 0] * (-0.5 * (centers[:, 0] * 2.0) * 1.0 + 0.5 * (centers[:, 1] * 2.0) * 0.0 + 0.5 * target[:, 0])
        m[:, 1, 2] = sxy[:, 1] * (-0.5 * (centers[:, 0] * 2.0) * 0.0 - 0.5 * (centers[:, 1] * 2.0) * 1.0 + 0.5 * target[:, 1])
        m64 = m.astype(np.float64)
        D = m64[:, 0, 0] * m64[:, 1, 1] - m64[:, 0, 1] * m64[:, 1, 0]
        D = np.where(D != 0, 1.0 / np.where(D != 0, D, 1.0), 0.0)
        inv = np.zeros((n, 6), dtype=np.float64)
        inv[:, 0] = m64[:, 1, 1] * D
        inv[:, 1] = m64[:, 0, 1] * -D
        inv[:, 3] = m64[:, 1, 0] * -D
        inv[:, 4] = m64[:, 0, 0] * D
        inv[:, 2] = -inv[:, 0] * m64[:, 0, 2] - inv[:, 1] * m64[:, 1, 2]
        inv[:, 5] = -inv[:, 3] * m64[:, 0, 2] - inv[:, 4] * m64[:, 1, 2]
        center_list, scale_list = list(centers), list(scales)
    else:
        inv = np.zeros((n, 6), dtype=np.float64)
        center_list, scale_list = [], []
        for i, (idx, box) in enumerate(boxes):
            center, scale = box2cs(box, (W, H))
            if use_udp:
                trans = get_warp_matrix(rotation, center * 2.0, size - 1.0, scale * 200.0)
            else:
                trans = get_affine_transform(center, scale, rotation, size)
            inv[i] = _invert_affine(trans).reshape(-1)
            center_list.append(center)
            scale_list.append(scale)
    return center_list, scale_list, inv


def preprocess_crops(images, boxes, image_size=(192, 256), use_udp=True, rotation=0.,
                     mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225), flip_pairs=None):
    """images: list of uint8 CUDA tensors [h,w,3] (RGB, as mmcv.imread(channel_order='rgb'));
    boxes: list of (image_index, (x, y, w, h[, score])).
    Returns (crops float32 CUDA [n,3,H,W], img_metas list) ready for ``TopDown.forward_test``."""
    _lib.require_cuda()
    n = len(boxes)
    W, H = int(image_size[0]), int(image_size[1])
    dev = images[0].device
    for im in images:
        if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3 or not im.is_cuda or not im.is_contiguous():
            raise _lib.VitposeLibError('images must be contiguous uint8 CUDA tensors of shape [h, w, 3]')
    hw = np.array([images[idx].shape[:2] for idx, _ in boxes], dtype=np.int32).reshape(n, 2)
    ptrs = np.array([images[idx].data_ptr() for idx, _ in boxes], dtype=np.int64)
    center_list, scale_list, inv = box_transforms(boxes, (W, H), use_udp, rotation)
    metas = []
    for i, (idx, box) in enumerate(boxes):
        meta = dict(center=center_list[i], scale=scale_list[i], rotation=rotation, image_file='', bbox_id=i,
                    bbox_score=float(box[4]) if len(box) > 4 else 1.0)
        if flip_pairs is not None:
            meta['flip_pairs'] = flip_pairs
        metas.append(meta)
    out = torch.empty(n, 3, H, W, device=dev, dtype=torch.float32)
    if n == 0:
        return out, metas
    d_inv = torch.from_numpy(inv).to(dev)
    d_hw = torch.from_numpy(hw).to(dev)
    d_ptrs = torch.from_numpy(ptrs).to(dev)
    m3 = (ctypes.c_float * 3)(*[float(v) for v in mean])
    s3 = (ctypes.c_float * 3)(*[float(v) for v in std])
    check(lib().vpb_warp_affine_normalize(ptr(d_ptrs), ptr(d_hw), ptr(d_inv), n, H, W, m3, s3, ptr(out),
                                          stream_ptr()), 'vpb_warp_affine_normalize')
    return out, metas
