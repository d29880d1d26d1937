"""Torch-tensor front ends of the C-ABI operators (plumbing only: pointers, shapes, current stream).
Every function launches hand-written sm_100a kernels from libvitpose_b200.so; nothing here computes."""
import numpy as np
import torch

from . import _lib
from ._lib import check, lib, ptr, stream_ptr

BF16 = torch.bfloat16


def _need(t, dtype, name):
    if t.dtype != dtype:
        raise TypeError(f'{name} must be {dtype}, got {t.dtype}')
    return t


def gemm(a, b, epilogue, bias=None, out=None, aux=None, period=0, max_ctas=0):
    """a bf16 [M,K], b bf16 [N,K] -> out per `epilogue` (see include/vitpose_b200.h)."""
    _need(a, BF16, 'a'); _need(b, BF16, 'b')
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K
    if epilogue in (_lib.EPI_BIAS_BF16, _lib.EPI_GELU_BF16):
        out = torch.empty(M, N, device=a.device, dtype=BF16) if out is None else out
        ldo = out.stride(0)
    elif epilogue in (_lib.EPI_RESID_F32, _lib.EPI_POS_F32):
        out = torch.empty(M, N, device=a.device, dtype=torch.float32) if out is None else out
        ldo = out.stride(0)
    elif epilogue == _lib.EPI_NCHW_F32:
        assert M % period == 0
        out = torch.empty(M // period, N, period, device=a.device, dtype=torch.float32) if out is None else out
        ldo = 0
    else:
        raise ValueError(epilogue)
    check(lib().vpb_gemm_bf16(ptr(a), ptr(b), M, N, K, epilogue, ptr(bias), ptr(out), ldo, ptr(aux), period,
                              max_ctas, stream_ptr()), 'vpb_gemm_bf16')
    return out


def gemm_layernorm(a, b, epilogue, bias, aux, gamma, beta, eps=1e-6, period=0, out=None):
    """Residual-stream GEMM + the LayerNorm that follows it, one kernel: returns (out fp32 [M,N], xn bf16 [M,N]).
    `out` may be the residual tensor itself (in-place update, as the forward pass does)."""
    _need(a, BF16, 'a'); _need(b, BF16, 'b'); _need(aux, torch.float32, 'aux')
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K and epilogue in (_lib.EPI_RESID_F32, _lib.EPI_POS_F32)
    out = torch.empty(M, N, device=a.device, dtype=torch.float32) if out is None else out
    xn = torch.empty(M, N, device=a.device, dtype=BF16)
    nbytes = lib().vpb_gemm_layernorm_scratch_bytes(M, N)
    scratch = torch.empty(nbytes, device=a.device, dtype=torch.uint8)
    check(lib().vpb_gemm_bf16_layernorm(ptr(a), ptr(b), M, N, K, epilogue, ptr(bias), ptr(out), ptr(aux), period,
                                        ptr(gamma), ptr(beta), float(eps), ptr(xn), ptr(scratch), nbytes,
                                        stream_ptr()), 'vpb_gemm_bf16_layernorm')
    return out, xn


def layernorm(x, gamma, beta, eps=1e-6, out=None):
    _need(x, torch.float32, 'x')
    M, D = x.shape
    out = torch.empty(M, D, device=x.device, dtype=BF16) if out is None else out
    check(lib().vpb_layernorm_bf16(ptr(x), ptr(gamma), ptr(beta), ptr(out), M, D, float(eps), stream_ptr()),
          'vpb_layernorm_bf16')
    return out


def im2col_patch16(img, flip=False):
    _need(img, torch.float32, 'img')
    n, c, H, W = img.shape
    assert c == 3
    Hp, Wp = (H + 4 - 16) // 16 + 1, (W + 4 - 16) // 16 + 1
    rows = (2 * n if flip else n) * Hp * Wp
    out = torch.empty(rows, 768, device=img.device, dtype=BF16)
    check(lib().vpb_im2col_patch16(ptr(img), ptr(out), n, H, W, int(flip), stream_ptr()), 'vpb_im2col_patch16')
    return out


def attention(qkv, heads, scale=None):
    """qkv bf16 [n, T, 3*heads*hd] -> bf16 [n, T, heads*hd]."""
    _need(qkv, BF16, 'qkv')
    n, T, C3 = qkv.shape
    hd = C3 // (3 * heads)
    scale = hd ** -0.5 if scale is None else scale
    out = torch.empty(n, T, heads * hd, device=qkv.device, dtype=BF16)
    check(lib().vpb_attention(ptr(qkv), ptr(out), n, T, heads, hd, float(scale), stream_ptr()), 'vpb_attention')
    return out


def deconv4x4s2_bn_relu(x_nhwc, wphase, scale, shift):
    _need(x_nhwc, BF16, 'x'); _need(wphase, BF16, 'wphase')
    n, h, w, cin = x_nhwc.shape
    cout = wphase.shape[1]
    out = torch.empty(n, 2 * h, 2 * w, cout, device=x_nhwc.device, dtype=BF16)
    check(lib().vpb_deconv4x4s2_bn_relu(ptr(x_nhwc), ptr(wphase), ptr(scale), ptr(shift), ptr(out), n, h, w, cin,
                                        cout, stream_ptr()), 'vpb_deconv4x4s2_bn_relu')
    return out


def conv3x3_nchw(x_nhwc, w9, bias):
    _need(x_nhwc, BF16, 'x'); _need(w9, BF16, 'w9')
    n, h, w, cin = x_nhwc.shape
    cout = w9.shape[0]
    out = torch.empty(n, cout, h, w, device=x_nhwc.device, dtype=torch.float32)
    check(lib().vpb_conv3x3_nchw(ptr(x_nhwc), ptr(w9), ptr(bias), ptr(out), n, h, w, cin, cout, stream_ptr()),
          'vpb_conv3x3_nchw')
    return out


def relu_upsample_nhwc(x_nhwc, factor):
    _need(x_nhwc, BF16, 'x')
    n, h, w, C = x_nhwc.shape
    out = torch.empty(n, h * factor, w * factor, C, device=x_nhwc.device, dtype=BF16)
    check(lib().vpb_relu_upsample_nhwc(ptr(x_nhwc), ptr(out), n, h, w, C, factor, stream_ptr()),
          'vpb_relu_upsample_nhwc')
    return out


def tokens_to_nchw(tokens, hp, wp):
    _need(tokens, BF16, 'tokens')
    n, T, D = tokens.shape
    out = torch.empty(n, D, hp, wp, device=tokens.device, dtype=torch.float32)
    check(lib().vpb_tokens_to_nchw_f32(ptr(tokens), ptr(out), n, T, D, stream_ptr()), 'vpb_tokens_to_nchw_f32')
    return out


def decode(hm, hm_flipped=None, flip_index=None, shift_heatmap=False, mode=_lib.DECODE_DEFAULT, kernel=11,
           use_udp=False, center=None, scale=None, want_merged=False, want_argmax=False):
    """Device-side decode.  hm fp32 [N,K,H,W] CUDA.  Returns dict of CUDA tensors."""
    _need(hm, torch.float32, 'hm')
    N, K, H, W = hm.shape
    dev = hm.device
    preds = torch.empty(N, K, 2, device=dev, dtype=torch.float32)
    maxvals = torch.empty(N, K, 1, device=dev, dtype=torch.float32)
    merged = torch.empty_like(hm) if want_merged else None
    amax = torch.empty(N, K, device=dev, dtype=torch.int32) if want_argmax else None
    apply_tf = center is not None
    check(lib().vpb_decode_heatmaps(ptr(hm), ptr(hm_flipped), ptr(flip_index), int(bool(shift_heatmap)), N, K, H, W,
                                    int(mode), int(kernel), int(bool(use_udp)), int(apply_tf), ptr(center),
                                    ptr(scale), ptr(preds), ptr(maxvals), ptr(merged), ptr(amax), stream_ptr()),
          'vpb_decode_heatmaps')
    return dict(preds=preds, maxvals=maxvals, merged=merged, argmax=amax)


def flip_back(hm_flipped, flip_index=None, shift_heatmap=False):
    """hm_flipped fp32 CUDA [N,K,H,W] -> flipped-back (and optionally shifted) heatmaps."""
    _need(hm_flipped, torch.float32, 'hm_flipped')
    N, K, H, W = hm_flipped.shape
    out = torch.empty_like(hm_flipped)
    check(lib().vpb_flip_back(ptr(hm_flipped), ptr(flip_index), ptr(out), N, K, H, W, int(bool(shift_heatmap)),
                              stream_ptr()), 'vpb_flip_back')
    return out


def transform_preds(coords, center, scale, heatmap_wh, use_udp=False):
    """coords fp32 CUDA [N,K,2], center/scale fp32 CUDA [N,2] -> image coordinates."""
    _need(coords, torch.float32, 'coords')
    N, K, _ = coords.shape
    out = torch.empty_like(coords)
    check(lib().vpb_transform_preds(ptr(coords), ptr(center), ptr(scale), ptr(out), N, K, int(heatmap_wh[0]),
                                    int(heatmap_wh[1]), int(bool(use_udp)), stream_ptr()), 'vpb_transform_preds')
    return out
