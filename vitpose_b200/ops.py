"""Torch-tensor front ends of the C-ABI operators (plumbing only: pointers, shapes, current stream).
Every function launches hand-written sm_100a kernels from libvitpose_b200.so; nothing here computes."""
import ctypes

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
    elif epilogue == _lib.EPI_ACCUM_F32:
        assert out is not None and out.dtype == torch.float32 and bias is None, 'accumulates into an existing fp32 out'
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


def gemm_atb_accum(at, bt, out):
    """out [M,N] fp32 += at^T @ bt for row-major bf16 at [K,M], bt [K,N] (weight gradient from dY and X). at / bt may
    be column slices of wider row-major matrices (unit column stride, any row pitch that is a multiple of 8)."""
    _need(at, BF16, 'at'); _need(bt, BF16, 'bt'); _need(out, torch.float32, 'out')
    K, M = at.shape
    N = bt.shape[1]
    assert bt.shape[0] == K and out.shape == (M, N)
    if at.is_contiguous() and bt.is_contiguous():
        check(lib().vpb_gemm_bf16_atb_accum(ptr(at), ptr(bt), M, N, K, ptr(out), out.stride(0), stream_ptr()),
              'vpb_gemm_bf16_atb_accum')
        return out
    assert at.stride(1) == 1 and bt.stride(1) == 1, 'operands need unit column stride'
    if not (at.is_cuda and bt.is_cuda):
        raise _lib.VitposeLibError('the vitpose_b200 kernels need CUDA tensors (no CPU fallback)')
    check(lib().vpb_gemm_bf16_atb_accum_ld(at.data_ptr(), at.stride(0), bt.data_ptr(), bt.stride(0), M, N, K, ptr(out),
                                           out.stride(0), stream_ptr()), 'vpb_gemm_bf16_atb_accum_ld')
    return out


class LnScratch:
    """Statistics-exchange buffer of the fused-LayerNorm GEMMs for one (M, N), initialised once; successive launches on
    it (same stream) take increasing epochs instead of two memsets each (see include/vitpose_b200.h)."""

    def __init__(self, M, N, device):
        self.M, self.N = M, N
        self.nbytes = lib().vpb_gemm_layernorm_scratch_bytes(M, N)
        self.buf = torch.empty(self.nbytes, device=device, dtype=torch.uint8)
        self.stream = None
        self.epoch = 0

    def next_epoch(self):
        st = stream_ptr()
        if self.epoch == 0 or self.epoch >= (1 << 30) or st != self.stream:
            if self.stream is not None and st != self.stream:
                torch.cuda.synchronize(self.buf.device)          # (the launches on the other stream may still read it)
            check(lib().vpb_gemm_layernorm_scratch_init(ptr(self.buf), self.nbytes, self.M, self.N, st),
                  'vpb_gemm_layernorm_scratch_init')
            self.stream, self.epoch = st, 0
        self.epoch += 1
        return self.epoch


def gemm_layernorm(a, b, epilogue, bias, aux, gamma, beta, eps=1e-6, period=0, out=None, row_scale=None,
                   rows_per_scale=0, scratch=None, xn=None):
    """Residual-stream GEMM + the LayerNorm that follows it, one kernel: returns (out fp32 [M,N], xn bf16 [M,N]).
    `out` may be the residual tensor itself (in-place update, as the forward pass does)."""
    _need(a, BF16, 'a'); _need(b, BF16, 'b'); _need(aux, torch.float32, 'aux')
    M, K = a.shape
    N = b.shape[0]
    assert b.shape[1] == K and epilogue in (_lib.EPI_RESID_F32, _lib.EPI_POS_F32)
    out = torch.empty(M, N, device=a.device, dtype=torch.float32) if out is None else out
    xn = torch.empty(M, N, device=a.device, dtype=BF16) if xn is None else xn
    assert xn.shape == (M, N) and xn.is_contiguous() and out.is_contiguous()
    if scratch is not None:        # an LnScratch the caller keeps across launches: no memsets
        assert (scratch.M, scratch.N) == (M, N)
        check(lib().vpb_gemm_bf16_layernorm_seq(ptr(a), ptr(b), M, N, K, epilogue, ptr(bias), ptr(out), ptr(aux), period,
                                                ptr(gamma), ptr(beta), float(eps), ptr(xn), ptr(scratch.buf),
                                                scratch.nbytes, scratch.next_epoch(), ptr(row_scale),
                                                int(rows_per_scale), stream_ptr()), 'vpb_gemm_bf16_layernorm_seq')
        return out, xn
    nbytes = lib().vpb_gemm_layernorm_scratch_bytes(M, N)
    scratch = torch.empty(nbytes, device=a.device, dtype=torch.uint8)
    check(lib().vpb_gemm_bf16_layernorm(ptr(a), ptr(b), M, N, K, epilogue, ptr(bias), ptr(out), ptr(aux), period,
                                        ptr(gamma), ptr(beta), float(eps), ptr(xn), ptr(scratch), nbytes,
                                        ptr(row_scale), int(rows_per_scale), stream_ptr()), 'vpb_gemm_bf16_layernorm')
    return out, xn


def fold_layernorm_linear(w, bias, gamma, beta):
    """(Wf bf16 [N, K], s [N], c [N]) of a Linear layer that applies the LayerNorm in front of it in its epilogue."""
    N, K = w.shape
    w = w.detach().float().contiguous()
    wf = torch.empty(N, K, device=w.device, dtype=torch.bfloat16)
    s = torch.empty(N, device=w.device, dtype=torch.float32)
    c = torch.empty(N, device=w.device, dtype=torch.float32)
    check(lib().vpb_fold_layernorm_linear(ptr(w), ptr(bias), ptr(gamma), ptr(beta), N, K, ptr(wf), ptr(s), ptr(c),
                                          stream_ptr()), 'vpb_fold_layernorm_linear')
    return wf, s, c


def gemm_stats_layout(N):
    """(column tiles per row, columns per tile) of the statistics gemm_resid_stats writes; (0, 0) = unsupported N."""
    parts, cols = ctypes.c_int(0), ctypes.c_int(0)
    lib().vpb_gemm_stats_layout(int(N), ctypes.byref(parts), ctypes.byref(cols))
    return parts.value, cols.value


def gemm_resid_stats(a, b, epilogue, bias, aux, period=0, out=None, row_scale=None, rows_per_scale=0, xb=None,
                     stats=None):
    """out fp32 = aux + a @ b.T + bias (in place when out is aux), xb = bf16(out), stats = per-tile (mean, M2) pairs:
    the producer side of a folded LayerNorm. Returns (out, xb, stats [rows padded to 128, parts, 2])."""
    M, K = a.shape
    N = b.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    if xb is None:
        xb = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    nbytes = lib().vpb_gemm_stats_bytes(M, N)
    parts, _ = gemm_stats_layout(N)
    if stats is None:
        stats = torch.empty(max(nbytes, 8) // 8, 2, device=a.device, dtype=torch.float32)
    check(lib().vpb_gemm_bf16_resid_stats(ptr(a), ptr(b), M, N, K, epilogue, ptr(bias), ptr(out), ptr(aux), period,
                                          ptr(xb), ptr(stats), nbytes, ptr(row_scale), int(rows_per_scale),
                                          stream_ptr()), 'vpb_gemm_bf16_resid_stats')
    return out, xb, stats.view(-1, parts, 2)


def gemm_lnfold(xb, wf, s, c, stats, epilogue=None, eps=1e-6, out=None):
    """bf16 act(LayerNorm(x) @ W.T + b) from the plain bf16 rows, the folded weight and the row statistics."""
    M, K = xb.shape
    N = wf.shape[0]
    parts, cols = stats.shape[1], K // stats.shape[1]
    if out is None:
        out = torch.empty(M, N, device=xb.device, dtype=torch.bfloat16)
    check(lib().vpb_gemm_bf16_lnfold(ptr(xb), ptr(wf), M, N, K, EPI_BIAS_BF16 if epilogue is None else epilogue, ptr(c),
                                     ptr(s), ptr(stats), parts, cols, float(eps), ptr(out), out.stride(0),
                                     stream_ptr()), 'vpb_gemm_bf16_lnfold')
    return out


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


# ---- backward-pass operators of the training step (include/vitpose_b200.h, "backward pass") ----
def transpose(x, batch=1):
    """bf16 [batch, R, C] (or [R, C]) -> [batch, C, R]."""
    _need(x, BF16, 'x')
    R, C = x.shape[-2:]
    out = torch.empty(*x.shape[:-2], C, R, device=x.device, dtype=BF16)
    check(lib().vpb_transpose_bf16(ptr(x), ptr(out), R, C, batch, stream_ptr()), 'vpb_transpose_bf16')
    return out


def cast_bf16(x, row_scale=None, rows_per_scale=0):
    """fp32 -> bf16; with row_scale (fp32 [rows / rows_per_scale]) rows of the 2-D x are scaled first."""
    _need(x, torch.float32, 'x')
    out = torch.empty(x.shape, device=x.device, dtype=BF16)
    row_len = x.shape[-1] if row_scale is not None else 0
    check(lib().vpb_cast_f32_bf16(ptr(x), ptr(out), x.numel(), ptr(row_scale), int(row_len), int(rows_per_scale),
                                  stream_ptr()), 'vpb_cast_f32_bf16')
    return out


def relu(x):
    _need(x, BF16, 'x')
    out = torch.empty_like(x)
    check(lib().vpb_relu_bf16(ptr(x), ptr(out), x.numel(), stream_ptr()), 'vpb_relu_bf16')
    return out


def relu_bwd(y, dy):
    """dy where y > 0 else 0 (y = relu output)."""
    _need(y, BF16, 'y'); _need(dy, BF16, 'dy')
    out = torch.empty_like(dy)
    check(lib().vpb_relu_bwd_bf16(ptr(y), ptr(dy), ptr(out), y.numel(), stream_ptr()), 'vpb_relu_bwd_bf16')
    return out


def simple_head_gather(z, bias, K, h, w, factor):
    """z fp32 [images, 9K, h*w] (tap maps) -> heatmaps fp32 [images, K, h*factor, w*factor]."""
    _need(z, torch.float32, 'z')
    images = z.shape[0]
    out = torch.empty(images, K, h * factor, w * factor, device=z.device, dtype=torch.float32)
    check(lib().vpb_simple_head_gather(ptr(z), ptr(bias), ptr(out), images, K, h, w, factor, stream_ptr()),
          'vpb_simple_head_gather')
    return out


def simple_head_gather_bwd(dout, h, w, factor, ldz):
    """dheatmaps fp32 [images, K, H, W] -> dz bf16 [images*h*w, ldz] (token-major, column k*9+t; pad columns zero)."""
    _need(dout, torch.float32, 'dout')
    images, K = dout.shape[:2]
    dz = torch.zeros(images * h * w, ldz, device=dout.device, dtype=BF16)
    check(lib().vpb_simple_head_gather_bwd(ptr(dout), ptr(dz), ldz, images, K, h, w, factor, stream_ptr()),
          'vpb_simple_head_gather_bwd')
    return dz


def cast_bf16_colsum(x, colsum, row_scale=None, rows_per_scale=0):
    """cast_bf16 of the 2-D x and colsum[C] (fp32) += column sums of the rounded result, in one pass."""
    _need(x, torch.float32, 'x'); _need(colsum, torch.float32, 'colsum')
    R, C = x.shape
    assert colsum.numel() == C
    out = torch.empty(x.shape, device=x.device, dtype=BF16)
    check(lib().vpb_cast_f32_bf16_colsum(ptr(x), ptr(out), R, C, ptr(row_scale), int(rows_per_scale), ptr(colsum),
                                         stream_ptr()), 'vpb_cast_f32_bf16_colsum')
    return out


def gemm_gelu_save(a, b, bias):
    """(gelu(a @ b.T + bias), a @ b.T + bias), both bf16: the MLP's first half in the training forward pass."""
    _need(a, BF16, 'a'); _need(b, BF16, 'b')
    M, K = a.shape
    N = b.shape[0]
    out = torch.empty(M, N, device=a.device, dtype=BF16)
    pre = torch.empty(M, N, device=a.device, dtype=BF16)
    check(lib().vpb_gemm_bf16_gelu_save(ptr(a), ptr(b), M, N, K, ptr(bias), ptr(out), N, ptr(pre), stream_ptr()),
          'vpb_gemm_bf16_gelu_save')
    return out, pre


def gemm_gelu_bwd(dy, wt, pre, colsum=None, out=None):
    """(dy @ wt.T) * gelu'(pre) as bf16 [M, N]; colsum (fp32 [N], optional) += its column sums."""
    _need(dy, BF16, 'dy'); _need(wt, BF16, 'wt'); _need(pre, BF16, 'pre')
    M, K = dy.shape
    N = wt.shape[0]
    assert pre.shape == (M, N) and pre.is_contiguous() and dy.is_contiguous()
    out = torch.empty(M, N, device=dy.device, dtype=BF16) if out is None else out
    assert out.shape == (M, N) and out.is_contiguous()
    check(lib().vpb_gemm_bf16_gelu_bwd(ptr(dy), ptr(wt), M, N, K, ptr(pre), ptr(out), N, ptr(colsum), stream_ptr()),
          'vpb_gemm_bf16_gelu_bwd')
    return out


def colsum_accumulate(x, out):
    """out[C] (fp32) += column sums of x [R, C] (bf16 or fp32)."""
    assert x.is_contiguous() and out.dtype == torch.float32
    R, C = x.shape
    check(lib().vpb_colsum_accumulate(ptr(x), int(x.dtype == torch.float32), R, C, ptr(out), stream_ptr()),
          'vpb_colsum_accumulate')
    return out


def gelu_fwd(pre):
    _need(pre, BF16, 'pre')
    out = torch.empty_like(pre)
    check(lib().vpb_gelu_fwd_bf16(ptr(pre), ptr(out), pre.numel(), stream_ptr()), 'vpb_gelu_fwd_bf16')
    return out


def gelu_bwd(pre, dh):
    _need(pre, BF16, 'pre'); _need(dh, BF16, 'dh')
    out = torch.empty_like(pre)
    check(lib().vpb_gelu_bwd_bf16(ptr(pre), ptr(dh), ptr(out), pre.numel(), stream_ptr()), 'vpb_gelu_bwd_bf16')
    return out


def layernorm_bwd(x, gamma, dy, dx_accum, dgamma, dbeta, eps=1e-6):
    _need(x, torch.float32, 'x'); _need(dy, BF16, 'dy'); _need(dx_accum, torch.float32, 'dx_accum')
    M, D = x.shape
    check(lib().vpb_layernorm_bwd(ptr(x), ptr(gamma), ptr(dy), ptr(dx_accum), ptr(dgamma), ptr(dbeta), M, D,
                                  float(eps), stream_ptr()), 'vpb_layernorm_bwd')


def attention_with_lse(qkv, heads, scale=None):
    """Forward attention that also returns the per-row log-sum-exp (base 2) the backward pass consumes."""
    _need(qkv, BF16, 'qkv')
    n, T, three = qkv.shape
    hd = three // 3 // heads
    scale = hd ** -0.5 if scale is None else scale
    out = torch.empty(n, T, heads * hd, device=qkv.device, dtype=BF16)
    lse = torch.empty(n, heads, T, device=qkv.device, dtype=torch.float32)
    check(lib().vpb_attention_lse(ptr(qkv), ptr(out), ptr(lse), n, T, heads, hd, float(scale), stream_ptr()),
          'vpb_attention_lse')
    return out, lse


def attention_bwd(qkv, out, lse, dout, heads, scale=None, dbias=None):
    """qkv bf16 [n,T,3*heads*hd], out / dout bf16 [n,T,heads*hd], lse fp32 [n,heads,T] -> dqkv bf16 [n,T,3*heads*hd];
    dbias (fp32 [3*heads*hd], optional) += column sums of dqkv over all tokens (attn.qkv's bias gradient)."""
    _need(qkv, BF16, 'qkv'); _need(out, BF16, 'out'); _need(dout, BF16, 'dout'); _need(lse, torch.float32, 'lse')
    n, T, three = qkv.shape
    hd = three // 3 // heads
    scale = hd ** -0.5 if scale is None else scale
    dqkv = torch.empty_like(qkv)
    if dbias is not None:
        _need(dbias, torch.float32, 'dbias')
        assert dbias.numel() == three
        check(lib().vpb_attention_bwd_bias(ptr(qkv), ptr(out), ptr(lse), ptr(dout), ptr(dqkv), ptr(dbias), n, T, heads,
                                           hd, float(scale), stream_ptr()), 'vpb_attention_bwd_bias')
    else:
        check(lib().vpb_attention_bwd(ptr(qkv), ptr(out), ptr(lse), ptr(dout), ptr(dqkv), n, T, heads, hd, float(scale),
                                      stream_ptr()), 'vpb_attention_bwd')
    return dqkv


def deconv4x4s2_raw(x, wphase):
    """ConvTranspose2d(k4,s2,p1) alone: x bf16 [n,h,w,cin] -> bf16 [n,2h,2w,cout]."""
    _need(x, BF16, 'x'); _need(wphase, BF16, 'wphase')
    n, h, w, cin = x.shape
    cout = wphase.shape[1]
    out = torch.empty(n, 2 * h, 2 * w, cout, device=x.device, dtype=BF16)
    ones = torch.ones(cout, device=x.device)
    zeros = torch.zeros(cout, device=x.device)
    check(lib().vpb_deconv4x4s2_raw(ptr(x), ptr(wphase), ptr(out), n, h, w, cin, cout, ptr(ones), ptr(zeros),
                                    stream_ptr()), 'vpb_deconv4x4s2_raw')
    return out


def bn_train_stats(raw, eps=1e-5, momentum=0.1, running_mean=None, running_var=None):
    """raw bf16 [..., C] -> (mean, rstd) fp32 [C] of the batch; updates the running statistics in place."""
    _need(raw, BF16, 'raw')
    C = raw.shape[-1]
    rows = raw.numel() // C
    scratch = torch.empty(4 * C, device=raw.device)
    mean, rstd = torch.empty(C, device=raw.device), torch.empty(C, device=raw.device)
    check(lib().vpb_bn_train_stats(ptr(raw), rows, C, float(eps), float(momentum), ptr(scratch), ptr(mean), ptr(rstd),
                                   ptr(running_mean), ptr(running_var), stream_ptr()), 'vpb_bn_train_stats')
    return mean, rstd


def bn_relu_fwd(raw, mean, rstd, gamma, beta):
    C = raw.shape[-1]
    act = torch.empty_like(raw)
    check(lib().vpb_bn_relu_fwd(ptr(raw), ptr(act), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta), raw.numel() // C, C,
                                stream_ptr()), 'vpb_bn_relu_fwd')
    return act


def bn_relu_bwd(raw, dact, mean, rstd, gamma, beta, dgamma, dbeta, eval_mode=False):
    """eval_mode: mean / rstd are running statistics (BatchNorm2d.eval() inside forward_train)."""
    C = raw.shape[-1]
    draw = torch.empty_like(raw)
    scratch = torch.empty(6 * C, device=raw.device)
    fn = lib().vpb_bn_relu_bwd_eval if eval_mode else lib().vpb_bn_relu_bwd
    check(fn(ptr(raw), ptr(dact), ptr(draw), ptr(mean), ptr(rstd), ptr(gamma), ptr(beta),
                                ptr(dgamma), ptr(dbeta), ptr(scratch), raw.numel() // C, C, stream_ptr()),
          'vpb_bn_relu_bwd')
    return draw


def nchw_to_rows(x, Kp):
    """fp32 [n, K, P] -> bf16 [n*P, Kp] (zero padded columns)."""
    _need(x, torch.float32, 'x')
    n, K, P = x.shape
    out = torch.empty(n * P, Kp, device=x.device, dtype=BF16)
    check(lib().vpb_nchw_f32_to_rows_bf16(ptr(x), ptr(out), n, K, P, Kp, stream_ptr()), 'vpb_nchw_f32_to_rows_bf16')
    return out


def deconv_gather_x(x):
    n, h, w, cin = x.shape
    out = torch.empty(4, n * h * w, 4 * cin, device=x.device, dtype=BF16)
    check(lib().vpb_deconv_gather_x(ptr(x), ptr(out), n, h, w, cin, stream_ptr()), 'vpb_deconv_gather_x')
    return out


def deconv_gather_dy(dy):
    n, h2, w2, cout = dy.shape
    h, w = h2 // 2, w2 // 2
    out = torch.empty(n * h * w, 16 * cout, device=dy.device, dtype=BF16)
    check(lib().vpb_deconv_gather_dy(ptr(dy), ptr(out), n, h, w, cout, stream_ptr()), 'vpb_deconv_gather_dy')
    return out


def deconv_phase_dy(dy):
    n, h2, w2, cout = dy.shape
    h, w = h2 // 2, w2 // 2
    out = torch.empty(4, n * h * w, cout, device=dy.device, dtype=BF16)
    check(lib().vpb_deconv_phase_dy(ptr(dy), ptr(out), n, h, w, cout, stream_ptr()), 'vpb_deconv_phase_dy')
    return out


def pose_pck_accuracy(output, target, weight, thr=0.05, normalize=None):
    """Device-side pose_pck_accuracy: output / target fp32 CUDA [N,K,H,W], weight fp32 CUDA [N,K] (> 0 = visible).
    Returns (acc [K], avg_acc [1], cnt [1] int32) as CUDA tensors (no host synchronisation)."""
    _need(output, torch.float32, 'output'); _need(target, torch.float32, 'target')
    N, K, H, W = output.shape
    n0, n1 = (float(H), float(W)) if normalize is None else (float(normalize[0]), float(normalize[1]))
    pred = decode(output, mode=_lib.DECODE_NONE)['preds']
    gt = decode(target, mode=_lib.DECODE_NONE)['preds']
    dev = output.device
    acc = torch.empty(K, device=dev, dtype=torch.float32)
    avg = torch.empty(1, device=dev, dtype=torch.float32)
    cnt = torch.empty(1, device=dev, dtype=torch.int32)
    w = weight.reshape(N, K).float().contiguous()
    check(lib().vpb_pose_pck_accuracy(ptr(pred), ptr(gt), ptr(w), N, K, n0, n1, float(thr), ptr(acc), ptr(avg),
                                      ptr(cnt), stream_ptr()), 'vpb_pose_pck_accuracy')
    return acc, avg, cnt


def deconv_pack_weight(w, want_dgrad=True):
    """ConvTranspose2d weight fp32 [Cin, Cout, 4, 4] -> (wp bf16 [4, Cout, 4*Cin], wd bf16 [Cin, 16*Cout] or None):
    the forward and input-gradient operands of the transposed convolution, one launch (engine.pack_deconv_weight /
    pack_deconv_weight_dgrad are the torch formulations, kept for the one-time inference repack and the tests)."""
    _need(w, torch.float32, 'w')
    cin, cout = w.shape[:2]
    assert tuple(w.shape[2:]) == (4, 4) and w.is_contiguous()
    wp = torch.empty(4, cout, 4 * cin, device=w.device, dtype=BF16)
    wd = torch.empty(cin, 16 * cout, device=w.device, dtype=BF16) if want_dgrad else None
    check(lib().vpb_deconv_pack_weight(ptr(w), ptr(wp), ptr(wd), cin, cout, stream_ptr()), 'vpb_deconv_pack_weight')
    return wp, wd


def deconv_unpack_wgrad(dwp, out):
    """packed fp32 weight gradient [4, Cout, 4*Cin] -> ``out`` [Cin, Cout, 4, 4] (contiguous fp32)."""
    _need(dwp, torch.float32, 'dwp'); _need(out, torch.float32, 'out')
    cin, cout = out.shape[:2]
    assert dwp.shape == (4, cout, 4 * cin) and out.is_contiguous() and dwp.is_contiguous()
    check(lib().vpb_deconv_unpack_wgrad(ptr(dwp), ptr(out), cin, cout, stream_ptr()), 'vpb_deconv_unpack_wgrad')
    return out
