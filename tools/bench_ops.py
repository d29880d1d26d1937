"""Per-kernel timing on the GPU box (CUDA events, warm-up, L2-cold by rotating buffers)."""
import argparse
import json
import math
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vitpose_b200 import _lib, ops  # noqa: E402

BF16 = torch.bfloat16


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--crops', type=int, default=128)
    ap.add_argument('--model', default='base')
    a = ap.parse_args()
    dev = torch.device('cuda:0')
    D, heads, depth = {'small': (384, 12, 12), 'base': (768, 12, 12), 'large': (1024, 16, 24),
                       'huge': (1280, 16, 32)}[a.model]
    M = a.crops * 2 * 192
    res = {}
    x = torch.randn(M, D, device=dev).to(BF16)
    for name, N, K, epi in (('qkv', 3 * D, D, _lib.EPI_BIAS_BF16), ('proj', D, D, _lib.EPI_RESID_F32),
                            ('fc1', 4 * D, D, _lib.EPI_GELU_BF16), ('fc1_nogelu', 4 * D, D, _lib.EPI_BIAS_BF16),
                            ('fc2', D, 4 * D, _lib.EPI_RESID_F32)):
        A = torch.randn(M, K, device=dev).to(BF16)
        B = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(BF16)
        bias = torch.randn(N, device=dev)
        if epi == _lib.EPI_RESID_F32:
            out = torch.randn(M, N, device=dev)
            fn = lambda: ops.gemm(A, B, epi, bias=bias, out=out, aux=out)
        else:
            out = torch.empty(M, N, device=dev, dtype=BF16)
            fn = lambda: ops.gemm(A, B, epi, bias=bias, out=out)
        ms = timeit(fn)
        res[name] = dict(ms=ms, tflops=2.0 * M * N * K / ms / 1e9)
        ref_ms = timeit(lambda: torch.matmul(A, B.t()))
        res[name]['cublas_ms'] = ref_ms
        res[name]['cublas_tflops'] = 2.0 * M * N * K / ref_ms / 1e9
        del A, B, out
    for name, K in (('proj_ln', D), ('fc2_ln', 4 * D)):      # residual GEMM + fused LayerNorm (one kernel)
        A = torch.randn(M, K, device=dev).to(BF16)
        B = (torch.randn(D, K, device=dev) / math.sqrt(K)).to(BF16)
        bias, gm, bt = torch.randn(D, device=dev), torch.ones(D, device=dev), torch.zeros(D, device=dev)
        out = torch.randn(M, D, device=dev)
        ms = timeit(lambda: ops.gemm_layernorm(A, B, _lib.EPI_RESID_F32, bias, out, gm, bt, out=out))
        res[name] = dict(ms=ms, tflops=2.0 * M * D * K / ms / 1e9)
        del A, B, out
    qkv = torch.randn(a.crops * 2, 192, 3 * D, device=dev).to(BF16)
    ms = timeit(lambda: ops.attention(qkv, heads))
    fl = 4.0 * a.crops * 2 * 192 * 192 * D
    res['attention'] = dict(ms=ms, tflops=fl / ms / 1e9)
    xf = torch.randn(M, D, device=dev)
    g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    ms = timeit(lambda: ops.layernorm(xf, g, b))
    res['layernorm'] = dict(ms=ms, gbs=M * D * 6 / ms / 1e6)
    img = torch.randn(a.crops, 3, 256, 192, device=dev)
    ms = timeit(lambda: ops.im2col_patch16(img, True))
    res['im2col'] = dict(ms=ms, gbs=(a.crops * 3 * 256 * 192 * 4 + M * 768 * 2) / ms / 1e6)
    from vitpose_b200.engine import pack_deconv_weight
    n2 = a.crops * 2
    f = torch.randn(n2, 16, 12, D, device=dev).to(BF16)
    w1 = pack_deconv_weight(torch.randn(D, 256, 4, 4, device=dev) * 0.02)
    sc, sh = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    ms = timeit(lambda: ops.deconv4x4s2_bn_relu(f, w1, sc, sh))
    res['deconv1'] = dict(ms=ms, tflops=2.0 * n2 * 192 * D * 256 * 16 / ms / 1e9)
    f2 = torch.randn(n2, 32, 24, 256, device=dev).to(BF16)
    w2 = pack_deconv_weight(torch.randn(256, 256, 4, 4, device=dev) * 0.02)
    ms = timeit(lambda: ops.deconv4x4s2_bn_relu(f2, w2, sc, sh))
    res['deconv2'] = dict(ms=ms, tflops=2.0 * n2 * 768 * 256 * 256 * 16 / ms / 1e9)
    f3 = torch.randn(n2 * 3072, 256, device=dev).to(BF16)
    wf = torch.randn(17, 256, device=dev).to(BF16)
    bf = torch.zeros(17, device=dev)
    ms = timeit(lambda: ops.gemm(f3, wf, _lib.EPI_NCHW_F32, bias=bf, period=3072))
    res['final1x1'] = dict(ms=ms, gbs=(n2 * 3072 * 256 * 2 + n2 * 17 * 3072 * 4) / ms / 1e6)
    hm = torch.rand(n2, 17, 64, 48, device=dev)
    fi = torch.arange(17, device=dev, dtype=torch.int32)
    c, s = torch.rand(a.crops, 2, device=dev), torch.rand(a.crops, 2, device=dev)
    for mode, nm in ((_lib.DECODE_UDP_DARK, 'decode_udp'), (_lib.DECODE_DEFAULT, 'decode_default'),
                     (_lib.DECODE_UNBIASED, 'decode_unbiased')):
        ms = timeit(lambda: ops.decode(hm[:a.crops], hm[a.crops:], fi, False, mode, 11, True, c, s))
        res[nm] = dict(ms=ms, gbs=n2 * 17 * 3072 * 4 / ms / 1e6)
    for k, v in res.items():
        print(k, json.dumps({kk: round(vv, 4) for kk, vv in v.items()}))
    per_layer = res['qkv']['ms'] + res['proj_ln']['ms'] + res['fc1']['ms'] + res['fc2_ln']['ms'] + res['attention']['ms']
    total = depth * per_layer + res['deconv1']['ms'] + res['deconv2']['ms'] + res['final1x1']['ms'] + res['im2col']['ms'] + res['decode_udp']['ms']
    print('est step ms', round(total, 3), 'crops/s', round(a.crops / total * 1e3, 1))


def bench_nms(images=5000, poses=20, K=17):
    """Rescoring + OKS NMS of a COCO-val-sized evaluation (5000 images) in one launch."""
    import time
    import numpy as np
    from vitpose_b200.core.post_processing import oks_nms_batched
    rng = np.random.RandomState(0)
    P = images * poses
    people = rng.rand(images, 5, K, 2).astype(np.float32) * 200
    who = rng.randint(5, size=(images, poses))
    kp = np.zeros((images, poses, K, 3), dtype=np.float32)
    kp[..., :2] = people[np.arange(images)[:, None], who] + rng.randn(images, poses, K, 2).astype(np.float32) * 3
    kp[..., 2] = rng.rand(images, poses, K)
    kp = kp.reshape(P, K, 3)
    areas, box = rng.rand(P) * 3e4 + 5e3, rng.rand(P).astype(np.float32)
    starts = np.arange(images + 1, dtype=np.int32) * poses
    oks_nms_batched(kp, areas, box, starts, 0.9, None, 0.2, rescore=True, rescore_vis_thr=0.2)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    keep, _ = oks_nms_batched(kp, areas, box, starts, 0.9, None, 0.2, rescore=True, rescore_vis_thr=0.2)
    t_gpu = time.perf_counter() - t0
    print('oks_nms', json.dumps(dict(images=images, poses=P, gpu_call_ms=round(t_gpu * 1e3, 2), kept=int(sum(len(k) for k in keep)))))


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'nms':
        bench_nms()
    else:
        main()


def bench_preprocess(crops=256):
    """HBM roofline of the fused warp+normalise kernel: bytes = fp32 crop written + source pixels touched."""
    from vitpose_b200 import pipelines as PL
    dev = torch.device('cuda:0')
    img = torch.randint(0, 256, (1080, 1920, 3), device=dev, dtype=torch.uint8)
    boxes = [(0, [100 + 5 * i, 80 + 2 * i, 300, 500]) for i in range(crops)]
    ms = timeit(lambda: PL.preprocess_crops([img], boxes))
    out_b = crops * 3 * 256 * 192 * 4
    src_b = crops * int(300 * 1.25 * 500 * 1.25 * 3)        # box area read at most once from HBM (then L2)
    print('preprocess', json.dumps(dict(ms=round(ms, 4), gbs=round((out_b + src_b) / ms / 1e6, 1), crops=crops,
                                        note='includes host-side matrix maths + 3 small H2D copies')))


if __name__ == '__main__' and os.environ.get('VPB_BENCH_PREPROCESS'):
    bench_preprocess()
