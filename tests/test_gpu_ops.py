"""Per-kernel parity (-m gpu): each sm_100a kernel, called through the C ABI, against a plain PyTorch
fp32 evaluation of the same op on the same bf16-rounded operands."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16 = torch.bfloat16


def _dev():
    return torch.device('cuda:0')


def _rand_bf16(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(BF16)


def _report(name, got, ref, tol, bf16_out=True):
    """|got - ref| <= tol + one bf16 ulp of |ref| (2^-8 relative) when the kernel's output is bf16."""
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    if bf16_out:
        err = (err - ref.abs() * 2.0 ** -8).clamp_min(0)
    m = err.max().item()
    if not (m <= tol):
        idx = np.unravel_index(int(err.argmax()), err.shape)
        bad = (err > tol).float().mean().item()
        rows_bad = (err > tol).any(dim=-1).nonzero().flatten()[:16].tolist() if err.ndim == 2 else None
        raise AssertionError(f'{name}: max err {m:.4g} > {tol} at {idx} (got {got[idx].item():.5g}, '
                             f'ref {ref[idx].item():.5g}); {bad * 100:.2f}% elements off; first bad rows {rows_bad}')
    return m


@pytest.mark.parametrize('M,N,K', [(128, 256, 64), (128, 256, 768), (256, 128, 128), (384, 768, 768),
                                   (1000, 384, 192), (192 * 8, 2304, 768), (192 * 6, 3072, 768),
                                   (192 * 6, 768, 3072), (192 * 3, 1152, 384), (192 * 2, 3840, 1280),
                                   (192 * 170, 768, 768)])
def test_gemm_bias_bf16(M, N, K):
    from vitpose_b200 import ops, _lib
    a, b = _rand_bf16((M, K), 1), _rand_bf16((N, K), 2, 1.0 / math.sqrt(K))
    bias = torch.randn(N, generator=torch.Generator().manual_seed(3))
    out = ops.gemm(a.to(_dev()), b.to(_dev()), _lib.EPI_BIAS_BF16, bias=bias.to(_dev()))
    torch.cuda.synchronize()
    ref = a.float() @ b.float().t() + bias
    _report(f'gemm {M}x{N}x{K}', out, ref, 0.03)


def test_gemm_gelu_bf16():
    from vitpose_b200 import ops, _lib
    M, N, K = 192 * 4, 1024, 256
    a, b = _rand_bf16((M, K), 4), _rand_bf16((N, K), 5, 2.0 / math.sqrt(K))
    bias = torch.randn(N, generator=torch.Generator().manual_seed(6)) * 0.5
    out = ops.gemm(a.to(_dev()), b.to(_dev()), _lib.EPI_GELU_BF16, bias=bias.to(_dev()))
    ref = F.gelu(a.float() @ b.float().t() + bias)
    _report('gemm+gelu', out, ref, 0.03)


def test_gemm_gelu_values_exact_erf():
    """The epilogue's GELU itself: an identity weight makes the pre-activations the (bf16) inputs, which sweep
    [-9, 9] densely plus large magnitudes; against fp64 x Phi(x) the error must stay at float rounding level
    before the bf16 rounding of the output."""
    from vitpose_b200 import ops, _lib
    K = 64
    x = torch.cat([torch.linspace(-9, 9, 128 * K - 256), torch.tensor([0.0, -0.0, 1e-6, -1e-6, 30.0, -30.0, 1e4, -1e4]),
                   torch.randn(248) * 3]).to(BF16).reshape(-1, K)
    eye = torch.eye(K).to(BF16)
    out = ops.gemm(x.to(_dev()), eye.to(_dev()), _lib.EPI_GELU_BF16, bias=torch.zeros(K, device=_dev()))
    xd = x.double()
    ref = 0.5 * xd * (1 + torch.erf(xd / math.sqrt(2)))
    got = out.cpu().double()
    # half a bf16 ulp of the exact value for the output rounding + 2e-6 for the formula (|error| <= 7e-7) and fp32
    ulp = 2.0 ** (torch.floor(torch.log2(ref.abs().clamp_min(1e-30))) - 7)
    tol = 0.5 * ulp + 2e-6
    assert ((got - ref).abs() <= tol).all(), ((got - ref).abs() / tol).max()
    assert (got[x.float() >= 30] == x.double()[x.float() >= 30]).all() and (got[x.float() <= -30] == 0).all()


def test_gemm_residual_f32_inplace():
    from vitpose_b200 import ops, _lib
    M, N, K = 192 * 5, 768, 768
    a, b = _rand_bf16((M, K), 7), _rand_bf16((N, K), 8, 1.0 / math.sqrt(K))
    bias = torch.randn(N, generator=torch.Generator().manual_seed(9))
    resid = torch.randn(M, N, generator=torch.Generator().manual_seed(10))
    x = resid.clone().to(_dev())
    ops.gemm(a.to(_dev()), b.to(_dev()), _lib.EPI_RESID_F32, bias=bias.to(_dev()), out=x, aux=x)
    ref = resid + a.float() @ b.float().t() + bias
    _report('gemm+residual', x, ref, 2e-3, bf16_out=False)


def test_gemm_pos_f32():
    from vitpose_b200 import ops, _lib
    T, n, D, K = 192, 3, 384, 768
    a, b = _rand_bf16((n * T, K), 11), _rand_bf16((D, K), 12, 1.0 / math.sqrt(K))
    bias = torch.randn(D, generator=torch.Generator().manual_seed(13))
    pos = torch.randn(T, D, generator=torch.Generator().manual_seed(14))
    out = ops.gemm(a.to(_dev()), b.to(_dev()), _lib.EPI_POS_F32, bias=bias.to(_dev()), aux=pos.to(_dev()), period=T)
    ref = (a.float() @ b.float().t() + bias).reshape(n, T, D) + pos
    _report('gemm+pos', out.reshape(n, T, D), ref, 2e-3, bf16_out=False)


@pytest.mark.parametrize('Kout,C', [(17, 256), (133, 256), (5, 64)])
def test_gemm_nchw_heatmap(Kout, C):
    from vitpose_b200 import ops, _lib
    n, P = 3, 3072
    a, b = _rand_bf16((n * P, C), 15), _rand_bf16((Kout, C), 16, 1.0 / math.sqrt(C))
    bias = torch.randn(Kout, generator=torch.Generator().manual_seed(17))
    out = ops.gemm(a.to(_dev()), b.to(_dev()), _lib.EPI_NCHW_F32, bias=bias.to(_dev()), period=P)
    ref = (a.float() @ b.float().t() + bias).reshape(n, P, Kout).permute(0, 2, 1)
    _report('gemm nchw', out, ref, 2e-3, bf16_out=False)


@pytest.mark.parametrize('M,D,K,offset', [
    (192 * 3, 128, 128, 0.0), (192 * 5, 384, 384, 0.5), (192 * 7, 768, 768, 0.0), (192 * 40, 768, 3072, 3.0),
    (192 * 200, 768, 768, 0.5), (192 * 200, 768, 3072, 0.0), (192 * 33, 1024, 1024, 0.0), (192 * 64, 1280, 5120, 1.0),
    (192, 768, 768, 0.0), (192 * 64, 384, 384, 0.5), (192 * 64, 384, 1536, 1.0)])
def test_gemm_residual_layernorm_fused(M, D, K, offset):
    """x = x + a @ w^T + bias (fp32, in place) and xn = LayerNorm(x) * gamma + beta from ONE kernel, against the
    two-step fp32 reference; row means up to 3 sigma to exercise the (mean, M2) merge across column tiles."""
    from vitpose_b200 import ops, _lib
    g = torch.Generator().manual_seed(M + D)
    a, w = _rand_bf16((M, K), 31), _rand_bf16((D, K), 32, 1.0 / math.sqrt(K))
    bias = torch.randn(D, generator=g)
    resid = torch.randn(M, D, generator=g) * 2 + offset * 2
    gamma, beta = torch.randn(D, generator=g), torch.randn(D, generator=g)
    x = resid.clone().to(_dev())
    out, xn = ops.gemm_layernorm(a.to(_dev()), w.to(_dev()), _lib.EPI_RESID_F32, bias.to(_dev()), x,
                                 gamma.to(_dev()), beta.to(_dev()), 1e-6, out=x)
    torch.cuda.synchronize()
    assert out.data_ptr() == x.data_ptr()
    if M * D * K <= 192 * 64 * 1280 * 5120:
        ref = resid.to(_dev()) + a.to(_dev()).float() @ w.to(_dev()).float().t() + bias.to(_dev())
        _report('gemm+residual (ln variant)', x, ref.cpu(), 2e-3, bf16_out=False)
    # LayerNorm of the kernel's own fp32 rows isolates the normalisation from GEMM rounding
    ref_n = F.layer_norm(x.cpu(), (D,), gamma, beta, 1e-6)
    _report('fused layernorm', xn, ref_n, 0.04)


def test_gemm_layernorm_shared_scratch_epochs():
    """Successive fused-LayerNorm launches on ONE scratch (ops.LnScratch: initialised once, increasing epochs — the
    training forward and vpb_vitpose_forward) give exactly what a freshly initialised scratch gives, through several
    region / tag cycles of the exchange protocol and with different K."""
    from vitpose_b200 import ops, _lib
    M, D = 192 * 9, 768
    g = torch.Generator().manual_seed(3)
    gamma, beta = torch.randn(D, generator=g).to(_dev()), torch.randn(D, generator=g).to(_dev())
    bias = torch.randn(D, generator=g).to(_dev())
    x0 = (torch.randn(M, D, generator=g) * 2).to(_dev())
    scratch = ops.LnScratch(M, D, _dev())
    xa, xb = x0.clone(), x0.clone()
    for i in range(9):
        K = 768 if i % 2 == 0 else 3072
        a = _rand_bf16((M, K), 100 + i).to(_dev())
        w = _rand_bf16((D, K), 200 + i, 1.0 / math.sqrt(K)).to(_dev())
        _, na = ops.gemm_layernorm(a, w, _lib.EPI_RESID_F32, bias, xa, gamma, beta, out=xa)
        _, nb = ops.gemm_layernorm(a, w, _lib.EPI_RESID_F32, bias, xb, gamma, beta, out=xb, scratch=scratch)
        torch.cuda.synchronize()
        assert torch.equal(xa, xb) and torch.equal(na, nb), f'launch {i} (epoch {scratch.epoch}) differs'
    assert scratch.epoch == 9


def test_gemm_pos_layernorm_fused():
    from vitpose_b200 import ops, _lib
    T, n, D, K = 192, 9, 768, 768
    g = torch.Generator().manual_seed(5)
    a, b = _rand_bf16((n * T, K), 11), _rand_bf16((D, K), 12, 1.0 / math.sqrt(K))
    bias, pos = torch.randn(D, generator=g), torch.randn(T, D, generator=g)
    gamma, beta = torch.randn(D, generator=g), torch.randn(D, generator=g)
    out, xn = ops.gemm_layernorm(a.to(_dev()), b.to(_dev()), _lib.EPI_POS_F32, bias.to(_dev()), pos.to(_dev()),
                                 gamma.to(_dev()), beta.to(_dev()), 1e-6, period=T)
    ref = (a.float() @ b.float().t() + bias).reshape(n, T, D) + pos
    _report('gemm+pos (ln variant)', out.reshape(n, T, D), ref, 2e-3, bf16_out=False)
    _report('fused layernorm', xn, F.layer_norm(out.cpu(), (D,), gamma, beta, 1e-6), 0.04)


def test_fold_layernorm_linear():
    """Wf = bf16(gamma o W), s = row sums of the ROUNDED Wf, c = b + W . beta (vpb_fold_layernorm_linear)."""
    from vitpose_b200 import ops
    g = torch.Generator().manual_seed(77)
    N, K = 2304, 768
    w, bias = torch.randn(N, K, generator=g) / math.sqrt(K), torch.randn(N, generator=g)
    gamma, beta = torch.randn(K, generator=g), torch.randn(K, generator=g)
    wf, s, c = ops.fold_layernorm_linear(w.to(_dev()), bias.to(_dev()), gamma.to(_dev()), beta.to(_dev()))
    ref_wf = (w * gamma).to(BF16)
    assert torch.equal(wf.cpu(), ref_wf)
    assert (s.cpu() - ref_wf.float().sum(1)).abs().max() < 1e-4
    assert (c.cpu() - (bias + w @ beta)).abs().max() < 1e-4


def _fold_check(M, D, K, N2, offset, epilogue, pos_period=0):
    """producer (fp32 rows + plain bf16 copy + per-tile (mean, M2)) and consumer (LayerNorm applied in the epilogue of
    the next Linear layer) of the folded LayerNorm, each against fp32 PyTorch on the same operands."""
    from vitpose_b200 import ops, _lib
    dev = _dev()
    g = torch.Generator().manual_seed(M + D + K)
    a, w = _rand_bf16((M, K), 31).to(dev), _rand_bf16((D, K), 32, 1.0 / math.sqrt(K)).to(dev)
    bias = torch.randn(D, generator=g).to(dev)
    if pos_period:
        aux = torch.randn(pos_period, D, generator=g).to(dev)
        out, xb, stats = ops.gemm_resid_stats(a, w, _lib.EPI_POS_F32, bias, aux, period=pos_period)
        ref = (a.float() @ w.float().t() + bias).reshape(-1, pos_period, D) + aux
        ref = ref.reshape(M, D)
    else:
        resid = (torch.randn(M, D, generator=g) * 2 + offset * 2).to(dev)
        x = resid.clone()
        out, xb, stats = ops.gemm_resid_stats(a, w, _lib.EPI_RESID_F32, bias, x, out=x)
        assert out.data_ptr() == x.data_ptr()
        ref = resid + a.float() @ w.float().t() + bias
    torch.cuda.synchronize()
    _report('fold producer rows', out, ref, 2e-3, bf16_out=False)
    assert torch.equal(xb, out.to(BF16)), 'the bf16 copy must be the rounded fp32 rows'
    parts, cols = ops.gemm_stats_layout(D)
    assert parts * cols == D and stats.shape[1] == parts
    tiles = out.reshape(M, parts, cols)
    mean = tiles.mean(-1)
    m2 = ((tiles - mean[..., None]) ** 2).sum(-1)
    assert (stats[:M, :, 0] - mean).abs().max().item() < 1e-4 * (1 + abs(offset) * 2)
    assert ((stats[:M, :, 1] - m2).abs() / m2.clamp_min(1.0)).max().item() < 1e-4
    # consumer
    w2 = (torch.randn(N2, D, generator=g) / math.sqrt(D)).to(dev)
    b2 = torch.randn(N2, generator=g).to(dev)
    gamma, beta = (1 + 0.3 * torch.randn(D, generator=g)).to(dev), (0.3 * torch.randn(D, generator=g)).to(dev)
    wf, s, c = ops.fold_layernorm_linear(w2, b2, gamma, beta)
    y = ops.gemm_lnfold(xb, wf, s, c, stats, epilogue=epilogue)
    torch.cuda.synchronize()
    # the same information the kernel has: statistics of the fp32 rows, the bf16-rounded rows and folded weight
    mu, var = out.mean(-1, keepdim=True), out.var(-1, unbiased=False, keepdim=True)
    rstd = torch.rsqrt(var + 1e-6)
    ref_y = rstd * (xb.float() @ wf.float().t() - mu * s) + c
    exact = F.layer_norm(out, (D,), gamma, beta, 1e-6) @ w2.t() + b2
    if epilogue == _lib.EPI_GELU_BF16:
        ref_y, exact = F.gelu(ref_y), F.gelu(exact)
    _report('fold consumer (same operands)', y, ref_y, 0.03)
    # against the unfolded fp32 evaluation: bf16 rounding of x (instead of LN(x)) scales with |mean| / sigma
    _report('fold consumer (exact LN)', y, exact, 0.06 * (1 + abs(offset)))


@pytest.mark.parametrize('M,D,K,N2,offset', [
    (192 * 7, 768, 768, 2304, 0.0), (192 * 40, 768, 3072, 3072, 1.0), (192 * 5, 384, 384, 1152, 0.5),
    (192 * 64, 384, 1536, 1536, 0.0), (192 * 33, 1024, 1024, 3072, 0.0), (192 * 16, 1280, 5120, 3840, 0.5),
    (192, 768, 768, 768, 0.0), (1000, 768, 768, 2304, 0.0)])
def test_gemm_folded_layernorm(M, D, K, N2, offset):
    from vitpose_b200 import _lib
    _fold_check(M, D, K, N2, offset, _lib.EPI_BIAS_BF16)


def test_gemm_folded_layernorm_gelu_and_pos():
    from vitpose_b200 import _lib
    _fold_check(192 * 9, 768, 768, 3072, 0.0, _lib.EPI_GELU_BF16)
    _fold_check(192 * 9, 768, 768, 2304, 0.0, _lib.EPI_BIAS_BF16, pos_period=192)
    _fold_check(192 * 6, 384, 768, 1152, 0.0, _lib.EPI_BIAS_BF16, pos_period=192)


@pytest.mark.parametrize('D', [128, 384, 768, 1024, 1280])
def test_layernorm(D):
    from vitpose_b200 import ops
    M = 192 * 3 + 5
    g = torch.Generator().manual_seed(D)
    x = torch.randn(M, D, generator=g) * 3 + 0.5
    gamma, beta = torch.randn(D, generator=g), torch.randn(D, generator=g)
    out = ops.layernorm(x.to(_dev()), gamma.to(_dev()), beta.to(_dev()), 1e-6)
    ref = F.layer_norm(x, (D,), gamma, beta, 1e-6)
    _report('layernorm', out, ref, 0.04)


@pytest.mark.parametrize('flip', [False, True])
def test_im2col(flip):
    from vitpose_b200 import ops
    n = 3
    img = torch.randn(n, 3, 256, 192, generator=torch.Generator().manual_seed(21))
    out = ops.im2col_patch16(img.to(_dev()), flip=flip).float().cpu()
    src = torch.cat([img, img.flip(3)]) if flip else img
    ref = F.unfold(src, kernel_size=16, stride=16, padding=2)          # [n, 768, 192]
    ref = ref.transpose(1, 2).reshape(-1, 768).to(BF16).float()
    assert out.shape == ref.shape
    assert torch.equal(out, ref), f'im2col mismatch: {(out - ref).abs().max()}'


@pytest.mark.parametrize('heads,hd', [(12, 64), (2, 64), (12, 32), (16, 80)])
def test_attention(heads, hd):
    from vitpose_b200 import ops
    n, T = 3, 192
    D = heads * hd
    qkv = _rand_bf16((n, T, 3 * D), 30 + hd, 1.5)
    out = ops.attention(qkv.to(_dev()), heads)
    torch.cuda.synchronize()
    q, k, v = qkv.float().reshape(n, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
    att = ((q * hd ** -0.5) @ k.transpose(-2, -1)).softmax(-1)
    ref = (att @ v).transpose(1, 2).reshape(n, T, D)
    _report(f'attention h{heads} d{hd}', out, ref, 0.03)


@pytest.mark.parametrize('heads,hd,n', [(16, 80, 40), (12, 64, 40)])
def test_attention_many_units_per_cta(heads, hd, n):
    """Several (crop, head, q-tile) units per persistent CTA: exercises the tile rings and the single O accumulator of
    the head_dim-80 variant (64-column SWIZZLE_128B box + 16-column SWIZZLE_32B box per operand)."""
    from vitpose_b200 import ops
    T, D = 192, heads * hd
    qkv = _rand_bf16((n, T, 3 * D), 77 + hd, 1.5).to(_dev())
    out = ops.attention(qkv, heads)
    q, k, v = qkv.float().reshape(n, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
    att = ((q * hd ** -0.5) @ k.transpose(-2, -1)).softmax(-1)
    ref = (att @ v).transpose(1, 2).reshape(n, T, D)
    _report(f'attention h{heads} d{hd} n{n}', out, ref.cpu(), 0.03)
    assert torch.equal(out, ops.attention(qkv, heads))          # deterministic


@pytest.mark.parametrize('n,h,w,cin,cout', [(2, 16, 12, 128, 64), (4, 16, 12, 768, 256), (3, 32, 24, 256, 256),
                                            (2, 32, 24, 64, 64)])
def test_deconv(n, h, w, cin, cout):
    from vitpose_b200 import ops
    from vitpose_b200.engine import pack_deconv_weight, fold_bn
    g = torch.Generator().manual_seed(cin + cout)
    x = _rand_bf16((n, cin, h, w), 40)
    wt = (torch.randn(cin, cout, 4, 4, generator=g) / math.sqrt(4 * cin)).to(BF16).float()
    gamma, beta = 1 + 0.1 * torch.randn(cout, generator=g), 0.1 * torch.randn(cout, generator=g)
    mean, var = 0.1 * torch.randn(cout, generator=g), 0.5 + torch.rand(cout, generator=g)
    scale, shift = fold_bn(gamma, beta, mean, var)
    out = ops.deconv4x4s2_bn_relu(x.permute(0, 2, 3, 1).contiguous().to(_dev()),
                                  pack_deconv_weight(wt).to(_dev()), scale.to(_dev()), shift.to(_dev()))
    ref = F.conv_transpose2d(x.float(), wt, None, stride=2, padding=1)
    ref = F.relu(F.batch_norm(ref, mean, var, gamma, beta, False, 0.0, 1e-5))
    _report('deconv', out.permute(0, 3, 1, 2), ref, 0.03)


@pytest.mark.parametrize('cin,cout', [(128, 5), (1024, 17)])
def test_conv3x3(cin, cout):
    from vitpose_b200 import ops
    n, h, w = 2, 64, 48
    g = torch.Generator().manual_seed(cin)
    x = _rand_bf16((n, cin, h, w), 50)
    wt = (torch.randn(cout, cin, 3, 3, generator=g) / math.sqrt(9 * cin)).to(BF16).float()
    bias = torch.randn(cout, generator=g)
    w9 = wt.permute(0, 2, 3, 1).reshape(cout, 9 * cin).to(BF16)
    out = ops.conv3x3_nchw(x.permute(0, 2, 3, 1).contiguous().to(_dev()), w9.to(_dev()), bias.to(_dev()))
    ref = F.conv2d(x.float(), wt, bias, padding=1)
    _report('conv3x3', out, ref, 3e-3, bf16_out=False)


def test_relu_upsample():
    from vitpose_b200 import ops
    n, h, w, C = 2, 16, 12, 128
    x = _rand_bf16((n, C, h, w), 60)
    out = ops.relu_upsample_nhwc(x.permute(0, 2, 3, 1).contiguous().to(_dev()), 4)
    ref = F.interpolate(F.relu(x.float()), scale_factor=4, mode='bilinear', align_corners=False)
    _report('relu_upsample', out.permute(0, 3, 1, 2), ref, 0.02)


def test_tokens_to_nchw():
    from vitpose_b200 import ops
    n, T, D = 2, 192, 384
    tok = _rand_bf16((n, T, D), 70)
    out = ops.tokens_to_nchw(tok.to(_dev()), 16, 12).cpu()
    ref = tok.float().permute(0, 2, 1).reshape(n, D, 16, 12)
    assert torch.equal(out, ref)
