"""Backward-pass operators (-m gpu) against torch.autograd of the same op in fp32 (on bf16-rounded operands).
These are the autograd nodes of the training-step configuration: Linear dgrad / wgrad as
GEMMs on transposed operands, GELU, LayerNorm, attention, BatchNorm(train)+ReLU and the transposed convolution."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF16 = torch.bfloat16


def _dev():
    return torch.device('cuda:0')


def _rand(shape, seed, scale=1.0, dtype=BF16):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(_dev())


def _close(name, got, ref, rel):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-12
    rms = ((got - ref).pow(2).mean().sqrt() / (ref.pow(2).mean().sqrt() + 1e-12)).item()
    assert err <= rel * scale and rms <= rel, f'{name}: max err {err:.4g} (ref absmax {scale:.4g}), rel rms {rms:.4g}'


def test_transpose_and_cast():
    from vitpose_b200 import ops
    x = _rand((3, 200, 136), 1)
    out = ops.transpose(x, batch=3)
    assert torch.equal(out, x.transpose(1, 2).contiguous())
    y = _rand((770, 64), 2)
    assert torch.equal(ops.transpose(y), y.t().contiguous())
    f = _rand((1000, 12), 3, dtype=torch.float32)
    assert torch.equal(ops.cast_bf16(f), f.to(BF16))


@pytest.mark.parametrize('R,C,f32', [(12288, 768, False), (1000, 2304, False), (64, 36864, True), (577, 18, False)])
def test_colsum(R, C, f32):
    from vitpose_b200 import ops
    x = _rand((R, C), 4, dtype=torch.float32 if f32 else BF16)
    out = torch.ones(C, device=_dev())
    ops.colsum_accumulate(x, out)
    _close('colsum', out, 1 + x.float().sum(0), 2e-4)


def test_gelu_fwd_bwd():
    from vitpose_b200 import ops
    pre = _rand((192 * 4, 1024), 5, 1.5)
    dh = _rand((192 * 4, 1024), 6)
    x = pre.float().requires_grad_(True)
    y = F.gelu(x)
    y.backward(dh.float())
    _close('gelu fwd', ops.gelu_fwd(pre), y.detach(), 1e-2)
    _close('gelu bwd', ops.gelu_bwd(pre, dh), x.grad, 1e-2)


def _gelu_grad(x):
    x = x.double()
    return 0.5 * (1 + torch.erf(x / math.sqrt(2))) + x * torch.exp(-0.5 * x * x) / math.sqrt(2 * math.pi)


@pytest.mark.parametrize('M,N,K', [(192 * 64, 3072, 768), (1000, 1536, 384), (192 * 3, 256, 64)])
def test_gemm_gelu_save_and_gelu_bwd(M, N, K):
    """The MLP fusions of the training step: fc1 + GELU that also stores the pre-activation, and fc2's input gradient
    with the GELU backward (and fc1's bias gradient) in its epilogue, against fp64 PyTorch on the same operands."""
    from vitpose_b200 import ops
    a, w = _rand((M, K), 21), _rand((N, K), 22, 1.0 / math.sqrt(K))
    bias = _rand((N,), 23, 0.5, dtype=torch.float32)
    h, pre = ops.gemm_gelu_save(a, w, bias)
    ref_pre = a.double() @ w.double().t() + bias.double()
    _close('pre-activation', pre, ref_pre, 1e-2)
    _close('gelu(pre)', h, F.gelu(ref_pre), 1e-2)
    # backward: dy [M, D], W2 [D, N] -> W2^T [N, D]
    D = 256
    dy, w2t = _rand((M, D), 24), _rand((N, D), 25, 1.0 / math.sqrt(D))
    db = torch.ones(N, device=_dev())
    dpre = ops.gemm_gelu_bwd(dy, w2t, pre, db)
    ref = (dy.double() @ w2t.double().t()) * _gelu_grad(pre)
    _close('dgelu', dpre, ref, 1e-2)
    _close('fc1 bias gradient', db, 1 + ref.sum(0), 2e-3)
    # without the column sums, and against the two-kernel path
    dpre2 = ops.gemm_gelu_bwd(dy, w2t, pre)
    assert torch.equal(dpre, dpre2)
    _close('two-kernel path', dpre, ops.gelu_bwd(pre, ops.gemm(dy, w2t, 0)), 2e-2)


@pytest.mark.parametrize('R,C,scaled', [(192 * 64, 768, True), (1000, 384, False), (192 * 5, 1280, True)])
def test_cast_bf16_colsum(R, C, scaled):
    from vitpose_b200 import ops
    x = _rand((R, C), 26, dtype=torch.float32)
    rps = 192 if scaled else 0
    scale = (torch.rand(R // 192 + 1, device=_dev()) * 2).contiguous() if scaled else None
    cs = torch.full((C,), 2.0, device=_dev())
    out = ops.cast_bf16_colsum(x, cs, scale, rps)
    ref = ops.cast_bf16(x, scale, rps)
    assert torch.equal(out, ref)
    _close('colsum of the rounded rows', cs, 2 + ref.float().sum(0), 2e-4)


@pytest.mark.parametrize('D', [128, 768])
def test_layernorm_bwd(D):
    from vitpose_b200 import ops
    M = 192 * 7 + 3
    x = _rand((M, D), 7, 2.0, torch.float32) + 0.3
    gamma, dy = _rand((D,), 8, dtype=torch.float32), _rand((M, D), 9)
    dx0 = _rand((M, D), 10, dtype=torch.float32)
    xr, gr, br = x.clone().requires_grad_(True), gamma.clone().requires_grad_(True), torch.zeros(D, device=_dev(), requires_grad=True)
    F.layer_norm(xr, (D,), gr, br, 1e-6).backward(dy.float())
    dx = dx0.clone()
    dg, db = torch.zeros(D, device=_dev()), torch.zeros(D, device=_dev())
    ops.layernorm_bwd(x, gamma, dy, dx, dg, db, 1e-6)
    _close('ln dx', dx - dx0, xr.grad, 2e-3)
    _close('ln dgamma', dg, gr.grad, 2e-3)
    _close('ln dbeta', db, br.grad, 2e-3)


@pytest.mark.parametrize('n,K,h,w,f', [(3, 17, 16, 12, 4), (2, 5, 8, 6, 2)])
def test_simple_head_gather_and_its_backward(n, K, h, w, f):
    """The un-materialised simple decoder (bilinear x f of the nine tap maps + their shifted sum) and its transpose,
    against F.interpolate + autograd (topdown_heatmap_simple_head.py:132-139,197-202 after the 1x1-tap split)."""
    from vitpose_b200 import ops
    g = torch.Generator().manual_seed(n * 100 + K)
    z = torch.randn(n, K * 9, h * w, generator=g).to(_dev()).requires_grad_(True)
    bias = torch.randn(K, generator=g).to(_dev())
    H, W = h * f, w * f
    up = F.interpolate(z.view(n, K * 9, h, w), scale_factor=f, mode='bilinear', align_corners=False).view(n, K, 9, H, W)
    pad = F.pad(up, (1, 1, 1, 1))
    ref = sum(pad[:, :, ky * 3 + kx, ky:ky + H, kx:kx + W] for ky in range(3) for kx in range(3)) + bias.view(1, K, 1, 1)
    out = ops.simple_head_gather(z.detach().contiguous(), bias, K, h, w, f)
    _close('gather', out, ref.detach(), 1e-5)
    dout = torch.randn(n, K, H, W, generator=g).to(_dev())
    ref.backward(dout)
    ldz = (9 * K + 7) // 8 * 8
    dz = ops.simple_head_gather_bwd(dout, h, w, f, ldz)                       # [n*h*w, ldz] bf16, column k*9+t
    ref_dz = z.grad.view(n, K * 9, h * w).permute(0, 2, 1).reshape(n * h * w, K * 9)
    _close('gather backward', dz[:, :9 * K], ref_dz, 1e-2)
    assert (dz[:, 9 * K:] == 0).all()


def test_relu_and_relu_bwd():
    from vitpose_b200 import ops
    x = _rand((192 * 3, 768), 51)
    x.view(-1)[:16] = torch.tensor([0.0, -0.0, 1e-30, -1e-30] * 4, device=_dev()).to(BF16)
    y = ops.relu(x)
    assert torch.equal(y, torch.relu(x))
    dy = _rand((192 * 3, 768), 52)
    assert torch.equal(ops.relu_bwd(y, dy), torch.where(y > 0, dy, torch.zeros_like(dy)))


@pytest.mark.parametrize('n,heads,hd', [(2, 2, 64), (3, 12, 64), (2, 3, 32), (3, 12, 32), (2, 2, 80), (2, 16, 80)])
def test_attention_bwd(n, heads, hd):
    """head_dim 32 / 64 / 80 = ViTPose-S / -B, -L / -H (vit.py:99-115 under autograd)"""
    from vitpose_b200 import ops
    T = 192
    D = heads * hd
    qkv = _rand((n, T, 3 * D), 11, 1.0)
    dout = _rand((n, T, D), 12, 1.0)
    out, lse = ops.attention_with_lse(qkv, heads)
    assert torch.equal(out, ops.attention(qkv, heads))
    x = qkv.float().requires_grad_(True)
    q, k, v = x.reshape(n, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
    a = ((q * hd ** -0.5) @ k.transpose(-2, -1)).softmax(-1)
    o = (a @ v).transpose(1, 2).reshape(n, T, D)
    o.backward(dout.float())
    _close('attention fwd', out, o.detach(), 2e-2)
    s_ref = ((q * hd ** -0.5) @ k.transpose(-2, -1)).detach()
    _close('attention lse', lse, torch.logsumexp(s_ref, -1) * 1.4426950408889634, 1e-3)
    dbias = torch.full((3 * D,), 0.5, device=_dev())
    dqkv = ops.attention_bwd(qkv, out, lse, dout, heads, dbias=dbias)
    assert torch.equal(dqkv, ops.attention_bwd(qkv, out, lse, dout, heads))
    _close('qkv bias gradient', dbias, 0.5 + dqkv.float().sum((0, 1)), 5e-3)
    g = x.grad.reshape(n, T, 3, D)
    d = dqkv.reshape(n, T, 3, D)
    for i, nm in enumerate('qkv'):
        _close(f'attention d{nm}', d[:, :, i], g[:, :, i], 3e-2)


def test_linear_dgrad_wgrad_as_gemms():
    """dX = dY W and dW += dY^T X through the forward GEMM kernel on transposed operands."""
    from vitpose_b200 import ops, _lib
    M, N, K = 192 * 16, 768, 3072
    x, w, dy = _rand((M, K), 13), _rand((N, K), 14, 1 / math.sqrt(K)), _rand((M, N), 15)
    dx = ops.gemm(dy, ops.transpose(w), _lib.EPI_BIAS_BF16)
    _close('dgrad', dx, dy.float() @ w.float(), 1e-2)
    dw = torch.ones(N, K, device=_dev())
    ops.gemm(ops.transpose(dy), ops.transpose(x), _lib.EPI_RESID_F32, out=dw, aux=dw)
    _close('wgrad', dw, 1 + dy.float().t() @ x.float(), 2e-3)
    # split-K accumulation (the form the training step uses), incl. a narrow output and a ragged K split
    for (n_, k_, m_) in ((N, K, M), (24, 256, 192 * 13), (256, 1024, 192 * 37)):
        x2, dy2 = _rand((m_, k_), 16), _rand((m_, n_), 17)
        dw2 = torch.ones(n_, k_, device=_dev())
        ops.gemm(ops.transpose(dy2), ops.transpose(x2), _lib.EPI_ACCUM_F32, out=dw2)
        _close(f'wgrad split-K {n_}x{k_}x{m_}', dw2, 1 + dy2.float().t() @ x2.float(), 2e-3)
        # the same product straight from dY and X (MN-major operands, no transposed copies)
        dw3 = torch.ones(n_, k_, device=_dev())
        ops.gemm_atb_accum(dy2, x2, dw3)
        _close(f'wgrad At.B {n_}x{k_}x{m_}', dw3, 1 + dy2.float().t() @ x2.float(), 2e-3)


def test_bn_train_relu_fwd_bwd():
    from vitpose_b200 import ops
    n, h, w, C = 4, 32, 24, 256
    raw = _rand((n, h, w, C), 16, 2.0) + 0.5
    dact = _rand((n, h, w, C), 17)
    gamma, beta = _rand((C,), 18, dtype=torch.float32), _rand((C,), 19, 0.5, torch.float32)
    rm, rv = torch.zeros(C, device=_dev()), torch.ones(C, device=_dev())
    x = raw.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    g, b = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm_ref, rv_ref = rm.clone(), rv.clone()
    y = F.relu(F.batch_norm(x, rm_ref, rv_ref, g, b, True, 0.1, 1e-5))
    y.backward(dact.float().permute(0, 3, 1, 2))
    mean, rstd = ops.bn_train_stats(raw, 1e-5, 0.1, rm, rv)
    _close('bn running_mean', rm, rm_ref, 1e-3)
    _close('bn running_var', rv, rv_ref, 1e-3)
    act = ops.bn_relu_fwd(raw, mean, rstd, gamma, beta)
    _close('bn+relu fwd', act, y.detach().permute(0, 2, 3, 1), 1e-2)
    dg, db = torch.zeros(C, device=_dev()), torch.zeros(C, device=_dev())
    draw = ops.bn_relu_bwd(raw, dact, mean, rstd, gamma, beta, dg, db)
    _close('bn dgamma', dg, g.grad, 5e-3)
    _close('bn dbeta', db, b.grad, 5e-3)
    _close('bn+relu bwd', draw, x.grad.permute(0, 2, 3, 1), 2e-2)
    # BatchNorm2d.eval() inside forward_train: running statistics, no batch-mean terms in the input gradient
    rm2, rv2 = _rand((C,), 20, 0.3, torch.float32), _rand((C,), 21, 0.2, torch.float32).abs() + 0.5
    x2 = raw.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    g2, b2 = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    y2 = F.relu(F.batch_norm(x2, rm2, rv2, g2, b2, False, 0.1, 1e-5))
    y2.backward(dact.float().permute(0, 3, 1, 2))
    rstd2 = torch.rsqrt(rv2 + 1e-5)
    _close('bn(eval)+relu fwd', ops.bn_relu_fwd(raw, rm2, rstd2, gamma, beta), y2.detach().permute(0, 2, 3, 1), 1e-2)
    dg2, db2 = torch.zeros(C, device=_dev()), torch.zeros(C, device=_dev())
    draw2 = ops.bn_relu_bwd(raw, dact, rm2, rstd2, gamma, beta, dg2, db2, eval_mode=True)
    _close('bn(eval) dgamma', dg2, g2.grad, 5e-3)
    _close('bn(eval) dbeta', db2, b2.grad, 5e-3)
    _close('bn(eval)+relu bwd', draw2, x2.grad.permute(0, 2, 3, 1), 2e-2)


@pytest.mark.parametrize('n,h,w,cin,cout', [(2, 16, 12, 128, 64), (3, 32, 24, 256, 256)])
def test_deconv_fwd_bwd(n, h, w, cin, cout):
    """ConvTranspose2d(k4,s2,p1): raw forward, dgrad and wgrad (in the packed 4-phase layout) as gathers + GEMMs."""
    from vitpose_b200 import ops, _lib
    from vitpose_b200.engine import pack_deconv_weight, pack_deconv_weight_dgrad
    x = _rand((n, h, w, cin), 20)
    wt = _rand((cin, cout, 4, 4), 21, 0.05)
    dy = _rand((n, 2 * h, 2 * w, cout), 22)
    xr, wr = x.float().permute(0, 3, 1, 2).contiguous().requires_grad_(True), wt.float().requires_grad_(True)
    y = F.conv_transpose2d(xr, wr, None, stride=2, padding=1)
    y.backward(dy.float().permute(0, 3, 1, 2))
    wp = pack_deconv_weight(wt)
    raw = ops.deconv4x4s2_raw(x, wp)
    _close('deconv raw', raw, y.detach().permute(0, 2, 3, 1), 1e-2)
    # dgrad: [pixels, 16*cout] x W2g[cin, 16*cout]^T
    dx = ops.gemm(ops.deconv_gather_dy(dy), pack_deconv_weight_dgrad(wp), _lib.EPI_BIAS_BF16)
    _close('deconv dgrad', dx.reshape(n, h, w, cin), xr.grad.permute(0, 2, 3, 1), 1e-2)
    # wgrad per phase: dWp[ph] [cout, 4*cin] = phase_dy[ph]^T . gather_x[ph]
    a = ops.transpose(ops.deconv_phase_dy(dy), batch=4)        # [4, cout, pixels]
    b = ops.transpose(ops.deconv_gather_x(x), batch=4)         # [4, 4*cin, pixels]
    dwp = torch.zeros(4, cout, 4 * cin, device=_dev())
    for ph in range(4):
        ops.gemm(a[ph], b[ph], _lib.EPI_ACCUM_F32, out=dwp[ph])
    ref = pack_deconv_weight(wr.grad)          # same re-layout applied to the reference gradient (fp32 in, bf16 out)
    _close('deconv wgrad', dwp, ref, 1e-2)


def test_cast_transpose_multi_matches_per_layer_ops():
    """One launch for the bf16 W / W^T copies of many layers == cast_bf16 + transpose per layer, bit for bit
    (ragged / odd shapes exercise partial 64 x 64 tiles and the element-wise path)."""
    import torch.nn as nn
    from vitpose_b200 import ops
    from vitpose_b200.training import _LinearBank
    torch.manual_seed(0)
    dev = torch.device('cuda:0')
    layers = [nn.Linear(45, 70), nn.Linear(768, 96), nn.Linear(33, 31, bias=False), nn.Conv2d(3, 40, 16, 16),
              nn.Linear(64, 2304), nn.Linear(130, 66)]
    layers = [l.to(dev) for l in layers]
    bank = _LinearBank(layers)
    for rep in range(2):                       # second pass: weights updated in place, table reused
        lins = bank.refresh()
        torch.cuda.synchronize()
        for l, lin in zip(layers, lins):
            w_ref = l.weight.detach().reshape(l.weight.shape[0], -1).to(torch.bfloat16)   # round to nearest even
            if w_ref.numel() % 4 == 0:
                assert torch.equal(w_ref, ops.cast_bf16(l.weight.detach().reshape(l.weight.shape[0], -1).contiguous()))
            assert torch.equal(lin.w, w_ref)
            assert torch.equal(lin.wt, w_ref.t().contiguous())
            assert lin.w.data_ptr() % 16 == 0 and lin.wt.data_ptr() % 16 == 0
        with torch.no_grad():
            for l in layers:
                l.weight.mul_(1.5).add_(0.01)


def test_deconv_weight_pack_kernels_match_torch_formulation():
    """vpb_deconv_pack_weight / vpb_deconv_unpack_wgrad (one launch each) == engine.pack_deconv_weight,
    pack_deconv_weight_dgrad and unpack_deconv_weight (the torch slice-copy formulations), bit for bit."""
    from vitpose_b200 import ops
    from vitpose_b200.engine import pack_deconv_weight, pack_deconv_weight_dgrad, unpack_deconv_weight
    g = torch.Generator().manual_seed(5)
    for cin, cout in ((768, 256), (256, 256), (64, 32)):
        w = (torch.randn(cin, cout, 4, 4, generator=g) * 0.05).to(_dev())
        wp, wd = ops.deconv_pack_weight(w)
        ref = pack_deconv_weight(w)
        assert torch.equal(wp, ref)
        assert torch.equal(wd, pack_deconv_weight_dgrad(ref))
        wp2, none = ops.deconv_pack_weight(w, want_dgrad=False)
        assert none is None and torch.equal(wp2, ref)
        dwp = torch.randn(4, cout, 4 * cin, generator=g).to(_dev())
        out = torch.empty(cin, cout, 4, 4, device=_dev())
        assert torch.equal(ops.deconv_unpack_wgrad(dwp, out), unpack_deconv_weight(dwp))
