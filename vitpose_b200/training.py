"""Training step of the ViTPose top-down path: ``TopDown.forward_train`` ->
``loss.backward()`` -> layer-decay AdamW, with every tensor operation in libvitpose_b200.so.

The reference trains through torch.autograd over eager modules (mmpose/models/detectors/top_down.py:143-161 for
``forward_train``; ViT.forward vit.py:313-337; TopdownHeatmapSimpleHead.forward simple_head.py:197-202; BatchNorm2d in
training mode inside the deconv stack, simple_head.py:324-333).  Here the whole network is ONE autograd node,
:class:`_NetworkFn`: its forward launches the training variant of the forward pass (activations kept, BatchNorm on
batch statistics, GELU unfused so the pre-activation survives), its backward launches the hand-written backward
pass and returns the gradient of every parameter, so ``loss.backward()`` / ``optimizer.step()`` of an unmodified
mmcv-style training loop work as before.  Python here only sequences C-ABI calls (the reference's runner is
Python too); there is no eager fallback.

Numerics: GEMM operands and the activation gradients between operators are bf16, the residual stream, its
gradient, all parameter gradients, LayerNorm / BatchNorm statistics and the optimizer state are fp32.
"""
import os

import torch

from . import _lib, ops

BF16 = torch.bfloat16
EPI_BIAS, EPI_RESID, EPI_POS, EPI_NCHW = _lib.EPI_BIAS_BF16, _lib.EPI_RESID_F32, _lib.EPI_POS_F32, _lib.EPI_NCHW_F32


def _wgrad(dy, x, out):
    """out [N, K] fp32 += dY^T X from dY [M, N] and X [M, K] as they are (row-major): both tiles are consumed as
    MN-major tensor-core operands and the contraction (all M token rows of the batch) is split over CTAs, because
    N x K alone is only a handful of tiles."""
    ops.gemm_atb_accum(dy, x, out)


class _Linear:
    """bf16 operand copies of one nn.Linear: W [out, in] for the forward, W^T [in, out] for the input gradient."""

    def __init__(self, lin, w=None, wt=None):
        if w is None:
            w = ops.cast_bf16(lin.weight.detach().reshape(lin.weight.shape[0], -1).contiguous())
            wt = ops.transpose(w)
        self.w, self.wt = w, wt
        self.b = lin.bias.detach() if lin.bias is not None else None


class _LinearBank:
    """Persistent bf16 W / W^T buffers of a list of linear layers, refreshed from the fp32 master parameters by ONE
    vpb_cast_transpose_multi launch per training step (one cast + one transpose launch per layer were launch-bound:
    125 launches, 0.82 ms per step). The descriptor table is rebuilt only when a parameter's storage moves. The
    buffers are shared by successive steps: a backward pass must run before the parameters it was recorded with are
    updated (the reference's autograd raises a version-counter error in that situation)."""

    _ENTRY = None

    def __init__(self, layers):
        import numpy as np
        if _LinearBank._ENTRY is None:
            _LinearBank._ENTRY = np.dtype([('src', np.uint64), ('w', np.uint64), ('wt', np.uint64),
                                           ('rows', np.int32), ('cols', np.int32)], align=True)
            assert _LinearBank._ENTRY.itemsize == 32
        self.layers = list(layers)
        dev = self.layers[0].weight.device
        shapes = [(l.weight.shape[0], l.weight.numel() // l.weight.shape[0]) for l in self.layers]
        total = sum(r * c for r, c in shapes)
        self.arena = torch.empty(2 * total + 16 * len(shapes), device=dev, dtype=torch.bfloat16)
        self.views, off = [], 0
        for r, c in shapes:
            w = self.arena[off:off + r * c].view(r, c)
            off += (r * c + 7) // 8 * 8                     # 16-byte aligned operands (TMA)
            wt = self.arena[off:off + r * c].view(c, r)
            off += (r * c + 7) // 8 * 8
            self.views.append((w, wt))
        self.shapes = shapes
        self._ptrs = None
        self._table = self._starts = None
        self._total_tiles = 0

    def _build_table(self):
        import numpy as np
        table = np.zeros(len(self.layers), dtype=self._ENTRY)
        starts = np.zeros(len(self.layers) + 1, dtype=np.int32)
        for i, (l, (w, wt), (r, c)) in enumerate(zip(self.layers, self.views, self.shapes)):
            if not l.weight.is_contiguous() or l.weight.dtype != torch.float32:
                raise _lib.VitposeLibError('training expects contiguous fp32 master weights')
            table[i] = (l.weight.data_ptr(), w.data_ptr(), wt.data_ptr(), r, c)
            starts[i + 1] = starts[i] + ((r + 63) // 64) * ((c + 63) // 64)
        dev = self.arena.device
        self._table = torch.from_numpy(table.view(np.uint8)).to(dev)
        self._starts = torch.from_numpy(starts).to(dev)
        self._total_tiles = int(starts[-1])

    def refresh(self):
        ptrs = tuple(l.weight.data_ptr() for l in self.layers)
        if ptrs != self._ptrs:
            self._build_table()
            self._ptrs = ptrs
        _lib.check(_lib.lib().vpb_cast_transpose_multi(_lib.ptr(self._table), _lib.ptr(self._starts), len(self.layers),
                                                       self._total_tiles, _lib.stream_ptr()),
                   'vpb_cast_transpose_multi')
        return [_Linear(l, w, wt) for l, (w, wt) in zip(self.layers, self.views)]


# VPB_TRAIN_FUSE=0: the un-fused MLP / bias-gradient kernels of round 1 (A/B measurements, parity tests of both paths)
FUSE_MLP = os.environ.get('VPB_TRAIN_FUSE', '1') != '0'


def drop_path_scales(bb, n, device):
    """[(s_attn, s_mlp)] per block: fp32 [n] factors mask / keep_prob of timm's drop_path (per-sample Bernoulli), or
    (None, None) where the branch is always kept. ``bb._drop_path_scales`` (same structure) overrides the random
    draw — the parity tests inject the masks the oracle uses."""
    forced = getattr(bb, '_drop_path_scales', None)
    if forced is not None:
        return forced
    rates = torch.linspace(0, bb.drop_path_rate, bb.depth)
    if not bb.training or float(rates.max()) <= 0.0:
        return [(None, None)] * bb.depth
    # all 2 * depth Bernoulli draws in one shot (a handful of launches instead of ~8 per block)
    keep = (1.0 - rates).to(device=device, dtype=torch.float32).view(-1, 1, 1)
    masks = torch.floor(keep + torch.rand(bb.depth, 2, n, device=device)).div_(keep).contiguous()
    # a frozen block is in eval mode (vit.py:257-259), where DropPath is the identity (vit.py:55-56)
    blocks = getattr(bb, 'blocks', None)
    frozen = [blocks is not None and not blocks[i].training for i in range(bb.depth)]
    return [(None, None) if p <= 0.0 or frozen[i] else (masks[i, 0], masks[i, 1]) for i, p in enumerate(rates.tolist())]


_GRAD_GROUPS = {}


def gradient_process_group():
    """(group, has_avg) for the gradient all-reduce. ``VPB_NCCL_MAX_CTAS=<n>`` gives the exchange an NCCL communicator
    of its own whose kernels are limited to n CTAs (``ncclConfig_t.maxCTAs``), because the all-reduce overlaps the
    persistent one-CTA-per-SM GEMMs of the backward pass. Measured in round 2 (2 GPUs, 64 crops per GPU, ReduceOp.AVG):
    15.28 ms per step with NCCL's default CTA count, 15.79 ms with 8 CTAs (14.79 ms on one GPU) — the default is kept
    (round 1, with a SUM + separate scaling pass: 17.5 ms default, 16.3 ms with 8 channels). Other backends (gloo in
    the CPU tests) use the default group and SUM + scale."""
    dist = torch.distributed
    key = dist.get_backend()
    if key not in _GRAD_GROUPS:
        group, has_avg = None, False
        if key == 'nccl':
            has_avg = True
            max_ctas = int(os.environ.get('VPB_NCCL_MAX_CTAS', '0'))
            if max_ctas > 0:
                opts = dist.ProcessGroupNCCL.Options()
                opts.config.max_ctas = max_ctas
                group = dist.new_group(backend='nccl', pg_options=opts)     # collective: every rank gets here together
        _GRAD_GROUPS[key] = (group, has_avg)
    return _GRAD_GROUPS[key]


def _param_list(model):
    """(name, parameter) in a fixed order: the inputs of the autograd node. Cached on the model (walking
    named_parameters() five times per step cost ~1 ms of host time); the Parameter objects of a built model are stable
    (load_state_dict copies in place) — `del model._vpb_param_list` after replacing sub-modules."""
    cached = model.__dict__.get('_vpb_param_list')
    if cached is None:
        cached = [(n, p) for n, p in model.named_parameters()]
        model.__dict__['_vpb_param_list'] = cached
    return cached


def _heads_of(model):
    """[(parameter-name prefix, head)]: the keypoint head, then TopDownMoE's associate heads (top_down_moe.py:75-89)."""
    out = [('keypoint_head.', model.keypoint_head)]
    for i, h in enumerate(getattr(model, 'associate_keypoint_heads', None) or ()):
        out.append((f'associate_keypoint_heads.{i}.', h))
    return out


def _head_kind(head):
    classic = head.num_deconv_layers == 2 and head.final_conv_kernel == 1
    simple = head.num_deconv_layers == 0 and head.final_conv_kernel == 3 and int(getattr(head, 'upsample', 0)) > 1
    if not (classic or simple):
        raise NotImplementedError('the training step is built for the classic decoder (2 deconv layers, 1x1 conv) '
                                  'and the simple decoder (upsample + 3x3 conv)')
    return 'simple' if simple else 'classic'


def _run_scratch(model, M, D, device):
    """statistics-exchange scratch of the fused-LayerNorm GEMMs for a run of M token rows (ViTPose+ batches)."""
    cache = model.__dict__.setdefault('_vpb_ln_scratch_runs', {})
    sc = cache.get((M, D))
    if sc is None or sc.buf.device != device:
        sc = cache[(M, D)] = ops.LnScratch(M, D, device)
    return sc


def _head_forward(head, xn, n, hp, wp, D, T):
    """TopdownHeatmapSimpleHead.forward (simple_head.py:197-202) on the normalised tokens xn bf16 [n * T, D], keeping
    what the backward pass needs. Returns (heatmaps fp32 [n, K, H, W], saved)."""
    if _head_kind(head) == 'simple':
        # ---- simple decoder (simple_head.py:132-139,197-202: ReLU -> bilinear x f -> Conv2d 3x3) without the
        # upsampled map: the 3x3 conv on the token grid as nine 1x1 convs (one GEMM with 9K columns), then a
        # bilinear gather of the nine tap maps (as vpb_vitpose_forward, csrc/api.cu)
        fl = head.final_layer
        K, f = fl.weight.shape[0], int(head.upsample)
        r = ops.relu(xn)
        w9 = ops.cast_bf16(fl.weight.detach().permute(0, 2, 3, 1).reshape(K * 9, D).contiguous())
        z = ops.gemm(r, w9, EPI_NCHW, period=T)                                  # [n, 9K, T] fp32
        hm = ops.simple_head_gather(z, fl.bias.detach().float().contiguous(), K, hp, wp, f)
        return hm, dict(simple=True, relu=r, w9=w9, K=K, f=f)
    # ---- classic decoder, BatchNorm2d in training mode (simple_head.py:324-333)
    cur, hs = xn.view(n, hp, wp, D), []
    for i in range(2):
        dw = head.deconv_layers[3 * i].weight.detach()
        bn = head.deconv_layers[3 * i + 1]
        wp_, wd_ = ops.deconv_pack_weight(dw.contiguous())     # forward + input-gradient operands, one launch
        raw = ops.deconv4x4s2_raw(cur, wp_)
        if bn.training:
            mean, rstd = ops.bn_train_stats(raw, bn.eps, bn.momentum, bn.running_mean, bn.running_var)
            bn.num_batches_tracked += 1
        else:
            mean, rstd = bn.running_mean.detach().clone(), torch.rsqrt(bn.running_var.detach() + bn.eps)
        act = ops.bn_relu_fwd(raw, mean, rstd, bn.weight.detach(), bn.bias.detach())
        hs.append(dict(x=cur, wp=wp_, wd=wd_, raw=raw, mean=mean, rstd=rstd, bn=bn, frozen_stats=not bn.training))
        cur = act
    fl = head.final_layer
    K = fl.weight.shape[0]
    wf = ops.cast_bf16(fl.weight.detach().reshape(K, -1).contiguous())
    P = cur.shape[1] * cur.shape[2]
    hm = ops.gemm(cur.view(n * P, -1), wf, EPI_NCHW, bias=fl.bias.detach(), period=P)
    return hm.view(n, K, cur.shape[1], cur.shape[2]), dict(simple=False, head=hs, act_last=cur, wf=wf, K=K, P=P)


def _head_backward(pfx, head, sv, dhm, s, g, zeros, scratch_zeros):
    """Backward of _head_forward: parameter gradients into g[pfx + ...], returns the gradient of xn (bf16 [M, D])."""
    n, T, D = s['n'], s['T'], s['D']
    K = sv['K']
    dev = dhm.device
    if sv['simple']:
        # ---- simple decoder: bias, then the transposed gather, then the tap GEMM's weight / input gradients
        hp, wp = s['hw']
        f = sv['f']
        P_out = hp * f * wp * f
        dhm32 = dhm.contiguous().float()
        Kp = (K + 7) // 8 * 8
        dbf = zeros(Kp)
        ops.colsum_accumulate(ops.nchw_to_rows(dhm32.view(n, K, P_out), Kp), dbf)
        g[pfx + 'final_layer.bias'] = dbf[:K]
        Kp9 = (9 * K + 7) // 8 * 8
        dz = ops.simple_head_gather_bwd(dhm32, hp, wp, f, Kp9)                       # [M, Kp9], column k*9+t
        dw9 = scratch_zeros(Kp9, D)
        _wgrad(dz, sv['relu'], dw9)
        dwf = zeros(K, D, 3, 3)
        dwf.copy_(dw9[:9 * K].view(K, 3, 3, D).permute(0, 3, 1, 2))
        g[pfx + 'final_layer.weight'] = dwf
        w9_pad = torch.zeros(Kp9, D, device=dev, dtype=BF16)
        w9_pad[:9 * K] = sv['w9']
        return ops.relu_bwd(sv['relu'], ops.gemm(dz, ops.transpose(w9_pad), EPI_BIAS))   # [M, D]
    # ---- final 1x1 conv: rows = pixels, columns = keypoints (zero padded to a multiple of 8)
    P = sv['P']
    Kp = (K + 7) // 8 * 8
    dy = ops.nchw_to_rows(dhm.contiguous().float().view(n, K, P), Kp)                # [n*P, Kp]
    act = sv['act_last'].view(n * P, -1)
    C = act.shape[1]
    dwf = zeros(Kp, C)
    _wgrad(dy, act, dwf)
    dbf = zeros(Kp)
    ops.colsum_accumulate(dy, dbf)
    g[pfx + 'final_layer.weight'] = dwf[:K].reshape(K, C, 1, 1)
    g[pfx + 'final_layer.bias'] = dbf[:K]
    wf_pad = torch.zeros(Kp, C, device=dev, dtype=BF16)
    wf_pad[:K] = sv['wf']
    dact = ops.gemm(dy, ops.transpose(wf_pad), EPI_BIAS)                             # [n*P, C]
    # ---- [ConvTranspose2d -> BatchNorm2d(train) -> ReLU] x 2, last to first
    for i in (1, 0):
        hsi = sv['head'][i]
        bn, raw, xin, wp_ = hsi['bn'], hsi['raw'], hsi['x'], hsi['wp']
        cout = raw.shape[-1]
        dgam, dbet = zeros(cout), zeros(cout)
        draw = ops.bn_relu_bwd(raw, dact.view(raw.shape), hsi['mean'], hsi['rstd'], bn.weight.detach(),
                               bn.bias.detach(), dgam, dbet, eval_mode=hsi['frozen_stats'])
        g[pfx + f'deconv_layers.{3 * i + 1}.weight'] = dgam
        g[pfx + f'deconv_layers.{3 * i + 1}.bias'] = dbet
        a_t = ops.deconv_phase_dy(draw)                                             # [4, pixels, cout]
        b_t = ops.deconv_gather_x(xin)                                               # [4, pixels, 4*cin]
        dwp = scratch_zeros(4, cout, wp_.shape[2])
        for ph in range(4):
            _wgrad(a_t[ph], b_t[ph], dwp[ph])
        g[pfx + f'deconv_layers.{3 * i}.weight'] = ops.deconv_unpack_wgrad(dwp, zeros(wp_.shape[2] // 4, cout, 4, 4))
        del a_t, b_t
        dact = ops.gemm(ops.deconv_gather_dy(draw), hsi['wd'], EPI_BIAS)             # [pixels_in, cin]
    return dact.view(-1, D)


class _NetworkFn(torch.autograd.Function):
    """img [N,3,H,W] fp32 (+ every parameter of backbone and head) -> heatmaps [N,K,H/4,W/4] fp32."""

    @staticmethod
    def forward(ctx, model, img, *params):
        bb, head = model.backbone, model.keypoint_head
        if not img.is_cuda:
            raise _lib.VitposeLibError('forward_train needs CUDA tensors (vitpose_b200 has no CPU path)')
        for _, head_i in _heads_of(model):
            _head_kind(head_i)
        img = img.contiguous().float()
        n = img.shape[0]
        D, heads, depth = bb.embed_dim, bb.num_heads, bb.depth
        hp, wp = bb.patch_embed.patch_shape
        T, M = hp * wp, n * hp * wp
        s = {}                                           # saved activations / operands for the backward pass
        # ---- operands from the current fp32 master parameters
        bank = getattr(model, '_vpb_linear_bank', None)
        # ViTPose+ (ViTMoE, vit_moe.py:77-115): the FFN output of a crop is cat(fc2(h), experts[dataset](h)). The crops
        # arrive sorted by dataset (TopDownMoE.forward_train); `runs` = [(dataset, crops)] in batch order.
        E = len(bb.blocks[0].mlp.experts) if hasattr(bb.blocks[0].mlp, 'experts') else 0
        runs = getattr(model, '_vpb_dataset_runs', None) or [(0, n)]
        if E == 0:
            runs = [(0, n)]
        elif sum(c for _, c in runs) != n or any(not 0 <= d < E for d, _ in runs):
            raise IndexError(f'dataset runs {runs} do not describe a batch of {n} crops over {E} experts')
        per_blk = 4 + E
        layers = [bb.patch_embed.proj] + [m for blk in bb.blocks
                                          for m in (blk.attn.qkv, blk.attn.proj, blk.mlp.fc1, blk.mlp.fc2,
                                                    *(blk.mlp.experts if E else ()))]
        if bank is None or bank.layers != layers or bank.arena.device != img.device:
            bank = _LinearBank(layers)
            model._vpb_linear_bank = bank
        lins = bank.refresh()                            # one launch: bf16 W and W^T of every linear layer
        pe = lins[0]
        pos = bb.pos_embed.detach()
        pos_tok = (pos[0, 1:] + pos[0, :1]).contiguous()
        blocks = []
        for i in range(len(bb.blocks)):
            q, pr, f1, f2 = lins[1 + per_blk * i:5 + per_blk * i]
            w = dict(qkv=q, proj=pr, fc1=f1, fc2=f2)
            if E:
                # the effective fc2 of dataset d (tools/model_split.py:36-40): rows / bias of fc2, then of experts[d]
                ex = lins[5 + per_blk * i:5 + per_blk * i + E]
                w['experts'] = ex
                w['cat'] = {d: (torch.cat([f2.w, ex[d].w], 0), torch.cat([f2.wt, ex[d].wt], 1),
                                torch.cat([f2.b, ex[d].b], 0)) for d in sorted({d for d, _ in runs})}
            blocks.append(w)
        # ---- ViT (vit.py:313-332). Stochastic depth (DropPath, vit.py:48-56,132,138-139,233): block i drops the
        # whole residual branch of a crop with probability linspace(0, drop_path_rate, depth)[i]; the surviving
        # branches are divided by keep_prob. The per-crop factor goes into the residual GEMM epilogue.
        scales = drop_path_scales(bb, n, img.device)
        patches = ops.im2col_patch16(img, flip=False)
        b0 = bb.blocks[0]
        # one statistics-exchange scratch for the 25 fused-LayerNorm GEMMs of a step (kept on the model across steps)
        lns = getattr(model, '_vpb_ln_scratch', None)
        if lns is None or (lns.M, lns.N) != (M, D) or lns.buf.device != img.device:
            lns = model._vpb_ln_scratch = ops.LnScratch(M, D, img.device)
        x, xn = ops.gemm_layernorm(patches, pe.w, EPI_POS, pe.b, pos_tok, b0.norm1.weight.detach(),
                                   b0.norm1.bias.detach(), 1e-6, period=T, scratch=lns)
        acts = []
        for l, blk in enumerate(bb.blocks):
            w = blocks[l]
            a = dict(x_in=x, xn1=xn)
            qkv = ops.gemm(xn, w['qkv'].w, EPI_BIAS, bias=w['qkv'].b)
            attn, lse = ops.attention_with_lse(qkv.view(n, T, 3 * D), heads)
            attn = attn.view(M, D)
            s1, s2 = scales[l]
            x_mid, xn2 = ops.gemm_layernorm(attn, w['proj'].w, EPI_RESID, w['proj'].b, x, blk.norm2.weight.detach(),
                                            blk.norm2.bias.detach(), 1e-6, row_scale=s1, rows_per_scale=T, scratch=lns)
            if FUSE_MLP:      # one kernel: h = gelu(pre) for fc2 and the bf16 pre-activation for the backward pass
                h, pre = ops.gemm_gelu_save(xn2, w['fc1'].w, w['fc1'].b)
            else:
                pre = ops.gemm(xn2, w['fc1'].w, EPI_BIAS, bias=w['fc1'].b)
                h = ops.gelu_fwd(pre)
            nxt = bb.blocks[l + 1].norm1 if l + 1 < depth else bb.last_norm
            if E:
                # one launch per run of crops with that dataset's effective fc2 (as vpb_moe_runs does in inference)
                x = torch.empty(M, D, device=img.device, dtype=torch.float32)
                xn = torch.empty(M, D, device=img.device, dtype=BF16)
                c0 = 0
                for d, cnt in runs:
                    r0, r1 = c0 * T, (c0 + cnt) * T
                    wc, _, bc = w['cat'][d]
                    ops.gemm_layernorm(h[r0:r1], wc, EPI_RESID, bc, x_mid[r0:r1], nxt.weight.detach(),
                                       nxt.bias.detach(), 1e-6, row_scale=None if s2 is None else s2[c0:c0 + cnt],
                                       rows_per_scale=T, scratch=_run_scratch(model, r1 - r0, D, img.device),
                                       out=x[r0:r1], xn=xn[r0:r1])
                    c0 += cnt
            else:
                x, xn = ops.gemm_layernorm(h, w['fc2'].w, EPI_RESID, w['fc2'].b, x_mid, nxt.weight.detach(),
                                           nxt.bias.detach(), 1e-6, row_scale=s2, rows_per_scale=T, scratch=lns)
            a.update(qkv=qkv, attn=attn, lse=lse, x_mid=x_mid, xn2=xn2, pre=pre, h=h, s1=s1, s2=s2)
            acts.append(a)
        s.update(patches=patches, acts=acts, x_final=x, blocks=blocks, pe=pe, runs=runs, E=E,
                 n=n, T=T, M=M, D=D, heads=heads, hw=(hp, wp))
        # ---- keypoint heads: TopDown has one; TopDownMoE runs the main head and every associate head on the whole
        # batch (top_down_moe.py:180-201) and returns one heatmap tensor per head
        outs, saved = [], []
        for _, head_i in _heads_of(model):
            hm, sv = _head_forward(head_i, xn, n, hp, wp, D, T)
            outs.append(hm)
            saved.append(sv)
        s['head_saved'] = saved
        ctx.s, ctx.model, ctx.names = s, model, [nm for nm, _ in _param_list(model)]
        return outs[0] if len(outs) == 1 else tuple(outs)

    @staticmethod
    def backward(ctx, *dhms):
        s, model = ctx.s, ctx.model
        bb = model.backbone
        n, T, M, D, heads = s['n'], s['T'], s['M'], s['D'], s['heads']
        runs, E = s['runs'], s['E']
        dev = next(d for d in dhms if d is not None).device
        g = {}                                           # parameter name -> fp32 gradient

        # Every parameter gradient lives in ONE zero-initialised arena, in the order the backward pass produces
        # them: one memset instead of ~270 fills, and with N > 1 ranks the arena is all-reduced (NCCL) in a few
        # large segments that are issued as soon as they are complete, so the exchange overlaps the rest of the
        # backward pass; when backward() returns the gradients are already averaged (what the DDP wrapper does).
        cap = sum((p.numel() + 63) // 64 * 64 for _, p in _param_list(model)) + (1 << 20)
        arena = torch.zeros(cap, device=dev, dtype=torch.float32)
        used = [0]

        def zeros(*shape):
            numel = 1
            for d in shape:
                numel *= d
            if used[0] + numel > cap:
                raise RuntimeError('gradient arena exhausted')
            t = arena[used[0]:used[0] + numel].view(*shape)
            used[0] += (numel + 63) // 64 * 64
            return t

        def scratch_zeros(*shape):                       # zero-initialised buffers that are not parameter gradients
            return torch.zeros(*shape, device=dev, dtype=torch.float32)

        world = 1
        if getattr(model, 'allreduce_in_backward', True) and torch.distributed.is_available() \
                and torch.distributed.is_initialized():
            world = torch.distributed.get_world_size()
        pending, sent = [], [0]
        group, has_avg = gradient_process_group() if world > 1 else (None, False)
        # optional bf16 exchange (VPB_GRAD_BF16=1 or model.grad_allreduce_dtype = torch.bfloat16): half the bytes on
        # NVLink; the reference's DDP exchanges fp32, so fp32 stays the default
        bf16_exchange = world > 1 and (getattr(model, 'grad_allreduce_dtype', None) == BF16 or
                                       os.environ.get('VPB_GRAD_BF16', '0') == '1')
        compressed = []

        def exchange(final=False):
            """all-reduce (average) the arena segment completed since the last call (at least 32 MB unless final)."""
            if world == 1 or used[0] == sent[0] or (not final and (used[0] - sent[0]) * 4 < (32 << 20)):
                return
            dist = torch.distributed
            seg = arena[sent[0]:used[0]]
            op = dist.ReduceOp.AVG if has_avg else dist.ReduceOp.SUM
            if bf16_exchange:
                half = seg.to(BF16)
                compressed.append((seg, half))
                pending.append(dist.all_reduce(half, op=op, group=group, async_op=True))
            else:
                pending.append(dist.all_reduce(seg, op=op, group=group, async_op=True))
            sent[0] = used[0]

        # ---- keypoint heads, each back to the gradient of the backbone features (bf16 [M, D]); a head whose output
        # took no part in the loss gets no gradient (its parameters' .grad stay None, as with autograd)
        dact = None
        for (pfx, head_i), sv, dhm in zip(_heads_of(model), s['head_saved'], dhms):
            if dhm is None:
                continue
            d_i = _head_backward(pfx, head_i, sv, dhm, s, g, zeros, scratch_zeros)
            if dact is None:
                dact = d_i
            else:                        # several heads: summed in fp32, rounded once
                dact = (dact.float() if dact.dtype == BF16 else dact).add_(d_i.float())
        if dact.dtype != BF16:
            dact = dact.to(BF16)
        # ---- last_norm, then the blocks in reverse
        exchange()
        dx = scratch_zeros(M, D)
        trainable = {nm for nm, p in _param_list(model) if p.requires_grad}

        def linear_bwd(name, lin, dy_bf16, x_bf16, want_dx=True, dbias=None):
            """gradients of y = x W^T + b given dy: dW, db into g[...]; returns dx (bf16). Frozen tensors
            (``requires_grad = False``: frozen_stages / freeze_attn / freeze_ffn, vit.py:249-284) are skipped.
            ``dbias``: the bias gradient if the kernel that produced dy already summed its columns."""
            if name + '.weight' in trainable:
                dw = zeros(*lin.w.shape)
                _wgrad(dy_bf16, x_bf16, dw)
                g[name + '.weight'] = dw
            if name + '.bias' in trainable:
                if dbias is None:
                    dbias = zeros(lin.w.shape[0])
                    ops.colsum_accumulate(dy_bf16, dbias)
                g[name + '.bias'] = dbias
            return ops.gemm(dy_bf16, lin.wt, EPI_BIAS) if want_dx else None

        def norm_bwd(pfx, norm, x_saved, dy_bf16, scale, bias_name):
            """LayerNorm backward into the residual gradient dx (+ dgamma / dbeta into g), then the bf16 output gradient
            of the residual branch that ends in this stream: dx times the stochastic-depth factor of the crop, and the
            bias gradient of the branch's last Linear layer (None if frozen; from the cast's own pass when FUSE_MLP)."""
            dg_, db_ = zeros(D), zeros(D)
            g[pfx + '.weight'], g[pfx + '.bias'] = dg_, db_
            ops.layernorm_bwd(x_saved, norm.weight.detach(), dy_bf16, dx, dg_, db_, 1e-6)
            if FUSE_MLP and bias_name in trainable:
                dbias = zeros(D)
                return ops.cast_bf16_colsum(dx, dbias, scale, T), dbias
            return ops.cast_bf16(dx, scale, T), None

        def moe_fc2_bwd(pfx, w, dyb, a, db2):
            """Backward of MoEMlp's second half (vit_moe.py:97-115): the shared fc2 owns the first D - part columns of
            the branch gradient dyb on every token, experts[d] the last `part` columns on the tokens of dataset d;
            every expert gets a gradient (zero if its dataset is not in the batch), as the reference's dense masked
            form does "to support ddp training". Returns (gradient of fc1's pre-activation, fc1's bias gradient)."""
            part = w['experts'][0].w.shape[0]
            Ds, H4 = D - part, w['fc2'].w.shape[1]
            h, pre = a['h'], a['pre']
            if pfx + 'fc2.weight' in trainable:
                dw = zeros(Ds, H4)
                _wgrad(dyb[:, :Ds], h, dw)
                g[pfx + 'fc2.weight'] = dw
            if pfx + 'fc2.bias' in trainable:
                if db2 is None:
                    db2 = zeros(D)
                    ops.colsum_accumulate(dyb, db2)
                g[pfx + 'fc2.bias'] = db2[:Ds]
            spans, c0 = [], 0
            for d, cnt in runs:
                spans.append((d, c0 * T, (c0 + cnt) * T))
                c0 += cnt
            for e in range(E):
                want_w, want_b = pfx + f'experts.{e}.weight' in trainable, pfx + f'experts.{e}.bias' in trainable
                dwe = zeros(part, H4) if want_w else None
                dbe = zeros(part) if want_b else None
                for d, r0, r1 in spans:
                    if d != e:
                        continue
                    if want_w:
                        _wgrad(dyb[r0:r1, Ds:], h[r0:r1], dwe)
                    if want_b:
                        if len(spans) == 1 and db2 is not None:
                            dbe.copy_(db2[Ds:])
                        else:
                            tmp = scratch_zeros(D)
                            ops.colsum_accumulate(dyb[r0:r1], tmp)
                            dbe.add_(tmp[Ds:])
                if want_w:
                    g[pfx + f'experts.{e}.weight'] = dwe
                if want_b:
                    g[pfx + f'experts.{e}.bias'] = dbe
            # input gradient through the effective fc2 of each run's dataset (+ GELU', + fc1's bias gradient)
            db1 = zeros(H4) if FUSE_MLP and pfx + 'fc1.bias' in trainable else None
            dpre = torch.empty(M, H4, device=dev, dtype=BF16)
            for d, r0, r1 in spans:
                wct = w['cat'][d][1]
                if FUSE_MLP:
                    ops.gemm_gelu_bwd(dyb[r0:r1], wct, pre[r0:r1], db1, out=dpre[r0:r1])
                else:
                    dpre[r0:r1] = ops.gelu_bwd(pre[r0:r1], ops.gemm(dyb[r0:r1], wct, EPI_BIAS))
            return dpre, db1

        L = len(bb.blocks)
        acts = s['acts']
        dyb, db2 = norm_bwd('backbone.last_norm', bb.last_norm, s['x_final'], dact, acts[L - 1]['s2'],
                            f'backbone.blocks.{L - 1}.mlp.fc2.bias')
        for l in range(L - 1, -1, -1):
            a, w, blk = acts[l], s['blocks'][l], bb.blocks[l]
            pfx = f'backbone.blocks.{l}.'
            # x_out = x_mid + fc2(gelu(fc1(norm2(x_mid))))            (vit.py:139); dyb = gradient of that branch
            if E:
                dpre, db1 = moe_fc2_bwd(pfx + 'mlp.', w, dyb, a, db2)
            elif FUSE_MLP:
                # fc2's input gradient with the GELU backward in its epilogue (+ fc1's bias gradient as column sums)
                linear_bwd(pfx + 'mlp.fc2', w['fc2'], dyb, a['h'], want_dx=False, dbias=db2)
                db1 = zeros(w['fc1'].w.shape[0]) if pfx + 'mlp.fc1.bias' in trainable else None
                dpre = ops.gemm_gelu_bwd(dyb, w['fc2'].wt, a['pre'], db1)
            else:
                dh = linear_bwd(pfx + 'mlp.fc2', w['fc2'], dyb, a['h'], dbias=db2)
                dpre, db1 = ops.gelu_bwd(a['pre'], dh), None
            dxn2 = linear_bwd(pfx + 'mlp.fc1', w['fc1'], dpre, a['xn2'], dbias=db1)
            dyb, dbp = norm_bwd(pfx + 'norm2', blk.norm2, a['x_mid'], dxn2, a['s1'], pfx + 'attn.proj.bias')
            # x_mid = x_in + proj(attention(qkv(norm1(x_in))))        (vit.py:138)
            dattn = linear_bwd(pfx + 'attn.proj', w['proj'], dyb, a['attn'], dbias=dbp)
            dbq = zeros(3 * D) if FUSE_MLP and pfx + 'attn.qkv.bias' in trainable else None    # summed by the kernel
            dqkv = ops.attention_bwd(a['qkv'].view(n, T, 3 * D), a['attn'].view(n, T, D), a['lse'],
                                     dattn.view(n, T, D), heads, dbias=dbq)
            dxn1 = linear_bwd(pfx + 'attn.qkv', w['qkv'], dqkv.view(M, 3 * D), a['xn1'], dbias=dbq)
            if l > 0:
                dyb, db2 = norm_bwd(pfx + 'norm1', blk.norm1, a['x_in'], dxn1, acts[l - 1]['s2'],
                                    f'backbone.blocks.{l - 1}.mlp.fc2.bias')
            else:       # the stream below block 0 is the patch embedding (no stochastic depth)
                dyb, db2 = norm_bwd(pfx + 'norm1', blk.norm1, a['x_in'], dxn1, None, 'backbone.patch_embed.proj.bias')
            acts[l] = None                                    # activations of this block are dead
            exchange()
        # ---- patch embed (vit.py:159-165) + pos embed (vit.py:320)
        linear_bwd('backbone.patch_embed.proj', s['pe'], dyb, s['patches'], want_dx=False, dbias=db2)
        if 'backbone.patch_embed.proj.weight' in g:
            g['backbone.patch_embed.proj.weight'] = g['backbone.patch_embed.proj.weight'].view(
                bb.patch_embed.proj.weight.shape)
        dpos = zeros(1, T + 1, D)                                          # [cls slot | tokens], contiguous
        dpos_tok = dpos[0, 1:].reshape(T * D)
        ops.colsum_accumulate(dx.view(n, T * D), dpos_tok)                 # sum over the crops
        ops.colsum_accumulate(dpos[0, 1:], dpos[0, 0])                     # the cls slot is added to every token
        g['backbone.pos_embed'] = dpos
        ctx.s = None
        exchange(final=True)
        for h in pending:
            h.wait()
        for seg, half in compressed:
            seg.copy_(half)
        if world > 1 and not has_avg:
            arena[:used[0]].mul_(1.0 / world)
        grads = []
        for nm, p in _param_list(model):
            gr = g.get(nm) if p.requires_grad else None
            grads.append(gr.reshape(p.shape) if gr is not None else None)
        return (None, None, *grads)


def network_heatmaps_train(model, img):
    """Differentiable forward of backbone + head for ``TopDown.forward_train``."""
    params = [p for _, p in _param_list(model)]
    return _NetworkFn.apply(model, img, *params)
