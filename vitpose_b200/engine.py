"""Host-side engine: one-time weight repack (reference state_dict -> device buffers in the layouts the
kernels want) and the per-batch launch of vpb_vitpose_forward + vpb_decode_heatmaps.

This is plumbing around the C ABI (pointers, workspace, streams); all arithmetic on the hot path is in
libvitpose_b200.so.  State-dict keys/layouts are the reference's, so checkpoints load
unchanged.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib
from ._lib import BlockFold, BlockWeights, ModelDesc, MoeRuns, Weights, check, lib, ptr, stream_ptr

BF16 = torch.bfloat16


def model_desc_from_cfg(backbone_cfg, head_cfg):
    """vpb_model_desc from the `backbone=` / `keypoint_head=` dicts of a ViTPose config."""
    img = backbone_cfg.get('img_size', 224)
    img_h, img_w = (img, img) if isinstance(img, int) else tuple(img)
    patch = backbone_cfg.get('patch_size', 16)
    if patch != 16 or backbone_cfg.get('ratio', 1) != 1:
        raise ValueError('the CUDA patch-embed kernel implements patch_size=16, ratio=1 (all ViTPose configs)')
    D = backbone_cfg.get('embed_dim', 768)
    d = ModelDesc()
    d.img_h, d.img_w = img_h, img_w
    d.embed_dim = D
    d.depth = backbone_cfg.get('depth', 12)
    d.num_heads = backbone_cfg.get('num_heads', 12)
    d.mlp_hidden = int(D * backbone_cfg.get('mlp_ratio', 4.))
    d.ln_eps = 1e-6
    d.has_last_norm = int(backbone_cfg.get('last_norm', True))
    if head_cfg is not None:
        nd = head_cfg.get('num_deconv_layers', 3)
        d.num_deconv = nd
        filters = list(head_cfg.get('num_deconv_filters', (256, 256, 256)))[:nd]
        for i in range(3):
            d.deconv_channels[i] = filters[i] if i < nd else 0
        kernels = list(head_cfg.get('num_deconv_kernels', (4, 4, 4)))[:nd]
        if any(k != 4 for k in kernels):
            raise ValueError('the CUDA deconv kernel implements kernel 4 / stride 2 / padding 1 (all ViTPose configs)')
        d.upsample = head_cfg.get('upsample', 0)
        extra = head_cfg.get('extra') or {}
        d.final_kernel = extra.get('final_conv_kernel', 1)
        d.num_keypoints = head_cfg['out_channels']
    return d


def pack_deconv_weight(w):
    """ConvTranspose2d weight [Cin, Cout, 4, 4] -> bf16 [4 phases, Cout, 4 taps * Cin] (see vitpose_b200.h)."""
    cin, cout = w.shape[:2]
    out = torch.empty(4, cout, 4 * cin, dtype=torch.float32, device=w.device)
    for py in range(2):
        for px in range(2):
            for ty in range(2):
                for tx in range(2):
                    kh = (1 if py == 0 else 2) if ty == 0 else (3 if py == 0 else 0)
                    kw = (1 if px == 0 else 2) if tx == 0 else (3 if px == 0 else 0)
                    t = ty * 2 + tx
                    out[py * 2 + px, :, t * cin:(t + 1) * cin] = w[:, :, kh, kw].t()
    return out.to(BF16).contiguous()


def pack_deconv_weight_dgrad(wp):
    """Packed deconv weight [4 phases, Cout, 4 taps * Cin] -> [Cin, 16 * Cout] with column (ph*4 + t)*Cout + co:
    the B operand of the transposed convolution's input gradient as one GEMM over the 16 (phase, tap) shifted
    copies of dY (vpb_deconv_gather_dy)."""
    _, cout, c4 = wp.shape
    cin = c4 // 4
    return wp.reshape(4, cout, 4, cin).permute(3, 0, 2, 1).reshape(cin, 16 * cout).contiguous()


def unpack_deconv_weight(wp, dtype=torch.float32, out=None):
    """Inverse of pack_deconv_weight: [4, Cout, 4*Cin] -> ConvTranspose2d layout [Cin, Cout, 4, 4]."""
    _, cout, c4 = wp.shape
    cin = c4 // 4
    w = torch.empty(cin, cout, 4, 4, dtype=dtype, device=wp.device) if out is None else out
    for py in range(2):
        for px in range(2):
            for ty in range(2):
                for tx in range(2):
                    kh = (1 if py == 0 else 2) if ty == 0 else (3 if py == 0 else 0)
                    kw = (1 if px == 0 else 2) if tx == 0 else (3 if px == 0 else 0)
                    t = ty * 2 + tx
                    w[:, :, kh, kw] = wp[py * 2 + px, :, t * cin:(t + 1) * cin].t().to(dtype)
    return w


def fold_bn(g, b, mean, var, eps=1e-5):
    scale = g / torch.sqrt(var + eps)
    return scale.float().contiguous(), (b - mean * scale).float().contiguous()


class PackedWeights:
    """Device buffers + the ctypes vpb_weights struct that points at them."""

    def __init__(self, state_dict, desc, device, backbone_prefix='backbone.', head_prefix='keypoint_head.',
                 share=None, private=()):
        """``share``: another PackedWeights built from the same module at the same weights version; every repacked
        tensor is taken from it instead of being converted again, except those made from state-dict keys that contain
        one of the ``private`` substrings (ViTPose+: only mlp.fc2 differs between the per-dataset networks)."""
        sd = state_dict
        dev = device
        D, L = desc.embed_dim, desc.depth
        self.keep = []          # owns every tensor the struct points to
        self.by_id = {}         # conversion order -> device tensor (what a later `share=` reuses)
        self.shared_bytes = 0
        pending = [None]        # state-dict key of the tensor being converted (set by g / h below)

        def convert(t, make):
            idx, key = len(self.by_id), pending[0]
            pending[0] = None
            if share is not None and key is not None and not any(p in key for p in private) and idx in share.by_id:
                out = share.by_id[idx]
                self.shared_bytes += out.numel() * out.element_size()
            else:
                out = make(t)
            self.by_id[idx] = out
            self.keep.append(out)
            return out

        def dev_f32(t):
            return convert(t, lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous())

        def dev_bf16(t):
            return convert(t, lambda t: t.detach().to(device=dev, dtype=torch.float32).to(BF16).contiguous())

        def g(k):
            pending[0] = backbone_prefix + k
            return sd[backbone_prefix + k]

        w = Weights()
        self.patch_w = dev_bf16(g('patch_embed.proj.weight').reshape(D, -1))
        w.patch_w = ptr(self.patch_w)
        w.patch_b = ptr(dev_f32(g('patch_embed.proj.bias')))
        pos = g('pos_embed').detach().float()
        w.pos = ptr(dev_f32(pos[0, 1:] + pos[0, :1]))
        self.blocks = (BlockWeights * L)()
        for i in range(L):
            b, bw = f'blocks.{i}.', self.blocks[i]
            bw.ln1_g, bw.ln1_b = ptr(dev_f32(g(b + 'norm1.weight'))), ptr(dev_f32(g(b + 'norm1.bias')))
            bw.qkv_w = ptr(dev_bf16(g(b + 'attn.qkv.weight')))
            qkv_b = sd.get(backbone_prefix + b + 'attn.qkv.bias')
            bw.qkv_b = ptr(dev_f32(qkv_b if qkv_b is not None else torch.zeros(3 * D)))
            bw.proj_w, bw.proj_b = ptr(dev_bf16(g(b + 'attn.proj.weight'))), ptr(dev_f32(g(b + 'attn.proj.bias')))
            bw.ln2_g, bw.ln2_b = ptr(dev_f32(g(b + 'norm2.weight'))), ptr(dev_f32(g(b + 'norm2.bias')))
            bw.fc1_w, bw.fc1_b = ptr(dev_bf16(g(b + 'mlp.fc1.weight'))), ptr(dev_f32(g(b + 'mlp.fc1.bias')))
            bw.fc2_w, bw.fc2_b = ptr(dev_bf16(g(b + 'mlp.fc2.weight'))), ptr(dev_f32(g(b + 'mlp.fc2.bias')))
        w.blocks = ctypes.cast(self.blocks, ctypes.POINTER(BlockWeights))
        # folded LayerNorm (include/vitpose_b200.h: vpb_block_fold): qkv / fc1 operands that carry norm1 / norm2
        # (opt-in, VPB_LN_FOLD=1: measured no faster end to end, see profiles/r02_summary.md §8)
        from .ops import fold_layernorm_linear
        use_fold = os.environ.get('VPB_LN_FOLD', '0') not in ('', '0')
        self.fold = (BlockFold * L)()
        for i in range(L if use_fold else 0):
            b, fw = f'blocks.{i}.', self.fold[i]
            for lin, norm, n_out in (('attn.qkv', 'norm1', 3 * D), ('mlp.fc1', 'norm2', desc.mlp_hidden)):
                bias = sd.get(backbone_prefix + b + lin + '.bias')
                wf, s, c = fold_layernorm_linear(
                    g(b + lin + '.weight').detach().to(device=dev, dtype=torch.float32),
                    dev_f32(bias if bias is not None else torch.zeros(n_out)),
                    dev_f32(g(b + norm + '.weight')), dev_f32(g(b + norm + '.bias')))
                self.keep += [wf, s, c]
                short = lin.split('.')[1]
                setattr(fw, short + '_wf', ptr(wf))
                setattr(fw, short + '_s', ptr(s))
                setattr(fw, short + '_c', ptr(c))
        if use_fold:
            w.fold = ctypes.cast(self.fold, ctypes.POINTER(BlockFold))
        if desc.has_last_norm:
            w.last_g, w.last_b = ptr(dev_f32(g('last_norm.weight'))), ptr(dev_f32(g('last_norm.bias')))
        if desc.num_keypoints > 0 and (head_prefix + 'final_layer.weight') in sd:
            h = lambda k: sd[head_prefix + k]
            for i in range(desc.num_deconv):
                wt = h(f'deconv_layers.{3 * i}.weight').detach().float().to(dev)
                w.deconv_w[i] = ptr(self._own(pack_deconv_weight(wt)))
                bn = f'deconv_layers.{3 * i + 1}.'
                s, t = fold_bn(h(bn + 'weight').float().to(dev), h(bn + 'bias').float().to(dev),
                               h(bn + 'running_mean').float().to(dev), h(bn + 'running_var').float().to(dev))
                w.deconv_scale[i], w.deconv_shift[i] = ptr(self._own(s)), ptr(self._own(t))
            fw = h('final_layer.weight').detach().float()
            K, cin, kh, kw = fw.shape
            if kh == 1:
                fw = fw.reshape(K, cin)
            else:   # [K, Cin, 3, 3] -> [K, (ky*3+kx)*Cin + ci]
                fw = fw.permute(0, 2, 3, 1).reshape(K, kh * kw * cin)
            w.final_w = ptr(dev_bf16(fw))
            w.final_b = ptr(dev_f32(h('final_layer.bias')))
        self.struct = w

    def _own(self, t):
        self.keep.append(t)
        return t


DECODE_MODES = {'none': _lib.DECODE_NONE, 'default': _lib.DECODE_DEFAULT, 'unbiased': _lib.DECODE_UNBIASED,
                'udp_dark': _lib.DECODE_UDP_DARK}


class VitPoseEngine:
    """Runs backbone + head + decode for batches of crops on one GPU."""

    def __init__(self, backbone_cfg, head_cfg, state_dict, device='cuda', max_batch=64, share_weights_from=None,
                 private_keys=()):
        _lib.require_cuda()
        lib()
        self.device = torch.device(device)
        self.desc = model_desc_from_cfg(backbone_cfg, head_cfg)
        self.weights = PackedWeights(state_dict, self.desc, self.device,
                                     share=None if share_weights_from is None else share_weights_from.weights,
                                     private=private_keys)
        self._ws = None
        self._ws_images = 0
        self.max_batch = max_batch

    # ---- workspace ------------------------------------------------------------------------------
    # ---- ViTPose+ (vpb_moe_runs): crops sorted by dataset, one mlp.fc2 launch per run ------------------
    def set_experts(self, experts):
        """experts: {dataset_idx: ([bf16 fc2 weight per block], [fp32 fc2 bias per block])} device tensors (kept here)."""
        self.experts = experts

    def set_moe_runs(self, runs):
        """runs: [(dataset_idx, number of crops), ...] covering, in order, the crops of the NEXT forward calls (None:
        back to the packed weights for every crop)."""
        self._moe_runs = runs

    def _moe_struct(self, n, flip):
        """ctypes vpb_moe_runs for a batch of n crops (+ their flipped copies) and the arrays it points to."""
        runs = getattr(self, '_moe_runs', None)
        if not runs:
            return None, None
        assert sum(c for _, c in runs) == n, 'the runs must cover the batch'
        passes = 2 if flip else 1
        R, depth = len(runs) * passes, self.desc.depth
        if R > _lib.MOE_MAX_RUNS:
            raise ValueError(f'{R} runs of datasets in one batch (max {_lib.MOE_MAX_RUNS}): sort the crops by dataset_idx')
        begin = (ctypes.c_int32 * (R + 1))()
        wp = (ctypes.c_void_p * (R * depth))()
        bp = (ctypes.c_void_p * (R * depth))()
        pos = 0
        for r, (d, c) in enumerate(list(runs) * passes):
            begin[r] = pos
            pos += c
            ws_, bs_ = self.experts[int(d)]
            for l in range(depth):
                wp[r * depth + l] = ws_[l].data_ptr()
                bp[r * depth + l] = bs_[l].data_ptr()
        begin[R] = pos
        mr = MoeRuns(R, ctypes.cast(begin, ctypes.POINTER(ctypes.c_int32)), ctypes.cast(wp, ctypes.POINTER(ctypes.c_void_p)),
                     ctypes.cast(bp, ctypes.POINTER(ctypes.c_void_p)))
        return mr, (begin, wp, bp)

    def _forward_call(self, img, n, flip, ws_ptr, ws_bytes, out_main, out_flip, feat):
        mr, keep = self._moe_struct(n, flip)
        w = self.weights.struct
        w.moe = ctypes.pointer(mr) if mr is not None else None
        try:
            check(lib().vpb_vitpose_forward(ctypes.byref(self.desc), ctypes.byref(w), ptr(img), n, int(flip), ws_ptr,
                                            ws_bytes, ptr(out_main), ptr(out_flip), ptr(feat), stream_ptr()),
                  'vpb_vitpose_forward')
        finally:
            w.moe = None
        del keep

    def _workspace(self, images):
        if self._ws is None or images > self._ws_images:
            nbytes = lib().vpb_workspace_bytes(ctypes.byref(self.desc), images)
            self._ws = torch.empty(nbytes + 1024, device=self.device, dtype=torch.uint8)
            self._ws_images = images
        base = self._ws.data_ptr()
        aligned = (base + 1023) // 1024 * 1024
        return aligned, self._ws.numel() - (aligned - base)

    @property
    def heatmap_size(self):
        """(H, W) of the maps vpb_vitpose_forward writes: the 16x16 token grid times 2 per transposed convolution
        (classic decoder, simple_head.py:306-337) or times ``upsample`` (simple decoder, simple_head.py:269-287)."""
        hp, wp = self.tokens_hw
        d = self.desc
        f = (1 << d.num_deconv) if d.num_deconv > 0 else (d.upsample if d.upsample > 0 else 1)
        return hp * f, wp * f

    @property
    def tokens_hw(self):
        return self.desc.img_h // 16, self.desc.img_w // 16

    # ---- network -----------------------------------------------------------------------------------
    def _check_crops(self, img):
        """The kernels index the crop buffer from desc.img_h / img_w: a mis-sized batch must be an error, not a read
        past the buffer."""
        if img.dtype != torch.float32 or not img.is_cuda:
            raise _lib.VitposeLibError('img must be a float32 CUDA tensor')
        if img.dim() != 4 or tuple(img.shape[1:]) != (3, self.desc.img_h, self.desc.img_w):
            raise ValueError(f'expected crops of shape [n,3,{self.desc.img_h},{self.desc.img_w}], got {tuple(img.shape)}')

    def forward_into(self, img, flip, out_main, out_flip):
        """Like forward_heatmaps, but writes the main-pass maps into ``out_main`` [n,K,h,w] and the raw
        flipped-pass maps into ``out_flip`` [n,K,h,w] (slices of larger, contiguous batch buffers)."""
        self._check_crops(img)
        img = img.contiguous()
        n = img.shape[0]
        H4, W4 = self.heatmap_size
        want = (n, self.desc.num_keypoints, H4, W4)
        for t in (out_main, out_flip) if flip else (out_main,):
            if tuple(t.shape) != want or t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError(f'heatmap buffer must be contiguous float32 {want}, got {tuple(t.shape)} {t.dtype}')
        images = 2 * n if flip else n
        ws_ptr, ws_bytes = self._workspace(images)
        self._forward_call(img, n, flip, ws_ptr, ws_bytes, out_main, out_flip if flip else None, None)

    def forward_heatmaps(self, img, flip=False, want_features=False, want_heatmaps=True):
        """img fp32 CUDA [n,3,H,W] -> raw heatmaps fp32 [(2n|n), K, H/4, W/4] (+ bf16 token features)."""
        self._check_crops(img)
        img = img.contiguous()
        n = img.shape[0]
        images = 2 * n if flip else n
        ws_ptr, ws_bytes = self._workspace(images)
        H4, W4 = self.heatmap_size
        hm = (torch.empty(images, self.desc.num_keypoints, H4, W4, device=self.device, dtype=torch.float32)
              if want_heatmaps else None)
        hp, wp = self.tokens_hw
        feat = (torch.empty(images, hp * wp, self.desc.embed_dim, device=self.device, dtype=BF16)
                if want_features else None)
        self._forward_call(img, n, flip, ws_ptr, ws_bytes, hm, None, feat)
        return hm, feat

    # ---- decode -------------------------------------------------------------------------------------
    def decode(self, hm, n, flip, flip_index, shift_heatmap, mode, kernel, use_udp, center, scale,
               want_merged=False):
        from . import ops
        hm_main = hm[:n]
        hm_flip = hm[n:2 * n] if flip else None
        return ops.decode(hm_main, hm_flip, flip_index if flip else None, shift_heatmap, mode, kernel, use_udp,
                          center, scale, want_merged=want_merged)


def decode_mode_from_cfg(test_cfg):
    """The branch keypoints_from_heatmaps takes for a test_cfg (top_down_eval.py:562-612)."""
    use_udp = bool(test_cfg.get('use_udp', False))
    post = test_cfg.get('post_process', 'default')
    unbiased = bool(test_cfg.get('unbiased_decoding', False))
    return resolve_decode_mode(post, unbiased, use_udp)


def resolve_decode_mode(post_process, unbiased, use_udp):
    if post_process is True:
        post_process = 'unbiased' if unbiased else 'default'
    elif post_process is False:
        post_process = None
    elif post_process == 'default' and unbiased:
        post_process = 'unbiased'
    if use_udp:
        return _lib.DECODE_UDP_DARK
    if post_process == 'unbiased':
        return _lib.DECODE_UNBIASED
    if post_process is None:
        return _lib.DECODE_NONE
    return _lib.DECODE_DEFAULT
