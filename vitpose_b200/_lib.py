"""ctypes binding of include/vitpose_b200.h.  There is NO CPU fallback: every wrapper raises when the
CUDA library is missing or when a call fails."""
import ctypes
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
# VPB_LIB=<path>: an alternative build of the same library (same-box A/B of compile-time variants)
LIB_PATH = os.environ.get('VPB_LIB') or os.path.join(HERE, 'libvitpose_b200.so')

c_void_p, c_int, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t

EPI_BIAS_BF16, EPI_GELU_BF16, EPI_RESID_F32, EPI_POS_F32, EPI_NCHW_F32 = range(5)
EPI_ACCUM_F32 = 10
ABI_VERSION = 5          # VPB_ABI_VERSION of include/vitpose_b200.h this module was written against
DECODE_NONE, DECODE_DEFAULT, DECODE_UNBIASED, DECODE_UDP_DARK = range(4)


class ModelDesc(ctypes.Structure):
    _fields_ = [('img_h', ctypes.c_int32), ('img_w', ctypes.c_int32), ('embed_dim', ctypes.c_int32),
                ('depth', ctypes.c_int32), ('num_heads', ctypes.c_int32), ('mlp_hidden', ctypes.c_int32),
                ('ln_eps', ctypes.c_float), ('has_last_norm', ctypes.c_int32), ('num_deconv', ctypes.c_int32),
                ('deconv_channels', ctypes.c_int32 * 3), ('upsample', ctypes.c_int32),
                ('final_kernel', ctypes.c_int32), ('num_keypoints', ctypes.c_int32)]


class BlockWeights(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ('ln1_g', 'ln1_b', 'qkv_w', 'qkv_b', 'proj_w', 'proj_b',
                                        'ln2_g', 'ln2_b', 'fc1_w', 'fc1_b', 'fc2_w', 'fc2_b')]


class BlockFold(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ('qkv_wf', 'qkv_s', 'qkv_c', 'fc1_wf', 'fc1_s', 'fc1_c')]


class MoeRuns(ctypes.Structure):
    """vpb_moe_runs: runs of images (sorted by dataset) with their own mlp.fc2 weights (ViTPose+)."""
    _fields_ = [('num_runs', ctypes.c_int32), ('image_begin', ctypes.POINTER(ctypes.c_int32)),
                ('fc2_w', ctypes.POINTER(c_void_p)), ('fc2_b', ctypes.POINTER(c_void_p))]


MOE_MAX_RUNS = 64


class Weights(ctypes.Structure):
    _fields_ = [('patch_w', c_void_p), ('patch_b', c_void_p), ('pos', c_void_p),
                ('blocks', ctypes.POINTER(BlockWeights)), ('last_g', c_void_p), ('last_b', c_void_p),
                ('deconv_w', c_void_p * 3), ('deconv_scale', c_void_p * 3), ('deconv_shift', c_void_p * 3),
                ('final_w', c_void_p), ('final_b', c_void_p), ('fold', ctypes.POINTER(BlockFold)),
                ('moe', ctypes.POINTER(MoeRuns))]


class VitposeLibError(RuntimeError):
    pass


_lib = None

_SIGS = {
    'vpb_abi_version': (c_int, []),
    'vpb_last_error': (ctypes.c_char_p, []),
    'vpb_profile_enable': (None, [c_int]),
    'vpb_profile_count': (c_int, []),
    'vpb_profile_get': (c_int, [c_int, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(c_float)]),
    'vpb_launch_count': (ctypes.c_longlong, []),
    'vpb_workspace_bytes': (c_size_t, [ctypes.POINTER(ModelDesc), c_int]),
    'vpb_vitpose_forward': (c_int, [ctypes.POINTER(ModelDesc), ctypes.POINTER(Weights), c_void_p, c_int, c_int,
                                    c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
    'vpb_decode_heatmaps': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                    c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p]),
    'vpb_flip_back': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'vpb_transform_preds': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    c_void_p]),
    'vpb_transpose_bf16': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    'vpb_cast_f32_bf16': (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_void_p, c_int, c_int, c_void_p]),
    'vpb_colsum_accumulate': (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    'vpb_gelu_fwd_bf16': (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_void_p]),
    'vpb_gelu_bwd_bf16': (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_void_p]),
    'vpb_layernorm_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float,
                                  c_void_p]),
    'vpb_attention_lse': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    'vpb_attention_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                  c_float, c_void_p]),
    'vpb_attention_bwd_bias': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                       c_int, c_float, c_void_p]),
    'vpb_deconv4x4s2_raw': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p]),
    'vpb_bn_train_stats': (c_int, [c_void_p, ctypes.c_longlong, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    'vpb_bn_relu_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int,
                                c_void_p]),
    'vpb_bn_relu_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p]),
    'vpb_bn_relu_bwd_eval': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, ctypes.c_longlong, c_int, c_void_p]),
    'vpb_nchw_f32_to_rows_bf16': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'vpb_deconv_gather_x': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'vpb_deconv_gather_dy': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'vpb_deconv_phase_dy': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'vpb_deconv_pack_weight': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'vpb_deconv_unpack_wgrad': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'vpb_warp_affine_normalize': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                          ctypes.POINTER(c_float), ctypes.POINTER(c_float), c_void_p, c_void_p]),
    'vpb_joints_mse_loss': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                    c_void_p]),
    'vpb_grad_sq_norm_accumulate': (c_int, [c_void_p, ctypes.c_longlong, c_void_p, c_void_p]),
    'vpb_adamw_step': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_float, c_float, c_float,
                               c_float, c_float, c_int, c_void_p, c_float, c_void_p]),
    'vpb_adamw_multi': (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_float, c_void_p, c_float,
                                c_void_p]),
    'vpb_cast_transpose_multi': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    'vpb_pose_pck_accuracy': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_float, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    'vpb_oks_nms': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, ctypes.c_double,
                            c_int, ctypes.c_double, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    'vpb_gemm_bf16': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                              c_int, c_int, c_void_p]),
    'vpb_gemm_bf16_atb_accum': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    'vpb_gemm_bf16_atb_accum_ld': (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                           c_void_p]),
    'vpb_gemm_layernorm_scratch_bytes': (c_size_t, [c_int, c_int]),
    'vpb_gemm_bf16_layernorm': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                        c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_size_t, c_void_p, c_int,
                                        c_void_p]),
    'vpb_relu_bf16': (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_void_p]),
    'vpb_relu_bwd_bf16': (c_int, [c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_void_p]),
    'vpb_simple_head_gather': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'vpb_simple_head_gather_bwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'vpb_gemm_bf16_gelu_save': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                        c_void_p]),
    'vpb_gemm_bf16_gelu_bwd': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                       c_void_p]),
    'vpb_cast_f32_bf16_colsum': (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    'vpb_fold_layernorm_linear': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                          c_void_p, c_void_p]),
    'vpb_gemm_stats_layout': (c_int, [c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    'vpb_gemm_stats_bytes': (c_size_t, [c_int, c_int]),
    'vpb_gemm_bf16_resid_stats': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                          c_int, c_void_p, c_void_p, c_size_t, c_void_p, c_int, c_void_p]),
    'vpb_gemm_bf16_lnfold': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                     c_int, c_int, c_float, c_void_p, c_int, c_void_p]),
    'vpb_gemm_layernorm_scratch_init': (c_int, [c_void_p, c_size_t, c_int, c_int, c_void_p]),
    'vpb_gemm_bf16_layernorm_seq': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                            c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_size_t,
                                            ctypes.c_uint, c_void_p, c_int, c_void_p]),
    'vpb_layernorm_bf16': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]),
    'vpb_im2col_patch16': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    'vpb_attention': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    'vpb_deconv4x4s2_bn_relu': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                        c_int, c_int, c_void_p]),
    'vpb_conv3x3_nchw': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p]),
    'vpb_relu_upsample_nhwc': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    'vpb_tokens_to_nchw_f32': (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGS)


def lib():
    """Loads libvitpose_b200.so (in-tree).  Raises if it has not been built — no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VitposeLibError(
            f'{LIB_PATH} is missing: build it with `python -m vitpose_b200.build` '
            '(or __graft_entry__.build()). There is no CPU/eager fallback for this path.')
    L = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    if L.vpb_abi_version() != ABI_VERSION:
        raise VitposeLibError('ABI version mismatch between _lib.py and libvitpose_b200.so')
    _lib = L
    return L


ABI_CALLS = [0]     # C-ABI calls that returned through check() (each launches at least one kernel)


def check(code, what):
    ABI_CALLS[0] += 1
    if code != 0:
        msg = lib().vpb_last_error().decode('utf-8', 'replace')
        raise VitposeLibError(f'{what} failed ({code}): {msg}')


def profile_records():
    """[(tag, ms)] of the launches recorded since vpb_profile_enable(1)."""
    L = lib()
    out = []
    tag, ms = ctypes.c_char_p(), c_float()
    for i in range(L.vpb_profile_count()):
        check(L.vpb_profile_get(i, ctypes.byref(tag), ctypes.byref(ms)), 'vpb_profile_get')
        out.append((tag.value.decode(), ms.value))
    return out


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or None)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise VitposeLibError('the vitpose_b200 kernels need CUDA tensors (no CPU fallback)')
    if not t.is_contiguous():
        raise VitposeLibError('tensor must be contiguous')
    return t.data_ptr()


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def stream_ptr():
    """cudaStream_t of torch's current stream on the current device (every kernel of the library is launched on it).
    The raw accessor avoids building a torch.cuda.Stream object per launch (~5 us each, 270 launches per training step)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def require_cuda():
    if not torch.cuda.is_available():
        raise VitposeLibError('vitpose_b200 runs only on a CUDA device (sm_100a); no CPU fallback exists')
