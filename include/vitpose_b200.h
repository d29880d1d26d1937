/* vitpose_b200 — C ABI of the B200-native ViTPose top-down inference hot path.
 *
 * The reference (MiraPurkrabek/ViTPose, a fork of mmpose 0.24) is 100 % Python and has no FFI: its
 * extension points are the mmcv registry classes and a few free functions.  This header is the
 * boundary a maintainer binds instead (ctypes stub in INTEGRATION.md); every entry point names the
 * reference code it replaces.  Plain pointers and sizes only — no torch types.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless stated otherwise; the caller owns all memory,
 *    including outputs and the workspace; the library never allocates device memory;
 *  - calls are asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream);
 *  - return value 0 = ok; non-zero = error, message via vpb_last_error() (thread-local);
 *  - "bf16" buffers are packed __nv_bfloat16; weights are [out_features, in_features] row-major,
 *    i.e. exactly torch's nn.Linear layout cast to bf16.
 */
#ifndef VITPOSE_B200_H_
#define VITPOSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPB_ABI_VERSION 5

int vpb_abi_version(void);
const char* vpb_last_error(void);

/* Optional profiler: while enabled, every kernel launched by vpb_vitpose_forward / vpb_decode_heatmaps is
 * bracketed by a CUDA-event pair on the launching stream. vpb_profile_get synchronises on record i and returns
 * its tag (static string) and duration in milliseconds. vpb_launch_count = kernels launched by those two
 * entry points since the library was loaded. */
void vpb_profile_enable(int on);
int vpb_profile_count(void);
int vpb_profile_get(int i, const char** tag, float* ms);
long long vpb_launch_count(void);

/* ---- model description ---------------------------------------------------------------------
 * Mirrors the `model = dict(...)` block of a ViTPose config
 * (configs/body/2d_kpt_sview_rgb_img/topdown_heatmap/coco/ViTPose_base_coco_256x192.py:52-84). */
typedef struct vpb_model_desc {
  int32_t img_h, img_w;        /* 256, 192 */
  int32_t embed_dim;           /* D */
  int32_t depth;               /* number of Blocks */
  int32_t num_heads;
  int32_t mlp_hidden;          /* int(D * mlp_ratio) */
  float ln_eps;                /* 1e-6, vit.py:212 */
  int32_t has_last_norm;       /* ViT(last_norm=True) */
  int32_t num_deconv;          /* classic decoder: 2; simple decoder: 0 */
  int32_t deconv_channels[3];  /* num_deconv_filters */
  int32_t upsample;            /* simple decoder: 4; classic: 0 */
  int32_t final_kernel;        /* 1 (classic) or 3 (simple) */
  int32_t num_keypoints;       /* out_channels */
} vpb_model_desc;

/* One transformer Block (vit.py:117-140). *_w are bf16, everything else fp32. */
typedef struct vpb_block_weights {
  const float* ln1_g; const float* ln1_b;
  const void* qkv_w;  const float* qkv_b;    /* [3D, D], [3D]  (attn.qkv)  */
  const void* proj_w; const float* proj_b;   /* [D, D], [D]    (attn.proj) */
  const float* ln2_g; const float* ln2_b;
  const void* fc1_w;  const float* fc1_b;    /* [4D, D], [4D]  (mlp.fc1)   */
  const void* fc2_w;  const float* fc2_b;    /* [D, 4D], [D]   (mlp.fc2)   */
} vpb_block_weights;

/* Optional per-Block operands of the "folded LayerNorm" forward (vpb_fold_layernorm_linear builds them): the two
 * Linear layers that follow a LayerNorm (attn.qkv after norm1, mlp.fc1 after norm2, vit.py:138-139) then read the plain
 * bf16 residual rows and apply the normalisation in their own epilogue,
 *   LN(x) W^T + b = rstd * (x (gamma o W)^T - mean * s) + c,   s_n = sum_k (gamma o W)[n,k],  c_n = b_n + sum_k beta_k W[n,k],
 * so that the GEMM in front (attn.proj / mlp.fc2 + residual) finishes each tile in one pass and no longer waits for
 * its sibling tiles' row statistics. *_wf bf16 [N, D], *_s / *_c fp32 [N]. */
typedef struct vpb_block_fold {
  const void* qkv_wf; const float* qkv_s; const float* qkv_c;
  const void* fc1_wf; const float* fc1_s; const float* fc1_c;
} vpb_block_fold;

/* ViTPose+ (ViTMoE, vit_moe.py:78-115): the per-dataset experts own the last part_features output columns of mlp.fc2,
 * every other weight is shared. A batch whose crops are sorted by dataset runs in ONE pass: all kernels on the whole
 * batch except mlp.fc2 (+ residual + the following LayerNorm), which is launched once per run of images with that
 * run's weight. image_begin counts images of the batch the forward call sees (2n of them with flip: the n originals, then
 * the n flipped copies — so a dataset that owns crops [a, b) owns the runs [a, b) and [n + a, n + b)). */
#define VPB_MOE_MAX_RUNS 64
typedef struct vpb_moe_runs {
  int32_t num_runs;                 /* 1 .. VPB_MOE_MAX_RUNS */
  const int32_t* image_begin;       /* HOST [num_runs + 1], increasing, image_begin[0] = 0, last = number of images */
  const void* const* fc2_w;         /* HOST [num_runs * depth]: bf16 [D, hidden] of run r, block l at [r * depth + l] */
  const float* const* fc2_b;        /* HOST [num_runs * depth]: fp32 [D] */
} vpb_moe_runs;

/* Repacked weights (built once by the host side, vitpose_b200/engine.py: PackedWeights):
 *  patch_w  bf16 [D, 768]      = patch_embed.proj.weight.reshape(D, 3*16*16)
 *  pos      fp32 [T, D]        = pos_embed[0, 1:] + pos_embed[0, :1]           (vit.py:320)
 *  deconv_w bf16 [4][Cout][4*Cin]: per output parity (py,px) the 2x2 taps of the k4/s2/p1 transposed conv,
 *           tap t=(ty,tx): ty=0 -> kh = (py==0 ? 1 : 2), ty=1 -> kh = (py==0 ? 3 : 0) (same for x);
 *           column = t*Cin + ci, value = W[ci, co, kh, kw]
 *  deconv_scale/shift fp32 [Cout]: BatchNorm2d(eval) folded: scale = g/sqrt(var+eps), shift = b - mean*scale
 *  final_w  bf16 [K, Cin] (1x1) or [K][9*Cin] (3x3, column = (ky*3+kx)*Cin + ci);  final_b fp32 [K] */
typedef struct vpb_weights {
  const void* patch_w; const float* patch_b; const float* pos;
  const vpb_block_weights* blocks;           /* HOST pointer to `depth` structs */
  const float* last_g; const float* last_b;
  const void* deconv_w[3]; const float* deconv_scale[3]; const float* deconv_shift[3];
  const void* final_w; const float* final_b;
  const vpb_block_fold* fold;                /* HOST pointer to `depth` structs, or NULL: LayerNorm in the producing GEMM */
  const vpb_moe_runs* moe;                   /* HOST pointer or NULL: blocks[l].fc2_* for every image */
} vpb_weights;

/* Bytes of scratch `vpb_vitpose_forward` needs for `images` crops (count the flipped copies too). */
size_t vpb_workspace_bytes(const vpb_model_desc* desc, int images);

/* backbone + head for n crops (+ their horizontal flips when flip != 0):
 *   img        fp32 [n, 3, img_h, img_w]  (NCHW, as the reference's `img` argument)
 *   heatmaps   fp32 [(flip ? 2n : n), K, img_h/4, img_w/4]; rows [n, 2n) are the RAW outputs of the flipped
 *              pass (not yet flipped back) — feed both halves to vpb_decode_heatmaps
 *   heatmaps_flipped  optional fp32 [n, K, img_h/4, img_w/4]: when non-NULL (and flip != 0) the flipped-pass maps
 *              are written here instead and `heatmaps` only needs n maps (lets a caller that pipelines chunks of
 *              a batch assemble contiguous [N,...] main / flipped tensors for one decode call)
 *   features   optional bf16 [(flip ? 2n : n), T, D] token-major backbone output (NULL to skip)
 * Replaces ViT.forward (vit.py:313-337), TopdownHeatmapSimpleHead.forward (simple_head.py:197-202) and the
 * second, flipped pass of TopDown.forward_test (top_down.py:179-186). */
int vpb_vitpose_forward(const vpb_model_desc* desc, const vpb_weights* w, const float* img, int n, int flip,
                        void* workspace, size_t workspace_bytes, float* heatmaps, float* heatmaps_flipped,
                        void* features, void* stream);

/* decode modes = the branches of keypoints_from_heatmaps (top_down_eval.py:562-612) */
enum { VPB_DECODE_NONE = 0, VPB_DECODE_DEFAULT = 1, VPB_DECODE_UNBIASED = 2, VPB_DECODE_UDP_DARK = 3 };

/* flip_back + shift + average + argmax + refine + transform_preds, fused.
 *   hm            fp32 [N,K,H,W]
 *   hm_flipped    fp32 [N,K,H,W] raw heatmaps of the flipped pass, or NULL (no flip test)
 *   flip_index    int32 [K] channel permutation of flip_back (post_transforms.py:138-141), or NULL
 *   shift_heatmap test_cfg['shift_heatmap'] (simple_head.py:223-224)
 *   mode/kernel   VPB_DECODE_*, test_cfg['modulate_kernel']
 *   use_udp       selects the (W-1)/(H-1) scaling of transform_preds (post_transforms.py:183-188)
 *   apply_transform 0 -> preds stay in heatmap pixels (host applies transform_preds itself)
 *   center, scale fp32 [N,2]
 *   preds         fp32 [N,K,2];  maxvals fp32 [N,K,1]
 *   merged_out    optional fp32 [N,K,H,W]: the averaged heatmap (`output_heatmap` of forward_test)
 *   argmax_out    optional int32 [N,K]: flat argmax index (bit-exact vs np.argmax)
 * Replaces TopdownHeatmapSimpleHead.inference_model's host half (simple_head.py:217-226),
 * top_down.py:187-188 and keypoints_from_heatmaps (top_down_eval.py:474-622, GaussianHeatmap branches). */
int vpb_decode_heatmaps(const float* hm, const float* hm_flipped, const int32_t* flip_index, int shift_heatmap,
                        int N, int K, int H, int W, int mode, int kernel, int use_udp, int apply_transform,
                        const float* center, const float* scale, float* preds, float* maxvals, float* merged_out,
                        int32_t* argmax_out, void* stream);

/* flip_back (post_transforms.py:110-147, GaussianHeatmap) + optional shift_heatmap (simple_head.py:223-224):
 * out[n,k,y,x] = in[n, flip_index[k], y, W-1-xs], xs = shift ? max(x-1,0) : x.  in/out fp32 [N,K,H,W]. */
int vpb_flip_back(const float* in, const int32_t* flip_index, float* out, int N, int K, int H, int W, int shift,
                  void* stream);
/* transform_preds (post_transforms.py:150-194) in float32: coords/out fp32 [N,K,2], center/scale fp32 [N,2],
 * heatmap size (W, H). */
int vpb_transform_preds(const float* coords, const float* center, const float* scale, float* out, int N, int K,
                        int W, int H, int use_udp, void* stream);

/* ---- individual operators (used by the per-kernel parity tests and by vpb_vitpose_forward) ---- */
enum {
  VPB_EPI_BIAS_BF16 = 0,  /* out bf16 [M,ldo]  = A.B^T + bias                         (attn.qkv)            */
  VPB_EPI_GELU_BF16 = 1,  /* out bf16 [M,ldo]  = gelu_erf(A.B^T + bias)               (mlp.fc1 + nn.GELU)   */
  VPB_EPI_RESID_F32 = 2,  /* out fp32 [M,ldo]  = aux[M,ldo] + A.B^T + bias            (attn.proj / mlp.fc2 + residual) */
  VPB_EPI_POS_F32 = 3,    /* out fp32 [M,ldo]  = A.B^T + bias + aux[row % period, N]  (patch embed + pos embed) */
  VPB_EPI_NCHW_F32 = 4,   /* out fp32 [M/period, N, period] = A.B^T + bias            (final 1x1 conv)      */
  VPB_EPI_ACCUM_F32 = 10  /* out fp32 [M,ldo] += A.B^T, K split over CTAs (atomic adds; bias must be NULL): weight
                             gradients dW += dY^T X, where M x N is small and K is every token of the batch    */
};
/* C = A[M,K] (bf16, row-major) x B[N,K]^T (bf16, row-major) on tcgen05 tensor cores. max_ctas <= 0: one per SM. */
int vpb_gemm_bf16(const void* A, const void* B, int M, int N, int K, int epilogue, const float* bias, void* out,
                  int ldo, const float* aux, int period, int max_ctas, void* stream);
/* out fp32 [M,ldo] += At^T . Bt for ROW-MAJOR At bf16 [K,M] and Bt bf16 [K,N] (M, N multiples of 8): a weight
 * gradient dW[N_out,K_in] += dY[tokens,N_out]^T . X[tokens,K_in] straight from dY and X — both tiles are consumed as
 * MN-major tensor-core operands (no transposed copies), the contraction over the tokens is split over CTAs. */
int vpb_gemm_bf16_atb_accum(const void* At, const void* Bt, int M, int N, int K, float* out, int ldo, void* stream);
/* The same with explicit row pitches lda >= M, ldb >= N (elements, multiples of 8): At / Bt are column slices of wider
 * row-major matrices. ViTPose+ (MoEMlp, mmpose/models/backbones/vit_moe.py:97-115): the shared fc2 owns the first
 * D - part_features columns of the FFN output gradient, the expert of a crop's dataset the last part_features. */
int vpb_gemm_bf16_atb_accum_ld(const void* At, int lda, const void* Bt, int ldb, int M, int N, int K, float* out,
                               int ldo, void* stream);
/* Residual-stream GEMM with the following LayerNorm fused into its epilogue:
 *   out fp32 [M,N] = (epilogue == VPB_EPI_RESID_F32 ? aux[M,N] : aux[row % period, N]) + A.B^T + bias
 *   xn  bf16 [M,N] = LayerNorm(out row, eps) * gamma + beta
 * i.e. `x = x + attn.proj(..)` / `x = x + mlp.fc2(..)` / patch embed + pos embed together with the norm1 / norm2 /
 * last_norm that consumes the result (vit.py:137-140, :320, :328). The CTAs that own the column tiles of one
 * 128-row block exchange per-row (mean, M2) through `scratch` (vpb_gemm_layernorm_scratch_bytes(M, N) bytes,
 * 16-byte aligned, contents don't care), so the rows are normalised while they are still in tensor memory.
 * out may alias aux (in-place residual update).
 * row_scale (optional fp32 [ceil(M / rows_per_scale)], VPB_EPI_RESID_F32 only): the GEMM branch of rows
 * [i*rows_per_scale, (i+1)*rows_per_scale) is multiplied by row_scale[i] before the residual add — stochastic depth,
 * `x + drop_path(branch)` with row_scale = mask / keep_prob per crop (vit.py:48-56,138-139). */
size_t vpb_gemm_layernorm_scratch_bytes(int M, int N);
int vpb_gemm_bf16_layernorm(const void* A, const void* B, int M, int N, int K, int epilogue, const float* bias,
                            float* out, const float* aux, int period, const float* gamma, const float* beta, float eps,
                            void* xn, void* scratch, size_t scratch_bytes, const float* row_scale, int rows_per_scale,
                            void* stream);
/* The same without the two memsets per call: initialise a scratch once (vpb_gemm_layernorm_scratch_init), then pass
 * epoch = 1, 2, 3, ... to successive launches that use it. All launches on one scratch must have the same (M, N) and be
 * ordered on one stream; re-initialise before the epoch wraps. (vpb_vitpose_forward does this internally.) */
int vpb_gemm_layernorm_scratch_init(void* scratch, size_t scratch_bytes, int M, int N, void* stream);
int vpb_gemm_bf16_layernorm_seq(const void* A, const void* B, int M, int N, int K, int epilogue, const float* bias,
                                float* out, const float* aux, int period, const float* gamma, const float* beta,
                                float eps, void* xn, void* scratch, size_t scratch_bytes, unsigned epoch,
                                const float* row_scale, int rows_per_scale, void* stream);
/* Folded LayerNorm, the three pieces (see vpb_block_fold):
 *  vpb_fold_layernorm_linear: W fp32 [N, K], bias [N] or NULL, gamma / beta [K] -> Wf bf16 [N, K], s [N], c [N];
 *  vpb_gemm_bf16_resid_stats: out = aux + A.B^T + bias like vpb_gemm_bf16_layernorm (same epilogues), but xb receives
 *    the PLAIN bf16 copy of the updated rows and `stats` one (mean, M2) float pair per row and column tile:
 *    stats[(row * parts + tile) * 2 + {0, 1}], parts / part_cols from vpb_gemm_stats_layout(N, ...), rows padded to 128
 *    (vpb_gemm_stats_bytes). Shapes the fused-LayerNorm tiles do not cover (N not a multiple of 192 / 256 / 128 ...)
 *    return -2;
 *  vpb_gemm_bf16_lnfold: out[M, N] bf16 = act(LN(x) W^T + b) from A = xb, B = Wf, (s, c), the statistics; epilogue =
 *    VPB_EPI_BIAS_BF16 or VPB_EPI_GELU_BF16. K must equal parts * part_cols. */
int vpb_fold_layernorm_linear(const float* W, const float* bias, const float* gamma, const float* beta, int N, int K,
                              void* Wf, float* s, float* c, void* stream);
int vpb_gemm_stats_layout(int N, int* parts, int* part_cols);
size_t vpb_gemm_stats_bytes(int M, int N);
int vpb_gemm_bf16_resid_stats(const void* A, const void* B, int M, int N, int K, int epilogue, const float* bias,
                              float* out, const float* aux, int period, void* xb, void* stats, size_t stats_bytes,
                              const float* row_scale, int rows_per_scale, void* stream);
int vpb_gemm_bf16_lnfold(const void* A, const void* Wf, int M, int N, int K, int epilogue, const float* c,
                         const float* s, const void* stats, int parts, int part_cols, float eps, void* out, int ldo,
                         void* stream);
int vpb_layernorm_bf16(const float* x, const float* gamma, const float* beta, void* y, int M, int D, float eps,
                       void* stream);
int vpb_im2col_patch16(const float* img, void* patches, int n, int H, int W, int flip, void* stream);
int vpb_attention(const void* qkv, void* out, int n, int T, int heads, int head_dim, float scale, void* stream);
int vpb_deconv4x4s2_bn_relu(const void* in, const void* wphase, const float* scale, const float* shift, void* out,
                            int n, int h, int w, int cin, int cout, void* stream);
int vpb_conv3x3_nchw(const void* in, const void* w9, const float* bias, float* out, int n, int h, int w, int cin,
                     int cout, void* stream);
int vpb_relu_upsample_nhwc(const void* in, void* out, int n, int h, int w, int C, int factor, void* stream);
int vpb_tokens_to_nchw_f32(const void* tokens, float* out, int n, int T, int D, void* stream);

/* ---- backward pass of the training step (`*_bwd`) ------------------------------------
 * The reference trains through torch.autograd (mmcv runner: loss.backward(), optimizer.step()); these are the
 * operators that replace the autograd nodes of the path's modules. The tensor-core work of the backward pass is
 * vpb_gemm_bf16 itself on transposed operands:
 *   dgrad  dX[M,K] = dY[M,N] . W[N,K]   = vpb_gemm_bf16(A = dY, B = W^T [K,N], VPB_EPI_BIAS_BF16 / _RESID_F32)
 *   wgrad  dW[N,K] += dY^T . X          = vpb_gemm_bf16_atb_accum(At = dY [M,N], Bt = X [M,K], out = dW)
 *                                         (or vpb_gemm_bf16(dY^T, X^T, VPB_EPI_ACCUM_F32) on transposed copies)
 * All matrices row-major; bf16 unless stated; gradients of parameters are fp32 and ACCUMULATED into their buffers. */
int vpb_transpose_bf16(const void* in, void* out, int R, int C, int batch, void* stream);   /* out[b][C,R] = in[b][R,C]^T */
/* out = bf16(in), optionally times row_scale[(i / row_len) / rows_per_scale] (gradient of a stochastic-depth branch) */
int vpb_cast_f32_bf16(const float* in, void* out, long long n, const float* row_scale, int row_len,
                      int rows_per_scale, void* stream);
/* out[C] += sum over the R rows of in[R,C] (bf16, or fp32 when is_f32): bias / pos-embed gradients */
int vpb_colsum_accumulate(const void* in, int is_f32, int R, int C, float* out, void* stream);
/* nn.GELU (exact erf, vit.py:71-76): out = gelu(pre); dpre = dh * gelu'(pre) */
/* Simple decoder (topdown_heatmap_simple_head.py:132-139 with num_deconv_layers = 0: ReLU -> bilinear x f ->
 * Conv2d 3x3) in its un-materialised form, as vpb_vitpose_forward runs it, and its backward:
 *   r = vpb_relu_bf16(tokens);  z fp32 [images, 9K, h*w] = vpb_gemm_bf16(r, W9 [9K, D] (row k*9 + ky*3 + kx),
 *   VPB_EPI_NCHW_F32, period h*w);  heatmaps = vpb_simple_head_gather(z, bias)  (bilinear gather of the nine tap maps);
 *   backward: dz bf16 [images*h*w, ldz] (token-major, column k*9+t, ldz >= 9K; pad columns are NOT written) =
 *   vpb_simple_head_gather_bwd(dheatmaps); then two GEMMs on dz (input / weight gradient of the tap GEMM) and
 *   vpb_relu_bwd_bf16(r, dr) = dr where r > 0. */
int vpb_relu_bf16(const void* in, void* out, long long n, void* stream);
int vpb_relu_bwd_bf16(const void* y, const void* dy, void* dx, long long n, void* stream);
int vpb_simple_head_gather(const float* z, const float* bias, float* out, int images, int K, int h, int w, int factor,
                           void* stream);
int vpb_simple_head_gather_bwd(const float* dout, void* dz, int ldz, int images, int K, int h, int w, int factor,
                               void* stream);
/* Training-step fusions of the MLP (vit.py:64-73: fc1 -> nn.GELU -> fc2):
 *  vpb_gemm_bf16_gelu_save: out bf16 [M, ldo] = gelu_erf(A.B^T + bias) AND pre_out bf16 [M, N] = A.B^T + bias (the
 *    activation the backward pass differentiates; the GELU is evaluated on the fp32 value, before that rounding);
 *  vpb_gemm_bf16_gelu_bwd:  out bf16 [M, ldo] = (A.B^T) * gelu'(pre[M, N]): A = dY of fc2, B = W2^T, so out = dL/d(pre);
 *    colsum (optional, fp32 [N]) += column sums of out's fp32 values = fc1's bias gradient;
 *  vpb_cast_f32_bf16_colsum: vpb_cast_f32_bf16 over [R, C] + colsum[C] += column sums of the rounded output (the bias
 *    gradient of the layer whose output gradient is being cast). */
int vpb_gemm_bf16_gelu_save(const void* A, const void* B, int M, int N, int K, const float* bias, void* out, int ldo,
                            void* pre_out, void* stream);
int vpb_gemm_bf16_gelu_bwd(const void* A, const void* B, int M, int N, int K, const void* pre, void* out, int ldo,
                           float* colsum, void* stream);
int vpb_cast_f32_bf16_colsum(const float* in, void* out, int R, int C, const float* row_scale, int rows_per_scale,
                             float* colsum, void* stream);
int vpb_gelu_fwd_bf16(const void* pre, void* out, long long n, void* stream);
int vpb_gelu_bwd_bf16(const void* pre, const void* dh, void* dpre, long long n, void* stream);
/* nn.LayerNorm backward (vit.py:125,133,328): x fp32 [M,D] (the saved input), dy bf16 [M,D];
 * dx_accum fp32 [M,D] += dL/dx (the residual stream's gradient), dgamma / dbeta fp32 [D] += */
int vpb_layernorm_bwd(const float* x, const float* gamma, const void* dy, float* dx_accum, float* dgamma,
                      float* dbeta, int M, int D, float eps, void* stream);
/* Attention.forward (vit.py:99-115) that also returns lse fp32 [n, heads, T] = log2 sum_j exp2(s_ij * scale * log2 e),
 * and its backward: qkv [n,T,3*heads*hd], out [n,T,heads*hd] and lse saved by the forward pass, dout = dL/dout ->
 * dqkv [n,T,3*heads*hd]. Backward: T = 192, head_dim 32 / 64 / 80 (ViTPose-S / -B, -L / -H). */
int vpb_attention_lse(const void* qkv, void* out, float* lse, int n, int T, int heads, int head_dim, float scale,
                      void* stream);
int vpb_attention_bwd(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, int n, int T,
                      int heads, int head_dim, float scale, void* stream);
/* the same, and dbias fp32 [3*heads*head_dim] += the column sums of dqkv over all tokens (attn.qkv's bias gradient) */
int vpb_attention_bwd_bias(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, float* dbias,
                           int n, int T, int heads, int head_dim, float scale, void* stream);
/* ConvTranspose2d(k4,s2,p1,bias=False) without BatchNorm / ReLU (training forward; ones / zeros: fp32 [cout]) */
int vpb_deconv4x4s2_raw(const void* in, const void* wphase, void* out, int n, int h, int w, int cin, int cout,
                        const float* ones, const float* zeros, void* stream);
/* nn.BatchNorm2d in training mode on NHWC rows [rows, C] (simple_head.py:324-333): batch mean / rstd (biased
 * variance), running statistics updated with `momentum` and the unbiased variance (NULL to skip);
 * scratch = 4*C floats (2*C fp64 accumulators, 8-byte aligned: the unordered atomic sums stay reproducible to the last
 * fp32 bit). Then act = relu(bn(raw)), and its backward (dgamma / dbeta fp32 [C] +=). */
int vpb_bn_train_stats(const void* raw, long long rows, int C, float eps, float momentum, float* sum_sumsq_scratch,
                       float* mean, float* rstd, float* running_mean, float* running_var, void* stream);
int vpb_bn_relu_fwd(const void* raw, void* act, const float* mean, const float* rstd, const float* gamma,
                    const float* beta, long long rows, int C, void* stream);
int vpb_bn_relu_bwd(const void* raw, const void* dact, void* draw, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, float* dgamma, float* dbeta, float* scratch /* 6*C floats */,
                    long long rows, int C, void* stream);
/* the same for a BatchNorm2d in EVAL mode inside forward_train (mean / rstd = the running statistics): no batch-mean
 * correction terms in draw; dgamma / dbeta as above */
int vpb_bn_relu_bwd_eval(const void* raw, const void* dact, void* draw, const float* mean, const float* rstd,
                         const float* gamma, const float* beta, float* dgamma, float* dbeta, float* scratch,
                         long long rows, int C, void* stream);
/* dL/dheatmaps fp32 [n,K,P] -> bf16 rows [n*P, Kp] (zero padded to Kp >= K): operand of the final conv's GEMMs */
int vpb_nchw_f32_to_rows_bf16(const float* in, void* out, int n, int K, int P, int Kp, void* stream);
/* Operand gathers for the backward of the 4-phase transposed convolution (layouts in csrc/train_bwd.cu):
 *   gather_x : x [n,h,w,cin]   -> [4 phases][n*h*w, 4*cin]   (wgrad: dWp[ph] = phase_dy[ph]^T . gather_x[ph])
 *   phase_dy : dy [n,2h,2w,cout] -> [4 phases][n*h*w, cout]
 *   gather_dy: dy [n,2h,2w,cout] -> [n*h*w, 16*cout]          (dgrad: dx = gather_dy . W2g^T) */
int vpb_deconv_gather_x(const void* x, void* out, int n, int h, int w, int cin, void* stream);
int vpb_deconv_gather_dy(const void* dy, void* out, int n, int h, int w, int cout, void* stream);
int vpb_deconv_phase_dy(const void* dy, void* out, int n, int h, int w, int cout, void* stream);
/* ConvTranspose2d weight fp32 [Cin, Cout, 4, 4] (nn.ConvTranspose2d layout, simple_head.py:324-333) -> the packed bf16
 * operands in one launch: wp [4 phases][Cout][4 taps * Cin] (vpb_weights.deconv_w) and, unless NULL, wd [Cin][16 * Cout]
 * with column (phase * 4 + tap) * Cout + co (B operand of the input gradient over vpb_deconv_gather_dy);
 * vpb_deconv_unpack_wgrad maps a packed fp32 weight gradient [4][Cout][4 * Cin] back to [Cin, Cout, 4, 4]. */
int vpb_deconv_pack_weight(const float* w, void* wp, void* wd, int cin, int cout, void* stream);
int vpb_deconv_unpack_wgrad(const float* dwp, float* dw, int cin, int cout, void* stream);

/* ---- post-decode evaluation step ----
 * Per-image rescoring + OKS NMS of the top-down COCO datasets (topdown_coco_dataset.py:476-503; oks_iou / oks_nms /
 * soft_oks_nms, mmpose/core/post_processing/nms.py:51-207). The P poses are grouped by image: group g is rows
 * [group_start[g], group_start[g+1]).
 *   kpts        fp32 [P,K,3] (x, y, score);  areas fp64 [P];  box_scores fp64 [P]
 *   var         fp64 [K] = (2 * sigma_k)^2
 *   rescore     bit 0: pose score = mean(joint scores > vis_thr) * box score; else box_scores are the pose scores.
 *               bit 1: the areas are float32 values (TopDownCocoDataset.evaluate passes boxes[:, 4]): their pair sum
 *               in the OKS denominator is rounded to float32 as NumPy does for float32 operands
 *   use_vis     != 0: only joints whose DETECTION score exceeds vis_thr enter the OKS (nms.py:80-82)
 *   soft        0: oks_nms (keep while OKS <= thr); 1: soft_oks_nms (Gaussian rescoring, at most max_dets per image)
 *   scores_out  fp64 [P] scores used;  keep int32 [P]: kept row indices of group g, in selection order, starting at
 *               keep[group_start[g]];  keep_count int32 [G].  max_group = largest group size (<= 2048). */
int vpb_oks_nms(const float* kpts, const double* areas, const double* box_scores, const int32_t* group_start, int G,
                int K, int max_group, const double* var, double thr, int use_vis, double vis_thr, int rescore, int soft,
                int max_dets, double* scores_out, int32_t* keep, int32_t* keep_count, void* stream);

/* ---- preprocessing ----
 * TopDownAffine + ToTensor + NormalizeTensor for n boxes (mmpose/datasets/pipelines/top_down_transform.py:295-364,
 * shared_transform.py:21-65): out[i] = ((warpAffine_u8(src[i], M_i, (out_w,out_h), INTER_LINEAR) / 255) - mean) / std
 * as fp32 NCHW, bit-identical to cv2.warpAffine on uint8 + torchvision to_tensor/normalize.
 *   src_ptrs  device array [n] of device pointers to uint8 HWC (3-channel, dense) images
 *   src_hw    device int32 [n,2] (height, width) of each source image
 *   inv_mats  device float64 [n,6]: the INVERSE (dst->src) 2x3 affine map, i.e. cv::invertAffineTransform of the
 *             matrix get_warp_matrix / get_affine_transform returns (computed by the host in double)
 *   mean3/std3 HOST float[3] */
int vpb_warp_affine_normalize(const unsigned char* const* src_ptrs, const int32_t* src_hw, const double* inv_mats,
                              int n, int out_h, int out_w, const float* mean3, const float* std3, float* out,
                              void* stream);

/* ---- training-step operators: loss, gradient norm, AdamW. The backward-pass operators of
 * the network itself (attention / LayerNorm / GELU / BatchNorm backward, weight-gradient GEMMs) follow below;
 * vitpose_b200/training.py sequences them into one autograd node. ----
 * JointsMSELoss.forward (mmpose/models/losses/mse_loss.py:24-45): loss[0] = loss_weight / K * sum_k mean_{n,hw}
 * ((output - target) * target_weight[n,k])^2; target_weight may be NULL (use_target_weight=False);
 * grad_output (optional, same shape as output) receives d loss / d output. */
int vpb_joints_mse_loss(const float* output, const float* target, const float* target_weight, int N, int K, int HW,
                        float loss_weight, float* loss, float* grad_output, void* stream);
/* adds sum(grad^2) to *sq_norm_accum (zero it once per step; sqrt = clip_grad_norm_'s total norm) */
int vpb_grad_sq_norm_accumulate(const float* grad, long long n, float* sq_norm_accum, void* stream);
/* torch.optim.AdamW update of one tensor with its group's lr / weight_decay (layer decay:
 * mmcv_custom/layer_decay_optimizer_constructor.py:6-78); when sq_norm != NULL gradients are scaled by
 * min(1, max_norm / (sqrt(*sq_norm) + 1e-6)) first (grad_clip=dict(max_norm=1.)). */
int vpb_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, const float* sq_norm,
                   float max_norm, void* stream);

/* pose_pck_accuracy's arithmetic (mmpose/core/evaluation/top_down_eval.py:8-60,133-215), the `acc_pose` value of
 * TopdownHeatmapSimpleHead.get_accuracy (simple_head.py:170-195): pred / gt fp32 [N,K,2] are the arg-max coordinates
 * of the output and target heatmaps (vpb_decode_heatmaps with VPB_DECODE_NONE, no transform), weight fp32 [N,K]
 * (> 0 = visible). norm0 / norm1 divide x / y (the reference passes (H, W)). acc fp32 [K] (-1 = no visible sample),
 * avg fp32 [1], cnt int32 [1]. */
int vpb_pose_pck_accuracy(const float* pred, const float* gt, const float* weight, int N, int K, float norm0,
                          float norm1, float thr, float* acc, float* avg, int32_t* cnt, void* stream);

/* Multi-tensor form: one call updates every parameter (and computes the clip norm first when sq_norm != NULL).
 * entries: DEVICE array of n descriptors; chunk_start: DEVICE int[n+1], prefix sum of ceil(n_i / 4096);
 * total_chunks = chunk_start[n]. `step` is the per-tensor AdamW step count (>= 1) for the bias correction.
 * sq_norm: NULL (no clipping) or DEVICE float[1 + total_chunks]: [0] receives sum(grad^2), the rest is scratch for the
 * per-chunk partials, which are added in a fixed order so that all data-parallel replicas clip identically. */
typedef struct vpb_tensor_entry {
  float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
  long long n; float lr; float weight_decay; int32_t step; int32_t pad_;
} vpb_tensor_entry;
int vpb_adamw_multi(const vpb_tensor_entry* entries, const int32_t* chunk_start, int n, int total_chunks, float beta1,
                    float beta2, float eps, float* sq_norm, float max_norm, void* stream);

/* bf16 operand copies of all linear layers of the training step in one launch: for every entry, w[rows, cols] =
 * bf16(src) and wt[cols, rows] = bf16(src)^T (round to nearest even). Replaces, per layer, the implicit fp32 -> bf16
 * weight use of nn.Linear in the reference's autograd graph (mmpose/models/backbones/vit.py:79-85, 100-121: attn.qkv /
 * attn.proj / mlp.fc1 / mlp.fc2; patch_embed.proj :143-165 as a [D, 3*16*16] matrix). `entries` and `tile_start`
 * ([n + 1] prefix sums of ceil(rows/64) * ceil(cols/64)) live in device memory. */
typedef struct vpb_cast_entry {
  const float* src; void* w; void* wt; int32_t rows; int32_t cols;
} vpb_cast_entry;
int vpb_cast_transpose_multi(const vpb_cast_entry* entries, const int32_t* tile_start, int n, int total_tiles,
                             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITPOSE_B200_H_ */
