// Internal C++ launchers behind the C-ABI (include/vitpose_b200.h). Device pointers only; every call is
// asynchronous on `stream`; return 0 on success, non-zero with vpb::get_last_error() set.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

struct vpb_tensor_entry;
struct vpb_cast_entry;

namespace vpb {

struct GemmParams;

// ---- GEMM (gemm.cu) ----
struct GemmMaps { CUtensorMap a, b, out, aux, ln; };
int gemm_pick_bn(int N, int epilogue);
int gemm_pick_cg(int M, int bn, int epilogue, int K);
int make_gemm_maps(GemmMaps* maps, const void* A, const void* B, int M, int N, int K, int lda, int ldb, int bn,
                   int epilogue, void* out, int ldo, const float* aux, int cg, int period);
int launch_gemm(const GemmMaps& maps, const GemmParams& p, int bn, int epilogue, int cg, int max_ctas,
                cudaStream_t stream);
// Consumer side of a folded LayerNorm (gemm.cuh): A holds the plain bf16 rows of the residual stream, `stats` the
// (mean, M2) pairs per row and column tile its producer left, `s` [N] the row sums of the gamma-folded weight; the
// folded bias c goes in as `bias`.
struct LnFoldIn { const void* stats; const float* s; int parts, part_cols; float eps; };
// Training-step extras of the bf16 epilogues: EPI_GELU_BF16 can also save the bf16 pre-activation (pre_out [M, N]);
// EPI_DGELU_BF16 (input gradient of mlp.fc2 times gelu'(pre)) reads it back (pre_in) and can add the column sums of
// its output to colsum_out [N] (fc1's bias gradient).
struct GemmTrainAux { void* pre_out; const void* pre_in; float* colsum_out; };
int gemm_bf16(const void* A, const void* B, int M, int N, int K, int epilogue, const float* bias, void* out, int ldo,
              const float* aux, int period, int max_ctas, cudaStream_t stream, const LnFoldIn* ln = nullptr,
              const GemmTrainAux* tr = nullptr);
// out[M, N] (fp32, leading dimension ldo) += At^T . Bt for ROW-MAJOR At [K, M], Bt [K, N]: the weight gradient
// dW = dY^T X straight from dY and X (MN-major tensor-core operands, K split over CTAs, atomic accumulation)
int gemm_bf16_atb_accum(const void* At, const void* Bt, int M, int N, int K, float* out, int ldo, int max_ctas,
                        cudaStream_t stream);
int gemm_bf16_atb_accum_ld(const void* At, int lda, const void* Bt, int ldb, int M, int N, int K, float* out, int ldo,
                           int max_ctas, cudaStream_t stream);
// Residual / patch-embed GEMM (epilogue EPI_RESID_F32 or EPI_POS_F32, out fp32 [M, N]) that also writes
// xn = LayerNorm(out) * gamma + beta as bf16 [M, N] from the same kernel. `scratch` (gemm_ln_scratch_bytes(M, N)
// bytes, 16-byte aligned) holds the per-row partial statistics the column tiles exchange; gemm_ln_scratch_init
// prepares it for a sequence of stream-ordered calls numbered epoch = 1, 2, ... (same M, N).
// Falls back to GEMM + layernorm_bf16 (two kernels) for shapes the fused kernel does not cover.
size_t gemm_ln_scratch_bytes(int M, int N);
int gemm_ln_scratch_init(void* scratch, int M, int N, cudaStream_t stream);
int gemm_bf16_ln(const void* A, const void* B, int M, int N, int K, int epilogue, const float* bias, float* out,
                 const float* aux, int period, const float* gamma, const float* beta, float eps, void* xn,
                 void* scratch, unsigned epoch, int max_ctas, cudaStream_t stream, const float* row_scale = nullptr,
                 int rows_per_scale = 0, void* fold_stats = nullptr);
// fold_stats != null: folded LayerNorm — xn receives the PLAIN bf16 copy of the updated rows and fold_stats
// (gemm_ln_stats_bytes(M, N) bytes, 16-byte aligned) one (mean, M2) float pair per row and column tile
// (gemm_ln_parts(N) tiles of gemm_ln_part_cols(N) columns); gamma / beta / scratch / epoch are not used.
int gemm_ln_parts(int N);
int gemm_ln_part_cols(int N);
size_t gemm_ln_stats_bytes(int M, int N);
// W fp32 [N, K], bias [N] (may be null), gamma / beta [K] -> Wf bf16 [N, K] = gamma o W, s [N] = row sums of the
// ROUNDED Wf, c [N] = bias + W . beta: the operands of a Linear layer that applies the preceding LayerNorm itself
int fold_layernorm_linear(const float* W, const float* bias, const float* gamma, const float* beta, int N, int K,
                          void* Wf, float* s, float* c, cudaStream_t stream);

// ---- elementwise / normalisation (elementwise.cu) ----
// img fp32 [n,3,H,W] -> patches bf16 [(flip?2n:n) * Hp*Wp, 768]; rows [n*Hp*Wp, 2n*Hp*Wp) hold the
// horizontally flipped crops (img.flip(3)) when flip != 0.
int im2col_patch16(const float* img, void* patches, int n, int H, int W, int flip, cudaStream_t stream);
// x fp32 [M, D] -> y bf16 [M, D], LayerNorm over D with affine (gamma, beta), eps
int layernorm_bf16(const float* x, const float* gamma, const float* beta, void* y, int M, int D, float eps,
                   cudaStream_t stream);
// tokens bf16 [n, T, D] -> fp32 NCHW [n, D, T]  (ViT.forward's permute for standalone backbone calls)
int tokens_to_nchw_f32(const void* tokens, float* out, int n, int T, int D, cudaStream_t stream);
// relu + bilinear x`factor` upsample (align_corners=False), bf16 NHWC [n,h,w,C] -> bf16 NHWC [n,h*f,w*f,C]
int relu_upsample_bilinear_nhwc(const void* in, void* out, int n, int h, int w, int C, int factor,
                                cudaStream_t stream);

// Simple decoder without the upsampled map (see elementwise.cu): relu on the bf16 features, then — after the tap GEMM
// z = relu(x) . W9^T (fp32 [images, 9K, h*w]) — the bilinear gather of the nine tap maps + bias -> fp32 NCHW heatmaps
int relu_bf16(const void* in, void* out, long long n, cudaStream_t stream);
int relu_bwd_bf16(const void* y, const void* dy, void* dx, long long n, cudaStream_t stream);
int simple_head_gather_bwd(const float* dout, void* dz, int ldz, int images, int K, int h, int w, int factor,
                           cudaStream_t stream);
int simple_head_gather(const float* z, const float* bias, float* out, int images, int K, int h, int w, int factor,
                       cudaStream_t stream);

// ---- attention (attention.cu) ----
// qkv bf16 [n, T, 3*heads*hd] (column order: which, head, d) -> out bf16 [n, T, heads*hd]
int attention_fwd(const void* qkv, void* out, int n, int T, int heads, int hd, float scale, int max_ctas,
                  cudaStream_t stream, float* lse = nullptr);

// ---- deconv / conv implicit GEMMs (conv.cu) ----
// ConvTranspose2d(k4,s2,p1,no bias) + folded BN + ReLU, NHWC bf16.
//   in [n,h,w,Cin] -> out [n,2h,2w,Cout]; wphase bf16 [4 phases][Cout][4 taps * Cin] (see pack in python),
//   scale/shift fp32 [Cout] applied as relu(acc * scale + shift)
// relu = 0: affine output without the activation (training forward: raw conv output for BatchNorm statistics)
int deconv4x4s2_affine(const void* in, const void* wphase, const float* scale, const float* shift, void* out, int n,
                       int h, int w, int cin, int cout, int relu, int max_ctas, cudaStream_t stream);
int deconv4x4s2_bn_relu(const void* in, const void* wphase, const float* scale, const float* shift, void* out,
                        int n, int h, int w, int cin, int cout, int max_ctas, cudaStream_t stream);
// Conv2d 3x3 pad 1 + bias, NHWC bf16 in [n,h,w,Cin], weights bf16 [Cout][9 taps * Cin] -> fp32 NCHW [n,Cout,h,w]
int conv3x3_nchw_out(const void* in, const void* w9, const float* bias, float* out, int n, int h, int w, int cin,
                     int cout, int max_ctas, cudaStream_t stream);

// ---- decode (decode.cu) ----
enum DecodeMode { DECODE_NONE = 0, DECODE_DEFAULT = 1, DECODE_UNBIASED = 2, DECODE_UDP_DARK = 3 };
int decode_heatmaps(const float* hm, const float* hm_flipped, const int* flip_index, int shift_heatmap, int N, int K,
                    int H, int W, int mode, int kernel, int use_udp, int apply_transform, const float* center,
                    const float* scale, float* preds, float* maxvals, float* merged_out, int* argmax_out,
                    cudaStream_t stream);

int flip_back(const float* in, const int* perm, float* out, int N, int K, int H, int W, int shift,
              cudaStream_t stream);
int transform_preds(const float* coords, const float* center, const float* scale, float* out, int N, int K, int W,
                    int H, int use_udp, cudaStream_t stream);

// ---- preprocessing (preprocess.cu) ----
int warp_affine_normalize(const unsigned char* const* src_ptrs, const int* src_hw, const double* inv_mats, int n,
                          int out_h, int out_w, const float* mean3, const float* std3, float* out,
                          cudaStream_t stream);

// ---- training-step operators (train_ops.cu) ----
int joints_mse_loss(const float* output, const float* target, const float* target_weight, int N, int K, int HW,
                    float loss_weight, float* loss, float* grad_output, cudaStream_t stream);
int grad_sq_norm_accumulate(const float* grad, long long n, float* sq_norm_accum, cudaStream_t stream);
int adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
               float beta2, float eps, float weight_decay, int step, const float* sq_norm, float max_norm,
               cudaStream_t stream);

// ---- backward pass of the training step (train_bwd.cu, attention_bwd.cu) ----
int transpose_bf16(const void* in, void* out, int R, int C, int batch, cudaStream_t stream);
// out bf16 [R, C] = in * row_scale (as cast_f32_bf16) and colsum[C] += column sums of the ROUNDED output: the bias
// gradient of the Linear layer whose output gradient this is, without a second pass over it
int cast_f32_bf16_colsum(const float* in, void* out, int R, int C, const float* row_scale, int rows_per_scale,
                         float* colsum, cudaStream_t stream);
int cast_f32_bf16(const float* in, void* out, long long n, cudaStream_t stream, const float* row_scale = nullptr,
                  int row_len = 0, int rows_per_scale = 0);
int colsum_accumulate(const void* in, int is_f32, int R, int C, float* out, cudaStream_t stream);
int colsum_sq_accumulate(const void* in, int R, int C, double* sum, double* sumsq, cudaStream_t stream);
int gelu_fwd_bf16(const void* pre, void* out, long long n, cudaStream_t stream);
int gelu_bwd_bf16(const void* pre, const void* dh, void* dpre, long long n, cudaStream_t stream);
int layernorm_bwd(const float* x, const float* gamma, const void* dy, float* dx_accum, float* dgamma, float* dbeta,
                  int M, int D, float eps, cudaStream_t stream);
int bn_finalize(const double* sum, const double* sumsq, float* mean, float* rstd, float* running_mean, float* running_var,
                int C, long long rows, float eps, float momentum, cudaStream_t stream);
int bn_relu_fwd(const void* raw, void* act, const float* mean, const float* rstd, const float* gamma, const float* beta,
                long long rows, int C, cudaStream_t stream);
int bn_relu_bwd_reduce(const void* raw, const void* dact, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, int R, int C, double* acc, float* sums_f, float* dbeta, float* dgamma,
                       cudaStream_t stream);
int bn_relu_bwd(const void* raw, const void* dact, void* draw, const float* mean, const float* rstd, const float* gamma,
                const float* beta, const float* dbeta, const float* dgamma, long long rows, int C, cudaStream_t stream);
int nchw_f32_to_rows_bf16(const float* in, void* out, int n, int K, int P, int Kp, cudaStream_t stream);
int deconv_gather_x(const void* x, void* out, int n, int h, int w, int cin, cudaStream_t stream);
int deconv_gather_dy(const void* dy, void* out, int n, int h, int w, int cout, cudaStream_t stream);
int deconv_phase_dy(const void* dy, void* out, int n, int h, int w, int cout, cudaStream_t stream);
// dbias (optional, fp32 [3*heads*hd]) += column sums of dqkv (fp32 values, before the bf16 rounding): attn.qkv's bias gradient
int attention_bwd(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, int n, int T,
                  int heads, int hd, float scale, cudaStream_t stream, float* dbias = nullptr);

int adamw_multi(const vpb_tensor_entry* entries, const int* chunk_start, int n, int total_chunks, float beta1,
                float beta2, float eps, float* sq_norm, float max_norm, cudaStream_t stream);
int cast_transpose_multi(const vpb_cast_entry* entries, const int* tile_start, int n, int total_tiles,
                         cudaStream_t stream);

int pose_pck_accuracy(const float* pred, const float* gt, const float* weight, int N, int K, float norm0, float norm1,
                      float thr, float* acc, float* avg, int* cnt, cudaStream_t stream);

// ---- post-decode evaluation step (nms.cu) ----
int oks_nms(const float* kpts, const double* areas, const double* box_scores, const int* group_start, int G, int K,
            int max_group, const double* var, double thr, int use_vis, double vis_thr, int rescore, int soft,
            int max_dets, double* scores_out, int* keep, int* keep_count, cudaStream_t stream);

// ConvTranspose2d weight fp32 [Cin, Cout, 4, 4] -> packed bf16 operands (wp: forward, wd: input gradient, may be null)
// and packed fp32 weight gradient [4][Cout][4 * Cin] -> [Cin, Cout, 4, 4] (train_bwd.cu)
int deconv_pack_weight(const float* w, void* wp, void* wd, int cin, int cout, cudaStream_t stream);
int deconv_unpack_wgrad(const float* dwp, float* dw, int cin, int cout, cudaStream_t stream);

}  // namespace vpb
