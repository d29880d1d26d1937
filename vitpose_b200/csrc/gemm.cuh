// Persistent warp-specialised bf16 GEMM for sm_100a: C[M,N] = A[M,K] * B[N,K]^T with fused epilogues.
//   warp 0      : TMA producer (A and B tiles, 128B-swizzled) through a STAGES-deep mbarrier ring
//   warp 1      : tcgen05.mma issuer (one elected lane), accumulators double-buffered in TMEM
//   warps 2..   : epilogue warpgroup(s). tcgen05.ld -> bias / GELU / residual (/ LayerNorm) in registers -> 128B-swizzled
//                 staging tile in shared memory -> TMA store (coalesced, clipped at the tensor edge by hardware).
//                 bf16 epilogues: two warpgroups split the column chunks of a tile; the short-K LayerNorm variant: two
//                 warpgroups on alternating tiles; everything else: one warpgroup.
//   warp 6 (11) : ring warp of the fp32 residual epilogues: TMA-loads the residual / positional chunk into a slot,
//                 TMA-stores the slot once the epilogue has updated it in place, recycles it.
// TMEM is released as soon as the last chunk of a tile has left it, so the next tile's MMAs overlap the epilogue.
// One CTA (or CTA pair, cta_group::2) per SM, tiles 128 (256) x BN, walked N-fastest so CTAs running together share
// the A panel in L2. Variants: EPI_ACCUM_F32 splits K over CTAs (weight gradients), OPM = 1 consumes transposed
// operands as MN-major tiles, EPI_*_LN* fuse the LayerNorm that follows a residual update.
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace vpb {

enum GemmEpilogue {
  EPI_BIAS_BF16 = 0,   // out bf16 [M, ldo]      = acc + bias
  EPI_GELU_BF16 = 1,   // out bf16 [M, ldo]      = gelu_erf(acc + bias)
  EPI_RESID_F32 = 2,   // out fp32 [M, ldo]      = aux[M, ldo] + acc + bias        (residual stream, may alias out)
  EPI_POS_F32 = 3,     // out fp32 [M, ldo]      = acc + bias + aux[(row % period), N] (patch embed + pos embed)
  EPI_NCHW_F32 = 4,    // out fp32 [M/period, N, period] = acc + bias             (final conv -> heatmaps)
  EPI_POSTMA_F32 = 5,  // as EPI_POS_F32 with the positional rows streamed by TMA (needs period % 64 == 0); internal
  // EPI_RESID_F32 / EPI_POSTMA_F32 that ALSO write LayerNorm(out row) * gamma + beta as bf16 [M, N] (N % BN == 0):
  // the LayerNorm that follows every residual update (vit.py:138-139, 328) fused into the producing GEMM.
  EPI_RESID_LN_F32 = 6,
  EPI_POSTMA_LN_F32 = 7,
  // the same for short-K GEMMs (attn.proj, patch embed), where the epilogue and not the MMA loop sets the pace:
  // 3 operand stages, a deep residual ring, and the normalise pass on its own warpgroup (see gemm_ln_split)
  EPI_RESID_LNS_F32 = 8,
  EPI_POSTMA_LNS_F32 = 9,
  // out fp32 [M, ldo] += A.B^T with the K range of every tile split over `ksplit` CTAs that accumulate with vector
  // atomics: weight gradients (small M x N, K = all tokens of the batch) would otherwise occupy a few SMs
  EPI_ACCUM_F32 = 10,
  // training: out bf16 [M, ldo] = acc * gelu'(pre), pre bf16 [M, N] the saved fc1 pre-activation (the input gradient of
  // mlp.fc2 with the backward of nn.GELU in its epilogue); optionally colsum[N] += column sums of out (fc1's bias gradient)
  EPI_DGELU_BF16 = 11,
  // training forward of mlp.fc1: EPI_GELU_BF16 + the bf16 pre-activation acc + bias stored to pre_out [M, N] (an
  // instance of its own so that the inference kernel carries none of it)
  EPI_GELU_SAVE_BF16 = 12,
  // consumer side of a folded LayerNorm (see GemmParams::ln_stats): EPI_BIAS_BF16 / EPI_GELU_BF16 of LN(A) from the
  // plain rows A, the producer's row statistics and the folded weight (instances of their own, as above)
  EPI_BIAS_LNIN_BF16 = 13,
  EPI_GELU_LNIN_BF16 = 14,
};
__host__ __device__ constexpr bool gemm_epi_ln(int epi) { return epi >= EPI_RESID_LN_F32 && epi <= EPI_POSTMA_LNS_F32; }
__host__ __device__ constexpr bool gemm_epi_pos(int epi) {
  return epi == EPI_POSTMA_F32 || epi == EPI_POSTMA_LN_F32 || epi == EPI_POSTMA_LNS_F32;
}
// epilogues that add an fp32 tile fetched by the residual-producer warp and store fp32 through the in-place ring
__host__ __device__ constexpr bool gemm_epi_adds_tile(int epi) {
  return epi == EPI_RESID_F32 || epi == EPI_POSTMA_F32 || gemm_epi_ln(epi);
}

struct GemmParams {
  int M, N, K;
  const float* bias;   // [N], may be null
  void* out;
  int ldo;             // leading dimension of out / aux in elements
  const float* aux;    // residual or positional table
  int period;          // tokens per crop (EPI_POS) / pixels per crop (EPI_NCHW)
  int ksplit;          // EPI_ACCUM_F32: CTAs per output tile (>= 1); 1 for every other epilogue
  // residual epilogues: out = aux + row_scale[row / scale_period] * (acc + bias); null = 1. Stochastic depth
  // (DropPath, vit.py:48-56,138-139): the branch of crop i is multiplied by mask_i / keep_prob.
  const float* row_scale;
  int scale_period;
  // fused LayerNorm (EPI_*_LN_F32): affine parameters [N] and the per-row partial statistics (mean, M2 of a tile's
  // BN columns, one 8-byte word each) that the CTAs owning the n-tiles of one row block exchange through global
  // memory. Launch e of a sequence writes region e & 1 and tags its words with bit (e >> 1) & 1 in the sign of M2,
  // so a reader can tell them from the words launch e - 2 left there (see gemm_ln_scratch_init).
  const float* ln_gamma;
  const float* ln_beta;
  unsigned long long* ln_part;   // two regions of [row blocks * 128, n_tiles] words, used alternately by launches
  size_t ln_region;              // words per region
  unsigned ln_epoch;
  float ln_eps;
  // "Folded" LayerNorm (round 2). The LayerNorm that follows a residual update is an affine map of the row followed
  // by a Linear layer, so it can be applied AFTER that layer's GEMM:
  //   LN(x) W^T + b = rstd * (x (gamma o W)^T - mean * s) + c,   s_n = sum_k (gamma o W)[n, k],  c_n = b_n + sum_k beta_k W[n, k]
  // The producer (EPI_*_LN* with ln_fold != 0) writes the fp32 rows, their plain bf16 copy (the next GEMM's A operand)
  // and one (mean, M2) pair per row and column tile — single pass, no exchange between CTAs, TMEM released as soon as
  // the tile is in registers. The consumer (EPI_BIAS_BF16 / EPI_GELU_BF16 with ln_stats != null) merges the pairs of
  // its rows and applies rstd, mean, s and c (passed as `bias`) in its epilogue.
  // training (EPI_GELU_SAVE_BF16): the bf16 pre-activation acc + bias [M, N] is stored too (through the aux tensor map)
  void* pre_out;
  // training (EPI_DGELU_BF16): the saved pre-activation [M, N], and optionally where to add the column sums of out
  const void* pre_in;
  float* colsum_out;
  int ln_fold;
  float2* ln_stats_out;     // producer: [row blocks * 128, n_tiles]
  const float2* ln_stats;   // consumer: [rows of A (padded to 128), ln_parts]
  const float* ln_s;        // consumer: [N]
  int ln_parts;             // consumer: column tiles per row of the producer
  int ln_part_cols;         //           and their width
  // Pipeline-depth switches (gemm_flags(): VPB_GEMM_FLAGS overrides the default for A/B runs; results identical):
  int flags;
  // optional [32] cycle counters of CTA 0 (VPB_GEMM_DEBUG=1, printed by the next launch), see gemm.cu
  long long* dbg;
};
// ring warp: prefetch the residual chunk of this group's NEXT tile into L2 while chunk c of the current tile loads
constexpr int GEMM_FLAG_PF_RESID = 1;
// TMA producer (CTAs that own n-tile 0): prefetch the A k-block of the NEXT tile's row block into L2
constexpr int GEMM_FLAG_PF_A = 2;
// ring warp: wait until a store has READ its slot right after issuing it and reload the slot at once, instead of one
// store later (one more chunk of load lead per ring)
constexpr int GEMM_FLAG_WAIT0 = 4;
// measured (B200, M = 98304, K = N = 768, L2-cold): WAIT0 209.7 -> 204.7 us; the two L2 prefetches cost 10 % (229 us)
constexpr int GEMM_DEFAULT_FLAGS = GEMM_FLAG_WAIT0;

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;   // 64 bf16 = one 128-byte swizzle row
// bf16 epilogues (bias / GELU) are instruction-heavy: two epilogue warpgroups (8 warps, 2 per scheduler) split
// the column chunks of a tile; the fp32 residual epilogue is memory-heavy and keeps one warpgroup.
__host__ __device__ constexpr bool gemm_epi_bf16(int epi) {
  return epi == EPI_BIAS_BF16 || epi == EPI_GELU_BF16 || epi == EPI_DGELU_BF16 || epi == EPI_GELU_SAVE_BF16 ||
         epi == EPI_BIAS_LNIN_BF16 || epi == EPI_GELU_LNIN_BF16;
}
__host__ __device__ constexpr bool gemm_epi_ln_in(int epi) { return epi == EPI_BIAS_LNIN_BF16 || epi == EPI_GELU_LNIN_BF16; }
__host__ __device__ constexpr bool gemm_epi_gelu(int epi) {
  return epi == EPI_GELU_BF16 || epi == EPI_GELU_SAVE_BF16 || epi == EPI_GELU_LNIN_BF16;
}
__host__ __device__ constexpr int gemm_epi_groups(int epi) { return gemm_epi_bf16(epi) ? 2 : 1; }
// The fp32 residual epilogues add warp 6 (and 11): the ring warp that streams the residual tile through in-place
// staging slots (loads ahead of the epilogue, stores behind it).
// The short-K LayerNorm variants run TWO such epilogue warpgroups (warps 2..5 and 7..10), each with its own
// residual ring and producer warp (6 and 11), on alternating tiles: with one warp per scheduler the epilogue of a
// tile is a latency-bound instruction stream several times longer than the tile's MMA loop.
__host__ __device__ constexpr bool gemm_ln_split(int epi) { return epi == EPI_RESID_LNS_F32 || epi == EPI_POSTMA_LNS_F32; }
__host__ __device__ constexpr int gemm_threads(int epi) {
  return 64 + 128 * gemm_epi_groups(epi) + (gemm_epi_adds_tile(epi) ? 32 : 0) + (gemm_ln_split(epi) ? 160 : 0);
}
constexpr int GEMM_RES_SLOTS = 4;   // upper bound (barrier arrays)
// Dynamic shared memory a kernel may plan with: the 227 KB opt-in limit minus 3 KB for the static arrays (barriers,
// bias / folded-LayerNorm columns). The dynamic array is declared __align__(1024) — ptxas pads the static part to the
// next KB, so the 128B-swizzle tiles need no run-time alignment slack.
constexpr int GEMM_SMEM_BUDGET = 232448 - 3072;
constexpr int GEMM_STAGING_BYTES = GEMM_BM * 128;   // one [128 rows x 128 B] TMA-store box

__host__ __device__ constexpr bool gemm_epi_staged(int epi) {
  return gemm_epi_bf16(epi) || gemm_epi_adds_tile(epi);
}
__host__ __device__ constexpr int gemm_tmem_cols(int bn) {
  return 2 * bn <= 32 ? 32 : 2 * bn <= 64 ? 64 : 2 * bn <= 128 ? 128 : 2 * bn <= 256 ? 256 : 512;
}
// per-CTA bytes of one pipeline stage; with CTA pairs (cg = 2) each CTA stages its 128 A rows and HALF of the B tile
__host__ __device__ constexpr int gemm_stage_bytes(int bn, int cg = 1) { return GEMM_BM * 128 + bn * 128 / cg; }
// Slots of an in-place residual ring. The MMA-bound LayerNorm variant gives one up to keep its operand stages; the
// short-K variants run 3 operand stages and split the rest of shared memory between their two rings.
#ifndef VPB_SPLIT_STAGES
#define VPB_SPLIT_STAGES 3
#endif
__host__ __device__ constexpr int gemm_split_stages(int bn, int cg) { return (cg == 2 && bn == 256) ? VPB_SPLIT_STAGES : 3; }
__host__ __device__ constexpr int gemm_res_slots(int bn, int epi, int cg) {
  if (!gemm_ln_split(epi)) return gemm_epi_ln(epi) ? 3 : 4;
  const int n = (GEMM_SMEM_BUDGET - gemm_split_stages(bn, cg) * gemm_stage_bytes(bn, cg) - 4096) / (2 * GEMM_STAGING_BYTES);
  return n > GEMM_RES_SLOTS ? GEMM_RES_SLOTS : n;
}
// shared memory for the epilogue: one staging box per bf16 epilogue group, or the residual ring (+ gamma/beta)
__host__ __device__ constexpr int gemm_epi_smem(int bn, int epi, int cg) {
  return !gemm_epi_staged(epi) ? 0
         : !gemm_epi_adds_tile(epi)
             ? ((epi == EPI_GELU_SAVE_BF16 || epi == EPI_DGELU_BF16) ? 4 : 2) * GEMM_STAGING_BYTES
             : (gemm_ln_split(epi) ? 2 : 1) * (gemm_res_slots(bn, epi, cg) * GEMM_STAGING_BYTES +
                                               (gemm_epi_ln(epi) ? 2048 : 0));
}
__host__ __device__ constexpr int gemm_num_stages(int bn, int epi, int cg = 1) {
  return gemm_ln_split(epi) ? gemm_split_stages(bn, cg)
         : ((GEMM_SMEM_BUDGET - gemm_epi_smem(bn, epi, cg)) / gemm_stage_bytes(bn, cg)) > 8
             ? 8
             : ((GEMM_SMEM_BUDGET - gemm_epi_smem(bn, epi, cg)) / gemm_stage_bytes(bn, cg));
}
__host__ __device__ constexpr int gemm_smem_bytes(int bn, int epi, int cg = 1) {
  return gemm_num_stages(bn, epi, cg) * gemm_stage_bytes(bn, cg) + gemm_epi_smem(bn, epi, cg);
}
// Decoupled operand rings (CTA pairs, 256-wide tiles): the A tiles (activations, streamed from HBM: ~2 us under load)
// and the B tiles (weights, L2-resident: a fraction of that) get rings of their own depth in the same shared memory,
// few B stages and as many A stages as the rest holds, instead of STAGES x (A + B). The loads in flight are what
// bounds these kernels (VPB_GEMM_DEBUG counters: the MMA issuer waits for operands 29 % (fc1) - 42 % (proj) of the
// time with the producer waiting for free stages).
// Measured (B200, M = 98304, one polling producer thread for both rings): SLOWER — fc1 411 -> 471 us, qkv 274 -> 325,
// fc2 + LN 411 -> 481, proj + LN 205 -> 219: the B tiles stall the issuer as long as the A tiles do (their latency
// under load is not shorter), and the polling thread is slower than two blocking waits. Kept switched off.
#ifndef VPB_GEMM_DEC
#define VPB_GEMM_DEC 0
#endif
#ifndef VPB_GEMM_SB
#define VPB_GEMM_SB 3
#endif
__host__ __device__ constexpr bool gemm_decoupled(int bn, int cg, int opm) {
  return VPB_GEMM_DEC != 0 && cg == 2 && bn == 256 && opm == 0;
}
__host__ __device__ constexpr int gemm_b_stages(int bn, int epi, int cg) {
  return gemm_num_stages(bn, epi, cg) <= 3 ? 2 : VPB_GEMM_SB;
}
__host__ __device__ constexpr int gemm_a_stages(int bn, int epi, int cg) {
  return (gemm_num_stages(bn, epi, cg) * gemm_stage_bytes(bn, cg) - gemm_b_stages(bn, epi, cg) * (bn * 128 / cg)) /
         (GEMM_BM * 128);
}

// Exact-erf GELU (nn.GELU default), gelu(x) = x Phi(x), written as
//   gelu(x) = max(x, 0) - 0.5 |x| erfc(|x| / sqrt 2),      erfc(a / sqrt 2) ~= exp2(q(a)),
// q = the degree-5 polynomial below (no constant term; minimax fit of the ABSOLUTE error of gelu over a in [0, 9],
// tools/fit_gelu.py): |error| <= 7e-7 for every finite input — float rounding level, the same as the Abramowitz &
// Stegun 7.1.26 form the round-1 kernel used — with ONE MUFU op (ex2) per element instead of two (rcp + ex2) and 13
// issue slots per PAIR of elements instead of 22: the fc1 epilogue is what the tile's MMA loop waits for.
// q is strictly decreasing on [0, inf) and -> -inf, so large |x| give erfc = 0 exactly (gelu = max(x, 0)).
#ifndef VPB_GELU_AS
#define VPB_GELU_AS 0      // 1: the Abramowitz & Stegun form (A/B measurements)
#endif
constexpr float GELU_Q1 = -1.1510006189346313f, GELU_Q2 = -0.4595956802368164f, GELU_Q3 = -0.05214685946702957f,
                GELU_Q4 = 0.007198837120085955f, GELU_Q5 = -0.000488122837850824f;

// two elements at once on Blackwell's packed fp32 pipe (FFMA2 / FMUL2 / FADD2); the MUFU ops stay scalar
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
#if VPB_GELU_AS
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 z = __fmul2_rn(ax, make_float2(0.70710678118654752f, 0.70710678118654752f));
  const float2 den = __ffma2_rn(make_float2(0.3275911f, 0.3275911f), z, make_float2(1.0f, 1.0f));
  const float2 t = make_float2(fast_rcp(den.x), fast_rcp(den.y));
  float2 poly = __ffma2_rn(make_float2(1.061405429f, 1.061405429f), t, make_float2(-1.453152027f, -1.453152027f));
  poly = __ffma2_rn(poly, t, make_float2(1.421413741f, 1.421413741f));
  poly = __ffma2_rn(poly, t, make_float2(-0.284496736f, -0.284496736f));
  poly = __ffma2_rn(poly, t, make_float2(0.254829592f, 0.254829592f));
  poly = __fmul2_rn(poly, t);
  const float2 zz = __fmul2_rn(__fmul2_rn(z, make_float2(-1.4426950408889634f, -1.4426950408889634f)), z);
  const float2 e = make_float2(fast_ex2(zz.x), fast_ex2(zz.y));
  const float2 npe = __fmul2_rn(poly, e);
  const float2 erf_abs = make_float2(1.0f - npe.x, 1.0f - npe.y);
  const float2 half_x = __fmul2_rn(x, make_float2(0.5f, 0.5f));
  const float2 erf_s = make_float2(copysignf(erf_abs.x, x.x), copysignf(erf_abs.y, x.y));
  return __ffma2_rn(half_x, erf_s, half_x);
#else
  const float2 a = make_float2(fabsf(x.x), fabsf(x.y));
  float2 q = __ffma2_rn(make_float2(GELU_Q5, GELU_Q5), a, make_float2(GELU_Q4, GELU_Q4));
  q = __ffma2_rn(q, a, make_float2(GELU_Q3, GELU_Q3));
  q = __ffma2_rn(q, a, make_float2(GELU_Q2, GELU_Q2));
  q = __ffma2_rn(q, a, make_float2(GELU_Q1, GELU_Q1));
  q = __fmul2_rn(q, a);
  const float2 e = make_float2(fast_ex2(q.x), fast_ex2(q.y));            // erfc(|x| / sqrt 2)
  const float2 t = __fmul2_rn(__fmul2_rn(a, make_float2(-0.5f, -0.5f)), e);
  return __ffma2_rn(__fadd2_rn(x, a), make_float2(0.5f, 0.5f), t);       // 0.5 (x + |x|) = max(x, 0), exactly
#endif
}

// d/dx [x Phi(x)] = Phi(x) + x phi(x) on two elements, with the same erfc polynomial as gelu_erf2:
// Phi(x) = 0.5 + copysign(0.5 - 0.5 erfc(|x| / sqrt 2), x), phi(x) = exp(-x^2 / 2) / sqrt(2 pi).
__device__ __forceinline__ float2 gelu_erf_grad2(float2 x) {
  const float2 a = make_float2(fabsf(x.x), fabsf(x.y));
  float2 q = __ffma2_rn(make_float2(GELU_Q5, GELU_Q5), a, make_float2(GELU_Q4, GELU_Q4));
  q = __ffma2_rn(q, a, make_float2(GELU_Q3, GELU_Q3));
  q = __ffma2_rn(q, a, make_float2(GELU_Q2, GELU_Q2));
  q = __ffma2_rn(q, a, make_float2(GELU_Q1, GELU_Q1));
  q = __fmul2_rn(q, a);
  const float2 e = make_float2(fast_ex2(q.x), fast_ex2(q.y));
  const float2 h = __ffma2_rn(e, make_float2(-0.5f, -0.5f), make_float2(0.5f, 0.5f));      // 0.5 - 0.5 erfc >= 0
  const float2 cdf = __fadd2_rn(make_float2(copysignf(h.x, x.x), copysignf(h.y, x.y)), make_float2(0.5f, 0.5f));
  const float2 gq = __fmul2_rn(__fmul2_rn(x, x), make_float2(-0.72134752044448170f, -0.72134752044448170f));
  const float2 g = make_float2(fast_ex2(gq.x), fast_ex2(gq.y));                              // exp(-x^2 / 2)
  return __ffma2_rn(__fmul2_rn(x, make_float2(0.3989422804014327f, 0.3989422804014327f)), g, cdf);
}

// Fused-LayerNorm statistics exchange: one 8-byte word per (row, n-tile) = {mean, M2 | tag << 31}; 8-byte accesses
// are single-copy atomic, so the word carries its own validity (the tag of this launch) — no fence, flag or barrier.
__device__ __forceinline__ void ln_publish_stats(const GemmParams& p, int m_blk, int r, int n_tiles, int n_blk,
                                                 float mean, float m2) {
  unsigned long long* part = p.ln_part + static_cast<size_t>(p.ln_epoch & 1u) * p.ln_region +
                             (static_cast<size_t>(m_blk) * GEMM_BM + r) * n_tiles;
  const uint32_t tag = ((p.ln_epoch >> 1) & 1u) << 31;
  st_relaxed_gpu_u64(part + n_blk,
                     (static_cast<unsigned long long>(__float_as_uint(fabsf(m2)) | tag) << 32) | __float_as_uint(mean));
}
// mean / rstd of row r over all N columns from the per-tile words (spins until every n-tile of the row has arrived;
// all CTAs of the persistent grid are resident, siblings run in the same step of the tile loop)
template <int BN>
__device__ __forceinline__ void ln_row_stats(const GemmParams& p, int m_blk, int r, int n_tiles, float& mu,
                                             float& rstd) {
  const unsigned long long* part = p.ln_part + static_cast<size_t>(p.ln_epoch & 1u) * p.ln_region +
                                   (static_cast<size_t>(m_blk) * GEMM_BM + r) * n_tiles;
  const uint32_t tag = ((p.ln_epoch >> 1) & 1u) << 31;
  float pm[10], pq[10];           // n_tiles <= 10 (checked on the host)
  mu = 0.0f;
#pragma unroll
  for (int j = 0; j < 10; ++j) {
    if (j < n_tiles) {
      unsigned long long w = ld_relaxed_gpu_u64(part + j);
      uint32_t spins = 0;
      while ((static_cast<uint32_t>(w >> 32) & 0x80000000u) != tag) {
        if (++spins > (1u << 22)) __trap();   // a sibling CTA never arrived: fail loudly instead of hanging
        __nanosleep(32);                      // 128 threads polling L2 flat out cost power the MMAs need
        w = ld_relaxed_gpu_u64(part + j);
      }
      pm[j] = __uint_as_float(static_cast<uint32_t>(w));
      pq[j] = __uint_as_float(static_cast<uint32_t>(w >> 32) & 0x7fffffffu);
      mu += pm[j];
    }
  }
  mu /= static_cast<float>(n_tiles);
  float var = 0.0f;
#pragma unroll
  for (int j = 0; j < 10; ++j) {
    if (j < n_tiles) {
      const float dm = pm[j] - mu;
      var += pq[j] + static_cast<float>(BN) * dm * dm;
    }
  }
  rstd = rsqrtf(var / static_cast<float>(p.N) + p.ln_eps);
}

// CG = 1: one CTA per 128 x BN tile. CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile —
// each CTA stages its own 128 A rows and half of the B tile, the leader CTA issues M=256 MMAs that read both halves,
// so per-SM shared-memory traffic per MAC drops by a third (1-CTA 128x256 tiles are smem-bandwidth bound:
// 96 B/clk of operand reads + 96 B/clk of TMA writes against 128 B/clk).
// OPM = 1: both operands are given TRANSPOSED, A^T [K, M] and B^T [K, N] row-major (what a weight gradient
// dW = dY^T X has: the contraction runs over the rows of dY and X). Their tiles are TMA-loaded as [64 k-rows] x
// [64-column chunks] and consumed as MN-major UMMA operands, so no transposed copies of the activations are made.
template <int BN, int EPI, int CG, int OPM = 0>
__global__ void __launch_bounds__(gemm_threads(EPI), 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_aux,
                    const __grid_constant__ CUtensorMap tma_ln, const GemmParams p) {
  constexpr int STAGES = gemm_num_stages(BN, EPI, CG);
  constexpr int A_BYTES = GEMM_BM * 128;
  constexpr int B_BYTES = BN * 128 / CG;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr bool DEC = gemm_decoupled(BN, CG, OPM);
  constexpr int SA = DEC ? gemm_a_stages(BN, EPI, CG) : STAGES;     // DEC: A ring [SA][A_BYTES] then B ring [SB][B_BYTES]
  constexpr int SB = DEC ? gemm_b_stages(BN, EPI, CG) : 1;
  static_assert(!DEC || (SA * A_BYTES + SB * B_BYTES <= STAGES * STAGE_BYTES && SA >= 2 && SB >= 2), "operand rings");
  constexpr int TMEM_COLS = gemm_tmem_cols(BN);
  constexpr uint32_t IDESC = umma_idesc_bf16(GEMM_BM * CG, BN, OPM, OPM);
  static_assert(CG == 1 || (CG == 2 && BN % 32 == 0), "CTA pairs split the B tile in two halves");
  static_assert(OPM == 0 || (BN % (64 * CG) == 0), "MN-major operands: 64-column chunks (per CTA of a pair)");
  constexpr int OPM_CHUNK = GEMM_BK * 128;      // [64 k-rows][64 columns] bf16
  constexpr bool STAGED = gemm_epi_staged(EPI);
  constexpr int CHUNK = (gemm_epi_adds_tile(EPI)) ? 32 : 64;     // columns per 128-byte staging row
  constexpr int GROUPS = gemm_epi_groups(EPI);                // epilogue warpgroups
  constexpr bool LN = gemm_epi_ln(EPI);                       // fused LayerNorm of the updated residual rows
  constexpr int RES_SLOTS = gemm_res_slots(BN, EPI, CG);
  constexpr bool LN_SPLIT = gemm_ln_split(EPI);               // two epilogue warpgroups on alternating tiles
  constexpr int TG = LN_SPLIT ? 2 : 1;                        // tile groups (epilogue warpgroup + ring + producer)
  static_assert(RES_SLOTS >= 2, "residual ring too small");
  static_assert(!LN || BN % 64 == 0, "the fused LayerNorm stores 64-column bf16 boxes");
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128 must be a multiple of 16 in [16,256]");
  static_assert(!STAGED || BN % CHUNK == 0, "staged epilogue needs BN to be a multiple of the chunk width");
  static_assert(STAGES >= 2, "pipeline needs at least two stages");

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (threadIdx.x == 0 && (smem_u32(smem_raw) & 1023u) != 0) __trap();   // (the swizzled TMA boxes rely on it)
  // staging boxes of [128 rows x 128 B]: one per epilogue group (bf16), or the in-place residual ring (fp32)
  uint8_t* s_out = smem + STAGES * STAGE_BYTES;
  constexpr int RING_BYTES = RES_SLOTS * GEMM_STAGING_BYTES;  // per tile group
  float* s_affine = reinterpret_cast<float*>(s_out + TG * RING_BYTES);   // LN only, per tile group: gamma[BN] beta[BN]
  __shared__ uint64_t full_bar[SA];      // (A + B of a stage, or the A ring when DEC)
  __shared__ uint64_t empty_bar[SA];
  __shared__ uint64_t fullb_bar[SB];     // DEC: the B ring
  __shared__ uint64_t emptyb_bar[SB];
  __shared__ uint64_t tfull_bar[2];
  __shared__ uint64_t tempty_bar[2];
  __shared__ uint64_t res_full[2][GEMM_RES_SLOTS];
  __shared__ uint64_t res_ready[2][GEMM_RES_SLOTS];   // slot updated in place by the epilogue -> ring warp stores it
  __shared__ uint32_t tmem_slot;
  constexpr int BIAS_PER_GROUP = (BN / CHUNK + GROUPS - 1) / GROUPS * CHUNK;   // columns a group's chunks cover
  __shared__ __align__(16) float s_bias[LN_SPLIT ? 2 : GROUPS][STAGED ? BIAS_PER_GROUP : 1];
  // consumer of a folded LayerNorm: s_n of this group's columns (bf16 epilogues only)
  __shared__ __align__(16) float s_lns[GROUPS][gemm_epi_ln_in(EPI) ? BIAS_PER_GROUP : 1];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + GEMM_BM * CG - 1) / (GEMM_BM * CG);   // (pairs of) 128-row blocks
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles * p.ksplit;
  const int k_blocks = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int kb_per = (k_blocks + p.ksplit - 1) / p.ksplit;   // K blocks per split (the host makes every split non-empty)
  const int cta_rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;     // both CTAs of a pair walk the same tiles

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(&fullb_bar[s], 1);
      mbar_init(&emptyb_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4 * GROUPS * CG);   // the leader's barrier also collects the peer's epilogue warps
    }
    for (int s = 0; s < GEMM_RES_SLOTS; ++s) {
      mbar_init(&res_full[0][s], 1);
      mbar_init(&res_ready[0][s], 1);
      mbar_init(&res_full[1][s], 1);
      mbar_init(&res_ready[1][s], 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (STAGED) tma_prefetch_desc(&tma_out);
    if (gemm_epi_adds_tile(EPI)) tma_prefetch_desc(&tma_aux);
    if (LN) tma_prefetch_desc(&tma_ln);
  }
  if (warp == 1) {
    if constexpr (CG == 2) tmem_alloc_pair(&tmem_slot, TMEM_COLS);
    else tmem_alloc(&tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();   // peer barriers initialised before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();                  // everything above overlapped the previous kernel's tail; its output is visible now
  pdl_launch_dependents();
  // cycle accounting of CTA 0 (one thread per role): time spent in each kind of wait, and the role's total
  const bool timing = p.dbg != nullptr && blockIdx.x == 0;
  auto twait = [&](uint64_t* bar, uint32_t parity, long long& acc) {
    if (timing) {
      const long long t0 = clock64();
      mbar_wait(bar, parity);
      acc += clock64() - t0;
    } else {
      mbar_wait(bar, parity);
    }
  };
  const long long t_role = timing ? clock64() : 0;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long w_empty = 0;
      if constexpr (DEC) {
        // One thread feeds both rings: it polls (test_wait, non-blocking) the next free slot of each and issues
        // whichever load can go, so a full B ring never holds back the A loads further ahead.
        int ta = tile0, ka = 0, tb = tile0, kb_ = 0, sa_i = 0, sb_i = 0;
        uint32_t pa = 0, pb = 0;
        const uint8_t* b_ring = smem + SA * A_BYTES;
        while (ta < num_tiles || tb < num_tiles) {
          bool progressed = false;
          if (ta < num_tiles && mbar_test_wait(&empty_bar[sa_i], pa ^ 1)) {
            const int m_blk = (ta / n_tiles) * CG + cta_rank;
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[sa_i], 2 * A_BYTES);
            tma_load_2d_pair(smem + sa_i * A_BYTES, &tma_a, mapa_shared(smem_u32(&full_bar[sa_i]), 0), ka * GEMM_BK,
                             m_blk * GEMM_BM);
            if (++sa_i == SA) { sa_i = 0; pa ^= 1; }
            if (++ka == k_blocks) { ka = 0; ta += tile_step; }
            progressed = true;
          }
          if (tb < num_tiles && mbar_test_wait(&emptyb_bar[sb_i], pb ^ 1)) {
            const int n_blk = tb % n_tiles;
            if (cta_rank == 0) mbar_arrive_expect_tx(&fullb_bar[sb_i], 2 * B_BYTES);
            tma_load_2d_pair(const_cast<uint8_t*>(b_ring) + sb_i * B_BYTES, &tma_b,
                             mapa_shared(smem_u32(&fullb_bar[sb_i]), 0), kb_ * GEMM_BK, n_blk * BN + cta_rank * (BN / 2));
            if (++sb_i == SB) { sb_i = 0; pb ^= 1; }
            if (++kb_ == k_blocks) { kb_ = 0; tb += tile_step; }
            progressed = true;
          }
          if (!progressed) {
            if (timing) w_empty += 64;
            __nanosleep(32);
          }
        }
      } else
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int m_blk = ((tile / p.ksplit) / n_tiles) * CG + cta_rank;
        const int n_blk = (tile / p.ksplit) % n_tiles;
        const int kb0 = (tile % p.ksplit) * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
        // L2 prefetch of the next tile's A panel (row block changes every tile when the grid is a whole number of row
        // blocks): the CTA that owns n-tile 0 issues it for its siblings too
        const int nt = tile + tile_step;
        const bool pf_a = OPM == 0 && (p.flags & GEMM_FLAG_PF_A) && n_blk == 0 && nt < num_tiles &&
                          (nt / p.ksplit) / n_tiles != (tile / p.ksplit) / n_tiles;
        const int m_next = ((nt / p.ksplit) / n_tiles) * CG + cta_rank;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (pf_a) tma_prefetch_l2_2d(&tma_a, kb * GEMM_BK, m_next * GEMM_BM);
          twait(&empty_bar[stage], phase ^ 1, w_empty);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if constexpr (CG == 2) {
            // both CTAs' bytes are credited to the LEADER's barrier, which the MMA issuer waits on
            const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
            if constexpr (OPM == 1) {     // [64 k-rows][64 columns] chunks: this CTA's 128 A columns, its half of B's
#pragma unroll
              for (int c = 0; c < GEMM_BM / 64; ++c)
                tma_load_2d_pair(sa + c * OPM_CHUNK, &tma_a, leader_full, m_blk * GEMM_BM + c * 64, kb * GEMM_BK);
#pragma unroll
              for (int c = 0; c < BN / 2 / 64; ++c)
                tma_load_2d_pair(sb + c * OPM_CHUNK, &tma_b, leader_full, n_blk * BN + cta_rank * (BN / 2) + c * 64,
                                 kb * GEMM_BK);
            } else {
            tma_load_2d_pair(sa, &tma_a, leader_full, kb * GEMM_BK, m_blk * GEMM_BM);
            tma_load_2d_pair(sb, &tma_b, leader_full, kb * GEMM_BK, n_blk * BN + cta_rank * (BN / 2));
            }
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            if constexpr (OPM == 1) {
#pragma unroll
              for (int c = 0; c < GEMM_BM / 64; ++c)
                tma_load_2d(sa + c * OPM_CHUNK, &tma_a, &full_bar[stage], m_blk * GEMM_BM + c * 64, kb * GEMM_BK);
#pragma unroll
              for (int c = 0; c < BN / 64; ++c)
                tma_load_2d(sb + c * OPM_CHUNK, &tma_b, &full_bar[stage], n_blk * BN + c * 64, kb * GEMM_BK);
            } else {
              tma_load_2d(sa, &tma_a, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
              tma_load_2d(sb, &tma_b, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (timing) { p.dbg[0] = w_empty; p.dbg[1] = clock64() - t_role; }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta_rank == 0) {      // with CTA pairs only the leader issues MMAs (for both CTAs)
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      long long w_tempty = 0, w_full = 0, w_fullb = 0;
      int sb_i = 0;
      uint32_t pb = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        twait(&tempty_bar[acc], acc_phase ^ 1, w_tempty);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        if constexpr (DEC) {
          for (int kb = 0; kb < k_blocks; ++kb) {
            twait(&full_bar[stage], phase, w_full);
            twait(&fullb_bar[sb_i], pb, w_fullb);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + stage * A_BYTES);
            const uint32_t b_addr = smem_u32(smem + SA * A_BYTES + sb_i * B_BYTES);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              umma_bf16_ss_pair(d_tmem, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), IDESC,
                                (kb > 0 || k != 0) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage], 3);
            umma_commit_pair(&emptyb_bar[sb_i], 3);
            if (++stage == SA) { stage = 0; phase ^= 1; }
            if (++sb_i == SB) { sb_i = 0; pb ^= 1; }
          }
        } else {
        const int kb0 = (tile % p.ksplit) * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
        for (int kb = kb0; kb < kb1; ++kb) {
          twait(&full_bar[stage], phase, w_full);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            if constexpr (CG == 2 && OPM == 1)
              umma_bf16_ss_pair(d_tmem, umma_desc_mn_sw128(a_addr + k * 2048, OPM_CHUNK),
                                umma_desc_mn_sw128(b_addr + k * 2048, OPM_CHUNK), IDESC, (kb > kb0 || k != 0) ? 1u : 0u);
            else if constexpr (CG == 2)
              umma_bf16_ss_pair(d_tmem, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), IDESC,
                                (kb > kb0 || k != 0) ? 1u : 0u);
            else if constexpr (OPM == 1)     // 16 k-rows = 2048 B per step; 64-column chunks OPM_CHUNK apart
              umma_bf16_ss(d_tmem, umma_desc_mn_sw128(a_addr + k * 2048, OPM_CHUNK),
                           umma_desc_mn_sw128(b_addr + k * 2048, OPM_CHUNK), IDESC, (kb > kb0 || k != 0) ? 1u : 0u);
            else
              umma_bf16_ss(d_tmem, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), IDESC,
                           (kb > kb0 || k != 0) ? 1u : 0u);
          }
          // frees this smem stage (in both CTAs of a pair) once the MMAs above have read it
          if constexpr (CG == 2) umma_commit_pair(&empty_bar[stage], 3);
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (CG == 2) umma_commit_pair(&tfull_bar[acc], 3);
        else umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (timing) { p.dbg[2] = w_tempty; p.dbg[3] = w_full; p.dbg[4] = clock64() - t_role; p.dbg[14] = w_fullb; }
    }
  } else if (gemm_epi_adds_tile(EPI) && (warp == 6 || (LN_SPLIT && warp == 11))) {
    // Ring warp of tile group tg: owns BOTH ends of the in-place staging ring. It TMA-loads the fp32 residual (or
    // positional) chunk c of tile (m_blk, n_blk) = [128 rows x 32 cols] into a slot, the epilogue threads update the
    // slot in place and flag it ready, and this warp TMA-stores it and recycles the slot once the store has read it —
    // so no epilogue thread ever waits for a store. With the fused LayerNorm the tile's bf16 boxes (64 columns) follow
    // its fp32 chunks through the same slots (nothing to load for those).
    if (lane == 0) {
      const int tg = warp == 11 ? 1 : 0;
      uint8_t* ring = s_out + tg * RING_BYTES;
      constexpr int NX = BN / 32, NCH = NX + (LN ? BN / 64 : 0);     // chunks per tile
      const int first_tile = tile0 + tg * tile_step, stride = tile_step * TG;
      // Order of a tile's chunks through the ring. Fused LayerNorm: the NX fp32 chunks, then the BN / 64 bf16 boxes
      // (pass 2). Folded LayerNorm: the bf16 box of 64 columns right after its two fp32 chunks.
      // (the folded form exists in the one-group variants only: the two-group kernels stay free of its registers)
      const bool fold = LN && !LN_SPLIT && p.ln_fold != 0;
      auto seq_is_box = [&](int q) { return LN && (fold ? (q % 3) == 2 : q >= NX); };
      auto seq_index = [&](int q) { return fold ? ((q % 3) == 2 ? q / 3 : (q / 3) * 2 + (q % 3)) : (q >= NX ? q - NX : q); };
      auto load = [&](int tile, int q, uint32_t slot) {
        if (seq_is_box(q)) {                // bf16 box: the slot only has to be free
          mbar_arrive(&res_full[tg][slot]);
          return;
        }
        const int c = seq_index(q);
        const int m_blk = ((tile / p.ksplit) / n_tiles) * CG + cta_rank;
        const int n_blk = (tile / p.ksplit) % n_tiles;
        mbar_arrive_expect_tx(&res_full[tg][slot], GEMM_STAGING_BYTES);
        if constexpr (!gemm_epi_pos(EPI)) {      // (the positional table is L2-resident anyway)
          const int nt = tile + stride;
          if ((p.flags & GEMM_FLAG_PF_RESID) && nt < num_tiles)
            tma_prefetch_l2_2d(&tma_aux, ((nt / p.ksplit) % n_tiles) * BN + c * 32,
                               (((nt / p.ksplit) / n_tiles) * CG + cta_rank) * GEMM_BM);
        }
        if constexpr (gemm_epi_pos(EPI)) {
          // positional table rows (token = row % period): two 64-row boxes, each inside one period (period % 64 == 0)
          const int t0 = (m_blk * GEMM_BM) % p.period, t1 = (m_blk * GEMM_BM + 64) % p.period;
          tma_load_2d(ring + slot * GEMM_STAGING_BYTES, &tma_aux, &res_full[tg][slot], n_blk * BN + c * 32, t0);
          tma_load_2d(ring + slot * GEMM_STAGING_BYTES + 64 * 128, &tma_aux, &res_full[tg][slot], n_blk * BN + c * 32,
                      t1);
        } else {
          tma_load_2d(ring + slot * GEMM_STAGING_BYTES, &tma_aux, &res_full[tg][slot], n_blk * BN + c * 32,
                      m_blk * GEMM_BM);
        }
      };
      int ld_tile = first_tile, ld_c = 0, st_tile = first_tile, st_c = 0;
      uint32_t nld = 0, nst = 0;
      long long w_ready = 0, w_read = 0;
      auto load_next = [&](uint32_t slot) {
        if (ld_tile >= num_tiles) return;
        load(ld_tile, ld_c, slot);
        ++nld;
        if (++ld_c == NCH) { ld_c = 0; ld_tile += stride; }
      };
      for (int i = 0; i < RES_SLOTS; ++i) load_next(static_cast<uint32_t>(i));
      while (st_tile < num_tiles) {
        const uint32_t slot = nst % RES_SLOTS;
        twait(&res_ready[tg][slot], (nst / RES_SLOTS) & 1, w_ready);  // the epilogue has finished chunk nst in place
        const int m_blk = ((st_tile / p.ksplit) / n_tiles) * CG + cta_rank;
        const int n_blk = (st_tile / p.ksplit) % n_tiles;
        if (!seq_is_box(st_c))
          tma_store_2d(&tma_out, ring + slot * GEMM_STAGING_BYTES, n_blk * BN + seq_index(st_c) * 32, m_blk * GEMM_BM);
        else
          tma_store_2d(&tma_ln, ring + slot * GEMM_STAGING_BYTES, n_blk * BN + seq_index(st_c) * 64, m_blk * GEMM_BM);
        tma_store_commit();
        if (p.flags & GEMM_FLAG_WAIT0) {    // this store has read its slot (a few hundred cycles): reload it now
          const long long t0 = timing ? clock64() : 0;
          tma_store_wait_read<0>();
          if (timing) w_read += clock64() - t0;
          load_next(slot);
        } else if (nst > 0) {               // the previous store has read its slot: reuse it for the next load
          tma_store_wait_read<1>();
          load_next((nst - 1) % RES_SLOTS);
        }
        ++nst;
        if (++st_c == NCH) { st_c = 0; st_tile += stride; }
      }
      tma_store_wait_all<0>();              // all output bytes committed before the CTA exits
      if (timing && tg == 0) { p.dbg[5] = w_ready; p.dbg[6] = w_read; p.dbg[7] = clock64() - t_role; }
    }
  } else {
    const int quad = warp & 3;             // TMEM lane quadrant this warp may read
    // epilogue warpgroup: bf16 epilogues split the column chunks of every tile between two groups (cgrp), the
    // short-K LayerNorm epilogue gives every other tile to each of its two groups
    const int grp = LN_SPLIT ? (warp >= 7 ? 1 : 0) : ((warp - 2) >> 2);
    const int cgrp = LN_SPLIT ? 0 : grp;
    const int etid = LN_SPLIT ? static_cast<int>(threadIdx.x) - (grp ? 224 : 64)
                              : static_cast<int>(threadIdx.x) - 64 - 128 * grp;   // 0..127 inside the warpgroup
    uint8_t* ring = s_out + (LN_SPLIT ? grp : 0) * RING_BYTES;
    uint64_t* rfull = res_full[LN_SPLIT ? grp : 0];
    uint64_t* rready = res_ready[LN_SPLIT ? grp : 0];
    float* s_gamma = s_affine + (LN_SPLIT ? grp : 0) * 2 * BN;
    float* s_beta = s_gamma + BN;
    const int r = quad * 32 + lane;        // row inside the tile == TMEM lane
    const int bar_id = 1 + grp;            // named barrier of this warpgroup
    auto arrive_tempty = [&](int a) {      // TMEM buffer drained: tell the MMA issuer (in the pair's leader CTA)
      if constexpr (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[a]), 0));
      else mbar_arrive(&tempty_bar[a]);
    };
    int acc = LN_SPLIT ? grp : 0;          // tile group g always finds its tiles in TMEM buffer g
    uint32_t acc_phase = 0;
    uint32_t chunk_seq = 0;                // running chunk counter of this group: staging buffer = chunk_seq & 1
    int const_n_blk = -1;                  // n-tile whose bias / gamma / beta columns are in shared memory
    long long w_tfull = 0, w_rfull = 0, w_sib = 0, w_bar = 0, n_epi_tiles = 0;
    // EPI_DGELU_BF16: the saved pre-activation of a chunk ([128 rows x 64 columns] bf16) arrives by TMA in the group's
    // second staging box (2 + grp), one chunk ahead: requested by the group's first thread as soon as all its threads
    // have read the previous one, on the (otherwise unused) barrier res_full[grp][0].
    uint32_t pre_phase = 0;
    auto request_pre = [&](int tile_, int c_) {
      const int mb = ((tile_ / p.ksplit) / n_tiles) * CG + cta_rank, nb = (tile_ / p.ksplit) % n_tiles;
      mbar_arrive_expect_tx(&res_full[grp][0], GEMM_STAGING_BYTES);
      tma_load_2d(s_out + (2 + grp) * GEMM_STAGING_BYTES, &tma_aux, &res_full[grp][0], nb * BN + c_ * 64, mb * GEMM_BM);
    };
    if constexpr (EPI == EPI_DGELU_BF16) {
      if (etid == 0 && tile0 < num_tiles && cgrp < BN / 64) request_pre(tile0, cgrp);
    }
    for (int tile = tile0 + (LN_SPLIT ? grp : 0) * tile_step; tile < num_tiles; tile += tile_step * TG) {
      const int m_blk = ((tile / p.ksplit) / n_tiles) * CG + cta_rank;
      const int n_blk = (tile / p.ksplit) % n_tiles;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
      float rs = 1.0f;                     // per-row scale of the branch (stochastic depth), rows past M clamp
      if (gemm_epi_adds_tile(EPI) && p.row_scale != nullptr)
        rs = __ldg(p.row_scale + min(m_blk * GEMM_BM + r, p.M - 1) / p.scale_period);
      const float2 rs2 = make_float2(rs, rs);
      (void)rs2;

      if constexpr (STAGED) {
        constexpr int NCHUNK = BN / CHUNK;
        // bias (and LayerNorm affine) columns of this n-tile -> smem, only when the n-tile changes (it never does when
        // the grid is a whole number of row blocks). All readers of the previous values are past a barrier.
        // (each group keeps only the columns of its own chunks: local chunk lc <-> tile chunk lc*GROUPS + grp)
        if (n_blk != const_n_blk) {
          const_n_blk = n_blk;
          for (int i = etid; i < BIAS_PER_GROUP; i += 128) {
            const int col = n_blk * BN + ((i / CHUNK) * GROUPS + cgrp) * CHUNK + (i % CHUNK);
            s_bias[grp][i] = (p.bias != nullptr && col < p.N) ? __ldg(p.bias + col) : 0.0f;
            if constexpr (gemm_epi_ln_in(EPI)) s_lns[grp][i] = col < p.N ? __ldg(p.ln_s + col) : 0.0f;
          }
          if constexpr (LN) {
            if (LN_SPLIT || p.ln_fold == 0)     // (the folded form leaves gamma / beta to the consumer's weights)
            for (int i = etid; i < BN; i += 128) {
              s_gamma[i] = __ldg(p.ln_gamma + n_blk * BN + i);
              s_beta[i] = __ldg(p.ln_beta + n_blk * BN + i);
            }
          }
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // constants visible to the whole group
        }
        // ---- consumer side of a folded LayerNorm: (mean, M2) pairs of this thread's row of A, merged over the producer's
        // column tiles. Requested BEFORE the wait for the accumulator so that the L2 round trip hides behind it.
        float ln_rstd = 1.0f, ln_nmr = 0.0f;
        constexpr bool ln_in = gemm_epi_ln_in(EPI);
        if constexpr (ln_in) {
          const int row = min(m_blk * GEMM_BM + r, p.M - 1);
          const float2* st = p.ln_stats + static_cast<size_t>(row) * p.ln_parts;
          float pm[10], pq[10], mu = 0.0f;
#pragma unroll
          for (int j = 0; j < 10; ++j)
            if (j < p.ln_parts) {
              const float2 w = __ldg(st + j);
              pm[j] = w.x;
              pq[j] = w.y;
              mu += w.x;
            }
          mu /= static_cast<float>(p.ln_parts);
          float var = 0.0f;
#pragma unroll
          for (int j = 0; j < 10; ++j)
            if (j < p.ln_parts) {
              const float dm = pm[j] - mu;
              var += pq[j] + static_cast<float>(p.ln_part_cols) * dm * dm;
            }
          ln_rstd = rsqrtf(var / static_cast<float>(p.ln_parts * p.ln_part_cols) + p.ln_eps);
          ln_nmr = -mu * ln_rstd;
        }
        twait(&tfull_bar[acc], acc_phase, w_tfull);
        ++n_epi_tiles;
        tc_fence_after();
        if (cgrp >= NCHUNK) {               // narrow tile: this group has no chunk, just release TMEM
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_tempty(acc);
        }
        if constexpr (LN) {
          if (!LN_SPLIT && p.ln_fold != 0) {
            // ---- folded LayerNorm: ONE pass. x' = residual + acc + bias goes out as fp32 through the in-place ring
            // and as bf16 (64-column boxes through the same ring, right after their two fp32 chunks); the row's
            // (mean, M2) over this tile's columns is left for the consumer GEMM. TMEM is free after the last load.
            float mean = 0.0f, m2 = 0.0f;
            uint32_t va[32], vb[32], pk[32];
            auto fold_chunk = [&](uint32_t(&v)[32], int c, uint32_t* pkh) {
              const uint32_t buf = chunk_seq % RES_SLOTS;
              const uint32_t srow = smem_u32(ring) + buf * GEMM_STAGING_BYTES + r * 128;
              const float4* bias4 = reinterpret_cast<const float4*>(&s_bias[grp][c * 32]);
              twait(&rfull[buf], (chunk_seq / RES_SLOTS) & 1, w_rfull);
              float2 s0 = make_float2(0.0f, 0.0f), s1 = make_float2(0.0f, 0.0f);
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const uint32_t pu = srow + ((u ^ (r & 7)) * 16);
                const float4 x = lds_f4(pu);
                const float4 bb = bias4[u];
                const float2 lo = __ffma2_rn(rs2,
                                             __fadd2_rn(make_float2(__uint_as_float(v[4 * u]), __uint_as_float(v[4 * u + 1])),
                                                        make_float2(bb.x, bb.y)),
                                             make_float2(x.x, x.y));
                const float2 hi = __ffma2_rn(rs2,
                                             __fadd2_rn(make_float2(__uint_as_float(v[4 * u + 2]), __uint_as_float(v[4 * u + 3])),
                                                        make_float2(bb.z, bb.w)),
                                             make_float2(x.z, x.w));
                sts_f4(pu, make_float4(lo.x, lo.y, hi.x, hi.y));
                v[4 * u + 0] = __float_as_uint(lo.x);
                v[4 * u + 1] = __float_as_uint(lo.y);
                v[4 * u + 2] = __float_as_uint(hi.x);
                v[4 * u + 3] = __float_as_uint(hi.y);
                pkh[2 * u] = pack_bf16x2(lo.x, lo.y);
                pkh[2 * u + 1] = pack_bf16x2(hi.x, hi.y);
                s0 = __fadd2_rn(s0, lo);
                s1 = __fadd2_rn(s1, hi);
              }
              fence_proxy_async_smem();
              asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
              if (etid == 0) mbar_arrive(&rready[buf]);
              const float2 ss = __fadd2_rn(s0, s1);
              const float mc = (ss.x + ss.y) * (1.0f / 32.0f);
              const float2 nmc = make_float2(-mc, -mc);
              float2 q0 = make_float2(0.0f, 0.0f), q1 = make_float2(0.0f, 0.0f);
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float2 d0 = __fadd2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), nmc);
                const float2 d1 = __fadd2_rn(make_float2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), nmc);
                q0 = __ffma2_rn(d0, d0, q0);
                q1 = __ffma2_rn(d1, d1, q1);
              }
              const float2 qq = __fadd2_rn(q0, q1);
              const float delta = mc - mean;
              const float wgt = 1.0f / static_cast<float>(c + 1);
              mean = fmaf(delta, wgt, mean);
              m2 += (qq.x + qq.y) + delta * delta * (32.0f * static_cast<float>(c) * wgt);
              ++chunk_seq;
            };
            tmem_ld_32x32b_x32(t_row, va);
#pragma unroll 1
            for (int c = 0; c < NCHUNK; c += 2) {
              tmem_ld_wait();
              tmem_ld_32x32b_x32(t_row + (c + 1) * 32, vb);
              fold_chunk(va, c, &pk[0]);
              tmem_ld_wait();
              if (c + 2 < NCHUNK) {
                tmem_ld_32x32b_x32(t_row + (c + 2) * 32, va);
              } else {                              // the whole tile is in registers: hand TMEM back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_tempty(acc);
              }
              fold_chunk(vb, c + 1, &pk[16]);
              // bf16 copy of the 64 columns just finished (same swizzled box layout as the normalised rows of pass 2)
              const uint32_t buf = chunk_seq % RES_SLOTS;
              const uint32_t srow = smem_u32(ring) + buf * GEMM_STAGING_BYTES + r * 128;
              twait(&rfull[buf], (chunk_seq / RES_SLOTS) & 1, w_bar);
#pragma unroll
              for (int u = 0; u < 8; ++u)
                sts_u4(srow + ((u ^ (r & 7)) * 16), pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
              fence_proxy_async_smem();
              asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
              if (etid == 0) mbar_arrive(&rready[buf]);
              ++chunk_seq;
            }
            p.ln_stats_out[(static_cast<size_t>(m_blk) * GEMM_BM + r) * n_tiles + n_blk] = make_float2(mean, m2);
          } else {
          // ---- pass 1: x' = residual + acc + bias. Stored as fp32 through the in-place ring exactly like
          // EPI_RESID_F32, written back into the TMEM accumulator for pass 2, and reduced to a running (mean, M2)
          // of this row over the tile's BN columns (two-pass inside a 32-column chunk, Chan's update across chunks).
          float mean = 0.0f, m2 = 0.0f;
          // one 32-column chunk (its accumulator values already in v)
          auto pass1_chunk = [&](uint32_t(&v)[32], int c) {
            const uint32_t buf = chunk_seq % RES_SLOTS;
            const uint32_t srow = smem_u32(ring) + buf * GEMM_STAGING_BYTES + r * 128;
            const float4* bias4 = reinterpret_cast<const float4*>(&s_bias[grp][c * 32]);
            twait(&rfull[buf], (chunk_seq / RES_SLOTS) & 1, w_rfull);
            float2 s0 = make_float2(0.0f, 0.0f), s1 = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint32_t pu = srow + ((u ^ (r & 7)) * 16);
              const float4 x = lds_f4(pu);
              const float4 bb = bias4[u];
              // residual + (acc + bias), as the unfused epilogue; packed fp32 adds are IEEE per lane
              const float2 lo = __ffma2_rn(rs2,
                                           __fadd2_rn(make_float2(__uint_as_float(v[4 * u]), __uint_as_float(v[4 * u + 1])),
                                                      make_float2(bb.x, bb.y)),
                                           make_float2(x.x, x.y));
              const float2 hi = __ffma2_rn(rs2,
                                           __fadd2_rn(make_float2(__uint_as_float(v[4 * u + 2]), __uint_as_float(v[4 * u + 3])),
                                                      make_float2(bb.z, bb.w)),
                                           make_float2(x.z, x.w));
              sts_f4(pu, make_float4(lo.x, lo.y, hi.x, hi.y));
              v[4 * u + 0] = __float_as_uint(lo.x);
              v[4 * u + 1] = __float_as_uint(lo.y);
              v[4 * u + 2] = __float_as_uint(hi.x);
              v[4 * u + 3] = __float_as_uint(hi.y);
              s0 = __fadd2_rn(s0, lo);
              s1 = __fadd2_rn(s1, hi);
            }
            tmem_st_32x32b_x32(t_row + c * 32, v);
            // the slot is complete: publish it to the async proxy and let thread 0 store it while the others go on
            fence_proxy_async_smem();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            if (etid == 0) mbar_arrive(&rready[buf]);      // the ring warp stores the slot and recycles it
            const float2 ss = __fadd2_rn(s0, s1);
            const float mc = (ss.x + ss.y) * (1.0f / 32.0f);
            const float2 nmc = make_float2(-mc, -mc);
            float2 q0 = make_float2(0.0f, 0.0f), q1 = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float2 d0 = __fadd2_rn(make_float2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), nmc);
              const float2 d1 = __fadd2_rn(make_float2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])), nmc);
              q0 = __ffma2_rn(d0, d0, q0);
              q1 = __ffma2_rn(d1, d1, q1);
            }
            const float2 qq = __fadd2_rn(q0, q1);
            const float delta = mc - mean;
            const float wgt = 1.0f / static_cast<float>(c + 1);          // n_b / (n_a + n_b), n_a = 32 c, n_b = 32
            mean = fmaf(delta, wgt, mean);
            m2 += (qq.x + qq.y) + delta * delta * (32.0f * static_cast<float>(c) * wgt);
            ++chunk_seq;
          };
          {
            // two register sets: the TMEM load of the next chunk is in flight while this one is processed
            uint32_t va[32], vb[32];
            tmem_ld_32x32b_x32(t_row, va);
#pragma unroll 1
            for (int c = 0; c < NCHUNK; c += 2) {
              tmem_ld_wait();
              tmem_ld_32x32b_x32(t_row + (c + 1) * 32, vb);
              pass1_chunk(va, c);
              tmem_ld_wait();
              if (c + 2 < NCHUNK) tmem_ld_32x32b_x32(t_row + (c + 2) * 32, va);
              pass1_chunk(vb, c + 1);
            }
          }
          tmem_st_wait();
          // ---- publish (mean, M2) for the CTAs that own the other n-tiles of this row block (and for pass 2)
          ln_publish_stats(p, m_blk, r, n_tiles, n_blk, mean, m2);
          float mu, rstd;
          const long long t_sib = timing ? clock64() : 0;
          ln_row_stats<BN>(p, m_blk, r, n_tiles, mu, rstd);
          if (timing) w_sib += clock64() - t_sib;
          const float nmr = -mu * rstd;
          // ---- pass 2: normalise the rows kept in TMEM, bf16 boxes of 64 columns through the same ring
#pragma unroll 1
          for (int c = 0; c < BN / 64; ++c, ++chunk_seq) {
            const uint32_t buf = chunk_seq % RES_SLOTS;
            uint32_t v[64];
            tmem_ld_32x32b_x32(t_row + c * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
            tmem_ld_32x32b_x32(t_row + c * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
            tmem_ld_wait();
            if (c == BN / 64 - 1) {               // the tile has left TMEM: hand the buffer back to the MMA warp
              tc_fence_before();
              __syncwarp();
              if (lane == 0) arrive_tempty(acc);
            }
            const float4* g4 = reinterpret_cast<const float4*>(s_gamma + c * 64);
            const float4* b4 = reinterpret_cast<const float4*>(s_beta + c * 64);
            const float2 rs2 = make_float2(rstd, rstd), nm2 = make_float2(nmr, nmr);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float4 gg = g4[j], bb = b4[j];
              const float2 y0 = __ffma2_rn(
                  __ffma2_rn(make_float2(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1])), rs2, nm2),
                  make_float2(gg.x, gg.y), make_float2(bb.x, bb.y));
              const float2 y1 = __ffma2_rn(
                  __ffma2_rn(make_float2(__uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3])), rs2, nm2),
                  make_float2(gg.z, gg.w), make_float2(bb.z, bb.w));
              v[2 * j] = pack_bf16x2(y0.x, y0.y);
              v[2 * j + 1] = pack_bf16x2(y1.x, y1.y);
            }
            const uint32_t srow = smem_u32(ring) + buf * GEMM_STAGING_BYTES + r * 128;
            twait(&rfull[buf], (chunk_seq / RES_SLOTS) & 1, w_bar);
#pragma unroll
            for (int u = 0; u < 8; ++u)
              sts_u4(srow + ((u ^ (r & 7)) * 16), v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
            fence_proxy_async_smem();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            if (etid == 0) mbar_arrive(&rready[buf]);
          }
          }
        } else {
        const float2 ln_r2 = make_float2(ln_rstd, ln_rstd), ln_m2 = make_float2(ln_nmr, ln_nmr);
        (void)ln_r2; (void)ln_m2;
#pragma unroll 1
        for (int c = cgrp; c < NCHUNK; c += GROUPS, ++chunk_seq) {
          // staging slot: bf16 epilogues own one box per group; the residual epilogue walks the in-place ring
          const uint32_t buf = gemm_epi_adds_tile(EPI) ? (chunk_seq % RES_SLOTS) : static_cast<uint32_t>(grp);
          uint32_t v[CHUNK];
          if constexpr (CHUNK == 64) {
            tmem_ld_32x32b_x32(t_row + c * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
            tmem_ld_32x32b_x32(t_row + c * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          } else {
            tmem_ld_32x32b_x32(t_row + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          }
          // training epilogues address global memory directly: this thread's row, this chunk's 64 columns (128 bytes)
          const int g_row = m_blk * GEMM_BM + r, g_col = n_blk * BN + c * CHUNK;
          uint4 pre8[EPI == EPI_DGELU_BF16 ? CHUNK / 8 : 1];
          if constexpr (EPI == EPI_DGELU_BF16) {      // this thread's row of the pre-activation box (zero past M / N)
            mbar_wait(&res_full[grp][0], pre_phase);
            pre_phase ^= 1u;
            const uint32_t prow = smem_u32(s_out) + (2 + grp) * GEMM_STAGING_BYTES + r * 128;
#pragma unroll
            for (int u = 0; u < CHUNK / 8; ++u) pre8[u] = lds_u4(prow + ((u ^ (r & 7)) * 16));
          }
          tmem_ld_wait();
          if (c + GROUPS >= NCHUNK) {       // this group's last chunk is in registers: hand TMEM back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_tempty(acc);
          }
          const uint32_t srow = smem_u32(ring) + buf * GEMM_STAGING_BYTES + r * 128;
          const float* bias_c = &s_bias[grp][(c / GROUPS) * CHUNK];
          if constexpr (gemm_epi_adds_tile(EPI)) {
            // residual chunk landed in the slot (TMA, issued by warp 6 well ahead); update it in place: every thread
            // reads and writes only its own 128-byte row, so no barrier is needed before the math
            twait(&rfull[buf], (chunk_seq / RES_SLOTS) & 1, w_rfull);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const uint32_t pu = srow + ((u ^ (r & 7)) * 16);
              float4 x = lds_f4(pu);
              // fmaf(1, t, x) == x + t exactly, so the unscaled path keeps its rounding
              x.x = fmaf(rs, __uint_as_float(v[4 * u + 0]) + bias_c[4 * u + 0], x.x);
              x.y = fmaf(rs, __uint_as_float(v[4 * u + 1]) + bias_c[4 * u + 1], x.y);
              x.z = fmaf(rs, __uint_as_float(v[4 * u + 2]) + bias_c[4 * u + 2], x.z);
              x.w = fmaf(rs, __uint_as_float(v[4 * u + 3]) + bias_c[4 * u + 3], x.w);
              sts_f4(pu, x);
            }
          } else {
            // all the math first, into packed registers (overwriting v) ...
            if constexpr (ln_in) {      // acc * rstd - mean * rstd * s + c   (folded LayerNorm of the A rows; `bias` holds c)
              const float* lns_c = &s_lns[grp][(c / GROUPS) * CHUNK];
#pragma unroll
              for (int j = 0; j < CHUNK / 2; ++j) {
                const float2 b2 = *reinterpret_cast<const float2*>(bias_c + 2 * j);
                const float2 s2 = *reinterpret_cast<const float2*>(lns_c + 2 * j);
                float2 f = __ffma2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), ln_r2,
                                      __ffma2_rn(ln_m2, s2, b2));
                if constexpr (gemm_epi_gelu(EPI)) f = gelu_erf2(f);
                v[j] = pack_bf16x2(f.x, f.y);
              }
            } else if constexpr (EPI == EPI_DGELU_BF16) {
              // acc * gelu'(pre); 32 columns at a time so that their fp32 values can be column-summed across the warp
              const uint32_t* prw = reinterpret_cast<const uint32_t*>(pre8);
#pragma unroll
              for (int hh = 0; hh < CHUNK / 32; ++hh) {
                float cs[32];
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                  const int j = hh * 16 + jj;
                  const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&prw[j]));
                  const float2 f = __fmul2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])),
                                              gelu_erf_grad2(x));
                  v[j] = pack_bf16x2(f.x, f.y);
                  cs[2 * jj] = f.x;
                  cs[2 * jj + 1] = f.y;
                }
                if (p.colsum_out != nullptr) {        // (rows past M hold acc = 0: A is zero-filled there)
                  const float sum = warp_colsum32(cs, lane);
                  const int col = g_col + hh * 32 + lane;
                  if (col < p.N) atomicAdd(p.colsum_out + col, sum);
                }
              }
            } else if constexpr (EPI == EPI_GELU_SAVE_BF16) {
              // training forward of fc1: the pre-activation the backward pass differentiates goes out through a second
              // staging box per group (boxes 2 + grp) and its own TMA store, then the activation as usual
              uint32_t pw[CHUNK / 2];
#pragma unroll
              for (int j = 0; j < CHUNK / 2; ++j) {
                const float2 b2 = *reinterpret_cast<const float2*>(bias_c + 2 * j);
                const float2 f = __fadd2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), b2);
                pw[j] = pack_bf16x2(f.x, f.y);
              }
              if (etid == 0) tma_store_wait_read<0>();       // both boxes of this group: their last stores have read them
              asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
              const uint32_t prow = smem_u32(ring) + (2 + grp) * GEMM_STAGING_BYTES + r * 128;
#pragma unroll
              for (int u = 0; u < 8; ++u)
                sts_u4(prow + ((u ^ (r & 7)) * 16), pw[4 * u], pw[4 * u + 1], pw[4 * u + 2], pw[4 * u + 3]);
#pragma unroll
              for (int j = 0; j < CHUNK / 2; ++j) {
                const float2 b2 = *reinterpret_cast<const float2*>(bias_c + 2 * j);
                const float2 f = gelu_erf2(
                    __fadd2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), b2));
                v[j] = pack_bf16x2(f.x, f.y);
              }
            } else {
#pragma unroll
            for (int j = 0; j < CHUNK / 2; ++j) {
              const float2 b2 = *reinterpret_cast<const float2*>(bias_c + 2 * j);
              float2 f = __fadd2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), b2);
              if constexpr (gemm_epi_gelu(EPI)) f = gelu_erf2(f);
              v[j] = pack_bf16x2(f.x, f.y);
            }
            }
            // ... then wait until the TMA store that last read this group's staging box has finished reading it
            // (that latency is now hidden behind the math), and write the row
            if constexpr (EPI != EPI_GELU_SAVE_BF16) {
              if (etid == 0) tma_store_wait_read<0>();
              asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            }
            if constexpr (EPI == EPI_DGELU_BF16) {     // every thread of the group has its pre-activation row: next box
              if (etid == 0) {
                if (c + GROUPS < NCHUNK) request_pre(tile, c + GROUPS);
                else if (tile + tile_step * TG < num_tiles) request_pre(tile + tile_step * TG, cgrp);
              }
            }
#pragma unroll
            for (int u = 0; u < 8; ++u)
              sts_u4(srow + ((u ^ (r & 7)) * 16), v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
          }
          fence_proxy_async_smem();         // generic-proxy smem writes -> visible to the TMA store
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (etid == 0) {
            if constexpr (gemm_epi_adds_tile(EPI)) {
              mbar_arrive(&rready[buf]);                   // the ring warp stores the slot and recycles it
            } else {
              tma_store_2d(&tma_out, ring + buf * GEMM_STAGING_BYTES, n_blk * BN + c * CHUNK, m_blk * GEMM_BM);
              if constexpr (EPI == EPI_GELU_SAVE_BF16)
                tma_store_2d(&tma_aux, ring + (2 + grp) * GEMM_STAGING_BYTES, n_blk * BN + c * CHUNK, m_blk * GEMM_BM);
              tma_store_commit();
            }
          }
        }
        }
      } else {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        const int row = m_blk * GEMM_BM + r;
        const bool row_ok = row < p.M;
#pragma unroll 1
        for (int c = 0; c < BN / 16; ++c) {
          uint32_t rr[16];
          tmem_ld_32x32b_x16(t_row + c * 16, rr);
          tmem_ld_wait();
          const int col0 = n_blk * BN + c * 16;
          if (row_ok && col0 < p.N) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(rr[j]);
            const int ncols = min(16, p.N - col0);
            if (p.bias != nullptr) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < ncols) v[j] += __ldg(p.bias + col0 + j);
            }
            if constexpr (EPI == EPI_ACCUM_F32) {
              float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col0;
              if (ncols == 16) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * j), "f"(v[4 * j]),
                               "f"(v[4 * j + 1]), "f"(v[4 * j + 2]), "f"(v[4 * j + 3])
                               : "memory");
              } else {
                for (int j = 0; j < ncols; ++j) atomicAdd(o + j, v[j]);
              }
            } else if constexpr (EPI == EPI_POS_F32) {
              const float* a = p.aux + static_cast<size_t>(row % p.period) * p.N + col0;
              float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col0;
              if (ncols == 16) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float4 x = reinterpret_cast<const float4*>(a)[j];
                  x.x += v[4 * j + 0]; x.y += v[4 * j + 1]; x.z += v[4 * j + 2]; x.w += v[4 * j + 3];
                  reinterpret_cast<float4*>(o)[j] = x;
                }
              } else {
                for (int j = 0; j < ncols; ++j) o[j] = a[j] + v[j];
              }
            } else {   // EPI_NCHW_F32
              const int img = row / p.period;
              const int pix = row - img * p.period;
              float* o = reinterpret_cast<float*>(p.out) + (static_cast<size_t>(img) * p.N + col0) * p.period + pix;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < ncols) o[static_cast<size_t>(j) * p.period] = v[j];   // lanes = consecutive pixels
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_tempty(acc);
      }
      if constexpr (LN_SPLIT) {
        acc_phase ^= 1;                     // the same buffer again, two tiles later
      } else {
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
    if (STAGED && !gemm_epi_adds_tile(EPI) && etid == 0) tma_store_wait_all<0>();   // output committed before exit
    if (timing && etid == 0 && grp == 0) {
      p.dbg[8] = w_tfull; p.dbg[9] = w_rfull; p.dbg[10] = w_sib; p.dbg[11] = w_bar;
      p.dbg[12] = clock64() - t_role; p.dbg[13] = n_epi_tiles;
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CG == 2) cluster_sync_all();   // the peer may still be reading our smem / signalling our barriers
  if (warp == 1) {
    if constexpr (CG == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace vpb
