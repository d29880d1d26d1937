// Backward of the fused multi-head self-attention (Attention.forward, mmpose/models/backbones/vit.py:99-115) for the
// training-step configurations (T = 192 tokens; head_dim 32 / 64 / 80 = ViTPose-S / -B, -L / -H).
//   S = (Q K^T) * scale, P = softmax(S), O = P V            (forward, attention.cu)
//   dV = P^T dO,  dP = dO V^T,  dS = P o (dP - delta),  delta_i = sum_d dO_id O_id,
//   dQ = scale * dS K,  dK = scale * dS^T Q
// One CTA per (crop, head). Q, K, V, dO and O of the head are TMA-loaded once as 128B-swizzled [192 x 64] tiles (plus
// a 16-column SWIZZLE_32B tile each for head_dim 80; TMEM column numbers below are those of head_dim <= 64); O lands
// in the dS area and is only read for delta. The query rows are processed in two 128-row tiles (the second is half
// empty). Per tile t, in issue order:
//   tcgen05.mma S  = Q_t K^T       -> TMEM [0,192)   ; 256 threads ((query row, half of the keys) each) form
//                                                       P = exp2(S * scale * log2e - lse) with the forward pass's
//                                                       log-sum-exp in ONE pass and write it (bf16) to smem
//   tcgen05.mma dP = dO_t V^T      -> same TMEM columns; meanwhile (tile 0) the threads form delta from the dO / O tiles
//   tcgen05.mma dV += P^T dO_t     -> TMEM [384,512)  (needs only P: runs while the threads form dS)
//                                                       dS = scale * P o (dP - delta) (bf16, smem)
//   tcgen05.mma S of tile t + 1    (head_dim <= 64: its columns are free once dP has been read)
//   tcgen05.mma dQ_t = dS K        -> TMEM [192,256)  (K consumed as an MN-major operand), committed on its own:
//                                                       drained while dK runs (tile 0: behind the wait for dP of tile 1)
//   tcgen05.mma dK += dS^T Q_t     -> TMEM [256,384): the A operands of dK / dV are the dS / P tiles read MN-major
//                                                 (transposed) straight from where the threads wrote them
// The bias gradient of attn.qkv (column sums of dQ / dK / dV) is taken from the staged bf16 output blocks. Block b
// prefetches the tiles of block b + #SMs into L2. VPB_ATTBWD_DEBUG=<cta>: clock64 stamps of every phase of that CTA.
// Nothing of size T x T ever touches HBM. Every MMA is M = 128: rows past the sequence end compute garbage from
// whatever follows the tile in shared memory, land in TMEM lanes that are never read, and never enter a contraction.
#include <cstdio>
#include <type_traits>
#include <cstdlib>
#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

namespace vpb {

constexpr int AB_T = 192;
constexpr int AB_THREADS = 288;                     // warp 0: TMA + MMA issue; warps 1..8: (row, column half) per thread
constexpr int AB_TILE = AB_T * 128;                 // SWIZZLE_128B box: [192 rows][128 B] = 64 columns
constexpr int AB_BOXB = AB_T * 32;                  // SWIZZLE_32B box: [192 rows][32 B] = the 16-column remainder of head_dim 80
constexpr int AB_CHUNK = 128 * 128;                 // [128 rows][64 keys] bf16

// head_dim 32 / 64: one 64-column box per operand (head_dim 32 uses its first 32 columns; the rest belongs to the next
// head and never enters an MMA). head_dim 80 (ViTPose-H): 64 columns + a 16-column SWIZZLE_32B box, as in the forward
// kernel (attention.cu); every MMA over the head dimension is then one N = 64 / K-step-4 piece plus one N = 16 / K = 16
// piece. Shared memory: P | dS | Q K V dO (| their 16-column boxes) | slack. M = 128 reads of the second query / key tile
// run past their 192-row tile into whatever follows (P -> dS, dS -> Q, dO -> slack): garbage rows land in TMEM lanes
// that are never read.
template <int HD>
struct AttnBwdCfg {
  static constexpr bool WIDE = HD > 64;
  static constexpr int HDM = WIDE ? 64 : HD;                        // columns of the 128-byte box the MMAs use
  static constexpr int SLACK = WIDE ? 2048 : 8192;                  // rows 192..255 of the last operand tile
  static constexpr int SMEM = 6 * AB_CHUNK + 4 * AB_TILE + (WIDE ? 4 * AB_BOXB : 0) + SLACK + 1024;
  // TMEM columns: S / dP [0,192). head_dim <= 64: dQ [192,256), dK [256,384), dV [384,512). head_dim 80: dK 2 x 80 and
  // dV 2 x 80 fill the rest, so dQ takes the S columns once dP has been consumed and the next tile's S waits for it.
  static constexpr int COL_DQ = WIDE ? 0 : 192;
  static constexpr int COL_DK = WIDE ? 192 : 256;
  static constexpr int COL_DV = WIDE ? 192 + 2 * HD : 384;
  static constexpr int QC = HD / 2;                                 // dQ columns per thread
  static constexpr int Q_PIECES = QC * 2 / 16;                      // 16-byte pieces of a thread's dQ row
  static constexpr int Q_PITCH = (Q_PIECES | 1) * 16;               // odd number of 16-byte units: conflict-free
  static constexpr int K_PIECES = HD * 2 / 16;                      // 16-byte pieces of a dK / dV row
  static constexpr int K_PITCH = HD * 2 + 16;
  static_assert(HD == 32 || HD == 64 || HD == 80, "attention backward handles head_dim 32 / 64 / 80");
  static_assert(COL_DV + 2 * HD <= 512 && 8 * 32 * Q_PITCH <= 3 * AB_CHUNK && 8 * 32 * K_PITCH <= 4 * AB_TILE, "budget");
};

struct AttnBwdParams {
  int n, heads;
  float scale, scale_log2e;
  const __nv_bfloat16* dO;    // [n, T, heads*hd]
  const __nv_bfloat16* O;     // [n, T, heads*hd] forward output
  const float* lse;           // [n, heads, T] log2-sum-exp of the scaled scores, written by the forward kernel
  __nv_bfloat16* dqkv;        // [n, T, 3*heads*hd]
  float* dbias;               // optional [3*heads*hd]: += column sums of dqkv over all tokens (attn.qkv's bias gradient)
  long long* dbg;             // VPB_ATTBWD_DEBUG=<cta>: clock64 stamps of that CTA (thread 0: [0,16), thread 32: [16,32))
  int dbg_cta;
  int prefetch;               // > 0: block b prefetches the tiles of block b + prefetch into L2
};

// fp32 columns [0, N) of this thread's TMEM lane, N a multiple of 8
template <int N>
__device__ __forceinline__ void tmem_ld_row(uint32_t taddr, float (&out)[N]) {
  static_assert(N % 8 == 0, "pieces of 8 columns");
  uint32_t v[N];                        // every load in flight, one wait
#pragma unroll
  for (int c = 0; c + 16 <= N; c += 16) tmem_ld_32x32b_x16(taddr + c, *reinterpret_cast<uint32_t(*)[16]>(&v[c]));
  if constexpr (N % 16 != 0) tmem_ld_32x32b_x8(taddr + (N - 8), *reinterpret_cast<uint32_t(*)[8]>(&v[N - 8]));
  tmem_ld_wait();
#pragma unroll
  for (int j = 0; j < N; ++j) out[j] = __uint_as_float(v[j]);
}

// dst[c] += sum over the first `rows` rows of a staged bf16 block [32 rows x NCOLS] (row pitch PITCH bytes), c < NCOLS:
// lane j owns the 32-bit words j, j + 32, ... of every row (two columns each; consecutive lanes read consecutive
// words: conflict-free). The sums are those of the bf16 values the kernel stores, i.e. exactly the column sums of dqkv.
// (A shuffle tree over the fp32 registers cost ~2.5x the instructions: 9 of 104 us per launch.)
template <int NCOLS, int PITCH>
__device__ __forceinline__ void staged_colsum_atomic(const uint8_t* stage, int rows, int lane, float* dst) {
  constexpr int NW = NCOLS / 2;
#pragma unroll
  for (int w0 = 0; w0 < NW; w0 += 32) {
    const int w = w0 + lane;
    if (w < NW) {
      // rows is warp-uniform but not known at compile time: fixed trip count with predicated loads, so that the loads
      // of a batch are in flight together (a `rows`-bounded loop ran as 32 dependent load -> add steps)
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const uint32_t x = i < rows ? *reinterpret_cast<const uint32_t*>(stage + i * PITCH + w * 4) : 0u;
        const uint32_t y = i + 1 < rows ? *reinterpret_cast<const uint32_t*>(stage + (i + 1) * PITCH + w * 4) : 0u;
        s0 += __uint_as_float(x << 16);
        s1 += __uint_as_float(x & 0xffff0000u);
        s2 += __uint_as_float(y << 16);
        s3 += __uint_as_float(y & 0xffff0000u);
      }
      s0 += s2;
      s1 += s3;
      atomicAdd(dst + 2 * w, s0);
      atomicAdd(dst + 2 * w + 1, s1);
    }
  }
}

#define AB_STAMP(slot)                                                                  \
  do {                                                                                  \
    if (p.dbg != nullptr && blockIdx.x == p.dbg_cta && (threadIdx.x == 0 || threadIdx.x == 32)) \
      p.dbg[(threadIdx.x == 0 ? 0 : 16) + (slot)] = clock64();                          \
  } while (0)

// 9 warps: one sub-partition holds three of them, so 16384 / (3 * 32) = 170 registers per thread is the ceiling
template <int HD>
__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                     const __grid_constant__ CUtensorMap tm_o, const __grid_constant__ CUtensorMap tm_qkvb,
                     const __grid_constant__ CUtensorMap tm_dob, const __grid_constant__ CUtensorMap tm_ob,
                     const AttnBwdParams p) {
  using C = AttnBwdCfg<HD>;
  constexpr bool WIDE = C::WIDE;
  constexpr int HDM = C::HDM;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_p = smem;                              // 3 chunks of [128 q][64 keys]
  uint8_t* s_ds = s_p + 3 * AB_CHUNK;               // 3 chunks
  uint8_t* s_q = s_ds + 3 * AB_CHUNK;
  uint8_t* s_k = s_q + AB_TILE;
  uint8_t* s_v = s_k + AB_TILE;
  uint8_t* s_do = s_v + AB_TILE;
  uint8_t* s_qb = s_do + AB_TILE;                   // head_dim 80: columns [64, 80) of Q, K, V, dO
  uint8_t* s_kb = s_qb + AB_BOXB;
  uint8_t* s_vb = s_kb + AB_BOXB;
  uint8_t* s_dob = s_vb + AB_BOXB;
  __shared__ uint64_t bar_load, bar_s, bar_sdone, bar_dp, bar_ds, bar_mma, bar_dq, bar_dqr;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_delta[2][2][128];              // [tile][half][row]: partial <dO_row, O_row>

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x % p.heads;
  const int crop = blockIdx.x / p.heads;
  const int ld_o = p.heads * HD, ld_qkv = 3 * ld_o;
  AB_STAMP(0);

  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_sdone, 256);
    mbar_init(&bar_dp, 1);
    mbar_init(&bar_ds, 256);
    mbar_init(&bar_mma, 1);
    mbar_init(&bar_dq, 256);
    mbar_init(&bar_dqr, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_o);
    if constexpr (WIDE) {
      tma_prefetch_desc(&tm_qkvb);
      tma_prefetch_desc(&tm_dob);
      tma_prefetch_desc(&tm_ob);
    }
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_slot;                // S / dP
  const uint32_t tmem_dq = tmem_s + C::COL_DQ;
  const uint32_t tmem_dk = tmem_s + C::COL_DK;      // two key tiles x HD columns
  const uint32_t tmem_dv = tmem_s + C::COL_DV;

  if (warp == 0) {
    if (lane == 0) {
      // the forward output O goes where dS will be written: it is only needed for delta = <dO_row, O_row>, which the
      // threads form while they wait for the first dP
      mbar_arrive_expect_tx(&bar_load, 5 * AB_TILE + (WIDE ? 5 * AB_BOXB : 0));
      tma_load_3d(s_q, &tm_qkv, &bar_load, head * HD, 0, crop);
      tma_load_3d(s_k, &tm_qkv, &bar_load, ld_o + head * HD, 0, crop);
      tma_load_3d(s_v, &tm_qkv, &bar_load, 2 * ld_o + head * HD, 0, crop);
      tma_load_3d(s_do, &tm_do, &bar_load, head * HD, 0, crop);
      tma_load_3d(s_ds, &tm_o, &bar_load, head * HD, 0, crop);
      if constexpr (WIDE) {
        tma_load_3d(s_ds + AB_TILE, &tm_ob, &bar_load, head * HD + 64, 0, crop);
        tma_load_3d(s_qb, &tm_qkvb, &bar_load, head * HD + 64, 0, crop);
        tma_load_3d(s_kb, &tm_qkvb, &bar_load, ld_o + head * HD + 64, 0, crop);
        tma_load_3d(s_vb, &tm_qkvb, &bar_load, 2 * ld_o + head * HD + 64, 0, crop);
        tma_load_3d(s_dob, &tm_dob, &bar_load, head * HD + 64, 0, crop);
      }
      // the tiles of the CTA that will follow on some SM one wave later: L2 hits instead of HBM latency at its start
      if (const int nxt = blockIdx.x + p.prefetch; p.prefetch > 0 && nxt < p.n * p.heads) {
        const int h2 = nxt % p.heads, c2 = nxt / p.heads;
        tma_prefetch_l2_3d(&tm_qkv, h2 * HD, 0, c2);
        tma_prefetch_l2_3d(&tm_qkv, ld_o + h2 * HD, 0, c2);
        tma_prefetch_l2_3d(&tm_qkv, 2 * ld_o + h2 * HD, 0, c2);
        tma_prefetch_l2_3d(&tm_do, h2 * HD, 0, c2);
        tma_prefetch_l2_3d(&tm_o, h2 * HD, 0, c2);
        if constexpr (WIDE) {
          tma_prefetch_l2_3d(&tm_qkvb, h2 * HD + 64, 0, c2);
          tma_prefetch_l2_3d(&tm_qkvb, ld_o + h2 * HD + 64, 0, c2);
          tma_prefetch_l2_3d(&tm_qkvb, 2 * ld_o + h2 * HD + 64, 0, c2);
          tma_prefetch_l2_3d(&tm_dob, h2 * HD + 64, 0, c2);
          tma_prefetch_l2_3d(&tm_ob, h2 * HD + 64, 0, c2);
        }
      }
      mbar_wait(&bar_load, 0);
      tc_fence_after();
      AB_STAMP(1);
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, AB_T);            // K-major A and B
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, HDM, 0, 1);      // B (K) MN-major
      constexpr uint32_t idesc_dq16 = umma_idesc_bf16(128, 16, 0, 1);
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, HDM, 1, 1);       // A (P^T / dS^T) and B MN-major
      constexpr uint32_t idesc_t16 = umma_idesc_bf16(128, 16, 1, 1);
      auto issue_s = [&](uint32_t q_t, uint32_t qb_t) {        // S = Q_t K^T
#pragma unroll
        for (int ks = 0; ks < HDM / 16; ++ks)
          umma_bf16_ss(tmem_s, umma_desc_k_sw128(q_t + ks * 32), umma_desc_k_sw128(smem_u32(s_k) + ks * 32), idesc_s,
                       ks != 0);
        if constexpr (WIDE) umma_bf16_ss(tmem_s, umma_desc_k_sw32(qb_t), umma_desc_k_sw32(smem_u32(s_kb)), idesc_s, 1u);
        umma_commit(&bar_s);
      };
      auto issue_tile = [&](auto tc) {
        constexpr int t = decltype(tc)::value;
        const uint32_t q_t = smem_u32(s_q) + t * AB_CHUNK, do_t = smem_u32(s_do) + t * AB_CHUNK;
        const uint32_t qb_t = smem_u32(s_qb) + t * 128 * 32, dob_t = smem_u32(s_dob) + t * 128 * 32;
        if constexpr (WIDE && t == 1) {   // dQ of tile 0 sits in the S columns until every thread has loaded it
          mbar_wait(&bar_dq, 0);
          tc_fence_after();
        }
        if constexpr (WIDE || t == 0) issue_s(q_t, qb_t);
        // dP = dO_t V^T into the same columns once every thread has consumed S
        mbar_wait(&bar_sdone, t);
        tc_fence_after();
        AB_STAMP(2 + 4 * t);
#pragma unroll
        for (int ks = 0; ks < HDM / 16; ++ks)
          umma_bf16_ss(tmem_s, umma_desc_k_sw128(do_t + ks * 32), umma_desc_k_sw128(smem_u32(s_v) + ks * 32), idesc_s,
                       ks != 0);
        if constexpr (WIDE) umma_bf16_ss(tmem_s, umma_desc_k_sw32(dob_t), umma_desc_k_sw32(smem_u32(s_vb)), idesc_s, 1u);
        umma_commit(&bar_dp);
        // dV += P^T dO_t needs only P: it runs on the tensor pipe while the threads form dS
        constexpr int nks = t == 0 ? 8 : (AB_T - 128) / 16;     // the valid queries of this tile
#pragma unroll
        for (int m = 0; m < 2; ++m) {
#pragma unroll
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t acc = (t | ks) != 0 ? 1u : 0u;
            const uint64_t a_p = umma_desc_mn_sw128(smem_u32(s_p) + 2 * m * AB_CHUNK + ks * 2048, AB_CHUNK);
            umma_bf16_ss(tmem_dv + m * HD, a_p, umma_desc_mn_sw128(do_t + ks * 2048, AB_TILE), idesc_t, acc);
            if constexpr (WIDE)
              umma_bf16_ss(tmem_dv + m * HD + 64, a_p, umma_desc_mn_sw32(dob_t + ks * 512), idesc_t16, acc);
          }
        }
        mbar_wait(&bar_ds, t);
        tc_fence_after();
        AB_STAMP(3 + 4 * t);
        // head_dim <= 64: the next tile's S goes first — its TMEM columns are free (every thread has read dP) and the
        // threads start on its P while dQ / dK of this tile are still running; dQ of tile 0 is drained later, behind
        // the wait for dP of tile 1. (head_dim 80: dQ occupies the S columns, the order stays S after dQ.)
        if constexpr (!WIDE && t == 0) issue_s(smem_u32(s_q) + AB_CHUNK, 0u);
        // dQ_t = dS K (contraction over the 192 keys)
#pragma unroll
        for (int ks = 0; ks < AB_T / 16; ++ks) {
          const uint64_t a = umma_desc_k_sw128(smem_u32(s_ds) + (ks / 4) * AB_CHUNK + (ks % 4) * 32);
          umma_bf16_ss(tmem_dq, a, umma_desc_mn_sw128(smem_u32(s_k) + ks * 2048, AB_TILE), idesc_dq, ks != 0);
          if constexpr (WIDE)
            umma_bf16_ss(tmem_dq + 64, a, umma_desc_mn_sw32(smem_u32(s_kb) + ks * 512), idesc_dq16, ks != 0);
        }
        umma_commit(&bar_dqr);          // dQ (and every MMA before it: dV) complete -> the threads drain dQ ...
        // ... while dK += dS^T Q_t runs (contraction over the valid queries of this tile)
#pragma unroll
        for (int m = 0; m < 2; ++m) {
#pragma unroll
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t acc = (t | ks) != 0 ? 1u : 0u;
            const uint64_t a_ds = umma_desc_mn_sw128(smem_u32(s_ds) + 2 * m * AB_CHUNK + ks * 2048, AB_CHUNK);
            umma_bf16_ss(tmem_dk + m * HD, a_ds, umma_desc_mn_sw128(q_t + ks * 2048, AB_TILE), idesc_t, acc);
            if constexpr (WIDE)
              umma_bf16_ss(tmem_dk + m * HD + 64, a_ds, umma_desc_mn_sw32(qb_t + ks * 512), idesc_t16, acc);
          }
        }
        umma_commit(&bar_mma);
        AB_STAMP(4 + 4 * t);
        // The threads rewrite P once dV has consumed it (bar_s of the next tile / bar_dqr: both commits follow the dV
        // MMAs) and dS after bar_mma (dK has consumed it).
      };
      issue_tile(std::integral_constant<int, 0>{});
      issue_tile(std::integral_constant<int, 1>{});
    }
  } else {
    const int quad = warp & 3;
    const int half = (warp - 1) >> 2;               // which half of the key columns (and of the output columns)
    const int r = quad * 32 + lane;                 // row inside a tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    constexpr int KH = AB_T / 2;                    // 96 keys per thread
    // The forward log-sum-exp of this thread's row in BOTH query tiles, requested up front. delta = <dO_row, O_row> is
    // formed from the TMA-loaded tiles while the first dP is in flight (O sits in the dS area; both tiles carry the same
    // swizzle, so the 16-byte pieces pair up by position). (Loading the
    // rows with 16-byte global loads instead cost ~5000 cycles per CTA after the operands had already landed.)
    float delta_t[2] = {0.f, 0.f}, lse_t[2] = {0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 2; ++t)
      if (t * 128 + r < AB_T)
        lse_t[t] = __ldg(p.lse + (static_cast<size_t>(crop) * p.heads + head) * AB_T + t * 128 + r);
    auto form_delta = [&]() {
      mbar_wait(&bar_load, 0);         // completed long ago: orders this thread's reads behind the TMA writes
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int token = t * 128 + r;
        if (token < AB_T) {
          float d = 0.f;
          auto dot16 = [&](const uint8_t* a, const uint8_t* b) {
            const uint4 x = *reinterpret_cast<const uint4*>(a), y = *reinterpret_cast<const uint4*>(b);
            const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 fx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xw[j]));
              const float2 fy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yw[j]));
              d = fmaf(fx.x, fy.x, fmaf(fx.y, fy.y, d));
            }
          };
          // each half takes half of the row's 16-byte pieces; the partial sums meet in shared memory
#pragma unroll
          for (int u = half * (HDM / 16); u < (half + 1) * (HDM / 16); ++u) {
            const int off = token * 128 + ((u ^ (token & 7)) * 16);
            dot16(s_do + off, s_ds + off);
          }
          if constexpr (WIDE) dot16(s_dob + token * 32 + half * 16, s_ds + AB_TILE + token * 32 + half * 16);
          s_delta[t][half][r] = d;
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");   // every row of O has been read: the dS area may be written
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (t * 128 + r < AB_T) delta_t[t] = s_delta[t][0][r] + s_delta[t][1][r];
    };
    // dQ rows of a tile: half of the head's columns per thread. A row-per-thread global store would touch 32 lines with
    // 16 bytes each per instruction, so the warp stages its 32-row block in shared memory and writes whole row pieces
    // back. The block lives in the part of the P (or dS) tile only this warp writes (its 32 rows of key chunk 0 / chunk
    // 2 for the two halves), so no other warp's next P / dS can touch it; row pitch an odd number of 16-byte units:
    // conflict-free both ways.
    auto drain_dq = [&](int t, uint8_t* area) {
      float v[C::QC];
      tmem_ld_row<C::QC>(tmem_dq + lane_off + half * C::QC, v);
      if constexpr (WIDE) {             // the S columns are free for the next tile's scores
        tc_fence_before();
        mbar_arrive(&bar_dq);
      }
      uint8_t* stage = area + 2 * half * AB_CHUNK + quad * 32 * 128;
      static_assert(32 * C::Q_PITCH <= 32 * 128, "the staging block fits the warp's own rows of one key chunk");
#pragma unroll
      for (int u = 0; u < C::Q_PIECES; ++u)
        *reinterpret_cast<uint4*>(stage + lane * C::Q_PITCH + u * 16) =
            make_uint4(pack_bf16x2(v[8 * u], v[8 * u + 1]), pack_bf16x2(v[8 * u + 2], v[8 * u + 3]),
                       pack_bf16x2(v[8 * u + 4], v[8 * u + 5]), pack_bf16x2(v[8 * u + 6], v[8 * u + 7]));
      __syncwarp();
      const int live_rows = min(32, max(0, AB_T - (t * 128 + quad * 32)));
      if (p.dbias != nullptr)           // bias gradient of attn.qkv: column sums over the live rows of this warp
        staged_colsum_atomic<C::QC, C::Q_PITCH>(stage, live_rows, lane, p.dbias + head * HD + half * C::QC);
      __nv_bfloat16* obase = p.dqkv + (static_cast<size_t>(crop) * AB_T + t * 128 + quad * 32) * ld_qkv +
                             head * HD + half * C::QC;
#pragma unroll
      for (int j = 0; j < C::Q_PIECES; ++j) {
        const int idx = j * 32 + lane, row = idx / C::Q_PIECES, ch = idx % C::Q_PIECES;
        if (t * 128 + quad * 32 + row < AB_T)
          *reinterpret_cast<uint4*>(obase + static_cast<size_t>(row) * ld_qkv + ch * 8) =
              *reinterpret_cast<const uint4*>(stage + row * C::Q_PITCH + ch * 16);
      }
      __syncwarp();
    };
    for (int t = 0; t < 2; ++t) {
      const float lse = t == 0 ? lse_t[0] : lse_t[1];
      AB_STAMP(1 + 6 * t);
      mbar_wait(&bar_s, t);
      tc_fence_after();
      AB_STAMP(2 + 6 * t);
      // Warps whose 32 rows all lie past the sequence end (second tile, rows 64..127) skip the arithmetic: P / dS rows of
      // dead queries never enter a contraction (dK / dV contract over the live queries only, dQ rows are per query).
      const bool warp_live = t * 128 + quad * 32 < AB_T;
      // P = exp2(s * scale * log2e - lse), one pass over this thread's 96 keys
      for (int c = half * KH; warp_live && c < (half + 1) * KH; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_s + lane_off + c, v);
        tmem_ld_wait();
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float e0 = fast_ex2(fmaf(__uint_as_float(v[j]), p.scale_log2e, -lse));
          const float e1 = fast_ex2(fmaf(__uint_as_float(v[j + 1]), p.scale_log2e, -lse));
          packed[j / 2] = pack_bf16x2(e0, e1);
        }
        const uint32_t row = smem_u32(s_p) + (c / 64) * AB_CHUNK + r * 128;
        const int u0 = (c % 64) / 8;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          sts_u4(row + (((u0 + u) ^ (r & 7)) * 16), packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
      }
      tc_fence_before();               // all tcgen05.ld of S done before dP overwrites the columns
      fence_proxy_async_smem();        // P (generic-proxy writes) visible to the tensor core: dV += P^T dO starts now
      mbar_arrive(&bar_sdone);
      AB_STAMP(3 + 6 * t);
      if (t == 0) {
        form_delta();
      } else {
        mbar_wait(&bar_mma, 0);        // dK of tile 0 has consumed the dS tile (and dQ of tile 0 is complete)
        if constexpr (!WIDE) {         // head_dim <= 64: dQ of tile 0 is drained here, behind the wait for dP of tile 1
          tc_fence_after();
          drain_dq(0, s_ds);
          tc_fence_before();
        }
      }
      const float delta = t == 0 ? delta_t[0] : delta_t[1];
      mbar_wait(&bar_dp, t);
      tc_fence_after();
      AB_STAMP(4 + 6 * t);
      // dS = scale * P o (dP - delta)
      for (int c = half * KH; warp_live && c < (half + 1) * KH; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_s + lane_off + c, v);
        tmem_ld_wait();
        const uint32_t prow = smem_u32(s_p) + (c / 64) * AB_CHUNK + r * 128;
        const uint32_t drow = smem_u32(s_ds) + (c / 64) * AB_CHUNK + r * 128;
        const int u0 = (c % 64) / 8;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t so = ((u0 + u) ^ (r & 7)) * 16;
          uint32_t pw[4];
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(pw[0]), "=r"(pw[1]), "=r"(pw[2]), "=r"(pw[3])
                       : "r"(prow + so));
          uint32_t o[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 pp = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pw[j]));
            const float d0 = pp.x * (__uint_as_float(v[8 * u + 2 * j]) - delta) * p.scale;
            const float d1 = pp.y * (__uint_as_float(v[8 * u + 2 * j + 1]) - delta) * p.scale;
            o[j] = pack_bf16x2(d0, d1);
          }
          sts_u4(drow + so, o[0], o[1], o[2], o[3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();         // dS (generic-proxy writes) visible to the tensor core
      mbar_arrive(&bar_ds);
      AB_STAMP(5 + 6 * t);
      if (WIDE || t == 1) {
        mbar_wait(&bar_dqr, t);
        tc_fence_after();
        AB_STAMP(6 + 6 * t);
        drain_dq(t, s_p);               // dV of this tile has consumed P (bar_dqr)
      }
      tc_fence_before();
    }
    mbar_wait(&bar_mma, 1);
    tc_fence_after();
    AB_STAMP(13);
    // dK / dV rows (keys): key tile m, lane r <-> key m*128 + r; half 0 stores dK, half 1 stores dV. Every operand tile
    // is dead by now (all MMAs have completed): the warp stages its 32-row block in the Q/K/V/dO area and stores whole
    // rows, HD / 8 lanes per row.
    for (int m = 0; m < 2; ++m) {
      const uint32_t base = (half == 0 ? tmem_dk : tmem_dv) + m * HD + lane_off;
      uint8_t* stage = s_q + (warp - 1) * (32 * C::K_PITCH);
      auto piece = [&](auto nc, int c) {
        constexpr int NC = decltype(nc)::value;
        float v[NC];
        tmem_ld_row<NC>(base + c, v);
#pragma unroll
        for (int u = 0; u < NC / 8; ++u)
          *reinterpret_cast<uint4*>(stage + lane * C::K_PITCH + c * 2 + u * 16) =
              make_uint4(pack_bf16x2(v[8 * u], v[8 * u + 1]), pack_bf16x2(v[8 * u + 2], v[8 * u + 3]),
                         pack_bf16x2(v[8 * u + 4], v[8 * u + 5]), pack_bf16x2(v[8 * u + 6], v[8 * u + 7]));
      };
#pragma unroll
      for (int c = 0; c + 32 <= HD; c += 32) piece(std::integral_constant<int, 32>{}, c);
      if constexpr (HD % 32 != 0) piece(std::integral_constant<int, 16>{}, HD - 16);
      __syncwarp();
      if (p.dbias != nullptr)
        staged_colsum_atomic<HD, C::K_PITCH>(stage, min(32, max(0, AB_T - (m * 128 + quad * 32))), lane,
                                             p.dbias + (1 + half) * ld_o + head * HD);
      __nv_bfloat16* obase = p.dqkv + (static_cast<size_t>(crop) * AB_T + m * 128 + quad * 32) * ld_qkv +
                             (1 + half) * ld_o + head * HD;
#pragma unroll
      for (int j = 0; j < C::K_PIECES; ++j) {
        const int idx = j * 32 + lane, row = idx / C::K_PIECES, ch = idx % C::K_PIECES;
        if (m * 128 + quad * 32 + row < AB_T)
          *reinterpret_cast<uint4*>(obase + static_cast<size_t>(row) * ld_qkv + ch * 8) =
              *reinterpret_cast<const uint4*>(stage + row * C::K_PITCH + ch * 16);
      }
      __syncwarp();
    }
  }
  AB_STAMP(14);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_s, 512);
  AB_STAMP(15);
}

template <int HD>
static int launch_attention_bwd(const CUtensorMap& tq, const CUtensorMap& tdo, const CUtensorMap& to,
                                const CUtensorMap& tqb, const CUtensorMap& tdob, const CUtensorMap& tob,
                                const AttnBwdParams& p, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    VPB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        AttnBwdCfg<HD>::SMEM));
    configured = true;
  }
  attention_bwd_kernel<HD><<<p.n * p.heads, AB_THREADS, AttnBwdCfg<HD>::SMEM, stream>>>(tq, tdo, to, tqb, tdob, tob, p);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int attention_bwd(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, int n, int T,
                  int heads, int hd, float scale, cudaStream_t stream, float* dbias) {
  VPB_REQUIRE(n > 0 && heads > 0 && lse != nullptr, "attention_bwd: empty problem / missing log-sum-exp");
  VPB_REQUIRE(T == AB_T && (hd == 32 || hd == 64 || hd == 80),
              "attention_bwd: built for T=%d, head_dim 32 / 64 / 80 (got T=%d, head_dim=%d)", AB_T, T, hd);
  const int ld_o = heads * hd, ld = 3 * ld_o;
  CUtensorMap tq, tdo, to, tqb, tdob, tob;
  uint64_t dims[3] = {(uint64_t)ld, (uint64_t)T, (uint64_t)n};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)T * ld * 2};
  uint32_t box[3] = {64, (uint32_t)T, 1};
  if (make_tma_desc(&tq, TMA_BF16, qkv, 3, dims, strides, box, TMA_SWIZZLE_128B)) return -1;
  uint64_t dims_o[3] = {(uint64_t)ld_o, (uint64_t)T, (uint64_t)n};
  uint64_t strides_o[2] = {(uint64_t)ld_o * 2, (uint64_t)T * ld_o * 2};
  if (make_tma_desc(&tdo, TMA_BF16, dout, 3, dims_o, strides_o, box, TMA_SWIZZLE_128B)) return -1;
  if (make_tma_desc(&to, TMA_BF16, out, 3, dims_o, strides_o, box, TMA_SWIZZLE_128B)) return -1;
  tqb = tq;
  tdob = tdo;
  tob = to;
  if (hd > 64) {      // the 16-column remainder of each operand
    uint32_t box_b[3] = {16, (uint32_t)T, 1};
    if (make_tma_desc(&tqb, TMA_BF16, qkv, 3, dims, strides, box_b, TMA_SWIZZLE_32B)) return -1;
    if (make_tma_desc(&tdob, TMA_BF16, dout, 3, dims_o, strides_o, box_b, TMA_SWIZZLE_32B)) return -1;
    if (make_tma_desc(&tob, TMA_BF16, out, 3, dims_o, strides_o, box_b, TMA_SWIZZLE_32B)) return -1;
  }
  AttnBwdParams p;
  p.n = n; p.heads = heads; p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.dO = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.O = reinterpret_cast<const __nv_bfloat16*>(out);
  p.lse = lse;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.dbias = dbias;
  p.dbg = nullptr;
  p.dbg_cta = 0;
  {
    static const int pf = getenv("VPB_ATTBWD_PREFETCH") ? atoi(getenv("VPB_ATTBWD_PREFETCH")) : sm_count();
    p.prefetch = pf;
  }
  {
    static const bool debug = getenv("VPB_ATTBWD_DEBUG") != nullptr;
    if (debug) p.dbg_cta = atoi(getenv("VPB_ATTBWD_DEBUG"));
    static long long* buf = nullptr;
    if (debug) {
      if (buf == nullptr) VPB_CHECK_CUDA(cudaMallocManaged(&buf, 32 * sizeof(long long)));
      else {      // stamps of the previous launch
        cudaStreamSynchronize(stream);
        const long long t0 = buf[0];
        fprintf(stderr, "attention_bwd stamped CTA cycles | issuer: loaded %lld, t0: S consumed %lld dS ready %lld MMAs issued %lld, "
                        "t1: %lld %lld %lld, end %lld | thread 32: t0 wait-S %lld..%lld P done %lld dP ready %lld dS done %lld "
                        "MMAs done %lld, t1 %lld..%lld %lld %lld %lld %lld, dQ drained %lld, dK/dV stored %lld, exit %lld\n",
                buf[1] - t0, buf[2] - t0, buf[3] - t0, buf[4] - t0, buf[6] - t0, buf[7] - t0, buf[8] - t0, buf[15] - t0,
                buf[17] - t0, buf[18] - t0, buf[19] - t0, buf[20] - t0, buf[21] - t0, buf[22] - t0, buf[23] - t0,
                buf[24] - t0, buf[25] - t0, buf[26] - t0, buf[27] - t0, buf[28] - t0, buf[29] - t0, buf[30] - t0,
                buf[31] - t0);
      }
      p.dbg = buf;
    }
  }
  if (hd == 32) return launch_attention_bwd<32>(tq, tdo, to, tqb, tdob, tob, p, stream);
  if (hd == 64) return launch_attention_bwd<64>(tq, tdo, to, tqb, tdob, tob, p, stream);
  return launch_attention_bwd<80>(tq, tdo, to, tqb, tdob, tob, p, stream);
}

}  // namespace vpb
