// Fused multi-head self-attention for the fixed short ViTPose sequence (T = 16 x 12 = 192 tokens).
// Reference: Attention.forward, mmpose/models/backbones/vit.py:99-115 —
//   q, k, v = split(qkv); attn = softmax((q * scale) @ k^T); out = (attn @ v) re-interleaved per head.
//
// One CTA per (crop, head, 128-row query tile). The whole key/value sequence of one head fits on chip, so
// there is no online-softmax rescaling:
//   TMA (3-D map over [crop, token, column], token dim zero-filled past T) -> Q, K, V tiles in 128B-swizzled smem
//   tcgen05.mma  S[128 x T]  = Q . K^T            (fp32 in TMEM, K-major operands)
//   4 warps, one query row per thread: tcgen05.ld S, row max, exp2, row sum, P (bf16) written to smem in the
//            canonical K-major swizzled layout (overlays the dead Q/K tiles)
//   tcgen05.mma  O[128 x hd] = P . V              (V consumed as an MN-major operand straight from its TMA tile;
//                                                  O overlays S in TMEM)
//   tcgen05.ld O, scale by 1/rowsum, bf16 store to out[crop, token, head*hd + :].
// The scores never touch HBM (the eager reference materialises [N, h, 192, 192] fp32 per block).
// Head dims that are not a multiple of 64 (32 for ViT-S, 80 for ViT-H) are loaded as 64-column boxes; the MMA
// reads only the first hd columns / K-steps of them.
#include <stdlib.h>

#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

namespace vpb {

constexpr int ATT_THREADS = 160;   // warp 0: TMA + MMA issue; warps 1..4: softmax + epilogue
constexpr int ATT_BM = 128;

struct AttnParams {
  int n, T, heads, hd;
  int ldo;              // heads * hd
  float scale_log2e;    // scale * log2(e)
  __nv_bfloat16* out;
  float* lse;           // optional [n, heads, T]: log2(sum_j exp2(s_ij * scale * log2e)) per query row (for the backward pass)
  long long* dbg_buf;   // optional [16] cycle counters of CTA 0 (VPB_ATT_DEBUG & 32), see tools/att_debug.py
  int dbg;              // profiling aid (VPB_ATT_DEBUG): 1 skip softmax math, 2 skip P.V MMAs, 4 skip S MMAs, 8 skip K/V loads
};

// Tn = T rounded up to 64 (P chunks); all sizes in bytes
__host__ __device__ constexpr int att_boxes(int hd) { return (hd + 63) / 64; }

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const __grid_constant__ CUtensorMap tm_q,
                                                                const __grid_constant__ CUtensorMap tm_kv,
                                                                const AttnParams p) {
  constexpr int NB = att_boxes(HD);               // 64-column boxes per operand
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_load, bar_s, bar_p, bar_o;
  __shared__ uint32_t tmem_slot;

  const int T = p.T;
  const int kv_box_bytes = T * 128;               // [T rows][128 B]
  const int q_box_bytes = ATT_BM * 128;
  // layout: [Q boxes][K boxes] (later overlaid by P) | [V boxes]
  const int p_chunks = (T + 63) / 64;
  const int qk_bytes = NB * (q_box_bytes + kv_box_bytes);
  const int p_bytes = p_chunks * ATT_BM * 128;
  const int region0 = qk_bytes > p_bytes ? qk_bytes : p_bytes;
  uint8_t* s_q = smem;
  uint8_t* s_k = smem + NB * q_box_bytes;
  uint8_t* s_p = smem;
  uint8_t* s_v = smem + ((region0 + 1023) & ~1023);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x;
  const int qt = unit & 1;
  const int head = (unit >> 1) % p.heads;
  const int crop = (unit >> 1) / p.heads;
  if (qt * ATT_BM >= T) return;                    // (only when T <= 128)

  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_p, 128);
    mbar_init(&bar_o, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_slot;              // S: columns [0, T); O overlays columns [0, HD)

  if (warp == 0) {
    if (lane == 0) {
      const int col_q = head * HD, col_k = p.heads * HD + head * HD, col_v = 2 * p.heads * HD + head * HD;
      mbar_arrive_expect_tx(&bar_load, NB * (q_box_bytes + 2 * kv_box_bytes));
      for (int b = 0; b < NB; ++b) {
        tma_load_3d(s_q + b * q_box_bytes, &tm_q, &bar_load, col_q + b * 64, qt * ATT_BM, crop);
        tma_load_3d(s_k + b * kv_box_bytes, &tm_kv, &bar_load, col_k + b * 64, 0, crop);
        tma_load_3d(s_v + b * kv_box_bytes, &tm_kv, &bar_load, col_v + b * 64, 0, crop);
      }
      mbar_wait(&bar_load, 0);
      tc_fence_after();
      // S = Q . K^T : M=128, N=T, K=HD
      const uint32_t idesc_s = umma_idesc_bf16(ATT_BM, T);
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        const uint32_t a = smem_u32(s_q + (ks / 4) * q_box_bytes) + (ks % 4) * 32;
        const uint32_t b = smem_u32(s_k + (ks / 4) * kv_box_bytes) + (ks % 4) * 32;
        umma_bf16_ss(tmem_s, umma_desc_k_sw128(a), umma_desc_k_sw128(b), idesc_s, ks != 0);
      }
      umma_commit(&bar_s);
      // wait for P (bf16, smem) from the softmax warps, then O = P . V : M=128, N=HD, K=T
      mbar_wait(&bar_p, 0);
      tc_fence_after();
      const uint32_t idesc_o = umma_idesc_bf16(ATT_BM, HD, 0, 1);
      const int ksteps = T / 16;
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t a = smem_u32(s_p + (ks / 4) * (ATT_BM * 128)) + (ks % 4) * 32;
        const uint32_t b = smem_u32(s_v) + ks * 2048;             // 16 tokens x 128 B per K step
        umma_bf16_ss(tmem_s, umma_desc_k_sw128(a), umma_desc_mn_sw128(b, kv_box_bytes), idesc_o, ks != 0);
      }
      umma_commit(&bar_o);
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;               // query row inside the tile == TMEM lane
    const uint32_t t_row = tmem_s + (static_cast<uint32_t>(quad * 32) << 16);
    mbar_wait(&bar_s, 0);
    tc_fence_after();
    // pass 1: row max of the raw scores
    float mx = -INFINITY;
    for (int c = 0; c < T; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    const float mxs = mx * p.scale_log2e;
    // pass 2: p = exp2(s * scale*log2e - max'), row sum, bf16 P into swizzled smem
    float sum = 0.f;
    for (int c = 0; c < T; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + c, v);
      tmem_ld_wait();
      uint32_t packed[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float e0 = exp2f(fmaf(__uint_as_float(v[j]), p.scale_log2e, -mxs));
        const float e1 = exp2f(fmaf(__uint_as_float(v[j + 1]), p.scale_log2e, -mxs));
        // the row sum must match what the tensor core will see: sum the bf16-rounded values
        const __nv_bfloat162 b2 = __floats2bfloat162_rn(e0, e1);
        sum += __low2float(b2) + __high2float(b2);
        packed[j / 2] = *reinterpret_cast<const uint32_t*>(&b2);
      }
      // columns [c, c+32) = half of 64-column chunk c/64: 16-byte units u0..u0+3, XOR-swizzled by (row & 7)
      uint8_t* chunk = s_p + (c / 64) * (ATT_BM * 128) + r * 128;
      const int u0 = (c % 64) / 8;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        *reinterpret_cast<uint4*>(chunk + (((u0 + u) ^ (r & 7)) * 16)) =
            make_uint4(packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
      }
    }
    tc_fence_before();          // our tcgen05.ld of S complete before the MMA warp overwrites S with O
    fence_proxy_async_smem();   // generic-proxy writes of P visible to the tensor core (async proxy)
    mbar_arrive(&bar_p);
    // epilogue: O row * 1/sum -> bf16
    mbar_wait(&bar_o, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    const int token = qt * ATT_BM + r;
    __nv_bfloat16* orow = p.out + (static_cast<size_t>(crop) * T + token) * p.ldo + head * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(t_row + c, v);
      tmem_ld_wait();
      if (token < T) {
        uint4 w0 = make_uint4(pack_bf16x2(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv),
                              pack_bf16x2(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv),
                              pack_bf16x2(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv),
                              pack_bf16x2(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv));
        uint4 w1 = make_uint4(pack_bf16x2(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv),
                              pack_bf16x2(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv),
                              pack_bf16x2(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv),
                              pack_bf16x2(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv));
        reinterpret_cast<uint4*>(orow + c)[0] = w0;
        reinterpret_cast<uint4*>(orow + c)[1] = w1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_s, 256);
}

// -------------------------------------------------------------------------------------------------------------
// Persistent, pipelined variant for head_dim 32 / 64 and T = 192 (ViT-S/B/L): one CTA per SM walks the
// (crop, head, q-tile) units.
//   warp 0     : TMA producer with two independent rings — {Q,K} tiles (dead as soon as S = Q.K^T has been issued and
//                completed) and V tiles (needed until P.V) — so loads run several units ahead of their use
//   warp 1     : MMA issuer: S(i+2) is issued right after softmax(i) has read its scores, P.V(i) when P(i) is ready
//   warps 2..9 : 256 softmax threads, thread (row, half) owns half of the key columns of one query row: a score row is
//                read from TMEM exactly once and stays in registers; halves exchange row max / sum through smem.
//                softmax(i+1) runs before the epilogue of unit i, overlapping the exp work with P.V(i).
// TMEM: S0 | S1 | O0 | O1 (2T + 2HD <= 512 columns). P (bf16) has its own smem tile, reused every unit once P.V of
// the previous unit has completed. No per-unit TMEM allocation, barrier init or CTA launch; exp2 runs only for query
// rows that exist (q-tile 1 of a 192-token sequence is half empty).
// -------------------------------------------------------------------------------------------------------------
// softmax threads = 128 * NSPLIT: thread (row, part) owns T/NSPLIT key columns of one query row
__host__ __device__ constexpr int att2_threads(int nsplit) { return 64 + 128 * nsplit; }
constexpr int ATT2_QK_DEPTH = 2;
// head_dim 80 (ViT-H) = a 64-column SWIZZLE_128B box + a 16-column SWIZZLE_32B box per operand; its larger tiles
// leave room for two V stages and one O accumulator
__host__ __device__ constexpr int att2_v_depth(int hd) { return hd > 64 ? 2 : 3; }
__host__ __device__ constexpr int att2_row_bytes(int hd) { return hd > 64 ? 160 : 128; }

template <int HD, int T_, int NSPLIT>
__global__ void __launch_bounds__(att2_threads(NSPLIT), 1)
attention_persistent_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                            const __grid_constant__ CUtensorMap tm_qb, const __grid_constant__ CUtensorMap tm_kvb,
                            const AttnParams p, const int num_units) {
  static_assert(HD == 32 || HD == 64 || HD == 80, "persistent attention handles head_dim 32 / 64 / 80");
  constexpr bool WIDE = HD > 64;                    // second, 16-column box per operand (tm_qb / tm_kvb)
  constexpr int OB = (2 * T_ + 2 * HD <= 512) ? 2 : 1;   // O accumulators that fit TMEM next to S0 | S1
  static_assert(T_ % 64 == 0 && 2 * T_ + OB * HD <= 512, "each softmax thread owns T/2 keys; S0,S1,O must fit TMEM");
  constexpr int T = T_;
  constexpr int ATT2_V_DEPTH = att2_v_depth(HD);
  constexpr int QA_BYTES = ATT_BM * 128, KVA_BYTES = T * 128;          // 64-column boxes
  constexpr int Q_BYTES = ATT_BM * att2_row_bytes(HD), KV_BYTES = T * att2_row_bytes(HD);
  constexpr int P_BYTES = (T / 64) * ATT_BM * 128;
  constexpr int QK_BYTES = Q_BYTES + KV_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_p = smem;
  uint8_t* s_qk = smem + P_BYTES;
  uint8_t* s_v = s_qk + ATT2_QK_DEPTH * QK_BYTES;
  __shared__ uint64_t qk_full[ATT2_QK_DEPTH], qk_free[ATT2_QK_DEPTH], v_full[ATT2_V_DEPTH], v_free[ATT2_V_DEPTH];
  __shared__ uint64_t s_full[2], o_full[2], o_free[2], p_full;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_max[NSPLIT][ATT_BM], s_sum[NSPLIT][ATT_BM];   // [column part][row], exchanged between the parts
  static_assert(T_ % (8 * NSPLIT) == 0 && (T_ / NSPLIT) % 16 == 0 && (WIDE ? NSPLIT == 2 : (HD / NSPLIT) % 16 == 0),
                "unsupported column split");

  const int q_tiles = (T + ATT_BM - 1) / ATT_BM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_local = num_units > static_cast<int>(blockIdx.x)
                          ? (num_units - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                static_cast<int>(gridDim.x)
                          : 0;

  // Query tile of the i-th unit of this CTA. With an even grid, unit % q_tiles would be the same for every i, i.e. a
  // CTA would get only full (128-row) or only partial (64-row, half the softmax warps idle) tiles; rotating the tile
  // index by i hands every CTA the same mix. The q_tiles units of one (crop, head) share i whenever the grid is a
  // multiple of q_tiles, so the rotation is a bijection.
  const bool rotate_qt = (gridDim.x % q_tiles) == 0;
  auto unit_qt = [&](int unit, int i) { return rotate_qt ? (unit + i) % q_tiles : unit % q_tiles; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < ATT2_QK_DEPTH; ++s) { mbar_init(&qk_full[s], 1); mbar_init(&qk_free[s], 1); }
    for (int s = 0; s < ATT2_V_DEPTH; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_free[s], 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&s_full[b], 1);
      mbar_init(&o_full[b], 1);
      mbar_init(&o_free[b], 128 * NSPLIT);
    }
    mbar_init(&p_full, 128 * NSPLIT);
    fence_mbar_init();
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    if (WIDE) {
      tma_prefetch_desc(&tm_qb);
      tma_prefetch_desc(&tm_kvb);
    }
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // cycle accounting of the waits (CTA 0 only, one thread per role), enabled with VPB_ATT_DEBUG & 32
  const bool timing = p.dbg_buf != nullptr && blockIdx.x == 0;
  long long t_wait[6] = {0, 0, 0, 0, 0, 0};
  auto timed_wait = [&](uint64_t* bar, uint32_t parity, int slot) {
    if (timing) {
      const long long t0 = clock64();
      mbar_wait(bar, parity);
      t_wait[slot] += clock64() - t0;
    } else {
      mbar_wait(bar, parity);
    }
  };
  const long long t_start = clock64();

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < n_local; ++i) {
        const int unit = blockIdx.x + i * gridDim.x;
        const int qt = unit_qt(unit, i);
        const int head = (unit / q_tiles) % p.heads;
        const int crop = (unit / q_tiles) / p.heads;
        const int sq = i % ATT2_QK_DEPTH, sv = i % ATT2_V_DEPTH;
        timed_wait(&qk_free[sq], ((i / ATT2_QK_DEPTH) & 1) ^ 1, 0);
        mbar_arrive_expect_tx(&qk_full[sq], QK_BYTES);
        tma_load_3d(s_qk + sq * QK_BYTES, &tm_q, &qk_full[sq], head * HD, qt * ATT_BM, crop);
        tma_load_3d(s_qk + sq * QK_BYTES + Q_BYTES, &tm_kv, &qk_full[sq], p.heads * HD + head * HD, 0, crop);
        if constexpr (WIDE) {     // columns [64, 80) of Q and K
          tma_load_3d(s_qk + sq * QK_BYTES + QA_BYTES, &tm_qb, &qk_full[sq], head * HD + 64, qt * ATT_BM, crop);
          tma_load_3d(s_qk + sq * QK_BYTES + Q_BYTES + KVA_BYTES, &tm_kvb, &qk_full[sq], p.heads * HD + head * HD + 64,
                      0, crop);
        }
        timed_wait(&v_free[sv], ((i / ATT2_V_DEPTH) & 1) ^ 1, 1);
        mbar_arrive_expect_tx(&v_full[sv], KV_BYTES);
        tma_load_3d(s_v + sv * KV_BYTES, &tm_kv, &v_full[sv], 2 * p.heads * HD + head * HD, 0, crop);
        if constexpr (WIDE)
          tma_load_3d(s_v + sv * KV_BYTES + KVA_BYTES, &tm_kvb, &v_full[sv], 2 * p.heads * HD + head * HD + 64, 0, crop);
      }
      if (timing) { p.dbg_buf[0] = t_wait[0]; p.dbg_buf[1] = t_wait[1]; }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BM, T);
      constexpr uint32_t idesc_o = umma_idesc_bf16(ATT_BM, WIDE ? 64 : HD, 0, 1);
      constexpr uint32_t idesc_o16 = umma_idesc_bf16(ATT_BM, 16, 0, 1);      // WIDE: output columns [64, 80)
      auto issue_s = [&](int i) {
        const int sq = i % ATT2_QK_DEPTH, b = i & 1;
        timed_wait(&qk_full[sq], (i / ATT2_QK_DEPTH) & 1, 0);
        tc_fence_after();
        const uint32_t qa = smem_u32(s_qk + sq * QK_BYTES), ka = qa + Q_BYTES;
        if (!(p.dbg & 4)) {
#pragma unroll
          for (int ks = 0; ks < (WIDE ? 4 : HD / 16); ++ks)
            umma_bf16_ss(tmem_base + b * T, umma_desc_k_sw128(qa + ks * 32), umma_desc_k_sw128(ka + ks * 32), idesc_s,
                         ks != 0);
          if constexpr (WIDE)       // fifth K step: the 16-column SWIZZLE_32B boxes
            umma_bf16_ss(tmem_base + b * T, umma_desc_k_sw32(qa + QA_BYTES), umma_desc_k_sw32(ka + KVA_BYTES), idesc_s, 1u);
        }
        umma_commit(&s_full[b]);
        umma_commit(&qk_free[sq]);      // Q/K tiles are dead once S has been computed
      };
      // S_b may be overwritten once softmax(i-2) has read it, which p_full(i-2) implies
      if (n_local > 0) issue_s(0);
      if (n_local > 1) issue_s(1);
      for (int i = 0; i < n_local; ++i) {
        const int sv = i % ATT2_V_DEPTH, b = i % OB;
        timed_wait(&p_full, i & 1, 1);                     // P(i) in smem, S read
        timed_wait(&o_free[b], ((i / OB) & 1) ^ 1, 2);     // O_b drained by the epilogue of unit i - OB
        timed_wait(&v_full[sv], (i / ATT2_V_DEPTH) & 1, 3);
        tc_fence_after();
        const uint32_t d = tmem_base + 2 * T + b * HD;
        const uint32_t pa = smem_u32(s_p), va = smem_u32(s_v + sv * KV_BYTES);
        for (int ks = 0; ks < ((p.dbg & 2) ? 0 : T / 16); ++ks) {
          const uint64_t pd = umma_desc_k_sw128(pa + (ks / 4) * (ATT_BM * 128) + (ks % 4) * 32);
          umma_bf16_ss(d, pd, umma_desc_mn_sw128(va + ks * 2048, KVA_BYTES), idesc_o, ks != 0);
          if constexpr (WIDE)       // 16 keys x 32 B per K step in the SWIZZLE_32B box
            umma_bf16_ss(d + 64, pd, umma_desc_mn_sw32(va + KVA_BYTES + ks * 512), idesc_o16, ks != 0);
        }
        umma_commit(&o_full[b]);
        umma_commit(&v_free[sv]);
        if (i + 2 < n_local) issue_s(i + 2);
      }
      if (timing) { for (int k = 0; k < 4; ++k) p.dbg_buf[2 + k] = t_wait[k]; }
    }
  } else {
    constexpr int KH = T_ / NSPLIT;                 // keys per thread
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;               // column part of this thread (0 .. NSPLIT-1)
    const int r = quad * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);

    // softmax in two parts: (1) TMEM load, row max, exchange between the two column halves; (2) exp2, bf16 P, row sum.
    // (Measured: running the previous unit's epilogue between the parts is slower — 0.133 vs 0.108 ms — the score
    // registers stay live across it.)
    uint32_t v[KH];
    float mx_row = 0.f;
    auto softmax_load_max = [&](int i) {
      const int unit = blockIdx.x + i * gridDim.x;
      const int qt = unit_qt(unit, i);
      const int b = i & 1;
      const bool warp_live = (qt * ATT_BM + quad * 32 < T) && !(p.dbg & 1);   // warps past the sequence end skip the math
      timed_wait(&s_full[b], (i >> 1) & 1, 0);
      tc_fence_after();
      float mx = -INFINITY;
      if (warp_live) {
#pragma unroll
        for (int c = 0; c + 32 <= KH; c += 32)
          tmem_ld_32x32b_x32(lane_base + b * T_ + half * KH + c, *reinterpret_cast<uint32_t(*)[32]>(&v[c]));
        if constexpr (KH % 32 == 16)
          tmem_ld_32x32b_x16(lane_base + b * T_ + half * KH + (KH - 16), *reinterpret_cast<uint32_t(*)[16]>(&v[KH - 16]));
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < KH; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
        s_max[half][r] = mx;
      }
      tc_fence_before();            // our tcgen05.ld of S_b are complete (ordered before the p_full arrive below)
      asm volatile("bar.sync %0, %1;" ::"r"(2 + quad), "n"(32 * NSPLIT) : "memory");   // only the warps sharing these rows
      if (warp_live) {
#pragma unroll
        for (int o = 1; o < NSPLIT; ++o) mx = fmaxf(mx, s_max[(half + o) % NSPLIT][r]);
      }
      mx_row = mx;
    };
    auto softmax_exp = [&](int i) -> float {
      const int unit = blockIdx.x + i * gridDim.x;
      const int qt = unit_qt(unit, i);
      const bool warp_live = (qt * ATT_BM + quad * 32 < T) && !(p.dbg & 1);
      float sum = 0.f;
      if (warp_live) {
        // all exponentials first, packed in place over the score registers ...
        const float mx = mx_row;
        const float2 sc = make_float2(p.scale_log2e, p.scale_log2e);
        const float2 nm = make_float2(-mx * p.scale_log2e, -mx * p.scale_log2e);
        float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < KH / 2; ++j) {
          const float2 a = __ffma2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), sc, nm);
          const __nv_bfloat162 b2 = __floats2bfloat162_rn(fast_ex2(a.x), fast_ex2(a.y));
          sum2 = __fadd2_rn(sum2, __bfloat1622float2(b2));      // sum what the tensor core will see
          v[j] = *reinterpret_cast<const uint32_t*>(&b2);
        }
        sum = sum2.x + sum2.y;
        s_sum[half][r] = sum;
      }
      // ... and only then wait for the single P tile: it is free once P.V of the previous unit has completed, and that
      // MMA ran while the exponentials above were being computed
      if (i > 0) timed_wait(&o_full[(i - 1) % OB], ((i - 1) / OB) & 1, 1);
      if (warp_live) {
#pragma unroll
        for (int g = 0; g < KH / 8; ++g) {            // 8 keys = one 16-byte unit of the swizzled P tile
          const int col0 = half * KH + 8 * g;
          uint8_t* dst = s_p + (col0 / 64) * (ATT_BM * 128) + r * 128 + ((((col0 % 64) / 8) ^ (r & 7)) * 16);
          *reinterpret_cast<uint4*>(dst) = make_uint4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
      }
      fence_proxy_async_smem();     // P (generic-proxy writes) visible to the tensor core
      mbar_arrive(&p_full);
      asm volatile("bar.sync %0, %1;" ::"r"(2 + quad), "n"(32 * NSPLIT) : "memory");
      if (warp_live) {
#pragma unroll
        for (int o = 1; o < NSPLIT; ++o) sum += s_sum[(half + o) % NSPLIT][r];
        if (p.lse != nullptr && half == 0 && qt * ATT_BM + r < T) {
          const int head = (unit / q_tiles) % p.heads;
          const int crop = (unit / q_tiles) / p.heads;
          p.lse[(static_cast<size_t>(crop) * p.heads + head) * T + qt * ATT_BM + r] =
              fmaf(mx_row, p.scale_log2e, log2f(sum));
        }
      }
      return warp_live ? 1.0f / sum : 0.f;
    };

    auto epilogue = [&](int i, float inv) {
      const int unit = blockIdx.x + i * gridDim.x;
      const int qt = unit_qt(unit, i);
      const int head = (unit / q_tiles) % p.heads;
      const int crop = (unit / q_tiles) / p.heads;
      const int b = i % OB;
      const int token = qt * ATT_BM + r;
      const bool warp_live = (qt * ATT_BM + quad * 32 < T) && !(p.dbg & 1);
      timed_wait(&o_full[b], (i / OB) & 1, 2);
      tc_fence_after();
      if (warp_live) {
        // output columns per thread: an even split, or 48 + 32 for head_dim 80
        constexpr int OC0 = WIDE ? 48 : HD / NSPLIT;
        const int oc_begin = half * OC0;
        const int oc = WIDE ? (half == 0 ? 48 : HD - 48) : OC0;
        __nv_bfloat16* orow = p.out + (static_cast<size_t>(crop) * T + token) * p.ldo + head * HD + oc_begin;
#pragma unroll 1
        for (int c = 0; c < oc; c += 16) {
          uint32_t o[16];
          tmem_ld_32x32b_x16(lane_base + 2 * T_ + b * HD + oc_begin + c, o);
          tmem_ld_wait();
          if (token < T) {
            uint32_t w8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j)
              w8[j] = pack_bf16x2(__uint_as_float(o[2 * j]) * inv, __uint_as_float(o[2 * j + 1]) * inv);
            reinterpret_cast<uint4*>(orow + c)[0] = make_uint4(w8[0], w8[1], w8[2], w8[3]);
            reinterpret_cast<uint4*>(orow + c)[1] = make_uint4(w8[4], w8[5], w8[6], w8[7]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&o_free[b]);
    };

    float inv_cur = 0.f;
    if (n_local > 0) {
      softmax_load_max(0);
      inv_cur = softmax_exp(0);
    }
    for (int i = 0; i < n_local; ++i) {
      // softmax of the next unit first: its exp work overlaps P.V of unit i, whose result the epilogue then drains
      float inv_next = 0.f;
      if (i + 1 < n_local) {
        softmax_load_max(i + 1);
        inv_next = softmax_exp(i + 1);
      }
      epilogue(i, inv_cur);
      inv_cur = inv_next;
    }
    if (timing && threadIdx.x == 64) {
      for (int k = 0; k < 3; ++k) p.dbg_buf[6 + k] = t_wait[k];
      p.dbg_buf[9] = clock64() - t_start;
      p.dbg_buf[10] = n_local;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// -------------------------------------------------------------------------------------------------------------
// Ping-pong persistent attention (T = 192). Same unit walk, TMA producer and MMA issuer as the kernel above, but
//   * the probabilities never touch shared memory: a softmax thread owns one whole query row, reads its scores from
//     TMEM in 32-column chunks (pass 1: row max; pass 2: exp2, bf16 pack, row sum) and writes the packed bf16 pairs
//     back over the first half of the same score columns with tcgen05.st; P.V then takes its A operand from TMEM;
//   * two softmax groups of 128 threads (warps 2-5 and 6-9) work on alternate units with their own S / P / O columns,
//     so one group's exp2 phase (MUFU-bound) overlaps the other's TMEM loads, row max and output epilogue, and no
//     barrier or shared-memory exchange between threads of a row is left.
// TMEM: S0 | S1 | O0 | O1; for head_dim 80 (2T + 2*80 > 512) O_g lives in the upper, already-consumed half of S_g and
// S(i+2) is issued only after the epilogue of unit i has drained it.
// -------------------------------------------------------------------------------------------------------------
constexpr int PP_QK_DEPTH = 2;
__host__ __device__ constexpr int pp_v_depth(int hd) { return hd > 64 ? 2 : 3; }
// output staging: one block of 32 rows per softmax warp; a warp stages HD / RS columns of its rows (row pitch = bytes of
// those columns + 16: conflict-free 16-byte row writes)
__host__ __device__ constexpr int pp_stage_pitch(int hd, int rs = 1) { return hd * 2 / rs + 16; }
__host__ __device__ constexpr int pp_stage_bytes(int hd, int rs = 1) { return 8 * rs * 32 * pp_stage_pitch(hd, rs); }
constexpr int PP_THREADS = 64 + 256;
// RS = 2 ("row split", A/B variant, measured slower — see attention_fwd): TWO threads per query row in each softmax
// group (8 warps per group; warps w and w + 4 of a group share a TMEM lane quarter and take keys [0, T/2) and
// [T/2, T)), halving the per-thread work of the softmax and of the epilogue. The two threads of a row exchange their partial row maximum
// and row sum through shared memory (one 64-thread named barrier per lane quarter, twice per unit); P of the second
// half is written behind ITS OWN score columns ([T/2, T/2 + T/4)), so no thread overwrites scores another one still
// reads, and P.V takes its K steps from the two pieces. Layout: warps 0-3 TMA / MMA / 2 idle (one warpgroup, so that
// setmaxnreg can hand its registers to the sixteen softmax warps 4-19). (A separate epilogue warpgroup was tried first
// in round 2 and dropped: 150.4 vs 139.0 us, profiles/r02_summary.md.)
__host__ __device__ constexpr int pp_threads(int rs) { return rs == 2 ? 128 + 512 : PP_THREADS; }
__host__ __device__ constexpr int pp_first_softmax_warp(int rs) { return rs == 2 ? 4 : 2; }

template <int HD, int T_, int RS>
__global__ void __launch_bounds__(pp_threads(RS), 1)
attention_pingpong_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                          const __grid_constant__ CUtensorMap tm_qb, const __grid_constant__ CUtensorMap tm_kvb,
                          const AttnParams p, const int num_units) {
  static_assert(HD == 32 || HD == 64 || HD == 80, "ping-pong attention handles head_dim 32 / 64 / 80");
  constexpr bool WIDE = HD > 64;
  constexpr int T = T_;
  constexpr bool O_ALIAS = 2 * T + 2 * HD > 512;
  static_assert(T % 32 == 0 && 2 * T <= 512 && (!O_ALIAS || T / 2 + HD <= T), "S0 | S1 (| O0 | O1) must fit TMEM");
  constexpr int QA_BYTES = ATT_BM * 128, KVA_BYTES = T * 128;
  constexpr int Q_BYTES = ATT_BM * att2_row_bytes(HD), KV_BYTES = T * att2_row_bytes(HD);
  constexpr int QK_BYTES = Q_BYTES + KV_BYTES;
  constexpr int PP_V_DEPTH = pp_v_depth(HD);
  constexpr int PITCH = pp_stage_pitch(HD, RS);
  static_assert(RS == 1 || (RS == 2 && !O_ALIAS && T % 64 == 0 && HD % 32 == 0), "row split: head_dim 32 / 64 only");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_qk = smem;
  uint8_t* s_v = s_qk + PP_QK_DEPTH * QK_BYTES;
  uint8_t* s_stage = s_v + PP_V_DEPTH * KV_BYTES;
  __shared__ uint64_t qk_full[PP_QK_DEPTH], qk_free[PP_QK_DEPTH], v_full[PP_V_DEPTH], v_free[PP_V_DEPTH];
  __shared__ uint64_t s_full[2], p_full[2], o_full[2], o_free[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_xmax[2][2][RS == 2 ? ATT_BM : 1];   // RS = 2: partial row maximum [group][key half][row]
  __shared__ float s_xsum[2][2][RS == 2 ? ATT_BM : 1];   //         partial row sum

  const int q_tiles = (T + ATT_BM - 1) / ATT_BM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_local = num_units > static_cast<int>(blockIdx.x)
                          ? (num_units - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) /
                                static_cast<int>(gridDim.x)
                          : 0;
  // rotate the query tile with i so every CTA and both softmax groups get the same mix of full and partial tiles,
  // the two groups holding one of each at any time (group g sees (i + 1) / 2 = g, g + 1, g + 2, ...)
  const bool rotate_qt = (gridDim.x % q_tiles) == 0;
  auto unit_qt = [&](int unit, int i) { return rotate_qt ? (unit + ((i + 1) >> 1)) % q_tiles : unit % q_tiles; };
  // A partial last tile (64 live rows for T = 192) is placed alternately at the bottom (tile rows [0, 64), TMA zero
  // fill above) and at the top (the box starts at token T - 128: rows [64, 128) are the live ones, rows [0, 64) repeat
  // tokens of the full tile and are ignored), so that its softmax work alternates between warps 0-1 and 2-3 of a
  // group, i.e. between the SM's sub-partition pairs, instead of always loading the same two MUFU pipes.
  struct Placement { int q0, r_lo, r_hi; };
  auto unit_rows = [&](int qt, int i) {
    const int live = (T - qt * ATT_BM < ATT_BM) ? T - qt * ATT_BM : ATT_BM;
    const bool high = live < ATT_BM && T >= ATT_BM && ((i >> 2) & 1);
    Placement pl;
    pl.q0 = high ? T - ATT_BM : qt * ATT_BM;
    pl.r_lo = high ? ATT_BM - live : 0;
    pl.r_hi = high ? ATT_BM : live;
    return pl;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < PP_QK_DEPTH; ++s) { mbar_init(&qk_full[s], 1); mbar_init(&qk_free[s], 1); }
    for (int s = 0; s < PP_V_DEPTH; ++s) { mbar_init(&v_full[s], 1); mbar_init(&v_free[s], 1); }
    for (int g = 0; g < 2; ++g) {
      mbar_init(&s_full[g], 1);
      mbar_init(&p_full[g], 128 * RS);
      mbar_init(&o_full[g], 1);
      mbar_init(&o_free[g], 128 * RS);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    if (WIDE) {
      tma_prefetch_desc(&tm_qb);
      tma_prefetch_desc(&tm_kvb);
    }
  }
  if (warp == 1) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();                  // the set-up above overlapped the previous kernel's tail
  pdl_launch_dependents();
  // cycle accounting (CTA 0 only, one thread per role), enabled with VPB_ATT_DEBUG & 32; printed by & 64.
  // The counters are declared inside each role (PP_TIMING_STATE) so that they are not live across the
  // setmaxnreg-limited branches of the other roles.
  const bool timing = p.dbg_buf != nullptr && blockIdx.x == 0;
#define PP_TIMING_STATE                                                \
  long long t_acc[6] = {0, 0, 0, 0, 0, 0};                             \
  const long long t_start = timing ? clock64() : 0;                    \
  (void)t_start;                                                       \
  auto timed_wait = [&](uint64_t* bar, uint32_t parity, int slot) {    \
    if (timing) {                                                      \
      const long long t0 = clock64();                                  \
      mbar_wait(bar, parity);                                          \
      t_acc[slot] += clock64() - t0;                                   \
    } else {                                                           \
      mbar_wait(bar, parity);                                          \
    }                                                                  \
  }
  auto s_col = [&](int g) { return static_cast<uint32_t>(g * T); };
  auto o_col = [&](int g) { return static_cast<uint32_t>(O_ALIAS ? g * T + T / 2 : 2 * T + g * HD); };

  constexpr int W0 = pp_first_softmax_warp(RS);
  // RS = 2: 640 threads start with 96 registers each; the softmax threads (two 32-column chunks of scores in flight plus
  // the packed probabilities) get what the producer / issuer / idle warps do not need. Each setmaxnreg is the first
  // instruction of its warpgroup's branch, so that the register allocation of the branch follows it.
  if (warp < W0) {
  if constexpr (RS == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");   // 32 + 4 x 112 = 5 x 96: the CTA's own pool
  if (warp == 0) {
    if (lane == 0) {
      PP_TIMING_STATE;
      for (int i = 0; i < n_local; ++i) {
        const int unit = blockIdx.x + i * gridDim.x;
        const int qt = unit_qt(unit, i);
        const int head = (unit / q_tiles) % p.heads;
        const int crop = (unit / q_tiles) / p.heads;
        const int sq = i % PP_QK_DEPTH, sv = i % PP_V_DEPTH;
        const int q0 = unit_rows(qt, i).q0;
        timed_wait(&qk_free[sq], ((i / PP_QK_DEPTH) & 1) ^ 1, 0);
        mbar_arrive_expect_tx(&qk_full[sq], QK_BYTES);
        tma_load_3d(s_qk + sq * QK_BYTES, &tm_q, &qk_full[sq], head * HD, q0, crop);
        tma_load_3d(s_qk + sq * QK_BYTES + Q_BYTES, &tm_kv, &qk_full[sq], p.heads * HD + head * HD, 0, crop);
        if constexpr (WIDE) {     // columns [64, 80) of Q and K
          tma_load_3d(s_qk + sq * QK_BYTES + QA_BYTES, &tm_qb, &qk_full[sq], head * HD + 64, q0, crop);
          tma_load_3d(s_qk + sq * QK_BYTES + Q_BYTES + KVA_BYTES, &tm_kvb, &qk_full[sq], p.heads * HD + head * HD + 64,
                      0, crop);
        }
        timed_wait(&v_free[sv], ((i / PP_V_DEPTH) & 1) ^ 1, 1);
        mbar_arrive_expect_tx(&v_full[sv], KV_BYTES);
        tma_load_3d(s_v + sv * KV_BYTES, &tm_kv, &v_full[sv], 2 * p.heads * HD + head * HD, 0, crop);
        if constexpr (WIDE)
          tma_load_3d(s_v + sv * KV_BYTES + KVA_BYTES, &tm_kvb, &v_full[sv], 2 * p.heads * HD + head * HD + 64, 0, crop);
      }
      if (timing) { p.dbg_buf[0] = t_acc[0]; p.dbg_buf[1] = t_acc[1]; }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      PP_TIMING_STATE;
      constexpr uint32_t idesc_s = umma_idesc_bf16(ATT_BM, T);
      constexpr uint32_t idesc_o = umma_idesc_bf16(ATT_BM, WIDE ? 64 : HD, 0, 1);
      constexpr uint32_t idesc_o16 = umma_idesc_bf16(ATT_BM, 16, 0, 1);
      auto issue_s = [&](int i) {
        const int sq = i % PP_QK_DEPTH, g = i & 1;
        // S_g: its scores were consumed before p_full(i-2) and P(i-2) was read by P.V(i-2), issued earlier by this
        // thread; with O_g aliased into S_g the epilogue of unit i-2 must have drained it as well
        if constexpr (O_ALIAS) timed_wait(&o_free[g], ((i >> 1) & 1) ^ 1, 2);
        timed_wait(&qk_full[sq], (i / PP_QK_DEPTH) & 1, 0);
        tc_fence_after();
        const uint32_t qa = smem_u32(s_qk + sq * QK_BYTES), ka = qa + Q_BYTES;
        if (!(p.dbg & 4)) {
#pragma unroll
          for (int ks = 0; ks < (WIDE ? 4 : HD / 16); ++ks)
            umma_bf16_ss(tmem_base + s_col(g), umma_desc_k_sw128(qa + ks * 32), umma_desc_k_sw128(ka + ks * 32),
                         idesc_s, ks != 0);
          if constexpr (WIDE)
            umma_bf16_ss(tmem_base + s_col(g), umma_desc_k_sw32(qa + QA_BYTES), umma_desc_k_sw32(ka + KVA_BYTES),
                         idesc_s, 1u);
        }
        umma_commit(&s_full[g]);
        umma_commit(&qk_free[sq]);
      };
      if (n_local > 0) issue_s(0);
      if (n_local > 1) issue_s(1);
      for (int i = 0; i < n_local; ++i) {
        const int sv = i % PP_V_DEPTH, g = i & 1;
        timed_wait(&p_full[g], (i >> 1) & 1, 1);                               // P(i) in TMEM, S(i) consumed
        if constexpr (!O_ALIAS) timed_wait(&o_free[g], ((i >> 1) & 1) ^ 1, 2);    // O_g drained by the epilogue of unit i-2
        timed_wait(&v_full[sv], (i / PP_V_DEPTH) & 1, 3);
        tc_fence_after();
        const uint32_t d = tmem_base + o_col(g);
        const uint32_t pa = tmem_base + s_col(g);
        const uint32_t va = smem_u32(s_v + sv * KV_BYTES);
#pragma unroll
        for (int ks = 0; ks < ((p.dbg & 2) ? 0 : T / 16); ++ks) {
          // P (bf16 pairs, 8 columns per K = 16 step): one piece at column 0, or with RS = 2 one behind each half of the
          // score columns (keys [0, T/2) at column 0, keys [T/2, T) at column T/2)
          const uint32_t pcol = (RS == 2 && ks >= T / 32) ? T / 2 + (ks - T / 32) * 8 : ks * 8;
          umma_bf16_ts(d, pa + pcol, umma_desc_mn_sw128(va + ks * 2048, KVA_BYTES), idesc_o, ks != 0);
          if constexpr (WIDE)
            umma_bf16_ts(d + 64, pa + pcol, umma_desc_mn_sw32(va + KVA_BYTES + ks * 512), idesc_o16, ks != 0);
        }
        umma_commit(&o_full[g]);
        umma_commit(&v_free[sv]);
        if (i + 2 < n_local) issue_s(i + 2);
      }
      if (timing) { for (int k = 0; k < 4; ++k) p.dbg_buf[2 + k] = t_acc[k]; }
    }
  }
  } else {
    if constexpr (RS == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    PP_TIMING_STATE;
    const int g = (warp - W0) / (4 * RS);           // softmax group: units i = g, g + 2, ...
    const int half = RS == 2 ? ((warp - W0) >> 2) & 1 : 0;   // RS = 2: which half of the keys / of the output columns
    const int quad = warp & 3;                      // TMEM lane quarter this warp may access
    const int r = quad * 32 + lane;                 // query row inside the tile
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    constexpr int NC = T / 32 / RS;                 // 32-column score chunks of this thread
    constexpr int CB = T / RS;                      // RS = 2: the second thread of a row starts at score column T / 2
    constexpr int OC = HD / RS;                     // output columns of this thread
    const uint32_t S = lane_base + s_col(g) + half * CB, O = lane_base + o_col(g) + half * OC;
    const int xbar = 1 + g * 4 + quad;              // RS = 2: named barrier of the two warps that share this lane quarter

    // The staged output block of a unit is written to global memory one unit later, while this group waits for the
    // next P.V to complete (it has nothing else to do then), instead of at the end of its own epilogue.
    uint8_t* const stage = s_stage + (warp - W0) * 32 * PITCH;
    __nv_bfloat16* pend_base = nullptr;       // null: nothing pending
    int pend_lo = 0, pend_hi = 0;
    auto flush_pending = [&]() {
      if (pend_base != nullptr && !(p.dbg & 8)) {
        constexpr int CH = OC / 8;            // 16-byte pieces per (part of a) row
#pragma unroll
        for (int e = lane; e < 32 * CH; e += 32) {
          const int row = e / CH, ch = e - row * CH;
          const int rr = quad * 32 + row;
          if (rr >= pend_lo && rr < pend_hi)
            *reinterpret_cast<uint4*>(pend_base + static_cast<size_t>(row) * p.ldo + ch * 8) =
                *reinterpret_cast<const uint4*>(stage + row * PITCH + ch * 16);
        }
        __syncwarp();                         // the block may be overwritten by the next epilogue
      }
      pend_base = nullptr;
    };
    const long long t_loop = timing ? clock64() : 0;
    long long t_body = 0;
    // VPB_ATT_DEBUG & 2048 (experiment): start group 1 half a unit late, so that the two groups' exp2 phases (which
    // share the four MUFU pipes) alternate instead of coinciding
    for (int i = g; i < n_local; i += 2) {
      const long long t_top = timing ? clock64() : 0;
      const int unit = blockIdx.x + i * gridDim.x;
      const int qt = unit_qt(unit, i);
      const int head = (unit / q_tiles) % p.heads;
      const int crop = (unit / q_tiles) / p.heads;
      const Placement pl = unit_rows(qt, i);
      const int token = pl.q0 + r;
      const bool rows_here = quad * 32 < pl.r_hi && quad * 32 + 32 > pl.r_lo;
      const bool warp_live = rows_here && !(p.dbg & 1);                         // warps without live rows skip the math
      const bool row_live = r >= pl.r_lo && r < pl.r_hi;
      const uint32_t ph = (i >> 1) & 1;
      float inv = 0.f;
      const bool tlive = timing && warp_live;      // counters: s_full wait, softmax math, o_full wait, epilogue
      timed_wait(&s_full[g], ph, warp_live ? 0 : 4);
      if ((p.dbg & 2048) && g == 1 && i == 1) __nanosleep(900);   // (experiment, see above; after the first S is there)
      long long t_mark = tlive ? clock64() : 0;
      tc_fence_after();
      if (warp_live) {
        uint32_t va[32], vb[32];
        // pass 1: row maximum
        float mx = -INFINITY;
        if (p.dbg & 1024) mx = 0.f;      // timing experiment (results invalid): no row-max pass over the scores
        else {
        tmem_ld_32x32b_x32(S, va);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          uint32_t(&cur)[32] = (c & 1) ? vb : va;
          uint32_t(&nxt)[32] = (c & 1) ? va : vb;
          tmem_ld_wait();
          if (c + 1 < NC) tmem_ld_32x32b_x32(S + 32 * (c + 1), nxt);
          float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            m0 = fmaxf(m0, fmaxf(__uint_as_float(cur[j]), __uint_as_float(cur[j + 1])));
            m1 = fmaxf(m1, fmaxf(__uint_as_float(cur[j + 2]), __uint_as_float(cur[j + 3])));
          }
          mx = fmaxf(mx, fmaxf(m0, m1));
        }
        }
        if constexpr (RS == 2) {             // the other half of the keys: exchange the partial maxima
          s_xmax[g][half][r] = mx;
          asm volatile("bar.sync %0, 64;" ::"r"(xbar) : "memory");
          mx = fmaxf(mx, s_xmax[g][half ^ 1][r]);
        }
        // pass 2: p = exp2(s * scale*log2e - max'), bf16 pairs written back over columns [16c, 16c + 16) — always
        // behind the columns still to be read
        const float2 sc = make_float2(p.scale_log2e, p.scale_log2e);
        const float2 nm = make_float2(-mx * p.scale_log2e, -mx * p.scale_log2e);
        float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
        tmem_ld_32x32b_x32(S, va);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          uint32_t(&cur)[32] = (c & 1) ? vb : va;
          uint32_t(&nxt)[32] = (c & 1) ? va : vb;
          tmem_ld_wait();
          if (c + 1 < NC) tmem_ld_32x32b_x32(S + 32 * (c + 1), nxt);
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 a = __ffma2_rn(make_float2(__uint_as_float(cur[2 * j]), __uint_as_float(cur[2 * j + 1])), sc, nm);
            const float2 e = make_float2(fast_ex2(a.x), fast_ex2(a.y));
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(e.x, e.y);
            // row sum of the unrounded values: two instructions per pair fewer than unpacking the bf16 pair again; the
            // difference to the sum of the rounded ones is ~2^-9 / sqrt(T) relative, far below the bf16 output rounding
            if (j & 1) sum_b = __fadd2_rn(sum_b, e);
            else       sum_a = __fadd2_rn(sum_a, e);
            pk[j] = *reinterpret_cast<const uint32_t*>(&b2);
          }
          tmem_st_32x32b_x16(S + 16 * c, pk);
        }
        tmem_st_wait();
        float sum = (sum_a.x + sum_b.x) + (sum_a.y + sum_b.y);
        if constexpr (RS == 2) {             // partial row sums (the barrier also orders the reads of s_xmax above
          s_xsum[g][half][r] = sum;          // before the next unit's writes)
          asm volatile("bar.sync %0, 64;" ::"r"(xbar) : "memory");
          const float other = s_xsum[g][half ^ 1][r];
          sum = half == 0 ? sum + other : other + sum;      // the same association in both threads of the row
        }
        inv = 1.0f / sum;
        if (p.lse != nullptr && row_live && half == 0)
          p.lse[(static_cast<size_t>(crop) * p.heads + head) * T + token] = fmaf(mx, p.scale_log2e, log2f(sum));
      }
      tc_fence_before();            // our tcgen05.ld / st of S_g are complete and ordered before the arrive
      mbar_arrive(&p_full[g]);
      if (tlive) { t_acc[1] += clock64() - t_mark; t_acc[5] += 1; }
      flush_pending();              // previous unit's output rows, while P.V of this one runs

      // epilogue: O_g / row sum -> bf16, one whole output row (HD * 2 bytes, contiguous) per thread
      timed_wait(&o_full[g], ph, warp_live ? 2 : 4);
      t_mark = tlive ? clock64() : 0;
      tc_fence_after();
      if (warp_live) {
        uint32_t o[OC];               // all loads in flight at once, one wait
#pragma unroll
        for (int c = 0; c < OC; c += 16) tmem_ld_32x32b_x16(O + c, *reinterpret_cast<uint32_t(*)[16]>(&o[c]));
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&o_free[g]);      // O_g is in registers: the next P.V (or S, when aliased) may overwrite it now
        // A thread owns a row, but a row-per-thread store touches 32 different 128-byte lines with 16 bytes each per
        // instruction (measured: 38 of 115 us). Stage the warp's 32 rows in shared memory and write them back with
        // consecutive lanes on consecutive 16-byte pieces of a row: whole lines per instruction.
#pragma unroll
        for (int c = 0; c < OC; c += 8) {
          uint32_t w4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            w4[j] = pack_bf16x2(__uint_as_float(o[c + 2 * j]) * inv, __uint_as_float(o[c + 2 * j + 1]) * inv);
          *reinterpret_cast<uint4*>(stage + lane * PITCH + c * 2) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        }
        __syncwarp();
        pend_base = p.out + (static_cast<size_t>(crop) * T + pl.q0 + quad * 32) * p.ldo + head * HD + half * OC;
        pend_lo = pl.r_lo;
        pend_hi = pl.r_hi;
      } else {
        tc_fence_before();
        mbar_arrive(&o_free[g]);
      }
      if (tlive) { t_acc[3] += clock64() - t_mark; }
      if (timing) t_body += clock64() - t_top;
    }
    flush_pending();
    if (timing && threadIdx.x == 32 * W0 + 64) {
      for (int k = 0; k < 4; ++k) p.dbg_buf[6 + k] = t_acc[k];
      p.dbg_buf[10] = clock64() - t_start;
      p.dbg_buf[11] = n_local;
      p.dbg_buf[12] = t_acc[4];
      p.dbg_buf[13] = t_acc[5];
      p.dbg_buf[14] = t_loop - t_start;
      p.dbg_buf[15] = t_body;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

#undef PP_TIMING_STATE

template <int HD, int T_, int RS>
static int launch_attention_pingpong(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& tqb,
                                     const CUtensorMap& tkvb, const AttnParams& p, int max_ctas, cudaStream_t stream) {
  constexpr int smem = PP_QK_DEPTH * (ATT_BM + T_) * att2_row_bytes(HD) + pp_v_depth(HD) * T_ * att2_row_bytes(HD) +
                       pp_stage_bytes(HD, RS) + 1024;
  static_assert(smem <= 227 * 1024 - 6 * 1024, "ping-pong attention tiles do not fit shared memory");
  auto kern = attention_pingpong_kernel<HD, T_, RS>;
  static bool configured = false;
  if (!configured) {
    VPB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int q_tiles = (T_ + ATT_BM - 1) / ATT_BM;
  const int units = p.n * p.heads * q_tiles;
  int grid = max_ctas > 0 ? max_ctas : sm_count();
  if (grid > units) grid = units;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(pp_threads(RS));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = pdl_launch_attr(&attr[0]);
  VPB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tq, tkv, tqb, tkvb, p, units));
  return 0;
}

template <int HD, int T_, int NSPLIT>
static int launch_attention_persistent(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& tqb,
                                       const CUtensorMap& tkvb, const AttnParams& p, int max_ctas,
                                       cudaStream_t stream) {
  constexpr int smem = (T_ / 64) * ATT_BM * 128 + ATT2_QK_DEPTH * (ATT_BM + T_) * att2_row_bytes(HD) +
                       att2_v_depth(HD) * T_ * att2_row_bytes(HD) + 1024;
  static_assert(smem <= 227 * 1024 - 6 * 1024, "persistent attention tiles do not fit shared memory");
  auto kern = attention_persistent_kernel<HD, T_, NSPLIT>;
  static bool configured = false;
  if (!configured) {
    VPB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int q_tiles = (T_ + ATT_BM - 1) / ATT_BM;
  const int units = p.n * p.heads * q_tiles;
  int grid = max_ctas > 0 ? max_ctas : sm_count();
  if (grid > units) grid = units;
  kern<<<grid, att2_threads(NSPLIT), smem, stream>>>(tq, tkv, tqb, tkvb, p, units);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int HD>
static int launch_attention(const CUtensorMap& tq, const CUtensorMap& tkv, const AttnParams& p, int smem,
                            cudaStream_t stream) {
  auto kern = attention_kernel<HD>;
  static int configured_smem = 0;
  if (configured_smem < smem) {
    VPB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  kern<<<p.n * p.heads * 2, ATT_THREADS, smem, stream>>>(tq, tkv, p);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int attention_fwd(const void* qkv, void* out, int n, int T, int heads, int hd, float scale, int max_ctas,
                  cudaStream_t stream, float* lse) {
  (void)max_ctas;
  VPB_REQUIRE(n > 0 && heads > 0, "attention: empty problem");
  VPB_REQUIRE(T % 16 == 0 && T >= 16 && T <= 256, "attention: T=%d must be a multiple of 16 in [16,256]", T);
  VPB_REQUIRE(hd % 16 == 0 && hd >= 16 && hd <= 128, "attention: head_dim=%d unsupported", hd);
  const int ld = 3 * heads * hd;
  VPB_REQUIRE(ld % 8 == 0, "attention: row pitch must be a multiple of 16 bytes");
  CUtensorMap tq, tkv;
  uint64_t dims[3] = {(uint64_t)ld, (uint64_t)T, (uint64_t)n};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)T * ld * 2};
  uint32_t box_q[3] = {64, ATT_BM, 1};
  uint32_t box_kv[3] = {64, (uint32_t)T, 1};
  if (make_tma_desc(&tq, TMA_BF16, qkv, 3, dims, strides, box_q, TMA_SWIZZLE_128B)) return -1;
  if (make_tma_desc(&tkv, TMA_BF16, qkv, 3, dims, strides, box_kv, TMA_SWIZZLE_128B)) return -1;
  AttnParams p;
  p.n = n; p.T = T; p.heads = heads; p.hd = hd; p.ldo = heads * hd;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.lse = lse;
  {
    const char* e = getenv("VPB_ATT_DEBUG");
    p.dbg = e ? atoi(e) : 0;
    static long long* dbg_buf = nullptr;
    if ((p.dbg & 32) && dbg_buf == nullptr) {
      VPB_CHECK_CUDA(cudaMallocManaged(&dbg_buf, 16 * sizeof(long long)));
    }
    p.dbg_buf = (p.dbg & 32) ? dbg_buf : nullptr;
    if (p.dbg & 64) {   // print the counters of the previous launch
      cudaStreamSynchronize(stream);
      if (dbg_buf)
        fprintf(stderr, "att cycles (CTA0): producer qk_free %lld v_free %lld | mma qk_full %lld p_full %lld o_free %lld "
                        "v_full %lld | softmax thread 128, live units: s_full %lld math %lld o_full %lld epilogue %lld "
                        "(%lld units), idle units %lld | total %lld, %lld units, loop entry at %lld, loop bodies %lld\n",
                dbg_buf[0], dbg_buf[1], dbg_buf[2], dbg_buf[3], dbg_buf[4], dbg_buf[5], dbg_buf[6], dbg_buf[7],
                dbg_buf[8], dbg_buf[9], dbg_buf[13], dbg_buf[12], dbg_buf[10], dbg_buf[11], dbg_buf[14], dbg_buf[15]);
    }
  }
  const int nb = att_boxes(hd);
  const int qk = nb * (ATT_BM * 128 + T * 128);
  const int pb = ((T + 63) / 64) * ATT_BM * 128;
  const int region0 = ((qk > pb ? qk : pb) + 1023) & ~1023;
  const int smem = region0 + nb * T * 128 + 1024;
  if (max_ctas >= 0) {   // max_ctas < 0 selects the per-unit kernel (kept for head_dim > 64 and for A/B tests)
    int rc = 1;
    const bool pingpong = !(p.dbg & 256);     // VPB_ATT_DEBUG & 256: the single-group kernel (P through smem), for A/B
    // VPB_ATT_DEBUG & 512: two threads per query row (RS = 2), for A/B. Measured SLOWER (B200, 512 image passes:
    // head_dim 64 166.1 vs 157.3 us, head_dim 32 211.6 vs 199.3 us): with half the work per thread the softmax phase of a
    // unit still takes ~3200 cycles — while both groups are in it the four MUFU pipes are saturated (2 x 24576 exp2 per
    // 3072 cycles at 16 / clk / SM), so more warps only add exchange barriers.
    const bool split = (p.dbg & 512) != 0;
    if (pingpong && hd == 32 && T == 192)
      rc = split ? launch_attention_pingpong<32, 192, 2>(tq, tkv, tq, tkv, p, max_ctas, stream)
                 : launch_attention_pingpong<32, 192, 1>(tq, tkv, tq, tkv, p, max_ctas, stream);
    if (pingpong && hd == 64 && T == 192)
      rc = split ? launch_attention_pingpong<64, 192, 2>(tq, tkv, tq, tkv, p, max_ctas, stream)
                 : launch_attention_pingpong<64, 192, 1>(tq, tkv, tq, tkv, p, max_ctas, stream);
    // head_dim 80: O_g aliases the consumed half of S_g (2T + 2*80 > 512 TMEM columns), so S(i+2) is issued after the
    // epilogue of unit i has loaded O_g into registers; still 65.6 us vs 85.8 us for the single-group kernel at 128
    // image passes once the output stores were deferred
    if (pingpong && hd == 80 && T == 192) {
      CUtensorMap tqb, tkvb;
      uint32_t box_qb[3] = {16, ATT_BM, 1}, box_kvb[3] = {16, (uint32_t)T, 1};
      if (make_tma_desc(&tqb, TMA_BF16, qkv, 3, dims, strides, box_qb, TMA_SWIZZLE_32B)) return -1;
      if (make_tma_desc(&tkvb, TMA_BF16, qkv, 3, dims, strides, box_kvb, TMA_SWIZZLE_32B)) return -1;
      rc = launch_attention_pingpong<80, 192, 1>(tq, tkv, tqb, tkvb, p, max_ctas, stream);
    }
    if (rc <= 0) return rc;
    if (hd == 32 && T == 192) rc = launch_attention_persistent<32, 192, 2>(tq, tkv, tq, tkv, p, max_ctas, stream);
    if (hd == 64 && T == 192) {
      // measured at 256 images x 12 heads: 2 column parts per row (8 softmax warps) 0.109 ms, 4 parts (16 warps) 0.134 ms
      rc = (p.dbg & 16) ? launch_attention_persistent<64, 192, 4>(tq, tkv, tq, tkv, p, max_ctas, stream)
                        : launch_attention_persistent<64, 192, 2>(tq, tkv, tq, tkv, p, max_ctas, stream);
    }
    if (hd == 80 && T == 192 && !(p.dbg & 128)) {
      // head_dim 80 (ViT-H): a second, 16-column SWIZZLE_32B box per operand
      CUtensorMap tqb, tkvb;
      uint32_t box_qb[3] = {16, ATT_BM, 1}, box_kvb[3] = {16, (uint32_t)T, 1};
      if (make_tma_desc(&tqb, TMA_BF16, qkv, 3, dims, strides, box_qb, TMA_SWIZZLE_32B)) return -1;
      if (make_tma_desc(&tkvb, TMA_BF16, qkv, 3, dims, strides, box_kvb, TMA_SWIZZLE_32B)) return -1;
      rc = launch_attention_persistent<80, 192, 2>(tq, tkv, tqb, tkvb, p, max_ctas, stream);
    }
    if (rc <= 0) return rc;
  }
  VPB_REQUIRE(lse == nullptr, "attention: the log-sum-exp output needs the persistent kernel (T=192, head_dim 32/64/80)");
  switch (hd) {
    case 32: return launch_attention<32>(tq, tkv, p, smem, stream);
    case 64: return launch_attention<64>(tq, tkv, p, smem, stream);
    case 80: return launch_attention<80>(tq, tkv, p, smem, stream);
    case 128: return launch_attention<128>(tq, tkv, p, smem, stream);
    default:
      set_last_error("attention: no kernel instance for head_dim=%d", hd);
      return -2;
  }
}

}  // namespace vpb
