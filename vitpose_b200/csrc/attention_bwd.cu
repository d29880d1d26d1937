// Backward of the fused multi-head self-attention (Attention.forward, mmpose/models/backbones/vit.py:99-115) for the
// training-step configuration (ViTPose-B: T = 192 tokens, head_dim 64).
//   S = (Q K^T) * scale, P = softmax(S), O = P V            (forward, attention.cu)
//   dV = P^T dO,  dP = dO V^T,  dS = P o (dP - delta),  delta_i = sum_d dO_id O_id,
//   dQ = scale * dS K,  dK = scale * dS^T Q
// One CTA per (crop, head). Q, K, V, dO of the head are TMA-loaded once as 128B-swizzled [192 x 64] tiles; the
// query rows are processed in two 128-row tiles (the second is half empty). Per tile:
//   tcgen05.mma S  = Q_t K^T       -> TMEM [0,192)   ; 256 threads ((query row, half of the keys) each) form
//                                                       P = exp2(S * scale * log2e - lse) with the forward pass's
//                                                       log-sum-exp in ONE pass and write it (bf16) to smem
//   tcgen05.mma dP = dO_t V^T      -> same TMEM columns; the threads form dS = scale * P o (dP - delta) (bf16, smem)
//   tcgen05.mma dQ_t = dS K        -> TMEM [192,256)  (K consumed as an MN-major operand)
//   tcgen05.mma dK += dS^T Q_t, dV += P^T dO_t -> TMEM [256,512): A operands are the P / dS tiles read MN-major
//                                                 (transposed) straight from where the threads wrote them
// Nothing of size T x T ever touches HBM. Every MMA is M = 128: rows past the sequence end compute garbage from
// whatever follows the tile in shared memory, land in TMEM lanes that are never read, and never enter a contraction.
#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

namespace vpb {

constexpr int AB_T = 192;
constexpr int AB_HD = 64;
constexpr int AB_THREADS = 288;                     // warp 0: TMA + MMA issue; warps 1..8: (row, column half) per thread
constexpr int AB_TILE = AB_T * 128;                 // [192 rows][128 B]
constexpr int AB_CHUNK = 128 * 128;                 // [128 rows][64 keys] bf16
constexpr int AB_SMEM = 4 * AB_TILE + 3 * AB_CHUNK + 4 * AB_CHUNK + 1024;   // Q K V dO | dS | P + one chunk of slack

struct AttnBwdParams {
  int n, heads;
  float scale, scale_log2e;
  const __nv_bfloat16* dO;    // [n, T, heads*64]
  const __nv_bfloat16* O;     // [n, T, heads*64] forward output
  const float* lse;           // [n, heads, T] log2-sum-exp of the scaled scores, written by the forward kernel
  __nv_bfloat16* dqkv;        // [n, T, 3*heads*64]
  float* dbias;               // optional [3*heads*64]: += column sums of dqkv over all tokens (attn.qkv's bias gradient)
};

__global__ void __launch_bounds__(AB_THREADS, 1)
attention_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                     const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_q = smem;
  uint8_t* s_k = s_q + AB_TILE;
  uint8_t* s_v = s_k + AB_TILE;
  uint8_t* s_do = s_v + AB_TILE;
  uint8_t* s_ds = s_do + AB_TILE;                   // 3 chunks of [128 q][64 keys]
  uint8_t* s_p = s_ds + 3 * AB_CHUNK;               // 3 chunks (+ 1 chunk of slack read by the M=128 key tile 1)
  __shared__ uint64_t bar_load, bar_s, bar_sdone, bar_dp, bar_ds, bar_mma;
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x % p.heads;
  const int crop = blockIdx.x / p.heads;
  const int ld_o = p.heads * AB_HD, ld_qkv = 3 * ld_o;

  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_sdone, 256);
    mbar_init(&bar_dp, 1);
    mbar_init(&bar_ds, 256);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_qkv);
    tma_prefetch_desc(&tm_do);
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_slot;                // S / dP
  const uint32_t tmem_dq = tmem_s + 192;
  const uint32_t tmem_dk = tmem_s + 256;            // two key tiles x 64 columns
  const uint32_t tmem_dv = tmem_s + 384;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bar_load, 4 * AB_TILE);
      tma_load_3d(s_q, &tm_qkv, &bar_load, head * AB_HD, 0, crop);
      tma_load_3d(s_k, &tm_qkv, &bar_load, ld_o + head * AB_HD, 0, crop);
      tma_load_3d(s_v, &tm_qkv, &bar_load, 2 * ld_o + head * AB_HD, 0, crop);
      tma_load_3d(s_do, &tm_do, &bar_load, head * AB_HD, 0, crop);
      mbar_wait(&bar_load, 0);
      tc_fence_after();
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, AB_T);            // K-major A and B
      constexpr uint32_t idesc_dq = umma_idesc_bf16(128, AB_HD, 0, 1);    // B (K) MN-major
      constexpr uint32_t idesc_t = umma_idesc_bf16(128, AB_HD, 1, 1);     // A (P^T / dS^T) and B MN-major
      for (int t = 0; t < 2; ++t) {
        const uint32_t q_t = smem_u32(s_q) + t * AB_CHUNK, do_t = smem_u32(s_do) + t * AB_CHUNK;
        // S = Q_t K^T
#pragma unroll
        for (int ks = 0; ks < AB_HD / 16; ++ks)
          umma_bf16_ss(tmem_s, umma_desc_k_sw128(q_t + ks * 32), umma_desc_k_sw128(smem_u32(s_k) + ks * 32), idesc_s,
                       ks != 0);
        umma_commit(&bar_s);
        // dP = dO_t V^T into the same columns once every thread has consumed S
        mbar_wait(&bar_sdone, t);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < AB_HD / 16; ++ks)
          umma_bf16_ss(tmem_s, umma_desc_k_sw128(do_t + ks * 32), umma_desc_k_sw128(smem_u32(s_v) + ks * 32), idesc_s,
                       ks != 0);
        umma_commit(&bar_dp);
        mbar_wait(&bar_ds, t);
        tc_fence_after();
        // dQ_t = dS K (contraction over the 192 keys)
        for (int ks = 0; ks < AB_T / 16; ++ks)
          umma_bf16_ss(tmem_dq, umma_desc_k_sw128(smem_u32(s_ds) + (ks / 4) * AB_CHUNK + (ks % 4) * 32),
                       umma_desc_mn_sw128(smem_u32(s_k) + ks * 2048, AB_TILE), idesc_dq, ks != 0);
        // dK += dS^T Q_t, dV += P^T dO_t (contraction over the valid queries of this tile)
        const int nks = t == 0 ? 8 : (AB_T - 128) / 16;
        for (int m = 0; m < 2; ++m) {
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t acc = (t | ks) != 0 ? 1u : 0u;
            umma_bf16_ss(tmem_dk + m * AB_HD,
                         umma_desc_mn_sw128(smem_u32(s_ds) + 2 * m * AB_CHUNK + ks * 2048, AB_CHUNK),
                         umma_desc_mn_sw128(q_t + ks * 2048, AB_TILE), idesc_t, acc);
            umma_bf16_ss(tmem_dv + m * AB_HD,
                         umma_desc_mn_sw128(smem_u32(s_p) + 2 * m * AB_CHUNK + ks * 2048, AB_CHUNK),
                         umma_desc_mn_sw128(do_t + ks * 2048, AB_TILE), idesc_t, acc);
          }
        }
        umma_commit(&bar_mma);
        // The next tile's S is issued right away: its TMEM columns are free since bar_ds (every thread has read dP), the
        // tensor pipe runs it behind the MMAs above, and the threads rewrite the P / dS tiles only after bar_mma — so
        // S of tile 1 is ready while the threads still drain dQ of tile 0.
      }
    }
  } else {
    const int quad = warp & 3;
    const int half = (warp - 1) >> 2;               // which half of the key columns (and of the output columns)
    const int r = quad * 32 + lane;                 // row inside a tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    constexpr int KH = AB_T / 2;                    // 96 keys per thread
    // delta = <dO_row, O_row> and the forward log-sum-exp of this thread's row in BOTH query tiles, requested up front:
    // the global-memory latency (16 x 16-byte loads per tile; 20 % of the kernel's stall samples when they sat at the
    // top of each tile) hides behind the TMA loads of Q / K / V / dO. Both halves of a row compute it: no exchange.
    float delta_t[2] = {0.f, 0.f}, lse_t[2] = {0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int token = t * 128 + r;
      if (token < AB_T) {
        const size_t off = (static_cast<size_t>(crop) * AB_T + token) * ld_o + head * AB_HD;
        const uint4* a = reinterpret_cast<const uint4*>(p.dO + off);
        const uint4* b = reinterpret_cast<const uint4*>(p.O + off);
        float d = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint4 x = __ldg(a + u), y = __ldg(b + u);
          const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, yw[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 fx = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xw[j]));
            const float2 fy = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yw[j]));
            d = fmaf(fx.x, fy.x, fmaf(fx.y, fy.y, d));
          }
        }
        delta_t[t] = d;
        lse_t[t] = __ldg(p.lse + (static_cast<size_t>(crop) * p.heads + head) * AB_T + token);
      }
    }
    for (int t = 0; t < 2; ++t) {
      const int token = t * 128 + r;
      const bool valid = token < AB_T;
      const float delta = t == 0 ? delta_t[0] : delta_t[1], lse = t == 0 ? lse_t[0] : lse_t[1];
      mbar_wait(&bar_s, t);
      tc_fence_after();
      // P = exp2(s * scale * log2e - lse), one pass over this thread's 96 keys
      for (int c = half * KH; c < (half + 1) * KH; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_s + lane_off + c, v);
        tmem_ld_wait();
        uint32_t packed[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float e0 = valid ? exp2f(fmaf(__uint_as_float(v[j]), p.scale_log2e, -lse)) : 0.f;
          const float e1 = valid ? exp2f(fmaf(__uint_as_float(v[j + 1]), p.scale_log2e, -lse)) : 0.f;
          packed[j / 2] = pack_bf16x2(e0, e1);
        }
        const uint32_t row = smem_u32(s_p) + (c / 64) * AB_CHUNK + r * 128;
        const int u0 = (c % 64) / 8;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          sts_u4(row + (((u0 + u) ^ (r & 7)) * 16), packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
      }
      tc_fence_before();               // all tcgen05.ld of S done before dP overwrites the columns
      mbar_arrive(&bar_sdone);
      mbar_wait(&bar_dp, t);
      tc_fence_after();
      // dS = scale * P o (dP - delta)
      for (int c = half * KH; c < (half + 1) * KH; c += 32) {
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
            o[j] = valid ? pack_bf16x2(d0, d1) : 0u;
          }
          sts_u4(drow + so, o[0], o[1], o[2], o[3]);
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();         // P and dS (generic-proxy writes) visible to the tensor core
      mbar_arrive(&bar_ds);
      mbar_wait(&bar_mma, t);
      tc_fence_after();
      // dQ rows of this tile: 32 of the 64 columns per thread. A row-per-thread global store would touch 32 lines
      // with 16 bytes each per instruction, so the warp stages its 32 x 64-byte block in shared memory (the Q rows
      // of tile 0, dead once bar_mma(0) has completed; 16-byte pieces XOR-swizzled: conflict-free both ways) and
      // writes it back with four lanes per row.
      {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_dq + lane_off + half * 32, v);
        tmem_ld_wait();
        if (p.dbias != nullptr) {       // bias gradient of attn.qkv: column sums over the live rows of this warp
          float cs[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) cs[j] = valid ? __uint_as_float(v[j]) : 0.f;
          const float sum = warp_colsum32(cs, lane);
          atomicAdd(p.dbias + head * AB_HD + half * 32 + lane, sum);
        }
        uint8_t* stage = s_q + (warp - 1) * 2048;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<uint4*>(stage + lane * 64 + ((u ^ ((lane >> 1) & 3)) * 16)) =
              make_uint4(pack_bf16x2(__uint_as_float(v[8 * u]), __uint_as_float(v[8 * u + 1])),
                         pack_bf16x2(__uint_as_float(v[8 * u + 2]), __uint_as_float(v[8 * u + 3])),
                         pack_bf16x2(__uint_as_float(v[8 * u + 4]), __uint_as_float(v[8 * u + 5])),
                         pack_bf16x2(__uint_as_float(v[8 * u + 6]), __uint_as_float(v[8 * u + 7])));
        __syncwarp();
        __nv_bfloat16* obase = p.dqkv + (static_cast<size_t>(crop) * AB_T + t * 128 + quad * 32) * ld_qkv +
                               head * AB_HD + half * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int row = 8 * j + (lane >> 2), ch = lane & 3;
          if (t * 128 + quad * 32 + row < AB_T)
            *reinterpret_cast<uint4*>(obase + static_cast<size_t>(row) * ld_qkv + ch * 8) =
                *reinterpret_cast<const uint4*>(stage + row * 64 + ((ch ^ ((row >> 1) & 3)) * 16));
        }
        __syncwarp();
      }
      tc_fence_before();
    }
    // dK / dV rows (keys): key tile m, lane r <-> key m*128 + r; half 0 stores dK, half 1 stores dV. Every operand tile
    // is dead by now (all MMAs have completed): the warp stages its 32 x 128-byte block in the K/V/dO area (row pitch
    // 144 bytes) and stores whole 128-byte lines, eight lanes per row.
    for (int m = 0; m < 2; ++m) {
      const uint32_t base = (half == 0 ? tmem_dk : tmem_dv) + m * AB_HD + lane_off;
      uint8_t* stage = s_k + (warp - 1) * (32 * 144);
#pragma unroll
      for (int c = 0; c < AB_HD; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(base + c, v);
        tmem_ld_wait();
        if (p.dbias != nullptr) {
          const bool key_ok = m * 128 + quad * 32 + lane < AB_T;
          float cs[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) cs[j] = key_ok ? __uint_as_float(v[j]) : 0.f;
          const float sum = warp_colsum32(cs, lane);
          atomicAdd(p.dbias + (1 + half) * ld_o + head * AB_HD + c + lane, sum);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          *reinterpret_cast<uint4*>(stage + lane * 144 + c * 2 + u * 16) =
              make_uint4(pack_bf16x2(__uint_as_float(v[8 * u]), __uint_as_float(v[8 * u + 1])),
                         pack_bf16x2(__uint_as_float(v[8 * u + 2]), __uint_as_float(v[8 * u + 3])),
                         pack_bf16x2(__uint_as_float(v[8 * u + 4]), __uint_as_float(v[8 * u + 5])),
                         pack_bf16x2(__uint_as_float(v[8 * u + 6]), __uint_as_float(v[8 * u + 7])));
      }
      __syncwarp();
      __nv_bfloat16* obase = p.dqkv + (static_cast<size_t>(crop) * AB_T + m * 128 + quad * 32) * ld_qkv +
                             (1 + half) * ld_o + head * AB_HD;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = 4 * j + (lane >> 3), ch = lane & 7;
        if (m * 128 + quad * 32 + row < AB_T)
          *reinterpret_cast<uint4*>(obase + static_cast<size_t>(row) * ld_qkv + ch * 8) =
              *reinterpret_cast<const uint4*>(stage + row * 144 + ch * 16);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_s, 512);
}

int attention_bwd(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, int n, int T,
                  int heads, int hd, float scale, cudaStream_t stream, float* dbias) {
  VPB_REQUIRE(n > 0 && heads > 0 && lse != nullptr, "attention_bwd: empty problem / missing log-sum-exp");
  VPB_REQUIRE(T == AB_T && hd == AB_HD, "attention_bwd: built for T=%d, head_dim=%d (got T=%d, head_dim=%d)", AB_T,
              AB_HD, T, hd);
  const int ld_o = heads * hd, ld = 3 * ld_o;
  CUtensorMap tq, tdo;
  uint64_t dims[3] = {(uint64_t)ld, (uint64_t)T, (uint64_t)n};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)T * ld * 2};
  uint32_t box[3] = {64, (uint32_t)T, 1};
  if (make_tma_desc(&tq, TMA_BF16, qkv, 3, dims, strides, box, TMA_SWIZZLE_128B)) return -1;
  uint64_t dims_o[3] = {(uint64_t)ld_o, (uint64_t)T, (uint64_t)n};
  uint64_t strides_o[2] = {(uint64_t)ld_o * 2, (uint64_t)T * ld_o * 2};
  if (make_tma_desc(&tdo, TMA_BF16, dout, 3, dims_o, strides_o, box, TMA_SWIZZLE_128B)) return -1;
  AttnBwdParams p;
  p.n = n; p.heads = heads; p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.dO = reinterpret_cast<const __nv_bfloat16*>(dout);
  p.O = reinterpret_cast<const __nv_bfloat16*>(out);
  p.lse = lse;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(dqkv);
  p.dbias = dbias;
  static bool configured = false;
  if (!configured) {
    VPB_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AB_SMEM));
    configured = true;
  }
  attention_bwd_kernel<<<n * heads, AB_THREADS, AB_SMEM, stream>>>(tq, tdo, p);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vpb
