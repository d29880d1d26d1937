// Persistent warp-specialised bf16 GEMM for sm_100a: C[M,N] = A[M,K] * B[N,K]^T with fused epilogues.
//   warp 0     : TMA producer (A and B tiles, 128B-swizzled, K-major) through a STAGES-deep mbarrier ring
//   warp 1     : tcgen05.mma issuer (one elected lane), accumulators double-buffered in TMEM
//   warps 2..5 : epilogue. tcgen05.ld -> bias / GELU / residual in registers -> 128B-swizzled staging tile in
//                shared memory -> TMA store (coalesced, clipped at the tensor edge by hardware). The fp32
//                residual tile is prefetched with TMA loads into a second staging ring. TMEM is released as soon
//                as the last chunk of a tile is in registers, so the next tile's MMAs overlap the epilogue.
// One CTA per SM, tiles 128 x BN, walked N-fastest so CTAs running together share the A panel in L2.
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
};
// epilogues that add an fp32 tile fetched by the residual-producer warp and store fp32 through the in-place ring
__host__ __device__ constexpr bool gemm_epi_adds_tile(int epi) { return epi == EPI_RESID_F32 || epi == EPI_POSTMA_F32; }

struct GemmParams {
  int M, N, K;
  const float* bias;   // [N], may be null
  void* out;
  int ldo;             // leading dimension of out / aux in elements
  const float* aux;    // residual or positional table
  int period;          // tokens per crop (EPI_POS) / pixels per crop (EPI_NCHW)
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;   // 64 bf16 = one 128-byte swizzle row
// bf16 epilogues (bias / GELU) are instruction-heavy: two epilogue warpgroups (8 warps, 2 per scheduler) split
// the column chunks of a tile; the fp32 residual epilogue is memory-heavy and keeps one warpgroup.
__host__ __device__ constexpr int gemm_epi_groups(int epi) { return (epi == 0 || epi == 1) ? 2 : 1; }
// EPI_RESID_F32 adds warp 6: a TMA producer that streams the fp32 residual tile through a ring of in-place
// staging slots, running ahead of the epilogue (and of the MMAs) by up to GEMM_RES_SLOTS chunks.
__host__ __device__ constexpr int gemm_threads(int epi) {
  return 64 + 128 * gemm_epi_groups(epi) + ((epi == 2 || epi == 5) ? 32 : 0);
}
constexpr int GEMM_RES_SLOTS = 4;
constexpr int GEMM_STAGING_BYTES = GEMM_BM * 128;   // one [128 rows x 128 B] TMA-store box

__host__ __device__ constexpr bool gemm_epi_staged(int epi) {
  return epi == EPI_BIAS_BF16 || epi == EPI_GELU_BF16 || epi == EPI_RESID_F32 || epi == EPI_POSTMA_F32;
}
__host__ __device__ constexpr int gemm_tmem_cols(int bn) {
  return 2 * bn <= 32 ? 32 : 2 * bn <= 64 ? 64 : 2 * bn <= 128 ? 128 : 2 * bn <= 256 ? 256 : 512;
}
// per-CTA bytes of one pipeline stage; with CTA pairs (cg = 2) each CTA stages its 128 A rows and HALF of the B tile
__host__ __device__ constexpr int gemm_stage_bytes(int bn, int cg = 1) { return GEMM_BM * 128 + bn * 128 / cg; }
// shared memory for the epilogue: one staging box per bf16 epilogue group, or the 4-slot residual ring
__host__ __device__ constexpr int gemm_epi_smem(int epi) {
  return !gemm_epi_staged(epi) ? 0 : (gemm_epi_adds_tile(epi) ? 4 : 2) * GEMM_STAGING_BYTES;
}
__host__ __device__ constexpr int gemm_num_stages(int bn, int epi, int cg = 1) {
  // 227 KB usable, minus 1 KB alignment slack and ~1 KB of static shared memory
  return ((230400 - gemm_epi_smem(epi)) / gemm_stage_bytes(bn, cg)) > 8
             ? 8
             : ((230400 - gemm_epi_smem(epi)) / gemm_stage_bytes(bn, cg));
}
__host__ __device__ constexpr int gemm_smem_bytes(int bn, int epi, int cg = 1) {
  return gemm_num_stages(bn, epi, cg) * gemm_stage_bytes(bn, cg) + gemm_epi_smem(epi) + 1024;
}

// Exact-erf GELU (nn.GELU default) with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, i.e. float
// rounding level): erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p z), z >= 0.
// ~17 issue slots per element instead of ~40 for erff(): the fc1 epilogue must stay under the tile's MMA time.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = fast_rcp(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = fast_ex2(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-poly, e, 1.0f);          // erf(|x|/sqrt2) in [0, 1]
  const float half_x = 0.5f * x;
  return fmaf(half_x, copysignf(erf_abs, x), half_x);  // 0.5 x (1 + erf(x/sqrt2))
}

// Same formula on two elements at once with Blackwell's packed fp32 pipe (FFMA2/FMUL2/FADD2): the FMA-pipe
// instruction count per element halves; the two MUFU ops (rcp, ex2) per element stay scalar.
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
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
}

// CG = 1: one CTA per 128 x BN tile. CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile —
// each CTA stages its own 128 A rows and half of the B tile, the leader CTA issues M=256 MMAs that read both halves,
// so per-SM shared-memory traffic per MAC drops by a third (1-CTA 128x256 tiles are smem-bandwidth bound:
// 96 B/clk of operand reads + 96 B/clk of TMA writes against 128 B/clk).
template <int BN, int EPI, int CG>
__global__ void __launch_bounds__(gemm_threads(EPI), 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_aux,
                    const GemmParams p) {
  constexpr int STAGES = gemm_num_stages(BN, EPI, CG);
  constexpr int A_BYTES = GEMM_BM * 128;
  constexpr int B_BYTES = BN * 128 / CG;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = gemm_tmem_cols(BN);
  constexpr uint32_t IDESC = umma_idesc_bf16(GEMM_BM * CG, BN);
  static_assert(CG == 1 || (CG == 2 && BN % 32 == 0), "CTA pairs split the B tile in two halves");
  constexpr bool STAGED = gemm_epi_staged(EPI);
  constexpr int CHUNK = (gemm_epi_adds_tile(EPI)) ? 32 : 64;     // columns per 128-byte staging row
  constexpr int GROUPS = gemm_epi_groups(EPI);                // epilogue warpgroups
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128 must be a multiple of 16 in [16,256]");
  static_assert(!STAGED || BN % CHUNK == 0, "staged epilogue needs BN to be a multiple of the chunk width");
  static_assert(STAGES >= 2, "pipeline needs at least two stages");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // staging boxes of [128 rows x 128 B]: one per epilogue group (bf16), or the in-place residual ring (fp32)
  uint8_t* s_out = smem + STAGES * STAGE_BYTES;
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tfull_bar[2];
  __shared__ uint64_t tempty_bar[2];
  __shared__ uint64_t res_full[GEMM_RES_SLOTS];
  __shared__ uint64_t res_empty[GEMM_RES_SLOTS];
  __shared__ uint32_t tmem_slot;
  constexpr int BIAS_PER_GROUP = (BN / CHUNK + GROUPS - 1) / GROUPS * CHUNK;   // columns a group's chunks cover
  __shared__ __align__(16) float s_bias[GROUPS][STAGED ? BIAS_PER_GROUP : 1];

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + GEMM_BM * CG - 1) / (GEMM_BM * CG);   // (pairs of) 128-row blocks
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int cta_rank = CG == 2 ? static_cast<int>(cluster_ctarank()) : 0;
  const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;     // both CTAs of a pair walk the same tiles

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4 * GROUPS * CG);   // the leader's barrier also collects the peer's epilogue warps
    }
    for (int s = 0; s < GEMM_RES_SLOTS; ++s) {
      mbar_init(&res_full[s], 1);
      mbar_init(&res_empty[s], 1);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    if (STAGED) tma_prefetch_desc(&tma_out);
    if (gemm_epi_adds_tile(EPI)) tma_prefetch_desc(&tma_aux);
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

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int m_blk = (tile / n_tiles) * CG + cta_rank;
        const int n_blk = tile % n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if constexpr (CG == 2) {
            // both CTAs' bytes are credited to the LEADER's barrier, which the MMA issuer waits on
            const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
            tma_load_2d_pair(sa, &tma_a, leader_full, kb * GEMM_BK, m_blk * GEMM_BM);
            tma_load_2d_pair(sb, &tma_b, leader_full, kb * GEMM_BK, n_blk * BN + cta_rank * (BN / 2));
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
            tma_load_2d(sa, &tma_a, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
            tma_load_2d(sb, &tma_b, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta_rank == 0) {      // with CTA pairs only the leader issues MMAs (for both CTAs)
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            if constexpr (CG == 2)
              umma_bf16_ss_pair(d_tmem, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), IDESC,
                                (kb | k) != 0 ? 1u : 0u);
            else
              umma_bf16_ss(d_tmem, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), IDESC,
                           (kb | k) != 0 ? 1u : 0u);
          }
          // frees this smem stage (in both CTAs of a pair) once the MMAs above have read it
          if constexpr (CG == 2) umma_commit_pair(&empty_bar[stage], 3);
          else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if constexpr (CG == 2) umma_commit_pair(&tfull_bar[acc], 3);
        else umma_commit(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (gemm_epi_adds_tile(EPI) && warp == 6) {
    // residual producer: chunk c of tile (m_blk, n_blk) = fp32 [128 rows x 32 cols] -> ring slot (in-place staging)
    if (lane == 0) {
      uint32_t seq = 0;
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int m_blk = (tile / n_tiles) * CG + cta_rank;
        const int n_blk = tile % n_tiles;
        for (int c = 0; c < BN / 32; ++c, ++seq) {
          const uint32_t slot = seq % GEMM_RES_SLOTS;
          mbar_wait(&res_empty[slot], ((seq / GEMM_RES_SLOTS) & 1) ^ 1);
          mbar_arrive_expect_tx(&res_full[slot], GEMM_STAGING_BYTES);
          if constexpr (EPI == EPI_POSTMA_F32) {
            // positional table rows (token = row % period): two 64-row boxes, each inside one period (period % 64 == 0)
            const int t0 = (m_blk * GEMM_BM) % p.period, t1 = (m_blk * GEMM_BM + 64) % p.period;
            tma_load_2d(s_out + slot * GEMM_STAGING_BYTES, &tma_aux, &res_full[slot], n_blk * BN + c * 32, t0);
            tma_load_2d(s_out + slot * GEMM_STAGING_BYTES + 64 * 128, &tma_aux, &res_full[slot], n_blk * BN + c * 32,
                        t1);
          } else {
            tma_load_2d(s_out + slot * GEMM_STAGING_BYTES, &tma_aux, &res_full[slot], n_blk * BN + c * 32,
                        m_blk * GEMM_BM);
          }
        }
      }
    }
  } else {
    const int quad = warp & 3;             // TMEM lane quadrant this warp may read
    const int grp = (warp - 2) >> 2;       // epilogue warpgroup
    const int etid = threadIdx.x - 64 - 128 * grp;   // 0..127 inside the warpgroup
    const int r = quad * 32 + lane;        // row inside the tile == TMEM lane
    const int bar_id = 1 + grp;            // named barrier of this warpgroup
    auto arrive_tempty = [&](int a) {      // TMEM buffer drained: tell the MMA issuer (in the pair's leader CTA)
      if constexpr (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[a]), 0));
      else mbar_arrive(&tempty_bar[a]);
    };
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t chunk_seq = 0;                // running chunk counter of this group: staging buffer = chunk_seq & 1
    for (int tile = tile0; tile < num_tiles; tile += tile_step) {
      const int m_blk = (tile / n_tiles) * CG + cta_rank;
      const int n_blk = tile % n_tiles;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;

      if constexpr (STAGED) {
        constexpr int NCHUNK = BN / CHUNK;
        // bias of this n-tile -> smem (all readers of the previous tile's bias are past its last barrier)
        // (each group keeps only the columns of its own chunks: local chunk lc <-> tile chunk lc*GROUPS + grp)
        for (int i = etid; i < BIAS_PER_GROUP; i += 128) {
          const int col = n_blk * BN + ((i / CHUNK) * GROUPS + grp) * CHUNK + (i % CHUNK);
          s_bias[grp][i] = (p.bias != nullptr && col < p.N) ? __ldg(p.bias + col) : 0.0f;
        }
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        if (grp >= NCHUNK) {                // narrow tile: this group has no chunk, just release TMEM
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_tempty(acc);
        }
#pragma unroll 1
        for (int c = grp; c < NCHUNK; c += GROUPS, ++chunk_seq) {
          // staging slot: bf16 epilogues own one box per group; the residual epilogue walks the in-place ring
          const uint32_t buf = gemm_epi_adds_tile(EPI) ? (chunk_seq % GEMM_RES_SLOTS) : static_cast<uint32_t>(grp);
          uint32_t v[CHUNK];
          if constexpr (CHUNK == 64) {
            tmem_ld_32x32b_x32(t_row + c * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
            tmem_ld_32x32b_x32(t_row + c * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          } else {
            tmem_ld_32x32b_x32(t_row + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          }
          tmem_ld_wait();
          if (c + GROUPS >= NCHUNK) {       // this group's last chunk is in registers: hand TMEM back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_tempty(acc);
          }
          uint8_t* srow = s_out + buf * GEMM_STAGING_BYTES + r * 128;
          const float* bias_c = &s_bias[grp][(c / GROUPS) * CHUNK];
          if constexpr (gemm_epi_adds_tile(EPI)) {
            // residual chunk landed in the slot (TMA, issued by warp 6 well ahead); update it in place: every thread
            // reads and writes only its own 128-byte row, so no barrier is needed before the math
            mbar_wait(&res_full[buf], (chunk_seq / GEMM_RES_SLOTS) & 1);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int pu = (u ^ (r & 7)) * 16;
              float4 x = *reinterpret_cast<const float4*>(srow + pu);
              x.x += __uint_as_float(v[4 * u + 0]) + bias_c[4 * u + 0];
              x.y += __uint_as_float(v[4 * u + 1]) + bias_c[4 * u + 1];
              x.z += __uint_as_float(v[4 * u + 2]) + bias_c[4 * u + 2];
              x.w += __uint_as_float(v[4 * u + 3]) + bias_c[4 * u + 3];
              *reinterpret_cast<float4*>(srow + pu) = x;
            }
          } else {
            // all the math first, into packed registers (overwriting v) ...
#pragma unroll
            for (int j = 0; j < CHUNK / 2; ++j) {
              const float2 b2 = *reinterpret_cast<const float2*>(bias_c + 2 * j);
              float2 f = __fadd2_rn(make_float2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])), b2);
              if constexpr (EPI == EPI_GELU_BF16) f = gelu_erf2(f);
              v[j] = pack_bf16x2(f.x, f.y);
            }
            // ... then wait until the TMA store that last read this group's staging box has finished reading it
            // (that latency is now hidden behind the math), and write the row
            if (etid == 0) tma_store_wait_read<0>();
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
#pragma unroll
            for (int u = 0; u < 8; ++u)
              *reinterpret_cast<uint4*>(srow + ((u ^ (r & 7)) * 16)) =
                  make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
          }
          fence_proxy_async_smem();         // generic-proxy smem writes -> visible to the TMA store
          asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
          if (etid == 0) {
            tma_store_2d(&tma_out, s_out + buf * GEMM_STAGING_BYTES, n_blk * BN + c * CHUNK, m_blk * GEMM_BM);
            tma_store_commit();
            if constexpr (gemm_epi_adds_tile(EPI)) {
              // hand the previous chunk's slot back to the residual producer once its store has read it
              if (chunk_seq > 0) {
                tma_store_wait_read<1>();
                mbar_arrive(&res_empty[(chunk_seq - 1) % GEMM_RES_SLOTS]);
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
            if constexpr (EPI == EPI_POS_F32) {
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
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (STAGED && etid == 0) tma_store_wait_all<0>();   // all output bytes committed before the CTA exits
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
