// Persistent warp-specialised bf16 GEMM for sm_100a: C[M,N] = A[M,K] * B[N,K]^T with fused epilogues.
//   warp 0     : TMA producer (A and B tiles, 128B-swizzled, K-major) through a STAGES-deep mbarrier ring
//   warp 1     : tcgen05.mma issuer (one elected lane), accumulators double-buffered in TMEM
//   warps 2..5 : epilogue (tcgen05.ld -> bias / GELU / residual / pos-embed / NCHW heatmap store)
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
  EPI_RELU_BF16 = 5,   // out bf16 [M, ldo]      = relu(acc * scale[n] + bias[n])  (unused by linear layers)
};

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
constexpr int GEMM_THREADS = 192;

__host__ __device__ constexpr int gemm_tmem_cols(int bn) {
  return 2 * bn <= 32 ? 32 : 2 * bn <= 64 ? 64 : 2 * bn <= 128 ? 128 : 2 * bn <= 256 ? 256 : 512;
}
__host__ __device__ constexpr int gemm_stage_bytes(int bn) { return GEMM_BM * 128 + bn * 128; }
__host__ __device__ constexpr int gemm_num_stages(int bn) {
  // keep the ring under ~192 KB so the rest of shared memory is free for the epilogue
  return (196608 / gemm_stage_bytes(bn)) > 8 ? 8 : (196608 / gemm_stage_bytes(bn));
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const GemmParams p) {
  constexpr int STAGES = gemm_num_stages(BN);
  constexpr int A_BYTES = GEMM_BM * 128;
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = gemm_tmem_cols(BN);
  constexpr uint32_t IDESC = umma_idesc_bf16(GEMM_BM, BN);
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M=128 must be a multiple of 16 in [16,256]");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tfull_bar[2];
  __shared__ uint64_t tempty_bar[2];
  __shared__ uint32_t tmem_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int k_blocks = (p.K + GEMM_BK - 1) / GEMM_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles;
        const int n_blk = tile % n_tiles;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(sa, &tma_a, &full_bar[stage], kb * GEMM_BK, m_blk * GEMM_BM);
          tma_load_2d(sb, &tma_b, &full_bar[stage], kb * GEMM_BK, n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
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
            umma_bf16_ss(d_tmem, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), IDESC,
                         (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);   // frees this smem stage once the MMAs above have read it
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);       // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    const int quad = warp & 3;             // TMEM lane quadrant this warp may read
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles;
      const int n_blk = tile % n_tiles;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * GEMM_BM + quad * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 16; ++c) {
        uint32_t r[16];
        tmem_ld_32x32b_x16(t_row + c * 16, r);
        tmem_ld_wait();
        const int col0 = n_blk * BN + c * 16;
        if (row_ok && col0 < p.N) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          const int ncols = min(16, p.N - col0);
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < ncols) v[j] += __ldg(p.bias + col0 + j);
          }
          if constexpr (EPI == EPI_BIAS_BF16 || EPI == EPI_GELU_BF16 || EPI == EPI_RELU_BF16) {
            if constexpr (EPI == EPI_GELU_BF16) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
            }
            if constexpr (EPI == EPI_RELU_BF16) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.0f);
            }
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<size_t>(row) * p.ldo + col0;
            if (ncols == 16) {
              uint4 w0 = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                    pack_bf16x2(v[6], v[7]));
              uint4 w1 = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                                    pack_bf16x2(v[14], v[15]));
              reinterpret_cast<uint4*>(o)[0] = w0;
              reinterpret_cast<uint4*>(o)[1] = w1;
            } else {
              for (int j = 0; j < ncols; ++j) o[j] = __float2bfloat16_rn(v[j]);
            }
          } else if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_POS_F32) {
            const float* a = (EPI == EPI_RESID_F32)
                                 ? p.aux + static_cast<size_t>(row) * p.ldo + col0
                                 : p.aux + static_cast<size_t>(row % p.period) * p.N + col0;
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
          } else if constexpr (EPI == EPI_NCHW_F32) {
            const int img = row / p.period;
            const int pix = row - img * p.period;
            float* o = reinterpret_cast<float*>(p.out) + (static_cast<size_t>(img) * p.N + col0) * p.period + pix;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < ncols) o[static_cast<size_t>(j) * p.period] = v[j];   // lanes = consecutive pixels: coalesced
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace vpb
