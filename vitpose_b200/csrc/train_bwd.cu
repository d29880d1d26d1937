// Backward-pass kernels of the ViTPose-B training step: everything around the tensor-core
// GEMMs, which are the forward kernel itself (gemm.cuh) applied to transposed operands:
//   dgrad  dX[M,K] = dY[M,N] . W[N,K]      = gemm(A = dY,   B = W^T [K,N])
//   wgrad  dW[N,K] = dY^T . X              = gemm(A = dY^T [N,M], B = X^T [K,M]), fp32, accumulating (EPI_RESID_F32)
// so the pieces here are layout changes (transpose, casts, gathers for the transposed convolutions), column
// reductions (bias / LayerNorm / BatchNorm parameter gradients, BatchNorm statistics) and the element-wise
// derivatives (GELU, LayerNorm, BatchNorm + ReLU). All of them are HBM-bound.
#include <cuda_bf16.h>

#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

namespace vpb {

// -------------------------------------------------------------------------------------------------
// out[C, R] = in[R, C]^T (bf16), 64 x 64 tiles through padded shared memory; R, C even.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in,
                                                             __nv_bfloat16* __restrict__ out, int R, int C,
                                                             long long in_batch_stride, long long out_batch_stride) {
  __shared__ __nv_bfloat16 tile[64][66];
  in += static_cast<size_t>(blockIdx.z) * in_batch_stride;
  out += static_cast<size_t>(blockIdx.z) * out_batch_stride;
  const int c0 = blockIdx.x * 64, r0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int i = ty; i < 64; i += 8) {
    const int r = r0 + i, c = c0 + 2 * tx;
    __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
    if (r < R && c < C) v = *reinterpret_cast<const __nv_bfloat162*>(in + static_cast<size_t>(r) * C + c);
    tile[i][2 * tx] = v.x;
    tile[i][2 * tx + 1] = v.y;
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 8) {
    const int c = c0 + i, r = r0 + 2 * tx;
    if (c < C && r < R) {
      __nv_bfloat162 v;
      v.x = tile[2 * tx][i];
      v.y = tile[2 * tx + 1][i];
      *reinterpret_cast<__nv_bfloat162*>(out + static_cast<size_t>(c) * R + r) = v;
    }
  }
}

int transpose_bf16(const void* in, void* out, int R, int C, int batch, cudaStream_t stream) {
  VPB_REQUIRE(R > 0 && C > 0 && batch > 0 && R % 2 == 0 && C % 2 == 0, "transpose: R=%d C=%d must be even", R, C);
  dim3 grid((C + 63) / 64, (R + 63) / 64, batch);
  transpose_bf16_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(in),
                                                  reinterpret_cast<__nv_bfloat16*>(out), R, C,
                                                  static_cast<long long>(R) * C, static_cast<long long>(R) * C);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// fp32 -> bf16 cast (master weights -> GEMM operands, fp32 gradient stream -> GEMM operand)
// -------------------------------------------------------------------------------------------------
__global__ void cast_f32_bf16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long long n4,
                                     const float* __restrict__ row_scale, int row_len4, int rows_per_scale) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = in[i];
  if (row_scale != nullptr) {        // gradient of a stochastic-depth branch: times mask_i / keep_prob of the crop
    const float s = __ldg(row_scale + (i / row_len4) / rows_per_scale);
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
  }
  out[i] = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}
int cast_f32_bf16(const float* in, void* out, long long n, cudaStream_t stream, const float* row_scale, int row_len,
                  int rows_per_scale) {
  VPB_REQUIRE(n > 0 && n % 4 == 0, "cast: n=%lld must be a positive multiple of 4", n);
  VPB_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0,
              "cast: unaligned buffers");
  VPB_REQUIRE(row_scale == nullptr || (row_len > 0 && row_len % 4 == 0 && rows_per_scale > 0),
              "cast: row_scale needs row_len %% 4 == 0 and rows_per_scale > 0");
  const long long n4 = n / 4;
  cast_f32_bf16_kernel<<<static_cast<unsigned>((n4 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<uint2*>(out), n4, row_scale, row_len > 0 ? row_len / 4 : 1,
      rows_per_scale > 0 ? rows_per_scale : 1);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// cast + column sums in one pass. Block = 32 x 8 threads: 128 columns (4 per thread), `rows_per_block` rows; the sums
// are taken over the bf16-ROUNDED values so that they equal colsum_accumulate() of the output.
__global__ void __launch_bounds__(256) cast_colsum_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                          int R, int C, int rows_per_block,
                                                          const float* __restrict__ row_scale, int rows_per_scale,
                                                          float* __restrict__ colsum) {
  __shared__ float red[8][128];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + 4 * tx;
  const int r_begin = blockIdx.y * rows_per_block, r_end = min(R, r_begin + rows_per_block);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (c < C) {
    for (int r = r_begin + ty; r < r_end; r += 8) {
      const size_t off = static_cast<size_t>(r) * C + c;
      float4 v = *reinterpret_cast<const float4*>(in + off);
      if (row_scale != nullptr) {
        const float s = __ldg(row_scale + r / rows_per_scale);
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
      }
      const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
      *reinterpret_cast<uint2*>(out + off) =
          make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
      const float2 flo = __bfloat1622float2(lo), fhi = __bfloat1622float2(hi);
      s0 += flo.x; s1 += flo.y; s2 += fhi.x; s3 += fhi.y;
    }
  }
  red[ty][4 * tx] = s0; red[ty][4 * tx + 1] = s1; red[ty][4 * tx + 2] = s2; red[ty][4 * tx + 3] = s3;
  __syncthreads();
  if (threadIdx.x < 128) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    const int cc = blockIdx.x * 128 + threadIdx.x;
    if (cc < C) atomicAdd(colsum + cc, s);
  }
}
int cast_f32_bf16_colsum(const float* in, void* out, int R, int C, const float* row_scale, int rows_per_scale,
                         float* colsum, cudaStream_t stream) {
  VPB_REQUIRE(R > 0 && C > 0 && C % 4 == 0 && in && out && colsum, "cast+colsum: C=%d must be a positive multiple of 4", C);
  VPB_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0,
              "cast+colsum: unaligned buffers");
  VPB_REQUIRE(row_scale == nullptr || rows_per_scale > 0, "cast+colsum: row_scale needs rows_per_scale > 0");
  const int col_blocks = (C + 127) / 128;
  int row_blocks = (4 * sm_count() + col_blocks - 1) / col_blocks;
  if (row_blocks < 1) row_blocks = 1;
  const int rows_per_block = ((R + row_blocks - 1) / row_blocks + 7) / 8 * 8;
  row_blocks = (R + rows_per_block - 1) / rows_per_block;
  cast_colsum_kernel<<<dim3(col_blocks, row_blocks), 256, 0, stream>>>(
      in, reinterpret_cast<__nv_bfloat16*>(out), R, C, rows_per_block, row_scale, rows_per_scale > 0 ? rows_per_scale : 1,
      colsum);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Column reductions over a row-major [R, C] matrix, accumulated (atomicAdd) into fp32 [C] vectors.
//   COLSUM_BF16 / COLSUM_F32 : out0 += sum_r a              (bias gradients, pos-embed gradient)
//   COLSUM_SQ_BF16           : out0 += sum a, out1 += sum a^2   (BatchNorm batch statistics)
//   COLSUM_BNBWD             : with dy' = b * (bn(a) > 0), xhat = (a - mean) * rstd:
//                              out0 += sum dy', out1 += sum dy' * xhat   (BatchNorm + ReLU backward reductions)
// Block = 32 x 8 threads: 64 columns, `rows_per_block` rows.
// -------------------------------------------------------------------------------------------------
enum ColsumMode { COLSUM_BF16 = 0, COLSUM_F32 = 1, COLSUM_SQ_BF16 = 2, COLSUM_BNBWD = 3 };

struct ColsumArgs {
  const void* a;
  const void* b;            // COLSUM_BNBWD: upstream gradient (bf16)
  const float* mean;        // COLSUM_BNBWD: per column
  const float* rstd;
  const float* gamma;
  const float* beta;
  float* out0;
  float* out1;
  int R, C, rows_per_block;
};

template <int MODE>
__global__ void __launch_bounds__(256) colsum_kernel(const ColsumArgs p) {
  __shared__ float red[2][8][64];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 64 + 2 * tx;
  const int r_begin = blockIdx.y * p.rows_per_block;
  const int r_end = min(p.R, r_begin + p.rows_per_block);
  float s0x = 0.f, s0y = 0.f, s1x = 0.f, s1y = 0.f;
  if (c < p.C) {
    float mx = 0.f, my = 0.f, rx = 0.f, ry = 0.f, gx = 0.f, gy = 0.f, bx = 0.f, by = 0.f;
    if (MODE == COLSUM_BNBWD) {
      mx = p.mean[c]; my = p.mean[c + 1];
      rx = p.rstd[c]; ry = p.rstd[c + 1];
      gx = p.gamma[c]; gy = p.gamma[c + 1];
      bx = p.beta[c]; by = p.beta[c + 1];
    }
    for (int r = r_begin + ty; r < r_end; r += 8) {
      const size_t off = static_cast<size_t>(r) * p.C + c;
      float ax, ay;
      if (MODE == COLSUM_F32) {
        const float2 v = *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(p.a) + off);
        ax = v.x; ay = v.y;
      } else {
        const float2 v = __bfloat1622float2(
            *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(p.a) + off));
        ax = v.x; ay = v.y;
      }
      if (MODE == COLSUM_BF16 || MODE == COLSUM_F32) {
        s0x += ax; s0y += ay;
      } else if (MODE == COLSUM_SQ_BF16) {
        s0x += ax; s0y += ay;
        s1x = fmaf(ax, ax, s1x); s1y = fmaf(ay, ay, s1y);
      } else {
        const float2 d = __bfloat1622float2(
            *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const __nv_bfloat16*>(p.b) + off));
        const float hx = (ax - mx) * rx, hy = (ay - my) * ry;
        const float dx = fmaf(hx, gx, bx) > 0.f ? d.x : 0.f;
        const float dy = fmaf(hy, gy, by) > 0.f ? d.y : 0.f;
        s0x += dx; s0y += dy;
        s1x = fmaf(dx, hx, s1x); s1y = fmaf(dy, hy, s1y);
      }
    }
  }
  red[0][ty][2 * tx] = s0x; red[0][ty][2 * tx + 1] = s0y;
  red[1][ty][2 * tx] = s1x; red[1][ty][2 * tx + 1] = s1y;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, col = threadIdx.x & 63;
    if (which == 1 && (MODE == COLSUM_BF16 || MODE == COLSUM_F32)) return;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[which][i][col];
    const int cc = blockIdx.x * 64 + col;
    if (cc < p.C) {
      // BatchNorm statistics feed the activations: accumulate them in fp64 so that the (unordered) atomic sum is
      // reproducible to the last fp32 bit; parameter gradients take plain fp32 atomics
      if (MODE == COLSUM_SQ_BF16 || MODE == COLSUM_BNBWD)
        atomicAdd(reinterpret_cast<double*>(which ? p.out1 : p.out0) + cc, static_cast<double>(s));
      else atomicAdd((which ? p.out1 : p.out0) + cc, s);
    }
  }
}

static int launch_colsum(int mode, const ColsumArgs& a, cudaStream_t stream) {
  VPB_REQUIRE(a.R > 0 && a.C > 0 && a.C % 2 == 0, "colsum: C=%d must be even", a.C);
  ColsumArgs p = a;
  // enough blocks to fill the machine, rows per block a multiple of 8
  const int col_blocks = (p.C + 63) / 64;
  int row_blocks = (4 * sm_count() + col_blocks - 1) / col_blocks;
  if (row_blocks < 1) row_blocks = 1;
  p.rows_per_block = ((p.R + row_blocks - 1) / row_blocks + 7) / 8 * 8;
  row_blocks = (p.R + p.rows_per_block - 1) / p.rows_per_block;
  dim3 grid(col_blocks, row_blocks);
  switch (mode) {
    case COLSUM_BF16: colsum_kernel<COLSUM_BF16><<<grid, 256, 0, stream>>>(p); break;
    case COLSUM_F32: colsum_kernel<COLSUM_F32><<<grid, 256, 0, stream>>>(p); break;
    case COLSUM_SQ_BF16: colsum_kernel<COLSUM_SQ_BF16><<<grid, 256, 0, stream>>>(p); break;
    default: colsum_kernel<COLSUM_BNBWD><<<grid, 256, 0, stream>>>(p); break;
  }
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int colsum_accumulate(const void* in, int is_f32, int R, int C, float* out, cudaStream_t stream) {
  ColsumArgs a{in, nullptr, nullptr, nullptr, nullptr, nullptr, out, nullptr, R, C, 0};
  return launch_colsum(is_f32 ? COLSUM_F32 : COLSUM_BF16, a, stream);
}
int colsum_sq_accumulate(const void* in, int R, int C, double* sum, double* sumsq, cudaStream_t stream) {
  ColsumArgs a{in, nullptr, nullptr, nullptr, nullptr, nullptr, reinterpret_cast<float*>(sum),
               reinterpret_cast<float*>(sumsq), R, C, 0};
  return launch_colsum(COLSUM_SQ_BF16, a, stream);
}
// sums of this layer into fp64 accumulators `acc` (2*C, zeroed here), then sums_f[0..C) = sum dy', sums_f[C..2C) =
// sum dy' * xhat as fp32 (what bn_relu_bwd_kernel consumes) and dbeta / dgamma += them
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ acc, float* __restrict__ sums_f,
                                       float* __restrict__ dbeta, float* __restrict__ dgamma, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float a = static_cast<float>(acc[c]), b = static_cast<float>(acc[C + c]);
  sums_f[c] = a;
  sums_f[C + c] = b;
  dbeta[c] += a;
  dgamma[c] += b;
}
int bn_relu_bwd_reduce(const void* raw, const void* dact, const float* mean, const float* rstd, const float* gamma,
                       const float* beta, int R, int C, double* acc, float* sums_f, float* dbeta, float* dgamma,
                       cudaStream_t stream) {
  VPB_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * C, stream));
  ColsumArgs a{raw, dact, mean, rstd, gamma, beta, reinterpret_cast<float*>(acc), reinterpret_cast<float*>(acc + C),
               R, C, 0};
  if (int e = launch_colsum(COLSUM_BNBWD, a, stream)) return e;
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(acc, sums_f, dbeta, dgamma, C);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// GELU (exact erf, nn.GELU default): forward on the stored pre-activation and its derivative
//   d/dx [x Phi(x)] = Phi(x) + x phi(x)
// -------------------------------------------------------------------------------------------------
// erf(x / sqrt 2) and exp(-x^2 / 2) from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, float rounding level), one rcp
// and one ex2 instead of erff()'s ~40 instructions: with erff() these two kernels were instruction-bound (47 us of
// issue time against 23 us of HBM time per layer). The exponential is also the Gaussian density of the derivative.
__device__ __forceinline__ void erf_and_gauss(float x, float& erf_s, float& gauss) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = fast_rcp(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  gauss = fast_ex2(-1.4426950408889634f * z * z);          // exp(-x^2 / 2)
  erf_s = copysignf(fmaf(-poly, gauss, 1.0f), x);
}

__global__ void gelu_fwd_kernel(const uint4* __restrict__ pre, uint4* __restrict__ out, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 v = pre[i];
  uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
    float ea, eb, ga, gb;
    erf_and_gauss(f.x, ea, ga);
    erf_and_gauss(f.y, eb, gb);
    const float ha = 0.5f * f.x, hb = 0.5f * f.y;
    w[j] = pack_bf16x2(fmaf(ha, ea, ha), fmaf(hb, eb, hb));
  }
  out[i] = make_uint4(w[0], w[1], w[2], w[3]);
}
__global__ void gelu_bwd_kernel(const uint4* __restrict__ pre, const uint4* __restrict__ dh, uint4* __restrict__ out,
                                long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 v = pre[i], g = dh[i];
  uint32_t w[4] = {v.x, v.y, v.z, v.w};
  const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 x = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
    const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[j]));
    float ea, eb, ga, gb;
    erf_and_gauss(x.x, ea, ga);
    erf_and_gauss(x.y, eb, gb);
    const float da = fmaf(0.5f, ea, 0.5f) + x.x * 0.3989422804014327f * ga;   // Phi(x) + x phi(x)
    const float db = fmaf(0.5f, eb, 0.5f) + x.y * 0.3989422804014327f * gb;
    w[j] = pack_bf16x2(d.x * da, d.y * db);
  }
  out[i] = make_uint4(w[0], w[1], w[2], w[3]);
}
int gelu_fwd_bf16(const void* pre, void* out, long long n, cudaStream_t stream) {
  VPB_REQUIRE(n > 0 && n % 8 == 0, "gelu: n=%lld must be a multiple of 8", n);
  const long long n8 = n / 8;
  gelu_fwd_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(pre), reinterpret_cast<uint4*>(out), n8);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int gelu_bwd_bf16(const void* pre, const void* dh, void* dpre, long long n, cudaStream_t stream) {
  VPB_REQUIRE(n > 0 && n % 8 == 0, "gelu: n=%lld must be a multiple of 8", n);
  const long long n8 = n / 8;
  gelu_bwd_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(pre), reinterpret_cast<const uint4*>(dh), reinterpret_cast<uint4*>(dpre), n8);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// LayerNorm backward. y = xhat * gamma + beta, xhat = (x - mean) * rstd. With g = dy * gamma:
//   dx = rstd * (g - mean_D(g) - xhat * mean_D(g * xhat)),  dgamma = sum_rows dy * xhat,  dbeta = sum_rows dy
// One warp per row (row in registers, statistics recomputed from the saved fp32 x); dx is ADDED to dx_accum
// (the gradient already flowing through the residual connection); dgamma / dbeta are accumulated per warp over
// `rows_per_warp` rows in registers and then added atomically.
// (Writing the bf16 branch gradient that the next Linear backward needs — cast_f32_bf16_colsum of dx_accum — from this
// kernel as well was tried in round 2: 70 us per launch against 33 + 9 us for the two kernels at M = 12288; the kernel
// runs one 8-warp block per SM and every extra dependent step of its row loop is exposed latency. Dropped.)
// -------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                            const __nv_bfloat16* __restrict__ dy,
                                                            float* __restrict__ dx_accum, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta, int M, float eps,
                                                            int rows_per_warp) {
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int row0 = warp_global * rows_per_warp;
  float4 gm[NV], dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    gm[i] = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // Two rows in flight: the loads of row r + 1 (x, dy and the dx_accum row it updates) are issued before row r is
  // processed, so the three shuffle reductions of a row hide the next row's memory latency (one 8-warp block per SM
  // cannot hide it by occupancy: 33 -> see profiles/r02_summary.md per launch at M = 12288).
  struct Row { float4 v[NV]; uint2 d[NV]; float4 a[NV]; };
  auto load_row = [&](Row& r, int row) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
    const uint2* dr = reinterpret_cast<const uint2*>(dy + static_cast<size_t>(row) * D);
    const float4* o = reinterpret_cast<const float4*>(dx_accum + static_cast<size_t>(row) * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      r.v[i] = xr[i * 32 + lane];
      r.d[i] = dr[i * 32 + lane];
      r.a[i] = o[i * 32 + lane];
    }
  };
  auto process_row = [&](Row& r, int row) {
    float4 d[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.d[i].x));
      const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&r.d[i].y));
      d[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
      s += (r.v[i].x + r.v[i].y) + (r.v[i].z + r.v[i].w);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    const float mean = s * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4& v = r.v[i];
      v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean;
      q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
    const float rstd = 1.0f / sqrtf(q * (1.0f / D) + eps);
    float c1 = 0.f, c2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4& v = r.v[i];
      v.x *= rstd; v.y *= rstd; v.z *= rstd; v.w *= rstd;        // xhat
      db[i].x += d[i].x; db[i].y += d[i].y; db[i].z += d[i].z; db[i].w += d[i].w;
      dg[i].x = fmaf(d[i].x, v.x, dg[i].x); dg[i].y = fmaf(d[i].y, v.y, dg[i].y);
      dg[i].z = fmaf(d[i].z, v.z, dg[i].z); dg[i].w = fmaf(d[i].w, v.w, dg[i].w);
      d[i].x *= gm[i].x; d[i].y *= gm[i].y; d[i].z *= gm[i].z; d[i].w *= gm[i].w;   // g = dy * gamma
      c1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      c2 += (d[i].x * v.x + d[i].y * v.y) + (d[i].z * v.z + d[i].w * v.w);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      c1 += __shfl_xor_sync(0xffffffffu, c1, off);
      c2 += __shfl_xor_sync(0xffffffffu, c2, off);
    }
    c1 *= (1.0f / D);
    c2 *= (1.0f / D);
    float4* o = reinterpret_cast<float4*>(dx_accum + static_cast<size_t>(row) * D);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 a = r.a[i];
      const float4& v = r.v[i];
      a.x += rstd * (d[i].x - c1 - v.x * c2);
      a.y += rstd * (d[i].y - c1 - v.y * c2);
      a.z += rstd * (d[i].z - c1 - v.z * c2);
      a.w += rstd * (d[i].w - c1 - v.w * c2);
      o[i * 32 + lane] = a;
    }
  };
  const int row_end = min(M, row0 + rows_per_warp);
  if constexpr (NV <= 6) {
    Row ra, rb;
    if (row0 < row_end) load_row(ra, row0);
    for (int row = row0; row < row_end; row += 2) {
      if (row + 1 < row_end) load_row(rb, row + 1);
      process_row(ra, row);
      if (row + 1 < row_end) {
        if (row + 2 < row_end) load_row(ra, row + 2);
        process_row(rb, row + 1);
      }
    }
  } else {          // D = 1024 / 1280: two rows do not fit the register file (372 / 744 bytes of spills): one at a time
    Row ra;
    for (int row = row0; row < row_end; ++row) {
      load_row(ra, row);
      process_row(ra, row);
    }
  }
  // block-level reduction of the per-warp dgamma / dbeta partials, then one atomic per column per block
  __shared__ float red[2][D];
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int w = 0; w < nwarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float4* rg = reinterpret_cast<float4*>(&red[0][0]) + i * 32 + lane;
        float4* rb = reinterpret_cast<float4*>(&red[1][0]) + i * 32 + lane;
        if (w == 0) {
          *rg = dg[i];
          *rb = db[i];
        } else {
          float4 a = *rg, c = *rb;
          a.x += dg[i].x; a.y += dg[i].y; a.z += dg[i].z; a.w += dg[i].w;
          c.x += db[i].x; c.y += db[i].y; c.z += db[i].z; c.w += db[i].w;
          *rg = a;
          *rb = c;
        }
      }
    }
    __syncthreads();
  }
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    atomicAdd(dgamma + c, red[0][c]);
    atomicAdd(dbeta + c, red[1][c]);
  }
}

int layernorm_bwd(const float* x, const float* gamma, const void* dy, float* dx_accum, float* dgamma, float* dbeta,
                  int M, int D, float eps, cudaStream_t stream) {
  VPB_REQUIRE(M > 0 && D > 0 && D % 128 == 0, "layernorm_bwd: D=%d must be a multiple of 128", D);
  const int warps = 8;
  // One wave of blocks: the kernel keeps two rows in flight per warp in registers (255 per thread), so ONE 8-warp block
  // is resident per SM; sizing the grid for two per SM ran 256 blocks as 148 + 108 (12 row times per warp instead of
  // 11 at 12288 rows). Few blocks also keep the parameter-gradient atomics (one per column per block) negligible.
  int rows_per_warp = (M + sm_count() * warps - 1) / (sm_count() * warps);
  if (rows_per_warp < 1) rows_per_warp = 1;
  const int total_warps = (M + rows_per_warp - 1) / rows_per_warp;
  dim3 grid((total_warps + warps - 1) / warps), block(warps * 32);
  const __nv_bfloat16* d = reinterpret_cast<const __nv_bfloat16*>(dy);
  switch (D / 128) {
#define VPB_LNB_CASE(NV_)                                                                                          \
  case NV_:                                                                                                        \
    layernorm_bwd_kernel<NV_><<<grid, block, 0, stream>>>(x, gamma, d, dx_accum, dgamma, dbeta, M, eps, rows_per_warp); \
    break;
    VPB_LNB_CASE(1) VPB_LNB_CASE(2) VPB_LNB_CASE(3) VPB_LNB_CASE(4) VPB_LNB_CASE(6) VPB_LNB_CASE(8) VPB_LNB_CASE(10)
#undef VPB_LNB_CASE
    default:
      set_last_error("layernorm_bwd: unsupported D=%d", D);
      return -2;
  }
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// BatchNorm2d in training mode + ReLU on NHWC bf16 rows [R = n*h*w, C]:
//   forward  act = relu((raw - mean) * rstd * gamma + beta)              (mean / rstd from colsum_sq_accumulate)
//   backward draw = gamma * rstd * (dy' - dbeta/R - xhat * dgamma/R),  dy' = dact * (act > 0)
// -------------------------------------------------------------------------------------------------
// eight consecutive per-channel fp32 parameters (c0 is a multiple of 8: 32-byte aligned) with two 16-byte loads
__device__ __forceinline__ void load8(const float* __restrict__ p, float (&o)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__global__ void bn_relu_fwd_kernel(const uint4* __restrict__ raw, uint4* __restrict__ act, const float* __restrict__ mean,
                                   const float* __restrict__ rstd, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, long long n8, int C) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const int c0 = static_cast<int>((i * 8) % C);
  const uint4 v = raw[i];
  uint32_t w[4] = {v.x, v.y, v.z, v.w};
  float pm[8], pr[8], pg[8], pb[8];             // the 8 channels' parameters: two 16-byte loads per array
  load8(mean + c0, pm);
  load8(rstd + c0, pr);
  load8(gamma + c0, pg);
  load8(beta + c0, pb);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
    const int c = 2 * j;
    const float a = fmaxf(fmaf((f.x - pm[c]) * pr[c], pg[c], pb[c]), 0.f);
    const float b = fmaxf(fmaf((f.y - pm[c + 1]) * pr[c + 1], pg[c + 1], pb[c + 1]), 0.f);
    w[j] = pack_bf16x2(a, b);
  }
  act[i] = make_uint4(w[0], w[1], w[2], w[3]);
}
__global__ void bn_relu_bwd_kernel(const uint4* __restrict__ raw, const uint4* __restrict__ dact, uint4* __restrict__ draw,
                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ dbeta, const float* __restrict__ dgamma, float inv_rows,
                                   long long n8, int C) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const int c0 = static_cast<int>((i * 8) % C);
  const uint4 v = raw[i], g = dact[i];
  uint32_t w[4] = {v.x, v.y, v.z, v.w};
  const uint32_t gw[4] = {g.x, g.y, g.z, g.w};
  float pm[8], pr[8], pg[8], pb[8], pdb[8], pdg[8];
  load8(mean + c0, pm);
  load8(rstd + c0, pr);
  load8(gamma + c0, pg);
  load8(beta + c0, pb);
  load8(dbeta + c0, pdb);
  load8(dgamma + c0, pdg);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
    const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[j]));
    float o[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int c = 2 * j + k;
      const float xh = ((k ? f.y : f.x) - pm[c]) * pr[c];
      const float dyp = fmaf(xh, pg[c], pb[c]) > 0.f ? (k ? d.y : d.x) : 0.f;
      o[k] = pg[c] * pr[c] * (dyp - pdb[c] * inv_rows - xh * pdg[c] * inv_rows);
    }
    w[j] = pack_bf16x2(o[0], o[1]);
  }
  draw[i] = make_uint4(w[0], w[1], w[2], w[3]);
}
int bn_relu_fwd(const void* raw, void* act, const float* mean, const float* rstd, const float* gamma, const float* beta,
                long long rows, int C, cudaStream_t stream) {
  VPB_REQUIRE(rows > 0 && C > 0 && C % 8 == 0, "bn_relu: C=%d must be a multiple of 8", C);
  VPB_REQUIRE(((reinterpret_cast<uintptr_t>(mean) | reinterpret_cast<uintptr_t>(rstd) | reinterpret_cast<uintptr_t>(gamma) |
                reinterpret_cast<uintptr_t>(beta)) & 15) == 0, "bn_relu: per-channel arrays must be 16-byte aligned");
  const long long n8 = rows * C / 8;
  bn_relu_fwd_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(raw), reinterpret_cast<uint4*>(act), mean, rstd, gamma, beta, n8, C);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int bn_relu_bwd(const void* raw, const void* dact, void* draw, const float* mean, const float* rstd, const float* gamma,
                const float* beta, const float* dbeta, const float* dgamma, long long rows, int C, cudaStream_t stream) {
  VPB_REQUIRE(rows > 0 && C > 0 && C % 8 == 0, "bn_relu: C=%d must be a multiple of 8", C);
  VPB_REQUIRE(((reinterpret_cast<uintptr_t>(mean) | reinterpret_cast<uintptr_t>(rstd) | reinterpret_cast<uintptr_t>(gamma) |
                reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(dbeta) |
                reinterpret_cast<uintptr_t>(dgamma)) & 15) == 0, "bn_relu: per-channel arrays must be 16-byte aligned");
  const long long n8 = rows * C / 8;
  bn_relu_bwd_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(raw), reinterpret_cast<const uint4*>(dact), reinterpret_cast<uint4*>(draw), mean,
      rstd, gamma, beta, dbeta, dgamma, 1.0f / static_cast<float>(rows), n8, C);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
// mean / rstd (and the running statistics update of nn.BatchNorm2d: momentum, unbiased variance) from the sums
__global__ void bn_finalize_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, int C, float rows, float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float m = static_cast<float>(sum[c] / rows);
  const float var = fmaxf(static_cast<float>(sumsq[c] / rows - (sum[c] / rows) * (sum[c] / rows)), 0.f);
  mean[c] = m;
  rstd[c] = 1.0f / sqrtf(var + eps);
  if (running_mean != nullptr) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * rows / fmaxf(rows - 1.f, 1.f);
  }
}
int bn_finalize(const double* sum, const double* sumsq, float* mean, float* rstd, float* running_mean, float* running_var,
                int C, long long rows, float eps, float momentum, cudaStream_t stream) {
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(sum, sumsq, mean, rstd, running_mean, running_var, C,
                                                         static_cast<float>(rows), eps, momentum);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Heatmap gradient fp32 NCHW [n, K, P] -> bf16 pixel-major [n*P, Kp] (Kp >= K, zero padded): the A operand of the
// final 1x1 conv's dgrad / (after a transpose) wgrad GEMMs.
// -------------------------------------------------------------------------------------------------
__global__ void nchw_to_rows_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int K, int P, int Kp) {
  __shared__ float tile[32][33];
  const int im = blockIdx.z;
  const int p0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int k = k0 + i, pp = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (k < K && pp < P) ? in[(static_cast<size_t>(im) * K + k) * P + pp] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int pp = p0 + i, k = k0 + threadIdx.x;
    if (pp < P && k < Kp) out[(static_cast<size_t>(im) * P + pp) * Kp + k] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}
int nchw_f32_to_rows_bf16(const float* in, void* out, int n, int K, int P, int Kp, cudaStream_t stream) {
  VPB_REQUIRE(n > 0 && K > 0 && P > 0 && Kp >= K, "nchw_to_rows: bad shape");
  dim3 grid((P + 31) / 32, (Kp + 31) / 32, n), block(32, 8);
  nchw_to_rows_kernel<<<grid, block, 0, stream>>>(in, reinterpret_cast<__nv_bfloat16*>(out), K, P, Kp);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Gathers for the backward of ConvTranspose2d(k4, s2, p1) in its 4-phase form (engine.pack_deconv_weight):
//   forward  Y[n, 2i+py, 2j+px, :] = sum_t X[n, i+dy(py,ty), j+dx(px,tx), :] . Wp[ph][:, t*Cin : (t+1)*Cin]^T
//   with d(p, 0) = 0, d(0, 1) = -1, d(1, 1) = +1.
// deconv_gather_x   : Xcol[ph][pix, t*Cin + ci] = X[n, i+dy, j+dx, ci]   (zero outside)   -> wgrad B operand (transposed later)
// deconv_gather_dy  : G[pix, (ph*4+t)*Cout + co] = dY[n, 2(i-dy)+py, 2(j-dx)+px, co] (zero when (i-dy, j-dx) is outside)
//                     -> dgrad A operand, against W2g[ci, (ph*4+t)*Cout + co] = Wp[ph][co][t*Cin+ci]
// deconv_phase_dy   : Yp[ph][pix, co] = dY[n, 2i+py, 2j+px, co]                          -> wgrad A operand (transposed later)
// One thread = 8 channels (16 bytes).
// -------------------------------------------------------------------------------------------------
__device__ __forceinline__ int deconv_shift(int parity, int tap) { return tap == 0 ? 0 : (parity == 0 ? -1 : 1); }

// IdxT = unsigned when every index fits 32 bits (always, at the sizes of the training step): 64-bit divisions by
// run-time values cost ~100 instructions each and made these copy kernels instruction-bound.
// One thread = one (input pixel, 8-channel group): the pixel's coordinates are decoded once, its 3 x 3 neighbourhood is
// read once (9 loads) and written to the 16 (phase, tap) slots that use it (one thread per 16-byte OUTPUT element with
// six run-time divisions each ran at 2.8 TB/s).
template <typename IdxT>
__global__ void deconv_gather_x_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int h, int w, int cin8,
                                       long long pixels_, long long total) {
  const IdxT idx = static_cast<IdxT>(blockIdx.x) * blockDim.x + threadIdx.x;      // over pixels * cin8
  if (static_cast<long long>(idx) >= total) return;
  const IdxT pixels = static_cast<IdxT>(pixels_);
  const int c8 = static_cast<int>(idx % cin8);
  const IdxT pix = idx / cin8;
  const int j = static_cast<int>(pix % w), i = static_cast<int>((pix / w) % h);
  const IdxT im = pix / (static_cast<IdxT>(w) * h);
  uint4 nb[3][3];                               // nb[dy + 1][dx + 1]
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int ii = i + dy, jj = j + dx;
      nb[dy + 1][dx + 1] = (ii >= 0 && ii < h && jj >= 0 && jj < w) ? x[((im * h + ii) * w + jj) * cin8 + c8]
                                                                  : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
  for (int ph = 0; ph < 4; ++ph)
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int dy = deconv_shift(ph >> 1, t >> 1), dx = deconv_shift(ph & 1, t & 1);
      out[((static_cast<IdxT>(ph) * pixels + pix) * 4 + t) * cin8 + c8] = nb[dy + 1][dx + 1];
    }
}
// One thread = one (input pixel, 8-channel group) and its 16 (phase, tap) sources: one row of 16 * Cout per pixel.
template <typename IdxT>
__global__ void deconv_gather_dy_kernel(const uint4* __restrict__ dy, uint4* __restrict__ out, int h, int w, int cout8,
                                        long long total) {
  const IdxT idx = static_cast<IdxT>(blockIdx.x) * blockDim.x + threadIdx.x;      // over pixels * cout8
  if (static_cast<long long>(idx) >= total) return;
  const int c8 = static_cast<int>(idx % cout8);
  const IdxT pix = idx / cout8;
  const int j = static_cast<int>(pix % w), i = static_cast<int>((pix / w) % h);
  const IdxT im = pix / (static_cast<IdxT>(w) * h);
#pragma unroll
  for (int pt = 0; pt < 16; ++pt) {
    const int ph = pt >> 2, t = pt & 3;
    const int ii = i - deconv_shift(ph >> 1, t >> 1), jj = j - deconv_shift(ph & 1, t & 1);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (ii >= 0 && ii < h && jj >= 0 && jj < w)
      v = dy[((im * 2 * h + 2 * ii + (ph >> 1)) * 2 * w + 2 * jj + (ph & 1)) * cout8 + c8];
    out[(pix * 16 + pt) * cout8 + c8] = v;
  }
}
template <typename IdxT>
__global__ void deconv_phase_dy_kernel(const uint4* __restrict__ dy, uint4* __restrict__ out, int h, int w, int cout8,
                                       long long pixels_, long long total) {
  const IdxT idx = static_cast<IdxT>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (static_cast<long long>(idx) >= total) return;
  const IdxT pixels = static_cast<IdxT>(pixels_);
  const int c8 = static_cast<int>(idx % cout8);
  IdxT rest = idx / cout8;
  const IdxT pix = rest % pixels;
  const int ph = static_cast<int>(rest / pixels);
  const int j = static_cast<int>(pix % w), i = static_cast<int>((pix / w) % h);
  const IdxT im = pix / (static_cast<IdxT>(w) * h);
  out[idx] = dy[((im * 2 * h + 2 * i + (ph >> 1)) * 2 * w + 2 * j + (ph & 1)) * cout8 + c8];
}
int deconv_gather_x(const void* x, void* out, int n, int h, int w, int cin, cudaStream_t stream) {
  VPB_REQUIRE(n > 0 && h > 0 && w > 0 && cin % 8 == 0, "deconv_gather_x: bad shape");
  const long long pixels = static_cast<long long>(n) * h * w, total = pixels * (cin / 8);      // threads
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (16 * total + 256 < (1ll << 31))      // the largest output index is < 16 * total
    deconv_gather_x_kernel<unsigned><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x),
                                                               reinterpret_cast<uint4*>(out), h, w, cin / 8, pixels, total);
  else
    deconv_gather_x_kernel<long long><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(x),
                                                                reinterpret_cast<uint4*>(out), h, w, cin / 8, pixels, total);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int deconv_gather_dy(const void* dy, void* out, int n, int h, int w, int cout, cudaStream_t stream) {
  VPB_REQUIRE(n > 0 && h > 0 && w > 0 && cout % 8 == 0, "deconv_gather_dy: bad shape");
  const long long total = static_cast<long long>(n) * h * w * (cout / 8);                        // threads
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (16 * total + 256 < (1ll << 31))      // the largest output index is < 16 * total, source indices < 4 * total
    deconv_gather_dy_kernel<unsigned><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(dy),
                                                                reinterpret_cast<uint4*>(out), h, w, cout / 8, total);
  else
    deconv_gather_dy_kernel<long long><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(dy),
                                                                 reinterpret_cast<uint4*>(out), h, w, cout / 8, total);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int deconv_phase_dy(const void* dy, void* out, int n, int h, int w, int cout, cudaStream_t stream) {
  VPB_REQUIRE(n > 0 && h > 0 && w > 0 && cout % 8 == 0, "deconv_phase_dy: bad shape");
  const long long pixels = static_cast<long long>(n) * h * w, total = 4 * pixels * (cout / 8);
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (total + 256 < (1ll << 31))           // source index = a permutation of [0, total)
    deconv_phase_dy_kernel<unsigned><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(dy),
                                                               reinterpret_cast<uint4*>(out), h, w, cout / 8, pixels, total);
  else
    deconv_phase_dy_kernel<long long><<<grid, 256, 0, stream>>>(reinterpret_cast<const uint4*>(dy),
                                                                reinterpret_cast<uint4*>(out), h, w, cout / 8, pixels, total);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------------------
// ConvTranspose2d weight [Cin, Cout, 4, 4] (fp32 master) <-> the packed layouts of the implicit GEMMs, one launch
// each (the training step repacks the weights it has just updated every iteration; as 4 x 16 strided torch copies
// this was the largest group of ATen kernels in the step):
//   wp  bf16 [4 phases][Cout][4 taps * Cin]  forward B operand (include/vitpose_b200.h, vpb_weights.deconv_w)
//   wd  bf16 [Cin][16 * Cout], column (phase * 4 + tap) * Cout + co: B operand of the input-gradient GEMM
//   and back: packed weight gradient fp32 [4][Cout][4 * Cin] -> [Cin, Cout, 4, 4].
// (phase, tap) <-> (kh, kw): py = 0: taps (ty 0, 1) read kh (1, 3); py = 1: kh (2, 0); same for x.
// -------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void deconv_k_to_phase_tap(int k, int& par, int& tap) {
  // kh 0 -> (py 1, ty 1), 1 -> (0, 0), 2 -> (1, 0), 3 -> (0, 1)
  par = (k == 0 || k == 2) ? 1 : 0;
  tap = (k == 0 || k == 3) ? 1 : 0;
}
__global__ void deconv_pack_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp,
                                          __nv_bfloat16* __restrict__ wd, int cin, int cout) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // over [cin, cout, 4, 4]
  if (idx >= cin * cout * 16) return;
  const int kw = idx & 3, kh = (idx >> 2) & 3, co = (idx >> 4) % cout, ci = (idx >> 4) / cout;
  int py, ty, px, tx;
  deconv_k_to_phase_tap(kh, py, ty);
  deconv_k_to_phase_tap(kw, px, tx);
  const int ph = py * 2 + px, t = ty * 2 + tx;
  const __nv_bfloat16 v = __float2bfloat16_rn(__ldg(w + idx));
  wp[(static_cast<size_t>(ph) * cout + co) * (4 * cin) + t * cin + ci] = v;
  if (wd != nullptr) wd[static_cast<size_t>(ci) * (16 * cout) + (ph * 4 + t) * cout + co] = v;
}
__global__ void deconv_unpack_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ dw, int cin, int cout) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // over [cin, cout, 4, 4]
  if (idx >= cin * cout * 16) return;
  const int kw = idx & 3, kh = (idx >> 2) & 3, co = (idx >> 4) % cout, ci = (idx >> 4) / cout;
  int py, ty, px, tx;
  deconv_k_to_phase_tap(kh, py, ty);
  deconv_k_to_phase_tap(kw, px, tx);
  dw[idx] = __ldg(dwp + (static_cast<size_t>(py * 2 + px) * cout + co) * (4 * cin) + (ty * 2 + tx) * cin + ci);
}
int deconv_pack_weight(const float* w, void* wp, void* wd, int cin, int cout, cudaStream_t stream) {
  VPB_REQUIRE(cin > 0 && cout > 0 && static_cast<long long>(cin) * cout * 16 < (1ll << 31), "deconv_pack_weight: bad shape");
  const int total = cin * cout * 16;
  deconv_pack_weight_kernel<<<(total + 255) / 256, 256, 0, stream>>>(w, reinterpret_cast<__nv_bfloat16*>(wp),
                                                                     reinterpret_cast<__nv_bfloat16*>(wd), cin, cout);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int deconv_unpack_wgrad(const float* dwp, float* dw, int cin, int cout, cudaStream_t stream) {
  VPB_REQUIRE(cin > 0 && cout > 0 && static_cast<long long>(cin) * cout * 16 < (1ll << 31), "deconv_unpack_wgrad: bad shape");
  const int total = cin * cout * 16;
  deconv_unpack_wgrad_kernel<<<(total + 255) / 256, 256, 0, stream>>>(dwp, dw, cin, cout);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vpb
