// Training-step operators of the ViTPose-B training config: the pieces that do not need the
// network backward — heatmap loss (+ its gradient), global gradient norm, and the AdamW update with per-group
// lr / weight decay (layer decay). All are single-pass, vectorised, HBM-bound kernels.
//   JointsMSELoss.forward      mmpose/models/losses/mse_loss.py:24-45
//   clip_grad_norm_(max_norm)  mmcv OptimizerHook (grad_clip=dict(max_norm=1.), ViTPose_base_coco_256x192.py:30)
//   AdamW step                 torch.optim.AdamW as configured at ViTPose_base_coco_256x192.py:16-28
#include <cuda_bf16.h>

#include "../../include/vitpose_b200.h"
#include "host_util.h"
#include "ops.h"

namespace vpb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// loss = loss_weight / (K * N * HW) * sum_{n,k,hw} ((o - t) * w[n,k])^2 ; grad_o = 2 * c * w^2 * (o - t)
__global__ void joints_mse_kernel(const float* __restrict__ out, const float* __restrict__ tgt,
                                  const float* __restrict__ w, float* __restrict__ loss, float* __restrict__ grad,
                                  int HW, long long total, float coef) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const float wk = w ? w[i / HW] : 1.0f;
    const float d = (out[i] - tgt[i]) * wk;
    acc = fmaf(d, d, acc);
    if (grad) grad[i] = 2.0f * coef * wk * d;
  }
  acc = warp_sum(acc);
  __shared__ float s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(loss, v * coef);
  }
}

int joints_mse_loss(const float* output, const float* target, const float* target_weight, int N, int K, int HW,
                    float loss_weight, float* loss, float* grad_output, cudaStream_t stream) {
  VPB_REQUIRE(N > 0 && K > 0 && HW > 0, "joints_mse: bad shape");
  const long long total = static_cast<long long>(N) * K * HW;
  VPB_CHECK_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), stream));
  const float coef = loss_weight / (static_cast<float>(K) * static_cast<float>(N) * static_cast<float>(HW));
  int blocks = static_cast<int>((total + 255) / 256);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  joints_mse_kernel<<<blocks, 256, 0, stream>>>(output, target, target_weight, loss, grad_output, HW, total, coef);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

__global__ void sq_norm_kernel(const float* __restrict__ g, long long n, float* __restrict__ out) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    acc = fmaf(g[i], g[i], acc);
  acc = warp_sum(acc);
  __shared__ float s[32];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? s[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

// accumulates sum(g^2) into *sq_norm_accum (caller zeroes it once per step, calls per tensor, takes sqrt)
int grad_sq_norm_accumulate(const float* grad, long long n, float* sq_norm_accum, cudaStream_t stream) {
  if (n <= 0) return 0;
  int blocks = static_cast<int>((n + 255) / 256);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  sq_norm_kernel<<<blocks, 256, 0, stream>>>(grad, n, sq_norm_accum);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// torch.optim.AdamW (decoupled weight decay, bias correction), gradient pre-scaled by clip coefficient
// min(1, max_norm / (sqrt(*sq_norm) + 1e-6)) when sq_norm != null.
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                             float wd, float bc1, float bc2_sqrt, const float* __restrict__ sq_norm, float max_norm) {
  float gs = 1.0f;
  if (sq_norm != nullptr) {
    const float coef = max_norm / (sqrtf(*sq_norm) + 1e-6f);
    gs = coef < 1.0f ? coef : 1.0f;
  }
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i] * gs;
    float pi = p[i] * (1.0f - lr * wd);
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

int adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr, float beta1,
               float beta2, float eps, float weight_decay, int step, const float* sq_norm, float max_norm,
               cudaStream_t stream) {
  if (n <= 0) return 0;
  VPB_REQUIRE(step >= 1, "adamw: step must be >= 1");
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, static_cast<float>(step)));
  int blocks = static_cast<int>((n + 255) / 256);
  const int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  adamw_kernel<<<blocks, 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                           bc1, bc2_sqrt, sq_norm, max_norm);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- multi-tensor forms: one launch for every parameter of the model -----------------------------------------
// `entries` (device) describes the tensors; `chunk_start` (device, n+1 ints) is the prefix sum of their chunk counts
// (a chunk = MT_CHUNK elements = one CTA). A CTA finds its tensor by binary search over chunk_start.
constexpr int MT_CHUNK = 256 * 16;

__device__ __forceinline__ int mt_find(const int* __restrict__ chunk_start, int n, int chunk) {
  int lo = 0, hi = n;                       // chunk_start[lo] <= chunk < chunk_start[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (chunk_start[mid] <= chunk) lo = mid; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) sq_norm_multi_kernel(const vpb_tensor_entry* __restrict__ entries,
                                                            const int* __restrict__ chunk_start, int n,
                                                            float* __restrict__ out) {
  const int t = mt_find(chunk_start, n, blockIdx.x);
  const vpb_tensor_entry e = entries[t];
  const long long base = static_cast<long long>(blockIdx.x - chunk_start[t]) * MT_CHUNK;
  float acc = 0.f;
  if ((reinterpret_cast<uintptr_t>(e.grad) & 15) == 0) {      // 16-byte loads; a fixed order either way
#pragma unroll
    for (int j = 0; j < MT_CHUNK / 1024; ++j) {
      const long long i = base + j * 1024 + threadIdx.x * 4;
      if (i + 3 < e.n) {
        const float4 g = *reinterpret_cast<const float4*>(e.grad + i);
        acc = fmaf(g.x, g.x, acc);
        acc = fmaf(g.y, g.y, acc);
        acc = fmaf(g.z, g.z, acc);
        acc = fmaf(g.w, g.w, acc);
      } else {
        for (long long q = i; q < e.n && q < i + 4; ++q) acc = fmaf(e.grad[q], e.grad[q], acc);
      }
    }
  } else {
    for (int k = threadIdx.x; k < MT_CHUNK; k += 256) {
      const long long i = base + k;
      if (i < e.n) acc = fmaf(e.grad[i], e.grad[i], acc);
    }
  }
  acc = warp_sum(acc);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? s[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[blockIdx.x] = v;          // per-chunk partial: no atomics, fixed order below
  }
}
// sum of the per-chunk partials in a fixed order (one CTA): every replica gets the same clip coefficient bit for bit
__global__ void __launch_bounds__(1024) sq_norm_finish_kernel(const float* __restrict__ partial, int n,
                                                              float* __restrict__ out) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 1024) acc += static_cast<double>(partial[i]);
  __shared__ double s[1024];
  s[threadIdx.x] = acc;
  __syncthreads();
  for (int w = 512; w > 0; w >>= 1) {
    if (threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = static_cast<float>(s[0]);
}

__global__ void __launch_bounds__(256) adamw_multi_kernel(const vpb_tensor_entry* __restrict__ entries,
                                                          const int* __restrict__ chunk_start, int n, float beta1,
                                                          float beta2, float eps, const float* __restrict__ sq_norm,
                                                          float max_norm, int allow_vec) {
  const int t = mt_find(chunk_start, n, blockIdx.x);
  const vpb_tensor_entry e = entries[t];
  float gs = 1.0f;
  if (sq_norm != nullptr) {
    const float coef = max_norm / (sqrtf(*sq_norm) + 1e-6f);
    gs = coef < 1.0f ? coef : 1.0f;
  }
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(e.step));
  const float bc2_sqrt = sqrtf(1.0f - powf(beta2, static_cast<float>(e.step)));
  const long long base = static_cast<long long>(blockIdx.x - chunk_start[t]) * MT_CHUNK;
  const float decay = 1.0f - e.lr * e.weight_decay, step_size = e.lr / bc1;
  // one element, exactly the arithmetic of torch's single-tensor AdamW
  auto update = [&](float g, float& pr, float& m, float& v) {
    const float gi = g * gs;
    const float pi = pr * decay;
    m = beta1 * m + (1.0f - beta1) * gi;
    v = beta2 * v + (1.0f - beta2) * gi * gi;
    const float denom = sqrtf(v) / bc2_sqrt + eps;
    pr = pi - step_size * (m / denom);
  };
  const bool vec = allow_vec && ((reinterpret_cast<uintptr_t>(e.grad) | reinterpret_cast<uintptr_t>(e.param) |
                     reinterpret_cast<uintptr_t>(e.exp_avg) | reinterpret_cast<uintptr_t>(e.exp_avg_sq)) & 15) == 0;
  if (vec) {      // 16-byte accesses, four independent loads per array in flight per thread
#pragma unroll
    for (int j = 0; j < MT_CHUNK / 1024; ++j) {
      const long long i = base + j * 1024 + threadIdx.x * 4;
      if (i + 3 < e.n) {
        const float4 g = *reinterpret_cast<const float4*>(e.grad + i);
        float4 pr = *reinterpret_cast<const float4*>(e.param + i);
        float4 m = *reinterpret_cast<const float4*>(e.exp_avg + i);
        float4 v = *reinterpret_cast<const float4*>(e.exp_avg_sq + i);
        update(g.x, pr.x, m.x, v.x);
        update(g.y, pr.y, m.y, v.y);
        update(g.z, pr.z, m.z, v.z);
        update(g.w, pr.w, m.w, v.w);
        *reinterpret_cast<float4*>(e.exp_avg + i) = m;
        *reinterpret_cast<float4*>(e.exp_avg_sq + i) = v;
        *reinterpret_cast<float4*>(e.param + i) = pr;
      } else {
        for (long long q = i; q < e.n && q < i + 4; ++q) update(e.grad[q], e.param[q], e.exp_avg[q], e.exp_avg_sq[q]);
      }
    }
  } else {
    for (int k = threadIdx.x; k < MT_CHUNK; k += 256) {
      const long long i = base + k;
      if (i >= e.n) break;
      update(e.grad[i], e.param[i], e.exp_avg[i], e.exp_avg_sq[i]);
    }
  }
}

int adamw_multi(const vpb_tensor_entry* entries, const int* chunk_start, int n, int total_chunks, float beta1,
                float beta2, float eps, float* sq_norm, float max_norm, cudaStream_t stream) {
  VPB_REQUIRE(n > 0 && total_chunks > 0 && entries && chunk_start, "adamw_multi: empty table");
  if (sq_norm != nullptr) {      // sq_norm[0] = result, sq_norm[1 .. total_chunks] = per-chunk partials
    sq_norm_multi_kernel<<<total_chunks, 256, 0, stream>>>(entries, chunk_start, n, sq_norm + 1);
    VPB_CHECK_CUDA(cudaGetLastError());
    sq_norm_finish_kernel<<<1, 1024, 0, stream>>>(sq_norm + 1, total_chunks, sq_norm);
    VPB_CHECK_CUDA(cudaGetLastError());
  }
  static const int allow_vec = getenv("VPB_ADAMW_SCALAR") == nullptr;      // A/B switch for measurements
  adamw_multi_kernel<<<total_chunks, 256, 0, stream>>>(entries, chunk_start, n, beta1, beta2, eps, sq_norm, max_norm,
                                                       allow_vec);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- bf16 operand copies of every linear layer in ONE launch -------------------------------------------------------
// The training step needs W (bf16, [out, in]) for the forward GEMM and W^T (bf16, [in, out]) for the input-gradient GEMM
// of each of its ~50 linear layers, refreshed from the fp32 master parameters after every optimizer step. One cast and
// one transpose launch per layer were launch-latency-bound (125 launches, 0.82 ms per step for 0.5 GB of traffic).
// One CTA = one 64 x 64 tile of one matrix: coalesced fp32 read (8 bytes per lane), coalesced bf16 write of W, shared-
// memory transpose, coalesced bf16 write of W^T (4 bytes per lane: 128-byte segments; with 32 x 32 tiles and 2-byte
// accesses the kernel ran at 2.8 TB/s). Rounding = round to nearest even, identical to cast_f32_bf16 + transpose_bf16.
// Matrices with an odd number of rows or columns take the element-wise path.
__global__ void __launch_bounds__(256) cast_transpose_multi_kernel(const vpb_cast_entry* __restrict__ entries,
                                                                   const int* __restrict__ tile_start, int n) {
  __shared__ __align__(4) __nv_bfloat16 tile[64][66];
  const int t = mt_find(tile_start, n, blockIdx.x);
  const vpb_cast_entry e = entries[t];
  const int local = blockIdx.x - tile_start[t];
  const int tiles_x = (e.cols + 63) / 64;
  const int ty = local / tiles_x, tx = local - ty * tiles_x;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;          // 32 x 8 threads, two columns per thread
  __nv_bfloat16* w = reinterpret_cast<__nv_bfloat16*>(e.w);
  __nv_bfloat16* wt = reinterpret_cast<__nv_bfloat16*>(e.wt);
  const bool vec = ((e.cols | e.rows) & 1) == 0;
  const int x = tx * 64 + 2 * lx;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int yl = ly + 8 * j, y = ty * 64 + yl;
    if (y >= e.rows) continue;
    const size_t off = static_cast<size_t>(y) * e.cols + x;
    if (vec) {
      if (x < e.cols) {
        const float2 f = *reinterpret_cast<const float2*>(e.src + off);
        const __nv_bfloat162 v = __floats2bfloat162_rn(f.x, f.y);
        *reinterpret_cast<__nv_bfloat162*>(w + off) = v;
        *reinterpret_cast<__nv_bfloat162*>(&tile[yl][2 * lx]) = v;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (x + q < e.cols) {
          const __nv_bfloat16 v = __float2bfloat16_rn(e.src[off + q]);
          w[off + q] = v;
          tile[yl][2 * lx + q] = v;
        }
    }
  }
  __syncthreads();
  const int xt = ty * 64 + 2 * lx;                                  // rows of W = columns of W^T (two per thread)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = ly + 8 * j, yt = tx * 64 + c;                     // column of W = row of W^T
    if (yt >= e.cols) continue;
    const size_t off = static_cast<size_t>(yt) * e.rows + xt;
    if (vec) {
      if (xt < e.rows) {
        __nv_bfloat162 v;
        v.x = tile[2 * lx][c];
        v.y = tile[2 * lx + 1][c];
        *reinterpret_cast<__nv_bfloat162*>(wt + off) = v;
      }
    } else {
#pragma unroll
      for (int q = 0; q < 2; ++q)
        if (xt + q < e.rows) wt[off + q] = tile[2 * lx + q][c];
    }
  }
}

int cast_transpose_multi(const vpb_cast_entry* entries, const int* tile_start, int n, int total_tiles,
                         cudaStream_t stream) {
  VPB_REQUIRE(n > 0 && total_tiles > 0 && entries && tile_start, "cast_transpose_multi: empty table");
  cast_transpose_multi_kernel<<<total_tiles, 256, 0, stream>>>(entries, tile_start, n);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- training-time accuracy: pose_pck_accuracy (mmpose/core/evaluation/top_down_eval.py:133-215) -------------
// pred / gt: arg-max coordinates [N,K,2] of output and target heatmaps (-1 where the map's maximum is <= 0);
// distances are normalised by (norm0, norm1) = (H, W) applied to (x, y) exactly as the reference does, computed in
// fp64 and stored as fp32 before the `< thr` test (NumPy stores them into a float32 array).
__global__ void pck_accuracy_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                    const float* __restrict__ weight, int N, int K, double norm0, double norm1,
                                    double thr, float* __restrict__ acc, float* __restrict__ avg, int* __restrict__ cnt) {
  __shared__ double s_sum;
  __shared__ int s_cnt;
  if (threadIdx.x == 0) { s_sum = 0.0; s_cnt = 0; }
  __syncthreads();
  // one warp per keypoint, lanes over the crops (the fp64 divide / sqrt chain per crop is latency-bound when one
  // thread walks all N crops: 75 us for 64 x 17); hit / valid counts are integers, so the reduction order is free
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int k = warp; k < K; k += nwarps) {
    int valid = 0, hit = 0;
    for (int n = lane; n < N; n += 32) {
      const size_t i = static_cast<size_t>(n) * K + k;
      if (!(weight[i] > 0.f)) continue;
      const double dx = (static_cast<double>(pred[2 * i]) - static_cast<double>(gt[2 * i])) / norm0;
      const double dy = (static_cast<double>(pred[2 * i + 1]) - static_cast<double>(gt[2 * i + 1])) / norm1;
      const float d = static_cast<float>(sqrt(dx * dx + dy * dy));
      ++valid;
      if (static_cast<double>(d) < thr) ++hit;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      valid += __shfl_xor_sync(0xffffffffu, valid, off);
      hit += __shfl_xor_sync(0xffffffffu, hit, off);
    }
    if (lane == 0) {
      const double a = valid > 0 ? static_cast<double>(hit) / valid : -1.0;
      acc[k] = static_cast<float>(a);
      if (valid > 0) {
        atomicAdd(&s_sum, a);
        atomicAdd(&s_cnt, 1);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    *avg = s_cnt > 0 ? static_cast<float>(s_sum / s_cnt) : 0.f;
    *cnt = s_cnt;
  }
}

int pose_pck_accuracy(const float* pred, const float* gt, const float* weight, int N, int K, float norm0, float norm1,
                      float thr, float* acc, float* avg, int* cnt, cudaStream_t stream) {
  VPB_REQUIRE(N > 0 && K > 0 && norm0 > 0 && norm1 > 0, "pck_accuracy: bad shape");
  pck_accuracy_kernel<<<1, 256, 0, stream>>>(pred, gt, weight, N, K, norm0, norm1, static_cast<double>(thr), acc, avg,
                                             cnt);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vpb
