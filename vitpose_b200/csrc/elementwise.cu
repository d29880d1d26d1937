// Memory-bound helpers around the tensor-core kernels: patch im2col (+ test-time flip), LayerNorm,
// token-major -> NCHW export, ReLU + bilinear upsample (simple decoder). All are single-pass, 8/16-byte
// vectorised, one read and one write of each element.
#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

namespace vpb {

// -------------------------------------------------------------------------------------------------
// PatchEmbed as a GEMM operand (reference: Conv2d(3, D, 16, stride 16, padding 2), vit.py:157-165).
// Row = (crop, patch_y, patch_x), column = c*256 + ky*16 + kx (the flattened conv weight order).
// The flipped crops of the flip test (top_down.py:180, img.flip(3)) are generated here as extra rows, so
// the flipped batch never exists in memory as an image.
// One thread = one (row, c, ky) segment of 16 pixels: 64-byte read, 32-byte write.
// -------------------------------------------------------------------------------------------------
__global__ void im2col_patch16_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int n, int H,
                                      int W, int Hp, int Wp, int total_segments) {
  const int seg = blockIdx.x * blockDim.x + threadIdx.x;
  if (seg >= total_segments) return;
  const int cky = seg % 48;                // c * 16 + ky
  const int row = seg / 48;
  const int c = cky >> 4, ky = cky & 15;
  const int T = Hp * Wp;
  const int im = row / T;
  const int t = row - im * T;
  const int pi = t / Wp, pj = t - pi * Wp;
  const bool flipped = im >= n;
  const int src_im = flipped ? im - n : im;
  const int y = pi * 16 - 2 + ky;
  float v[16];
  if (y < 0 || y >= H) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.f;
  } else {
    const float* rowp = img + (static_cast<size_t>(src_im * 3 + c) * H + y) * W;
    // output kx = 0..15 reads x = x0 + kx (plain) or x = W-1-(x0+kx) (flipped), x0 = 16*pj - 2
    const int x0 = pj * 16 - 2;
    const int lo = flipped ? (W - 1 - (x0 + 15)) : x0;    // ascending source range [lo, lo+15], lo is even
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int x = lo + 2 * i;
      float2 f = make_float2(0.f, 0.f);
      if (x >= 0 && x + 1 < W) {
        f = __ldg(reinterpret_cast<const float2*>(rowp + x));
      } else {
        if (x >= 0 && x < W) f.x = __ldg(rowp + x);
        if (x + 1 >= 0 && x + 1 < W) f.y = __ldg(rowp + x + 1);
      }
      if (flipped) { v[15 - 2 * i] = f.x; v[14 - 2 * i] = f.y; }
      else         { v[2 * i] = f.x; v[2 * i + 1] = f.y; }
    }
  }
  uint4 w0 = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                        pack_bf16x2(v[6], v[7]));
  uint4 w1 = make_uint4(pack_bf16x2(v[8], v[9]), pack_bf16x2(v[10], v[11]), pack_bf16x2(v[12], v[13]),
                        pack_bf16x2(v[14], v[15]));
  uint4* o = reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * 768 + cky * 16);
  o[0] = w0;
  o[1] = w1;
}

// Tiled variant (W <= 256, 16-byte aligned image rows): one CTA = one (image, channel, patch row): the 16 image rows
// y = 16*pi - 2 .. 16*pi + 13 are read once with coalesced float4 loads and kept as bf16 in shared memory; every
// (patch, channel) block of the output is 512 contiguous bytes, written by one warp with 16 bytes per lane — for the
// plain AND the flipped batch from the same tile (the flipped rows read it mirrored). The per-thread kernel above read
// 8 bytes per lane from 32 different lines per instruction, and the image twice.
constexpr int IM2COL_MAX_W = 256;
__global__ void __launch_bounds__(256) im2col_patch16_tiled_kernel(const float* __restrict__ img,
                                                                   __nv_bfloat16* __restrict__ out, int n, int H, int W,
                                                                   int Hp, int Wp, int flip) {
  __shared__ __align__(16) __nv_bfloat16 tile[16][IM2COL_MAX_W + 8];
  const int pi = blockIdx.x % Hp;
  const int c = (blockIdx.x / Hp) % 3;
  const int im = blockIdx.x / (3 * Hp);
  const int W4 = W / 4;
  for (int idx = threadIdx.x; idx < 16 * W4; idx += 256) {
    const int ky = idx / W4, x4 = idx - ky * W4;
    const int y = pi * 16 - 2 + ky;
    uint2 w = make_uint2(0u, 0u);
    if (y >= 0 && y < H) {
      const float4 f = __ldg(reinterpret_cast<const float4*>(img + (static_cast<size_t>(im * 3 + c) * H + y) * W) + x4);
      w = make_uint2(pack_bf16x2(f.x, f.y), pack_bf16x2(f.z, f.w));
    }
    *reinterpret_cast<uint2*>(&tile[ky][4 * x4]) = w;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ky = lane >> 1, h = lane & 1;                  // lane -> 16 bytes (kx = 8h .. 8h+7) of row ky
  const int T = Hp * Wp;
  const int units = (flip ? 2 : 1) * Wp;                   // (variant, patch column)
  for (int u = warp; u < units; u += 8) {
    const int flipped = u >= Wp;
    const int pj = flipped ? u - Wp : u;
    uint32_t w[4];
    if (!flipped) {
      const int x0 = pj * 16 - 2 + 8 * h;                  // even; source x = x0 + j
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = x0 + 2 * j;
        w[j] = (x >= 0 && x + 1 < W) ? *reinterpret_cast<const uint32_t*>(&tile[ky][x]) : 0u;   // x, x+1 in or out together
      }
    } else {
      const int xs = W + 1 - pj * 16 - 8 * h;              // odd; source x = xs - j  (x = W-1-(16*pj-2+kx))
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = xs - 2 * j;                          // pair (x, x-1), x odd
        const uint32_t v = (x < W && x - 1 >= 0) ? *reinterpret_cast<const uint32_t*>(&tile[ky][x - 1]) : 0u;
        w[j] = (v >> 16) | (v << 16);                      // element x first, then x-1
      }
    }
    const size_t row = static_cast<size_t>(flipped ? im + n : im) * T + pi * Wp + pj;
    *reinterpret_cast<uint4*>(out + row * 768 + c * 256 + ky * 16 + 8 * h) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

int im2col_patch16(const float* img, void* patches, int n, int H, int W, int flip, cudaStream_t stream) {
  VPB_REQUIRE(n > 0 && H % 16 == 0 && W % 16 == 0 && W % 2 == 0, "im2col: bad shape n=%d H=%d W=%d", n, H, W);
  VPB_REQUIRE((reinterpret_cast<uintptr_t>(img) & 7) == 0 && (reinterpret_cast<uintptr_t>(patches) & 15) == 0,
              "im2col: misaligned pointers");
  const int Hp = (H + 4 - 16) / 16 + 1, Wp = (W + 4 - 16) / 16 + 1;
  if (W <= IM2COL_MAX_W && (reinterpret_cast<uintptr_t>(img) & 15) == 0 && !getenv("VPB_IM2COL_SIMPLE")) {
    im2col_patch16_tiled_kernel<<<static_cast<unsigned>(n) * 3 * Hp, 256, 0, stream>>>(
        img, reinterpret_cast<__nv_bfloat16*>(patches), n, H, W, Hp, Wp, flip);
    VPB_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const long long rows = static_cast<long long>(flip ? 2 * n : n) * Hp * Wp;
  const long long total = rows * 48;
  VPB_REQUIRE(total < (1ll << 31), "im2col: batch too large");
  const int threads = 256;
  im2col_patch16_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0, stream>>>(
      img, reinterpret_cast<__nv_bfloat16*>(patches), n, H, W, Hp, Wp, static_cast<int>(total));
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// LayerNorm (nn.LayerNorm(D, eps=1e-6), vit.py:125,133,212,242): one warp per token row, the row lives in
// registers (NV float4 per lane), two-pass mean / variance in fp32, bf16 output feeding the next GEMM's
// A operand. Reads M*D*4 bytes, writes M*D*2.
// -------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                        int M, float eps) {
  constexpr int D = NV * 128;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * D);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[i * 32 + lane];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
  const float mean = s * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) q += __shfl_xor_sync(0xffffffffu, q, off);
  const float rstd = 1.0f / sqrtf(q * (1.0f / D) + eps);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<size_t>(row) * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
    const float o0 = (v[i].x - mean) * rstd * g.x + b.x;
    const float o1 = (v[i].y - mean) * rstd * g.y + b.y;
    const float o2 = (v[i].z - mean) * rstd * g.z + b.z;
    const float o3 = (v[i].w - mean) * rstd * g.w + b.w;
    yr[i * 32 + lane] = make_uint2(pack_bf16x2(o0, o1), pack_bf16x2(o2, o3));
  }
}

int layernorm_bf16(const float* x, const float* gamma, const float* beta, void* y, int M, int D, float eps,
                   cudaStream_t stream) {
  VPB_REQUIRE(M > 0 && D > 0 && D % 128 == 0, "layernorm: D=%d must be a multiple of 128", D);
  const int warps = 8;
  dim3 grid((M + warps - 1) / warps), block(warps * 32);
  __nv_bfloat16* yo = reinterpret_cast<__nv_bfloat16*>(y);
  switch (D / 128) {
#define VPB_LN_CASE(NV_) \
  case NV_: layernorm_kernel<NV_><<<grid, block, 0, stream>>>(x, gamma, beta, yo, M, eps); break;
    VPB_LN_CASE(1) VPB_LN_CASE(2) VPB_LN_CASE(3) VPB_LN_CASE(4) VPB_LN_CASE(5) VPB_LN_CASE(6) VPB_LN_CASE(8)
    VPB_LN_CASE(10) VPB_LN_CASE(12) VPB_LN_CASE(16)
#undef VPB_LN_CASE
    default:
      set_last_error("layernorm: unsupported D=%d", D);
      return -2;
  }
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// ViT.forward's final permute (vit.py:330): tokens [n, T, D] bf16 -> [n, D, T] fp32 through a padded
// 32x32 shared-memory transpose. Only used when the backbone is called standalone; the fused path keeps
// token-major (= NHWC) features for the head.
// -------------------------------------------------------------------------------------------------
__global__ void tokens_to_nchw_kernel(const __nv_bfloat16* __restrict__ tok, float* __restrict__ out, int T, int D) {
  __shared__ float tile[32][33];
  const int im = blockIdx.z;
  const int t0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int t = t0 + i, d = d0 + threadIdx.x;
    tile[i][threadIdx.x] = (t < T && d < D) ? __bfloat162float(tok[(static_cast<size_t>(im) * T + t) * D + d]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int d = d0 + i, t = t0 + threadIdx.x;
    if (t < T && d < D) out[(static_cast<size_t>(im) * D + d) * T + t] = tile[threadIdx.x][i];
  }
}

int tokens_to_nchw_f32(const void* tokens, float* out, int n, int T, int D, cudaStream_t stream) {
  VPB_REQUIRE(n > 0 && T > 0 && D > 0, "tokens_to_nchw: bad shape");
  dim3 grid((T + 31) / 32, (D + 31) / 32, n), block(32, 8);
  tokens_to_nchw_kernel<<<grid, block, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(tokens), out, T, D);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Simple decoder input transform (topdown_heatmap_simple_head.py:278-287):
// F.interpolate(relu(x), scale_factor=f, mode='bilinear', align_corners=False) on NHWC bf16.
// src = (dst + 0.5) / f - 0.5 clamped at 0 (PyTorch's area_pixel_compute_source_index), neighbours clamped
// to the last row/column. One thread = 8 channels of one output pixel (16-byte accesses).
// -------------------------------------------------------------------------------------------------
// IdxT = unsigned when the output fits 32-bit indexing (64-bit divisions by run-time values cost ~100 instructions each
// and outweighed the interpolation itself).
template <typename IdxT>
__global__ void relu_upsample_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int h,
                                     int w, int C, int f, long long total_vec) {
  const IdxT idx = static_cast<IdxT>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (static_cast<long long>(idx) >= total_vec) return;
  const int cv = C / 8;
  const int c8 = static_cast<int>(idx % cv);
  IdxT pix = idx / cv;
  const int W2 = w * f, H2 = h * f;
  const int ox = static_cast<int>(pix % W2);
  pix /= W2;
  const int oy = static_cast<int>(pix % H2);
  const int im = static_cast<int>(pix / H2);
  const float inv = 1.0f / f;
  float sy = (oy + 0.5f) * inv - 0.5f, sx = (ox + 0.5f) * inv - 0.5f;
  sy = sy < 0.f ? 0.f : sy;
  sx = sx < 0.f ? 0.f : sx;
  const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly = sy - y0, lx = sx - x0, hy = 1.f - ly, hx = 1.f - lx;
  auto ld = [&](int yy, int xx) {
    return __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<size_t>(im) * h + yy) * w + xx) * C) + c8);
  };
  const uint4 a = ld(y0, x0), b = ld(y0, x1), c = ld(y1, x0), d = ld(y1, x1);
  const uint32_t* ap = &a.x; const uint32_t* bp = &b.x; const uint32_t* cp = &c.x; const uint32_t* dp = &d.x;
  uint32_t o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 fa = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(ap + i));
    const float2 fb = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(bp + i));
    const float2 fc = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(cp + i));
    const float2 fd = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(dp + i));
    const float r0 = hy * (hx * fmaxf(fa.x, 0.f) + lx * fmaxf(fb.x, 0.f)) +
                     ly * (hx * fmaxf(fc.x, 0.f) + lx * fmaxf(fd.x, 0.f));
    const float r1 = hy * (hx * fmaxf(fa.y, 0.f) + lx * fmaxf(fb.y, 0.f)) +
                     ly * (hx * fmaxf(fc.y, 0.f) + lx * fmaxf(fd.y, 0.f));
    o[i] = pack_bf16x2(r0, r1);
  }
  reinterpret_cast<uint4*>(out)[idx] = make_uint4(o[0], o[1], o[2], o[3]);
}

int relu_upsample_bilinear_nhwc(const void* in, void* out, int n, int h, int w, int C, int factor,
                                cudaStream_t stream) {
  VPB_REQUIRE(n > 0 && C % 8 == 0 && factor >= 1, "relu_upsample: bad shape n=%d C=%d factor=%d", n, C, factor);
  const long long total = static_cast<long long>(n) * h * factor * w * factor * (C / 8);
  const int threads = 256;
  const unsigned grid = static_cast<unsigned>((total + threads - 1) / threads);
  if (total + threads < (1ll << 31))
    relu_upsample_kernel<unsigned><<<grid, threads, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), h, w, C, factor, total);
  else
    relu_upsample_kernel<long long><<<grid, threads, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(in), reinterpret_cast<__nv_bfloat16*>(out), h, w, C, factor, total);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Simple decoder without the upsampled map. conv3x3(pad 1)(upsample_x4(relu(x))) is linear in r = relu(x) and the
// bilinear upsample acts on every channel separately, so
//   heatmap[k, Y, X] = bias[k] + sum_{tap t = (ky, kx)} [ (Y+ky-1, X+kx-1) inside the upsampled map ]
//                                 * upsample(z_t[k])[Y+ky-1, X+kx-1],      z_t[k] = sum_c W[k, c, ky, kx] * r[c]
// i.e. nine 1x1 convolutions on the 16 x 12 token grid (ONE GEMM with 9K output columns) followed by a gather that
// interpolates the nine small maps. 16x fewer FLOPs than the convolution on the 64 x 48 map, the
// [images, 64, 48, D] bf16 map (12.9 GB for ViTPose-L at 1024 crops + flips) is never written, and the interpolation
// runs in fp32 instead of on a bf16-rounded map. (topdown_heatmap_simple_head.py:278-287 + final 3x3 conv :132-139)
// -------------------------------------------------------------------------------------------------
__global__ void relu_bf16_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, long long n_vec) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_vec) return;
  uint4 v = __ldg(in + i);
  uint32_t* w = &v.x;
  const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(w + j);
    x = __hmax2(x, zero);
    w[j] = *reinterpret_cast<uint32_t*>(&x);
  }
  out[i] = v;
}

int relu_bf16(const void* in, void* out, long long n, cudaStream_t stream) {
  VPB_REQUIRE(n > 0 && n % 8 == 0, "relu_bf16: element count %lld must be a positive multiple of 8", n);
  const long long n_vec = n / 8;
  relu_bf16_kernel<<<static_cast<unsigned>((n_vec + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), n_vec);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// z fp32 [images, K * 9, h * w] (column k * 9 + t of the tap GEMM, token-contiguous) -> out fp32 [images, K, h*f, w*f].
// One CTA per (image, keypoint). The bilinear upsample is separable: the nine tap maps are first interpolated along x
// into shared memory (u[t][i][xx], 9 x h x W values), then every output pixel needs 2 multiply-adds per tap with the
// row weights of its three source rows (tables in shared memory) — ~40 instructions per pixel instead of ~250 with a
// per-tap 2-D interpolation (1.06 -> see profiles/r02_summary.md for 1024 images x 17 keypoints).
__global__ void __launch_bounds__(256) simple_head_gather_kernel(const float* __restrict__ z,
                                                                  const float* __restrict__ bias, float* __restrict__ out,
                                                                  int K, int h, int w, int f) {
  extern __shared__ float s_dyn[];
  const int img = blockIdx.x / K, k = blockIdx.x % K;
  const int T = h * w, H = h * f, W = w * f;
  float* s_z = s_dyn;                            // [9][h * w]
  float* s_u = s_z + 9 * T;                      // [9][h][W]: tap maps interpolated along x
  int* s_y0 = reinterpret_cast<int*>(s_u + 9 * h * W);   // [H]: upper source row (the lower one is min(y0 + 1, h - 1))
  float* s_ly = reinterpret_cast<float*>(s_y0 + H);      // [H]: weight of the lower row
  const float* src = z + (static_cast<size_t>(img) * K + k) * 9 * T;
  for (int i = threadIdx.x; i < 9 * T; i += blockDim.x) s_z[i] = __ldg(src + i);
  const float inv = 1.0f / f;
  for (int y = threadIdx.x; y < H; y += blockDim.x) {
    float sy = (y + 0.5f) * inv - 0.5f;          // PyTorch's area_pixel_compute_source_index (align_corners=False)
    sy = sy < 0.f ? 0.f : sy;
    const int y0 = static_cast<int>(sy);
    s_y0[y] = y0;
    s_ly[y] = sy - y0;
  }
  __syncthreads();
  {
    // (xx, ti) walked incrementally: no division by the run-time width in the loop
    const int step_x = static_cast<int>(blockDim.x) % W, step_r = static_cast<int>(blockDim.x) / W;
    int xx = static_cast<int>(threadIdx.x) % W, ti = static_cast<int>(threadIdx.x) / W;   // ti = t * h + i
    for (int e = threadIdx.x; e < 9 * h * W; e += blockDim.x) {
      float sx = (xx + 0.5f) * inv - 0.5f;
      sx = sx < 0.f ? 0.f : sx;
      const int x0 = static_cast<int>(sx), x1 = x0 + (x0 < w - 1 ? 1 : 0);
      const float lx = sx - x0;
      const float* row = s_z + ti * w;
      s_u[e] = (1.f - lx) * row[x0] + lx * row[x1];
      xx += step_x;
      ti += step_r;
      if (xx >= W) { xx -= W; ++ti; }
    }
  }
  __syncthreads();
  const float b = bias != nullptr ? __ldg(bias + k) : 0.f;
  float* o = out + (static_cast<size_t>(img) * K + k) * H * W;
  const int step_x = static_cast<int>(blockDim.x) % W, step_y = static_cast<int>(blockDim.x) / W;
  int X = static_cast<int>(threadIdx.x) % W, Y = static_cast<int>(threadIdx.x) / W;
  for (int p = threadIdx.x; p < H * W; p += blockDim.x, X += step_x, Y += step_y) {
    if (X >= W) { X -= W; ++Y; }
    float acc = b;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = Y + ky - 1;
      if (yy < 0 || yy >= H) continue;           // zero padding of the conv on the upsampled map
      const int y0 = s_y0[yy], y1 = y0 + (y0 < h - 1 ? 1 : 0);
      const float ly = s_ly[yy], hy = 1.f - ly;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = X + kx - 1;
        if (xx < 0 || xx >= W) continue;
        const float* u = s_u + (ky * 3 + kx) * h * W;
        acc += hy * u[y0 * W + xx] + ly * u[y1 * W + xx];
      }
    }
    o[p] = acc;
  }
}

int simple_head_gather(const float* z, const float* bias, float* out, int images, int K, int h, int w, int factor,
                       cudaStream_t stream) {
  VPB_REQUIRE(images > 0 && K > 0 && factor >= 1, "simple_head_gather: bad shape");
  const size_t smem = (static_cast<size_t>(9) * h * w + static_cast<size_t>(9) * h * w * factor + 2 * static_cast<size_t>(h) * factor) *
                      sizeof(float);
  VPB_REQUIRE(smem <= 48 * 1024, "simple_head_gather: token grid %d x %d (x%d) too large", h, w, factor);
  simple_head_gather_kernel<<<static_cast<unsigned>(images) * K, 256, smem, stream>>>(z, bias, out, K, h, w, factor);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Backward of simple_head_gather (training of the simple decoder): dout fp32 [images, K, h*f, w*f] ->
// dz bf16 [images * h * w, ldz] (TOKEN-major rows, column k * 9 + t: the A operand of the tap GEMM's input- and
// weight-gradient GEMMs). out = sum_t shift_t(bilinear(z_t)), so dz_t = bilinear^T(shift_t^T(dout)), and the transposed
// interpolation is separable like the forward one: first along y into shared memory (v[t][i][xx]), then along x.
// One CTA per (image, keypoint); every (t, i, xx) / (t, i, j) element looks at the <= 3 f source rows / columns whose
// interpolation touches it, with the same index / weight formulas as the forward kernel.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) simple_head_gather_bwd_kernel(const float* __restrict__ dout,
                                                                      __nv_bfloat16* __restrict__ dz, int ldz, int K,
                                                                      int h, int w, int f) {
  extern __shared__ float s_dyn[];
  const int img = blockIdx.x / K, k = blockIdx.x % K;
  const int T = h * w, H = h * f, W = w * f;
  float* s_d = s_dyn;                           // [H][W]
  float* s_v = s_d + H * W;                     // [9][h][W]
  const float* src = dout + (static_cast<size_t>(img) * K + k) * H * W;
  for (int i = threadIdx.x; i < H * W; i += blockDim.x) s_d[i] = __ldg(src + i);
  __syncthreads();
  const float inv = 1.0f / f;
  auto src_index = [&](int o, int n_src, int& i0, int& i1, float& l) {   // as the forward kernel (align_corners=False)
    float s = (o + 0.5f) * inv - 0.5f;
    s = s < 0.f ? 0.f : s;
    i0 = static_cast<int>(s);
    i1 = i0 + (i0 < n_src - 1 ? 1 : 0);
    l = s - i0;
  };
  for (int e = threadIdx.x; e < 9 * h * W; e += blockDim.x) {
    const int xx = e % W, i = (e / W) % h, t = e / (W * h);
    const int ky = t / 3, kx = t % 3;
    const int Xo = xx - kx + 1;                 // output column whose tap t reads upsampled column xx
    float acc = 0.f;
    if (Xo >= 0 && Xo < W) {
      const int lo = max(0, f * (i - 1)), hi = min(H - 1, f * (i + 2));
      for (int yy = lo; yy <= hi; ++yy) {
        int y0, y1;
        float ly;
        src_index(yy, h, y0, y1, ly);
        const float wgt = (y0 == i ? 1.f - ly : 0.f) + (y1 == i ? ly : 0.f);
        const int Yo = yy - ky + 1;
        if (wgt != 0.f && Yo >= 0 && Yo < H) acc = fmaf(wgt, s_d[Yo * W + Xo], acc);
      }
    }
    s_v[e] = acc;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 9 * T; e += blockDim.x) {
    const int j = e % w, i = (e / w) % h, t = e / T;
    const float* row = s_v + (t * h + i) * W;
    float acc = 0.f;
    const int lo = max(0, f * (j - 1)), hi = min(W - 1, f * (j + 2));
    for (int xx = lo; xx <= hi; ++xx) {
      int x0, x1;
      float lx;
      src_index(xx, w, x0, x1, lx);
      const float wgt = (x0 == j ? 1.f - lx : 0.f) + (x1 == j ? lx : 0.f);
      acc = fmaf(wgt, row[xx], acc);
    }
    dz[(static_cast<size_t>(img) * T + i * w + j) * ldz + k * 9 + t] = __float2bfloat16_rn(acc);
  }
}

int simple_head_gather_bwd(const float* dout, void* dz, int ldz, int images, int K, int h, int w, int factor,
                           cudaStream_t stream) {
  VPB_REQUIRE(images > 0 && K > 0 && factor >= 1 && ldz >= 9 * K, "simple_head_gather_bwd: bad shape (ldz %d < 9 K)", ldz);
  const size_t smem = (static_cast<size_t>(h) * factor * w * factor + static_cast<size_t>(9) * h * w * factor) * sizeof(float);
  VPB_REQUIRE(smem <= 48 * 1024, "simple_head_gather_bwd: token grid %d x %d (x%d) too large", h, w, factor);
  simple_head_gather_bwd_kernel<<<static_cast<unsigned>(images) * K, 256, smem, stream>>>(
      dout, reinterpret_cast<__nv_bfloat16*>(dz), ldz, K, h, w, factor);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// dx = dy where y > 0 else 0 (y = relu(x) as the forward pass stored it), 8 bf16 per thread
__global__ void relu_bwd_bf16_kernel(const uint4* __restrict__ y, const uint4* __restrict__ dy, uint4* __restrict__ dx,
                                     long long n_vec) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_vec) return;
  const uint4 a = __ldg(y + i);
  uint4 d = __ldg(dy + i);
  const uint32_t* aw = &a.x;
  uint32_t* dw = &d.x;
#pragma unroll
  for (int j = 0; j < 4; ++j) {       // bf16 > 0: sign bit clear and not (+)zero
    const uint32_t lo = aw[j] & 0xffffu, hi = aw[j] >> 16;
    const uint32_t mlo = (lo != 0u && (lo & 0x8000u) == 0u) ? 0xffffu : 0u;
    const uint32_t mhi = (hi != 0u && (hi & 0x8000u) == 0u) ? 0xffff0000u : 0u;
    dw[j] &= (mlo | mhi);
  }
  dx[i] = d;
}
int relu_bwd_bf16(const void* y, const void* dy, void* dx, long long n, cudaStream_t stream) {
  VPB_REQUIRE(n > 0 && n % 8 == 0, "relu_bwd_bf16: element count %lld must be a positive multiple of 8", n);
  const long long n_vec = n / 8;
  relu_bwd_bf16_kernel<<<static_cast<unsigned>((n_vec + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const uint4*>(y), reinterpret_cast<const uint4*>(dy), reinterpret_cast<uint4*>(dx), n_vec);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------------
// Weights of a Linear layer that applies the LayerNorm in front of it in its own epilogue (gemm.cuh, "folded"
// LayerNorm): Wf = bf16(gamma o W), s_n = sum_k float(Wf[n, k]) (the rounded values the tensor cores multiply by, so that
// the mean term cancels exactly), c_n = b_n + sum_k beta_k W[n, k]. One CTA per output row; one-time weight repack.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fold_ln_linear_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, int K,
                                                             __nv_bfloat16* __restrict__ Wf, float* __restrict__ s_out,
                                                             float* __restrict__ c_out) {
  const int n = blockIdx.x;
  const float* w = W + static_cast<size_t>(n) * K;
  float s = 0.f, c = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float wk = __ldg(w + k);
    const __nv_bfloat16 wf = __float2bfloat16_rn(wk * __ldg(gamma + k));
    Wf[static_cast<size_t>(n) * K + k] = wf;
    s += __bfloat162float(wf);
    c = fmaf(__ldg(beta + k), wk, c);
  }
  __shared__ float red[2][8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float ss = 0.f, cc = 0.f;
    for (int i = 0; i < 8; ++i) { ss += red[0][i]; cc += red[1][i]; }
    s_out[n] = ss;
    c_out[n] = cc + (bias != nullptr ? bias[n] : 0.f);
  }
}

int fold_layernorm_linear(const float* W, const float* bias, const float* gamma, const float* beta, int N, int K,
                          void* Wf, float* s, float* c, cudaStream_t stream) {
  VPB_REQUIRE(N > 0 && K > 0 && W && gamma && beta && Wf && s && c, "fold_layernorm_linear: bad argument");
  fold_ln_linear_kernel<<<N, 256, 0, stream>>>(W, bias, gamma, beta, K, reinterpret_cast<__nv_bfloat16*>(Wf), s, c);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vpb
