// Fused test-time preprocessing that feeds the hot path:
//   TopDownAffine   cv2.warpAffine(img, trans, (W, H), flags=INTER_LINEAR)      top_down_transform.py:324-347
//   ToTensor        uint8 HWC -> float CHW / 255                                 shared_transform.py:21-36
//   NormalizeTensor (x - mean) / std                                             shared_transform.py:39-65
// One thread per output pixel: the inverse map is evaluated exactly as OpenCV does on uint8 images (double-precision
// products rounded half-to-even to 10-bit fixed point, 5-bit bilinear fractions, 15-bit weights, rounding shift,
// constant-0 border), so the uint8 warp — and therefore the float crop — is bit-identical to the CPU pipeline.
// Reads 4 source pixels (x3 channels) per output pixel, writes the fp32 NCHW crop coalesced along x.
#include "host_util.h"
#include "ops.h"

namespace vpb {

struct WarpParams {
  const unsigned char* const* src;   // [n] device pointers to uint8 HWC images
  const int* src_hw;                 // [n,2] height, width
  const double* inv;                 // [n,6] inverse (dst -> src) affine map, row-major 2x3
  float* out;                        // [n,3,H,W]
  int n, H, W;
  float mean[3], std[3];
};

__global__ void __launch_bounds__(256) warp_affine_normalize_kernel(const WarpParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  const int i = blockIdx.z;
  if (x >= p.W) return;
  const double* M = p.inv + 6 * i;
  const int h = p.src_hw[2 * i], w = p.src_hw[2 * i + 1];
  const unsigned char* s = p.src[i];
  // cvRound == round half to even
  const long long adelta = __double2ll_rn(M[0] * x * 1024.0);
  const long long bdelta = __double2ll_rn(M[3] * x * 1024.0);
  const long long X0 = __double2ll_rn((M[1] * y + M[2]) * 1024.0) + 16;
  const long long Y0 = __double2ll_rn((M[4] * y + M[5]) * 1024.0) + 16;
  const long long X = (X0 + adelta) >> 5, Y = (Y0 + bdelta) >> 5;
  long long sx = X >> 5, sy = Y >> 5;
  sx = sx < -32768 ? -32768 : (sx > 32767 ? 32767 : sx);      // saturate_cast<short>
  sy = sy < -32768 ? -32768 : (sy > 32767 ? 32767 : sy);
  const int fx = static_cast<int>(X & 31), fy = static_cast<int>(Y & 31);
  const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
  const int x0 = static_cast<int>(sx), y0 = static_cast<int>(sy);
  const bool okx0 = x0 >= 0 && x0 < w, okx1 = x0 + 1 >= 0 && x0 + 1 < w;
  const bool oky0 = y0 >= 0 && y0 < h, oky1 = y0 + 1 >= 0 && y0 + 1 < h;
  const size_t pitch = static_cast<size_t>(w) * 3;
  const unsigned char* r0 = s + static_cast<size_t>(oky0 ? y0 : 0) * pitch;
  const unsigned char* r1 = s + static_cast<size_t>(oky1 ? y0 + 1 : 0) * pitch;
  const size_t plane = static_cast<size_t>(p.H) * p.W;
  float* o = p.out + (static_cast<size_t>(i) * 3) * plane + static_cast<size_t>(y) * p.W + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int p00 = (oky0 && okx0) ? r0[x0 * 3 + c] : 0;
    const int p01 = (oky0 && okx1) ? r0[(x0 + 1) * 3 + c] : 0;
    const int p10 = (oky1 && okx0) ? r1[x0 * 3 + c] : 0;
    const int p11 = (oky1 && okx1) ? r1[(x0 + 1) * 3 + c] : 0;
    const int v = (p00 * w00 + p01 * w01 + p10 * w10 + p11 * w11 + (1 << 14)) >> 15;
    const float f = __fdiv_rn(static_cast<float>(v & 255), 255.0f);
    o[c * plane] = __fdiv_rn(__fsub_rn(f, p.mean[c]), p.std[c]);
  }
}

int warp_affine_normalize(const unsigned char* const* src_ptrs, const int* src_hw, const double* inv_mats, int n,
                          int out_h, int out_w, const float* mean3, const float* std3, float* out,
                          cudaStream_t stream) {
  VPB_REQUIRE(n >= 0 && out_h > 0 && out_w > 0, "warp_affine_normalize: bad shape");
  if (n == 0) return 0;
  WarpParams p;
  p.src = src_ptrs; p.src_hw = src_hw; p.inv = inv_mats; p.out = out; p.n = n; p.H = out_h; p.W = out_w;
  for (int c = 0; c < 3; ++c) { p.mean[c] = mean3[c]; p.std[c] = std3[c]; }
  dim3 block(256), grid((out_w + 255) / 256, out_h, n);
  if (out_w <= 192) { block = dim3(192); grid = dim3((out_w + 191) / 192, out_h, n); }
  warp_affine_normalize_kernel<<<grid, block, 0, stream>>>(p);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vpb
