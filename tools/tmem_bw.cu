// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM (cycles per 32x32b.x32 access = 4 KB per warp) with
// 1, 4 (one per TMEM lane quadrant) and 8 (two per quadrant) warps issuing back to back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/tmem_bw tools/tmem_bw.cu && tools/bin/tmem_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../vitpose_b200/csrc/ptx.cuh"
using namespace vpb;

template <int MODE>   // 0: ld.x32, wait every access; 1: ld.x32, two in flight; 2: st.x32; 3: ld.x16
__global__ void __launch_bounds__(256, 1) k(long long* out, int active_warps, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < active_warps) {
    uint32_t va[32], vb[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) va[j] = vb[j] = j;
    if (MODE != 2) { tmem_st_32x32b_x32(base, va); tmem_st_wait(); }
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t col = (i * 32) & 255;
      if (MODE == 0) {
        tmem_ld_32x32b_x32(base + col, va);
        tmem_ld_wait();
        acc += va[i & 31];
      } else if (MODE == 1) {
        tmem_ld_32x32b_x32(base + col, va);
        tmem_ld_32x32b_x32(base + ((col + 32) & 255), vb);
        tmem_ld_wait();
        acc += va[i & 31] + vb[(i + 1) & 31];
      } else if (MODE == 2) {
        va[0] = i;
        tmem_st_32x32b_x32(base + col, va);
        tmem_st_wait();
      } else {
        uint32_t r[16];
        tmem_ld_32x32b_x16(base + col, r);
        tmem_ld_wait();
        acc += r[i & 15];
      }
    }
    t1 = clock64();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
  if ((threadIdx.x & 31) == 0 && warp < active_warps) {
    out[(blockIdx.x * 8 + warp) * 2] = t1 - t0;
    out[(blockIdx.x * 8 + warp) * 2 + 1] = acc;
  }
}

template <int MODE>
void run(const char* name, int bytes_per_iter, long long* d, long long* h) {
  const int iters = 4000;
  for (int w : {1, 4, 8}) {
    k<MODE><<<148, 256>>>(d, w, iters);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, 148 * 8 * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int b = 0; b < 148; ++b)
      for (int i = 0; i < w; ++i) mx = h[(b * 8 + i) * 2] > mx ? h[(b * 8 + i) * 2] : mx;
    const double cyc = mx / iters;
    printf("%-28s warps=%d: %.1f cycles per access per warp -> %.1f B/clk/SM\n", name, w, cyc, w * bytes_per_iter / cyc);
  }
}

int main() {
  long long *d, *h = new long long[148 * 16];
  cudaMalloc(&d, 148 * 16 * sizeof(long long));
  run<0>("tcgen05.ld x32 (wait each)", 4096, d, h);
  run<1>("tcgen05.ld x32 (2 in flight)", 8192, d, h);
  run<3>("tcgen05.ld x16 (wait each)", 2048, d, h);
  run<2>("tcgen05.st x32 (wait each)", 4096, d, h);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
