// Host-side helpers shared by the C-ABI translation units: error reporting and TMA descriptors.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace vpb {

void set_last_error(const char* fmt, ...);
const char* get_last_error();

#define VPB_CHECK_CUDA(expr)                                                                   \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      vpb::set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -1;                                                                               \
    }                                                                                          \
  } while (0)

#define VPB_REQUIRE(cond, ...)             \
  do {                                     \
    if (!(cond)) {                         \
      vpb::set_last_error(__VA_ARGS__);    \
      return -2;                           \
    }                                      \
  } while (0)

enum TmaSwizzle { TMA_SWIZZLE_NONE = 0, TMA_SWIZZLE_32B = 1, TMA_SWIZZLE_64B = 2, TMA_SWIZZLE_128B = 3 };
enum TmaDtype { TMA_BF16 = 0, TMA_F32 = 1 };

// Encodes a tiled tensor map (rank <= 5). dims/box are innermost-first, in elements; strides_bytes has
// rank-1 entries (stride of dim 1..rank-1). Out-of-bounds elements read as zero. Returns 0 on success.
int make_tma_desc(CUtensorMap* out, TmaDtype dtype, const void* base, int rank, const uint64_t* dims,
                  const uint64_t* strides_bytes, const uint32_t* box, TmaSwizzle swizzle);

int sm_count();

// Programmatic dependent launch for the kernels of the forward sequence (VPB_PDL=0 switches it off, A/B): fills
// `attr` and returns 1 when enabled, else returns 0. Only kernels that call pdl_wait() may be launched with it.
int pdl_launch_attr(cudaLaunchAttribute* attr);

}  // namespace vpb
