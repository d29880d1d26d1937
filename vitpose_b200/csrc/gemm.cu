#include "gemm.cuh"
#include "host_util.h"
#include "ops.h"

namespace vpb {

template <int BN, int EPI>
static int launch_gemm_inst(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int max_ctas,
                            cudaStream_t stream) {
  constexpr int smem = gemm_num_stages(BN) * gemm_stage_bytes(BN) + 1024;
  static bool configured = false;
  auto kern = gemm_bf16_tn_kernel<BN, EPI>;
  if (!configured) {
    VPB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (p.N + BN - 1) / BN;
  int grid = m_tiles * n_tiles;
  int cap = max_ctas > 0 ? max_ctas : sm_count();
  if (grid > cap) grid = cap;
  kern<<<grid, GEMM_THREADS, smem, stream>>>(ta, tb, p);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int gemm_pick_bn(int N, int epilogue) {
  if (epilogue == EPI_NCHW_F32) return N <= 32 ? 32 : (N <= 144 ? 144 : 256);
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  if (N <= 64) return 64;
  return (N % 256) > 128 || N > 1024 ? 256 : 128;
}

int make_gemm_maps(CUtensorMap* ta, CUtensorMap* tb, const void* A, const void* B, int M, int N, int K, int lda,
                   int ldb, int bn) {
  uint64_t dims_a[2] = {(uint64_t)K, (uint64_t)M};
  uint64_t str_a[1] = {(uint64_t)lda * 2};
  uint32_t box_a[2] = {GEMM_BK, GEMM_BM};
  if (make_tma_desc(ta, TMA_BF16, A, 2, dims_a, str_a, box_a, TMA_SWIZZLE_128B)) return -1;
  uint64_t dims_b[2] = {(uint64_t)K, (uint64_t)N};
  uint64_t str_b[1] = {(uint64_t)ldb * 2};
  uint32_t box_b[2] = {GEMM_BK, (uint32_t)bn};
  if (make_tma_desc(tb, TMA_BF16, B, 2, dims_b, str_b, box_b, TMA_SWIZZLE_128B)) return -1;
  return 0;
}

int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int bn, int epilogue,
                int max_ctas, cudaStream_t stream) {
#define VPB_GEMM_CASE(BN_, EPI_) \
  if (bn == BN_ && epilogue == EPI_) return launch_gemm_inst<BN_, EPI_>(ta, tb, p, max_ctas, stream);
  VPB_GEMM_CASE(256, EPI_BIAS_BF16)
  VPB_GEMM_CASE(128, EPI_BIAS_BF16)
  VPB_GEMM_CASE(64, EPI_BIAS_BF16)
  VPB_GEMM_CASE(256, EPI_GELU_BF16)
  VPB_GEMM_CASE(128, EPI_GELU_BF16)
  VPB_GEMM_CASE(64, EPI_GELU_BF16)
  VPB_GEMM_CASE(256, EPI_RESID_F32)
  VPB_GEMM_CASE(128, EPI_RESID_F32)
  VPB_GEMM_CASE(64, EPI_RESID_F32)
  VPB_GEMM_CASE(256, EPI_POS_F32)
  VPB_GEMM_CASE(128, EPI_POS_F32)
  VPB_GEMM_CASE(64, EPI_POS_F32)
  VPB_GEMM_CASE(32, EPI_NCHW_F32)
  VPB_GEMM_CASE(144, EPI_NCHW_F32)
  VPB_GEMM_CASE(256, EPI_NCHW_F32)
#undef VPB_GEMM_CASE
  set_last_error("gemm: no kernel instance for BN=%d epilogue=%d", bn, epilogue);
  return -2;
}

int gemm_bf16(const void* A, const void* B, int M, int N, int K, int epilogue, const float* bias, void* out, int ldo,
              const float* aux, int period, int max_ctas, cudaStream_t stream) {
  VPB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  VPB_REQUIRE(K % 8 == 0, "gemm: K=%d must be a multiple of 8 (16-byte TMA row pitch)", K);
  VPB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
              "gemm: operands must be 16-byte aligned");
  if (epilogue != EPI_NCHW_F32)
    VPB_REQUIRE(ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0, "gemm: out must be 16B aligned, ldo%%8==0");
  if (epilogue == EPI_RESID_F32 || epilogue == EPI_POS_F32) VPB_REQUIRE(aux != nullptr, "gemm: aux is null");
  if (epilogue == EPI_POS_F32 || epilogue == EPI_NCHW_F32) VPB_REQUIRE(period > 0, "gemm: period must be > 0");
  const int bn = gemm_pick_bn(N, epilogue);
  CUtensorMap ta, tb;
  if (make_gemm_maps(&ta, &tb, A, B, M, N, K, K, K, bn)) return -1;
  GemmParams p{M, N, K, bias, out, ldo, aux, period};
  return launch_gemm(ta, tb, p, bn, epilogue, max_ctas, stream);
}

}  // namespace vpb
