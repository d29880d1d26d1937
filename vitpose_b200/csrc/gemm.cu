#include <stdlib.h>

#include "gemm.cuh"
#include "host_util.h"
#include "ops.h"

namespace vpb {

template <int BN, int EPI, int CG>
static int launch_gemm_inst(const GemmMaps& maps, const GemmParams& p, int max_ctas, cudaStream_t stream) {
  constexpr int smem = gemm_smem_bytes(BN, EPI, CG);
  static bool configured = false;
  auto kern = gemm_bf16_tn_kernel<BN, EPI, CG>;
  if (!configured) {
    VPB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int m_tiles = (p.M + GEMM_BM * CG - 1) / (GEMM_BM * CG);
  const int n_tiles = (p.N + BN - 1) / BN;
  int grid = m_tiles * n_tiles * CG;
  int cap = max_ctas > 0 ? max_ctas : sm_count();
  cap -= cap % CG;
  if (cap < CG) cap = CG;
  if (grid > cap) grid = cap;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(gemm_threads(EPI));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VPB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, maps.a, maps.b, maps.out, maps.aux, p));
  return 0;
}

// CTA pairs pay off on the large transformer GEMMs (full 256-wide N tiles, many tiles); everything else stays 1-CTA.
// VPB_GEMM_CG=1 forces single-CTA tiles (A/B experiments).
int gemm_pick_cg(int M, int bn, int epilogue, int K) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("VPB_GEMM_CG");
    forced = e ? atoi(e) : 0;
  }
  if (forced == 1) return 1;
  if (!(bn == 256 && gemm_epi_staged(epilogue) && M >= 1024)) return 1;
  // measured (B200, M = 49152): qkv 1135 -> 1269, fc1 w/o GELU 1152 -> 1284, fc2 1080 -> 1218 TFLOP/s with pairs;
  // the short-K residual GEMM (attn.proj, K = D) is bound by its fp32 residual traffic and is ~3 % faster unpaired
  if (forced != 2 && epilogue == EPI_RESID_F32 && K < 1536) return 1;
  return 2;
}

int gemm_pick_bn(int N, int epilogue) {
  if (epilogue == EPI_NCHW_F32) return N <= 32 ? 32 : (N <= 144 ? 144 : 256);
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  if (N <= 64) return 64;
  return (N % 256) > 128 || N > 1024 ? 256 : 128;
}

int make_gemm_maps(GemmMaps* maps, const void* A, const void* B, int M, int N, int K, int lda, int ldb, int bn,
                   int epilogue, void* out, int ldo, const float* aux, int cg, int period) {
  uint64_t dims_a[2] = {(uint64_t)K, (uint64_t)M};
  uint64_t str_a[1] = {(uint64_t)lda * 2};
  uint32_t box_a[2] = {GEMM_BK, GEMM_BM};
  if (make_tma_desc(&maps->a, TMA_BF16, A, 2, dims_a, str_a, box_a, TMA_SWIZZLE_128B)) return -1;
  uint64_t dims_b[2] = {(uint64_t)K, (uint64_t)N};
  uint64_t str_b[1] = {(uint64_t)ldb * 2};
  uint32_t box_b[2] = {GEMM_BK, (uint32_t)(bn / cg)};   // each CTA of a pair loads half of the B tile
  if (make_tma_desc(&maps->b, TMA_BF16, B, 2, dims_b, str_b, box_b, TMA_SWIZZLE_128B)) return -1;
  maps->out = maps->a;   // placeholders for the epilogues that store directly
  maps->aux = maps->a;
  if (gemm_epi_staged(epilogue)) {
    const bool f32 = gemm_epi_adds_tile(epilogue);
    uint64_t dims_o[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t str_o[1] = {(uint64_t)ldo * (f32 ? 4 : 2)};
    uint32_t box_o[2] = {f32 ? 32u : 64u, GEMM_BM};
    if (make_tma_desc(&maps->out, f32 ? TMA_F32 : TMA_BF16, out, 2, dims_o, str_o, box_o, TMA_SWIZZLE_128B)) return -1;
    if (epilogue == EPI_RESID_F32 &&
        make_tma_desc(&maps->aux, TMA_F32, aux, 2, dims_o, str_o, box_o, TMA_SWIZZLE_128B))
      return -1;
    if (epilogue == EPI_POSTMA_F32) {   // positional table [period, N] fp32, fetched as 64-row boxes
      uint64_t dims_p[2] = {(uint64_t)N, (uint64_t)period};
      uint64_t str_p[1] = {(uint64_t)N * 4};
      uint32_t box_p[2] = {32u, 64u};
      if (make_tma_desc(&maps->aux, TMA_F32, aux, 2, dims_p, str_p, box_p, TMA_SWIZZLE_128B)) return -1;
    }
  }
  return 0;
}

int launch_gemm(const GemmMaps& maps, const GemmParams& p, int bn, int epilogue, int cg, int max_ctas,
                cudaStream_t stream) {
  if (cg == 2) {
    if (bn == 256 && epilogue == EPI_BIAS_BF16) return launch_gemm_inst<256, EPI_BIAS_BF16, 2>(maps, p, max_ctas, stream);
    if (bn == 256 && epilogue == EPI_GELU_BF16) return launch_gemm_inst<256, EPI_GELU_BF16, 2>(maps, p, max_ctas, stream);
    if (bn == 256 && epilogue == EPI_RESID_F32) return launch_gemm_inst<256, EPI_RESID_F32, 2>(maps, p, max_ctas, stream);
    if (bn == 256 && epilogue == EPI_POSTMA_F32) return launch_gemm_inst<256, EPI_POSTMA_F32, 2>(maps, p, max_ctas, stream);
    set_last_error("gemm: no CTA-pair kernel instance for BN=%d epilogue=%d", bn, epilogue);
    return -2;
  }
#define VPB_GEMM_CASE(BN_, EPI_) \
  if (bn == BN_ && epilogue == EPI_) return launch_gemm_inst<BN_, EPI_, 1>(maps, p, max_ctas, stream);
  VPB_GEMM_CASE(256, EPI_BIAS_BF16)
  VPB_GEMM_CASE(128, EPI_BIAS_BF16)
  VPB_GEMM_CASE(64, EPI_BIAS_BF16)
  VPB_GEMM_CASE(256, EPI_GELU_BF16)
  VPB_GEMM_CASE(128, EPI_GELU_BF16)
  VPB_GEMM_CASE(64, EPI_GELU_BF16)
  VPB_GEMM_CASE(256, EPI_RESID_F32)
  VPB_GEMM_CASE(128, EPI_RESID_F32)
  VPB_GEMM_CASE(64, EPI_RESID_F32)
  VPB_GEMM_CASE(256, EPI_POSTMA_F32)
  VPB_GEMM_CASE(128, EPI_POSTMA_F32)
  VPB_GEMM_CASE(64, EPI_POSTMA_F32)
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
  if (epilogue == EPI_RESID_F32)
    VPB_REQUIRE((reinterpret_cast<uintptr_t>(aux) & 15) == 0, "gemm: residual must be 16-byte aligned");
  if (epilogue == EPI_POS_F32 || epilogue == EPI_NCHW_F32) VPB_REQUIRE(period > 0, "gemm: period must be > 0");
  // patch embed + pos embed: stream the positional rows with TMA when a period is a whole number of 64-row boxes
  if (epilogue == EPI_POS_F32 && period % 64 == 0 && ldo == N && (reinterpret_cast<uintptr_t>(aux) & 15) == 0)
    epilogue = EPI_POSTMA_F32;
  const int bn = gemm_pick_bn(N, epilogue);
  const int cg = gemm_pick_cg(M, bn, epilogue, K);
  GemmMaps maps;
  if (make_gemm_maps(&maps, A, B, M, N, K, K, K, bn, epilogue, out, ldo, aux, cg, period)) return -1;
  GemmParams p{M, N, K, bias, out, ldo, aux, period};
  return launch_gemm(maps, p, bn, epilogue, cg, max_ctas, stream);
}

}  // namespace vpb
