#include <stdlib.h>

#include "gemm.cuh"
#include "host_util.h"
#include "ops.h"

namespace vpb {

static long long* gemm_debug_buf(cudaStream_t stream, int bn, int epi);
static bool gemm_cooperative();

template <int BN, int EPI, int CG, int OPM = 0>
static int launch_gemm_inst(const GemmMaps& maps, const GemmParams& p, int max_ctas, cudaStream_t stream) {
  constexpr int smem = gemm_smem_bytes(BN, EPI, CG);
  static bool configured = false;
  auto kern = gemm_bf16_tn_kernel<BN, EPI, CG, OPM>;
  if (!configured) {
    VPB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  const int m_tiles = (p.M + GEMM_BM * CG - 1) / (GEMM_BM * CG);
  const int n_tiles = (p.N + BN - 1) / BN;
  int grid = m_tiles * n_tiles * CG * p.ksplit;
  int cap = max_ctas > 0 ? max_ctas : sm_count();
  cap -= cap % CG;
  if (cap < CG) cap = CG;
  if (gemm_epi_ln(EPI)) {
    // the n-tiles of a row block exchange LayerNorm statistics: keep them on CTAs that run in the same step of the
    // persistent loop (grid a whole number of row blocks), and every CTA must be resident (spin-wait on siblings)
    VPB_REQUIRE(cap >= n_tiles * CG, "gemm+layernorm: needs at least %d resident CTAs", n_tiles * CG);
    // When that rounding would idle >= 5 % of the SMs (ViTPose-H: five column tiles x CTA pairs -> 140 of 148) keep every
    // SM busy instead: a row block whose tiles straddle the end of a step then has its early tiles wait up to one tile
    // period for the late ones (no cycle: statistics are published before the wait, and the cooperative launch keeps
    // every CTA resident). Measured (profiles/r02_summary.md §12): H proj + LN 219 -> 212 us, fc2 + LN 520 -> 509 us;
    // B (144 of 148 SMs) unchanged, so it keeps the rounded grid. VPB_LN_FULLGRID=0 / 1 forces either.
    static const int full_grid = [] { const char* e = getenv("VPB_LN_FULLGRID"); return e ? (atoi(e) != 0 ? 1 : 0) : -1; }();
    const int rem = cap % (n_tiles * CG);
    const bool keep_all = (full_grid == 1 || (full_grid == -1 && rem * 20 >= cap)) && p.ln_fold == 0 && gemm_cooperative();
    if (!keep_all) cap -= rem;
  }
  if (grid > cap) grid = cap;
  GemmParams pd = p;
  pd.dbg = gemm_debug_buf(stream, BN, EPI);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(gemm_threads(EPI));
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[3];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  if (gemm_epi_ln(EPI) && p.ln_fold == 0 && gemm_cooperative()) {
    // The CTAs that own the column tiles of one row block spin on each other's LayerNorm statistics: they must all
    // be resident. A cooperative launch makes the driver guarantee that (it places the whole grid at once, whatever
    // else runs on the device — a second stream, NCCL's CTAs, MPS) or fail the launch loudly, instead of relying on
    // the grid being sized to an otherwise idle GPU. VPB_COOP=0 launches normally (Nsight Compute cannot profile
    // cooperative cluster launches).
    attr[1].id = cudaLaunchAttributeCooperative;
    attr[1].val.cooperative = 1;
    cfg.numAttrs = 2;
    // VPB_COOP_PDL=1 (experiment): programmatic dependent launch on top of the cooperative launch — accepted by the
    // driver, results identical, no gain (21.43 / 21.44 ms per step against 21.39 / 21.18 without): stays off
    static const bool coop_pdl = [] { const char* e = getenv("VPB_COOP_PDL"); return e && atoi(e) != 0; }();
    if (coop_pdl) cfg.numAttrs += pdl_launch_attr(&attr[2]);
  } else {
    cfg.numAttrs = 1 + pdl_launch_attr(&attr[1]);
  }
  VPB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, maps.a, maps.b, maps.out, maps.aux, maps.ln, pd));
  return 0;
}

// Pipeline-depth switches of the residual GEMMs (GEMM_FLAG_*), VPB_GEMM_FLAGS=<int> overrides the default (A/B runs).
static int gemm_flags() {
  static int flags = -1;
  if (flags < 0) {
    const char* e = getenv("VPB_GEMM_FLAGS");
    flags = e ? atoi(e) : GEMM_DEFAULT_FLAGS;
  }
  return flags;
}

// VPB_GEMM_DEBUG=1: per-role wait-cycle counters of CTA 0 (managed memory); every launch prints the counters the
// previous launch left (tools/gemm_time.py). Measured with them in round 2 (profiles/r02_summary.md): the MMA issuer
// of the paired 256 x 256 GEMMs waits for operands 29 % (fc1) to 42 % (proj + LayerNorm, 3 stages) of the kernel.
static long long* gemm_debug_buf(cudaStream_t stream, int bn, int epi) {
  static int on = -1;
  static long long* buf = nullptr;
  if (on < 0) {
    const char* e = getenv("VPB_GEMM_DEBUG");
    on = e ? atoi(e) : 0;
  }
  if (!on) return nullptr;
  if (buf == nullptr) {
    if (cudaMallocManaged(&buf, 32 * sizeof(long long)) != cudaSuccess) return nullptr;
    for (int i = 0; i < 32; ++i) buf[i] = 0;
  } else {
    cudaStreamSynchronize(stream);
    fprintf(stderr, "gemm cycles (CTA0, prev launch -> now BN=%d epi=%d): producer wait_empty %lld total %lld | mma wait_tempty %lld "
                    "wait_full %lld total %lld | ring0 wait_ready %lld wait_read %lld total %lld | epi0 wait_tfull %lld "
                    "wait_rfull %lld sibling %lld wait_slot_p2 %lld total %lld tiles %lld | mma wait_full_B %lld\n",
            bn, epi, buf[0], buf[1], buf[2], buf[3], buf[4], buf[5], buf[6], buf[7], buf[8], buf[9], buf[10], buf[11],
            buf[12], buf[13], buf[14]);
  }
  return buf;
}

static bool gemm_cooperative() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("VPB_COOP");
    on = e ? atoi(e) : 1;
  }
  return on != 0;
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
  // 128-wide pair tiles (M = 256 x N = 128 per pair) exist for the fused-LayerNorm epilogues only: ViTPose-S, D = 384
  // = three 128-column tiles per row block (VPB_GEMM_PAIR128=0: single-CTA tiles as in round 1, A/B)
  static int pair128 = -1;
  if (pair128 < 0) {
    const char* e = getenv("VPB_GEMM_PAIR128");
    pair128 = (e && atoi(e) == 0) ? 0 : 1;
  }
  if ((bn == 128 || bn == 192) && pair128 && gemm_epi_ln(epilogue) && !gemm_epi_pos(epilogue) && M >= 1024) return 2;
  if (!(bn == 256 && gemm_epi_staged(epilogue) && M >= 1024)) return 1;
  // measured (B200, M = 49152): qkv 1135 -> 1269, fc1 w/o GELU 1152 -> 1284, fc2 1080 -> 1218 TFLOP/s with pairs;
  // the short-K residual GEMM (attn.proj, K = D) is bound by its fp32 residual traffic and is ~3 % faster unpaired
  if (forced != 2 && epilogue == EPI_RESID_F32 && K < 1536) return 1;
  return 2;
}

int gemm_pick_bn(int N, int epilogue) {
  if (epilogue == EPI_NCHW_F32) return N <= 32 ? 32 : (N <= 144 ? 144 : 256);
  if (N % 256 == 0) return 256;
  // wide but not a multiple of 256 (ViTPose-S qkv, N = 1152 = 4.5 tiles): a clipped last 256-wide tile on CTA pairs
  // beats nine 128-wide single-CTA tiles (VPB_GEMM_WIDE=0 restores those, A/B)
  static int wide = -1;
  if (wide < 0) {
    const char* e = getenv("VPB_GEMM_WIDE");
    wide = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (wide && N >= 1024 && (N % 256) >= 128 && gemm_epi_staged(epilogue) && !gemm_epi_adds_tile(epilogue)) return 256;
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
  maps->ln = maps->a;
  if (gemm_epi_staged(epilogue)) {
    const bool f32 = gemm_epi_adds_tile(epilogue);
    uint64_t dims_o[2] = {(uint64_t)N, (uint64_t)M};
    uint64_t str_o[1] = {(uint64_t)ldo * (f32 ? 4 : 2)};
    uint32_t box_o[2] = {f32 ? 32u : 64u, GEMM_BM};
    if (make_tma_desc(&maps->out, f32 ? TMA_F32 : TMA_BF16, out, 2, dims_o, str_o, box_o, TMA_SWIZZLE_128B)) return -1;
    if ((epilogue == EPI_RESID_F32 || epilogue == EPI_RESID_LN_F32 || epilogue == EPI_RESID_LNS_F32) &&
        make_tma_desc(&maps->aux, TMA_F32, aux, 2, dims_o, str_o, box_o, TMA_SWIZZLE_128B))
      return -1;
    if (gemm_epi_pos(epilogue)) {   // positional table [period, N] fp32, fetched as 64-row boxes
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
    if (bn == 256 && epilogue == EPI_BIAS_LNIN_BF16) return launch_gemm_inst<256, EPI_BIAS_LNIN_BF16, 2>(maps, p, max_ctas, stream);
    if (bn == 256 && epilogue == EPI_GELU_LNIN_BF16) return launch_gemm_inst<256, EPI_GELU_LNIN_BF16, 2>(maps, p, max_ctas, stream);
    if (bn == 256 && epilogue == EPI_GELU_SAVE_BF16) return launch_gemm_inst<256, EPI_GELU_SAVE_BF16, 2>(maps, p, max_ctas, stream);
    if (bn == 256 && epilogue == EPI_DGELU_BF16) return launch_gemm_inst<256, EPI_DGELU_BF16, 2>(maps, p, max_ctas, stream);
    if (bn == 256 && epilogue == EPI_RESID_F32) return launch_gemm_inst<256, EPI_RESID_F32, 2>(maps, p, max_ctas, stream);
    if (bn == 256 && epilogue == EPI_POSTMA_F32) return launch_gemm_inst<256, EPI_POSTMA_F32, 2>(maps, p, max_ctas, stream);
    if (bn == 256 && epilogue == EPI_RESID_LN_F32) return launch_gemm_inst<256, EPI_RESID_LN_F32, 2>(maps, p, max_ctas, stream);
    if (bn == 256 && epilogue == EPI_POSTMA_LN_F32) return launch_gemm_inst<256, EPI_POSTMA_LN_F32, 2>(maps, p, max_ctas, stream);
    if (bn == 256 && epilogue == EPI_RESID_LNS_F32) return launch_gemm_inst<256, EPI_RESID_LNS_F32, 2>(maps, p, max_ctas, stream);
    if (bn == 256 && epilogue == EPI_POSTMA_LNS_F32) return launch_gemm_inst<256, EPI_POSTMA_LNS_F32, 2>(maps, p, max_ctas, stream);
    if (bn == 192 && epilogue == EPI_RESID_LN_F32) return launch_gemm_inst<192, EPI_RESID_LN_F32, 2>(maps, p, max_ctas, stream);
    if (bn == 192 && epilogue == EPI_RESID_LNS_F32) return launch_gemm_inst<192, EPI_RESID_LNS_F32, 2>(maps, p, max_ctas, stream);
    if (bn == 128 && epilogue == EPI_RESID_LN_F32) return launch_gemm_inst<128, EPI_RESID_LN_F32, 2>(maps, p, max_ctas, stream);
    if (bn == 128 && epilogue == EPI_RESID_LNS_F32) return launch_gemm_inst<128, EPI_RESID_LNS_F32, 2>(maps, p, max_ctas, stream);
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
  VPB_GEMM_CASE(256, EPI_BIAS_LNIN_BF16)
  VPB_GEMM_CASE(128, EPI_BIAS_LNIN_BF16)
  VPB_GEMM_CASE(64, EPI_BIAS_LNIN_BF16)
  VPB_GEMM_CASE(256, EPI_GELU_LNIN_BF16)
  VPB_GEMM_CASE(128, EPI_GELU_LNIN_BF16)
  VPB_GEMM_CASE(64, EPI_GELU_LNIN_BF16)
  VPB_GEMM_CASE(256, EPI_GELU_SAVE_BF16)
  VPB_GEMM_CASE(128, EPI_GELU_SAVE_BF16)
  VPB_GEMM_CASE(64, EPI_GELU_SAVE_BF16)
  VPB_GEMM_CASE(256, EPI_DGELU_BF16)
  VPB_GEMM_CASE(128, EPI_DGELU_BF16)
  VPB_GEMM_CASE(64, EPI_DGELU_BF16)
  VPB_GEMM_CASE(256, EPI_RESID_F32)
  VPB_GEMM_CASE(128, EPI_RESID_F32)
  VPB_GEMM_CASE(64, EPI_RESID_F32)
  VPB_GEMM_CASE(256, EPI_POSTMA_F32)
  VPB_GEMM_CASE(128, EPI_POSTMA_F32)
  VPB_GEMM_CASE(64, EPI_POSTMA_F32)
  VPB_GEMM_CASE(256, EPI_RESID_LN_F32)
  VPB_GEMM_CASE(192, EPI_RESID_LN_F32)
  VPB_GEMM_CASE(128, EPI_RESID_LN_F32)
  VPB_GEMM_CASE(64, EPI_RESID_LN_F32)
  VPB_GEMM_CASE(256, EPI_POSTMA_LN_F32)
  VPB_GEMM_CASE(192, EPI_POSTMA_LN_F32)
  VPB_GEMM_CASE(128, EPI_POSTMA_LN_F32)
  VPB_GEMM_CASE(64, EPI_POSTMA_LN_F32)
  VPB_GEMM_CASE(256, EPI_RESID_LNS_F32)
  VPB_GEMM_CASE(192, EPI_RESID_LNS_F32)
  VPB_GEMM_CASE(128, EPI_RESID_LNS_F32)
  VPB_GEMM_CASE(64, EPI_RESID_LNS_F32)
  VPB_GEMM_CASE(256, EPI_POSTMA_LNS_F32)
  VPB_GEMM_CASE(192, EPI_POSTMA_LNS_F32)
  VPB_GEMM_CASE(128, EPI_POSTMA_LNS_F32)
  VPB_GEMM_CASE(64, EPI_POSTMA_LNS_F32)
  VPB_GEMM_CASE(256, EPI_ACCUM_F32)
  VPB_GEMM_CASE(128, EPI_ACCUM_F32)
  VPB_GEMM_CASE(64, EPI_ACCUM_F32)
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

// Split of the contraction for the accumulating epilogue (weight gradients: few output tiles, long K). Every work item
// (tile, split) goes to one CTA of a persistent grid of `ctas`, so the cost is waves x (K blocks per item + the item's
// fixed part: pipeline fill and the fp32 atomic epilogue, about 8 K blocks' worth). Picks the split count that
// minimises it — e.g. fc1's gradient at 64 crops (72 tiles, 192 K blocks): 2 splits = one wave of 144 items instead of
// the 5 splits = 2.4 -> 3 waves that "about two waves" used to give; at least 8 K blocks per split.
static int pick_ksplit(int tiles, int k_blocks, int ctas) {
  int best = 1;
  long best_cost = -1;
  for (int s = 1; s <= 32 && s * 8 <= k_blocks; ++s) {
    const int kb_per = (k_blocks + s - 1) / s;
    const int splits = (k_blocks + kb_per - 1) / kb_per;
    if (splits != s) continue;                            // (same item size as a smaller s)
    const long waves = (static_cast<long>(tiles) * splits + ctas - 1) / ctas;
    const long cost = waves * (kb_per + 8);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

int gemm_bf16(const void* A, const void* B, int M, int N, int K, int epilogue, const float* bias, void* out, int ldo,
              const float* aux, int period, int max_ctas, cudaStream_t stream, const LnFoldIn* ln,
              const GemmTrainAux* tr) {
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
  GemmParams p{M, N, K, bias, out, ldo, aux, period, 1, nullptr, 1, nullptr, nullptr, nullptr, 0, 0u, 0.0f};
  p.flags = gemm_flags();
  if (ln != nullptr && ln->stats != nullptr) {     // A = plain bf16 rows of a residual stream, LayerNorm applied here
    VPB_REQUIRE(epilogue == EPI_BIAS_BF16 || epilogue == EPI_GELU_BF16, "gemm: folded LayerNorm needs a bf16 epilogue");
    VPB_REQUIRE(ln->s != nullptr && bias != nullptr && ln->parts >= 1 && ln->parts <= 10 && ln->part_cols > 0 &&
                    ln->parts * ln->part_cols == K,
                "gemm: folded LayerNorm: bad statistics layout (parts %d x %d columns, K = %d)", ln->parts, ln->part_cols, K);
    p.ln_stats = reinterpret_cast<const float2*>(ln->stats);
    p.ln_s = ln->s;
    p.ln_parts = ln->parts;
    p.ln_part_cols = ln->part_cols;
    p.ln_eps = ln->eps;
    epilogue = epilogue == EPI_BIAS_BF16 ? EPI_BIAS_LNIN_BF16 : EPI_GELU_LNIN_BF16;    // (same tiles and tensor maps)
  }
  if (epilogue == EPI_DGELU_BF16)
    VPB_REQUIRE(tr != nullptr && tr->pre_in != nullptr && bias == nullptr && N % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(tr->pre_in) & 15) == 0,
                "gemm: the gelu-backward epilogue needs the saved pre-activation [M, N] (16-byte aligned, N %% 8 == 0) and no bias");
  if (tr != nullptr) {
    if (tr->pre_out != nullptr && epilogue == EPI_GELU_BF16) epilogue = EPI_GELU_SAVE_BF16;
    VPB_REQUIRE(tr->pre_out == nullptr || (epilogue == EPI_GELU_SAVE_BF16 && N % 8 == 0 &&
                                           (reinterpret_cast<uintptr_t>(tr->pre_out) & 15) == 0),
                "gemm: a pre-activation output needs the GELU epilogue, N %% 8 == 0 and a 16-byte aligned buffer");
    VPB_REQUIRE(tr->colsum_out == nullptr || epilogue == EPI_DGELU_BF16, "gemm: column sums come with the gelu-backward epilogue");
    if (tr->pre_out != nullptr || tr->pre_in != nullptr) {     // [M, N] bf16, same boxes as the output
      uint64_t dims_o[2] = {(uint64_t)N, (uint64_t)M};
      uint64_t str_o[1] = {(uint64_t)N * 2};
      uint32_t box_o[2] = {64u, GEMM_BM};
      void* pre = tr->pre_out != nullptr ? tr->pre_out : const_cast<void*>(tr->pre_in);
      if (make_tma_desc(&maps.aux, TMA_BF16, pre, 2, dims_o, str_o, box_o, TMA_SWIZZLE_128B)) return -1;
    }
    p.pre_out = tr->pre_out;
    p.pre_in = tr->pre_in;
    p.colsum_out = tr->colsum_out;
  }
  if (epilogue == EPI_ACCUM_F32) {
    VPB_REQUIRE(bias == nullptr && ldo % 4 == 0, "gemm: the accumulating epilogue takes no bias and needs ldo %% 4 == 0");
    const int tiles = ((M + GEMM_BM - 1) / GEMM_BM) * ((N + bn - 1) / bn);
    p.ksplit = pick_ksplit(tiles, (K + GEMM_BK - 1) / GEMM_BK, max_ctas > 0 ? max_ctas : sm_count());
  }
  return launch_gemm(maps, p, bn, epilogue, cg, max_ctas, stream);
}

// ---- residual GEMM + fused LayerNorm ------------------------------------------------------------------------
static int ln_row_blocks(int M) { return (M + 255) / 256 * 2; }          // 128-row blocks, padded to CTA pairs
// column tile of the fused-LayerNorm GEMMs: the widest of 256 / 192 / 128 / 64 that divides N (ViTPose-S, N = 384: two
// 192-wide tiles — 22 % fewer operand bytes per FLOP than three 128-wide ones; VPB_LN_BN192=0: the round-1 choice)
static int ln_bn(int N) {
  static int use192 = -1;
  if (use192 < 0) {
    const char* e = getenv("VPB_LN_BN192");
    use192 = (e && atoi(e) == 0) ? 0 : 1;
  }
  if (N % 256 == 0) return 256;
  if (use192 && N % 192 == 0) return 192;
  return N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 0);
}
static size_t ln_region_words(int M, int N) {     // one 8-byte word per (row, n-tile), region padded to 256 bytes
  const int bn = ln_bn(N) >= 128 ? 128 : 64;      // room for the narrowest tiles a variant may pick
  const size_t n_tiles = (N + bn - 1) / bn;
  return (static_cast<size_t>(ln_row_blocks(M)) * 128 * n_tiles + 31) / 32 * 32;
}
size_t gemm_ln_scratch_bytes(int M, int N) { return 2 * ln_region_words(M, N) * 8; }
// folded LayerNorm: column tiles per row the producer writes statistics for, their width, and the buffer size
int gemm_ln_parts(int N) { return ln_bn(N) ? N / ln_bn(N) : 0; }
int gemm_ln_part_cols(int N) { return ln_bn(N); }
size_t gemm_ln_stats_bytes(int M, int N) { return static_cast<size_t>(ln_row_blocks(M)) * 128 * gemm_ln_parts(N) * 8; }
// Launch e uses region e & 1 and expects tag (e >> 1) & 1 in the words its siblings write: region 1 (first used by
// e = 1, tag 0) starts with all tag bits set, region 0 (first used by e = 2, tag 1) with all tag bits clear.
int gemm_ln_scratch_init(void* scratch, int M, int N, cudaStream_t stream) {
  const size_t bytes = ln_region_words(M, N) * 8;
  VPB_CHECK_CUDA(cudaMemsetAsync(scratch, 0x00, bytes, stream));
  VPB_CHECK_CUDA(cudaMemsetAsync(static_cast<uint8_t*>(scratch) + bytes, 0xFF, bytes, stream));
  return 0;
}

int gemm_bf16_ln(const void* A, const void* B, int M, int N, int K, int epilogue, const float* bias, float* out,
                 const float* aux, int period, const float* gamma, const float* beta, float eps, void* xn,
                 void* scratch, unsigned epoch, int max_ctas, cudaStream_t stream, const float* row_scale,
                 int rows_per_scale, void* fold_stats) {
  VPB_REQUIRE(row_scale == nullptr || (rows_per_scale > 0 && epilogue == EPI_RESID_F32),
              "gemm+layernorm: row_scale needs rows_per_scale > 0 and the residual epilogue");
  VPB_REQUIRE(epilogue == EPI_RESID_F32 || epilogue == EPI_POS_F32, "gemm+layernorm: epilogue %d has no fused form",
              epilogue);
  const bool fold = fold_stats != nullptr;   // xn = plain bf16 copy of the rows + per-tile (mean, M2): see gemm.cuh
  VPB_REQUIRE((fold || (gamma && beta)) && xn && out && aux, "gemm+layernorm: null argument");
  const int bn = ln_bn(N);
  static int disabled = -1;   // VPB_LN_FUSED=0: GEMM + separate LayerNorm kernel (A/B measurements)
  if (disabled < 0) {
    const char* e = getenv("VPB_LN_FUSED");
    disabled = (e && atoi(e) == 0) ? 1 : 0;
  }
  const bool pos_ok = epilogue != EPI_POS_F32 || (period > 0 && period % 64 == 0);
  const bool fused = (fold || (!disabled && scratch != nullptr && epoch > 0)) && bn != 0 && pos_ok && K % 8 == 0 &&
                     N % 8 == 0 &&
                     ((reinterpret_cast<uintptr_t>(aux) | reinterpret_cast<uintptr_t>(out) |
                       reinterpret_cast<uintptr_t>(xn) | reinterpret_cast<uintptr_t>(fold ? fold_stats : scratch)) & 15) == 0;
  if (fold && !fused) {
    set_last_error("gemm + folded layernorm: unsupported shape N=%d K=%d period=%d", N, K, period);
    return -2;
  }
  if (!fused && row_scale != nullptr) {
    set_last_error("gemm+layernorm: row_scale is only implemented in the fused kernel (N=%d K=%d)", N, K);
    return -2;
  }
  if (!fused) {   // two kernels (still on the GPU): shapes the fused epilogue does not cover
    if (int e = gemm_bf16(A, B, M, N, K, epilogue, bias, out, N, aux, period, max_ctas, stream)) return e;
    return layernorm_bf16(out, gamma, beta, xn, M, N, eps, stream);
  }
  VPB_REQUIRE(M > 0 && K > 0, "gemm+layernorm: empty problem");
  VPB_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
              "gemm: operands must be 16-byte aligned");
  // short K: the epilogue sets the pace -> two epilogue warpgroups with their own residual rings on alternating
  // tiles (needs the smaller operand stages of CTA pairs or of <=128-wide tiles to fit in shared memory).
  // measured (B200, M = 49152, K = N = 768): 0.117 ms, against 0.138 ms with one epilogue group (pairs or single
  // CTAs) and 0.153 ms with 128-wide tiles; 0.076 + 0.038 ms for the unfused GEMM + LayerNorm kernels
  // (the split variant's 3 operand stages cost a long-K GEMM more than the second epilogue group gains: fc2 + LN
  // 0.211 ms unsplit, 0.240 ms split at M = 49152)
  const bool short_k = K < 1536;
  int cg = gemm_pick_cg(M, bn, EPI_RESID_LNS_F32, K);
  if (bn <= 192 && epilogue != EPI_RESID_F32) cg = 1;     // (128 / 192-wide pair tiles: residual epilogues only)
  const bool split = !fold && short_k && (bn <= 192 || cg == 2);   // (the folded form: one-group kernels only)
  const int epi = epilogue == EPI_RESID_F32 ? (split ? EPI_RESID_LNS_F32 : EPI_RESID_LN_F32)
                                            : (split ? EPI_POSTMA_LNS_F32 : EPI_POSTMA_LN_F32);
  if (!split && short_k && !fold) cg = 1;     // short-K residual GEMMs are slightly faster unpaired (see gemm_pick_cg)
  GemmMaps maps;
  if (make_gemm_maps(&maps, A, B, M, N, K, K, K, bn, epi, out, N, aux, cg, period)) return -1;
  uint64_t dims_o[2] = {(uint64_t)N, (uint64_t)M};
  uint64_t str_o[1] = {(uint64_t)N * 2};
  uint32_t box_o[2] = {64u, GEMM_BM};
  if (make_tma_desc(&maps.ln, TMA_BF16, xn, 2, dims_o, str_o, box_o, TMA_SWIZZLE_128B)) return -1;
  VPB_REQUIRE(N / bn <= 10, "gemm+layernorm: at most 10 column tiles per row (N=%d)", N);
  GemmParams p{M, N, K, bias, out, N, aux, period, 1, row_scale, rows_per_scale > 0 ? rows_per_scale : 1, gamma, beta,
               reinterpret_cast<unsigned long long*>(scratch),
               ln_region_words(M, N), epoch, eps};
  p.flags = gemm_flags();
  if (fold) {
    p.ln_fold = 1;
    p.ln_stats_out = reinterpret_cast<float2*>(fold_stats);
  }
  return launch_gemm(maps, p, bn, epi, cg, max_ctas, stream);
}

// ---- out[M, N] += At^T . Bt for row-major At [K, M], Bt [K, N] (weight gradients) ------------------------------
int gemm_bf16_atb_accum(const void* At, const void* Bt, int M, int N, int K, float* out, int ldo, int max_ctas,
                        cudaStream_t stream) {
  return gemm_bf16_atb_accum_ld(At, M, Bt, N, M, N, K, out, ldo, max_ctas, stream);
}

// the same with explicit row pitches (elements): At / Bt may be column slices of wider row-major matrices — the
// gradient of a Linear layer whose output columns are a slice of dY (ViTPose+ experts, vit_moe.py:107-111)
int gemm_bf16_atb_accum_ld(const void* At, int lda, const void* Bt, int ldb, int M, int N, int K, float* out, int ldo,
                           int max_ctas, cudaStream_t stream) {
  VPB_REQUIRE(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  VPB_REQUIRE(M % 8 == 0 && N % 8 == 0 && lda >= M && ldb >= N && lda % 8 == 0 && ldb % 8 == 0,
              "gemm(At, Bt): M=%d, N=%d and the row pitches %d / %d must be multiples of 8 (16-byte row pitch)", M, N,
              lda, ldb);
  VPB_REQUIRE(((reinterpret_cast<uintptr_t>(At) | reinterpret_cast<uintptr_t>(Bt) | reinterpret_cast<uintptr_t>(out)) & 15) == 0 &&
                  ldo % 4 == 0, "gemm(At, Bt): operands / out must be 16-byte aligned, ldo %% 4 == 0");
  const int bn = N > 128 ? 256 : (N > 64 ? 128 : 64);
  GemmMaps maps;
  uint64_t dims_a[2] = {(uint64_t)M, (uint64_t)K};
  uint64_t str_a[1] = {(uint64_t)lda * 2};
  uint32_t box[2] = {64u, (uint32_t)GEMM_BK};
  if (make_tma_desc(&maps.a, TMA_BF16, At, 2, dims_a, str_a, box, TMA_SWIZZLE_128B)) return -1;
  uint64_t dims_b[2] = {(uint64_t)N, (uint64_t)K};
  uint64_t str_b[1] = {(uint64_t)ldb * 2};
  if (make_tma_desc(&maps.b, TMA_BF16, Bt, 2, dims_b, str_b, box, TMA_SWIZZLE_128B)) return -1;
  maps.out = maps.a;
  maps.aux = maps.a;
  maps.ln = maps.a;
  GemmParams p{M, N, K, nullptr, out, ldo, nullptr, 0, 1, nullptr, 1, nullptr, nullptr, nullptr, 0, 0u, 0.0f};
  // CTA pairs (256 x 256 tiles, each CTA stages its 128 A columns and half of the B chunk) when the output has whole
  // pair tiles: the Linear layers of the backbone (VPB_WGRAD_PAIR=0: single-CTA tiles, A/B)
  static const bool pair_ok = [] { const char* e = getenv("VPB_WGRAD_PAIR"); return !(e && atoi(e) == 0); }();
  const int cg = (pair_ok && bn == 256 && M % 256 == 0 && N % 256 == 0) ? 2 : 1;
  const int tiles = ((M + GEMM_BM * cg - 1) / (GEMM_BM * cg)) * ((N + bn - 1) / bn);
  p.ksplit = pick_ksplit(tiles, (K + GEMM_BK - 1) / GEMM_BK, (max_ctas > 0 ? max_ctas : sm_count()) / cg);
  if (cg == 2) return launch_gemm_inst<256, EPI_ACCUM_F32, 2, 1>(maps, p, max_ctas, stream);
  if (bn == 256) return launch_gemm_inst<256, EPI_ACCUM_F32, 1, 1>(maps, p, max_ctas, stream);
  if (bn == 128) return launch_gemm_inst<128, EPI_ACCUM_F32, 1, 1>(maps, p, max_ctas, stream);
  return launch_gemm_inst<64, EPI_ACCUM_F32, 1, 1>(maps, p, max_ctas, stream);
}

}  // namespace vpb
