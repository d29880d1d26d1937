// Implicit-GEMM convolutions of the keypoint head on tcgen05 tensor cores (NHWC bf16 activations).
//
//   MODE_DECONV : ConvTranspose2d(Cin, Cout, 4, stride 2, padding 1, bias=False) + BatchNorm2d(eval) + ReLU
//                 (topdown_heatmap_simple_head.py:306-337, _get_deconv_cfg topdown_heatmap_base_head.py:105-120).
//                 A k4/s2/p1 transposed conv is four independent 2x2 convolutions, one per output parity
//                 (py, px): out[2i+py, 2j+px] = sum over taps (dy, dx) of in[i+dy, j+dx] . W[:, :, kh, kw] with
//                 py=0: (dy,kh) in {(0,1), (-1,3)},  py=1: (dy,kh) in {(+1,0), (0,2)}  (same for x).
//                 BN is folded to a per-channel scale/shift applied in the epilogue.
//   MODE_CONV3  : Conv2d(Cin, K, 3, padding 1) + bias of the simple decoder (:132-139), fp32 NCHW heatmaps out.
//
// The A operand is never materialised: each K step is one TMA box {64 channels, w_box, h_box, n_box} of the
// NHWC input fetched at the tap's (dx, dy) offset — out-of-image coordinates are zero-filled by TMA, which is
// exactly the conv padding. A box holds 384 pixels = three 128-row UMMA tiles that share one weight tile, so
// every weight byte staged in shared memory feeds 3x the math.
#include <cstdlib>
#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

namespace vpb {

// warp 0: TMA producer, warp 1: MMA issuer, warps 2-9: two epilogue warpgroups. The accumulators (384 of the 512 TMEM
// columns) are single-buffered, so the epilogue of an item is NOT overlapped with the next item's MMAs: the two groups
// each drain half of the channels of every pixel tile (round 2; one group of four warps before).
#ifndef VPB_CV_EPI_GROUPS
#define VPB_CV_EPI_GROUPS 2
#endif
constexpr int CV_EPI_GROUPS = VPB_CV_EPI_GROUPS;
constexpr int CV_THREADS = 64 + 128 * CV_EPI_GROUPS;
constexpr int CV_SUB = 3;                 // 128-row sub-tiles per super tile
constexpr int CV_ROWS = CV_SUB * 128;     // 384 pixels
constexpr int MODE_DECONV = 0;
constexpr int MODE_CONV3 = 1;

struct ConvParams {
  int n, h, w, cin, cout;
  int w_box, h_box, n_box;       // pixels of an item = w_box * h_box * n_box (w_box divides w)
  int tiles_x, tiles_y;          // w / w_box, h / h_box
  int super_tiles;               // ceil(n / n_box) * tiles_y * tiles_x
  int n_tiles;                   // ceil(cout / BN)
  const float* scale;            // [cout] (deconv: folded BN scale)
  const float* shift;            // [cout] (deconv: folded BN shift; conv3: bias)
  float floor;                   // deconv: 0 = ReLU, -inf = no activation (training: BatchNorm needs the raw output)
  void* out;
};

// SUB = 128-row sub-tiles per item. <128 channels, 3 sub-tiles> moves (48 + 16) KB of operands per 768 MMA cycles;
// <256 channels, 2 sub-tiles> (the whole Cout = 256 of the ViTPose decoders in one tile: the activation box is
// fetched once instead of once per 128-channel tile) moves (32 + 32) KB per 1024 cycles: 25 % fewer L2 -> SM bytes per
// MMA cycle, which is what bounds this kernel. 256 pixels are a strip w_box x h_box x n_box of the image (w_box < w).
template <int BN, int MODE, int SUB = CV_SUB>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_w,
                  const ConvParams p) {
  constexpr int A_BYTES = SUB * 128 * 128;
  constexpr int B_BYTES = BN * 128;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGES = (196608 / STAGE_BYTES) > 6 ? 6 : (196608 / STAGE_BYTES);
  constexpr int TMEM_COLS = SUB * BN <= 128 ? 128 : (SUB * BN <= 256 ? 256 : 512);
  constexpr uint32_t IDESC = umma_idesc_bf16(128, BN);
  constexpr int NTAPS = MODE == MODE_DECONV ? 4 : 9;
  constexpr int NPHASE = MODE == MODE_DECONV ? 4 : 1;
  static_assert(SUB * BN <= 512, "accumulators must fit TMEM");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar[STAGES];
  __shared__ uint64_t empty_bar[STAGES];
  __shared__ uint64_t tfull_bar, tempty_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float s_ss[2][2][BN];                 // [item parity][scale | shift][channel of the n-tile]
  __shared__ __align__(16) uint8_t s_ostage[4 * CV_EPI_GROUPS][32 * 64];   // per epilogue warp: 32 pixels x 64 bytes

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cin_chunks = p.cin / 64;
  const int k_steps = NTAPS * cin_chunks;
  const int num_items = p.super_tiles * NPHASE * p.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar, 1);
    mbar_init(&tempty_bar, 4 * CV_EPI_GROUPS);
    fence_mbar_init();
    tma_prefetch_desc(&tm_in);
    tma_prefetch_desc(&tm_w);
  }
  if (warp == 1) tmem_alloc(&tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  pdl_wait();                  // the set-up above overlapped the previous kernel's tail
  pdl_launch_dependents();

  // item -> (super tile, phase, n tile); n tile fastest so co-running CTAs share the activation box in L2
  auto decode_item = [&](int item, int& st, int& phase, int& nt) {
    nt = item % p.n_tiles;
    const int rest = item / p.n_tiles;
    phase = rest % NPHASE;
    st = rest / NPHASE;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t ph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int st, phase, nt;
        decode_item(item, st, phase, nt);
        const int x0 = (st % p.tiles_x) * p.w_box;
        const int n0 = (st / p.tiles_x / p.tiles_y) * p.n_box;
        const int y0 = (st / p.tiles_x % p.tiles_y) * p.h_box;
        const int py = phase >> 1, px = phase & 1;
        for (int ks = 0; ks < k_steps; ++ks) {
          const int tap = ks / cin_chunks, cc = ks - tap * cin_chunks;
          int dy, dx;
          if (MODE == MODE_DECONV) {
            const int ty = tap >> 1, tx = tap & 1;        // tap 0 -> the dy=0 (resp. dx=0) neighbour
            dy = ty == 0 ? 0 : (py == 0 ? -1 : 1);
            dx = tx == 0 ? 0 : (px == 0 ? -1 : 1);
          } else {
            dy = tap / 3 - 1;
            dx = tap % 3 - 1;
          }
          mbar_wait(&empty_bar[stage], ph ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_4d(sa, &tm_in, &full_bar[stage], cc * 64, x0 + dx, y0 + dy, n0);
          tma_load_2d(sa + A_BYTES, &tm_w, &full_bar[stage], tap * p.cin + cc * 64, phase * p.cout + nt * BN);
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0;
      uint32_t ph = 0, acc_ph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        mbar_wait(&tempty_bar, acc_ph ^ 1);
        tc_fence_after();
        for (int ks = 0; ks < k_steps; ++ks) {
          mbar_wait(&full_bar[stage], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int sub = 0; sub < SUB; ++sub) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_bf16_ss(tmem_base + sub * BN, umma_desc_k_sw128(a_addr + sub * (128 * 128) + k * 32),
                           umma_desc_k_sw128(b_addr + k * 32), IDESC, (ks | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar);
        acc_ph ^= 1;
      }
    }
  } else {
    const int quad = warp & 3;
    const int etid = threadIdx.x - 64;
    const int egrp = (warp - 2) >> 2;            // epilogue group: channel chunks egrp, egrp + CV_EPI_GROUPS, ...
    uint32_t acc_ph = 0;
    uint32_t it = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      int st, phase, nt;
      decode_item(item, st, phase, nt);
      const int x0 = (st % p.tiles_x) * p.w_box;
      const int n0 = (st / p.tiles_x / p.tiles_y) * p.n_box;
      const int y0 = (st / p.tiles_x % p.tiles_y) * p.h_box;
      const int py = phase >> 1, px = phase & 1;
      // per-channel scale / shift (folded BN) or bias of this n-tile -> smem, double-buffered by item parity
      float* sc = s_ss[it & 1][0];
      float* sh = s_ss[it & 1][1];
      for (int i = etid; i < BN; i += 128 * CV_EPI_GROUPS) {
        const int co = nt * BN + i;
        sc[i] = (p.scale != nullptr && co < p.cout) ? __ldg(p.scale + co) : 1.0f;
        sh[i] = (p.shift != nullptr && co < p.cout) ? __ldg(p.shift + co) : 0.0f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(128 * CV_EPI_GROUPS) : "memory");
      mbar_wait(&tfull_bar, acc_ph);
      tc_fence_after();
#pragma unroll 1
      for (int sub = 0; sub < SUB; ++sub) {
        const int pr = sub * 128 + quad * 32 + lane;            // pixel row inside the box (x fastest)
        const int xx = x0 + pr % p.w_box;
        const int yy = (pr / p.w_box) % p.h_box;
        const int nn = pr / (p.w_box * p.h_box);
        const int img = n0 + nn, iy = y0 + yy;
        const bool ok = img < p.n && iy < p.h;
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + sub * BN;
        if constexpr (MODE == MODE_DECONV && BN % 32 == 0) {
          // A thread owns a pixel, but pixel-per-thread stores put 16 bytes into 32 different lines per instruction.
          // Each warp stages its 32 pixels x 32 channels in shared memory (16-byte pieces XOR-swizzled, conflict-free
          // both ways) and writes them back with four lanes per pixel: 64 contiguous bytes, whole sectors.
          const int oy = 2 * iy + py, ox = 2 * xx + px;
          const long long my_off =
              ok ? ((static_cast<long long>(img) * (2 * p.h) + oy) * (2 * p.w) + ox) * p.cout + nt * BN : -1;
          long long poff[4];                        // output offsets of the pixels this lane writes back
#pragma unroll
          for (int j = 0; j < 4; ++j) poff[j] = __shfl_sync(0xffffffffu, my_off, 8 * j + (lane >> 2));
          uint8_t* stage = s_ostage[warp - 2];
          __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(p.out);
#pragma unroll 1
          for (int c = 32 * egrp; c < BN; c += 32 * CV_EPI_GROUPS) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_row + c, r);
            tmem_ld_wait();
            if (nt * BN + c >= p.cout) continue;     // warp-uniform
            if (nt * BN + c + 32 <= p.cout) {        // 64 contiguous bytes per pixel (warp-uniform)
              uint32_t w[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float a = fmaxf(fmaf(__uint_as_float(r[2 * j]), sc[c + 2 * j], sh[c + 2 * j]), p.floor);
                const float b =
                    fmaxf(fmaf(__uint_as_float(r[2 * j + 1]), sc[c + 2 * j + 1], sh[c + 2 * j + 1]), p.floor);
                w[j] = pack_bf16x2(a, b);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u)
                *reinterpret_cast<uint4*>(stage + lane * 64 + ((u ^ ((lane >> 1) & 3)) * 16)) =
                    make_uint4(w[4 * u], w[4 * u + 1], w[4 * u + 2], w[4 * u + 3]);
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int row = 8 * j + (lane >> 2), ch = lane & 3;
                if (poff[j] >= 0)
                  *reinterpret_cast<uint4*>(obase + poff[j] + c + ch * 8) =
                      *reinterpret_cast<const uint4*>(stage + row * 64 + ((ch ^ ((row >> 1) & 3)) * 16));
              }
              __syncwarp();
            } else if (ok) {
              __nv_bfloat16* orow = obase + my_off;
              for (int j = 0; j < p.cout - (nt * BN + c); ++j)
                orow[c + j] = __float2bfloat16_rn(fmaxf(fmaf(__uint_as_float(r[j]), sc[c + j], sh[c + j]), p.floor));
            }
          }
        } else {
#pragma unroll 1
          for (int c = egrp; c < BN / 16; c += CV_EPI_GROUPS) {
            uint32_t r[16];
            tmem_ld_32x32b_x16(t_row + c * 16, r);
            tmem_ld_wait();
            const int co0 = nt * BN + c * 16;
            if (!ok || co0 >= p.cout) continue;
            const int ncols = min(16, p.cout - co0);
            if (MODE == MODE_DECONV) {
              const int oy = 2 * iy + py, ox = 2 * xx + px;
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) +
                                 ((static_cast<size_t>(img) * (2 * p.h) + oy) * (2 * p.w) + ox) * p.cout + co0;
              for (int j = 0; j < ncols; ++j)
                o[j] = __float2bfloat16_rn(
                    fmaxf(fmaf(__uint_as_float(r[j]), sc[c * 16 + j], sh[c * 16 + j]), p.floor));
            } else {
              // fp32 NCHW heatmaps: lanes hold consecutive pixels -> 128-byte coalesced stores per channel
              float* o = reinterpret_cast<float*>(p.out) +
                         ((static_cast<size_t>(img) * p.cout + co0) * p.h + iy) * p.w + xx;
              const size_t plane = static_cast<size_t>(p.h) * p.w;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < ncols) o[j * plane] = __uint_as_float(r[j]) + sh[c * 16 + j];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar);
      acc_ph ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

static int pick_boxes(int n, int h, int w, ConvParams& p, int box_rows = CV_ROWS) {
  // the widest strip of image columns whose pixels fill whole image rows of the box: w_box divides w and box_rows
  p.w_box = 0;
  for (int wb = w; wb >= 1; --wb)
    if (w % wb == 0 && box_rows % wb == 0 && wb <= 256) {
      const int rows = box_rows / wb;             // image rows per item if an item stays inside one image
      if ((rows <= h && h % rows == 0) || rows % h == 0) { p.w_box = wb; break; }
    }
  if (p.w_box == 0) return -1;
  const int rows = box_rows / p.w_box;
  if (rows <= h) {
    p.h_box = rows; p.n_box = 1;
  } else {
    p.h_box = h; p.n_box = rows / h;
  }
  if (p.h_box > 256 || p.n_box > 256) return -1;
  p.tiles_x = w / p.w_box;
  p.tiles_y = h / p.h_box;
  p.super_tiles = ((n + p.n_box - 1) / p.n_box) * p.tiles_y * p.tiles_x;
  return 0;
}

template <int BN, int MODE, int SUB = CV_SUB>
static int launch_conv(const void* in, const void* wts, ConvParams& p, int k_total, int w_rows, int max_ctas,
                       cudaStream_t stream) {
  constexpr int STAGE_BYTES = SUB * 128 * 128 + BN * 128;
  constexpr int STAGES = (196608 / STAGE_BYTES) > 6 ? 6 : (196608 / STAGE_BYTES);
  constexpr int smem = STAGES * STAGE_BYTES + 1024;
  CUtensorMap tin, tw;
  uint64_t dims[4] = {(uint64_t)p.cin, (uint64_t)p.w, (uint64_t)p.h, (uint64_t)p.n};
  uint64_t str[3] = {(uint64_t)p.cin * 2, (uint64_t)p.w * p.cin * 2, (uint64_t)p.h * p.w * p.cin * 2};
  uint32_t box[4] = {64, (uint32_t)p.w_box, (uint32_t)p.h_box, (uint32_t)p.n_box};
  if (make_tma_desc(&tin, TMA_BF16, in, 4, dims, str, box, TMA_SWIZZLE_128B)) return -1;
  uint64_t wd[2] = {(uint64_t)k_total, (uint64_t)w_rows};
  uint64_t ws[1] = {(uint64_t)k_total * 2};
  uint32_t wb[2] = {64, (uint32_t)BN};
  if (make_tma_desc(&tw, TMA_BF16, wts, 2, wd, ws, wb, TMA_SWIZZLE_128B)) return -1;
  p.n_tiles = (p.cout + BN - 1) / BN;
  auto kern = conv_igemm_kernel<BN, MODE, SUB>;
  static bool configured = false;
  if (!configured) {
    VPB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  int grid = p.super_tiles * (MODE == MODE_DECONV ? 4 : 1) * p.n_tiles;
  const int cap = max_ctas > 0 ? max_ctas : sm_count();
  if (grid > cap) grid = cap;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(CV_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = pdl_launch_attr(&attr[0]);
  VPB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tin, tw, p));
  return 0;
}

int deconv4x4s2_bn_relu(const void* in, const void* wphase, const float* scale, const float* shift, void* out,
                        int n, int h, int w, int cin, int cout, int max_ctas, cudaStream_t stream) {
  return deconv4x4s2_affine(in, wphase, scale, shift, out, n, h, w, cin, cout, 1, max_ctas, stream);
}

int deconv4x4s2_affine(const void* in, const void* wphase, const float* scale, const float* shift, void* out, int n,
                       int h, int w, int cin, int cout, int relu, int max_ctas, cudaStream_t stream) {
  VPB_REQUIRE(n > 0 && cin % 64 == 0 && cout % 8 == 0, "deconv: need Cin %% 64 == 0, Cout %% 8 == 0 (Cin=%d Cout=%d)",
              cin, cout);
  ConvParams p{};
  p.n = n; p.h = h; p.w = w; p.cin = cin; p.cout = cout; p.scale = scale; p.shift = shift; p.out = out;
  p.floor = relu ? 0.f : -INFINITY;
  // Cout = 256 (every ViTPose decoder): one 256-channel tile over 256-pixel strips (VPB_DECONV_WIDE=0: the 128-channel
  // tiles over 384 pixels, for A/B)
  static const bool wide = getenv("VPB_DECONV_WIDE") == nullptr || atoi(getenv("VPB_DECONV_WIDE")) != 0;
  if (wide && cout % 256 == 0 && pick_boxes(n, h, w, p, 256) == 0)
    return launch_conv<256, MODE_DECONV, 2>(in, wphase, p, 4 * cin, 4 * cout, max_ctas, stream);
  VPB_REQUIRE(pick_boxes(n, h, w, p) == 0, "deconv: %dx%d input does not tile into 384-pixel TMA boxes", h, w);
  if (cout % 128 == 0 || cout > 64) return launch_conv<128, MODE_DECONV>(in, wphase, p, 4 * cin, 4 * cout, max_ctas, stream);
  return launch_conv<64, MODE_DECONV>(in, wphase, p, 4 * cin, 4 * cout, max_ctas, stream);
}

int conv3x3_nchw_out(const void* in, const void* w9, const float* bias, float* out, int n, int h, int w, int cin,
                     int cout, int max_ctas, cudaStream_t stream) {
  VPB_REQUIRE(n > 0 && cin % 64 == 0, "conv3x3: need Cin %% 64 == 0 (Cin=%d)", cin);
  ConvParams p{};
  p.n = n; p.h = h; p.w = w; p.cin = cin; p.cout = cout; p.scale = nullptr; p.shift = bias; p.out = out;
  VPB_REQUIRE(pick_boxes(n, h, w, p) == 0, "conv3x3: %dx%d input does not tile into 384-pixel TMA boxes", h, w);
  if (cout <= 32) return launch_conv<32, MODE_CONV3>(in, w9, p, 9 * cin, cout, max_ctas, stream);
  return launch_conv<144, MODE_CONV3>(in, w9, p, 9 * cin, cout, max_ctas, stream);
}

}  // namespace vpb
