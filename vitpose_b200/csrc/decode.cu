// Fused heatmap decode: flip-back(+shift) + average -> first-index argmax -> refinement
// {none | quarter offset | DARK/unbiased Taylor | UDP-DARK} -> transform_preds, one CTA per (crop, keypoint) map.
//
// Replaces (reference, host NumPy/OpenCV): flip_back post_transforms.py:110-147, the shift at
// topdown_heatmap_simple_head.py:223-224, the average at top_down.py:187-188, _get_max_preds
// top_down_eval.py:63-95, the refinement branches of keypoints_from_heatmaps :562-612 (_taylor :298-332,
// post_dark_udp :335-396, _gaussian_blur :399-438) and transform_preds post_transforms.py:150-194.
//
// HBM-bound: each fp32 map (and its flipped partner) is read exactly once with 16-byte loads into shared
// memory; everything else (argmax, 7-point windowed blur or full separable blur, 2x2 solve) runs out of
// shared memory. Arithmetic order follows the reference's float32 NumPy expressions (no FMA contraction
// where it could change an argmax or a sign), so indices and maxvals are bit-exact.
#include <math.h>

#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

namespace vpb {

constexpr int DEC_THREADS = 128;
constexpr int DEC_MAX_PIX = 16384;  // 128x96 (a three-deconv head) and the reference's 64x64 KAT; ViTPose maps are 64x48
constexpr int DEC_MAX_TAPS = 31;

struct DecodeParams {
  const float* hm;
  const float* hmf;         // raw heatmaps of the flipped pass, or null
  const int* flip_index;    // [K] channel permutation, or null (identity)
  int shift;
  int N, K, H, W;
  int mode, ksize, use_udp, apply_transform;
  const float* center;
  const float* scale;
  float* preds;
  float* maxvals;
  float* merged_out;
  int* argmax_out;
  float taps[DEC_MAX_TAPS];
};

__device__ __forceinline__ int reflect101(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}
__device__ __forceinline__ int clampi(int i, int lo, int hi) { return i < lo ? lo : (i > hi ? hi : i); }

// merged value of map `map` at (y, x), recomputed from global memory (used only for the reference's
// flat-index underflow quirk, where a dead map reads three stencil points from the previous map)
__device__ float merged_from_global(const DecodeParams& p, int map, int y, int x) {
  const int n = map / p.K, k = map - n * p.K;
  float v = p.hm[(static_cast<size_t>(map) * p.H + y) * p.W + x];
  if (p.hmf != nullptr) {
    const int xs = p.shift ? (x > 0 ? x - 1 : 0) : x;
    const int ks = p.flip_index ? p.flip_index[k] : k;
    const float f = p.hmf[(static_cast<size_t>(n * p.K + ks) * p.H + y) * p.W + (p.W - 1 - xs)];
    v = __fmul_rn(__fadd_rn(v, f), 0.5f);
  }
  return v;
}

__global__ void __launch_bounds__(DEC_THREADS) decode_kernel(const DecodeParams p) {
  // dynamic shared memory sized to the actual map (two 64x48 maps = 24 KB -> 8 CTAs per SM instead of 6):
  // s_map = this map (merged in place), s_aux = the flipped partner / the row-blurred map (UNBIASED)
  extern __shared__ __align__(16) float s_dyn[];
  float* s_map = s_dyn;
  float* s_aux = s_dyn + ((p.H * p.W + 3) & ~3);
  __shared__ float s_red_v[DEC_THREADS / 32];
  __shared__ int s_red_i[DEC_THREADS / 32];
  __shared__ float s_part[7 * 32];

  const int map = blockIdx.x;
  const int n = map / p.K;
  const int k = map - n * p.K;
  const int H = p.H, W = p.W, HW = H * W;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  // ---- 1. load (both passes), merge, optional store of the merged map -------------------------
  // Each map is one contiguous HW*4-byte row in HBM: thread 0 fetches it (and its flipped partner) with a bulk
  // asynchronous copy straight into shared memory — two instructions per CTA instead of ~12 LDG.128 per thread.
  const float* src = p.hm + static_cast<size_t>(map) * HW;
  const float* srcf = nullptr;
  if (p.hmf != nullptr) {
    const int ks = p.flip_index ? p.flip_index[k] : k;
    srcf = p.hmf + static_cast<size_t>(n * p.K + ks) * HW;
  }
  const bool bulk = (HW % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(srcf)) & 15) == 0;
  if (bulk) {
    __shared__ uint64_t bar;
    if (tid == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = static_cast<uint32_t>(HW) * 4u;
      mbar_arrive_expect_tx(&bar, srcf ? 2 * bytes : bytes);
      bulk_load_1d(s_map, src, bytes, &bar);
      if (srcf) bulk_load_1d(s_aux, srcf, bytes, &bar);
    }
    mbar_wait(&bar, 0);
  } else {
    for (int i = tid; i < HW; i += DEC_THREADS) s_map[i] = __ldg(src + i);
    if (srcf)
      for (int i = tid; i < HW; i += DEC_THREADS) s_aux[i] = __ldg(srcf + i);
    __syncthreads();
  }
  // ---- 2. flip merge fused with the arg-max scan (first index wins on ties, np.argmax) -----------------------
  // One pass over the map: 16-byte groups, the flipped partner read mirrored (components reversed), the merged
  // values written back for the refinement stage, the running maximum kept in registers.
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  auto consider = [&](float v, int i) {
    if (v > best || best_i == 0x7fffffff) { best = v; best_i = i; }   // increasing i: strict > keeps the first
  };
  if (HW % 4 == 0 && W % 4 == 0) {
    const int W4 = W / 4;
    float4* m4 = reinterpret_cast<float4*>(s_map);
    const float4* a4 = reinterpret_cast<const float4*>(s_aux);
    for (int g = tid; g < HW / 4; g += DEC_THREADS) {
      float4 a = m4[g];
      if (srcf != nullptr) {
        const int y = g / W4, x4 = g - y * W4;
        float4 b;
        if (!p.shift) {
          const float4 t = a4[y * W4 + (W4 - 1 - x4)];
          b = make_float4(t.w, t.z, t.y, t.x);
        } else {                       // out[x] = flipped[max(x - 1, 0)]: not 16-byte aligned, scalar reads
          const float* row = s_aux + y * W;
          const int x0 = 4 * x4;
          b = make_float4(row[W - 1 - (x0 > 0 ? x0 - 1 : 0)], row[W - 1 - x0], row[W - 2 - x0], row[W - 3 - x0]);
        }
        // (a + b) * 0.5 exactly as NumPy float32: one rounded add, then an exact halving
        a.x = __fmul_rn(__fadd_rn(a.x, b.x), 0.5f);
        a.y = __fmul_rn(__fadd_rn(a.y, b.y), 0.5f);
        a.z = __fmul_rn(__fadd_rn(a.z, b.z), 0.5f);
        a.w = __fmul_rn(__fadd_rn(a.w, b.w), 0.5f);
        m4[g] = a;
      }
      consider(a.x, 4 * g);
      consider(a.y, 4 * g + 1);
      consider(a.z, 4 * g + 2);
      consider(a.w, 4 * g + 3);
    }
  } else {
    if (srcf != nullptr) {
      for (int y = warp; y < H; y += DEC_THREADS / 32) {
        const float* arow = s_aux + y * W;
        float* mrow = s_map + y * W;
        for (int x = lane; x < W; x += 32) {
          const int xs = p.shift ? (x > 0 ? x - 1 : 0) : x;
          mrow[x] = __fmul_rn(__fadd_rn(mrow[x], arow[W - 1 - xs]), 0.5f);
        }
      }
      __syncthreads();
    }
    for (int i = tid; i < HW; i += DEC_THREADS) consider(s_map[i], i);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, off);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  if (lane == 0) { s_red_v[warp] = best; s_red_i[warp] = best_i; }
  __syncthreads();
  if (p.merged_out != nullptr) {       // every thread's merged values are in shared memory after the barrier above
    float* dst = p.merged_out + static_cast<size_t>(map) * HW;
    const bool vec = (HW % 4 == 0) && (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    if (vec) {
      for (int i = tid; i < HW / 4; i += DEC_THREADS)
        reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(s_map)[i];
    } else {
      for (int i = tid; i < HW; i += DEC_THREADS) dst[i] = s_map[i];
    }
  }
  best = s_red_v[0];
  best_i = s_red_i[0];
#pragma unroll
  for (int w = 1; w < DEC_THREADS / 32; ++w) {
    const float ov = s_red_v[w];
    const int oi = s_red_i[w];
    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  const float maxval = best;
  const bool alive = maxval > 0.0f;
  const int ix = alive ? best_i % W : -1;
  const int iy = alive ? best_i / W : -1;
  float cx = static_cast<float>(ix), cy = static_cast<float>(iy);

  // ---- 3. refinement -----------------------------------------------------------------------------
  const int ks = p.ksize, r = (ks - 1) / 2;
  if (p.mode == DECODE_UDP_DARK) {
    // Only 7 values of the blurred log-map are ever read: evaluate the separable blur there.
    // stencil point `pt` of this map: (map, y, x) it reads (every thread derives its own; no broadcast needed)
    auto stencil_point = [&](int pt, int& pm, int& py, int& px) {
      const int dxs = (0x2424 >> (2 * pt)) & 3;    // dx = {0,1,-1,0,0,1,-1} as 2-bit codes (2 = -1)
      const int dys = (0x2640 >> (2 * pt)) & 3;    // dy = {0,0,0,1,-1,1,-1}
      pm = map;
      if (alive) {
        px = clampi(ix + (dxs == 2 ? -1 : dxs), 0, W - 1);
        py = clampi(iy + (dys == 2 ? -1 : dys), 0, H - 1);
      } else {
        // coords = -1: the reference's flat gather underflows into the previous map's last padded row
        const int prev = (map + p.N * p.K - 1) % (p.N * p.K);
        if (pt == 2 || pt == 6) { pm = prev; py = H - 1; px = W - 1; }
        else if (pt == 4)       { pm = prev; py = H - 1; px = 0; }
        else                    { py = 0; px = 0; }
      }
    };
    for (int t = tid; t < 7 * ks; t += DEC_THREADS) {
      const int pt = t / ks, tr = t - pt * ks;
      int pm, py, px;
      stencil_point(pt, pm, py, px);
      const int yy = reflect101(py + tr - r, H);
      float acc = 0.0f;
      for (int tx = 0; tx < ks; ++tx) {
        const int xx = reflect101(px + tx - r, W);
        const float v = (pm == map) ? s_map[yy * W + xx] : merged_from_global(p, pm, yy, xx);
        acc = __fadd_rn(acc, __fmul_rn(p.taps[tx], v));
      }
      s_part[pt * 32 + tr] = acc;
    }
    __syncthreads();
    if (warp == 0) {                      // column pass + log on lanes 0..6, gathered into lane 0 by shuffles
      float lv = 0.0f;
      if (lane < 7) {
        float acc = 0.0f;
        for (int tr = 0; tr < ks; ++tr) acc = __fadd_rn(acc, __fmul_rn(p.taps[tr], s_part[lane * 32 + tr]));
        acc = fminf(fmaxf(acc, 0.001f), 50.0f);
        lv = logf(acc);
      }
      const float v0 = __shfl_sync(0xffffffffu, lv, 0), xp = __shfl_sync(0xffffffffu, lv, 1),
                  xm = __shfl_sync(0xffffffffu, lv, 2), yp = __shfl_sync(0xffffffffu, lv, 3),
                  ym = __shfl_sync(0xffffffffu, lv, 4), xpyp = __shfl_sync(0xffffffffu, lv, 5),
                  xmym = __shfl_sync(0xffffffffu, lv, 6);
      if (lane == 0) {
        const float dx = __fmul_rn(0.5f, __fsub_rn(xp, xm));
        const float dy = __fmul_rn(0.5f, __fsub_rn(yp, ym));
        const float two_v0 = __fmul_rn(2.0f, v0);
        const float dxx = __fadd_rn(__fsub_rn(xp, two_v0), xm);
        const float dyy = __fadd_rn(__fsub_rn(yp, two_v0), ym);
        float t = __fsub_rn(xpyp, xp);
        t = __fsub_rn(t, yp);
        t = __fadd_rn(t, v0);
        t = __fadd_rn(t, v0);
        t = __fsub_rn(t, xm);
        t = __fsub_rn(t, ym);
        t = __fadd_rn(t, xmym);
        const float dxy = __fmul_rn(0.5f, t);
        // float64 solve of (H + eps32 * I) delta = g, as np.linalg.inv on the float64-promoted Hessian
        const double eps = 1.1920928955078125e-07;
        const double a = static_cast<double>(dxx) + eps, b = static_cast<double>(dxy),
                     d = static_cast<double>(dyy) + eps;
        const double det = a * d - b * b;
        const double ddx = (d * static_cast<double>(dx) - b * static_cast<double>(dy)) / det;
        const double ddy = (a * static_cast<double>(dy) - b * static_cast<double>(dx)) / det;
        cx = static_cast<float>(static_cast<double>(cx) - ddx);
        cy = static_cast<float>(static_cast<double>(cy) - ddy);
      }
    }
  } else if (p.mode == DECODE_UNBIASED) {
    // full zero-bordered separable blur (the rescale needs the max of the blurred map)
    float bmax = -INFINITY;
    if (ks == 11 && W % 4 == 0) {
      // Register-blocked 11-tap passes (the shipped kernel size). Same per-output summation order as the generic
      // loop below (acc = 0; acc += tap[t] * v, t ascending, zeros outside the map), so results are bit-identical.
      const int W4 = W / 4;
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* m4 = reinterpret_cast<const float4*>(s_map);
      float4* a4 = reinterpret_cast<float4*>(s_aux);
      // rows: one 4-pixel group per thread-iteration; its 14 inputs come from five 16-byte loads
      for (int g = tid; g < H * W4; g += DEC_THREADS) {
        const int y = g / W4, x4 = g - y * W4;
        float in[20];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const int gx = x4 + q - 2;
          const float4 v = (gx >= 0 && gx < W4) ? m4[y * W4 + gx] : zero4;
          in[4 * q] = v.x; in[4 * q + 1] = v.y; in[4 * q + 2] = v.z; in[4 * q + 3] = v.w;
        }
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float acc = 0.0f;
#pragma unroll
          for (int tx = 0; tx < 11; ++tx) acc = __fadd_rn(acc, __fmul_rn(p.taps[tx], in[c + tx + 3]));
          o[c] = acc;
        }
        a4[g] = make_float4(o[0], o[1], o[2], o[3]);
      }
      __syncthreads();
      // columns: one (4-pixel column group, 8-row segment) per thread, an 11-row window sliding down in registers
      float4* o4 = reinterpret_cast<float4*>(s_map);
      const int nseg = (H + 7) / 8;
      for (int item = tid; item < W4 * nseg; item += DEC_THREADS) {
        const int seg = item / W4, x4 = item - seg * W4;
        const int y0 = seg * 8;
        float4 win[11];
#pragma unroll
        for (int j = 0; j < 10; ++j) {
          const int yy = y0 + j - 5;
          win[j + 1] = (yy >= 0 && yy < H) ? a4[yy * W4 + x4] : zero4;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
          for (int j = 0; j < 10; ++j) win[j] = win[j + 1];
          const int yy = y0 + i + 5;
          win[10] = (yy < H) ? a4[yy * W4 + x4] : zero4;
          if (y0 + i < H) {
            float4 acc = zero4;
#pragma unroll
            for (int ty = 0; ty < 11; ++ty) {
              acc.x = __fadd_rn(acc.x, __fmul_rn(p.taps[ty], win[ty].x));
              acc.y = __fadd_rn(acc.y, __fmul_rn(p.taps[ty], win[ty].y));
              acc.z = __fadd_rn(acc.z, __fmul_rn(p.taps[ty], win[ty].z));
              acc.w = __fadd_rn(acc.w, __fmul_rn(p.taps[ty], win[ty].w));
            }
            o4[(y0 + i) * W4 + x4] = acc;      // the un-blurred map is no longer needed
            bmax = fmaxf(bmax, fmaxf(fmaxf(acc.x, acc.y), fmaxf(acc.z, acc.w)));
          }
        }
      }
    } else {
      for (int i = tid; i < HW; i += DEC_THREADS) {
        const int y = i / W, x = i - y * W;
        float acc = 0.0f;
        for (int tx = 0; tx < ks; ++tx) {
          const int xx = x + tx - r;
          const float v = (xx >= 0 && xx < W) ? s_map[y * W + xx] : 0.0f;
          acc = __fadd_rn(acc, __fmul_rn(p.taps[tx], v));
        }
        s_aux[i] = acc;
      }
      __syncthreads();
      for (int i = tid; i < HW; i += DEC_THREADS) {
        const int y = i / W, x = i - y * W;
        float acc = 0.0f;
        for (int ty = 0; ty < ks; ++ty) {
          const int yy = y + ty - r;
          const float v = (yy >= 0 && yy < H) ? s_aux[yy * W + x] : 0.0f;
          acc = __fadd_rn(acc, __fmul_rn(p.taps[ty], v));
        }
        s_map[i] = acc;            // the un-blurred map is no longer needed
        bmax = fmaxf(bmax, acc);
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, off));
    __syncthreads();             // s_red_v reuse + s_map writes visible
    if (lane == 0) s_red_v[warp] = bmax;
    __syncthreads();
    if (tid == 0 && ix > 1 && ix < W - 2 && iy > 1 && iy < H - 2) {
      bmax = fmaxf(fmaxf(s_red_v[0], s_red_v[1]), fmaxf(s_red_v[2], s_red_v[3]));
      const float ratio = __fdiv_rn(maxval, bmax);
      auto L = [&](int yy, int xx) { return logf(fmaxf(__fmul_rn(s_map[yy * W + xx], ratio), 1e-10f)); };
      const float c0 = L(iy, ix);
      const float dx = __fmul_rn(0.5f, __fsub_rn(L(iy, ix + 1), L(iy, ix - 1)));
      const float dy = __fmul_rn(0.5f, __fsub_rn(L(iy + 1, ix), L(iy - 1, ix)));
      const float dxx = __fmul_rn(0.25f, __fadd_rn(__fsub_rn(L(iy, ix + 2), __fmul_rn(2.0f, c0)), L(iy, ix - 2)));
      const float dxy = __fmul_rn(
          0.25f, __fadd_rn(__fsub_rn(__fsub_rn(L(iy + 1, ix + 1), L(iy - 1, ix + 1)), L(iy + 1, ix - 1)),
                           L(iy - 1, ix - 1)));
      const float dyy = __fmul_rn(0.25f, __fadd_rn(__fsub_rn(L(iy + 2, ix), __fmul_rn(2.0f, c0)), L(iy - 2, ix)));
      const float det = __fsub_rn(__fmul_rn(dxx, dyy), __fmul_rn(dxy, dxy));
      if (det != 0.0f) {
        const double a = dxx, b = dxy, d = dyy;
        const double dd = a * d - b * b;
        const float ox = static_cast<float>(-(d * dx - b * dy) / dd);
        const float oy = static_cast<float>(-(a * dy - b * dx) / dd);
        cx = __fadd_rn(cx, ox);
        cy = __fadd_rn(cy, oy);
      }
    }
  } else if (p.mode == DECODE_DEFAULT) {
    if (tid == 0 && ix > 1 && ix < W - 1 && iy > 1 && iy < H - 1) {
      const float ddx = __fsub_rn(s_map[iy * W + ix + 1], s_map[iy * W + ix - 1]);
      const float ddy = __fsub_rn(s_map[(iy + 1) * W + ix], s_map[(iy - 1) * W + ix]);
      cx += (ddx > 0.0f) ? 0.25f : (ddx < 0.0f ? -0.25f : 0.0f);   // np.sign(0) = 0
      cy += (ddy > 0.0f) ? 0.25f : (ddy < 0.0f ? -0.25f : 0.0f);
    }
  }

  // ---- 4. transform_preds (float32, NumPy evaluation order) ------------------------------------
  if (tid == 0) {
    if (p.apply_transform) {
      const float sx200 = __fmul_rn(p.scale[2 * n + 0], 200.0f);
      const float sy200 = __fmul_rn(p.scale[2 * n + 1], 200.0f);
      const float den_x = p.use_udp ? static_cast<float>(W) - 1.0f : static_cast<float>(W);
      const float den_y = p.use_udp ? static_cast<float>(H) - 1.0f : static_cast<float>(H);
      const float kx = __fdiv_rn(sx200, den_x), ky = __fdiv_rn(sy200, den_y);
      cx = __fsub_rn(__fadd_rn(__fmul_rn(cx, kx), p.center[2 * n + 0]), __fmul_rn(sx200, 0.5f));
      cy = __fsub_rn(__fadd_rn(__fmul_rn(cy, ky), p.center[2 * n + 1]), __fmul_rn(sy200, 0.5f));
    }
    p.preds[2 * map + 0] = cx;
    p.preds[2 * map + 1] = cy;
    p.maxvals[map] = maxval;
    if (p.argmax_out != nullptr) p.argmax_out[map] = best_i;
  }
}

// Standalone flip_back (post_transforms.py:110-147) (+ optional shift_heatmap, simple_head.py:223-224):
// out[n,k,y,x] = in[n, perm[k], y, W-1-xs], xs = shift ? max(x-1,0) : x. One read, one write per element.
__global__ void flip_back_kernel(const float* __restrict__ in, const int* __restrict__ perm, float* __restrict__ out,
                                 int K, int H, int W, int shift, long long total) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = static_cast<int>(i % W);
  const long long row = i / W;                 // (n*K + k)*H + y
  const int y = static_cast<int>(row % H);
  const long long map = row / H;
  const int k = static_cast<int>(map % K);
  const long long n = map / K;
  const int ks = perm ? perm[k] : k;
  const int xs = shift ? (x > 0 ? x - 1 : 0) : x;
  out[i] = __ldg(in + ((n * K + ks) * H + y) * W + (W - 1 - xs));
}

int flip_back(const float* in, const int* perm, float* out, int N, int K, int H, int W, int shift,
              cudaStream_t stream) {
  VPB_REQUIRE(N >= 0 && K > 0 && H > 0 && W > 0, "flip_back: bad shape");
  const long long total = static_cast<long long>(N) * K * H * W;
  if (total == 0) return 0;
  flip_back_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(in, perm, out, K, H, W, shift, total);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Standalone transform_preds (post_transforms.py:150-194) on float32: coords [N,K,2] with per-crop center/scale.
__global__ void transform_preds_kernel(const float* __restrict__ coords, const float* __restrict__ center,
                                       const float* __restrict__ scale, float* __restrict__ out, int K, float den_x,
                                       float den_y, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = i / K;
  const float sx200 = __fmul_rn(scale[2 * n + 0], 200.0f), sy200 = __fmul_rn(scale[2 * n + 1], 200.0f);
  const float kx = __fdiv_rn(sx200, den_x), ky = __fdiv_rn(sy200, den_y);
  out[2 * i + 0] = __fsub_rn(__fadd_rn(__fmul_rn(coords[2 * i + 0], kx), center[2 * n + 0]), __fmul_rn(sx200, 0.5f));
  out[2 * i + 1] = __fsub_rn(__fadd_rn(__fmul_rn(coords[2 * i + 1], ky), center[2 * n + 1]), __fmul_rn(sy200, 0.5f));
}

int transform_preds(const float* coords, const float* center, const float* scale, float* out, int N, int K, int W,
                    int H, int use_udp, cudaStream_t stream) {
  VPB_REQUIRE(N >= 0 && K > 0, "transform_preds: bad shape");
  if (N == 0) return 0;
  const float dx = use_udp ? W - 1.0f : static_cast<float>(W), dy = use_udp ? H - 1.0f : static_cast<float>(H);
  transform_preds_kernel<<<(N * K + 127) / 128, 128, 0, stream>>>(coords, center, scale, out, K, dx, dy, N * K);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static void gaussian_taps_host(int ksize, float* taps) {
  // cv2.getGaussianKernel(ksize, sigma<=0) as float32 (fixed tables for ksize <= 9, else sampled Gaussian)
  static const float t1[] = {1.f};
  static const float t3[] = {0.25f, 0.5f, 0.25f};
  static const float t5[] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
  static const float t7[] = {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f};
  static const float t9[] = {4.f / 256, 13.f / 256, 30.f / 256, 51.f / 256, 60.f / 256,
                             51.f / 256, 30.f / 256, 13.f / 256, 4.f / 256};
  const float* tab = ksize == 1 ? t1 : ksize == 3 ? t3 : ksize == 5 ? t5 : ksize == 7 ? t7 : ksize == 9 ? t9 : nullptr;
  if (tab) {
    for (int i = 0; i < ksize; ++i) taps[i] = tab[i];
    return;
  }
  const double sigma = 0.3 * ((ksize - 1) * 0.5 - 1.0) + 0.8;
  const double s2 = -0.5 / (sigma * sigma);
  double w[DEC_MAX_TAPS], sum = 0.0;
  for (int i = 0; i < ksize; ++i) {
    const double x = i - (ksize - 1) * 0.5;
    w[i] = exp(s2 * x * x);
    sum += w[i];
  }
  for (int i = 0; i < ksize; ++i) taps[i] = static_cast<float>(w[i] / sum);
}

int decode_heatmaps(const float* hm, const float* hm_flipped, const int* flip_index, int shift_heatmap, int N, int K,
                    int H, int W, int mode, int kernel, int use_udp, int apply_transform, const float* center,
                    const float* scale, float* preds, float* maxvals, float* merged_out, int* argmax_out,
                    cudaStream_t stream) {
  VPB_REQUIRE(N >= 0 && K > 0 && H > 0 && W > 0, "decode: bad shape N=%d K=%d H=%d W=%d", N, K, H, W);
  if (N == 0) return 0;
  VPB_REQUIRE(H * W <= DEC_MAX_PIX, "decode: heatmap %dx%d exceeds %d pixels", H, W, DEC_MAX_PIX);
  VPB_REQUIRE(mode >= DECODE_NONE && mode <= DECODE_UDP_DARK, "decode: unknown mode %d", mode);
  if (mode == DECODE_UNBIASED || mode == DECODE_UDP_DARK) {
    VPB_REQUIRE(kernel > 0 && kernel % 2 == 1 && kernel <= DEC_MAX_TAPS, "decode: kernel=%d must be odd in [1,%d]",
                kernel, DEC_MAX_TAPS);
    VPB_REQUIRE(kernel / 2 < H && kernel / 2 < W, "decode: kernel=%d too large for %dx%d", kernel, H, W);
  }
  VPB_REQUIRE(!apply_transform || (center != nullptr && scale != nullptr), "decode: center/scale missing");
  VPB_REQUIRE((reinterpret_cast<uintptr_t>(hm) & 15) == 0 && (reinterpret_cast<uintptr_t>(hm_flipped) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(merged_out) & 15) == 0,
              "decode: heatmap pointers must be 16-byte aligned");
  DecodeParams p;
  p.hm = hm; p.hmf = hm_flipped; p.flip_index = flip_index; p.shift = shift_heatmap;
  p.N = N; p.K = K; p.H = H; p.W = W;
  p.mode = mode; p.ksize = kernel > 0 ? kernel : 1; p.use_udp = use_udp; p.apply_transform = apply_transform;
  p.center = center; p.scale = scale; p.preds = preds; p.maxvals = maxvals;
  p.merged_out = merged_out; p.argmax_out = argmax_out;
  for (int i = 0; i < DEC_MAX_TAPS; ++i) p.taps[i] = 0.f;
  if (mode == DECODE_UNBIASED || mode == DECODE_UDP_DARK) gaussian_taps_host(p.ksize, p.taps);
  const bool need_aux = hm_flipped != nullptr || mode == DECODE_UNBIASED;
  const size_t smem = static_cast<size_t>((H * W + 3) & ~3) * sizeof(float) * (need_aux ? 2 : 1);
  if (smem > 48 * 1024) {     // maps larger than 64 x 48 with a flipped partner: opt in to more dynamic shared memory
    static bool configured = false;
    if (!configured) {
      VPB_CHECK_CUDA(cudaFuncSetAttribute(decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          2 * DEC_MAX_PIX * static_cast<int>(sizeof(float))));
      configured = true;
    }
  }
  decode_kernel<<<N * K, DEC_THREADS, smem, stream>>>(p);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vpb
