// Post-decode evaluation step of the top-down COCO datasets, per image:
//   rescoring   mmpose/datasets/datasets/top_down/topdown_coco_dataset.py:476-490
//   oks_nms / soft_oks_nms   mmpose/core/post_processing/nms.py:51-207
// One CTA per image (its poses are a contiguous group). The reference is O(P^2 K) NumPy per image on the host; here
// the greedy loop runs in shared memory with the pairwise OKS evaluated by the CTA's threads. Arithmetic follows
// NumPy's: squared distances in fp32 (no FMA), everything else in fp64, np.sum's 8-accumulator pairwise order, the
// OKS value rounded to fp32 before the threshold test.
#include "host_util.h"
#include "ops.h"

namespace vpb {

constexpr int NMS_THREADS = 128;
constexpr int NMS_MAX_K = 136;

// np.add.reduce over a contiguous fp64 vector (numpy/core/src/umath/loops_utils.h.src, pairwise_sum):
// n < 8: sequential; n <= 128: eight accumulators then a fixed tree, tail sequential; else split at n/2 rounded
// down to a multiple of 8.
__device__ __forceinline__ double np_sum_block(const double* a, int n) {      // n <= 128
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; ++i) res += a[i];
    return res;
  }
  double r[8];
  for (int j = 0; j < 8; ++j) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int j = 0; j < 8; ++j) r[j] += a[i + j];
  double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  for (; i < n; ++i) res += a[i];
  return res;
}
__device__ __forceinline__ double np_pairwise_sum(const double* a, int n) {   // n <= NMS_MAX_K < 256: one split at most
  if (n <= 128) return np_sum_block(a, n);
  int n2 = n / 2;
  n2 -= n2 % 8;
  return np_sum_block(a, n2) + np_sum_block(a + n2, n - n2);
}

__device__ float oks_pair(const float* __restrict__ g, const float* __restrict__ d, double a_g, double a_d,
                          const double* __restrict__ var, int K, bool use_vis, double vis_thr, bool area_f32 = false) {
  double e[NMS_MAX_K];
  int n = 0;
  // (a_g + a_d) / 2 + np.spacing(1): with float32 areas (the dataset path hands over the float32 `boxes[:, 4]`) the
  // sum and the halving are float32 operations, only np.spacing(1) widens the result
  const double half = area_f32 ? static_cast<double>(__fmul_rn(__fadd_rn(static_cast<float>(a_g), static_cast<float>(a_d)), 0.5f))
                               : (a_g + a_d) / 2;
  const double denom = half + 2.220446049250313e-16;     // np.spacing(1)
  for (int k = 0; k < K; ++k) {
    if (use_vis && !(static_cast<double>(d[3 * k + 2]) > vis_thr)) continue;
    const float dx = __fsub_rn(d[3 * k], g[3 * k]), dy = __fsub_rn(d[3 * k + 1], g[3 * k + 1]);
    const float sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    e[n++] = exp(-(static_cast<double>(sq) / var[k] / denom / 2));
  }
  if (n == 0) return 0.f;
  return static_cast<float>(np_pairwise_sum(e, n) / n);
}

struct NmsParams {
  const float* kpts;        // [P, K, 3]
  const double* areas;      // [P]
  const double* box_scores; // [P] (rescore != 0) or final scores
  const int* group_start;   // [G + 1]
  const double* var;        // [K] = (2 sigma)^2
  int K;
  double thr, vis_thr;
  int use_vis, rescore, soft, max_dets;
  double* scores_out;       // [P] scores used for the ordering (after rescoring)
  int* keep;                // [P]: kept global indices of group g at keep[group_start[g] ...], selection order
  int* keep_count;          // [G]
};

__global__ void __launch_bounds__(NMS_THREADS) oks_nms_kernel(const NmsParams p) {
  extern __shared__ unsigned char smem_raw[];
  const int g = blockIdx.x;
  const int lo = p.group_start[g], P = p.group_start[g + 1] - lo;
  double* sc = reinterpret_cast<double*>(smem_raw);        // [P] current scores
  int* order = reinterpret_cast<int*>(sc + P);             // [P] local indices, descending score
  int* alive = order + P;                                  // [P]
  __shared__ int s_sel, s_kept;
  if (P <= 0) {
    if (threadIdx.x == 0) p.keep_count[g] = 0;
    return;
  }
  // ---- scores: rescoring = mean of the visible joint scores (fp32, joint order) times the box score
  for (int i = threadIdx.x; i < P; i += NMS_THREADS) {
    double s = p.box_scores[lo + i];
    if (p.rescore & 1) {
      const float* kp = p.kpts + static_cast<size_t>(lo + i) * p.K * 3;
      float acc = 0.f;
      int cnt = 0;
      for (int k = 0; k < p.K; ++k)
        if (static_cast<double>(kp[3 * k + 2]) > p.vis_thr) { acc = __fadd_rn(acc, kp[3 * k + 2]); ++cnt; }
      if (cnt != 0) acc = acc / static_cast<float>(cnt);
      s = static_cast<double>(__fmul_rn(acc, static_cast<float>(s)));
    }
    sc[i] = s;
    alive[i] = 1;
    p.scores_out[lo + i] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) s_kept = 0;
  if (!p.soft) {
    // order = argsort(scores)[::-1]; ties: the later index first
    for (int i = threadIdx.x; i < P; i += NMS_THREADS) {
      int rank = 0;
      for (int j = 0; j < P; ++j) rank += (sc[j] > sc[i]) || (sc[j] == sc[i] && j > i);
      order[rank] = i;
    }
    __syncthreads();
    for (int pos = 0; pos < P; ++pos) {
      const int i = order[pos];
      if (alive[i]) {                                   // uniform across the CTA
        if (threadIdx.x == 0) p.keep[lo + s_kept++] = lo + i;
        const float* gk = p.kpts + static_cast<size_t>(lo + i) * p.K * 3;
        for (int q = pos + 1 + threadIdx.x; q < P; q += NMS_THREADS) {
          const int j = order[q];
          if (!alive[j]) continue;
          const float o = oks_pair(gk, p.kpts + static_cast<size_t>(lo + j) * p.K * 3, p.areas[lo + i],
                                   p.areas[lo + j], p.var, p.K, p.use_vis != 0, p.vis_thr, (p.rescore & 2) != 0);
          if (!(static_cast<double>(o) <= p.thr)) alive[j] = 0;
        }
      }
      __syncthreads();
    }
  } else {
    for (int it = 0; it < p.max_dets; ++it) {
      if (threadIdx.x == 0) {                           // arg-max of the current scores among the remaining poses
        int best = -1;
        for (int j = 0; j < P; ++j)
          if (alive[j] && (best < 0 || sc[j] > sc[best] || (sc[j] == sc[best] && j > best))) best = j;
        s_sel = best;
        if (best >= 0) { p.keep[lo + s_kept++] = lo + best; alive[best] = 0; }
      }
      __syncthreads();
      const int i = s_sel;
      if (i < 0) break;
      const float* gk = p.kpts + static_cast<size_t>(lo + i) * p.K * 3;
      for (int j = threadIdx.x; j < P; j += NMS_THREADS) {
        if (!alive[j]) continue;
        const float o = oks_pair(gk, p.kpts + static_cast<size_t>(lo + j) * p.K * 3, p.areas[lo + i], p.areas[lo + j],
                                 p.var, p.K, p.use_vis != 0, p.vis_thr, (p.rescore & 2) != 0);
        // scores * np.exp(-overlap**2 / thr): the exponent stays float32 (a Python float does not widen an array)
        const float t = __fdiv_rn(-__fmul_rn(o, o), static_cast<float>(p.thr));
        sc[j] = sc[j] * static_cast<double>(expf(t));
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) p.keep_count[g] = s_kept;
}

int oks_nms(const float* kpts, const double* areas, const double* box_scores, const int* group_start, int G, int K,
            int max_group, const double* var, double thr, int use_vis, double vis_thr, int rescore, int soft,
            int max_dets, double* scores_out, int* keep, int* keep_count, cudaStream_t stream) {
  VPB_REQUIRE(G > 0 && K > 0 && K <= NMS_MAX_K, "oks_nms: K=%d must be in [1, %d]", K, NMS_MAX_K);
  VPB_REQUIRE(max_group > 0 && max_group <= 2048, "oks_nms: at most 2048 poses per image (got %d)", max_group);
  NmsParams p{kpts, areas, box_scores, group_start, var, K, thr, vis_thr, use_vis, rescore, soft, max_dets,
              scores_out, keep, keep_count};
  const size_t smem = static_cast<size_t>(max_group) * (sizeof(double) + 2 * sizeof(int));
  oks_nms_kernel<<<G, NMS_THREADS, smem, stream>>>(p);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace vpb
