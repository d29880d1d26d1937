// Fused multi-head self-attention for the fixed short ViTPose sequence (T = 16 x 12 = 192 tokens).
// Reference: Attention.forward, mmpose/models/backbones/vit.py:99-115 —
//   q, k, v = split(qkv); attn = softmax((q * scale) @ k^T); out = (attn @ v) re-interleaved per head.
//
// One CTA per (crop, head, 128-row query tile). The whole key/value sequence of one head fits on chip, so
// there is no online-softmax rescaling:
//   TMA (3-D map over [crop, token, column], token dim zero-filled past T) -> Q, K, V tiles in 128B-swizzled smem
//   tcgen05.mma  S[128 x T]  = Q . K^T            (fp32 in TMEM, K-major operands)
//   4 warps, one query row per thread: tcgen05.ld S, row max, exp2, row sum, P (bf16) written to smem in the
//            canonical K-major swizzled layout (overlays the dead Q/K tiles)
//   tcgen05.mma  O[128 x hd] = P . V              (V consumed as an MN-major operand straight from its TMA tile;
//                                                  O overlays S in TMEM)
//   tcgen05.ld O, scale by 1/rowsum, bf16 store to out[crop, token, head*hd + :].
// The scores never touch HBM (the eager reference materialises [N, h, 192, 192] fp32 per block).
// Head dims that are not a multiple of 64 (32 for ViT-S, 80 for ViT-H) are loaded as 64-column boxes; the MMA
// reads only the first hd columns / K-steps of them.
#include "host_util.h"
#include "ops.h"
#include "ptx.cuh"

namespace vpb {

constexpr int ATT_THREADS = 160;   // warp 0: TMA + MMA issue; warps 1..4: softmax + epilogue
constexpr int ATT_BM = 128;

struct AttnParams {
  int n, T, heads, hd;
  int ldo;              // heads * hd
  float scale_log2e;    // scale * log2(e)
  __nv_bfloat16* out;
};

// Tn = T rounded up to 64 (P chunks); all sizes in bytes
__host__ __device__ constexpr int att_boxes(int hd) { return (hd + 63) / 64; }

template <int HD>
__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const __grid_constant__ CUtensorMap tm_q,
                                                                const __grid_constant__ CUtensorMap tm_kv,
                                                                const AttnParams p) {
  constexpr int NB = att_boxes(HD);               // 64-column boxes per operand
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar_load, bar_s, bar_p, bar_o;
  __shared__ uint32_t tmem_slot;

  const int T = p.T;
  const int kv_box_bytes = T * 128;               // [T rows][128 B]
  const int q_box_bytes = ATT_BM * 128;
  // layout: [Q boxes][K boxes] (later overlaid by P) | [V boxes]
  const int p_chunks = (T + 63) / 64;
  const int qk_bytes = NB * (q_box_bytes + kv_box_bytes);
  const int p_bytes = p_chunks * ATT_BM * 128;
  const int region0 = qk_bytes > p_bytes ? qk_bytes : p_bytes;
  uint8_t* s_q = smem;
  uint8_t* s_k = smem + NB * q_box_bytes;
  uint8_t* s_p = smem;
  uint8_t* s_v = smem + ((region0 + 1023) & ~1023);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x;
  const int qt = unit & 1;
  const int head = (unit >> 1) % p.heads;
  const int crop = (unit >> 1) / p.heads;
  if (qt * ATT_BM >= T) return;                    // (only when T <= 128)

  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_p, 128);
    mbar_init(&bar_o, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = tmem_slot;              // S: columns [0, T); O overlays columns [0, HD)

  if (warp == 0) {
    if (lane == 0) {
      const int col_q = head * HD, col_k = p.heads * HD + head * HD, col_v = 2 * p.heads * HD + head * HD;
      mbar_arrive_expect_tx(&bar_load, NB * (q_box_bytes + 2 * kv_box_bytes));
      for (int b = 0; b < NB; ++b) {
        tma_load_3d(s_q + b * q_box_bytes, &tm_q, &bar_load, col_q + b * 64, qt * ATT_BM, crop);
        tma_load_3d(s_k + b * kv_box_bytes, &tm_kv, &bar_load, col_k + b * 64, 0, crop);
        tma_load_3d(s_v + b * kv_box_bytes, &tm_kv, &bar_load, col_v + b * 64, 0, crop);
      }
      mbar_wait(&bar_load, 0);
      tc_fence_after();
      // S = Q . K^T : M=128, N=T, K=HD
      const uint32_t idesc_s = umma_idesc_bf16(ATT_BM, T);
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        const uint32_t a = smem_u32(s_q + (ks / 4) * q_box_bytes) + (ks % 4) * 32;
        const uint32_t b = smem_u32(s_k + (ks / 4) * kv_box_bytes) + (ks % 4) * 32;
        umma_bf16_ss(tmem_s, umma_desc_k_sw128(a), umma_desc_k_sw128(b), idesc_s, ks != 0);
      }
      umma_commit(&bar_s);
      // wait for P (bf16, smem) from the softmax warps, then O = P . V : M=128, N=HD, K=T
      mbar_wait(&bar_p, 0);
      tc_fence_after();
      const uint32_t idesc_o = umma_idesc_bf16(ATT_BM, HD, 0, 1);
      const int ksteps = T / 16;
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t a = smem_u32(s_p + (ks / 4) * (ATT_BM * 128)) + (ks % 4) * 32;
        const uint32_t b = smem_u32(s_v) + ks * 2048;             // 16 tokens x 128 B per K step
        umma_bf16_ss(tmem_s, umma_desc_k_sw128(a), umma_desc_mn_sw128(b, kv_box_bytes), idesc_o, ks != 0);
      }
      umma_commit(&bar_o);
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;               // query row inside the tile == TMEM lane
    const uint32_t t_row = tmem_s + (static_cast<uint32_t>(quad * 32) << 16);
    mbar_wait(&bar_s, 0);
    tc_fence_after();
    // pass 1: row max of the raw scores
    float mx = -INFINITY;
    for (int c = 0; c < T; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + c, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
    }
    const float mxs = mx * p.scale_log2e;
    // pass 2: p = exp2(s * scale*log2e - max'), row sum, bf16 P into swizzled smem
    float sum = 0.f;
    for (int c = 0; c < T; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_row + c, v);
      tmem_ld_wait();
      uint32_t packed[16];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float e0 = exp2f(fmaf(__uint_as_float(v[j]), p.scale_log2e, -mxs));
        const float e1 = exp2f(fmaf(__uint_as_float(v[j + 1]), p.scale_log2e, -mxs));
        // the row sum must match what the tensor core will see: sum the bf16-rounded values
        const __nv_bfloat162 b2 = __floats2bfloat162_rn(e0, e1);
        sum += __low2float(b2) + __high2float(b2);
        packed[j / 2] = *reinterpret_cast<const uint32_t*>(&b2);
      }
      // columns [c, c+32) = half of 64-column chunk c/64: 16-byte units u0..u0+3, XOR-swizzled by (row & 7)
      uint8_t* chunk = s_p + (c / 64) * (ATT_BM * 128) + r * 128;
      const int u0 = (c % 64) / 8;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        *reinterpret_cast<uint4*>(chunk + (((u0 + u) ^ (r & 7)) * 16)) =
            make_uint4(packed[4 * u], packed[4 * u + 1], packed[4 * u + 2], packed[4 * u + 3]);
      }
    }
    tc_fence_before();          // our tcgen05.ld of S complete before the MMA warp overwrites S with O
    fence_proxy_async_smem();   // generic-proxy writes of P visible to the tensor core (async proxy)
    mbar_arrive(&bar_p);
    // epilogue: O row * 1/sum -> bf16
    mbar_wait(&bar_o, 0);
    tc_fence_after();
    const float inv = 1.0f / sum;
    const int token = qt * ATT_BM + r;
    __nv_bfloat16* orow = p.out + (static_cast<size_t>(crop) * T + token) * p.ldo + head * HD;
#pragma unroll
    for (int c = 0; c < HD; c += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(t_row + c, v);
      tmem_ld_wait();
      if (token < T) {
        uint4 w0 = make_uint4(pack_bf16x2(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv),
                              pack_bf16x2(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv),
                              pack_bf16x2(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv),
                              pack_bf16x2(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv));
        uint4 w1 = make_uint4(pack_bf16x2(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv),
                              pack_bf16x2(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv),
                              pack_bf16x2(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv),
                              pack_bf16x2(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv));
        reinterpret_cast<uint4*>(orow + c)[0] = w0;
        reinterpret_cast<uint4*>(orow + c)[1] = w1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_s, 256);
}

template <int HD>
static int launch_attention(const CUtensorMap& tq, const CUtensorMap& tkv, const AttnParams& p, int smem,
                            cudaStream_t stream) {
  auto kern = attention_kernel<HD>;
  static int configured_smem = 0;
  if (configured_smem < smem) {
    VPB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured_smem = smem;
  }
  kern<<<p.n * p.heads * 2, ATT_THREADS, smem, stream>>>(tq, tkv, p);
  VPB_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int attention_fwd(const void* qkv, void* out, int n, int T, int heads, int hd, float scale, int max_ctas,
                  cudaStream_t stream) {
  (void)max_ctas;
  VPB_REQUIRE(n > 0 && heads > 0, "attention: empty problem");
  VPB_REQUIRE(T % 16 == 0 && T >= 16 && T <= 256, "attention: T=%d must be a multiple of 16 in [16,256]", T);
  VPB_REQUIRE(hd % 16 == 0 && hd >= 16 && hd <= 128, "attention: head_dim=%d unsupported", hd);
  const int ld = 3 * heads * hd;
  VPB_REQUIRE(ld % 8 == 0, "attention: row pitch must be a multiple of 16 bytes");
  CUtensorMap tq, tkv;
  uint64_t dims[3] = {(uint64_t)ld, (uint64_t)T, (uint64_t)n};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)T * ld * 2};
  uint32_t box_q[3] = {64, ATT_BM, 1};
  uint32_t box_kv[3] = {64, (uint32_t)T, 1};
  if (make_tma_desc(&tq, TMA_BF16, qkv, 3, dims, strides, box_q, TMA_SWIZZLE_128B)) return -1;
  if (make_tma_desc(&tkv, TMA_BF16, qkv, 3, dims, strides, box_kv, TMA_SWIZZLE_128B)) return -1;
  AttnParams p;
  p.n = n; p.T = T; p.heads = heads; p.hd = hd; p.ldo = heads * hd;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  const int nb = att_boxes(hd);
  const int qk = nb * (ATT_BM * 128 + T * 128);
  const int pb = ((T + 63) / 64) * ATT_BM * 128;
  const int region0 = ((qk > pb ? qk : pb) + 1023) & ~1023;
  const int smem = region0 + nb * T * 128 + 1024;
  switch (hd) {
    case 32: return launch_attention<32>(tq, tkv, p, smem, stream);
    case 64: return launch_attention<64>(tq, tkv, p, smem, stream);
    case 80: return launch_attention<80>(tq, tkv, p, smem, stream);
    case 128: return launch_attention<128>(tq, tkv, p, smem, stream);
    default:
      set_last_error("attention: no kernel instance for head_dim=%d", hd);
      return -2;
  }
}

}  // namespace vpb
