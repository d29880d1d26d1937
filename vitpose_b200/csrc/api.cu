// extern "C" boundary (include/vitpose_b200.h) and the forward orchestration: the launch sequence that
// replaces ViT.forward + TopdownHeatmapSimpleHead.forward for a batch of crops and their flips.
#include "../../include/vitpose_b200.h"

#include "gemm.cuh"
#include "host_util.h"
#include "ops.h"

using namespace vpb;

#include <stdlib.h>

#include <vector>

namespace {

// Optional per-launch profiler: CUDA-event pairs recorded on the launching stream around every kernel of the
// forward / decode sequence, so bench.py can report per-kernel durations measured inside the timed step.
struct ProfRecord { const char* tag; cudaEvent_t a, b; };
struct Profiler {
  bool on = false;
  std::vector<ProfRecord> recs;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  cudaEvent_t get() {
    if (used == pool.size()) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
    return pool[used++];
  }
} g_prof;
long long g_launches = 0;

template <typename F>
int prof_run(const char* tag, cudaStream_t stream, F&& f) {
  ++g_launches;
  if (!g_prof.on) return f();
  ProfRecord r{tag, g_prof.get(), g_prof.get()};
  cudaEventRecord(r.a, stream);
  int e = f();
  cudaEventRecord(r.b, stream);
  g_prof.recs.push_back(r);
  return e;
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Workspace {
  // offsets in bytes
  size_t patches, x, xn, qkv, attn, hidden, head_a, head_b, ln_scratch, moe_scratch, total;
};

Workspace plan_workspace(const vpb_model_desc& d, int images) {
  const size_t T = static_cast<size_t>(d.img_h / 16) * (d.img_w / 16);
  const size_t rows = T * images;
  const size_t D = d.embed_dim;
  Workspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  // the im2col patches are dead once the patch-embed GEMM has run: they share the MLP hidden buffer
  const size_t hidden_bytes = rows * static_cast<size_t>(d.mlp_hidden) * 2;
  const size_t patch_bytes = rows * 768 * 2;
  w.hidden = take(hidden_bytes > patch_bytes ? hidden_bytes : patch_bytes);
  w.patches = w.hidden;
  w.x = take(rows * D * 4);
  w.xn = take(rows * D * 2);
  w.qkv = take(rows * 3 * D * 2);
  w.attn = take(rows * D * 2);
  // head activations (NHWC bf16): classic = two deconv outputs, simple = the upsampled feature map
  const size_t hp = static_cast<size_t>(d.img_h / 16), wp = static_cast<size_t>(d.img_w / 16);
  size_t a = 0, b = 0;
  if (d.num_deconv > 0) {
    a = static_cast<size_t>(images) * (2 * hp) * (2 * wp) * d.deconv_channels[0] * 2;
    if (d.num_deconv > 1) b = static_cast<size_t>(images) * (4 * hp) * (4 * wp) * d.deconv_channels[1] * 2;
    if (d.num_deconv > 2) {
      const size_t c = static_cast<size_t>(images) * (8 * hp) * (8 * wp) * d.deconv_channels[2] * 2;
      a = a > c ? a : c;
    }
  } else if (d.upsample > 0) {
    a = static_cast<size_t>(images) * (hp * d.upsample) * (wp * d.upsample) * D * 2;
  }
  w.head_a = take(a);
  w.head_b = take(b);
  // per-row LayerNorm partial statistics + arrival counters of the residual GEMMs with the fused LayerNorm
  // (+ one small slice per ViTPose+ run: each run's fc2 launch exchanges statistics in a scratch of its own)
  w.ln_scratch = take(gemm_ln_scratch_bytes(static_cast<int>(rows), static_cast<int>(D)));
  w.moe_scratch = take(gemm_ln_scratch_bytes(static_cast<int>(rows), static_cast<int>(D)) +
                       VPB_MOE_MAX_RUNS * (gemm_ln_scratch_bytes(256, static_cast<int>(D)) + 1024));
  w.total = off;
  return w;
}

}  // namespace

extern "C" {

int vpb_abi_version(void) { return VPB_ABI_VERSION; }
const char* vpb_last_error(void) { return get_last_error(); }

void vpb_profile_enable(int on) {
  g_prof.on = on != 0;
  g_prof.recs.clear();
  g_prof.used = 0;
}
int vpb_profile_count(void) { return static_cast<int>(g_prof.recs.size()); }
int vpb_profile_get(int i, const char** tag, float* ms) {
  if (i < 0 || i >= static_cast<int>(g_prof.recs.size())) return -2;
  const ProfRecord& r = g_prof.recs[i];
  VPB_CHECK_CUDA(cudaEventSynchronize(r.b));
  VPB_CHECK_CUDA(cudaEventElapsedTime(ms, r.a, r.b));
  *tag = r.tag;
  return 0;
}
long long vpb_launch_count(void) { return g_launches; }

size_t vpb_workspace_bytes(const vpb_model_desc* desc, int images) {
  if (!desc || images <= 0) return 0;
  return plan_workspace(*desc, images).total;
}

int vpb_vitpose_forward(const vpb_model_desc* desc, const vpb_weights* w, const float* img, int n, int flip,
                        void* workspace, size_t workspace_bytes, float* heatmaps, float* heatmaps_flipped,
                        void* features, void* stream_) {
  VPB_REQUIRE(desc && w && img && workspace, "forward: null argument");
  VPB_REQUIRE(n > 0, "forward: n must be positive");
  const vpb_model_desc& d = *desc;
  VPB_REQUIRE(d.img_h % 16 == 0 && d.img_w % 16 == 0, "forward: image size must be a multiple of the 16-px patch");
  VPB_REQUIRE(d.embed_dim % d.num_heads == 0, "forward: embed_dim %% num_heads != 0");
  cudaStream_t stream = as_stream(stream_);
  const int images = flip ? 2 * n : n;
  const int hp = d.img_h / 16, wp = d.img_w / 16, T = hp * wp;
  const int rows = images * T;
  const int D = d.embed_dim, hd = D / d.num_heads;
  const Workspace ws = plan_workspace(d, images);
  VPB_REQUIRE(workspace_bytes >= ws.total, "forward: workspace too small (%zu < %zu)", workspace_bytes, ws.total);
  VPB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "forward: workspace must be 1024-byte aligned");
  uint8_t* base = reinterpret_cast<uint8_t*>(workspace);
  void* patches = base + ws.patches;
  float* x = reinterpret_cast<float*>(base + ws.x);
  void* xn = base + ws.xn;
  void* qkv = base + ws.qkv;
  void* attn = base + ws.attn;
  void* hidden = base + ws.hidden;

  // The LayerNorm that follows every update of the residual stream (norm1 / norm2 of the next sub-block,
  // vit.py:138-139, and last_norm, vit.py:328) is produced by the epilogue of the GEMM that writes the stream.
  VPB_REQUIRE(d.has_last_norm, "forward: last_norm=False is not supported");
  VPB_REQUIRE(d.depth > 0, "forward: depth must be positive");
  void* ln_scratch = base + ws.ln_scratch;
  if (int e = gemm_ln_scratch_init(ln_scratch, rows, D, stream)) return e;
  unsigned epoch = 0;
  // Folded LayerNorm (vpb_block_fold): the residual GEMMs leave plain bf16 rows + per-tile (mean, M2) and qkv / fc1
  // normalise in their epilogue. The statistics live in the first half of the scratch above (the fused LayerNorm of the
  // last fc2 — last_norm has no Linear layer behind it to fold into — runs as epoch 1 and uses the second half).
  // Measured on ViTPose-B (profiles/r02_summary.md §8): proj -21 us, fc2 -6 us per launch, but qkv +17 us and fc1 +22 us
  // (their epilogues gain an FFMA2 and a shared-memory read per pair of outputs on a power-capped GPU) -> the step is
  // no faster. Kept behind VPB_LN_FOLD=1 (and only when the caller supplies w->fold).
  static const bool fold_enabled = [] {
    const char* e = getenv("VPB_LN_FOLD");
    return e != nullptr && atoi(e) != 0;
  }();
  const int parts = gemm_ln_parts(D), part_cols = gemm_ln_part_cols(D);
  // ---- ViTPose+ runs (vpb_moe_runs): row ranges with their own fc2 weights and their own statistics scratch
  const vpb_moe_runs* moe = w->moe;
  struct MoeRun { int row0, rows; void* scratch; };
  MoeRun runs[VPB_MOE_MAX_RUNS];
  int n_runs = 0;
  if (moe != nullptr) {
    VPB_REQUIRE(moe->num_runs >= 1 && moe->num_runs <= VPB_MOE_MAX_RUNS && moe->image_begin && moe->fc2_w && moe->fc2_b,
                "forward: bad vpb_moe_runs (num_runs %d)", moe->num_runs);
    VPB_REQUIRE(moe->image_begin[0] == 0 && moe->image_begin[moe->num_runs] == images,
                "forward: vpb_moe_runs must cover the %d images of the batch exactly", images);
    uint8_t* sp = base + ws.moe_scratch;
    for (int r = 0; r < moe->num_runs; ++r) {
      const int i0 = moe->image_begin[r], i1 = moe->image_begin[r + 1];
      VPB_REQUIRE(i1 > i0, "forward: vpb_moe_runs: run %d is empty", r);
      runs[r] = MoeRun{i0 * T, (i1 - i0) * T, sp};
      if (int e = gemm_ln_scratch_init(sp, runs[r].rows, D, stream)) return e;
      sp += (gemm_ln_scratch_bytes(runs[r].rows, D) + 1023) / 1024 * 1024;
    }
    n_runs = moe->num_runs;
  }
  unsigned moe_epoch = 0;
  const bool fold = fold_enabled && w->fold != nullptr && moe == nullptr && parts > 0 && parts <= 10 && T % 64 == 0 &&
                    D % 8 == 0;
  void* stats = ln_scratch;
  const LnFoldIn ln_in{stats, nullptr, parts, part_cols, d.ln_eps};

  // PatchEmbed (vit.py:159-165) + pos embed (vit.py:320) + blocks[0].norm1
  if (int e = prof_run("im2col", stream, [&] { return im2col_patch16(img, patches, n, d.img_h, d.img_w, flip, stream); })) return e;
  if (int e = prof_run("gemm_patch_ln", stream, [&] {
        if (fold)
          return gemm_bf16_ln(patches, w->patch_w, rows, D, 768, EPI_POS_F32, w->patch_b, x, w->pos, T, nullptr, nullptr,
                              d.ln_eps, xn, nullptr, 0, 0, stream, nullptr, 0, stats);
        return gemm_bf16_ln(patches, w->patch_w, rows, D, 768, EPI_POS_F32, w->patch_b, x, w->pos, T, w->blocks[0].ln1_g,
                            w->blocks[0].ln1_b, d.ln_eps, xn, ln_scratch, ++epoch, 0, stream);
      }))
    return e;

  const float scale = 1.0f / sqrtf(static_cast<float>(hd));
  for (int l = 0; l < d.depth; ++l) {
    const vpb_block_weights& b = w->blocks[l];
    // x = x + proj(attn(LN1(x)))            (vit.py:138), then LN2(x) for the MLP
    if (int e = prof_run("gemm_qkv", stream, [&] {
          if (fold) {
            LnFoldIn in = ln_in;
            in.s = w->fold[l].qkv_s;
            return gemm_bf16(xn, w->fold[l].qkv_wf, rows, 3 * D, D, EPI_BIAS_BF16, w->fold[l].qkv_c, qkv, 3 * D, nullptr, 0, 0,
                             stream, &in);
          }
          return gemm_bf16(xn, b.qkv_w, rows, 3 * D, D, EPI_BIAS_BF16, b.qkv_b, qkv, 3 * D, nullptr, 0, 0, stream);
        }))
      return e;
    if (int e = prof_run("attention", stream, [&] { return attention_fwd(qkv, attn, images, T, d.num_heads, hd, scale, 0, stream); })) return e;
    if (int e = prof_run("gemm_proj_ln", stream, [&] {
          if (fold)
            return gemm_bf16_ln(attn, b.proj_w, rows, D, D, EPI_RESID_F32, b.proj_b, x, x, 0, nullptr, nullptr, d.ln_eps, xn,
                                nullptr, 0, 0, stream, nullptr, 0, stats);
          return gemm_bf16_ln(attn, b.proj_w, rows, D, D, EPI_RESID_F32, b.proj_b, x, x, 0, b.ln2_g, b.ln2_b, d.ln_eps, xn,
                              ln_scratch, ++epoch, 0, stream);
        }))
      return e;
    // x = x + fc2(gelu(fc1(LN2(x))))        (vit.py:139), then the next block's LN1 (or last_norm, vit.py:328)
    if (int e = prof_run("gemm_fc1", stream, [&] {
          if (fold) {
            LnFoldIn in = ln_in;
            in.s = w->fold[l].fc1_s;
            return gemm_bf16(xn, w->fold[l].fc1_wf, rows, d.mlp_hidden, D, EPI_GELU_BF16, w->fold[l].fc1_c, hidden,
                             d.mlp_hidden, nullptr, 0, 0, stream, &in);
          }
          return gemm_bf16(xn, b.fc1_w, rows, d.mlp_hidden, D, EPI_GELU_BF16, b.fc1_b, hidden, d.mlp_hidden, nullptr, 0, 0,
                           stream);
        }))
      return e;
    const float* ng = l + 1 < d.depth ? w->blocks[l + 1].ln1_g : w->last_g;
    const float* nb = l + 1 < d.depth ? w->blocks[l + 1].ln1_b : w->last_b;
    if (int e = prof_run("gemm_fc2_ln", stream, [&] {
          if (n_runs > 0) {      // ViTPose+: one launch per run of images, each with its dataset's fc2
            ++moe_epoch;
            for (int r = 0; r < n_runs; ++r) {
              const size_t r0 = static_cast<size_t>(runs[r].row0);
              const uint8_t* a = static_cast<const uint8_t*>(hidden) + r0 * d.mlp_hidden * 2;
              float* xr = x + r0 * D;
              void* xnr = static_cast<uint8_t*>(xn) + r0 * D * 2;
              if (int e2 = gemm_bf16_ln(a, moe->fc2_w[r * d.depth + l], runs[r].rows, D, d.mlp_hidden, EPI_RESID_F32,
                                        moe->fc2_b[r * d.depth + l], xr, xr, 0, ng, nb, d.ln_eps, xnr, runs[r].scratch,
                                        moe_epoch, 0, stream))
                return e2;
            }
            return 0;
          }
          if (fold && l + 1 < d.depth)
            return gemm_bf16_ln(hidden, b.fc2_w, rows, D, d.mlp_hidden, EPI_RESID_F32, b.fc2_b, x, x, 0, nullptr, nullptr,
                                d.ln_eps, xn, nullptr, 0, 0, stream, nullptr, 0, stats);
          return gemm_bf16_ln(hidden, b.fc2_w, rows, D, d.mlp_hidden, EPI_RESID_F32, b.fc2_b, x, x, 0, ng, nb, d.ln_eps, xn,
                              ln_scratch, ++epoch, 0, stream);
        }))
      return e;
  }
  // token-major [images, T, D] == NHWC [images, hp, wp, D] for the head
  void* feat = xn;
  if (features != nullptr)
    VPB_CHECK_CUDA(cudaMemcpyAsync(features, feat, static_cast<size_t>(rows) * D * 2, cudaMemcpyDeviceToDevice, stream));
  if (heatmaps == nullptr) return 0;

  const int K = d.num_keypoints;
  const bool split = flip && heatmaps_flipped != nullptr;   // flipped-pass maps go to their own buffer
  if (d.num_deconv > 0) {
    VPB_REQUIRE(d.final_kernel == 1, "forward: classic decoder expects a 1x1 final conv");
    const void* cur = feat;
    int ch = D, h = hp, wd = wp;
    void* bufs[2] = {base + ws.head_a, base + ws.head_b};
    for (int i = 0; i < d.num_deconv; ++i) {
      void* o = bufs[i & 1];
      if (int e = prof_run("deconv", stream, [&] { return deconv4x4s2_bn_relu(cur, w->deconv_w[i], w->deconv_scale[i], w->deconv_shift[i], o, images, h, wd,
                                      ch, d.deconv_channels[i], 0, stream); }))
        return e;
      cur = o;
      ch = d.deconv_channels[i];
      h *= 2;
      wd *= 2;
    }
    // final 1x1 conv (simple_head.py:132-139) as a GEMM over pixels with an NCHW fp32 epilogue
    // (two launches when the flipped pass goes to its own buffer)
    const int parts = split ? 2 : 1, per = images / parts;
    for (int part = 0; part < parts; ++part) {
      const void* a = static_cast<const uint8_t*>(cur) + static_cast<size_t>(part) * per * h * wd * ch * 2;
      float* o = part == 0 ? heatmaps : heatmaps_flipped;
      if (int e = prof_run("final_conv1x1", stream, [&] {
            return gemm_bf16(a, w->final_w, per * h * wd, K, ch, EPI_NCHW_F32, w->final_b, o, 0, nullptr, h * wd, 0, stream);
          }))
        return e;
    }
  } else {
    VPB_REQUIRE(d.final_kernel == 3 && d.upsample > 0, "forward: simple decoder expects upsample + 3x3 final conv");
    static int fused = -1;      // VPB_SIMPLE_FUSED=0: materialise the upsampled map and run the 3x3 implicit GEMM (A/B)
    if (fused < 0) {
      const char* e = getenv("VPB_SIMPLE_FUSED");
      fused = (e && atoi(e) == 0) ? 0 : 1;
    }
    if (fused && 9 * K <= 256 && (9 * T * (1 + d.upsample) + 2 * hp * d.upsample) * 4 <= 48 * 1024) {
      // nine 1x1 convolutions on the token grid as ONE GEMM (final_w [K][9 * D] viewed as [9K, D]: row k * 9 + t),
      // then the bilinear gather of the tap maps (elementwise.cu): the upsampled map is never materialised
      void* r = base + ws.head_a;                                         // relu(features) bf16 [rows, D]
      float* z = reinterpret_cast<float*>(base + ws.head_a + align_up(static_cast<size_t>(rows) * D * 2, 1024));
      if (int e = prof_run("relu", stream, [&] { return relu_bf16(feat, r, static_cast<long long>(rows) * D, stream); })) return e;
      if (int e = prof_run("simple_tap_gemm", stream, [&] {
            return gemm_bf16(r, w->final_w, rows, 9 * K, D, EPI_NCHW_F32, nullptr, z, 0, nullptr, T, 0, stream);
          }))
        return e;
      const int parts = split ? 2 : 1, per = images / parts;
      for (int part = 0; part < parts; ++part) {
        float* o = part == 0 ? heatmaps : heatmaps_flipped;
        const float* zp = z + static_cast<size_t>(part) * per * 9 * K * T;
        if (int e = prof_run("simple_gather", stream, [&] {
              return simple_head_gather(zp, w->final_b, o, per, K, hp, wp, d.upsample, stream);
            }))
          return e;
      }
      return 0;
    }
    void* up = base + ws.head_a;
    if (int e = prof_run("relu_upsample", stream, [&] { return relu_upsample_bilinear_nhwc(feat, up, images, hp, wp, D, d.upsample, stream); })) return e;
    const int parts = split ? 2 : 1, per = images / parts;
    const int h = hp * d.upsample, wd = wp * d.upsample;
    for (int part = 0; part < parts; ++part) {
      const void* a = static_cast<const uint8_t*>(up) + static_cast<size_t>(part) * per * h * wd * D * 2;
      float* o = part == 0 ? heatmaps : heatmaps_flipped;
      if (int e = prof_run("final_conv3x3", stream, [&] {
            return conv3x3_nchw_out(a, w->final_w, w->final_b, o, per, h, wd, D, K, 0, stream);
          }))
        return e;
    }
  }
  return 0;
}

int vpb_decode_heatmaps(const float* hm, const float* hm_flipped, const int32_t* flip_index, int shift_heatmap,
                        int N, int K, int H, int W, int mode, int kernel, int use_udp, int apply_transform,
                        const float* center, const float* scale, float* preds, float* maxvals, float* merged_out,
                        int32_t* argmax_out, void* stream) {
  cudaStream_t st = as_stream(stream);
  return prof_run("decode", st, [&] {
    return decode_heatmaps(hm, hm_flipped, flip_index, shift_heatmap, N, K, H, W, mode, kernel, use_udp,
                           apply_transform, center, scale, preds, maxvals, merged_out, argmax_out, st);
  });
}

int vpb_flip_back(const float* in, const int32_t* flip_index, float* out, int N, int K, int H, int W, int shift,
                  void* stream) {
  return flip_back(in, flip_index, out, N, K, H, W, shift, as_stream(stream));
}
int vpb_transform_preds(const float* coords, const float* center, const float* scale, float* out, int N, int K,
                        int W, int H, int use_udp, void* stream) {
  return transform_preds(coords, center, scale, out, N, K, W, H, use_udp, as_stream(stream));
}

int vpb_gemm_bf16(const void* A, const void* B, int M, int N, int K, int epilogue, const float* bias, void* out,
                  int ldo, const float* aux, int period, int max_ctas, void* stream) {
  return gemm_bf16(A, B, M, N, K, epilogue, bias, out, ldo, aux, period, max_ctas, as_stream(stream));
}
int vpb_gemm_bf16_atb_accum(const void* At, const void* Bt, int M, int N, int K, float* out, int ldo, void* stream) {
  return gemm_bf16_atb_accum(At, Bt, M, N, K, out, ldo, 0, as_stream(stream));
}
int vpb_gemm_bf16_atb_accum_ld(const void* At, int lda, const void* Bt, int ldb, int M, int N, int K, float* out,
                               int ldo, void* stream) {
  return gemm_bf16_atb_accum_ld(At, lda, Bt, ldb, M, N, K, out, ldo, 0, as_stream(stream));
}
int vpb_gemm_bf16_layernorm(const void* A, const void* B, int M, int N, int K, int epilogue, const float* bias,
                            float* out, const float* aux, int period, const float* gamma, const float* beta, float eps,
                            void* xn, void* scratch, size_t scratch_bytes, const float* row_scale, int rows_per_scale,
                            void* stream_) {
  cudaStream_t stream = as_stream(stream_);
  VPB_REQUIRE(M > 0 && N > 0, "gemm+layernorm: empty problem");
  VPB_REQUIRE(scratch != nullptr && scratch_bytes >= gemm_ln_scratch_bytes(M, N),
              "gemm+layernorm: scratch too small (%zu < %zu)", scratch_bytes, gemm_ln_scratch_bytes(M, N));
  if (int e = gemm_ln_scratch_init(scratch, M, N, stream)) return e;
  return gemm_bf16_ln(A, B, M, N, K, epilogue, bias, out, aux, period, gamma, beta, eps, xn, scratch, 1u, 0, stream,
                      row_scale, rows_per_scale);
}
int vpb_relu_bf16(const void* in, void* out, long long n, void* stream) { return relu_bf16(in, out, n, as_stream(stream)); }
int vpb_relu_bwd_bf16(const void* y, const void* dy, void* dx, long long n, void* stream) {
  return relu_bwd_bf16(y, dy, dx, n, as_stream(stream));
}
int vpb_simple_head_gather(const float* z, const float* bias, float* out, int images, int K, int h, int w, int factor,
                           void* stream) {
  return simple_head_gather(z, bias, out, images, K, h, w, factor, as_stream(stream));
}
int vpb_simple_head_gather_bwd(const float* dout, void* dz, int ldz, int images, int K, int h, int w, int factor,
                               void* stream) {
  return simple_head_gather_bwd(dout, dz, ldz, images, K, h, w, factor, as_stream(stream));
}
int vpb_gemm_bf16_gelu_save(const void* A, const void* B, int M, int N, int K, const float* bias, void* out, int ldo,
                            void* pre_out, void* stream) {
  const GemmTrainAux tr{pre_out, nullptr, nullptr};
  return gemm_bf16(A, B, M, N, K, EPI_GELU_BF16, bias, out, ldo, nullptr, 0, 0, as_stream(stream), nullptr, &tr);
}
int vpb_gemm_bf16_gelu_bwd(const void* A, const void* B, int M, int N, int K, const void* pre, void* out, int ldo,
                           float* colsum, void* stream) {
  const GemmTrainAux tr{nullptr, pre, colsum};
  return gemm_bf16(A, B, M, N, K, EPI_DGELU_BF16, nullptr, out, ldo, nullptr, 0, 0, as_stream(stream), nullptr, &tr);
}
int vpb_cast_f32_bf16_colsum(const float* in, void* out, int R, int C, const float* row_scale, int rows_per_scale,
                             float* colsum, void* stream) {
  return cast_f32_bf16_colsum(in, out, R, C, row_scale, rows_per_scale, colsum, as_stream(stream));
}
int vpb_fold_layernorm_linear(const float* W, const float* bias, const float* gamma, const float* beta, int N, int K,
                              void* Wf, float* s, float* c, void* stream) {
  return fold_layernorm_linear(W, bias, gamma, beta, N, K, Wf, s, c, as_stream(stream));
}
int vpb_gemm_stats_layout(int N, int* parts, int* part_cols) {
  VPB_REQUIRE(parts && part_cols, "gemm_stats_layout: null argument");
  *parts = N > 0 ? gemm_ln_parts(N) : 0;
  *part_cols = N > 0 ? gemm_ln_part_cols(N) : 0;
  return *parts > 0 ? 0 : -2;
}
size_t vpb_gemm_stats_bytes(int M, int N) { return M > 0 && N > 0 ? gemm_ln_stats_bytes(M, N) : 0; }
int vpb_gemm_bf16_resid_stats(const void* A, const void* B, int M, int N, int K, int epilogue, const float* bias,
                              float* out, const float* aux, int period, void* xb, void* stats, size_t stats_bytes,
                              const float* row_scale, int rows_per_scale, void* stream) {
  VPB_REQUIRE(M > 0 && N > 0, "gemm+stats: bad shape");
  VPB_REQUIRE(stats != nullptr && gemm_ln_parts(N) > 0 && stats_bytes >= gemm_ln_stats_bytes(M, N),
              "gemm+stats: statistics buffer too small (%zu < %zu)", stats_bytes, gemm_ln_stats_bytes(M, N));
  return gemm_bf16_ln(A, B, M, N, K, epilogue, bias, out, aux, period, nullptr, nullptr, 0.f, xb, nullptr, 0, 0,
                      as_stream(stream), row_scale, rows_per_scale, stats);
}
int vpb_gemm_bf16_lnfold(const void* A, const void* Wf, int M, int N, int K, int epilogue, const float* c,
                         const float* s, const void* stats, int parts, int part_cols, float eps, void* out, int ldo,
                         void* stream) {
  VPB_REQUIRE(stats != nullptr && s != nullptr && c != nullptr, "gemm+folded layernorm: null argument");
  const LnFoldIn in{stats, s, parts, part_cols, eps};
  return gemm_bf16(A, Wf, M, N, K, epilogue, c, out, ldo, nullptr, 0, 0, as_stream(stream), &in);
}
int vpb_gemm_layernorm_scratch_init(void* scratch, size_t scratch_bytes, int M, int N, void* stream) {
  VPB_REQUIRE(M > 0 && N > 0 && scratch != nullptr && scratch_bytes >= gemm_ln_scratch_bytes(M, N),
              "gemm+layernorm: scratch too small (%zu < %zu)", scratch_bytes, gemm_ln_scratch_bytes(M > 0 ? M : 1, N > 0 ? N : 1));
  return gemm_ln_scratch_init(scratch, M, N, as_stream(stream));
}
int vpb_gemm_bf16_layernorm_seq(const void* A, const void* B, int M, int N, int K, int epilogue, const float* bias,
                                float* out, const float* aux, int period, const float* gamma, const float* beta,
                                float eps, void* xn, void* scratch, size_t scratch_bytes, unsigned epoch,
                                const float* row_scale, int rows_per_scale, void* stream) {
  VPB_REQUIRE(M > 0 && N > 0, "gemm+layernorm: empty problem");
  VPB_REQUIRE(epoch >= 1u && scratch != nullptr && scratch_bytes >= gemm_ln_scratch_bytes(M, N),
              "gemm+layernorm: epoch must be >= 1 and the scratch %zu bytes", gemm_ln_scratch_bytes(M, N));
  return gemm_bf16_ln(A, B, M, N, K, epilogue, bias, out, aux, period, gamma, beta, eps, xn, scratch, epoch, 0,
                      as_stream(stream), row_scale, rows_per_scale);
}
size_t vpb_gemm_layernorm_scratch_bytes(int M, int N) { return M > 0 && N > 0 ? gemm_ln_scratch_bytes(M, N) : 0; }
int vpb_layernorm_bf16(const float* x, const float* gamma, const float* beta, void* y, int M, int D, float eps,
                       void* stream) {
  return layernorm_bf16(x, gamma, beta, y, M, D, eps, as_stream(stream));
}
int vpb_im2col_patch16(const float* img, void* patches, int n, int H, int W, int flip, void* stream) {
  return im2col_patch16(img, patches, n, H, W, flip, as_stream(stream));
}
int vpb_attention(const void* qkv, void* out, int n, int T, int heads, int head_dim, float scale, void* stream) {
  return attention_fwd(qkv, out, n, T, heads, head_dim, scale, 0, as_stream(stream));
}
int vpb_deconv4x4s2_bn_relu(const void* in, const void* wphase, const float* scale, const float* shift, void* out,
                            int n, int h, int w, int cin, int cout, void* stream) {
  return deconv4x4s2_bn_relu(in, wphase, scale, shift, out, n, h, w, cin, cout, 0, as_stream(stream));
}
int vpb_conv3x3_nchw(const void* in, const void* w9, const float* bias, float* out, int n, int h, int w, int cin,
                     int cout, void* stream) {
  return conv3x3_nchw_out(in, w9, bias, out, n, h, w, cin, cout, 0, as_stream(stream));
}
int vpb_relu_upsample_nhwc(const void* in, void* out, int n, int h, int w, int C, int factor, void* stream) {
  return relu_upsample_bilinear_nhwc(in, out, n, h, w, C, factor, as_stream(stream));
}
int vpb_tokens_to_nchw_f32(const void* tokens, float* out, int n, int T, int D, void* stream) {
  return tokens_to_nchw_f32(tokens, out, n, T, D, as_stream(stream));
}

int vpb_joints_mse_loss(const float* output, const float* target, const float* target_weight, int N, int K, int HW,
                        float loss_weight, float* loss, float* grad_output, void* stream) {
  return joints_mse_loss(output, target, target_weight, N, K, HW, loss_weight, loss, grad_output, as_stream(stream));
}
int vpb_grad_sq_norm_accumulate(const float* grad, long long n, float* sq_norm_accum, void* stream) {
  return grad_sq_norm_accumulate(grad, n, sq_norm_accum, as_stream(stream));
}
int vpb_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                   float beta1, float beta2, float eps, float weight_decay, int step, const float* sq_norm,
                   float max_norm, void* stream) {
  return adamw_step(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, sq_norm, max_norm,
                    as_stream(stream));
}

int vpb_adamw_multi(const vpb_tensor_entry* entries, const int32_t* chunk_start, int n, int total_chunks, float beta1,
                    float beta2, float eps, float* sq_norm, float max_norm, void* stream) {
  return adamw_multi(entries, chunk_start, n, total_chunks, beta1, beta2, eps, sq_norm, max_norm, as_stream(stream));
}

int vpb_cast_transpose_multi(const vpb_cast_entry* entries, const int32_t* tile_start, int n, int total_tiles,
                             void* stream) {
  cudaStream_t st = as_stream(stream);
  return prof_run("cast_transpose_multi", st, [&] { return cast_transpose_multi(entries, tile_start, n, total_tiles, st); });
}

int vpb_pose_pck_accuracy(const float* pred, const float* gt, const float* weight, int N, int K, float norm0,
                          float norm1, float thr, float* acc, float* avg, int32_t* cnt, void* stream) {
  return pose_pck_accuracy(pred, gt, weight, N, K, norm0, norm1, thr, acc, avg, cnt, as_stream(stream));
}

int vpb_oks_nms(const float* kpts, const double* areas, const double* box_scores, const int32_t* group_start, int G,
                int K, int max_group, const double* var, double thr, int use_vis, double vis_thr, int rescore, int soft,
                int max_dets, double* scores_out, int32_t* keep, int32_t* keep_count, void* stream) {
  cudaStream_t st = as_stream(stream);
  return prof_run("oks_nms", st, [&] {
    return oks_nms(kpts, areas, box_scores, group_start, G, K, max_group, var, thr, use_vis, vis_thr, rescore, soft,
                   max_dets, scores_out, keep, keep_count, st);
  });
}

// ---- backward-pass operators of the training step ----
int vpb_transpose_bf16(const void* in, void* out, int R, int C, int batch, void* stream) {
  return transpose_bf16(in, out, R, C, batch, as_stream(stream));
}
int vpb_cast_f32_bf16(const float* in, void* out, long long n, const float* row_scale, int row_len,
                      int rows_per_scale, void* stream) {
  return cast_f32_bf16(in, out, n, as_stream(stream), row_scale, row_len, rows_per_scale);
}
int vpb_colsum_accumulate(const void* in, int is_f32, int R, int C, float* out, void* stream) {
  return colsum_accumulate(in, is_f32, R, C, out, as_stream(stream));
}
int vpb_gelu_fwd_bf16(const void* pre, void* out, long long n, void* stream) {
  return gelu_fwd_bf16(pre, out, n, as_stream(stream));
}
int vpb_gelu_bwd_bf16(const void* pre, const void* dh, void* dpre, long long n, void* stream) {
  return gelu_bwd_bf16(pre, dh, dpre, n, as_stream(stream));
}
int vpb_layernorm_bwd(const float* x, const float* gamma, const void* dy, float* dx_accum, float* dgamma,
                      float* dbeta, int M, int D, float eps, void* stream) {
  return layernorm_bwd(x, gamma, dy, dx_accum, dgamma, dbeta, M, D, eps, as_stream(stream));
}
int vpb_attention_lse(const void* qkv, void* out, float* lse, int n, int T, int heads, int head_dim, float scale,
                      void* stream) {
  return attention_fwd(qkv, out, n, T, heads, head_dim, scale, 0, as_stream(stream), lse);
}
int vpb_attention_bwd(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, int n, int T,
                      int heads, int head_dim, float scale, void* stream) {
  return attention_bwd(qkv, out, lse, dout, dqkv, n, T, heads, head_dim, scale, as_stream(stream));
}
int vpb_attention_bwd_bias(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, float* dbias,
                           int n, int T, int heads, int head_dim, float scale, void* stream) {
  return attention_bwd(qkv, out, lse, dout, dqkv, n, T, heads, head_dim, scale, as_stream(stream), dbias);
}
int vpb_deconv4x4s2_raw(const void* in, const void* wphase, void* out, int n, int h, int w, int cin, int cout,
                        const float* ones, const float* zeros, void* stream) {
  return deconv4x4s2_affine(in, wphase, ones, zeros, out, n, h, w, cin, cout, 0, 0, as_stream(stream));
}
int vpb_bn_train_stats(const void* raw, long long rows, int C, float eps, float momentum, float* sum_sumsq_scratch,
                       float* mean, float* rstd, float* running_mean, float* running_var, void* stream_) {
  cudaStream_t stream = as_stream(stream_);
  VPB_REQUIRE(rows > 0 && rows < (1ll << 31) && C > 0, "bn_train_stats: bad shape");
  VPB_REQUIRE((reinterpret_cast<uintptr_t>(sum_sumsq_scratch) & 7) == 0, "bn_train_stats: scratch must be 8-byte aligned");
  double* acc = reinterpret_cast<double*>(sum_sumsq_scratch);     // 2*C fp64 accumulators
  VPB_CHECK_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 2 * C, stream));
  if (int e = colsum_sq_accumulate(raw, static_cast<int>(rows), C, acc, acc + C, stream)) return e;
  return bn_finalize(acc, acc + C, mean, rstd, running_mean, running_var, C, rows, eps, momentum, stream);
}
int vpb_bn_relu_fwd(const void* raw, void* act, const float* mean, const float* rstd, const float* gamma,
                    const float* beta, long long rows, int C, void* stream) {
  return bn_relu_fwd(raw, act, mean, rstd, gamma, beta, rows, C, as_stream(stream));
}
int vpb_bn_relu_bwd(const void* raw, const void* dact, void* draw, const float* mean, const float* rstd,
                    const float* gamma, const float* beta, float* dgamma, float* dbeta, float* scratch, long long rows,
                    int C, void* stream_) {
  cudaStream_t stream = as_stream(stream_);
  VPB_REQUIRE(rows > 0 && rows < (1ll << 31), "bn_relu_bwd: bad shape");
  VPB_REQUIRE(scratch != nullptr && (reinterpret_cast<uintptr_t>(scratch) & 7) == 0, "bn_relu_bwd: scratch (6*C floats, 8-byte aligned)");
  double* acc = reinterpret_cast<double*>(scratch);          // 2*C fp64 accumulators
  float* sums = scratch + 4 * C;                             // 2*C fp32: this layer's sum dy', sum dy' * xhat
  if (int e = bn_relu_bwd_reduce(raw, dact, mean, rstd, gamma, beta, static_cast<int>(rows), C, acc, sums, dbeta, dgamma,
                                 stream))
    return e;
  return bn_relu_bwd(raw, dact, draw, mean, rstd, gamma, beta, sums, sums + C, rows, C, stream);
}
// BatchNorm2d in EVAL mode (running statistics passed as mean / rstd) + ReLU: the statistics do not depend on the
// batch, so draw = gamma * rstd * dy' without the two batch-mean correction terms; dgamma / dbeta as in training mode
int vpb_bn_relu_bwd_eval(const void* raw, const void* dact, void* draw, const float* mean, const float* rstd,
                         const float* gamma, const float* beta, float* dgamma, float* dbeta, float* scratch,
                         long long rows, int C, void* stream_) {
  cudaStream_t stream = as_stream(stream_);
  VPB_REQUIRE(rows > 0 && rows < (1ll << 31), "bn_relu_bwd: bad shape");
  VPB_REQUIRE(scratch != nullptr && (reinterpret_cast<uintptr_t>(scratch) & 7) == 0, "bn_relu_bwd: scratch (6*C floats, 8-byte aligned)");
  double* acc = reinterpret_cast<double*>(scratch);
  float* sums = scratch + 4 * C;
  if (int e = bn_relu_bwd_reduce(raw, dact, mean, rstd, gamma, beta, static_cast<int>(rows), C, acc, sums, dbeta, dgamma,
                                 stream))
    return e;
  VPB_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, stream));
  return bn_relu_bwd(raw, dact, draw, mean, rstd, gamma, beta, sums, sums + C, rows, C, stream);
}
int vpb_nchw_f32_to_rows_bf16(const float* in, void* out, int n, int K, int P, int Kp, void* stream) {
  return nchw_f32_to_rows_bf16(in, out, n, K, P, Kp, as_stream(stream));
}
int vpb_deconv_gather_x(const void* x, void* out, int n, int h, int w, int cin, void* stream) {
  return deconv_gather_x(x, out, n, h, w, cin, as_stream(stream));
}
int vpb_deconv_gather_dy(const void* dy, void* out, int n, int h, int w, int cout, void* stream) {
  return deconv_gather_dy(dy, out, n, h, w, cout, as_stream(stream));
}
int vpb_deconv_phase_dy(const void* dy, void* out, int n, int h, int w, int cout, void* stream) {
  return deconv_phase_dy(dy, out, n, h, w, cout, as_stream(stream));
}

int vpb_deconv_pack_weight(const float* w, void* wp, void* wd, int cin, int cout, void* stream) {
  return deconv_pack_weight(w, wp, wd, cin, cout, as_stream(stream));
}
int vpb_deconv_unpack_wgrad(const float* dwp, float* dw, int cin, int cout, void* stream) {
  return deconv_unpack_wgrad(dwp, dw, cin, cout, as_stream(stream));
}

int vpb_warp_affine_normalize(const unsigned char* const* src_ptrs, const int32_t* src_hw, const double* inv_mats,
                              int n, int out_h, int out_w, const float* mean3, const float* std3, float* out,
                              void* stream) {
  cudaStream_t st = as_stream(stream);
  return prof_run("warp_affine_normalize", st, [&] {
    return warp_affine_normalize(src_ptrs, src_hw, inv_mats, n, out_h, out_w, mean3, std3, out, st);
  });
}

}  // extern "C"
