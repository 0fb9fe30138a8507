// Model-level C-ABI entry points: the layer loops of the two towers in C++, for hosts that are not Python.
//
//   omc_vit_forward      InternVITVisionTower.forward + feature_select + (pixel shuffle) + mm_projector
//                        (omchat/model/multimodal_encoder/internVIT_encoder.py:35-56, intern_vit_6b/modeling_intern_vit.py:
//                        90-102,138-222,268-279, multimodal_projector/builder.py:54-61 = encode_images, omchat_arch.py:50-53)
//   omc_decoder_prefill  Qwen2Model.forward over packed sequences + final norm + lm_head on each sequence's last row
//                        (transformers modeling_qwen2.py:280-310,353-414,470-472), filling the paged KV cache
//
// Both are plain sequences of the op-level launches declared above them in include/omchat_b200.h - the same kernels in the
// same order as omchat_b200/model/vision.py and decoder.py issue them, so results are bit-identical to the Python host path
// (tests/test_model_capi_gpu.py) - with caller-provided workspaces, no allocation, no host synchronisation: a call can be
// captured in a CUDA graph. The decode step already has its model-level entry (omc_decode_step).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "omc_internal.h"

using namespace omc;

namespace {
constexpr int kSsqParts = 32;  // capacity (in N tiles) of a sums-of-squares buffer: hidden 3200 / 3584 at 128-wide tiles = 25 / 28
inline long long align256(long long x) { return (x + 255) & ~255LL; }
struct Carver {
  uint8_t* p;
  void* take(long long bytes) {
    void* r = p;
    p += align256(bytes);
    return r;
  }
};
#define OMC_TRY(expr)          \
  do {                         \
    const int rc_ = (expr);    \
    if (rc_ != OMC_OK) return rc_; \
  } while (0)
}  // namespace

// ------------------------------------------------------------------------------------------------ vision tower + projector
static void vit_sizes(const omc_vit_desc* d, int n, long long* rows, long long* sizes) {
  const long long P = (long long)(d->image_size / d->patch_size) * (d->image_size / d->patch_size);
  *rows = (long long)n * (P + 1);
  const long long G = d->image_size / d->patch_size, down = d->pixel_shuffle_down;
  const long long L = (G / down) * (G / down);
  sizes[0] = (long long)n * P * d->patch_k * 2;                  // im2col columns
  sizes[1] = (long long)n * P * d->hidden * 2;                   // patch embeddings
  sizes[2] = *rows * d->hidden * 2;                              // residual stream h
  sizes[3] = *rows * d->hidden * 2;                              // normed rows
  // attention width: heads * 128 when the caller zero-padded narrower heads (attn_head_dim = 128), else hidden
  const long long Ca = d->attn_head_dim == 128 ? (long long)d->heads * 128 : (long long)d->hidden;
  sizes[4] = *rows * 3 * Ca * 2;                                 // qkv
  sizes[5] = *rows * Ca * 2;                                     // attention output
  sizes[6] = *rows * d->inter * 2;                               // MLP activation
  sizes[7] = (long long)n * L * d->hidden * down * down * 2;     // selected (+ shuffled) features
  sizes[8] = (long long)n * L * d->proj_hidden * 2;              // projector hidden
  sizes[9] = ((long long)n + 1) * 4;                             // cu_seqlens
  sizes[10] = d->norm_folded ? 2 * kSsqParts * *rows * 4 : 0;    // two [parts][rows] fp32 sums-of-squares buffers
}

extern "C" long long omc_vit_workspace_bytes(const omc_vit_desc* d, int max_crops) {
  if (d == nullptr || max_crops <= 0 || d->patch_size <= 0) return -1;
  long long rows, s[11], total = 0;
  vit_sizes(d, max_crops, &rows, s);
  for (int i = 0; i < 11; ++i) total += align256(s[i]);
  return total + 256;
}

__global__ void iota_scaled_kernel(int32_t* out, int n, int step) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i * step;
}

extern "C" int omc_vit_forward(const omc_vit_desc* d, const void* pixels, int pixels_are_f32, int n_crops, void* workspace,
                               void* feats_out, void* stream) {
  if (d == nullptr || pixels == nullptr || workspace == nullptr || feats_out == nullptr)
    return set_error(OMC_ERR_ARG, "omc_vit_forward: null argument");
  if (n_crops <= 0) return OMC_OK;
  if (d->image_size % d->patch_size != 0 || d->heads <= 0 || d->hidden % d->heads != 0)
    return set_error(OMC_ERR_SHAPE, "omc_vit_forward: the image must be a whole number of patches, hidden a multiple of heads");
  const int Dh = d->hidden / d->heads;
  const bool padded = d->attn_head_dim == 128 && Dh < 128;  // heads zero-padded to 128 by the caller
  if (d->attn_head_dim != 0 && d->attn_head_dim != 128) return set_error(OMC_ERR_ARG, "omc_vit_forward: attn_head_dim must be 0 or 128");
  if (!padded && Dh != 128 && Dh != 64)
    return set_error(OMC_ERR_SHAPE, "omc_vit_forward: head_dim must be 128 or 64 (or zero-padded to attn_head_dim = 128)");
  if (padded && d->qk_norm) return set_error(OMC_ERR_SHAPE, "omc_vit_forward: no QK-norm over zero-padded heads");
  const int Da = padded ? 128 : Dh;  // head width the attention kernel runs at
  const bool layer_norm = d->norm_type == 1;
  if (d->norm_type != 0 && d->norm_type != 1) return set_error(OMC_ERR_ARG, "omc_vit_forward: norm_type must be 0 (rms) or 1 (layer)");
  if (layer_norm && (d->norm_folded || d->norm1_b == nullptr || d->norm2_b == nullptr))
    return set_error(OMC_ERR_ARG, "omc_vit_forward: layer_norm needs norm1_b / norm2_b and norm_folded = 0");
  const int G = d->image_size / d->patch_size, P = G * G, C = d->hidden, S = P + 1;
  const int down = d->pixel_shuffle_down;
  if (down < 1 || G % down != 0) return set_error(OMC_ERR_SHAPE, "omc_vit_forward: bad pixel_shuffle_down");
  long long rows, s[11];
  vit_sizes(d, n_crops, &rows, s);
  Carver cv{static_cast<uint8_t*>(workspace) + ((256 - (reinterpret_cast<uintptr_t>(workspace) & 255)) & 255)};
  void* cols = cv.take(s[0]);
  void* patch = cv.take(s[1]);
  void* h = cv.take(s[2]);
  void* xn = cv.take(s[3]);
  __nv_bfloat16* qkv = static_cast<__nv_bfloat16*>(cv.take(s[4]));
  void* attn = cv.take(s[5]);
  void* act = cv.take(s[6]);
  void* sel = cv.take(s[7]);
  void* ph = cv.take(s[8]);
  int32_t* cu = static_cast<int32_t*>(cv.take(s[9]));
  float* ssq_a = static_cast<float*>(cv.take(s[10] / 2));
  float* ssq_b = static_cast<float*>(cv.take(s[10] / 2));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  iota_scaled_kernel<<<(n_crops + 1 + 127) / 128, 128, 0, st>>>(cu, n_crops + 1, S);
  OMC_TRY(check_launch("iota"));
  // embeddings: conv14/14 as im2col + GEMM, then CLS + position embeddings (modeling_intern_vit.py:90-102)
  OMC_TRY(omc_vit_im2col(pixels, pixels_are_f32, cols, d->patch_k, n_crops, d->image_size, d->image_size, stream));
  OMC_TRY(omc_gemm_bf16(cols, d->patch_k, d->patch_w, d->patch_k, patch, C, n_crops * P, C, d->patch_k, d->patch_b, nullptr,
                        nullptr, 0, OMC_EPI_NONE, 0, 0, stream));
  OMC_TRY(omc_vit_assemble(patch, d->cls, d->pos, h, n_crops, P, C, stream));
  const float scale = 1.0f / sqrtf((float)Dh);
  const int M = (int)rows;
  const int Ca = d->heads * Da;  // attention width
  auto qkv_bias = [&](int li) -> const void* { return d->qkv_b != nullptr ? d->qkv_b[li] : nullptr; };
  auto pre_norm = [&](const void* const* w, const void* const* b, int li) {
    return layer_norm ? omc_layernorm(h, C, w[li], b[li], xn, C, M, C, d->eps, stream)
                      : omc_rmsnorm(h, C, w[li], xn, C, M, C, d->eps, stream);
  };
  const bool fold = d->norm_folded != 0;
  int parts_a = 1, parts_b = 1;  // sums-of-squares partials the last producer of each buffer wrote
  if (fold) OMC_TRY(omc_row_ssq_rows(h, C, rows, C, ssq_a, stream));
  auto norm_in = [&](const float* ssq, int parts) {
    omc_gemm_norm nf{};
    nf.ssq_in = ssq; nf.ssq_in_ld = rows; nf.ssq_in_parts = parts; nf.norm_dim = C; nf.eps = d->eps;
    return nf;
  };
  auto norm_out = [&](float* ssq) {
    omc_gemm_norm nf{};
    nf.ssq_out = ssq; nf.ssq_out_ld = rows; nf.ssq_out_max_parts = kSsqParts;
    return nf;
  };
  for (int li = 0; li < d->n_layers; ++li) {
    // h += ls1 * proj(attn(qk_norm(qkv(norm1(h)))));  h += ls2 * fc2(gelu(fc1(norm2(h))))   (:138-155, 187-191, 218-220)
    if (fold) {
      omc_gemm_norm nf = norm_in(ssq_a, parts_a);
      OMC_TRY(omc_gemm_bf16_norm(h, C, d->qkv_w[li], C, qkv, 3LL * Ca, M, 3 * Ca, C, qkv_bias(li), nullptr, nullptr, 0, OMC_EPI_NONE,
                                 0, 0, &nf, stream));
    } else {
      OMC_TRY(pre_norm(d->norm1, d->norm1_b, li));
      OMC_TRY(omc_gemm_bf16(xn, C, d->qkv_w[li], C, qkv, 3LL * Ca, M, 3 * Ca, C, qkv_bias(li), nullptr, nullptr, 0, OMC_EPI_NONE, 0,
                            0, stream));
    }
    if (d->qk_norm) {
      if (fold) {
        OMC_TRY(omc_rmsnorm_pair(qkv, 3LL * C, d->q_norm[li], d->k_norm[li], M, C, d->eps, stream));
      } else {
        OMC_TRY(omc_rmsnorm(qkv, 3LL * C, d->q_norm[li], qkv, 3LL * C, M, C, d->eps, stream));
        OMC_TRY(omc_rmsnorm(qkv + C, 3LL * C, d->k_norm[li], qkv + C, 3LL * C, M, C, d->eps, stream));
      }
    }
    OMC_TRY(omc_attention_fwd_hd(qkv, 3LL * Ca, qkv + Ca, 3LL * Ca, qkv + 2 * Ca, 3LL * Ca, attn, Ca, cu, n_crops, S, rows, d->heads,
                                 d->heads, Da, 0, scale, stream));
    if (fold) {
      omc_gemm_norm no = norm_out(ssq_b);
      OMC_TRY(omc_gemm_bf16_norm(attn, Ca, d->proj_w[li], Ca, h, C, M, C, Ca, d->proj_b[li], d->ls1[li], h, C, OMC_EPI_RES, 0, 0, &no, stream));
      parts_b = no.ssq_out_parts;
      omc_gemm_norm nf = norm_in(ssq_b, parts_b);
      OMC_TRY(omc_gemm_bf16_norm(h, C, d->fc1_w[li], C, act, d->inter, M, d->inter, C, d->fc1_b[li], nullptr, nullptr, 0,
                                 OMC_EPI_GELU, 0, 0, &nf, stream));
      omc_gemm_norm no2 = norm_out(ssq_a);
      OMC_TRY(omc_gemm_bf16_norm(act, d->inter, d->fc2_w[li], d->inter, h, C, M, C, d->inter, d->fc2_b[li], d->ls2[li], h, C,
                                 OMC_EPI_RES, 0, 0, &no2, stream));
      parts_a = no2.ssq_out_parts;
    } else {
      OMC_TRY(omc_gemm_bf16(attn, Ca, d->proj_w[li], Ca, h, C, M, C, Ca, d->proj_b[li], d->ls1[li], h, C, OMC_EPI_RES, 0, 0, stream));
      OMC_TRY(pre_norm(d->norm2, d->norm2_b, li));
      OMC_TRY(omc_gemm_bf16(xn, C, d->fc1_w[li], C, act, d->inter, M, d->inter, C, d->fc1_b[li], nullptr, nullptr, 0, OMC_EPI_GELU,
                            0, 0, stream));
      OMC_TRY(omc_gemm_bf16(act, d->inter, d->fc2_w[li], d->inter, h, C, M, C, d->inter, d->fc2_b[li], d->ls2[li], h, C, OMC_EPI_RES,
                            0, 0, stream));
    }
  }
  // feature select 'patch' (+ pixel shuffle), then the mlp2x_gelu projector (internVIT_encoder.py:35-43, builder.py:54-61)
  OMC_TRY(omc_select_pixel_shuffle(h, sel, n_crops, G, C, down, stream));
  const int L = (G / down) * (G / down), Cin = C * down * down, H = d->proj_hidden;
  OMC_TRY(omc_gemm_bf16(sel, Cin, d->p_w0, Cin, ph, H, n_crops * L, H, Cin, d->p_b0, nullptr, nullptr, 0, OMC_EPI_GELU, 0, 0, stream));
  OMC_TRY(omc_gemm_bf16(ph, H, d->p_w2, H, feats_out, H, n_crops * L, H, H, d->p_b2, nullptr, nullptr, 0, OMC_EPI_NONE, 0, 0, stream));
  return OMC_OK;
}

// ------------------------------------------------------------------------------------------------ decoder prefill
extern "C" long long omc_decoder_prefill_workspace_bytes(const omc_decode_desc* d, int T, int n_seq) {
  if (d == nullptr || T <= 0 || n_seq <= 0) return -1;
  const long long C = d->hidden, qw = (long long)(d->q_heads + 2 * d->kv_heads) * 128;
  return align256((long long)T * C * 2) + align256((long long)T * qw * 2) + align256((long long)T * d->q_heads * 128 * 2) +
         align256((long long)T * d->inter * 2) + 2 * align256((long long)n_seq * C * 2) +
         2 * align256((long long)kSsqParts * T * 4) + 256;
}

extern "C" int omc_decoder_prefill(const omc_decode_desc* d, const float* inv_freq, void* embeds, const int32_t* pos_ids,
                                   const int32_t* seq_ids, const int32_t* cu_seqlens, int n_seq, int T, int max_len,
                                   const int64_t* last_rows, void* workspace, float* last_logits, int norm_folded,
                                   void* stream) {
  if (d == nullptr || inv_freq == nullptr || embeds == nullptr || workspace == nullptr)
    return set_error(OMC_ERR_ARG, "omc_decoder_prefill: null argument");
  if (T <= 0 || n_seq <= 0) return OMC_OK;
  const int C = d->hidden, Hq = d->q_heads, Hkv = d->kv_heads, I = d->inter;
  const long long qw = (long long)(Hq + 2 * Hkv) * 128;
  Carver cv{static_cast<uint8_t*>(workspace) + ((256 - (reinterpret_cast<uintptr_t>(workspace) & 255)) & 255)};
  void* xn = cv.take((long long)T * C * 2);
  __nv_bfloat16* qkv = static_cast<__nv_bfloat16*>(cv.take((long long)T * qw * 2));
  void* attn = cv.take((long long)T * Hq * 128 * 2);
  void* act = cv.take((long long)T * I * 2);
  void* hl = cv.take((long long)n_seq * C * 2);
  void* hn = cv.take((long long)n_seq * C * 2);
  float* ssq_a = static_cast<float*>(cv.take((long long)kSsqParts * T * 4));
  float* ssq_b = static_cast<float*>(cv.take((long long)kSsqParts * T * 4));
  void* h = embeds;  // the residual stream is updated in place
  const bool fold = norm_folded != 0;
  int parts_a = 1, parts_b = 1;
  if (fold) OMC_TRY(omc_row_ssq_rows(h, C, T, C, ssq_a, stream));
  auto norm_in = [&](const float* ssq, int parts) {
    omc_gemm_norm nf{};
    nf.ssq_in = ssq; nf.ssq_in_ld = T; nf.ssq_in_parts = parts; nf.norm_dim = C; nf.eps = d->eps;
    return nf;
  };
  auto norm_out = [&](float* ssq) {
    omc_gemm_norm nf{};
    nf.ssq_out = ssq; nf.ssq_out_ld = T; nf.ssq_out_max_parts = kSsqParts;
    return nf;
  };
  for (int li = 0; li < d->n_layers; ++li) {
    // Qwen2DecoderLayer.forward modeling_qwen2.py:280-310
    __nv_bfloat16* pool = static_cast<__nv_bfloat16*>(d->kv_pool) + (long long)li * d->kv_layer_stride;
    if (fold) {
      omc_gemm_norm nf = norm_in(ssq_a, parts_a);
      OMC_TRY(omc_gemm_bf16_norm(h, C, d->qkv_w[li], C, qkv, qw, T, (int)qw, C, d->qkv_b[li], nullptr, nullptr, 0, OMC_EPI_NONE, 0, 0,
                                 &nf, stream));
    } else {
      OMC_TRY(omc_rmsnorm(h, C, d->ln1[li], xn, C, T, C, d->eps, stream));
      OMC_TRY(omc_gemm_bf16(xn, C, d->qkv_w[li], C, qkv, qw, T, (int)qw, C, d->qkv_b[li], nullptr, nullptr, 0, OMC_EPI_NONE, 0, 0, stream));
    }
    OMC_TRY(omc_rope_kv_store(qkv, qw, pos_ids, seq_ids, T, Hq, Hkv, inv_freq, pool, d->block_table, d->max_pages, d->page_size, stream));
    OMC_TRY(omc_attention_fwd(qkv, qw, qkv + (long long)Hq * 128, qw, qkv + (long long)(Hq + Hkv) * 128, qw, attn, (long long)Hq * 128,
                              cu_seqlens, n_seq, max_len, T, Hq, Hkv, 1, d->attn_scale, stream));
    if (fold) {
      omc_gemm_norm no = norm_out(ssq_b);
      OMC_TRY(omc_gemm_bf16_norm(attn, (long long)Hq * 128, d->o_w[li], (long long)Hq * 128, h, C, T, C, Hq * 128, nullptr, nullptr, h,
                                 C, OMC_EPI_RES, 0, 0, &no, stream));
      parts_b = no.ssq_out_parts;
      omc_gemm_norm nf = norm_in(ssq_b, parts_b);
      OMC_TRY(omc_gemm_bf16_norm(h, C, d->gate_up_w[li], C, act, I, T, 2 * I, C, nullptr, nullptr, nullptr, 0, OMC_EPI_SWIGLU, 0, 0,
                                 &nf, stream));
      omc_gemm_norm no2 = norm_out(ssq_a);
      OMC_TRY(omc_gemm_bf16_norm(act, I, d->down_w[li], I, h, C, T, C, I, nullptr, nullptr, h, C, OMC_EPI_RES, 0, 0, &no2, stream));
      parts_a = no2.ssq_out_parts;
    } else {
      OMC_TRY(omc_gemm_bf16(attn, (long long)Hq * 128, d->o_w[li], (long long)Hq * 128, h, C, T, C, Hq * 128, nullptr, nullptr, h, C,
                            OMC_EPI_RES, 0, 0, stream));
      OMC_TRY(omc_rmsnorm(h, C, d->ln2[li], xn, C, T, C, d->eps, stream));
      OMC_TRY(omc_gemm_bf16(xn, C, d->gate_up_w[li], C, act, I, T, 2 * I, C, nullptr, nullptr, nullptr, 0, OMC_EPI_SWIGLU, 0, 0, stream));
      OMC_TRY(omc_gemm_bf16(act, I, d->down_w[li], I, h, C, T, C, I, nullptr, nullptr, h, C, OMC_EPI_RES, 0, 0, stream));
    }
  }
  if (last_logits != nullptr && last_rows != nullptr) {
    // final norm + lm_head on each sequence's last position (modeling_qwen2.py:411,470-472); the row gather is the
    // embedding-lookup kernel with the residual stream as its table
    OMC_TRY(omc_embed_lookup(last_rows, n_seq, h, C, hl, C, T, stream));
    OMC_TRY(omc_rmsnorm(hl, C, d->final_norm, hn, C, n_seq, C, d->eps, stream));
    OMC_TRY(omc_gemm_bf16(hn, C, d->lm_head, C, last_logits, d->vocab, n_seq, d->vocab, C, nullptr, nullptr, nullptr, 0,
                          OMC_EPI_NONE, 1, 0, stream));
  }
  return OMC_OK;
}
