/* omchat_b200 — C-ABI of the B200-native OmChat multimodal forward pass.
 *
 * The reference (om-ai-lab/OmChat) has no FFI: its hot path is Python calling torch.nn modules. This header is the
 * boundary a maintainer would bind instead (ctypes stub in INTEGRATION.md). Every entry point replaces one group of
 * reference call sites (cited per function, paths relative to the reference root). Conventions:
 *   - plain device pointers + sizes, no torch types; bf16 = raw uint16 storage; row-major; leading dims in ELEMENTS
 *   - every function is asynchronous on `stream` (a cudaStream_t passed as void*), never synchronises, never allocates
 *     device memory, never frees caller memory; outputs go to caller-allocated buffers
 *   - returns 0 on success or a negative OMC_ERR_* code; omc_last_error() gives the message (thread-local)
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns OMC_ERR_CUDA
 */
#ifndef OMCHAT_B200_H_
#define OMCHAT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OMC_OK 0
#define OMC_ERR_ARG (-1)    /* invalid argument / unsupported mode */
#define OMC_ERR_SHAPE (-2)  /* shape constraint violated (cf. ValueError in modeling_intern_vit.py:329-338) */
#define OMC_ERR_ALIGN (-3)  /* pointer / leading-dimension alignment */
#define OMC_ERR_CUDA (-4)   /* CUDA runtime error (message holds cudaGetErrorString) */
#define OMC_ERR_DRIVER (-5) /* driver entry point (tensor-map encode) unavailable/failed */

/* GEMM epilogues */
#define OMC_EPI_NONE 0   /* Y = X W^T (+ bias) */
#define OMC_EPI_GELU 1   /* Y = gelu_erf(X W^T + bias) */
#define OMC_EPI_RES 2    /* Y = res + scale * (X W^T + bias)   (scale, bias optional) */
#define OMC_EPI_SWIGLU 3 /* W rows alternate gate_i, up_i (row 2i, 2i+1): Y[:, i] = silu(gate_i) * up_i, Y is [M, N/2] */

const char* omc_last_error(void);
int omc_version(void);
/* number of SMs of the current device (148 on B200), <0 on error */
int omc_num_sms(void);

/* ---- dense linear layers: tcgen05/TMEM/TMA GEMM -------------------------------------------------------------------
 * Y[M,N] = X[M,K] W[N,K]^T with fused epilogue. Replaces nn.Linear at intern_vit_6b/modeling_intern_vit.py:124
 * (qkv), :136 (proj, with ls1+residual of :218), :183-190 (fc1+GELU, fc2 with ls2+residual of :220), the conv
 * patch-embed :73-75,92 after omc_vit_im2col, multimodal_projector/builder.py:57-61, and the Qwen2 q/k/v/o/gate/up/down
 * projections (transformers models/qwen2/modeling_qwen2.py:46-48,219-221,245).
 * tile_cfg: 0 = auto; else (BN) | (cta_group << 16) | (max_ctas << 20). out_is_f32: write fp32 (EPI_NONE only). */
int omc_gemm_bf16(const void* X, long long ldx, const void* W, long long ldw, void* out, long long ldo, int M, int N,
                  int K, const void* bias, const void* scale, const void* res, long long ldr, int epi, int out_is_f32,
                  int tile_cfg, void* stream);
/* The same GEMM with the RMSNorm in front of the layer FOLDED in (InternRMSNorm norm1 / norm2, modeling_intern_vit.py:39-44,
 * 218-220; Qwen2RMSNorm input / post-attention / final norm, modeling_qwen2.py:258-263,280-310,411): the caller multiplies the
 * norm weight into W's columns once at load time, X is the RAW residual stream, and the epilogue scales row m by
 * rsqrt(sum_q ssq_in[q][m] / norm_dim + eps) before bias / activation. The sums of squares come from the EPI_RES GEMM that
 * wrote X: with ssq_out set, every N tile of that GEMM stores the sum of squares of the bf16 values it wrote for each row
 * (ssq_out[tile][m], fp32) and reports how many tiles there were in ssq_out_parts (host-side, at launch time: it depends on
 * the tile configuration). For rows that no GEMM produced (embeddings) use omc_row_ssq_rows. */
typedef struct omc_gemm_norm {
  const float* ssq_in;    /* [ssq_in_parts][ssq_in_ld] or NULL */
  long long ssq_in_ld;    /* >= M */
  int ssq_in_parts, norm_dim;
  float eps;
  int ssq_out_max_parts;  /* capacity of ssq_out in parts */
  float* ssq_out;         /* [ssq_out_max_parts][ssq_out_ld] or NULL (EPI_RES / EPI_NONE with bf16 output) */
  long long ssq_out_ld;   /* >= M */
  int ssq_out_parts;      /* OUT: parts written by this launch */
  int reserved;
} omc_gemm_norm;
int omc_gemm_bf16_norm(const void* X, long long ldx, const void* W, long long ldw, void* out, long long ldo, int M, int N,
                       int K, const void* bias, const void* scale, const void* res, long long ldr, int epi, int out_is_f32,
                       int tile_cfg, omc_gemm_norm* nf, void* stream);
/* ssq[m] = sum_k x[m,k]^2 for any number of rows (one warp per row). */
int omc_row_ssq_rows(const void* x, long long ldx, long long rows, int C, float* ssq, void* stream);


/* Skinny variant for M <= 64 (batched decode steps too large for the GEMV kernels; HBM-bound): operands swapped so the
 * 128-row tcgen05 tile runs along N (all TMA bytes are weight bytes), small N split along K over the CTAs of a thread-block
 * cluster whose fp32 partial tiles are summed in rank order through distributed shared memory (bit-identical from call to
 * call). Same epilogues and argument meaning as omc_gemm_bf16; out may alias res. `workspace` / `workspace_bytes` are
 * kept for ABI stability and ignored (the first version reduced through a global workspace); NULL / 0 are fine. */
long long omc_gemm_skinny_workspace_bytes(int max_n);
int omc_gemm_skinny_bf16(const void* X, long long ldx, const void* W, long long ldw, void* out, long long ldo, int M, int N,
                         int K, const void* bias, const void* scale, const void* res, long long ldr, int epi,
                         int out_is_f32, void* workspace, long long workspace_bytes, void* stream);

/* ---- decode-time linear layers: split-K GEMV, 128-bit loads, warp-shuffle reductions ---------------------------
 * out[B,N] = epi( norm(x)[B,K] W[N,K]^T ), B <= 8. If norm_w != NULL the input is RMS-normalised first
 * (Qwen2RMSNorm modeling_qwen2.py:258-263 fused in). epi: NONE (+bias), RES (+res), SWIGLU (interleaved W).
 * out_is_f32 selects fp32 output (lm_head logits, modeling_qwen2.py:470-472). */
int omc_gemv_bf16(const void* x, long long ldx, const void* W, long long ldw, void* out, long long ldo, int B, int N,
                  int K, const void* norm_w, float eps, const void* bias, const void* res, long long ldr, int epi,
                  int out_is_f32, void* stream);

/* ---- row ops --------------------------------------------------------------------------------------------------
 * InternRMSNorm / Qwen2RMSNorm: out = w * cast_bf16(x * rsqrt(mean(x^2) + eps)), statistics in fp32
 * (modeling_intern_vit.py:39-44, modeling_qwen2.py:258-263). Also used in place with ldx=3*C for the full-width
 * QK-norm of modeling_intern_vit.py:143-146,161-165. */
int omc_rmsnorm(const void* x, long long ldx, const void* w, void* out, long long ldo, int rows, int C, float eps,
                void* stream);

/* torch.nn.LayerNorm on bf16 rows - the norm_type = 'layer_norm' option of the InternViT-300M tower
 * (intern_vit_300m/modeling_intern_vit.py:61-64,209-210): out = bf16((x - mean) * rsqrt(var + eps) * w + b), fp32 statistics
 * (biased variance), one rounding. b may be NULL (no bias). */
int omc_layernorm(const void* x, long long ldx, const void* w, const void* b, void* out, long long ldo, int rows, int C,
                  float eps, void* stream);

/* In place on two adjacent C-wide column segments of every row - x[m, 0:C] with w_a, x[m, C:2C] with w_b - in ONE launch:
 * InternAttention's q_norm / k_norm over all heads flattened (modeling_intern_vit.py:143-146) on the packed qkv rows. */
int omc_rmsnorm_pair(void* x, long long ldx, const void* w_a, const void* w_b, int rows, int C, float eps, void* stream);

/* ---- vision tower glue -----------------------------------------------------------------------------------------
 * im2col for the 14x14/stride-14 patch conv (modeling_intern_vit.py:73-75,92): pixels [B,3,H,W] (fp32 if
 * pixels_are_f32 else bf16) -> cols [B*(H/14)*(W/14), ldc] bf16, column index = c*196 + ky*14 + kx (the conv weight's
 * own flattening), zero padded to ldc. */
int omc_vit_im2col(const void* pixels, int pixels_are_f32, void* cols, long long ldc, int B, int H, int W,
                   void* stream);
/* hidden[b,0,:] = cls + pos[0]; hidden[b,1+i,:] = patch[b*P+i,:] + pos[1+i]  (modeling_intern_vit.py:94-101;
 * the bicubic resize of :82-88 is the identity at the native grid and is not re-applied). */
int omc_vit_assemble(const void* patch, const void* cls, const void* pos, void* hidden, int B, int P, int C,
                     void* stream);
/* feature select 'patch' (internVIT_encoder.py:35-43: drop CLS) fused with InternVL-style pixel shuffle.
 * hidden [B, 1+G*G, C] -> out [B, (G*r)^2, C/(r*r)] with r = 1/down (down = 1: plain CLS drop, down = 2: ratio 0.5).
 * Pure gather: bit-exact. */
int omc_select_pixel_shuffle(const void* hidden, void* out, int B, int G, int C, int down, void* stream);

/* ---- any-resolution image preprocessing (the step in front of the path) -------------------------------
 * process_anyres_image (omchat/mm_utils.py:119-158): PIL bicubic resize (Pillow Resample.c, 8 bpc fixed point) of an RGB
 * uint8 HWC image, black-canvas padding, 448x448 patches + the whole image at 448x448, CLIPImageProcessor rescale/normalise
 * (internVIT_encoder.py:26-29). omc_resample_u8 = one separable pass (vertical = 0: width changes, 1: height changes) with
 * HOST-computed Pillow weights: coefs int32 [out, ksize] (22-bit fixed point), bounds int32 [out, 2] = (first tap, taps).
 * Bit-exact against Pillow. omc_anyres_pack gathers the crops [1 + patches, 3, crop, crop] (fp32 or bf16) through a
 * [3][256] fp32 look-up table (value -> normalised float, built by the host with the reference formulas). */
int omc_resample_u8(const void* src, int src_h, int src_w, void* dst, int dst_h, int dst_w, const int32_t* coefs,
                    const int32_t* bounds, int ksize, int vertical, void* stream);
int omc_anyres_pack(const void* thumb, const void* resized, int new_w, int new_h, int target_w, int target_h, int paste_x,
                    int paste_y, int crop, const float* lut, void* out, int out_is_bf16, void* stream);

/* ---- attention ---------------------------------------------------------------------------------------------------
 * Flash-style softmax(Q K^T * scale) V with head_dim 128, fp32 softmax, over packed variable-length sequences.
 * q rows [total, Hq, 128] with row stride ldq (elements), k/v rows [total, Hkv, 128] with strides ldk/ldv, out row
 * stride ldo. Sequence s covers rows cu_seqlens[s]..cu_seqlens[s+1] (int32, device). causal=0: ViT attention
 * (modeling_intern_vit.py:148-152 / flash_attention.py:43-55); causal=1 with Hq = 7*Hkv: Qwen2 GQA prefill
 * (modeling_qwen2.py:161-184,229-243). total_rows = rows of the packed q/k/v/out buffers (= cu_seqlens[num_seqs]; the
 * bound of the TMA tensor maps). Every full 128-row query tile runs on the tcgen05/TMEM kernel (S and O accumulators and
 * the P operand in tensor memory, K/V tiles by TMA); the ragged tail rows run on the mma.sync kernel.
 * omc_attention_set_impl: 0 = tcgen05 kernel (default), 1 = mma.sync kernel for everything (A/B baseline; env
 * OMCHAT_B200_ATTN_LEGACY=1), 2 = the first tcgen05 kernel (128-key tiles, single-buffered S). */
int omc_attention_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                      void* out, long long ldo, const int32_t* cu_seqlens, int num_seqs, int max_seqlen,
                      long long total_rows, int Hq, int Hkv, int causal, float scale, void* stream);
/* The same with an explicit head_dim: 128, or 64 - the 16 x 64 heads of the InternViT-300M tower
 * (intern_vit_300m/configuration_intern_vit.py:66-67; InternAttention, modeling_intern_vit.py:138-172) on the default tcgen05
 * kernel (one 64-column TMA box per tile, 4 k-steps for S = Q K^T, N = 64 for O += P V). q / k / v / out rows are [total, H, head_dim]. */
int omc_attention_fwd_hd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                         void* out, long long ldo, const int32_t* cu_seqlens, int num_seqs, int max_seqlen,
                         long long total_rows, int Hq, int Hkv, int head_dim, int causal, float scale, void* stream);
int omc_attention_set_impl(int impl);
/* diagnostic: device buffer of 12 x 8 uint64 filled by CTA (0,0,0) of the tcgen05 kernel with per-warp clock accumulators
 * (tools/attn_check.py prof-clocks); NULL = off */
int omc_attention_set_prof(void* dev_buf);

/* ---- Qwen2 decoder glue ----------------------------------------------------------------------------------------
 * RoPE (rotate-half, pairs (i, i+64), theta, fp32 angles: modeling_qwen2.py:102-113,124-146) applied in place to the
 * q and k slices of qkv [T, (Hq+2*Hkv)*128] (inv_freq: 64 fp32 on device, computed by the host exactly as
 * Qwen2RotaryEmbedding does), then K and V of every token are appended to the paged cache
 * (replaces DynamicCache.update, modeling_qwen2.py:227). kv_pool: [num_pages, 2, Hkv, page_size, 128] bf16 for ONE
 * layer; token t of sequence seq_ids[t] at position pos[t] goes to page block_table[seq_ids[t]*max_pages +
 * pos[t]/page_size], slot pos[t]%page_size. */
int omc_rope_kv_store(void* qkv, long long ldqkv, const int32_t* pos, const int32_t* seq_ids, int T, int Hq, int Hkv,
                      const float* inv_freq, void* kv_pool, const int32_t* block_table, int max_pages, int page_size,
                      void* stream);
/* One decode step of GQA attention over the paged cache, with the new token's RoPE + cache append fused in.
 * qkv [B, (Hq+2*Hkv)*128]: the new token's q|k|v projections BEFORE RoPE (inv_freq != NULL), or q already rotated with
 * K/V already in the cache (inv_freq == NULL). ctx_lens[B] int32 on device = context length INCLUDING the new token.
 * Keys are split over `splits` CTAs per (sequence, kv head) and merged by the last CTA to finish (log-sum-exp).
 * workspace: omc_decode_attn_workspace_bytes() bytes, zero-initialised once by the caller (self-resetting counters).
 * page_size must be a multiple of 16, Hq/Hkv <= 8. */
int omc_decode_attn_splits(int B, int Hkv, int max_ctx);
long long omc_decode_attn_workspace_bytes(int B, int Hq, int Hkv, int splits);
int omc_paged_decode_attn(const void* qkv, long long ldq, const float* inv_freq, void* kv_pool,
                          const int32_t* block_table, int max_pages, int page_size, const int32_t* ctx_lens, int B,
                          int Hq, int Hkv, int splits, float scale, void* out, long long ldo, void* workspace,
                          void* stream);
/* embed_tokens gather (omchat_arch.py:139): out[t,:] = table[ids[t],:]; ids int64. */
int omc_embed_lookup(const int64_t* ids, int T, const void* table, int C, void* out, long long ldo, int vocab,
                     void* stream);
/* Image-token splice (omchat_arch.py:115-195), integer placement on device, bit-exact.
 * ids: packed int64 [S_total]; seq_offsets int32 [n_seq+1] into ids; every id == image_token (-200) is replaced,
 * in batch-major order, by the L rows of the next image feature block feats[img, L, C]; all other ids gather
 * table rows. Outputs: embeds [T_total, C] packed, pos_ids int32 [T_total] (0..len-1 per sequence), seq_ids int32
 * [T_total], out_offsets int32 [n_seq+1]. max_len > 0 truncates each spliced sequence (omchat_arch.py:161-164).
 * A sequence without placeholders still consumes one feature block (omchat_arch.py:122-129).
 * workspace: int32 [S_total + n_seq + 2]. T_capacity is the row capacity of the outputs
 * (S_total + n_img*(L-1) always suffices). vocab > 0: text ids outside [0, vocab) yield zero rows (omc_embed_lookup too)
 * instead of an out-of-bounds read - nn.Embedding raises there (omchat_arch.py:139); vocab <= 0 disables the check. */
int omc_splice(const int64_t* ids, const int32_t* seq_offsets, int n_seq, int S_total, long long image_token,
               const void* table, const void* feats, int n_img, int L, int C, int max_len, void* embeds,
               int32_t* pos_ids, int32_t* seq_ids, int32_t* out_offsets, int32_t* workspace, int T_capacity,
               int vocab, void* stream);
/* greedy sampling (HF GenerationMixin argmax as driven by cli.py:60-70): next[b] = argmax_v logits[b, v]
 * (lowest index on ties), logits fp32 [B, V] with row stride ldl. workspace: 128*B floats. */
int omc_argmax(const float* logits, long long ldl, int B, int V, int64_t* next, float* workspace, void* stream);

/* ---- weight-streaming GEMM of the batched decode step (omchat_b200/csrc/gemm_stream.cu) --------------------------------
 * Replaces, for M <= 64 rows, the nn.Linear calls of Qwen2Attention / Qwen2MLP / lm_head (transformers modeling_qwen2.py:
 * 46-48, 219-221, 245, 470-472) together with the Qwen2RMSNorm in front of them (:258-263): Y[M,N] = epi(rstd[m] * (X W'^T)).
 * omc_pack_weight re-lays an [N, K] row-major bf16 weight once, at load time, as [N/128][K/64] tiles of 128 x 64 elements in
 * the shared-memory image of the MMA (128-byte swizzle), optionally scaling column k by col_scale[k] (the RMSNorm weight that
 * precedes the layer). `packed` must be 1024-byte aligned and omc_packed_weight_bytes(N, K) long.
 * omc_gemm_stream: X bf16 [M, K] (row stride ldx), packed weights, out bf16 (or fp32) [M, N or N/2 for SwiGLU].
 *   ssq_in  != NULL: fp32 [ssq_parts][64] partial sums of squares of X's rows over norm_dim columns; row m of the product is
 *                    scaled by rsqrt(sum_parts / norm_dim + eps) before bias / activation  (the folded RMSNorm)
 *   ssq_out != NULL: (bf16 output, not SwiGLU) fp32 [ceil(N/128)][64] partial sums of squares of the rows written
 *   workspace: omc_gemm_stream_workspace_bytes() bytes, zeroed once by the caller, shared by consecutive launches of a stream
 *   pdl != 0: launch with programmatic stream serialization (the kernel prefetches weights before it waits for its
 *             predecessor; every access to activations happens after griddepcontrol.wait).
 * omc_row_ssq: sums of squares of rows that were not produced by an EPI_RES epilogue (embedding rows; all-reduced rows
 *   under tensor parallelism): ssq[0][m] = sum_k x[m,k]^2, ssq[1..parts-1][m] = 0.
 * omc_gemm_stream, tp != NULL (row-parallel o_proj / down_proj under tensor parallelism, one process per GPU): the all-reduce over the
 *             tp->size ranks happens inside the epilogue over NVLink peer memory - tp->bufs[r] is rank r's exchange buffer
 *             (omc_gemm_stream_xchg_bytes() bytes from omc_peer_alloc, zeroed; peers' buffers mapped with omc_peer_open),
 *             tp->channel 0 / 1 separates the two row-parallel ops of a layer. Every rank passes the same residual replica and
 *             gets the same bits of output and of ssq_out. Replaces the NCCL all-reduce of the reference-shaped TP plan. */
typedef struct omc_tp_xchg {
  int rank, size, channel, reserved;
  void* bufs[8];
} omc_tp_xchg;
long long omc_gemm_stream_xchg_bytes(void);
long long omc_packed_weight_bytes(int N, int K);
int omc_pack_weight(const void* W, long long ldw, int N, int K, const void* col_scale, void* packed, void* stream);
long long omc_gemm_stream_workspace_bytes(void);
int omc_gemm_stream(const void* X, long long ldx, int M, const void* Wp, int N, int K, void* out, long long ldo,
                    int out_is_f32, const void* bias, const void* res, long long ldr, int epi, const float* ssq_in,
                    int ssq_parts, int norm_dim, float eps, float* ssq_out, void* workspace, int pdl, const omc_tp_xchg* tp,
                    void* stream);
int omc_row_ssq(const void* x, long long ldx, int rows, int C, float* ssq, int parts, int pdl, void* stream);
/* profiling hook: the next max_launches omc_gemm_stream launches write %globaltimer stamps [2 * SMs][16] each into buf
 * (0 entry, 1 weights issued, 2 dependency satisfied, 3 first stage landed, 4 last MMA issued, 5 accumulator read, 6 K parts
 * staged, 7 stores issued, 8 partial tile pushed to the peers, 9 peers' tiles arrived); buf = NULL switches it off.
 * Measurement only (tools/prof_stream.py). */
int omc_gemm_stream_set_prof(void* buf, int max_launches);
/* spin-wait watchdog record: 8 ints of pinned, device-mapped host memory (or NULL). A wait that lasts 10 s writes
 * {kind (1 = stream-K partial, 2 + 16 * rank = tensor-parallel exchange), tile, peer / part, channel, block, grid} and traps. */
int omc_gemm_stream_set_debug(void* pinned_host_record);

/* ---- persistent decode step ("megakernel", omchat_b200/csrc/decode_mega.cu) ------------------------------------------
 * One cooperative launch = one whole Qwen2 decode step for 1..4 sequences: embed_tokens (omchat_arch.py:139), every
 * Qwen2DecoderLayer (modeling_qwen2.py:280-310: RMSNorm, q/k/v + bias, RoPE, paged KV append + GQA attention, o_proj,
 * SwiGLU MLP, residuals), final norm + lm_head (:411,470-472) and the greedy argmax of GenerationMixin (cli.py:60-70).
 * The caller describes the model once (omc_decode_desc: plain device pointers and sizes), builds a plan on the HOST
 * (omc_decode_plan_build is pure CPU code), copies the plan bytes to the device and launches omc_decode_step per token.
 * Per-layer pointer arrays are HOST arrays of device pointers. Weight matrices must be contiguous [N, K] bf16;
 * gate_up_w is the [2*inter, hidden] matrix with rows alternating gate_i, up_i (OMC_EPI_SWIGLU layout).
 * State: tokens[batch] (in: the tokens to feed, out: the greedy next tokens), ctx_lens[batch] (in: tokens already in
 * the cache, incremented by the kernel), token_hist[hist_capacity, batch] / hist_pos[1] (optional history of the sampled
 * tokens, appended at *hist_pos), h/qkv/attn/act/logits scratch of [batch, hidden | (q+2kv)*128 | q*128 | inter | vocab].
 * workspace: omc_decode_workspace_bytes(desc) bytes, ZERO-INITIALISED once by the caller. grid = CTAs = omc_num_sms(). */
typedef struct omc_decode_desc {
  int32_t n_layers, batch, hidden, q_heads, kv_heads, inter, vocab, vocab_offset;
  int32_t page_size, max_pages, grid, hist_capacity;
  int32_t rope_positions;
  int32_t l2_prefetch_stages; /* how many ring stages (one slot each, per CTA) the L2 prefetch runs ahead; 0 = off */
  int32_t ring_slot_bytes;    /* bytes of one shared-memory ring slot (multiple of 128); 0 = default (14336 = 2 rows of
                                 K = 3584; longer rows travel as equal K chunks of at most one slot) */
  int32_t tune;               /* A/B switches for measurements, 0 = defaults. bit 0: FFMA dot products instead of mma.sync;
                                 bit 1: two-row stages for split-K ops; bit 2 / bit 3: never / always cut the MLP into
                                 K-chunk sub-ops (default: batch >= 3); bits 4-7: cap on the number of ring slots;
                                 bit 8: cut only down_proj into K-chunk sub-ops */
  float eps, attn_scale;
  const void* embed;
  const void* final_norm;
  const void* lm_head;
  const float* rope_cs; /* fp32 [rope_positions][64][2]: (cos, sin)(pos * inv_freq[i]), computed by the host exactly as
                           Qwen2RotaryEmbedding does (modeling_qwen2.py:102-113); rope_positions >= max_pages * page_size */
  const void* const* ln1;
  const void* const* qkv_w;
  const void* const* qkv_b;
  const void* const* o_w;
  const void* const* ln2;
  const void* const* gate_up_w;
  const void* const* down_w;
  void* kv_pool; /* layer 0 pool [num_pages, 2, kv_heads, page_size, 128]; layer l at + l * kv_layer_stride elements */
  long long kv_layer_stride;
  const int32_t* block_table;
  int32_t* ctx_lens;
  int64_t* tokens;
  int64_t* token_hist;
  int32_t* hist_pos;
  void* h;
  void* qkv;
  void* attn;
  void* act;
  float* logits;
  void* workspace;
  int32_t* status; /* optional int32[4], zeroed by the caller, device-accessible (pinned host memory is fine): a watchdog
                      inside the kernel writes {code, CTA, detail, thread} here before trapping; NULL = in workspace */
  void* prof;      /* optional uint64[grid][5*n_layers+2][8] device buffer for tools/prof_mega.py: per-CTA, per-op
                      %globaltimer stamps (op start, activations staged, op end), SM id, and warp 0's clock64 cycles in
                      stage wait / dot products / refill issue / reduce + epilogue; NULL = off */
  /* Tensor parallelism INSIDE the persistent kernel (tp_size 2..8, one process per GPU; 0/1 = single GPU). The weight
   * pointers above are this rank's Megatron shards (q/k/v and gate/up column-parallel: q_heads/kv_heads/inter are the LOCAL
   * counts; o_proj/down_proj row-parallel; lm_head vocab-parallel with vocab = local rows and vocab_offset = first row).
   * The all-reduce after o_proj and down_proj (the per-layer NCCL all-reduce of a Megatron decoder) is done by the kernel
   * itself: every CTA pushes the fp32 partial sums of the rows it owns into the peers' exchange buffers with 8-byte
   * {value, tag} stores over NVLink and polls its own buffer for theirs; greedy sampling exchanges (max, index) the same
   * way. xchg[p] = rank p's exchange buffer (omc_decode_xchg_bytes bytes, zero-initialised, allocated with
   * omc_peer_alloc and mapped into this process with omc_peer_open; xchg[tp_rank] = the local one). All ranks must
   * launch omc_decode_step with the same epoch sequence. */
  int32_t tp_rank, tp_size;
  void* xchg[8];
} omc_decode_desc;
long long omc_decode_plan_bytes(int n_layers);
long long omc_decode_workspace_bytes(const omc_decode_desc* desc); /* uses batch, hidden, heads, inter, grid */
int omc_decode_plan_build(const omc_decode_desc* desc, void* plan_host);
/* epoch: a counter the caller increments on every launch that uses the same workspace (it tags the in-flight activations) */
int omc_decode_step(const void* plan_host, const void* plan_dev, unsigned int epoch, void* stream);
long long omc_decode_xchg_bytes(const omc_decode_desc* desc); /* uses batch, hidden, tp_size */

/* ---- model-level entry points (omchat_b200/csrc/model_capi.cu): the layer loops in C++, for hosts that are not Python ----
 * omc_vit_forward = OmChatMetaForCausalLM.encode_images (omchat/model/omchat_arch.py:50-53): InternVITVisionTower.forward
 * (multimodal_encoder/internVIT_encoder.py:45-56 -> intern_vit_6b/modeling_intern_vit.py:90-102,138-222,268-279), feature
 * select 'patch' (:35-43) (+ pixel shuffle), mm_projector mlp2x_gelu (multimodal_projector/builder.py:54-61).
 *   pixels [n, 3, S, S] fp32 or bf16 -> feats_out bf16 [n, (S/patch/down)^2, proj_hidden]. Weight pointers are bf16 device
 *   tensors in nn.Linear's [out, in] layout (patch_w: [hidden, patch_k] = the conv weight flattened, K zero-padded to patch_k);
 *   per-layer pointers are arrays of n_layers. workspace: omc_vit_workspace_bytes(desc, n) bytes, 256-byte aligned.
 * omc_decoder_prefill = Qwen2Model.forward on PACKED sequences (transformers modeling_qwen2.py:280-310,353-414) + final norm
 *   and lm_head on each sequence's last row (:411,470-472), writing K/V into the paged cache. Uses the omc_decode_desc
 *   fields n_layers, hidden, q_heads, kv_heads, inter, vocab, page_size, max_pages, eps, attn_scale, final_norm, lm_head,
 *   ln1 .. down_w, kv_pool, kv_layer_stride, block_table. embeds [T, hidden] bf16 is the residual stream (updated in place),
 *   pos_ids / seq_ids int32 [T], cu_seqlens int32 [n_seq + 1], last_rows int64 [n_seq] (packed row of each sequence's last
 *   token), last_logits fp32 [n_seq, vocab] (NULL: cache fill only). workspace: omc_decoder_prefill_workspace_bytes().
 * Neither allocates nor synchronises: both can be captured in a CUDA graph. Same kernels, same order, same bits as the Python
 * host path (omchat_b200/model/vision.py, decoder.py). The decode step's model-level entry is omc_decode_step above. */
typedef struct omc_vit_desc {
  int32_t n_layers, hidden, heads, inter, image_size, patch_size, patch_k, qk_norm;
  int32_t pixel_shuffle_down, proj_hidden;
  float eps;
  int32_t norm_folded; /* 1: qkv_w / fc1_w carry norm1 / norm2 in their columns (W * g): the loop uses omc_gemm_bf16_norm and no
                          stand-alone RMSNorm in front of them (the product's default path); 0: plain weights + omc_rmsnorm */
  const void* patch_w;
  const void* patch_b;
  const void* cls;
  const void* pos;
  const void* const* norm1;
  const void* const* qkv_w;
  const void* const* q_norm;
  const void* const* k_norm;
  const void* const* proj_w;
  const void* const* proj_b;
  const void* const* ls1;
  const void* const* norm2;
  const void* const* fc1_w;
  const void* const* fc1_b;
  const void* const* fc2_w;
  const void* const* fc2_b;
  const void* const* ls2;
  const void* p_w0;
  const void* p_b0;
  const void* p_w2;
  const void* p_b2;
  /* InternViT-300M variant (intern_vit_300m/modeling_intern_vit.py:61-64,131,209-210); all zero / NULL for the 6B tower */
  int32_t norm_type;     /* 0: InternRMSNorm; 1: nn.LayerNorm with norm1_b / norm2_b (requires norm_folded = 0) */
  int32_t attn_head_dim; /* 0: hidden / heads (128, or 64 = the 300M tower, native). 128 with hidden / heads < 128: every head of qkv_w's rows (and of
                            qkv_b) and of proj_w's columns is zero-padded to 128 dims by the caller, qkv_w is [3 * heads * 128,
                            hidden], proj_w [hidden, heads * 128]; q.k and P.V are unchanged, the softmax scale stays
                            (hidden / heads)^-0.5 */
  const void* const* norm1_b;
  const void* const* norm2_b;
  const void* const* qkv_b; /* NULL: no qkv bias (config.qkv_bias = false) */
} omc_vit_desc;
long long omc_vit_workspace_bytes(const omc_vit_desc* desc, int max_crops);
int omc_vit_forward(const omc_vit_desc* desc, const void* pixels, int pixels_are_f32, int n_crops, void* workspace,
                    void* feats_out, void* stream);
long long omc_decoder_prefill_workspace_bytes(const omc_decode_desc* desc, int T, int n_seq);
int omc_decoder_prefill(const omc_decode_desc* desc, const float* inv_freq, void* embeds, const int32_t* pos_ids,
                        const int32_t* seq_ids, const int32_t* cu_seqlens, int n_seq, int T, int max_len,
                        const int64_t* last_rows, void* workspace, float* last_logits, int norm_folded, void* stream);
/* norm_folded = 1: desc->qkv_w / gate_up_w carry the input / post-attention RMSNorm weights in their columns (ln1 / ln2 are
 * then unused); final_norm + lm_head stay separate. */

/* ---- Qwen2-MoE sparse MLP block --------------------------------------------------------------------------------------
 * The language model behind omchat/model/language_model/omchat_qwen2_moe.py:28-117 is transformers' Qwen2MoeForCausalLM; its
 * sparse block (modeling_qwen2_moe.py:295-374) is: router softmax over all experts (fp32) -> top-k (optionally renormalised)
 * -> every token through its k expert SwiGLU MLPs, weighted -> + sigmoid(shared_expert_gate(x)) * shared_expert(x).
 * Device-side plan, no host synchronisation (capturable in a CUDA graph):
 *   omc_moe_route    x [T, C] bf16, router_w [E, C], shared_gate_w [C] or NULL -> topk_ids int32 [T, k], topk_w fp32 [T, k],
 *                    shared_gate fp32 [T] (sigmoid), counts int32 [E] += histogram (counts must be zero on entry: zero it
 *                    once, omc_moe_plan re-zeroes it). E <= 128, k <= 8. norm_w == NULL: x holds the post-attention-normed
 *                    rows; norm_w != NULL: x is the raw residual stream, the kernel applies post_attention_layernorm
 *                    (Qwen2MoeRMSNorm, modeling_qwen2_moe.py:70-75, same rounding as omc_rmsnorm) itself and also writes the
 *                    normed rows to xn_out [T, ldn] for omc_moe_scatter / the shared expert.
 *   omc_moe_select   the selection alone, for logits computed elsewhere: logits fp32 [T, ld], columns 0..E-1 = router logits,
 *                    column E = the shared expert's gate logit when has_gate. At prefill sizes the host computes them with
 *                    omc_gemm_bf16 (fp32 out) against [router_w; shared_gate_w; zero rows up to a multiple of 128] - the 61 dot
 *                    products per token cost 170 us per 8192 tokens on the CUDA cores and ~10 us on the tensor cores.
 *   omc_moe_plan     counts -> seg_start int32 [E] (first row of every expert's segment, segments padded to whole 128-row
 *                    tiles), tile_expert int32 [max_tiles] (expert of every 128-row tile, -1 = unused), cursor [E] = 0,
 *                    counts = 0. max_tiles >= omc_moe_max_tiles(T, k, E) = T * k / 128 + E.
 *   omc_moe_scatter  copies row t of x to its k slots of xperm [max_tiles * 128, ldp] and records them in slot_of int32 [T, k]
 *   omc_gemm_bf16_grouped   out[r, :] = epi(xperm[r, :] . W[tile_expert[r / 128]]^T) for W = [n_experts * N, K] stacked expert
 *                    matrices on the tcgen05 GEMM (epi OMC_EPI_SWIGLU on interleaved gate/up rows, or OMC_EPI_NONE); tiles
 *                    with tile_expert < 0 are skipped without touching the weights. M_max = max_tiles * 128, N % 128 == 0.
 *                    active_tiles_hint: an upper bound of the tiles that can be active (min(max_tiles, T * k); 0 = M_max / 128),
 *                    used only to pick the tile width (narrower tiles when few experts' weights are streamed).
 *   omc_moe_combine  h[t] += sum_j topk_w[t, j] * yperm[slot_of[t, j]] + shared_gate[t] * shared_y[t]  (shared_y may be NULL),
 *                    fp32 accumulate, one bf16 rounding: the block's output + the decoder layer's residual add. ssq_out != NULL
 *                    (T <= 64): also leaves the new rows' sums of squares in omc_row_ssq's [ssq_parts][64] layout for the
 *                    folded RMSNorm of the next omc_gemm_stream.
 *   omc_moe_plan_scatter   omc_moe_plan + omc_moe_scatter; for T * top_k <= 16 (small decode steps) as ONE single-CTA launch. */
int omc_moe_max_tiles(int T, int top_k, int n_experts);
/* 1: the routing kernels and the grouped GEMMs are launched as programmatic dependents of their predecessors
 * (cudaLaunchAttributeProgrammaticStreamSerialization; every one of them starts with griddepcontrol.launch_dependents +
 * griddepcontrol.wait). 0 (default): plain launches - measured, the overlap buys nothing at these sizes (batch-1 step 2.02 ms
 * with, 1.93 ms without; batch 32 6.74 vs 6.84 ms). */
int omc_moe_set_pdl(int on);
int omc_moe_route(const void* x, long long ldx, int T, int C, const void* norm_w, float eps, void* xn_out, long long ldn,
                  const void* router_w, const void* shared_gate_w, int n_experts, int top_k, int norm_topk, int32_t* topk_ids,
                  float* topk_w, float* shared_gate, int32_t* counts, void* stream);
int omc_moe_select(const float* logits, long long ld, int T, int n_experts, int top_k, int norm_topk, int has_gate,
                   int32_t* topk_ids, float* topk_w, float* shared_gate, int32_t* counts, void* stream);
int omc_moe_plan(int32_t* counts, int n_experts, int max_tiles, int32_t* seg_start, int32_t* cursor, int32_t* tile_expert,
                 void* stream);
int omc_moe_scatter(const void* x, long long ldx, int T, int C, const int32_t* topk_ids, int top_k, const int32_t* seg_start,
                    int32_t* cursor, void* xperm, long long ldp, int32_t* slot_of, void* stream);
int omc_gemm_bf16_grouped(const void* X, long long ldx, int M_max, const void* W, long long ldw, int n_experts, int N, int K,
                          const int32_t* tile_expert, int active_tiles_hint, void* out, long long ldo, int epi, void* stream);
int omc_moe_combine(void* h, long long ldh, int T, int C, const void* yperm, long long ldy, const int32_t* slot_of,
                    const float* topk_w, int top_k, const void* shared_y, long long lds, const float* shared_gate, float* ssq_out,
                    int ssq_parts, void* stream);
int omc_moe_plan_scatter(int32_t* counts, int n_experts, int max_tiles, int32_t* seg_start, int32_t* cursor, int32_t* tile_expert,
                         const void* x, long long ldx, int T, int C, const int32_t* topk_ids, int top_k, void* xperm, long long ldp,
                         int32_t* slot_of, void* stream);

/* ---- peer (NVLink) memory for the tensor-parallel decode step ------------------------------------------------------
 * Replaces the NCCL communicator a Megatron-style decoder would hand to its all-reduce: one exchange buffer per rank,
 * visible to every rank of the node. omc_peer_alloc: cudaMalloc + zero-fill on the current device and export a 64-byte
 * CUDA IPC handle; omc_peer_open: map another process's buffer (enables peer access); omc_peer_close / omc_peer_free undo
 * them. The handles travel between the processes through the caller's own channel (torch.distributed here). */
int omc_peer_alloc(long long bytes, void** ptr, void* handle64);
int omc_peer_open(const void* handle64, void** ptr);
int omc_peer_close(void* ptr);
int omc_peer_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* OMCHAT_B200_H_ */
