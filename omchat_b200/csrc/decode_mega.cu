// Persistent decode step ("megakernel") for small batches (B <= 4): ONE cooperative launch runs a whole Qwen2 decode step
//   embed -> 28 x [RMSNorm+QKV GEMV, RoPE + KV append + paged GQA attention, O GEMV (+res), RMSNorm+gate/up GEMV
//   (SwiGLU), down GEMV (+res)] -> final RMSNorm + lm_head GEMV -> greedy argmax (+ token history, ctx_lens += 1)
// with one CTA per SM. A decode step is HBM-bound weight streaming (14.1 GB per token for Qwen2-7B in bf16), so the
// kernel is organised around keeping HBM busy across op boundaries:
//   weight ring  this CTA's row slab of every weight matrix, op after op, is one long sequence of "stages" streamed into a
//                shared-memory ring with cp.async.bulk (TMA, mbarrier complete_tx, L2 evict-first). The ring refills
//                itself: the warp that finishes reading a slot immediately issues the copy of the stage that will
//                occupy it next (stage + nslots), whichever op that belongs to. Weights are immutable, so the stream runs
//                ahead across phase boundaries: the ring (~170 KB/SM, 25 MB chip-wide) stays full while the CTA sits in a
//                grid barrier, stages activations or runs the attention phase, which hides those behind HBM streaming.
//   8 warps      per op: grid barrier -> stage the activation vector(s) in shared memory (RMSNorm fused) -> each warp
//                owns whole "units" (R weight rows x full K) of the ring, dot products with fp32 accumulation,
//                warp-shuffle reduction, fused epilogue (bias / residual / SwiGLU / fp32 logits + running argmax).
// Attention runs inside the same kernel: (sequence, kv-head) items are split over CTAs by key range, the 7 query heads
// of a group share each K/V row read, partial (m, l, O) are merged by the last CTA to finish (self-resetting counters).
// Grid barriers are a monotonic global counter (release add / acquire spin), reset by the last CTA to leave the kernel.
//
// Reference call sites replaced: transformers models/qwen2/modeling_qwen2.py:280-310 (decoder layer), :206-246 (attention,
// RoPE :124-146, cache update :227), :46-48 (MLP), :258-263 (RMSNorm), :411,470-472 (final norm + lm_head) and the HF
// GenerationMixin greedy argmax driven by cli.py:60-70; omchat_arch.py:139 (embed_tokens).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "omc_internal.h"
#include "ptx.cuh"

namespace omc {

typedef __nv_bfloat16 bf16;

constexpr int kMegaThreads = 256;
constexpr int kConsumers = 256;
constexpr int kCWarps = 8;
constexpr int kRMax = 4;            // weight rows per ring stage (upper bound)
constexpr int kSlotBytes = 19456;   // ring slot: 2 rows of K=3584 (14336 B) or half a row of K=18944 (18944 B)
constexpr int kMaxSlots = 12;
constexpr int kMaxOps = 192;
constexpr int kAttnKeysPerCta = 128;
constexpr int kPartStride = 130;    // O[128], m, l
constexpr int kAttnScratchBytes = kCWarps * 8 * kPartStride * 4;
constexpr int kSmemLimit = 227 * 1024;
constexpr unsigned long long kWaitLimitNs = 4000000000ull;  // a protocol bug must end in a trap, never in a hung GPU

enum { OP_GEMV = 1, OP_ATTN = 2, OP_FINAL = 3 };
enum { F_OUT_F32 = 1, F_X_EMBED = 2, F_ARGMAX = 4 };

struct MegaOp {  // 112 bytes
  int32_t type, N, K, epi;
  int32_t R, ksplit, gran, flags;
  const bf16* W;
  const bf16* x;
  const bf16* norm_w;
  const bf16* bias;
  const bf16* res;
  void* out;
  bf16* aux;  // GEMV + F_X_EMBED: residual stream to seed (h);  ATTN: this layer's KV pool
  int32_t ldx, ldo, ldr, pad[3];
};
static_assert(sizeof(MegaOp) == 112, "MegaOp layout");

struct MegaPlan {  // header, followed by n_ops MegaOp
  int32_t n_ops, B, C, Hq, Hkv, G, page_size, max_pages;
  int32_t grid, nsplit_max, vocab_offset, hist_capacity, kmax, nslots, region_a_bytes, smem_bytes;
  float eps, scale_log2, pad0, pad1;
  const bf16* embed;
  const float* inv_freq;
  const int32_t* block_table;
  int32_t* ctx_lens;
  int64_t* tokens;
  int64_t* token_hist;
  int32_t* hist_pos;
  unsigned int* bar_ctr;
  unsigned int* exit_ctr;
  float* attn_part;
  unsigned int* attn_ctr;
  float* amax_val;
  int32_t* amax_idx;
  int32_t* err_flag;
};
static_assert(sizeof(MegaPlan) % 16 == 0, "MegaPlan must keep the op array 16-byte aligned");

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(kEvictFirst)
      : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void consumer_sync() { __syncthreads(); }

// err_flag[0..3] = {code, CTA, detail, thread}; the flag may live in pinned host memory so it survives the trap
__device__ __noinline__ void mega_fail(int32_t* err_flag, int code, int detail = 0) {
  if (err_flag && atomicCAS(reinterpret_cast<int*>(err_flag), 0, code) == 0) {
    volatile int32_t* e = err_flag;
    e[1] = (int)blockIdx.x;
    e[2] = detail;
    e[3] = (int)threadIdx.x;
  }
  __threadfence_system();
  __trap();
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// bounded mbarrier wait (wall-clock bound, checked every 256 polls)
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity, int32_t* err_flag, int code, int detail) {
  if (mbar_try_wait(bar, parity)) return;
  const unsigned long long t0 = global_ns();
  unsigned int spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0 && global_ns() - t0 > kWaitLimitNs) mega_fail(err_flag, code, detail);
  }
}

__device__ __forceinline__ void slab_of(const MegaOp& op, int cta, int grid, int& row0, int& rows) {
  const long long ng = op.N / op.gran;
  const int g0 = (int)(ng * cta / grid), g1 = (int)(ng * (cta + 1) / grid);
  row0 = g0 * op.gran;
  rows = (g1 - g0) * op.gran;
}

__device__ __forceinline__ void unpack8(uint4 v, float* f) {
  float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ float dot8(uint4 w, const float* x, float acc) {
  float f[8];
  unpack8(w, f);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc = fmaf(f[e], x[e], acc);
  return acc;
}
__device__ __forceinline__ float silu_m(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }
__device__ __forceinline__ float ldcg_bf16(const bf16* p) {
  return __bfloat162float(__ushort_as_bfloat16(__ldcg(reinterpret_cast<const unsigned short*>(p))));
}

// Position of a stage index in the op list: the GEMV op that owns it and this CTA's slab of that op.
struct StageCursor {
  int op_i;            // index of the current GEMV op (n_ops = exhausted)
  uint32_t base, cnt;  // stages [base, base + cnt) belong to op_i
  int row0, rows;
};

struct MegaCtx {
  const MegaPlan* P;     // global
  const MegaOp* ops;     // shared copy
  uint64_t* full;
  volatile uint32_t* gen;  // gen[slot] = number of copies issued into the slot so far (monotonic: no parity aliasing)
  float* red;            // [16] floats of block-reduction scratch
  int* flag;             // [4] ints
  uint8_t* region_a;     // activation vectors / attention scratch
  uint8_t* ring;
  int nslots, cta, grid, n_ops;
  unsigned int bar_target;
};

__device__ __forceinline__ void cursor_load(const MegaCtx& c, StageCursor& k) {
  while (k.op_i < c.n_ops && c.ops[k.op_i].type != OP_GEMV) ++k.op_i;
  if (k.op_i < c.n_ops) {
    const MegaOp& op = c.ops[k.op_i];
    slab_of(op, c.cta, c.grid, k.row0, k.rows);
    k.cnt = (uint32_t)(((k.rows + op.R - 1) / op.R) * op.ksplit);
  } else {
    k.cnt = 0;
  }
}
// Issue the bulk copy of stage s (if it exists) into ring slot s % nslots. Called by ONE lane; `k` is the caller's cursor,
// stage indices passed through one cursor are strictly increasing.
__device__ __forceinline__ void issue_stage(const MegaCtx& c, StageCursor& k, uint32_t s) {
  while (k.op_i < c.n_ops && s >= k.base + k.cnt) {
    k.base += k.cnt;
    ++k.op_i;
    cursor_load(c, k);
  }
  if (k.op_i >= c.n_ops) return;
  const MegaOp& op = c.ops[k.op_i];
  const uint32_t rel = s - k.base;
  const int u = (int)(rel / (uint32_t)op.ksplit), ks = (int)(rel % (uint32_t)op.ksplit);
  const int r = u * op.R;
  const int rows_here = min(op.R, k.rows - r);
  const int Kc = op.K / op.ksplit;
  const uint32_t bytes = (uint32_t)rows_here * (uint32_t)Kc * 2u;
  const uint32_t slot = s % (uint32_t)c.nslots;
  fence_proxy_async();  // generic-proxy reads of this slot (previous stage) are ordered before the async-proxy refill
  mbar_arrive_expect_tx(&c.full[slot], bytes);
  bulk_g2s(c.ring + (size_t)slot * kSlotBytes, op.W + (size_t)(k.row0 + r) * op.K + (size_t)ks * Kc, bytes, &c.full[slot]);
  __threadfence_block();
  c.gen[slot] = s / (uint32_t)c.nslots + 1u;  // publish: the barrier is now in the phase that carries stage s
}
// Wait until stage s has landed in its slot. A slot is shared by stages s, s + nslots, ... that different warps consume,
// and an mbarrier parity wait is only meaningful for the phase in flight, so first wait (monotonic counter) until the
// copy of stage s has actually been issued, then for its bytes.
__device__ __forceinline__ void wait_stage(const MegaCtx& c, uint32_t s, uint32_t slot) {
  const uint32_t g = s / (uint32_t)c.nslots;
  if (c.gen[slot] < g + 1u) {
    const unsigned long long t0 = global_ns();
    unsigned int spins = 0;
    while (c.gen[slot] < g + 1u) {
      if ((++spins & 1023u) == 0 && global_ns() - t0 > kWaitLimitNs) mega_fail(c.P->err_flag, 4, (int)s);
    }
  }
  __threadfence_block();
  mbar_wait_bounded(&c.full[slot], g & 1u, c.P->err_flag, 3, (int)s);
}

// ------------------------------------------------------------------------------------------------ grid barrier
__device__ __forceinline__ void grid_sync(MegaCtx& c, int ctid) {
  consumer_sync();
  c.bar_target += (unsigned int)c.grid;
  if (ctid == 0) {
    __threadfence();
    red_release_gpu_add(c.P->bar_ctr, 1u);
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    while (ld_acquire_gpu(c.P->bar_ctr) < c.bar_target) {
      if ((++spins & 1023u) == 0) {
        if (t0 == 0) t0 = global_ns();
        else if (global_ns() - t0 > kWaitLimitNs) mega_fail(c.P->err_flag, 2, (int)c.bar_target);
      }
    }
    __threadfence();
  }
  consumer_sync();
}

// ------------------------------------------------------------------------------------------------ activation staging
// x[b, :K] (bf16, produced earlier in this launch by other CTAs: read through L2) -> shared memory, optionally
// RMS-normalised: out = w * bf16(x * rsqrt(mean(x^2) + eps))  [modeling_qwen2.py:258-263].
template <int NB>
__device__ void stage_x(MegaCtx& c, const MegaOp& op, int ctid) {
  const MegaPlan& P = *c.P;
  const int K = op.K, nvec = K >> 3;
  uint4* xs = reinterpret_cast<uint4*>(c.region_a);
#pragma unroll 1
  for (int b = 0; b < NB; ++b) {
    if (b >= P.B) break;
    const uint4* src;
    if (op.flags & F_X_EMBED) src = reinterpret_cast<const uint4*>(P.embed + (long long)P.tokens[b] * P.C);
    else src = reinterpret_cast<const uint4*>(op.x + (long long)b * op.ldx);
    uint4* dst = xs + (long long)b * nvec;
    if (op.norm_w == nullptr) {
      for (int i = ctid; i < nvec; i += kConsumers) dst[i] = __ldcg(src + i);
    } else {
      uint4 v[2];
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int i = ctid + j * kConsumers;
        if (i < nvec) {
          v[j] = __ldcg(src + i);
          float f[8];
          unpack8(v[j], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) ss = fmaf(f[e], f[e], ss);
        }
      }
      if (op.aux != nullptr && (op.flags & F_X_EMBED)) {
        // seed the residual stream h with the raw embedding row: every CTA writes its own slice
        const int lo = (int)((long long)nvec * c.cta / c.grid), hi = (int)((long long)nvec * (c.cta + 1) / c.grid);
        uint4* hdst = reinterpret_cast<uint4*>(op.aux + (long long)b * op.ldr);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int i = ctid + j * kConsumers;
          if (i >= lo && i < hi) __stcg(hdst + i, v[j]);
        }
      }
      ss = warp_sum(ss);
      consumer_sync();  // c.red free (previous b / previous user done)
      if ((ctid & 31) == 0) c.red[ctid >> 5] = ss;
      consumer_sync();
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kCWarps; ++w) tot += c.red[w];
      const float rstd = rsqrtf(tot / (float)K + P.eps);
      const uint4* wv = reinterpret_cast<const uint4*>(op.norm_w);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int i = ctid + j * kConsumers;
        if (i < nvec) {
          const uint4 g = __ldg(wv + i);
          uint32_t xi[4] = {v[j].x, v[j].y, v[j].z, v[j].w}, gi[4] = {g.x, g.y, g.z, g.w}, oo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float2 a = unpack_bf16(xi[q]), w2 = unpack_bf16(gi[q]);
            float2 n = unpack_bf16(pack_bf16(a.x * rstd, a.y * rstd));
            oo[q] = pack_bf16(n.x * w2.x, n.y * w2.y);
          }
          dst[i] = make_uint4(oo[0], oo[1], oo[2], oo[3]);
        }
      }
    }
  }
  consumer_sync();
}

// ------------------------------------------------------------------------------------------------ GEMV consumer
template <int NB>
__device__ void gemv_consume(MegaCtx& c, const MegaOp& op, uint32_t sc_base, int row0, int rows, int cw, int lane,
                             StageCursor& refill, float& best_v, int& best_i) {
  const MegaPlan& P = *c.P;
  const int R = op.R, ksplit = op.ksplit;
  const int Kc = op.K / ksplit, nv = Kc >> 3, nvK = op.K >> 3;
  const int units = (rows + R - 1) / R;
  const uint4* xs = reinterpret_cast<const uint4*>(c.region_a);
#pragma unroll 1
  for (int u = cw; u < units; u += kCWarps) {
    const int r = u * R;
    const int rows_here = min(R, rows - r);
    float acc[kRMax][NB];
#pragma unroll
    for (int i = 0; i < kRMax; ++i)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[i][b] = 0.f;
#pragma unroll 1
    for (int ks = 0; ks < ksplit; ++ks) {
      const uint32_t s = sc_base + (uint32_t)(u * ksplit + ks);
      const uint32_t slot = s % (uint32_t)c.nslots;
      wait_stage(c, s, slot);
      const uint4* wb = reinterpret_cast<const uint4*>(c.ring + (size_t)slot * kSlotBytes);
      const uint4* xb = xs + ks * nv;
      if (rows_here == 2) {
#pragma unroll 2
        for (int j = lane; j < nv; j += 32) {
          const uint4 w0 = wb[j], w1 = wb[nv + j];
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            float xf[8];
            unpack8(xb[b * nvK + j], xf);
            acc[0][b] = dot8(w0, xf, acc[0][b]);
            acc[1][b] = dot8(w1, xf, acc[1][b]);
          }
        }
      } else if (rows_here == 1) {
#pragma unroll 4
        for (int j = lane; j < nv; j += 32) {
          const uint4 w0 = wb[j];
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            float xf[8];
            unpack8(xb[b * nvK + j], xf);
            acc[0][b] = dot8(w0, xf, acc[0][b]);
          }
        }
      } else {
#pragma unroll 1
        for (int j = lane; j < nv; j += 32) {
          float xf[NB][8];
#pragma unroll
          for (int b = 0; b < NB; ++b) unpack8(xb[b * nvK + j], xf[b]);
#pragma unroll
          for (int i = 0; i < kRMax; ++i) {
            if (i < rows_here) {
              const uint4 w = wb[i * nv + j];
#pragma unroll
              for (int b = 0; b < NB; ++b) acc[i][b] = dot8(w, xf[b], acc[i][b]);
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) issue_stage(c, refill, s + (uint32_t)c.nslots);  // this slot is free again: refill it
    }
#pragma unroll
    for (int i = 0; i < kRMax; ++i)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[i][b] = warp_sum(acc[i][b]);

    // ---- epilogue: lane (i * NB + b) owns output (row r + i, sequence b)
    const int grow = row0 + r;
    if (op.epi == EPI_SWIGLU) {
#pragma unroll
      for (int pr = 0; pr < kRMax / 2; ++pr)
#pragma unroll
        for (int b = 0; b < NB; ++b)
          if (2 * pr < rows_here && b < P.B && lane == pr * NB + b) {
            const float val = silu_m(acc[2 * pr][b]) * acc[2 * pr + 1][b];
            static_cast<bf16*>(op.out)[(long long)b * op.ldo + (grow >> 1) + pr] = __float2bfloat16(val);
          }
    } else {
#pragma unroll
      for (int i = 0; i < kRMax; ++i)
#pragma unroll
        for (int b = 0; b < NB; ++b)
          if (i < rows_here && b < P.B && lane == i * NB + b) {
            const int row = grow + i;
            float val = acc[i][b];
            if (op.bias) val += __bfloat162float(op.bias[row]);
            if (op.epi == EPI_RES) val += ldcg_bf16(op.res + (long long)b * op.ldr + row);
            if (op.flags & F_OUT_F32) static_cast<float*>(op.out)[(long long)b * op.ldo + row] = val;
            else static_cast<bf16*>(op.out)[(long long)b * op.ldo + row] = __float2bfloat16(val);
            if ((op.flags & F_ARGMAX) && (val > best_v || (val == best_v && row < best_i))) {
              best_v = val;
              best_i = row;
            }
          }
    }
  }
}

// ------------------------------------------------------------------------------------------------ attention phase
// One decode step of GQA attention over the paged cache for item (sequence b, kv head), keys split over CTAs.
// Thread layout: one warp per key, lane owns dims [4*lane, 4*lane+4) of the 128-wide head (8-byte loads, a K or V row is
// one coalesced 256-byte warp access); the G query heads of the group are all evaluated against each K/V row read.
// RoPE (rotate-half, pairs (i, i+64) = lanes (l, l^16)) of the new q/k and the cache append are fused in; K is rounded
// to bf16 before use exactly like the cached copy later steps will read.
__device__ __forceinline__ void unpack4(uint2 v, float* f) {
  const float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}

__device__ void attn_phase(MegaCtx& c, const MegaOp& op, int ctid) {
  const MegaPlan& P = *c.P;
  const int items = P.B * P.Hkv;
  const int item = c.cta % items, split = c.cta / items;
  if (split >= P.nsplit_max) return;
  const int b = item / P.Hkv, kvh = item % P.Hkv, G = P.G;
  const int n_cached = __ldcg(P.ctx_lens + b);  // keys already in the cache; the new token sits at position n_cached
  int nsplit = (n_cached + kAttnKeysPerCta - 1) / kAttnKeysPerCta;
  nsplit = max(1, min(nsplit, P.nsplit_max));
  if (split >= nsplit) return;
  const int per = (n_cached + nsplit - 1) / nsplit;
  const int k0 = split * per, k1 = min(n_cached, k0 + per);
  const bool has_new = (split == nsplit - 1);

  const int cw = ctid >> 5, lane = ctid & 31;
  const bf16* qrow = op.x + (long long)b * op.ldx;
  bf16* pool = op.aux;
  const long long page_stride = 2LL * P.Hkv * P.page_size * 128;
  const long long v_off = (long long)P.Hkv * P.page_size * 128;
  const float sl2 = P.scale_log2;

  // rotary factors of the new position for this lane's 4 dims
  float cs[4], sn[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) sincosf((float)n_cached * P.inv_freq[(lane & 15) * 4 + e], &sn[e], &cs[e]);
  const float sgn = (lane < 16) ? -1.f : 1.f;

  auto load_rot = [&](const bf16* head, float* outv) {
    float own[4], par[4];
    unpack4(__ldcg(reinterpret_cast<const uint2*>(head + lane * 4)), own);
    unpack4(__ldcg(reinterpret_cast<const uint2*>(head + (lane ^ 16) * 4)), par);
#pragma unroll
    for (int e = 0; e < 4; ++e) outv[e] = bf16_round(own[e] * cs[e] + sgn * par[e] * sn[e]);
  };

  float q[8][4];
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    if (h < G) load_rot(qrow + (kvh * G + h) * 128, q[h]);
    else q[h][0] = q[h][1] = q[h][2] = q[h][3] = 0.f;
  }
  float m[8], l[8], o[8][4];
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    m[h] = -INFINITY;
    l[h] = 0.f;
    o[h][0] = o[h][1] = o[h][2] = o[h][3] = 0.f;
  }
  auto consume_key = [&](const float* kf, const float* vf) {
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      if (h < G) {
        float s = q[h][0] * kf[0];
        s = fmaf(q[h][1], kf[1], s);
        s = fmaf(q[h][2], kf[2], s);
        s = fmaf(q[h][3], kf[3], s);
        s = warp_sum(s);
        const float m_new = fmaxf(m[h], s);
        const float corr = exp2f((m[h] - m_new) * sl2);  // m = -inf -> 0
        const float p = exp2f((s - m_new) * sl2);
        m[h] = m_new;
        l[h] = l[h] * corr + p;
#pragma unroll
        for (int e = 0; e < 4; ++e) o[h][e] = fmaf(o[h][e], corr, p * vf[e]);
      }
    }
  };

  constexpr int U = 4;  // keys in flight per warp
#pragma unroll 1
  for (int kb = k0 + cw; kb < k1; kb += kCWarps * U) {
    uint2 kr[U], vr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int kk = kb + kCWarps * u;
      if (kk < k1) {
        const int page = __ldg(P.block_table + (long long)b * P.max_pages + kk / P.page_size);
        const bf16* kp = pool + (long long)page * page_stride + ((long long)kvh * P.page_size + kk % P.page_size) * 128 + lane * 4;
        kr[u] = __ldcg(reinterpret_cast<const uint2*>(kp));
        vr[u] = __ldcg(reinterpret_cast<const uint2*>(kp + v_off));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (kb + kCWarps * u < k1) {  // warp-uniform
        float kf[4], vf[4];
        unpack4(kr[u], kf);
        unpack4(vr[u], vf);
        consume_key(kf, vf);
      }
    }
  }
  if (has_new && cw == 0) {
    // the new token: rotate K, append K/V to the cache, attend to it
    float kf[4], vf[4];
    load_rot(qrow + (P.Hq + kvh) * 128, kf);
    const uint2 vraw = __ldcg(reinterpret_cast<const uint2*>(qrow + (P.Hq + P.Hkv + kvh) * 128 + lane * 4));
    unpack4(vraw, vf);
    const int page = __ldg(P.block_table + (long long)b * P.max_pages + n_cached / P.page_size);
    bf16* kp = pool + (long long)page * page_stride + ((long long)kvh * P.page_size + n_cached % P.page_size) * 128 + lane * 4;
    *reinterpret_cast<uint2*>(kp) = make_uint2(pack_bf16(kf[0], kf[1]), pack_bf16(kf[2], kf[3]));
    *reinterpret_cast<uint2*>(kp + v_off) = vraw;
    consume_key(kf, vf);
  }
  float* part = reinterpret_cast<float*>(c.region_a);  // [8 warps][8 heads][130]
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    if (h < G) {
      float* dst = part + (cw * 8 + h) * kPartStride;
      *reinterpret_cast<float2*>(dst + lane * 4) = make_float2(o[h][0], o[h][1]);
      *reinterpret_cast<float2*>(dst + lane * 4 + 2) = make_float2(o[h][2], o[h][3]);
      if (lane == 0) {
        dst[128] = m[h];
        dst[129] = l[h];
      }
    }
  }
  consumer_sync();
  // ---- merge the 8 warps: thread -> (head = ctid / 32, dims 4 * (ctid % 32) ..)
  const int h = ctid >> 5, d4 = (ctid & 31) * 4;
  float acc4[4] = {0.f, 0.f, 0.f, 0.f}, m_tot = -INFINITY, l_tot = 0.f;
  if (h < G) {
    for (int w = 0; w < kCWarps; ++w) m_tot = fmaxf(m_tot, part[(w * 8 + h) * kPartStride + 128]);
    for (int w = 0; w < kCWarps; ++w) {
      const float* pw = part + (w * 8 + h) * kPartStride;
      const float sc = (pw[128] == -INFINITY) ? 0.f : exp2f((pw[128] - m_tot) * sl2);
      l_tot += pw[129] * sc;
#pragma unroll
      for (int e = 0; e < 4; ++e) acc4[e] += pw[d4 + e] * sc;
    }
  }
  bf16* outp = static_cast<bf16*>(op.out) + (long long)b * op.ldo + (kvh * G + h) * 128 + d4;
  if (nsplit == 1) {
    if (h < G) {
      const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
      *reinterpret_cast<uint2*>(outp) = make_uint2(pack_bf16(acc4[0] * inv, acc4[1] * inv), pack_bf16(acc4[2] * inv, acc4[3] * inv));
    }
    return;
  }
  float* wsb = P.attn_part + ((long long)item * P.nsplit_max) * 8 * kPartStride;
  if (h < G) {
    float* dst = wsb + ((long long)split * 8 + h) * kPartStride;
    __stcg(reinterpret_cast<float2*>(dst + d4), make_float2(acc4[0], acc4[1]));
    __stcg(reinterpret_cast<float2*>(dst + d4 + 2), make_float2(acc4[2], acc4[3]));
    if ((ctid & 31) == 0) {
      __stcg(dst + 128, m_tot);
      __stcg(dst + 129, l_tot);
    }
  }
  __threadfence();
  consumer_sync();
  if (ctid == 0) {
    const unsigned int prev = atomicAdd(P.attn_ctr + item, 1u);
    const int last = (prev == (unsigned int)nsplit - 1) ? 1 : 0;
    if (last) P.attn_ctr[item] = 0u;  // self-reset for the next layer / launch
    c.flag[0] = last;
  }
  consumer_sync();
  const int is_last = c.flag[0];
  consumer_sync();  // flag may be rewritten by a later phase
  if (!is_last) return;
  __threadfence();
  if (h < G) {
    float mt = -INFINITY, lt = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) mt = fmaxf(mt, __ldcg(wsb + ((long long)sp * 8 + h) * kPartStride + 128));
    acc4[0] = acc4[1] = acc4[2] = acc4[3] = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) {
      const float* ps = wsb + ((long long)sp * 8 + h) * kPartStride;
      const float ms = __ldcg(ps + 128);
      const float sc = (ms == -INFINITY) ? 0.f : exp2f((ms - mt) * sl2);
      lt += __ldcg(ps + 129) * sc;
      const float2 o01 = __ldcg(reinterpret_cast<const float2*>(ps + d4));
      const float2 o23 = __ldcg(reinterpret_cast<const float2*>(ps + d4 + 2));
      acc4[0] += o01.x * sc; acc4[1] += o01.y * sc; acc4[2] += o23.x * sc; acc4[3] += o23.y * sc;
    }
    const float inv = lt > 0.f ? 1.f / lt : 0.f;
    *reinterpret_cast<uint2*>(outp) = make_uint2(pack_bf16(acc4[0] * inv, acc4[1] * inv), pack_bf16(acc4[2] * inv, acc4[3] * inv));
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int NB>
__global__ void __launch_bounds__(kMegaThreads, 1) decode_mega_kernel(const MegaPlan* __restrict__ plan) {
  extern __shared__ __align__(128) uint8_t mega_smem[];
  const MegaPlan& P = *plan;
  const int n_ops = P.n_ops;
  // layout: ops | barriers + scratch (2 KB) | region A | ring
  MegaOp* s_ops = reinterpret_cast<MegaOp*>(mega_smem);
  const int ops_bytes = (n_ops * (int)sizeof(MegaOp) + 127) & ~127;
  uint64_t* full = reinterpret_cast<uint64_t*>(mega_smem + ops_bytes);
  volatile uint32_t* gen = reinterpret_cast<volatile uint32_t*>(full + kMaxSlots);  // kMaxSlots counters
  float* red = reinterpret_cast<float*>(const_cast<uint32_t*>(gen) + kMaxSlots + 4);  // 16 floats
  int* flag = reinterpret_cast<int*>(red + 16);             // 4 ints
  float* am_v = reinterpret_cast<float*>(flag + 4);         // [8 warps][16 lanes]
  int* am_i = reinterpret_cast<int*>(am_v + kCWarps * 16);
  uint8_t* region_a = mega_smem + ops_bytes + 2048;
  uint8_t* ring = region_a + P.region_a_bytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {  // copy the op list into shared memory (16-byte chunks), init the ring barriers
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(plan) + sizeof(MegaPlan));
    uint4* dst = reinterpret_cast<uint4*>(s_ops);
    const int n16 = n_ops * (int)sizeof(MegaOp) / 16;
    for (int i = tid; i < n16; i += kMegaThreads) dst[i] = __ldg(src + i);
    if (tid == 0) {
      for (int s = 0; s < P.nslots; ++s) {
        mbar_init(&full[s], 1);
        gen[s] = 0u;
      }
      fence_barrier_init();
    }
  }
  __syncthreads();

  MegaCtx c;
  c.P = plan; c.ops = s_ops; c.full = full; c.gen = gen; c.red = red; c.flag = flag;
  c.region_a = region_a; c.ring = ring; c.nslots = P.nslots; c.cta = blockIdx.x; c.grid = gridDim.x; c.n_ops = n_ops;
  c.bar_target = 0;

  // every warp keeps its own cursor into the stage sequence for the refills it issues
  StageCursor refill;
  refill.op_i = 0; refill.base = 0; refill.cnt = 0; refill.row0 = 0; refill.rows = 0;
  cursor_load(c, refill);
  if (tid == 0) {  // prime the ring: stages 0 .. nslots-1
    StageCursor k = refill;
    for (int s = 0; s < P.nslots; ++s) issue_stage(c, k, (uint32_t)s);
  }

  const int ctid = tid, cw = warp;
  uint32_t sc_base = 0;
  float best_v = -INFINITY;
  int best_i = 0x7fffffff;
#pragma unroll 1
  for (int i = 0; i < n_ops; ++i) {
    const MegaOp& op = s_ops[i];
    if (i > 0) grid_sync(c, ctid);
    if (op.type == OP_GEMV) {
      stage_x<NB>(c, op, ctid);
      int row0, rows;
      slab_of(op, c.cta, c.grid, row0, rows);
      gemv_consume<NB>(c, op, sc_base, row0, rows, cw, lane, refill, best_v, best_i);
      sc_base += (uint32_t)(((rows + op.R - 1) / op.R) * op.ksplit);
      if (op.flags & F_ARGMAX) {
        // CTA-level partial argmax per sequence: lane (i * NB + b) tracked sequence b = lane % NB
        if (lane < 16) {
          am_v[cw * 16 + lane] = best_v;
          am_i[cw * 16 + lane] = best_i;
        }
        consumer_sync();
        if (ctid < NB && ctid < P.B) {
          float bv = -INFINITY;
          int bi = 0x7fffffff;
          for (int w = 0; w < kCWarps; ++w)
            for (int ln = ctid; ln < kRMax * NB && ln < 16; ln += NB) {
              const float v = am_v[w * 16 + ln];
              const int ix = am_i[w * 16 + ln];
              if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
            }
          __stcg(P.amax_val + c.cta * NB + ctid, bv);
          __stcg(P.amax_idx + c.cta * NB + ctid, bi);
        }
        best_v = -INFINITY;
        best_i = 0x7fffffff;
      }
    } else if (op.type == OP_ATTN) {
      attn_phase(c, op, ctid);
    } else if (op.type == OP_FINAL) {
      // greedy sampling: lowest index among the maxima (HF argmax), token history, context lengths
      if (c.cta == 0 && cw == 0) {
        const int pos = P.hist_pos ? __ldcg(P.hist_pos) : 0;
        for (int b = 0; b < P.B; ++b) {
          float bv = -INFINITY;
          int bi = 0x7fffffff;
          for (int k = lane; k < c.grid; k += 32) {
            const float v = __ldcg(P.amax_val + k * NB + b);
            const int ix = __ldcg(P.amax_idx + k * NB + b);
            if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
          }
          if (lane == 0) {
            const long long tok = (long long)bi + P.vocab_offset;
            P.tokens[b] = tok;
            if (P.token_hist && pos < P.hist_capacity) P.token_hist[(long long)pos * P.B + b] = tok;
            P.ctx_lens[b] = P.ctx_lens[b] + 1;
          }
        }
        if (lane == 0 && P.hist_pos) *P.hist_pos = pos + 1;
      }
    }
  }
  // ===================== exit: the last CTA to leave resets the barrier counters for the next launch =====================
  consumer_sync();
  if (ctid == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(P.exit_ctr, 1u);
    if (prev == (unsigned int)c.grid - 1) {
      *P.bar_ctr = 0u;
      *P.exit_ctr = 0u;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
static int pick_ksplit(int K) {
  const int nvec = K / 8;
  for (int d = 1; d <= nvec; ++d)
    if (nvec % d == 0 && (long long)(K / d) * 2 <= kSlotBytes) return d;
  return -1;
}

struct WsLayout {
  long long bar, exitc, err, attn_ctr, amax_val, amax_idx, attn_part, total;
};
static WsLayout ws_layout(int grid) {
  WsLayout w;
  long long off = 0;
  w.bar = off; off += 128;
  w.exitc = off; off += 128;
  w.err = off; off += 128;
  w.attn_ctr = off; off += 4LL * grid;  // one counter per (sequence, kv head) item, items <= grid
  off = (off + 127) & ~127LL;
  w.amax_val = off; off += 4LL * grid * 4;
  w.amax_idx = off; off += 4LL * grid * 4;
  off = (off + 127) & ~127LL;
  w.attn_part = off; off += 4LL * grid * 8 * kPartStride;
  w.total = (off + 127) & ~127LL;
  return w;
}

}  // namespace omc

using namespace omc;

extern "C" long long omc_decode_plan_bytes(int n_layers) {
  if (n_layers < 0) return -1;
  return (long long)sizeof(MegaPlan) + (long long)(5 * n_layers + 2) * (long long)sizeof(MegaOp);
}

extern "C" long long omc_decode_workspace_bytes(int grid) {
  if (grid <= 0) return -1;
  return ws_layout(grid).total;
}

extern "C" int omc_decode_plan_build(const omc_decode_desc* d, void* plan_host) {
  if (d == nullptr || plan_host == nullptr) return set_error(OMC_ERR_ARG, "omc_decode_plan_build: null argument");
  if (d->batch < 1 || d->batch > 4) return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: batch must be 1..4");
  if (d->n_layers < 0 || 5 * d->n_layers + 2 > kMaxOps) return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: too many layers");
  if (d->grid < 1) return set_error(OMC_ERR_ARG, "omc_decode_plan_build: grid must be the number of CTAs (SMs)");
  if (d->kv_heads < 1 || d->q_heads % d->kv_heads != 0 || d->q_heads / d->kv_heads > 8)
    return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: q heads per kv head must be 1..8");
  if (d->batch * d->kv_heads > d->grid) return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: batch * kv_heads exceeds the grid");
  if (d->hidden % 8 != 0 || d->hidden > 4096) return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: hidden must be a multiple of 8, <= 4096");
  if (d->inter % 8 != 0 || d->vocab < 1) return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: inter % 8 != 0 or empty vocab");
  if (d->page_size < 1 || d->max_pages < 1) return set_error(OMC_ERR_ARG, "omc_decode_plan_build: bad paging parameters");
  const int C = d->hidden, Hq = d->q_heads, Hkv = d->kv_heads, I = d->inter, B = d->batch;
  const int qw = (Hq + 2 * Hkv) * 128, aw = Hq * 128;
  if (aw > 4096 * 8) return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: too many heads");
  MegaPlan* P = static_cast<MegaPlan*>(plan_host);
  memset(P, 0, sizeof(MegaPlan));
  MegaOp* ops = reinterpret_cast<MegaOp*>(P + 1);
  int n = 0, kmax = 0;
  bool bad = false;
  auto gemv = [&](const void* W, int N, int K, const void* x, int ldx, const void* norm_w, const void* bias, const void* res,
                  int ldr, void* out, int ldo, int epi, int flags, void* aux) {
    MegaOp& o = ops[n++];
    memset(&o, 0, sizeof(o));
    o.type = OP_GEMV; o.N = N; o.K = K; o.epi = epi; o.flags = flags;
    o.gran = (epi == EPI_SWIGLU) ? 2 : 1;
    o.ksplit = pick_ksplit(K);
    if (o.ksplit < 1 || K % 8 != 0 || N % o.gran != 0) { bad = true; o.ksplit = 1; }
    int R = 1;
    if (o.ksplit == 1) {
      R = kSlotBytes / (K * 2);
      if (R > kRMax) R = kRMax;
      if (o.gran == 2) R &= ~1;
      if (R < o.gran) bad = true;
    } else if (o.gran == 2) bad = true;  // a (gate, up) pair must fit one ring stage
    o.R = R < 1 ? 1 : R;
    if (norm_w != nullptr && K > 4096) bad = true;
    o.W = (const bf16*)W; o.x = (const bf16*)x; o.norm_w = (const bf16*)norm_w; o.bias = (const bf16*)bias;
    o.res = (const bf16*)res; o.out = out; o.aux = (bf16*)aux; o.ldx = ldx; o.ldo = ldo; o.ldr = ldr;
    if (K > kmax) kmax = K;
  };
  for (int li = 0; li < d->n_layers; ++li) {
    gemv(d->qkv_w[li], qw, C, d->h, C, d->ln1[li], d->qkv_b[li], nullptr, C, d->qkv, qw, EPI_NONE,
         li == 0 ? F_X_EMBED : 0, li == 0 ? d->h : nullptr);
    MegaOp& a = ops[n++];
    memset(&a, 0, sizeof(a));
    a.type = OP_ATTN; a.x = (const bf16*)d->qkv; a.ldx = qw; a.out = d->attn; a.ldo = aw;
    a.aux = static_cast<bf16*>(d->kv_pool) + (long long)li * d->kv_layer_stride;
    gemv(d->o_w[li], C, aw, d->attn, aw, nullptr, nullptr, d->h, C, d->h, C, EPI_RES, 0, nullptr);
    gemv(d->gate_up_w[li], 2 * I, C, d->h, C, d->ln2[li], nullptr, nullptr, C, d->act, I, EPI_SWIGLU, 0, nullptr);
    gemv(d->down_w[li], C, I, d->act, I, nullptr, nullptr, d->h, C, d->h, C, EPI_RES, 0, nullptr);
  }
  gemv(d->lm_head, d->vocab, C, d->h, C, d->final_norm, nullptr, nullptr, C, d->logits, d->vocab, EPI_NONE,
       F_OUT_F32 | F_ARGMAX | (d->n_layers == 0 ? F_X_EMBED : 0), nullptr);
  MegaOp& f = ops[n++];
  memset(&f, 0, sizeof(f));
  f.type = OP_FINAL;
  if (bad) return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: a layer shape does not fit the ring (K % 8, SwiGLU pair > slot, norm K > 4096)");
  P->n_ops = n; P->B = B; P->C = C; P->Hq = Hq; P->Hkv = Hkv; P->G = Hq / Hkv;
  P->page_size = d->page_size; P->max_pages = d->max_pages; P->grid = d->grid;
  P->nsplit_max = d->grid / (B * Hkv);
  P->vocab_offset = d->vocab_offset; P->hist_capacity = d->hist_capacity; P->kmax = kmax;
  int region_a = B * kmax * 2;
  if (region_a < kAttnScratchBytes) region_a = kAttnScratchBytes;
  region_a = (region_a + 127) & ~127;
  const int ops_bytes = (n * (int)sizeof(MegaOp) + 127) & ~127;
  int nslots = (kSmemLimit - ops_bytes - 2048 - region_a) / kSlotBytes;
  if (nslots > kMaxSlots) nslots = kMaxSlots;
  if (nslots < 2) return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: activations leave no room for the weight ring");
  P->nslots = nslots; P->region_a_bytes = region_a;
  P->smem_bytes = ops_bytes + 2048 + region_a + nslots * kSlotBytes;
  P->eps = d->eps; P->scale_log2 = d->attn_scale * 1.4426950408889634f;
  P->embed = (const bf16*)d->embed; P->inv_freq = d->inv_freq; P->block_table = d->block_table; P->ctx_lens = d->ctx_lens;
  P->tokens = d->tokens; P->token_hist = d->token_hist; P->hist_pos = d->hist_pos;
  const WsLayout w = ws_layout(d->grid);
  uint8_t* ws = static_cast<uint8_t*>(d->workspace);
  P->bar_ctr = reinterpret_cast<unsigned int*>(ws + w.bar);
  P->exit_ctr = reinterpret_cast<unsigned int*>(ws + w.exitc);
  P->err_flag = d->status ? d->status : reinterpret_cast<int32_t*>(ws + w.err);
  P->attn_ctr = reinterpret_cast<unsigned int*>(ws + w.attn_ctr);
  P->amax_val = reinterpret_cast<float*>(ws + w.amax_val);
  P->amax_idx = reinterpret_cast<int32_t*>(ws + w.amax_idx);
  P->attn_part = reinterpret_cast<float*>(ws + w.attn_part);
  return OMC_OK;
}

template <int NB>
static int launch_mega(const MegaPlan* host, const void* plan_dev, cudaStream_t st) {
  static int attr_smem = 0;
  if (host->smem_bytes > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(decode_mega_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, host->smem_bytes);
    if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
    attr_smem = host->smem_bytes;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(host->grid);
  cfg.blockDim = dim3(kMegaThreads);
  cfg.dynamicSmemBytes = host->smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: the grid barriers depend on it
  attrs[0].val.cooperative = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  const MegaPlan* arg = static_cast<const MegaPlan*>(plan_dev);
  cudaError_t e = cudaLaunchKernelEx(&cfg, decode_mega_kernel<NB>, arg);
  if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
  return OMC_OK;
}

extern "C" int omc_decode_step(const void* plan_host, const void* plan_dev, void* stream) {
  if (plan_host == nullptr || plan_dev == nullptr) return set_error(OMC_ERR_ARG, "omc_decode_step: null plan");
  const MegaPlan* P = static_cast<const MegaPlan*>(plan_host);
  if (P->n_ops < 2 || P->n_ops > kMaxOps || P->grid < 1) return set_error(OMC_ERR_ARG, "omc_decode_step: plan not built");
  cudaStream_t st = (cudaStream_t)stream;
  switch (P->B) {
    case 1: return launch_mega<1>(P, plan_dev, st);
    case 2: return launch_mega<2>(P, plan_dev, st);
    case 3: return launch_mega<3>(P, plan_dev, st);
    case 4: return launch_mega<4>(P, plan_dev, st);
    default: return set_error(OMC_ERR_SHAPE, "omc_decode_step: batch must be 1..4");
  }
}
