// Persistent decode step ("megakernel") for small batches (B <= 4): ONE cooperative launch runs a whole Qwen2 decode step
//   embed -> 28 x [RMSNorm+QKV GEMV, RoPE + KV append + paged GQA attention, O GEMV (+res), RMSNorm+gate/up GEMV
//   (SwiGLU), down GEMV (+res)] -> final RMSNorm + lm_head GEMV -> greedy argmax (+ token history, ctx_lens += 1)
// with one CTA per SM. A decode step is HBM-bound weight streaming (14.1 GB per token for Qwen2-7B in bf16), so the
// kernel is organised around keeping HBM busy across op boundaries:
//   weight ring  this CTA's row slab of every weight matrix, op after op, is one long sequence of "stages" (2 whole rows
//                of K = 3584, or one K chunk of a longer row: <= one 14 KB slot) streamed into a 12-slot shared-memory
//                ring with cp.async.bulk (TMA, mbarrier complete_tx, L2 evict-first). The ring refills itself: the warp
//                that finishes reading a slot immediately issues the copy of the stage that will occupy it next
//                (stage + nslots), whichever op that belongs to. Weights are immutable, so the stream runs ahead across
//                op boundaries while the CTA waits for activations or runs the attention phase. The ring is latency-
//                bound (step time follows the bytes in flight), so the rest of shared memory is kept small: the op list
//                is a 32-entry sliding window, everything the refill path reads sits at a compile-time offset.
//   8 warps      per op: gather the activation vector(s) into shared memory (RMSNorm fused) -> each warp owns whole
//                "units" (R weight rows x a K chunk) of the ring; dot products on the tensor pipe (ldmatrix + mma.sync
//                with the diagonal trick of mma_rows), fp32 accumulation, warp-shuffle reduction, fused epilogue (bias /
//                residual / SwiGLU / fp32 logits + running argmax).
//   no barriers  there is no grid-wide barrier. Every activation that crosses CTAs travels in a flag-in-data buffer (the
//                LL idea of NCCL): each 32-bit word is {16-bit sequence tag, bf16 value} (64-bit {tag, fp32} for the
//                attention partials), written with one store and polled by the readers, so "data has arrived" is the
//                synchronisation and a producer->consumer hand-off costs one L2 round trip. Buffers are double-buffered by
//                layer parity; a writer can only reach a buffer again after every reader of its previous contents has
//                moved two phases on (it needs their outputs first), so tags never alias. The residual stream stays in
//                the shared memory of the CTA that owns those rows (o_proj and down_proj partition rows identically).
// Attention runs inside the same kernel, on the tensor pipe too: (sequence, kv-head) items are split over all CTAs by key
// range, the K/V rows are fetched with cp.async into swizzled tiles before the CTA waits for q, the 7 query heads of a
// group are the rows of mma.sync fragments (S = Q K^T, softmax in registers, O = P V with every warp owning 16 output
// dims), and partial (m, l, O) are merged head by head by designated CTAs that poll the partials (no counters).
// DESIGN.md section 4.3 has the measurements behind each of these choices and the variants that were rejected.
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

#ifndef OMC_MEGA_DETAIL
#define OMC_MEGA_DETAIL 0  // 1: warp 0 also records a clock64 breakdown of the GEMV loop (tools/prof_mega.py)
#endif
constexpr int kMegaThreads = 256;
constexpr int kCWarps = 8;
constexpr int kRMax = 4;            // weight rows per ring stage (upper bound)
constexpr int kSlotBytesDefault = 14336;  // ring slot: 2 rows of K=3584; a K=18944 row travels as 3 chunks (6400|6400|6144)
constexpr int kSlotBytesMax = 32768;
constexpr int kMaxSlots = 16;
constexpr int kMaxSub = 8;                      // K-chunk sub-ops a wide MLP may be cut into
constexpr int kOpsPerLayerMax = 3 + 2 * kMaxSub;
constexpr int kMaxOps = 640;
constexpr int kAttnKeysPerCta = 32;
constexpr int kPartStride = 130;    // O[128], m, l
constexpr int kAttnScratchBytes = (8 + 2 * 48) * 256;  // attention phase: Q tile [8][128] + K and V tiles [48][128] bf16 (26 KB)
constexpr int kHRows = 64;          // residual rows one CTA can own
constexpr int kBiasRows = 128;      // slab rows whose bias is staged in shared memory (larger slabs read it from L2)
constexpr int kMaxBt = 1024;        // block-table entries cached in shared memory (batch * max_pages)
constexpr int kProfStride = 8;       // uint64 per (CTA, op) in the optional profile buffer
constexpr int kMaxTp = 8;           // tensor-parallel ranks one step can span (peer exchange buffers over NVLink)
constexpr int kXpBytes = kHRows * 4 * 4;  // this CTA's own row-parallel partial sums [kHRows][4 sequences] fp32
// shared-memory front: MegaPlan header (256 B) | the kernel's MegaCtx (256 B) | ring barriers + generations | a WINDOW of
// the op list and of the per-CTA slab table (the whole list, 16 KB for 28 layers, would cost the ring a slot: the ops
// live in global memory and a sliding window of kWin entries - op i at index i % kWin - is refreshed one op per iteration)
constexpr int kWin = 32;        // ops [cur, cur + kWin - 3] are valid while op `cur` runs; refills look < 24 ops ahead
constexpr int kOpsOff = 704;
constexpr int kTabOff = kOpsOff + kWin * 96;
constexpr int kHdrBytes = (kTabOff + kWin * 24 + 127) & ~127;
constexpr int kFullOff = 512;   // uint64 full[kMaxSlots]: mbarriers of the ring slots
constexpr int kGenOff = 640;    // uint32 gen[kMaxSlots]: copies issued into each slot so far (monotonic: no parity aliasing)
constexpr int kMetaFixed = 2048 + 4 * kHRows * 2 + kXpBytes;  // barriers/scratch | residual slab | partials | block table (sized per plan)
__host__ __device__ __forceinline__ int bt_bytes_of(int B, int max_pages) { return (B * max_pages * 4 + 127) & ~127; }
constexpr int kSmemLimit = 227 * 1024;
constexpr unsigned long long kWaitLimitNs = 4000000000ull;  // a protocol bug must end in a trap, never in a hung GPU

enum { OP_GEMV = 1, OP_ATTN = 2, OP_FINAL = 3 };
enum { F_OUT_F32 = 1, F_X_EMBED = 2, F_ARGMAX = 4, F_X_KEEP = 8, F_ADD_PART = 16 };
// epilogues of the K-chunk sub-ops of a row-parallel GEMV (private to this kernel; the public OMC_EPI_* stop at 3):
// the first chunk stores the row's partial sum in shared memory, middle chunks add to it, the last chunk (EPI_RES +
// F_ADD_PART) adds it to its own sum before the usual epilogue
enum { EPI_PART_SET = 8, EPI_PART_ADD = 9 };

struct MegaOp {  // 88 bytes (the op list lives in shared memory: every byte here is a byte less of weight ring)
  const bf16* W;
  const bf16* norm_w;
  const bf16* bias;
  const uint32_t* x_ll;  // input vector(s) [B][ldx], flag-in-data words {tag16, bf16}; null with F_X_EMBED
  uint32_t* out_ll;      // output vector(s) [B][ldo] in the same format (null for lm_head)
  float* out_f32;        // lm_head: fp32 logits [B][ldo]
  bf16* pool;            // ATTN: this layer's KV pool
  int32_t N, K, ldx, ldo;
  int32_t kc0;           // elements per K chunk (chunk ks = [ks*kc0, min(K, (ks+1)*kc0)))
  int32_t ldw;           // row pitch of W in elements (= K unless the op is a K-chunk sub-op of a wider matrix)
  int16_t in_op;         // index of the op whose tag the input carries
  uint8_t type, epi, R, ksplit, gran, flags, parity;
  uint8_t xslot;         // 1 + peer exchange slot of a row-parallel op under tensor parallelism (0 = none)
  int32_t pad;
};
static_assert(sizeof(MegaOp) == 96, "MegaOp layout");
// Per-CTA view of a GEMV op, built once per launch in shared memory: this CTA's row slab and where its stages sit in the
// CTA's stage sequence (locating a stage must not cost 64-bit divisions on the refill path)
struct SlabEnt {  // 24 bytes
  const bf16* w0;      // first weight row of the slab
  uint32_t base, cnt;  // stages [base, base + cnt) of this CTA's stream belong to the op (cnt = 0: not a GEMV)
  int32_t row0, rows;
};
static_assert(sizeof(SlabEnt) == 24 && (kWin & (kWin - 1)) == 0, "window layout");

struct MegaPlan {  // header, followed by n_ops MegaOp
  int32_t n_ops, B, C, Hq, Hkv, G, page_size, max_pages;
  int32_t grid, nsplit_max, vocab_offset, hist_capacity, kmax, nslots, region_a_bytes, smem_bytes;
  float eps, scale_log2;
  int32_t pf_stages, slot_bytes;
  int32_t scalar_gemv, pad2[3];  // scalar_gemv: A/B switch, 1 = FFMA dot products instead of the mma.sync path
  const bf16* embed;
  const float* rope_cs;  // [positions][64][2] fp32 (cos, sin)
  const int32_t* block_table;
  int32_t* ctx_lens;
  int64_t* tokens;
  int64_t* token_hist;
  int32_t* hist_pos;
  uint2* attn_part;   // [2][grid][8][130] {fp32 bits, tag32}
  uint2* amax_part;   // [grid][4][2]      {value bits | index, tag32}
  int32_t* err_flag;
  unsigned long long* prof;  // optional [grid][n_ops][8]: globaltimer stamps op start, inputs staged, op end, SM id, then
                             // warp 0's clock64 cycles spent in: stage wait, dot products, refill issue, reduce + epilogue
  int32_t tp_rank, tp_size;
  uint2* xchg[kMaxTp];  // tensor parallelism: rank p's exchange buffer as mapped into THIS process (xchg[tp_rank] = own).
                        // layout (uint2 {fp32 bits | index, tag32}): partials [4 slots][tp][B][C], then argmax [tp][B][2]
};
static_assert(sizeof(MegaPlan) % 16 == 0, "MegaPlan must keep the op array 16-byte aligned");
static_assert(sizeof(MegaPlan) == 256 && sizeof(MegaOp) == 96, "shared-memory front layout (and lib.py's op parser)");

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(kEvictFirst)
      : "memory");
}
__device__ __forceinline__ uint4 ld_poll_v4(const void* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}
// 8-byte flag-in-data word to / from a peer GPU's memory over NVLink (single-copy atomic: value and tag travel together)
__device__ __forceinline__ void st_sys_v2(uint2* p, uint2 v) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint2 ld_sys_v2(const uint2* p) {
  uint2 r;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  return r;
}
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
// wall-clock watchdog for the polling loops: call once per failed poll
struct Watchdog {
  unsigned long long t0;
  unsigned int spins;
  __device__ __forceinline__ Watchdog() : t0(0), spins(0) {}
  __device__ __forceinline__ void tick(int32_t* err_flag, int code, int detail) {
    if ((++spins & 255u) == 0) {
      if (t0 == 0) t0 = global_ns();
      else if (global_ns() - t0 > kWaitLimitNs) mega_fail(err_flag, code, detail);
    }
  }
};
__device__ __forceinline__ void mbar_wait_bounded(uint64_t* bar, uint32_t parity, int32_t* err_flag, int code, int detail) {
  Watchdog wd;
  while (!mbar_try_wait(bar, parity)) wd.tick(err_flag, code, detail);
}

__device__ __forceinline__ uint32_t tag32_of(uint32_t epoch, int op_idx) { return epoch * 1024u + (uint32_t)op_idx + 1u; }
__device__ __forceinline__ uint32_t tag16_of(uint32_t t32) { return ((t32 % 65535u) + 1u) << 16; }  // pre-shifted, never 0
__device__ __forceinline__ bool ll4_ok(uint4 v, uint32_t tag) {
  return ((v.x ^ tag) < 65536u) & ((v.y ^ tag) < 65536u) & ((v.z ^ tag) < 65536u) & ((v.w ^ tag) < 65536u);
}
// 4 flag-in-data words -> 4 packed bf16
__device__ __forceinline__ uint2 ll4_data(uint4 v) {
  return make_uint2((v.x & 0xffffu) | (v.y << 16), (v.z & 0xffffu) | (v.w << 16));
}
__device__ __forceinline__ uint32_t ll4_word(float x, uint32_t tag) {
  return tag | (uint32_t)__bfloat16_as_ushort(__float2bfloat16(x));
}

__device__ __forceinline__ void slab_rows(int N, int gran, int cta, int grid, int& row0, int& rows) {
  const long long ng = N / gran;
  const int g0 = (int)(ng * cta / grid), g1 = (int)(ng * (cta + 1) / grid);
  row0 = g0 * gran;
  rows = (g1 - g0) * gran;
}

// this CTA's slab of a GEMV op and its stage count; `base` = stage index where the op starts in the CTA's stream
__device__ __forceinline__ SlabEnt make_slab(const MegaOp& op, uint32_t base) {
  SlabEnt e;
  e.w0 = nullptr; e.base = base; e.cnt = 0; e.row0 = 0; e.rows = 0;
  if (op.type == OP_GEMV) {
    slab_rows(op.N, op.gran, (int)blockIdx.x, (int)gridDim.x, e.row0, e.rows);
    e.cnt = (uint32_t)(((e.rows + op.R - 1) / op.R) * op.ksplit);
    e.w0 = op.W + (size_t)e.row0 * op.ldw;
  }
  return e;
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

__device__ __forceinline__ void unpack8(uint4 v, float* f) {
  float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void unpack4(uint2 v, float* f) {
  const float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
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

// The dynamic shared memory of the kernel. Everything the refill path touches sits at an offset known at compile time or
// kept in a register (RingHot): after every mbarrier / bulk-copy asm (memory clobber) the compiler must reload whatever it
// reads through a pointer, and a pointer that itself lives in shared memory makes that a chain of dependent LDS behind
// the ldmatrix traffic of the other warps - measured ~0.5 us per refill before this layout.
extern __shared__ __align__(128) uint8_t mega_smem[];
struct RingHot {  // by value, in registers
  int ring_off;  // byte offset of ring slot 0
  int slot_bytes, nslots, n_ops;
};
__device__ __forceinline__ uint64_t* ring_full(uint32_t slot) { return reinterpret_cast<uint64_t*>(mega_smem + kFullOff) + slot; }
__device__ __forceinline__ volatile uint32_t* ring_gen(uint32_t slot) {
  return reinterpret_cast<volatile uint32_t*>(mega_smem + kGenOff) + slot;
}
__device__ __forceinline__ MegaOp* win_ops() { return reinterpret_cast<MegaOp*>(mega_smem + kOpsOff); }    // [kWin]
__device__ __forceinline__ SlabEnt* win_tab() { return reinterpret_cast<SlabEnt*>(mega_smem + kTabOff); }  // [kWin]

// Position of a stage index in the op list: the op that owns it (a hint that only moves forward; n_ops = exhausted)
struct StageCursor {
  int op_i;
};

struct MegaCtx {
  const MegaPlan* P;     // shared copy of the plan header
  const MegaOp* gops;    // the whole op list in global memory (the window is refreshed from it)
  float* red;            // [16] floats of block-reduction scratch
  float* am_v;           // [8 warps][16]
  int* am_i;
  int* s_ctx;            // [4] context lengths at kernel start
  bf16* s_h;             // [4][kHRows] residual rows owned by this CTA
  float* s_bias;         // [kBiasRows] bias of this CTA's slab for the current op (fetched while waiting for x)
  int* s_bt;             // block table copy [B][max_pages]
  float* s_xp;           // [kHRows][4] own partial sums of a row-parallel op (tensor parallelism)
  uint8_t* region_a;     // activation vectors / attention scratch
  int cta, grid, n_ops, pf_stages, scalar_gemv;
  unsigned long long* prof_op;  // this CTA's profile record of the current op (null = profiling off)
  uint32_t epoch;
};

// Locate stage s: advance cursor k (stage indices passed through one cursor are strictly increasing) and return the
// global source address and byte count of the stage, or false when the op list is exhausted.
static_assert(sizeof(MegaCtx) <= 256, "MegaCtx must fit its 256-byte shared-memory slot");

__device__ __forceinline__ bool locate_stage(const RingHot& h, int cur_op, StageCursor& k, uint32_t s, const bf16*& src,
                                             uint32_t& bytes, int& ncopies, uint32_t& pitch_bytes) {
  int oi = max(k.op_i, cur_op);  // entries of finished ops have been recycled: never look behind the running op
  uint32_t base = 0, cnt = 0;
  const SlabEnt* tab = win_tab();
  const int hi = min(h.n_ops, cur_op + kWin - 2);
  while (oi < hi) {
    const uint2 bc = *reinterpret_cast<const uint2*>(&tab[oi & (kWin - 1)].base);
    base = bc.x; cnt = bc.y;
    if (s < base + cnt) break;
    ++oi;
  }
  k.op_i = oi;
  if (oi >= h.n_ops) return false;
  if (oi >= hi) __trap();  // the stage lies beyond the op window (host-side checks make this unreachable)
  const MegaOp& op = win_ops()[oi & (kWin - 1)];
  const SlabEnt& e = tab[oi & (kWin - 1)];
  const uint32_t rel = s - base;
  const int ksplit = op.ksplit, R = op.R, K = op.K, ldw = op.ldw;
  int u = (int)rel, ks = 0;
  if (ksplit > 1) {
    // split-K stage order: unit-major (the chunks of one row are consecutive stages, consumed back to back by one warp).
    // Tried: chunk-major inside groups of 8 units so that consecutive stages go to different warps - 2.5 % slower (the
    // op's end skew grew from 4.3 to 7 us); and the ring depth matters in whole units: 12 slots = 4 rows of 3 chunks, 10
    // slots cost 48 % of down_proj's time.
    u = (int)(rel / (uint32_t)ksplit);
    ks = (int)rel - u * ksplit;
  }
  const int r = u * R;
  const int rows_here = min(R, e.rows - r);
  const int kbeg = ks * op.kc0, Kc = min(op.kc0, K - kbeg);
  // whole rows of a dense matrix are one contiguous copy; a K chunk of several rows is one copy per row (row pitch ldw)
  const bool dense = (ksplit == 1) & (ldw == K);
  ncopies = dense ? 1 : rows_here;
  bytes = (uint32_t)((dense ? rows_here : 1) * Kc) * 2u;
  pitch_bytes = (uint32_t)ldw * 2u;
  src = e.w0 + ((size_t)(uint32_t)r * (uint32_t)ldw + (uint32_t)kbeg);
  return true;
}
// Issue the bulk copy of stage s (if it exists) into ring slot s % nslots, and ask L2 for the stage `pf_stages` further
// down the stream: the shared-memory ring bounds how far the copies can run ahead (~150 KB/SM), the L2 prefetch lets HBM
// keep streaming through the latency-bound phases (qkv -> attention -> o_proj) up to ~60 MB chip-wide ahead of use.
// Called by ONE lane with that warp's cursors.
__device__ __forceinline__ void issue_stage(const RingHot& h, int cur_op, int pf_stages, StageCursor& k, StageCursor& kpf,
                                            uint32_t s) {
  const bf16* src;
  uint32_t bytes, pitch;
  int ncopies;
  if (pf_stages > 0 && locate_stage(h, cur_op, kpf, s + (uint32_t)pf_stages, src, bytes, ncopies, pitch))
    for (int i = 0; i < ncopies; ++i)
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const uint8_t*>(src) + (size_t)i * pitch),
                   "r"(bytes)
                   : "memory");
  if (!locate_stage(h, cur_op, k, s, src, bytes, ncopies, pitch)) return;
  const uint32_t g = s / (uint32_t)h.nslots, slot = s - g * (uint32_t)h.nslots;
  fence_proxy_async();  // generic-proxy reads of this slot (previous stage) are ordered before the async-proxy refill
  mbar_arrive_expect_tx(ring_full(slot), bytes * (uint32_t)ncopies);
  uint8_t* dst = mega_smem + h.ring_off + slot * (uint32_t)h.slot_bytes;
  for (int i = 0; i < ncopies; ++i)
    bulk_g2s(dst + (size_t)i * bytes, reinterpret_cast<const uint8_t*>(src) + (size_t)i * pitch, bytes, ring_full(slot));
  __threadfence_block();
  *ring_gen(slot) = g + 1u;  // publish: the barrier is now in the phase that carries stage s
}
// Wait until stage s has landed in its slot. A slot is shared by stages s, s + nslots, ... that different warps consume,
// and an mbarrier parity wait is only meaningful for the phase in flight, so first wait (monotonic counter) until the
// copy of stage s has actually been issued, then for its bytes.
__device__ __forceinline__ void wait_stage(const MegaCtx& c, uint32_t s, uint32_t g, uint32_t slot) {
  Watchdog wd;
  while (*ring_gen(slot) < g + 1u) wd.tick(c.P->err_flag, 4, (int)s);
  __threadfence_block();
  mbar_wait_bounded(ring_full(slot), g & 1u, c.P->err_flag, 3, (int)s);
}

// ------------------------------------------------------------------------------------------------ activation staging
// Gather x[b, :K] into shared memory as bf16, optionally RMS-normalised: out = w * bf16(x * rsqrt(mean(x^2) + eps))
// [modeling_qwen2.py:258-263]. The source is either the embedding row of this step's token (first op) or a
// flag-in-data vector another phase is producing right now: polling it IS the synchronisation with the producers.
template <int NB>
__device__ void stage_x(MegaCtx& c, const MegaOp& op, int op_idx, int ctid) {
  const MegaPlan& P = *c.P;
  const int K = op.K;
  bf16* xs = reinterpret_cast<bf16*>(c.region_a);
  const uint32_t tag = (op.flags & F_X_EMBED) ? 0u : tag16_of(tag32_of(c.epoch, op.in_op));
  // bias of this CTA's slab and the norm weights (K <= 4096: at most 4 chunks per thread) are fetched before the wait so
  // that they do not cost a dependent round trip later
  if (op.bias != nullptr) {
    const int r0 = win_tab()[op_idx & (kWin - 1)].row0, nr = win_tab()[op_idx & (kWin - 1)].rows;
    if (ctid < nr && ctid < kBiasRows) c.s_bias[ctid] = __bfloat162float(op.bias[r0 + ctid]);
  }
  // (the norm weights go to shared memory behind the activation vectors with cp.async: no registers held across the wait)
  const uint2* gs = reinterpret_cast<const uint2*>(xs + (long long)NB * K);
  if (op.norm_w != nullptr) {
    for (int i = ctid; i < (K >> 3); i += kMegaThreads)
      cp_async16(const_cast<uint2*>(gs) + 2 * i, op.norm_w + 8 * i, true);
    cp_async_commit();
  }
#pragma unroll 1
  for (int b = 0; b < NB; ++b) {
    if (b >= P.B) break;
    uint2* dst = reinterpret_cast<uint2*>(xs + (long long)b * K);  // 4 bf16 per element
    const int n4 = K >> 2;
    float ss = 0.f;
    if (op.flags & F_X_EMBED) {
      const uint2* src = reinterpret_cast<const uint2*>(P.embed + (long long)P.tokens[b] * P.C);
      for (int i = ctid; i < n4; i += kMegaThreads) {
        const uint2 v = __ldg(src + i);
        dst[i] = v;
        float f[4];
        unpack4(v, f);
        ss += f[0] * f[0] + f[1] * f[1] + f[2] * f[2] + f[3] * f[3];
      }
      // seed the residual rows this CTA owns (the slab of the [C, *] row-parallel ops)
      int h0, hr;
      slab_rows(P.C, 1, c.cta, c.grid, h0, hr);
      if (ctid < hr) c.s_h[b * kHRows + ctid] = P.embed[(long long)P.tokens[b] * P.C + h0 + ctid];
    } else {
      const uint4* src = reinterpret_cast<const uint4*>(op.x_ll + (long long)b * op.ldx);
      constexpr int UN = 10;
#pragma unroll 1
      for (int base = ctid; base < n4; base += kMegaThreads * UN) {
        uint4 v[UN];
        Watchdog wd;
        bool ok;
        do {
          ok = true;
#pragma unroll
          for (int u = 0; u < UN; ++u) {
            const int i = base + u * kMegaThreads;
            if (i < n4) v[u] = ld_poll_v4(src + i);
          }
#pragma unroll
          for (int u = 0; u < UN; ++u) {
            const int i = base + u * kMegaThreads;
            if (i < n4 && !ll4_ok(v[u], tag)) ok = false;
          }
          if (!ok) wd.tick(P.err_flag, 5, op.in_op);  // (a nanosleep back-off here was measured: no change)
        } while (!ok);
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int i = base + u * kMegaThreads;
          if (i < n4) {
            const uint2 d = ll4_data(v[u]);
            dst[i] = d;
            float f[4];
            unpack4(d, f);
            ss += f[0] * f[0] + f[1] * f[1] + f[2] * f[2] + f[3] * f[3];
          }
        }
      }
    }
    if (op.norm_w != nullptr) {
      ss = warp_sum(ss);
      cp_async_wait<0>();
      __syncthreads();  // c.red free, raw vector and norm weights fully in shared memory
      if ((ctid & 31) == 0) c.red[ctid >> 5] = ss;
      __syncthreads();
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < kCWarps; ++w) tot += c.red[w];
      const float rstd = rsqrtf(tot / (float)K + P.eps);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = ctid + j * kMegaThreads;
        if (i < n4) {
          const uint2 g = gs[i], v = dst[i];
          const float2 a0 = unpack_bf16(v.x), a1 = unpack_bf16(v.y), g0 = unpack_bf16(g.x), g1 = unpack_bf16(g.y);
          const float2 n0 = unpack_bf16(pack_bf16(a0.x * rstd, a0.y * rstd)), n1 = unpack_bf16(pack_bf16(a1.x * rstd, a1.y * rstd));
          dst[i] = make_uint2(pack_bf16(n0.x * g0.x, n0.y * g0.y), pack_bf16(n1.x * g1.x, n1.y * g1.y));
        }
      }
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ GEMV consumer
// Dot products of one or two weight rows (a K chunk of `nsteps` x 128 elements each, in the ring) with the staged
// activation vector(s) on the tensor pipe. The scalar form (LDS.128 + 8 bf16->fp32 unpacks + 8 FFMA per 16 weight bytes)
// costs ~45 issue slots per 512 B of weights and made the short ops (qkv, o_proj: data already in the ring when their
// input arrives) instruction-bound; here 512 B of weights cost one ldmatrix + one mma.sync:
//   A (16 x 16, row-major) = rows 0-7: the eight 16-byte pieces [0,8) of a 128-element group of weight row r0 (k 0-7) and
//                            pieces [8,16) (k 8-15); rows 8-15: the same pieces of row r1
//   B (16 x 8, "col")      = column n: pieces n and 8 + n of the same 128-element group of x
// so D[m][m] (m < 8) is the partial dot product of row r0 over pieces m and 8 + m, D[8 + m][m] that of row r1; the other
// 112 entries of D are discarded - the tensor pipe is idle anyway and the point is the issue slots. Each 8x8 ldmatrix
// tile is 128 contiguous bytes of shared memory: conflict-free. fp32 accumulation inside the MMA (as in the GEMMs).
// acc0[b] / acc1[b] receive this lane's share (sum over lanes = the dot product, reduced later with warp_sum).
template <int NB>
__device__ __forceinline__ void mma_rows(uint32_t w_addr, uint32_t row_pitch, bool two, uint32_t x_addr, uint32_t x_pitch,
                                         int nsteps, int lane, float* acc0, float* acc1) {
  const int mat = lane >> 3, ri = lane & 7;
  const uint32_t a_addr = w_addr + (((mat & 1) && two) ? row_pitch : 0u) + (uint32_t)((mat >> 1) * 128 + ri * 16);
  const uint32_t b_addr = x_addr + (uint32_t)lane * 16u;  // tiles: step s pieces 0-7, 8-15, step s+1 pieces 0-7, 8-15
  float d[2][NB][4];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int b = 0; b < NB; ++b) d[h][b][0] = d[h][b][1] = d[h][b][2] = d[h][b][3] = 0.f;
  int s = 0;
#pragma unroll 2
  for (; s + 1 < nsteps; s += 2) {
    uint32_t a0[4], a1[4];
    ldmatrix_x4(a_addr + (uint32_t)s * 256u, a0[0], a0[1], a0[2], a0[3]);
    ldmatrix_x4(a_addr + (uint32_t)s * 256u + 256u, a1[0], a1[1], a1[2], a1[3]);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      uint32_t x0, x1, x2, x3;
      ldmatrix_x4(b_addr + (uint32_t)b * x_pitch + (uint32_t)s * 256u, x0, x1, x2, x3);
      mma_bf16_16816(d[0][b], a0, x0, x1);
      mma_bf16_16816(d[1][b], a1, x2, x3);
    }
  }
  if (s < nsteps) {  // odd tail (the second half of the x tiles read past the chunk: still inside shared memory, unused)
    uint32_t a0[4];
    ldmatrix_x4(a_addr + (uint32_t)s * 256u, a0[0], a0[1], a0[2], a0[3]);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      uint32_t x0, x1, x2, x3;
      ldmatrix_x4(b_addr + (uint32_t)b * x_pitch + (uint32_t)s * 256u, x0, x1, x2, x3);
      mma_bf16_16816(d[0][b], a0, x0, x1);
    }
  }
  // this lane holds D[g][2t], D[g][2t+1], D[g+8][2t], D[g+8][2t+1] (g = lane / 4, t = lane % 4): keep the diagonal
  const int g = lane >> 2, t = lane & 3;
  const bool on = (g >> 1) == t, odd = g & 1;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const float v0 = odd ? d[0][b][1] + d[1][b][1] : d[0][b][0] + d[1][b][0];
    const float v1 = odd ? d[0][b][3] + d[1][b][3] : d[0][b][2] + d[1][b][2];
    if (on) {
      acc0[b] += v0;
      if (two) acc1[b] += v1;
    }
  }
}

template <int NB, bool TP>
__device__ void gemv_consume(MegaCtx& c, const RingHot& h, const MegaOp& op, int op_idx, uint32_t sc_base, int row0, int rows,
                             int cw, int lane, StageCursor& refill, StageCursor& refill_pf, float& best_v, int& best_i) {
  const MegaPlan& P = *c.P;
  const int R = op.R, ksplit = op.ksplit;
  // loop invariants the epilogue and the refill need: read them from shared memory once, not once per unit
  const int epi = op.epi, oflags = op.flags, ldo = op.ldo, PB = P.B, pf_stages = c.pf_stages;
  const bool scalar = c.scalar_gemv != 0;
  const bf16* const bias = op.bias;
  uint32_t* const out_ll = op.out_ll;
  float* const out_f32 = op.out_f32;
  bf16* const s_h = c.s_h;
  const float* const s_bias = c.s_bias;
  const int nv0 = op.kc0 >> 3, nvK = op.K >> 3;
  const int units = (rows + R - 1) / R;
  const uint4* xs = reinterpret_cast<const uint4*>(c.region_a);
  const uint32_t otag = tag16_of(tag32_of(c.epoch, op_idx));
  // row-parallel op under tensor parallelism (o_proj / down_proj): partial sums are exchanged between the GPUs
  // (compiled out of the single-GPU instantiation: the weight-streaming loop below is register-bound)
  const bool xon = TP && op.xslot > 0 && epi == EPI_RES;
  const uint32_t xtag = TP ? tag32_of(c.epoch, op_idx) : 0u;
  const long long xslot_base = TP ? (long long)(op.xslot - 1) * P.tp_size : 0;                 // [slot][src rank][B][C]
  const long long xoff_w = TP ? (xslot_base + P.tp_rank) * (long long)P.B * P.C : 0;            // where peers find OUR partials
#if OMC_MEGA_DETAIL
  const bool pw = c.prof_op != nullptr && cw == 0 && lane == 0;
#else
  constexpr bool pw = false;  // the cycle breakdown costs registers this kernel does not have: compile-time switch
#endif
  long long pc_wait = 0, pc_dot = 0, pc_issue = 0, pc_epi = 0;
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
      const uint32_t sg = s / (uint32_t)h.nslots, slot = s - sg * (uint32_t)h.nslots;
      long long tk0 = 0;
      if (pw) tk0 = clock64();
      wait_stage(c, s, sg, slot);
      if (pw) { const long long t = clock64(); pc_wait += t - tk0; tk0 = t; }
      const uint4* wb = reinterpret_cast<const uint4*>(mega_smem + h.ring_off + slot * (uint32_t)h.slot_bytes);
      const uint4* xb = xs + ks * nv0;
      const int nv = min(nv0, nvK - ks * nv0);  // this chunk's length in 16-byte vectors (the last chunk may be shorter)
      if ((nv & 15) == 0 && !scalar) {
        // tensor-pipe path: the chunk is a whole number of 128-element groups
        const uint32_t w_addr = smem_u32(wb), x_addr = smem_u32(xb);
#pragma unroll
        for (int i = 0; i < kRMax; i += 2)
          if (i < rows_here)
            mma_rows<NB>(w_addr + (uint32_t)(i * nv) * 16u, (uint32_t)nv * 16u, i + 1 < rows_here, x_addr, (uint32_t)nvK * 16u,
                         nv >> 4, lane, acc[i], acc[i + 1]);
      } else if (rows_here == 2) {
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
        float a2[NB];  // second accumulator chain: two independent FMA streams per lane
#pragma unroll
        for (int b = 0; b < NB; ++b) a2[b] = 0.f;
        int j = lane;
#pragma unroll 2
        for (; j + 32 < nv; j += 64) {
          const uint4 w0 = wb[j], w1 = wb[j + 32];
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            float xf[8];
            unpack8(xb[b * nvK + j], xf);
            acc[0][b] = dot8(w0, xf, acc[0][b]);
            unpack8(xb[b * nvK + j + 32], xf);
            a2[b] = dot8(w1, xf, a2[b]);
          }
        }
        for (; j < nv; j += 32) {
          const uint4 w0 = wb[j];
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            float xf[8];
            unpack8(xb[b * nvK + j], xf);
            acc[0][b] = dot8(w0, xf, acc[0][b]);
          }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[0][b] += a2[b];
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
      if (pw) { const long long t = clock64(); pc_dot += t - tk0; tk0 = t; }
      if (lane == 0) issue_stage(h, op_idx, pf_stages, refill, refill_pf, s + (uint32_t)h.nslots);  // slot free: refill it
      if (pw) pc_issue += clock64() - tk0;
    }
    long long te0 = 0;
    if (pw) te0 = clock64();
#pragma unroll
    for (int i = 0; i < kRMax; ++i)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[i][b] = warp_sum(acc[i][b]);

    // ---- epilogue: lane (i * NB + b) owns output (row r + i, sequence b)
    const int lrow = r;            // row index inside this CTA's slab
    const int grow = row0 + r;     // global row
    if (epi == EPI_SWIGLU) {
#pragma unroll
      for (int pr = 0; pr < kRMax / 2; ++pr)
#pragma unroll
        for (int b = 0; b < NB; ++b)
          if (2 * pr < rows_here && b < PB && lane == pr * NB + b) {
            const float val = silu_m(acc[2 * pr][b]) * acc[2 * pr + 1][b];
            __stcg(out_ll + (long long)b * ldo + (grow >> 1) + pr, ll4_word(val, otag));
          }
    } else {
#pragma unroll
      for (int i = 0; i < kRMax; ++i)
#pragma unroll
        for (int b = 0; b < NB; ++b)
          if (i < rows_here && b < PB && lane == i * NB + b) {
            const int row = grow + i;
            float val = acc[i][b];
            if (bias) val += (lrow + i < kBiasRows) ? s_bias[lrow + i] : __bfloat162float(bias[row]);
            if (epi >= EPI_PART_SET) {  // K-chunk sub-op: keep the row's running sum in shared memory
              float* pp = c.s_xp + (lrow + i) * 4 + b;
              *pp = (epi == EPI_PART_SET) ? val : *pp + val;
              continue;
            }
            if (oflags & F_ADD_PART) val += c.s_xp[(lrow + i) * 4 + b];
            if (TP && xon) {
              // tensor parallelism: this is rank tp_rank's PARTIAL sum over its K shard. Push it into every peer's
              // exchange buffer over NVLink ({fp32, tag} in one 8-byte store) and keep the own copy; the all-reduce is
              // finished below, after the warp has consumed all its units (the weight ring keeps streaming meanwhile).
              const uint2 wv = make_uint2(__float_as_uint(val), xtag);
              const long long off = xoff_w + (long long)b * P.C + row;
              for (int p = 0; p < P.tp_size; ++p)
                if (p != P.tp_rank) st_sys_v2(P.xchg[p] + off, wv);
              c.s_xp[(lrow + i) * 4 + b] = val;
              continue;
            }
            if (epi == EPI_RES) {
              // residual stream: bf16, resident in the shared memory of the CTA that owns the row
              bf16* hp = s_h + b * kHRows + lrow + i;
              val += __bfloat162float(*hp);
              *hp = __float2bfloat16(val);
            }
            if (oflags & F_OUT_F32) {
              out_f32[(long long)b * ldo + row] = val;
              if ((oflags & F_ARGMAX) && (val > best_v || (val == best_v && row < best_i))) {
                best_v = val;
                best_i = row;
              }
            } else {
              __stcg(out_ll + (long long)b * ldo + row, ll4_word(val, otag));
            }
          }
    }
    if (pw) pc_epi += clock64() - te0;
  }
  if (pw) {
    c.prof_op[4] = (unsigned long long)pc_wait;
    c.prof_op[5] = (unsigned long long)pc_dot;
    c.prof_op[6] = (unsigned long long)pc_issue;
    c.prof_op[7] = (unsigned long long)pc_epi;
  }
  if (TP && xon) {
    // ---- finish the all-reduce: h[row] += sum over ranks (fixed rank order -> bit-identical replicas on every GPU) of
    // the partial sums; the peers' values are polled out of OUR exchange buffer (they pushed them), one lane per
    // (row, sequence). Then the new residual row is broadcast to this GPU's CTAs like in the single-GPU path.
    __syncwarp();
    const uint2* xin = P.xchg[P.tp_rank];
#pragma unroll 1
    for (int u = cw; u < units; u += kCWarps) {
      const int r = u * R;
      const int rows_here = min(R, rows - r);
      const int i = lane / NB, b = lane - i * NB;
      if (i < rows_here && b < P.B) {
        const int row = row0 + r + i;
        float tot = 0.f;
        for (int p = 0; p < P.tp_size; ++p) {
          float v;
          if (p == P.tp_rank) {
            v = c.s_xp[(r + i) * 4 + b];
          } else {
            const uint2* src = xin + ((xslot_base + p) * (long long)P.B + b) * P.C + row;
            uint2 w;
            Watchdog wd;
            while (true) {
              w = ld_sys_v2(src);
              if (w.y == xtag) break;
              wd.tick(P.err_flag, 11, p);
            }
            v = __uint_as_float(w.x);
          }
          tot += v;
        }
        bf16* hp = c.s_h + b * kHRows + r + i;
        tot += __bfloat162float(*hp);
        *hp = __float2bfloat16(tot);
        __stcg(op.out_ll + (long long)b * op.ldo + row, ll4_word(tot, otag));
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
// Every global access here costs a full L2 round trip under the weight stream, so the phase is organised as few
// dependent round trips as possible: [K/V + rotary table] (issued before the wait) -> [q poll] -> compute -> [partials of
// all splits polled in one batch by the merging CTA] -> output.
// byte offset of 16-byte chunk `chunk` (0..15) of row `row` in a [rows][128] bf16 tile, XOR-swizzled (ldmatrix conflict-free)
__device__ __forceinline__ uint32_t swz128(int row, int chunk) { return (uint32_t)(row * 256 + ((chunk ^ (row & 7)) << 4)); }

constexpr int kAttnTile = 48;  // keys per shared-memory tile (one tile covers contexts up to ~1700 on 148 CTAs)

template <int G>
__device__ __noinline__ void attn_phase(MegaCtx& c, const MegaOp& op, int op_idx, int ctid) {
  // Tensor-pipe formulation (the scalar one - warp per key, 35 shuffles + 14 exp2 per key and warp - was issue-bound at
  // ~0.8 us per key): the G query heads of the kv group are the M rows of mma.sync.m16n8k16 (rows G..15 are padding), the
  // CTA's keys the N columns of S = Q K^T and the K rows of O = P V. Every warp computes the whole S (<= 64 keys, the tensor
  // pipe is idle anyway) and owns 16 of the 128 output dims, so there is no merge between the warps.
  const MegaPlan& P = *c.P;
  const int items = P.B * P.Hkv;
  const int item = c.cta % items, split = c.cta / items;
  if (split >= P.nsplit_max) return;
  const int b = item / P.Hkv, kvh = item % P.Hkv;
  const int n_cached = c.s_ctx[b];  // keys already in the cache; the new token sits at position n_cached
  int nsplit = (n_cached + kAttnKeysPerCta - 1) / kAttnKeysPerCta;
  nsplit = max(1, min(nsplit, P.nsplit_max));
  if (split >= nsplit) return;
  const int per = (n_cached + nsplit - 1) / nsplit;
  const int k0 = split * per, k1 = min(n_cached, k0 + per);
  const bool has_new = (split == nsplit - 1);
  const int nk_all = max(k1 - k0, 0) + (has_new ? 1 : 0);  // keys this CTA attends to (>= 1 when has_new)

  const int cw = ctid >> 5, lane = ctid & 31, g = lane >> 2, tq = lane & 3;
  bf16* pool = op.pool;
  const int* bt = c.s_bt + b * P.max_pages;
  const int page_size = P.page_size, Hkv = P.Hkv;
  const long long page_stride = 2LL * Hkv * page_size * 128;
  const long long v_off = (long long)Hkv * page_size * 128;
  const float sl2 = P.scale_log2;
  const uint32_t itag = tag16_of(tag32_of(c.epoch, op.in_op));
  const uint32_t otag16 = tag16_of(tag32_of(c.epoch, op_idx));
  const uint32_t otag32 = tag32_of(c.epoch, op_idx);

  uint8_t* sQ = c.region_a;                    // [8][128] bf16, swizzled (2 KB)
  uint8_t* sK = sQ + 8 * 256;                  // [64][128]
  uint8_t* sV = sK + kAttnTile * 256;          // [64][128]
  // cached keys [kt0, kt0 + cnt) of this CTA's range -> tile rows 0..cnt-1 (cp.async, 16 bytes per request); rows up to the
  // next multiple of 16 are zero-filled (P = 0 there, but 0 * garbage must stay 0). `skip_row`: the row the new token's
  // K/V will be written to by hand.
  auto load_tile = [&](int kt0, int cnt, int skip_row) {
    const int rows16 = (max(cnt, skip_row + 1) + 15) & ~15;
    for (int i = ctid; i < rows16 * 32; i += kMegaThreads) {
      const int row = i >> 5, part = i & 31, chunk = part & 15;
      if (row == skip_row) continue;
      const bool ok = row < cnt;
      const int kk = kt0 + (ok ? row : 0);
      const bf16* src = pool + (long long)bt[kk / page_size] * page_stride + ((long long)kvh * page_size + kk % page_size) * 128 +
                        chunk * 8 + ((part & 16) ? v_off : 0);
      cp_async16(((part & 16) ? sV : sK) + swz128(row, chunk), ok ? src : pool, ok);
    }
    cp_async_commit();
  };
  const int ntiles = (nk_all + kAttnTile - 1) / kAttnTile;
  {
    // first tile: the cache does not depend on this step's qkv, fetch before waiting for q
    const int cnt = min(k1 - k0, kAttnTile);
    load_tile(k0, max(cnt, 0), (has_new && ntiles == 1) ? nk_all - 1 : -1);
  }
  // rotary factors of the new position for this lane's 4 dims: table of (cos, sin)(pos * inv_freq) computed in fp32 by
  // the host exactly as Qwen2RotaryEmbedding does (modeling_qwen2.py:102-113)
  float cs[4], sn[4];
  {
    const float4* t = reinterpret_cast<const float4*>(P.rope_cs + ((long long)n_cached * 64 + (lane & 15) * 4) * 2);
    const float4 t0 = __ldg(t), t1 = __ldg(t + 1);
    cs[0] = t0.x; sn[0] = t0.y; cs[1] = t0.z; sn[1] = t0.w;
    cs[2] = t1.x; sn[2] = t1.y; cs[3] = t1.z; sn[3] = t1.w;
  }
  const float sgn = (lane < 16) ? -1.f : 1.f;
  // RoPE (rotate-half: dims (i, i + 64) = lanes (l, l ^ 16)) of 4 dims per lane; result rounded to bf16 like the reference
  auto rot = [&](uint4 own_ll) -> uint2 {  // all 32 lanes must call (shuffle)
    const uint2 d = ll4_data(own_ll);
    const uint2 pd = make_uint2(__shfl_xor_sync(0xffffffffu, d.x, 16), __shfl_xor_sync(0xffffffffu, d.y, 16));
    float own[4], par[4], r[4];
    unpack4(d, own);
    unpack4(pd, par);
#pragma unroll
    for (int e = 0; e < 4; ++e) r[e] = own[e] * cs[e] + sgn * par[e] * sn[e];
    return make_uint2(pack_bf16(r[0], r[1]), pack_bf16(r[2], r[3]));
  };

  // ---- wait for q: warp h < G polls head h (this lane's 4 dims), rotates it and parks it in the Q tile; warp 7 of the CTA
  // that owns the new token does the same for k (rotated) and v and appends both to the cache
  const uint32_t* qll = op.x_ll + (long long)b * op.ldx;
  const int r_new = (nk_all - 1) % kAttnTile;  // tile row of the new token (in the last tile)
  if (cw < G) {
    uint4 qo;
    Watchdog wd;
    while (true) {
      qo = ld_poll_v4(qll + (kvh * G + cw) * 128 + lane * 4);
      if (ll4_ok(qo, itag)) break;
      wd.tick(P.err_flag, 6, op.in_op);
    }
    const uint2 qr = rot(qo);
    *reinterpret_cast<uint2*>(sQ + swz128(cw, lane >> 1) + (lane & 1) * 8) = qr;
  }
  uint2 knew_r = make_uint2(0u, 0u), vnew_r = make_uint2(0u, 0u);
  if (has_new && cw == kCWarps - 1) {
    uint4 ko, vo;
    Watchdog wd;
    while (true) {
      ko = ld_poll_v4(qll + (P.Hq + kvh) * 128 + lane * 4);
      vo = ld_poll_v4(qll + (P.Hq + Hkv + kvh) * 128 + lane * 4);
      if (ll4_ok(ko, itag) && ll4_ok(vo, itag)) break;
      wd.tick(P.err_flag, 6, op.in_op);
    }
    knew_r = rot(ko);
    vnew_r = ll4_data(vo);
    bf16* kp = pool + (long long)bt[n_cached / page_size] * page_stride + ((long long)kvh * page_size + n_cached % page_size) * 128 +
               lane * 4;
    *reinterpret_cast<uint2*>(kp) = knew_r;
    *reinterpret_cast<uint2*>(kp + v_off) = vnew_r;
    if (ntiles == 1) {
      *reinterpret_cast<uint2*>(sK + swz128(r_new, lane >> 1) + (lane & 1) * 8) = knew_r;
      *reinterpret_cast<uint2*>(sV + swz128(r_new, lane >> 1) + (lane & 1) * 8) = vnew_r;
    }
  }
#if OMC_MEGA_DETAIL
  if (c.prof_op && ctid == 0) c.prof_op[4] = global_ns();  // q arrived
#endif
  cp_async_wait<0>();
  __syncthreads();

  uint32_t qf[8][4];
  {
    const uint32_t q_sa = smem_u32(sQ);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)  // rows 8..15 of the A tile alias rows 0..7: their results are never read
      ldmatrix_x4(q_sa + swz128(lane & 7, ks * 2 + (lane >> 4)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
  }
  const uint32_t k_sa = smem_u32(sK), v_sa = smem_u32(sV);
  float m_run = -INFINITY, l_run = 0.f;  // of head g (row g of the fragments)
  float o[2][4];
#pragma unroll
  for (int n = 0; n < 2; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll 1
  for (int t = 0; t < ntiles; ++t) {
    const int nk = min(nk_all - t * kAttnTile, kAttnTile);  // keys in this tile (the new token included)
    if (t > 0) {
      // further tiles of a long context: single-buffered (load, wait, compute)
      __syncthreads();
      const bool last = (t == ntiles - 1);
      const int cnt = nk - ((has_new && last) ? 1 : 0);
      load_tile(k0 + t * kAttnTile, cnt, (has_new && last) ? r_new : -1);
      if (has_new && last && cw == kCWarps - 1) {
        *reinterpret_cast<uint2*>(sK + swz128(r_new, lane >> 1) + (lane & 1) * 8) = knew_r;
        *reinterpret_cast<uint2*>(sV + swz128(r_new, lane >> 1) + (lane & 1) * 8) = vnew_r;
      }
      cp_async_wait<0>();
      __syncthreads();
    }
    // S = Q K^T: 8-key n-tiles in pairs
    float sc[kAttnTile / 8][4];
#pragma unroll
    for (int np = 0; np < kAttnTile / 16; ++np) {
#pragma unroll
      for (int e = 0; e < 4; ++e) sc[2 * np][e] = sc[2 * np + 1][e] = 0.f;
      if (np * 16 < nk) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          uint32_t b0, b1, b2, b3;
          const int key = np * 16 + (lane & 7) + ((lane >> 4) << 3);
          const int chunk = ks * 2 + ((lane >> 3) & 1);
          ldmatrix_x4(k_sa + swz128(key, chunk), b0, b1, b2, b3);
          mma_bf16_16816(sc[2 * np], qf[ks], b0, b1);
          mma_bf16_16816(sc[2 * np + 1], qf[ks], b2, b3);
        }
      }
    }
    // mask, row max of head g over the tile (this lane: keys nt*8 + 2*tq, +1), online softmax
    float mx = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < kAttnTile / 8; ++nt) {
      const int kl = nt * 8 + tq * 2;
      sc[nt][0] = (kl < nk) ? sc[nt][0] : -INFINITY;
      sc[nt][1] = (kl + 1 < nk) ? sc[nt][1] : -INFINITY;
      mx = fmaxf(mx, fmaxf(sc[nt][0], sc[nt][1]));
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_new = fmaxf(m_run, mx);  // finite: every tile holds at least one valid key
    const float corr = exp2f((m_run - m_new) * sl2);  // m_run = -inf -> 0
    m_run = m_new;
    const float msc = m_new * sl2;
    uint32_t pf[kAttnTile / 16][4];
    float lsum = 0.f;
#pragma unroll
    for (int kk = 0; kk < kAttnTile / 16; ++kk) {
      const float p0 = exp2f(sc[2 * kk][0] * sl2 - msc), p1 = exp2f(sc[2 * kk][1] * sl2 - msc);
      const float p2 = exp2f(sc[2 * kk + 1][0] * sl2 - msc), p3 = exp2f(sc[2 * kk + 1][1] * sl2 - msc);
      lsum += (p0 + p1) + (p2 + p3);
      pf[kk][0] = pack_bf16(p0, p1);
      pf[kk][1] = 0u;  // padding rows g + 8
      pf[kk][2] = pack_bf16(p2, p3);
      pf[kk][3] = 0u;
    }
    l_run = l_run * corr + lsum;
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      o[n][0] *= corr;
      o[n][1] *= corr;
    }
    // O[:, 16*cw .. 16*cw+16) += P V
#pragma unroll
    for (int kk = 0; kk < kAttnTile / 16; ++kk) {
      if (kk * 16 < nk) {
        uint32_t b0, b1, b2, b3;
        const int key = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        const int chunk = cw * 2 + (lane >> 4);
        ldmatrix_x4_trans(v_sa + swz128(key, chunk), b0, b1, b2, b3);
        mma_bf16_16816(o[0], pf[kk], b0, b1);
        mma_bf16_16816(o[1], pf[kk], b2, b3);
      }
    }
  }
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
#if OMC_MEGA_DETAIL
  if (c.prof_op && ctid == 0) c.prof_op[7] = global_ns();  // P V done
#endif
  // this lane: head g, dims 16*cw + 8*n + 2*tq, +1
  uint32_t* outp = op.out_ll + (long long)b * op.ldo + kvh * G * 128;  // this kv group's heads
  if (nsplit == 1) {
    if (g < G) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
#pragma unroll
      for (int n = 0; n < 2; ++n)
        __stcg(reinterpret_cast<uint2*>(outp + g * 128 + cw * 16 + n * 8 + tq * 2),
               make_uint2(ll4_word(o[n][0] * inv, otag16), ll4_word(o[n][1] * inv, otag16)));
    }
    return;
  }
  // ---- publish this CTA's partial {fp32, tag}: unnormalised O relative to the raw score maximum m, then (m, l)
  uint2* wsb = P.attn_part + (((long long)op.parity * P.grid + (long long)item * P.nsplit_max) * 8) * kPartStride;
  if (g < G) {
    uint2* dst = wsb + ((long long)split * 8 + g) * kPartStride;
#pragma unroll
    for (int n = 0; n < 2; ++n)
      __stcg(reinterpret_cast<uint4*>(dst + cw * 16 + n * 8 + tq * 2),
             make_uint4(__float_as_uint(o[n][0]), otag32, __float_as_uint(o[n][1]), otag32));
    if (cw == 0 && tq == 0)
      __stcg(reinterpret_cast<uint4*>(dst + 128), make_uint4(__float_as_uint(m_run), otag32, __float_as_uint(l_run), otag32));
  }
#if OMC_MEGA_DETAIL
  if (c.prof_op && ctid == 0) c.prof_op[5] = global_ns();  // own partial published
#endif
  float* part = reinterpret_cast<float*>(c.region_a);  // merge scratch [8 warps][4 heads][130] (aliases the Q/K/V tiles)
  // ---- merge: head hh belongs to the CTA of split (hh % nsplit). The (head, split) partials this CTA must read are
  // dealt round-robin to its 8 warps, each warp polls its share in ONE batch, warps combine through shared memory.
  int n_mine = 0;
  for (int hh = split; hh < G; hh += nsplit) ++n_mine;
  if (n_mine == 0) return;          // CTA-uniform
  __syncthreads();                  // everyone is done with the Q/K/V tiles: reuse them as [8 warps][4 heads][130]
  constexpr int MS = 5;             // (head, split) pairs per warp: 8 * 5 = 40 >= nsplit_max (37) or 4 heads x 7 splits
  float a4[MS][4], ml_m[MS], ml_l[MS];
  int pair_h[MS];
  {
    uint4 va[MS], vb[MS], vm[MS];
    Watchdog wd;
    bool ok;
    do {
      ok = true;
#pragma unroll
      for (int t = 0; t < MS; ++t) {
        const int pidx = cw + kCWarps * t;  // pair index = wi * nsplit + sp
        pair_h[t] = -1;
        if (pidx < n_mine * nsplit) {
          const int wi = pidx / nsplit, sp = pidx - wi * nsplit;
          pair_h[t] = wi;
          const uint2* src = wsb + ((long long)sp * 8 + (split + wi * nsplit)) * kPartStride;
          va[t] = ld_poll_v4(src + lane * 4);
          vb[t] = ld_poll_v4(src + lane * 4 + 2);
          vm[t] = ld_poll_v4(src + 128);
        }
      }
#pragma unroll
      for (int t = 0; t < MS; ++t)
        if (pair_h[t] >= 0 && !(va[t].y == otag32 && va[t].w == otag32 && vb[t].y == otag32 && vb[t].w == otag32 &&
                                vm[t].y == otag32 && vm[t].w == otag32))
          ok = false;
      if (!ok) wd.tick(P.err_flag, 9, split);
    } while (!ok);
#pragma unroll
    for (int t = 0; t < MS; ++t) {
      a4[t][0] = __uint_as_float(va[t].x); a4[t][1] = __uint_as_float(va[t].z);
      a4[t][2] = __uint_as_float(vb[t].x); a4[t][3] = __uint_as_float(vb[t].z);
      ml_m[t] = __uint_as_float(vm[t].x);
      ml_l[t] = __uint_as_float(vm[t].z);
    }
  }
  // per warp: combine its pairs head by head (at most 4 heads per CTA), then across warps through shared memory
#pragma unroll 1
  for (int wi = 0; wi < n_mine; ++wi) {
    float mw = -INFINITY;
#pragma unroll
    for (int t = 0; t < MS; ++t)
      if (pair_h[t] == wi) mw = fmaxf(mw, ml_m[t]);
    float lw = 0.f, ow[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < MS; ++t)
      if (pair_h[t] == wi) {
        const float sc = (ml_m[t] == -INFINITY) ? 0.f : exp2f((ml_m[t] - mw) * sl2);
        lw += ml_l[t] * sc;
#pragma unroll
        for (int e = 0; e < 4; ++e) ow[e] += a4[t][e] * sc;
      }
    float* dst = part + (cw * 4 + wi) * kPartStride;
    *reinterpret_cast<float2*>(dst + lane * 4) = make_float2(ow[0], ow[1]);
    *reinterpret_cast<float2*>(dst + lane * 4 + 2) = make_float2(ow[2], ow[3]);
    if (lane == 0) {
      dst[128] = mw;
      dst[129] = lw;
    }
  }
  __syncthreads();
  if (cw < n_mine) {  // warp wi finishes head split + wi * nsplit
    const int wi = cw, hh = split + wi * nsplit;
    float mt = -INFINITY;
    for (int w = 0; w < kCWarps; ++w) mt = fmaxf(mt, part[(w * 4 + wi) * kPartStride + 128]);
    float lt = 0.f, of[4] = {0.f, 0.f, 0.f, 0.f};
    for (int w = 0; w < kCWarps; ++w) {
      const float* pw = part + (w * 4 + wi) * kPartStride;
      const float sc = (pw[128] == -INFINITY) ? 0.f : exp2f((pw[128] - mt) * sl2);
      lt += pw[129] * sc;
#pragma unroll
      for (int e = 0; e < 4; ++e) of[e] += pw[lane * 4 + e] * sc;
    }
    const float inv = lt > 0.f ? 1.f / lt : 0.f;
    __stcg(reinterpret_cast<uint4*>(outp + hh * 128 + lane * 4),
           make_uint4(ll4_word(of[0] * inv, otag16), ll4_word(of[1] * inv, otag16), ll4_word(of[2] * inv, otag16),
                      ll4_word(of[3] * inv, otag16)));
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int NB, int G, bool TP>
__global__ void __launch_bounds__(kMegaThreads, 1) decode_mega_kernel(const MegaPlan* __restrict__ plan, uint32_t epoch) {
  extern __shared__ __align__(128) uint8_t mega_smem[];
  // layout: plan header | ctx | ring barriers | op window | slab window | meta (scratch 2 KB, residual slab, partials, block table) | region A | ring
  // (the header is copied too: a plan field read from global memory costs an L2 round trip under the weight stream)
  MegaPlan* s_plan = reinterpret_cast<MegaPlan*>(mega_smem);
  for (int i = threadIdx.x; i < (int)(sizeof(MegaPlan) / 8); i += kMegaThreads)
    reinterpret_cast<uint2*>(s_plan)[i] = __ldg(reinterpret_cast<const uint2*>(plan) + i);
  __syncthreads();
  const MegaPlan& P = *s_plan;
  const int n_ops = P.n_ops;
  MegaOp* s_ops = win_ops();
  SlabEnt* s_tab = win_tab();
  const MegaOp* gops = reinterpret_cast<const MegaOp*>(reinterpret_cast<const uint8_t*>(plan) + sizeof(MegaPlan));
  const int n_win = min(n_ops, kWin);
  uint8_t* meta = mega_smem + kHdrBytes;
  float* red = reinterpret_cast<float*>(meta);  // 16 floats
  int* s_ctx = reinterpret_cast<int*>(red + 16);                                    // 4 ints
  float* am_v = reinterpret_cast<float*>(s_ctx + 4);                                // [8 warps][16 lanes]
  int* am_i = reinterpret_cast<int*>(am_v + kCWarps * 16);
  float* s_bias = reinterpret_cast<float*>(am_i + kCWarps * 16);  // kBiasRows floats (meta scratch: 1104 + 512 <= 2048)
  bf16* s_h = reinterpret_cast<bf16*>(meta + 2048);
  float* s_xp = reinterpret_cast<float*>(meta + 2048 + 4 * kHRows * 2);
  int* s_bt = reinterpret_cast<int*>(meta + kMetaFixed);
  uint8_t* region_a = meta + kMetaFixed + bt_bytes_of(P.B, P.max_pages);
  uint8_t* ring = region_a + P.region_a_bytes;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {  // copy the op list, context lengths and block table into shared memory, init the ring barriers
    const uint2* src = reinterpret_cast<const uint2*>(gops);
    uint2* dst = reinterpret_cast<uint2*>(s_ops);
    const int n8 = n_win * (int)sizeof(MegaOp) / 8;
    for (int i = tid; i < n8; i += kMegaThreads) dst[i] = __ldg(src + i);
    for (int i = tid; i < P.B * P.max_pages; i += kMegaThreads) s_bt[i] = P.block_table[i];
    if (tid < P.B) s_ctx[tid] = P.ctx_lens[tid];
    if (tid == 0) {
      for (int s = 0; s < P.nslots; ++s) {
        mbar_init(ring_full(s), 1);
        *ring_gen(s) = 0u;
      }
      fence_barrier_init();
    }
  }
  __syncthreads();
  // this CTA's slab of the first kWin ops, then (one thread) the running stage index where each op starts
  for (int i = tid; i < n_win; i += kMegaThreads) s_tab[i] = make_slab(s_ops[i], 0u);
  __syncthreads();
  if (tid == 0) {
    uint32_t run = 0;
    for (int i = 0; i < n_win; ++i) {
      s_tab[i].base = run;
      run += s_tab[i].cnt;
    }
  }
  __syncthreads();

  // The context lives in shared memory, not in registers or on the stack: it is passed by reference into the phase
  // functions, and a local-memory copy means LDL round trips to L2 (the L1 left beside 227 KB of shared memory does not
  // hold it under the weight stream) on the refill path of every ring stage.
  MegaCtx& c = *reinterpret_cast<MegaCtx*>(mega_smem + 256);
  if (tid == 0) {
  c.P = s_plan; c.gops = gops; c.red = red; c.am_v = am_v; c.am_i = am_i; c.s_ctx = s_ctx;
  c.s_h = s_h; c.s_bias = s_bias; c.s_bt = s_bt; c.s_xp = s_xp; c.region_a = region_a; c.cta = blockIdx.x;
  c.grid = gridDim.x; c.n_ops = n_ops; c.epoch = epoch; c.pf_stages = P.pf_stages; c.scalar_gemv = P.scalar_gemv;
  c.prof_op = nullptr;
  }
  __syncthreads();

  RingHot hot;
  hot.ring_off = (int)(ring - mega_smem);
  hot.slot_bytes = P.slot_bytes; hot.nslots = P.nslots; hot.n_ops = n_ops;
  // every warp keeps its own cursor into the stage sequence for the refills it issues
  StageCursor refill;
  refill.op_i = 0;
  StageCursor refill_pf = refill;
  if (tid == 0) {  // prime the ring (stages 0 .. nslots-1) and the L2 prefetch window behind it
    StageCursor k = refill, kpf = refill;
    const bf16* src;
    uint32_t bytes, pitch;
    int ncopies;
    for (int s = P.nslots; s < P.nslots + P.pf_stages; ++s)
      if (locate_stage(hot, 0, kpf, (uint32_t)s, src, bytes, ncopies, pitch))
        for (int i = 0; i < ncopies; ++i)
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<const uint8_t*>(src) + (size_t)i * pitch),
                       "r"(bytes)
                       : "memory");
    StageCursor kpf2 = refill;
    for (int s = 0; s < P.nslots; ++s) issue_stage(hot, 0, P.pf_stages, k, kpf2, (uint32_t)s);
  }

  const int ctid = tid, cw = warp;
  uint32_t sc_base = 0;
  float best_v = -INFINITY;
  int best_i = 0x7fffffff;
  unsigned long long* prof = P.prof ? P.prof + ((size_t)c.cta * n_ops) * kProfStride : nullptr;
#pragma unroll 1
  for (int i = 0; i < n_ops; ++i) {
    const MegaOp& op = s_ops[i & (kWin - 1)];
    if (cw == kCWarps - 1 && i >= 1) {
      // slide the op window (warp 7, off the critical path: cp.async, no register staging): op i-1 is finished, its
      // entry takes op j1 = i-1+kWin; the slab entry of j1-1, copied one iteration ago, is built now. An entry is
      // first read (by a refill that runs ahead, or by the op loop) many block-wide barriers after it was written.
      const int j1 = i - 1 + kWin, j2 = j1 - 1;
      cp_async_wait<0>();
      __syncwarp();
      if (lane == 0 && j2 >= kWin && j2 < n_ops) {
        const SlabEnt& prev = s_tab[(j2 - 1) & (kWin - 1)];
        s_tab[j2 & (kWin - 1)] = make_slab(s_ops[j2 & (kWin - 1)], prev.base + prev.cnt);
      }
      if (j1 < n_ops && lane < (int)sizeof(MegaOp) / 8)
        cp_async8(reinterpret_cast<uint2*>(&s_ops[j1 & (kWin - 1)]) + lane, reinterpret_cast<const uint2*>(gops + j1) + lane);
      cp_async_commit();
    }
    if (prof && ctid == 0) {
      c.prof_op = prof + i * kProfStride;
      prof[i * kProfStride + 0] = global_ns();
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      prof[i * kProfStride + 3] = smid;
    }
    if (op.type == OP_GEMV) {
      if (!(op.flags & F_X_KEEP)) stage_x<NB>(c, op, i, ctid);  // F_X_KEEP: the previous op staged the same vector
      if (prof && ctid == 0) prof[i * kProfStride + 1] = global_ns();
      const int row0 = s_tab[i & (kWin - 1)].row0, rows = s_tab[i & (kWin - 1)].rows;
      gemv_consume<NB, TP>(c, hot, op, i, sc_base, row0, rows, cw, lane, refill, refill_pf, best_v, best_i);
      sc_base += (uint32_t)(((rows + op.R - 1) / op.R) * op.ksplit);
      if (op.flags & F_ARGMAX) {
        // CTA-level partial argmax per sequence: lane (i * NB + b) tracked sequence b = lane % NB
        if (lane < 16) {
          am_v[cw * 16 + lane] = best_v;
          am_i[cw * 16 + lane] = best_i;
        }
        __syncthreads();
        if (ctid < NB && ctid < P.B) {
          float bv = -INFINITY;
          int bi = 0x7fffffff;
          for (int w = 0; w < kCWarps; ++w)
            for (int ln = ctid; ln < kRMax * NB && ln < 16; ln += NB) {
              const float v = am_v[w * 16 + ln];
              const int ix = am_i[w * 16 + ln];
              if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
            }
          const uint32_t t = tag32_of(epoch, i);
          __stcg(reinterpret_cast<uint4*>(P.amax_part + ((long long)c.cta * 4 + ctid) * 2),
                 make_uint4(__float_as_uint(bv), t, (uint32_t)bi, t));
        }
        best_v = -INFINITY;
        best_i = 0x7fffffff;
      }
      __syncthreads();  // region A (activations) and the residual slab are reused by the next op
    } else if (op.type == OP_ATTN) {
      attn_phase<G>(c, op, i, ctid);
      __syncthreads();
    } else if (op.type == OP_FINAL) {
      // greedy sampling: lowest index among the maxima (HF argmax), token history, context lengths
      if (c.cta == 0 && cw == 0) {
        const uint32_t t = tag32_of(epoch, op.in_op);
        const int pos = P.hist_pos ? *P.hist_pos : 0;
        for (int b = 0; b < P.B; ++b) {
          float bv = -INFINITY;
          int bi = 0x7fffffff;
          for (int k = lane; k < c.grid; k += 32) {
            uint4 v;
            Watchdog wd;
            while (true) {
              v = ld_poll_v4(P.amax_part + ((long long)k * 4 + b) * 2);
              if (v.y == t && v.w == t) break;
              wd.tick(P.err_flag, 10, k);
            }
            const float val = __uint_as_float(v.x);
            const int ix = (int)v.z;
            if (val > bv || (val == bv && ix < bi)) { bv = val; bi = ix; }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
          }
          if (bi == 0x7fffffff) bi = 0;  // all-NaN logits: index 0 like torch.argmax, never the sentinel (embed row out of range)
          if (TP) {
            // vocab-parallel lm_head: every rank pushes its (max, global index) to all ranks (itself included), then
            // picks the best of the tp_size candidates — lowest index among equal maxima, like a single argmax would
            const uint32_t xt = tag32_of(epoch, i);
            const long long abase = 4LL * P.tp_size * P.B * P.C;
            if (lane < P.tp_size) {
              uint2* dst = P.xchg[lane] + abase + ((long long)P.tp_rank * P.B + b) * 2;
              st_sys_v2(dst, make_uint2(__float_as_uint(bv), xt));
              st_sys_v2(dst + 1, make_uint2((uint32_t)(bi + P.vocab_offset), xt));
            }
            bv = -INFINITY;
            bi = 0x7fffffff;
            if (lane < P.tp_size) {
              const uint2* src = P.xchg[P.tp_rank] + abase + ((long long)lane * P.B + b) * 2;
              uint2 v0, v1;
              Watchdog wd;
              while (true) {
                v0 = ld_sys_v2(src);
                v1 = ld_sys_v2(src + 1);
                if (v0.y == xt && v1.y == xt) break;
                wd.tick(P.err_flag, 12, lane);
              }
              bv = __uint_as_float(v0.x);
              bi = (int)v1.x;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
              if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (bi == 0x7fffffff) bi = P.vocab_offset;
            bi -= P.vocab_offset;  // the shared tail below adds it back
          }
          if (lane == 0) {
            const long long tok = (long long)bi + P.vocab_offset;
            P.tokens[b] = tok;
            if (P.token_hist && pos < P.hist_capacity) P.token_hist[(long long)pos * P.B + b] = tok;
            P.ctx_lens[b] = s_ctx[b] + 1;
          }
        }
        if (lane == 0 && P.hist_pos) *P.hist_pos = pos + 1;
      }
    }
    if (prof && ctid == 0) prof[i * kProfStride + 2] = global_ns();
  }
}

// ------------------------------------------------------------------------------------------------ host side
// Split a K-element row into the fewest chunks that fit a ring slot; chunks are kc0 elements (a multiple of 256 = one
// 16-byte vector per lane per 8 rounds when possible, else of 8), the last one takes what is left.
static int pick_ksplit(int K, int slot_bytes, int rows, int* kc0) {
  const int cap = slot_bytes / (2 * rows);  // elements of one row per slot
  if (K < 8 || K % 8 != 0 || cap < 8) return -1;
  const int d = (K + cap - 1) / cap;
  int per = (K + d - 1) / d;
  int c = (per + 255) & ~255;
  if (c > cap) c = (per + 7) & ~7;
  if (c > cap) return -1;
  *kc0 = c;
  return (K + c - 1) / c;
}

struct WsLayout {
  long long err, h1, qkv, attn, h2, act, part, amax, total;
};
static WsLayout ws_layout(int grid, int B, int C, int qw, int aw, int I) {
  WsLayout w;
  long long off = 0;
  auto take = [&](long long bytes) {
    const long long at = off;
    off = (off + bytes + 127) & ~127LL;
    return at;
  };
  w.err = take(128);
  w.h1 = take(2LL * B * C * 4);
  w.qkv = take(2LL * B * qw * 4);
  w.attn = take(2LL * B * aw * 4);
  w.h2 = take(2LL * B * C * 4);
  w.act = take(2LL * B * I * 4);
  w.part = take(2LL * grid * 8 * kPartStride * 8);
  w.amax = take((long long)grid * 4 * 2 * 8);
  w.total = off;
  return w;
}

}  // namespace omc

using namespace omc;

extern "C" long long omc_decode_plan_bytes(int n_layers) {
  if (n_layers < 0) return -1;
  return (long long)sizeof(MegaPlan) + (long long)(kOpsPerLayerMax * n_layers + 2) * (long long)sizeof(MegaOp);
}

extern "C" long long omc_decode_workspace_bytes(const omc_decode_desc* d) {
  if (d == nullptr || d->grid < 1 || d->batch < 1) return -1;
  return ws_layout(d->grid, d->batch, d->hidden, (d->q_heads + 2 * d->kv_heads) * 128, d->q_heads * 128, d->inter).total;
}

extern "C" long long omc_decode_xchg_bytes(const omc_decode_desc* d) {
  if (d == nullptr || d->batch < 1 || d->hidden < 1 || d->tp_size < 1 || d->tp_size > kMaxTp) return -1;
  // partial sums [4 slots][tp][B][C] + argmax candidates [tp][B][2], 8-byte {value, tag} words
  return (4LL * d->tp_size * d->batch * d->hidden + 2LL * d->tp_size * d->batch) * 8;
}

extern "C" int omc_decode_plan_build(const omc_decode_desc* d, void* plan_host) {
  if (d == nullptr || plan_host == nullptr) return set_error(OMC_ERR_ARG, "omc_decode_plan_build: null argument");
  if (d->batch < 1 || d->batch > 4) return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: batch must be 1..4");
  if (d->n_layers < 0 || kOpsPerLayerMax * d->n_layers + 2 > kMaxOps) return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: too many layers");
  if (d->grid < 1) return set_error(OMC_ERR_ARG, "omc_decode_plan_build: grid must be the number of CTAs (SMs)");
  {
    const int g = d->kv_heads >= 1 && d->q_heads % d->kv_heads == 0 ? d->q_heads / d->kv_heads : 0;
    if (!(g == 1 || g == 2 || g == 4 || g == 6 || g == 7 || g == 8))
      return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: q heads per kv head must be one of 1, 2, 4, 6, 7, 8");
  }
  if (d->batch * d->kv_heads > d->grid) return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: batch * kv_heads exceeds the grid");
  if (d->hidden % 8 != 0 || d->hidden > 4096) return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: hidden must be a multiple of 8, <= 4096");
  if ((d->hidden + d->grid - 1) / d->grid + 1 > kHRows)
    return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: hidden / grid exceeds the residual rows one CTA can own");
  if (d->inter % 8 != 0 || d->vocab < 1) return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: inter % 8 != 0 or empty vocab");
  if (d->page_size < 1 || d->max_pages < 1 || d->batch * d->max_pages > kMaxBt)
    return set_error(OMC_ERR_ARG, "omc_decode_plan_build: bad paging parameters (batch * max_pages must be <= 1024)");
  if (d->workspace == nullptr) return set_error(OMC_ERR_ARG, "omc_decode_plan_build: workspace is null");
  const int tp = d->tp_size > 1 ? d->tp_size : 1;
  if (tp > kMaxTp || (tp > 1 && (d->tp_rank < 0 || d->tp_rank >= tp)))
    return set_error(OMC_ERR_ARG, "omc_decode_plan_build: tp_size must be <= 8 and 0 <= tp_rank < tp_size");
  for (int p = 0; p < tp && tp > 1; ++p)
    if (d->xchg[p] == nullptr)
      return set_error(OMC_ERR_ARG, "omc_decode_plan_build: xchg[p] (peer exchange buffers, omc_decode_xchg_bytes each) is null");
  if (d->rope_cs == nullptr || d->rope_positions < d->max_pages * d->page_size)
    return set_error(OMC_ERR_ARG, "omc_decode_plan_build: rope_cs must cover max_pages * page_size positions");
  const int C = d->hidden, Hq = d->q_heads, Hkv = d->kv_heads, I = d->inter, B = d->batch;
  const int qw = (Hq + 2 * Hkv) * 128, aw = Hq * 128;
  const WsLayout w = ws_layout(d->grid, B, C, qw, aw, I);
  uint8_t* ws = static_cast<uint8_t*>(d->workspace);
  auto ll = [&](long long base, int parity, int width) {
    return reinterpret_cast<uint32_t*>(ws + base) + (long long)parity * B * width;
  };
  int slot_bytes = d->ring_slot_bytes > 0 ? d->ring_slot_bytes : kSlotBytesDefault;
  if (slot_bytes % 128 != 0 || slot_bytes < 1024 || slot_bytes > kSlotBytesMax)
    return set_error(OMC_ERR_ARG, "omc_decode_plan_build: ring_slot_bytes must be a multiple of 128 in [1024, 32768]");
  MegaPlan* P = static_cast<MegaPlan*>(plan_host);
  memset(P, 0, sizeof(MegaPlan));
  MegaOp* ops = reinterpret_cast<MegaOp*>(P + 1);
  int n = 0, kmax = 0, knorm = 0;
  bool bad = false;
  auto gemv = [&](const void* W, int N, int K, const uint32_t* x_ll, int ldx, int in_op, const void* norm_w, const void* bias,
                  uint32_t* out_ll, int ldo, int epi, int flags, int ldw = 0) -> int {
    MegaOp& o = ops[n];
    memset(&o, 0, sizeof(o));
    o.type = OP_GEMV; o.N = N; o.K = K; o.epi = (uint8_t)epi; o.flags = (uint8_t)flags;
    o.ldw = ldw > 0 ? ldw : K;
    o.gran = (epi == EPI_SWIGLU) ? 2 : 1;
    o.kc0 = K;
    // stage geometry: R whole rows when two of them fit a slot, else K chunks of one row
    int R = slot_bytes / (K * 2);
    if (R > kRMax) R = kRMax;
    if (o.gran == 2) R &= ~1;
    int ksp = 1;
    if (R < 2) {
      // split-K stages: one row per stage by default. (Two-row stages - the same K chunk of two rows, one bulk copy per row,
      // activation fragments shared by both rows - measured 0.5 % slower: the ring is latency-bound, not LDS-bound.)
      R = 1;
      ksp = pick_ksplit(K, slot_bytes, 1, &o.kc0);
      if (d->tune & 2) {
        R = 2;
        ksp = pick_ksplit(K, slot_bytes, 2, &o.kc0);
      }
    }
    o.ksplit = (uint8_t)ksp;
    if (ksp < 1 || ksp > 255 || K % 8 != 0 || N % o.gran != 0 || R < o.gran) { bad = true; o.ksplit = 1; o.kc0 = K; }
    o.R = (uint8_t)(R < 1 ? 1 : R);
    if (norm_w != nullptr && K > 4096) bad = true;
    o.W = (const bf16*)W; o.norm_w = (const bf16*)norm_w; o.bias = (const bf16*)bias;
    o.x_ll = x_ll; o.out_ll = out_ll; o.ldx = ldx; o.ldo = ldo; o.in_op = (int16_t)in_op;
    if (K > kmax) kmax = K;
    if (norm_w != nullptr && K > knorm) knorm = K;
    return n++;
  };
  int prev = -1;  // op that produced the current residual-stream broadcast (h1)
  for (int li = 0; li < d->n_layers; ++li) {
    const int par = li & 1;
    const int i_qkv = gemv(d->qkv_w[li], qw, C, li == 0 ? nullptr : ll(w.h1, par, C), C, prev, d->ln1[li], d->qkv_b[li],
                           ll(w.qkv, par, qw), qw, EPI_NONE, li == 0 ? F_X_EMBED : 0);
    MegaOp& a = ops[n];
    memset(&a, 0, sizeof(a));
    a.type = OP_ATTN; a.x_ll = ll(w.qkv, par, qw); a.ldx = qw; a.in_op = (int16_t)i_qkv; a.out_ll = ll(w.attn, par, aw); a.ldo = aw;
    a.parity = (uint8_t)par;
    a.pool = static_cast<bf16*>(d->kv_pool) + (long long)li * d->kv_layer_stride;
    const int i_attn = n++;
    const int i_o = gemv(d->o_w[li], C, aw, ll(w.attn, par, aw), aw, i_attn, nullptr, nullptr, ll(w.h2, par, C), C, EPI_RES, 0);
    ops[i_o].xslot = (uint8_t)(1 + par);  // row-parallel under TP: partial sums cross the GPUs through exchange slot par
    // MLP. For batches of 3 and 4 sequences down_proj is cut into nsub K-chunk sub-ops (a down_proj row of I elements is
    // nsub ring slots long): down_0 .. down_{nsub-1} are [C, kc] slices of down_w (row pitch I) whose row sums accumulate
    // in shared memory; the last one adds the residual and broadcasts the new hidden state. The activation vectors staged
    // per op shrink from B * I to B * kc elements, which is what leaves room for a ring at those batch sizes (B = 4: 4 -> 10
    // slots, 6.36 -> 5.27 ms per step; B = 3: 5.02 -> 4.29). gate_up can be cut the same way (gate_up_j produces
    // act[j*kc, (j+1)*kc), gate_up_0 stages the normed input and the others keep it) so that down_j only waits for
    // gate_up_j - measured 3 % slower than cutting down_proj alone, and at B = 1 either cut costs 9 % (2.71 -> 2.95 ms):
    // every extra op costs ~1.8 us of staging and barriers plus its own ramp and tail, so fewer, longer ops win whenever
    // the ring is deep enough. B = 2: 3.20 ms uncut (9 slots), 3.53 with down_proj cut (13 slots).
    int kc = I;
    int nsub = ((long long)I * 2 > slot_bytes) ? pick_ksplit(I, slot_bytes, 1, &kc) : 1;
    // cut: 0 = one gate_up and one down op, 1 = both cut into sub-ops, 2 = only down_proj cut (gate_up stays whole; the
    // staged activation vector still shrinks to kc elements). tune bits 2 / 3 / 8 force 0 / 1 / 2.
    int cut = B >= 3 ? 2 : 0;
    if (d->tune & 4) cut = 0;
    else if (d->tune & 8) cut = 1;
    else if (d->tune & 256) cut = 2;
    if (nsub > kMaxSub || kc % 8 != 0 || cut == 0) nsub = 1;
    if (nsub <= 1) {
      const int i_gu = gemv(d->gate_up_w[li], 2 * I, C, ll(w.h2, par, C), C, i_o, d->ln2[li], nullptr, ll(w.act, par, I), I,
                            EPI_SWIGLU, 0);
      prev = gemv(d->down_w[li], C, I, ll(w.act, par, I), I, i_gu, nullptr, nullptr, ll(w.h1, (li + 1) & 1, C), C, EPI_RES, 0);
    } else {
      int i_gu[kMaxSub];
      for (int j = 0; j < nsub; ++j) {
        if (cut == 2 && j > 0) { i_gu[j] = i_gu[0]; continue; }
        const int a0 = j * kc, len = cut == 2 ? I : (I - a0 < kc ? I - a0 : kc);
        i_gu[j] = gemv(static_cast<const bf16*>(d->gate_up_w[li]) + (size_t)2 * a0 * C, 2 * len, C, ll(w.h2, par, C), C, i_o,
                       d->ln2[li], nullptr, ll(w.act, par, I) + a0, I, EPI_SWIGLU, j > 0 ? F_X_KEEP : 0);
      }
      for (int j = 0; j < nsub; ++j) {
        const int a0 = j * kc, len = I - a0 < kc ? I - a0 : kc;
        const bool last = (j == nsub - 1);
        prev = gemv(static_cast<const bf16*>(d->down_w[li]) + a0, C, len, ll(w.act, par, I) + a0, I, i_gu[j], nullptr, nullptr,
                    last ? ll(w.h1, (li + 1) & 1, C) : nullptr, C, last ? EPI_RES : (j == 0 ? EPI_PART_SET : EPI_PART_ADD),
                    last ? F_ADD_PART : 0, I);
      }
    }
    ops[prev].xslot = (uint8_t)(3 + par);
  }
  const int i_head = gemv(d->lm_head, d->vocab, C, d->n_layers == 0 ? nullptr : ll(w.h1, d->n_layers & 1, C), C, prev,
                          d->final_norm, nullptr, nullptr, d->vocab, EPI_NONE,
                          F_OUT_F32 | F_ARGMAX | (d->n_layers == 0 ? F_X_EMBED : 0));
  ops[i_head].out_f32 = d->logits;
  MegaOp& f = ops[n++];
  memset(&f, 0, sizeof(f));
  f.type = OP_FINAL;
  f.in_op = (int16_t)i_head;
  if (bad) return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: a layer shape does not fit the ring (K % 8, SwiGLU pair > slot, norm K > 4096)");
  P->n_ops = n; P->B = B; P->C = C; P->Hq = Hq; P->Hkv = Hkv; P->G = Hq / Hkv;
  P->page_size = d->page_size; P->max_pages = d->max_pages; P->grid = d->grid;
  P->nsplit_max = d->grid / (B * Hkv);
  P->vocab_offset = d->vocab_offset; P->hist_capacity = d->hist_capacity; P->kmax = kmax;
  int region_a = B * kmax * 2;
  if (region_a < (B + 1) * knorm * 2) region_a = (B + 1) * knorm * 2;  // normed ops stage the norm weights behind x
  if (region_a < kAttnScratchBytes) region_a = kAttnScratchBytes;
  region_a = (region_a + 127) & ~127;
  const int ops_bytes = kHdrBytes;  // plan header, context, ring barriers, op window + slab window
  // the refill path may look nslots stages ahead: that must stay inside the op window, so no GEMV op may be empty for
  // a CTA (every op then contributes >= 1 stage and 16 stages span at most 16 GEMV ops + their attention ops)
  for (int i = 0; i < n; ++i)
    if (ops[i].type == OP_GEMV && ops[i].N / ops[i].gran < d->grid)
      return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: a weight matrix has fewer row groups than the grid has CTAs");
  const int meta_bytes = kMetaFixed + bt_bytes_of(B, d->max_pages);
  int nslots = (kSmemLimit - ops_bytes - meta_bytes - region_a) / slot_bytes;
  if (nslots > kMaxSlots) nslots = kMaxSlots;
  {
    // a split-K op streams its rows as ksplit consecutive stages consumed by one warp: the ring works in whole rows.
    // Measured on down_proj (3 chunks per row): 12 slots 21.3 us, 13 slots 26.8 us, 10 slots 31.5 us per layer.
    int ksp_max = 1;
    for (int i = 0; i < n; ++i)
      if (ops[i].type == OP_GEMV && ops[i].ksplit > ksp_max) ksp_max = ops[i].ksplit;
    if (ksp_max > 1 && nslots >= 2 * ksp_max) nslots -= nslots % ksp_max;
  }
  if (((d->tune >> 4) & 15) >= 2 && nslots > ((d->tune >> 4) & 15)) nslots = (d->tune >> 4) & 15;  // A/B: cap on the ring depth
  if (nslots < 2) return set_error(OMC_ERR_SHAPE, "omc_decode_plan_build: activations leave no room for the weight ring");
  P->nslots = nslots; P->region_a_bytes = region_a;
  P->pf_stages = d->l2_prefetch_stages < 0 ? 0 : d->l2_prefetch_stages;
  P->slot_bytes = slot_bytes;
  P->scalar_gemv = d->tune & 1;
  P->smem_bytes = ops_bytes + meta_bytes + region_a + nslots * slot_bytes;
  P->eps = d->eps; P->scale_log2 = d->attn_scale * 1.4426950408889634f;
  P->embed = (const bf16*)d->embed; P->rope_cs = d->rope_cs; P->block_table = d->block_table; P->ctx_lens = d->ctx_lens;
  P->tokens = d->tokens; P->token_hist = d->token_hist; P->hist_pos = d->hist_pos;
  P->err_flag = d->status ? d->status : reinterpret_cast<int32_t*>(ws + w.err);
  P->prof = static_cast<unsigned long long*>(d->prof);
  P->tp_rank = tp > 1 ? d->tp_rank : 0;
  P->tp_size = tp;
  for (int p = 0; p < tp && tp > 1; ++p) P->xchg[p] = static_cast<uint2*>(d->xchg[p]);
  P->attn_part = reinterpret_cast<uint2*>(ws + w.part);
  P->amax_part = reinterpret_cast<uint2*>(ws + w.amax);
  return OMC_OK;
}

template <int NB, int G, bool TP>
static int launch_mega(const MegaPlan* host, const void* plan_dev, uint32_t epoch, cudaStream_t st) {
  static int attr_smem_dev[kMaxDevices] = {};
  int& attr_smem = attr_smem_dev[cur_device()];
  if (host->smem_bytes > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(decode_mega_kernel<NB, G, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, host->smem_bytes);
    if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
    attr_smem = host->smem_bytes;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(host->grid);
  cfg.blockDim = dim3(kMegaThreads);
  cfg.dynamicSmemBytes = host->smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: they wait for each other's data
  attrs[0].val.cooperative = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  const MegaPlan* arg = static_cast<const MegaPlan*>(plan_dev);
  cudaError_t e = cudaLaunchKernelEx(&cfg, decode_mega_kernel<NB, G, TP>, arg, epoch);
  if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
  return OMC_OK;
}

extern "C" int omc_decode_step(const void* plan_host, const void* plan_dev, unsigned int epoch, void* stream) {
  if (plan_host == nullptr || plan_dev == nullptr) return set_error(OMC_ERR_ARG, "omc_decode_step: null plan");
  const MegaPlan* P = static_cast<const MegaPlan*>(plan_host);
  if (P->n_ops < 2 || P->n_ops > kMaxOps || P->grid < 1) return set_error(OMC_ERR_ARG, "omc_decode_step: plan not built");
  cudaStream_t st = (cudaStream_t)stream;
#define OMC_MEGA_CASE(NB_, G_, TP_) \
  if (P->B == NB_ && P->G == G_) return launch_mega<NB_, G_, TP_>(P, plan_dev, epoch, st);
#define OMC_MEGA_G(G_, TP_) OMC_MEGA_CASE(1, G_, TP_) OMC_MEGA_CASE(2, G_, TP_) OMC_MEGA_CASE(3, G_, TP_) OMC_MEGA_CASE(4, G_, TP_)
  if (P->tp_size > 1) {
    // tensor-parallel instantiations (in-kernel NVLink all-reduce): Qwen2-7B at TP 2/4 (G = 7) and TP 8 (G = 4), G = 2 tests
    OMC_MEGA_G(2, true) OMC_MEGA_G(4, true) OMC_MEGA_G(7, true)
    return set_error(OMC_ERR_SHAPE, "omc_decode_step: tensor parallelism needs batch 1..4 and 2, 4 or 7 q heads per kv head");
  }
  OMC_MEGA_G(1, false) OMC_MEGA_G(2, false) OMC_MEGA_G(4, false) OMC_MEGA_G(6, false) OMC_MEGA_G(7, false) OMC_MEGA_G(8, false)
#undef OMC_MEGA_G
#undef OMC_MEGA_CASE
  return set_error(OMC_ERR_SHAPE, "omc_decode_step: batch must be 1..4 and q heads per kv head one of 1, 2, 4, 6, 7, 8");
}

