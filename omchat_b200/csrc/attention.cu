// Flash-style attention kernels, head_dim 128, bf16 in / fp32 softmax + accumulate.
//   attention_fwd_kernel      : packed var-len sequences, non-causal (InternViT, 1025 tokens) or causal GQA (Qwen2
//                               prefill). 64 query rows per CTA (4 warps x 16), 64-key tiles double-buffered with
//                               cp.async, ldmatrix + mma.sync.m16n8k16 with XOR-swizzled shared memory.
//   paged_decode_attn_kernel  : one decode step over the paged KV cache. The 7 query heads that share a KV head are
//                               packed into the M dimension of the MMA so K/V are streamed from HBM exactly once;
//                               keys are split over warps and CTAs and merged with a log-sum-exp combine by the last
//                               CTA to finish. RoPE of the new q/k and the cache append are fused in.
// Reference call sites: intern_vit_6b/modeling_intern_vit.py:148-152, flash_attention.py:43-55;
// transformers models/qwen2/modeling_qwen2.py:124-146,161-184,227-243.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <math.h>
#include <stdint.h>

#include "omc_internal.h"
#include "ptx.cuh"

namespace omc {

typedef __nv_bfloat16 bf16;
constexpr int kHD = 128;            // head dim
constexpr int kRowBytes = kHD * 2;  // 256 B per (token, head) row
constexpr float kLog2e = 1.4426950408889634f;

// byte offset of 16-byte chunk `chunk` (0..15) of row `row` in a [rows][128] bf16 tile, XOR-swizzled
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)(row * kRowBytes + ((chunk ^ (row & 7)) << 4)); }

// S(16 x 8*NT) += Q(16x128) K^T for NT key n-tiles starting at key row `key0` of the K tile in smem.
template <int NT>
__device__ __forceinline__ void qk_tile(const uint32_t (&qf)[8][4], uint32_t k_smem, int key0, float (&s)[NT][4]) {
  const int lane = lane_id();
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
    for (int np = 0; np < NT / 2; ++np) {
      uint32_t b0, b1, b2, b3;
      const int key = key0 + np * 16 + (lane & 7) + ((lane >> 4) << 3);
      const int chunk = ks * 2 + ((lane >> 3) & 1);
      ldmatrix_x4(k_smem + swz(key, chunk), b0, b1, b2, b3);
      mma_bf16_16816(s[2 * np], qf[ks], b0, b1);
      mma_bf16_16816(s[2 * np + 1], qf[ks], b2, b3);
    }
  }
}

// O(16x128) += P(16 x 16*KS) V for KS key k-steps starting at key row `key0` of the V tile in smem.
template <int KS>
__device__ __forceinline__ void pv_tile(const uint32_t (&pf)[KS][4], uint32_t v_smem, int key0, float (&o)[16][4]) {
  const int lane = lane_id();
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int dp = 0; dp < 8; ++dp) {
      uint32_t b0, b1, b2, b3;
      const int key = key0 + ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
      const int chunk = dp * 2 + (lane >> 4);
      ldmatrix_x4_trans(v_smem + swz(key, chunk), b0, b1, b2, b3);
      mma_bf16_16816(o[2 * dp], pf[ks], b0, b1);
      mma_bf16_16816(o[2 * dp + 1], pf[ks], b2, b3);
    }
  }
}

// ==================================================================================================== prefill / ViT
constexpr int kAttM = 64, kAttN = 64, kAttThreads = 128;
constexpr int kAttSmem = kAttM * kRowBytes + 4 * kAttN * kRowBytes;  // Q + 2x(K,V) = 80 KB

struct AttnParams {
  const bf16 *q, *k, *v;
  bf16* out;
  long long ldq, ldk, ldv, ldo;
  const int32_t* cu;
  int Hq, Hkv, causal;
  float scale_log2;
  int tail_mode;  // 1: only the ragged tail (rows >= 128 * (len / 128)) — the full tiles run on the tcgen05 kernel
};

// cooperative cp.async of `rows` rows (256 B each) from global (row stride ld elements) into a swizzled tile;
// rows >= valid_rows are zero-filled.
__device__ __forceinline__ void load_tile_async(uint8_t* smem_tile, const bf16* g, long long ld, int rows, int valid_rows,
                                                int tid, int nthreads) {
  for (int c = tid; c < rows * 16; c += nthreads) {
    const int r = c >> 4, ch = c & 15;
    const bool ok = r < valid_rows;
    const bf16* src = g + (long long)(ok ? r : 0) * ld + ch * 8;
    cp_async16(smem_tile + swz(r, ch), src, ok);
  }
}

__global__ void __launch_bounds__(kAttThreads) attention_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  uint8_t* sQ = att_smem;
  uint8_t* sK[2] = {att_smem + kAttM * kRowBytes, att_smem + kAttM * kRowBytes + 2 * kAttN * kRowBytes};
  uint8_t* sV[2] = {sK[0] + kAttN * kRowBytes, sK[1] + kAttN * kRowBytes};

  const int seq = blockIdx.z, head = blockIdx.y;
  const int row0 = p.cu[seq], len = p.cu[seq + 1] - row0;
  // heavy (late) causal tiles first
  const int mt = p.tail_mode ? (len / 128) * (128 / kAttM) + (int)blockIdx.x
                             : (p.causal ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x);
  const int m0 = mt * kAttM;
  if (m0 >= len) return;
  const int kvh = head / (p.Hq / p.Hkv);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;

  const bf16* qg = p.q + (long long)(row0 + m0) * p.ldq + head * kHD;
  const bf16* kg = p.k + (long long)row0 * p.ldk + kvh * kHD;
  const bf16* vg = p.v + (long long)row0 * p.ldv + kvh * kHD;

  int kv_end = len;
  if (p.causal) kv_end = min(len, m0 + kAttM);
  const int n_tiles = (kv_end + kAttN - 1) / kAttN;

  load_tile_async(sQ, qg, p.ldq, kAttM, len - m0, tid, kAttThreads);
  load_tile_async(sK[0], kg, p.ldk, kAttN, kv_end, tid, kAttThreads);
  load_tile_async(sV[0], vg, p.ldv, kAttN, kv_end, tid, kAttThreads);
  cp_async_commit();

  uint32_t qf[8][4];
  float o[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int j = 0; j < n_tiles; ++j) {
    const int buf = j & 1;
    __syncthreads();  // everyone is done reading buffer buf^1 (iteration j-1)
    if (j + 1 < n_tiles) {
      const int k0 = (j + 1) * kAttN;
      load_tile_async(sK[buf ^ 1], kg + (long long)k0 * p.ldk, p.ldk, kAttN, kv_end - k0, tid, kAttThreads);
      load_tile_async(sV[buf ^ 1], vg + (long long)k0 * p.ldv, p.ldv, kAttN, kv_end - k0, tid, kAttThreads);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int r = warp * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
        ldmatrix_x4(smem_u32(sQ) + swz(r, ks * 2 + (lane >> 4)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    qk_tile<8>(qf, smem_u32(sK[buf]), 0, s);

    const int key_base = j * kAttN;
    const int qrow0 = m0 + warp * 16 + g;  // rows qrow0 and qrow0 + 8
    const bool need_mask = (key_base + kAttN > kv_end) || (p.causal && key_base + kAttN > m0 + warp * 16);
    if (need_mask) {
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = key_base + nt * 8 + tq * 2 + (e & 1);
          const int qr = qrow0 + ((e >> 1) << 3);
          if (key >= kv_end || (p.causal && key > qr)) s[nt][e] = -INFINITY;
        }
      }
    }
    // online softmax (rows g and g+8 of this warp's 16-row slab)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float corr[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      msc[r] = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;
      corr[r] = (m_run[r] == -INFINITY) ? 0.f : exp2f(m_run[r] * p.scale_log2 - msc[r]);
      m_run[r] = m_new;
    }
    uint32_t pf[4][4];
    float ls[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = exp2f(s[nt][0] * p.scale_log2 - msc[0]);
      const float p1 = exp2f(s[nt][1] * p.scale_log2 - msc[0]);
      const float p2 = exp2f(s[nt][2] * p.scale_log2 - msc[1]);
      const float p3 = exp2f(s[nt][3] * p.scale_log2 - msc[1]);
      ls[0] += p0 + p1;
      ls[1] += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * corr[r] + ls[r];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      o[i][0] *= corr[0];
      o[i][1] *= corr[0];
      o[i][2] *= corr[1];
      o[i][3] *= corr[1];
    }
    pv_tile<4>(pf, smem_u32(sV[buf]), 0, o);
  }

  // finalize: O / l, stage through sQ (each warp only touches its own 16 rows), 16-byte coalesced stores
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  __syncwarp();
#pragma unroll
  for (int dt = 0; dt < 16; ++dt) {
    const int r0 = warp * 16 + g, r1 = r0 + 8;
    const int chunk = dt, within = tq * 4;  // n-tile dt covers d = dt*8 .. dt*8+7 = one 16-byte chunk
    *reinterpret_cast<uint32_t*>(sQ + swz(r0, chunk) + within) = pack_bf16(o[dt][0] * inv0, o[dt][1] * inv0);
    *reinterpret_cast<uint32_t*>(sQ + swz(r1, chunk) + within) = pack_bf16(o[dt][2] * inv1, o[dt][3] * inv1);
  }
  __syncwarp();
  bf16* og = p.out + (long long)(row0 + m0) * p.ldo + head * kHD;
  for (int c = lane; c < 16 * 16; c += 32) {
    const int r = warp * 16 + (c >> 4), ch = c & 15;
    if (m0 + r < len) *reinterpret_cast<uint4*>(og + (long long)r * p.ldo + ch * 8) = *reinterpret_cast<const uint4*>(sQ + swz(r, ch));
  }
}

// ==================================================================================================== paged decode
constexpr int kDecThreads = 128;
constexpr int kDecTile = 16;  // keys per warp tile (page_size must be a multiple)
constexpr int kDecStages = 3;  // cp.async stages per warp: a warp is latency-bound on its own tile stream (2 stages: one
                               // 8 KB tile in flight per warp, batch-32 attention at 2.4 TB/s)
constexpr int kDecWarpBuf = kDecStages * 2 * kDecTile * kRowBytes;  // stages x (K,V) x 16 rows = 24 KB per warp
constexpr int kDecQBytes = 16 * kRowBytes;
constexpr int kDecPartStride = kHD + 2;  // O[128], m, l
constexpr int kDecMaxStagedPages = 1024;  // page ids of one sequence staged in shared memory (longer tables: global look-ups)
constexpr int kDecSmem = kDecQBytes + 4 * kDecWarpBuf;  // 100 KB: two CTAs per SM (the warp-merge scratch aliases warp 0's tiles)
static_assert(4 * 8 * kDecPartStride * 4 <= kDecWarpBuf, "merge scratch must fit the tile buffer of warp 0");

struct DecParams {
  const bf16* qkv;  // [B, ldq]: q heads | k heads | v heads of the NEW token (pre-RoPE) when fused != 0, else q only
  long long ldq;
  bf16* pool;  // [pages, 2, Hkv, page_size, 128]
  const int32_t* block_table;
  int max_pages, page_size;
  const int32_t* ctx_lens;
  const float* inv_freq;  // non-null => fused RoPE + append of the new token
  int Hq, Hkv, G, splits;
  float scale_log2;
  bf16* out;
  long long ldo;
  float* ws;              // [B, Hkv, splits, G, 130]
  unsigned int* counters; // [B * Hkv], zero-initialised, self-resetting
};

__global__ void __launch_bounds__(kDecThreads) paged_decode_attn_kernel(const DecParams p) {
  extern __shared__ __align__(128) uint8_t dec_smem[];
  uint8_t* sQ = dec_smem;
  uint8_t* sKV = dec_smem + kDecQBytes;
  float* sPart = reinterpret_cast<float*>(dec_smem + kDecQBytes);  // [4 warps][8 rows][130], aliases warp 0's K/V tiles
  __shared__ int s_is_last;

  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int G = p.G;
  // Programmatic dependent launch (csrc/gemm_stream.cu): let the next kernel's CTAs move in as ours retire. Everything that
  // does not depend on THIS step's qkv rows runs before griddepcontrol.wait, i.e. while the qkv GEMM is still streaming: the
  // context length (incremented at the top of the step, several completed kernels ago), the sequence's page ids (staged in
  // shared memory: no dependent table look-up in front of every tile fetch) and the first K/V tiles of every warp's pipeline
  // (the cache rows of earlier tokens are immutable; the new token's slot is patched in shared memory below).
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int ctx = p.ctx_lens[b];
  const int pos_new = ctx - 1;
  const bool fused = p.inv_freq != nullptr;
  const bf16* qrow = p.qkv + (long long)b * p.ldq;
  __shared__ int s_pages[kDecMaxStagedPages];
  {
    const int n_pages = (ctx + p.page_size - 1) / p.page_size;
    for (int i = tid; i < n_pages && i < kDecMaxStagedPages; i += kDecThreads) s_pages[i] = p.block_table[(long long)b * p.max_pages + i];
  }
  // ---- Q tile [16][128]: rows 0..G-1 = query heads of this KV group (rotated), rest zero
  for (int c = tid; c < 16 * 16; c += kDecThreads) *reinterpret_cast<uint4*>(sQ + swz(c >> 4, c & 15)) = make_uint4(0, 0, 0, 0);
  __syncthreads();

  // ---- this CTA's key-tile range, tiles dealt round-robin to the 4 warps
  const int tiles_ctx = (ctx + kDecTile - 1) / kDecTile;
  const int per_split = (tiles_ctx + p.splits - 1) / p.splits;
  const int t_begin = split * per_split;
  const int t_end = min(tiles_ctx, t_begin + per_split);

  uint8_t* wbuf = sKV + warp * kDecWarpBuf;
  auto stage_k = [&](int st) { return wbuf + st * (2 * kDecTile * kRowBytes); };
  auto stage_v = [&](int st) { return wbuf + st * (2 * kDecTile * kRowBytes) + kDecTile * kRowBytes; };
  auto issue = [&](int tile, int st) {
    const int key0 = tile * kDecTile;
    const int pi = key0 / p.page_size;
    const int page = pi < kDecMaxStagedPages ? s_pages[pi] : p.block_table[(long long)b * p.max_pages + pi];
    const int slot = key0 % p.page_size;
    const bf16* kg = p.pool + ((((long long)page * 2 + 0) * p.Hkv + kvh) * p.page_size + slot) * kHD;
    const bf16* vg = p.pool + ((((long long)page * 2 + 1) * p.Hkv + kvh) * p.page_size + slot) * kHD;
    for (int c = lane; c < kDecTile * 16; c += 32) {
      const int r = c >> 4, ch = c & 15;
      cp_async16(stage_k(st) + swz(r, ch), kg + r * kHD + ch * 8, true);
      cp_async16(stage_v(st) + swz(r, ch), vg + r * kHD + ch * 8, true);
    }
  };
  int tile = t_begin + warp, it = 0;
#pragma unroll
  for (int s = 0; s < kDecStages - 1; ++s) {  // one commit group per stage, empty groups keep the count uniform
    if (tile + 4 * s < t_end) issue(tile + 4 * s, s);
    cp_async_commit();
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");  // from here on: this step's q / k / v rows
  for (int e = tid; e < G * 64; e += kDecThreads) {
    const int r = e >> 6, i = e & 63;
    const bf16* hp = qrow + (kvh * G + r) * kHD;
    float x0 = __bfloat162float(hp[i]), x1 = __bfloat162float(hp[i + 64]);
    if (fused) {
      float sn, cs;
      sincosf((float)pos_new * p.inv_freq[i], &sn, &cs);
      const float y0 = x0 * cs - x1 * sn, y1 = x1 * cs + x0 * sn;
      x0 = y0;
      x1 = y1;
    }
    *reinterpret_cast<bf16*>(sQ + swz(r, i >> 3) + (i & 7) * 2) = __float2bfloat16(x0);
    *reinterpret_cast<bf16*>(sQ + swz(r, (i + 64) >> 3) + (i & 7) * 2) = __float2bfloat16(x1);
  }
  __syncthreads();
  uint32_t qf[8][4];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    const int r = (lane & 7) + (((lane >> 3) & 1) << 3);
    ldmatrix_x4(smem_u32(sQ) + swz(r, ks * 2 + (lane >> 4)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
  }

  float o[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;  // row g only (rows >= 8 are padding)

  for (; tile < t_end; tile += 4, ++it) {
    const int st = it % kDecStages;
    if (tile + 4 * (kDecStages - 1) < t_end) issue(tile + 4 * (kDecStages - 1), (it + kDecStages - 1) % kDecStages);
    cp_async_commit();
    cp_async_wait<kDecStages - 1>();
    __syncwarp();
    const int key0 = tile * kDecTile;
    if (fused && pos_new >= key0 && pos_new < key0 + kDecTile) {
      // the new token lives in this tile: rotate its K, take its V from the qkv row, patch the smem tile and append
      // both to the cache (the stale cache slot fetched by cp.async is overwritten here before any use)
      const int r = pos_new - key0;
      const bf16* kp = qrow + (p.Hq + kvh) * kHD;
      const bf16* vp = qrow + (p.Hq + p.Hkv + kvh) * kHD;
      const int page = p.block_table[(long long)b * p.max_pages + pos_new / p.page_size];
      const int slot = pos_new % p.page_size;
      bf16* kdst = p.pool + ((((long long)page * 2 + 0) * p.Hkv + kvh) * p.page_size + slot) * kHD;
      bf16* vdst = p.pool + ((((long long)page * 2 + 1) * p.Hkv + kvh) * p.page_size + slot) * kHD;
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int i = lane + h2 * 32;
        float sn, cs;
        sincosf((float)pos_new * p.inv_freq[i], &sn, &cs);
        const float x0 = __bfloat162float(kp[i]), x1 = __bfloat162float(kp[i + 64]);
        const bf16 y0 = __float2bfloat16(x0 * cs - x1 * sn), y1 = __float2bfloat16(x1 * cs + x0 * sn);
        *reinterpret_cast<bf16*>(stage_k(st) + swz(r, i >> 3) + (i & 7) * 2) = y0;
        *reinterpret_cast<bf16*>(stage_k(st) + swz(r, (i + 64) >> 3) + (i & 7) * 2) = y1;
        kdst[i] = y0;
        kdst[i + 64] = y1;
      }
      if (lane < 16) {
        const uint4 vv = reinterpret_cast<const uint4*>(vp)[lane];
        *reinterpret_cast<uint4*>(stage_v(st) + swz(r, lane)) = vv;
        reinterpret_cast<uint4*>(vdst)[lane] = vv;
      }
      __syncwarp();
    }
    float s[2][4];
    s[0][0] = s[0][1] = s[0][2] = s[0][3] = s[1][0] = s[1][1] = s[1][2] = s[1][3] = 0.f;
    qk_tile<2>(qf, smem_u32(stage_k(st)), 0, s);
    if (key0 + kDecTile > ctx) {
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (key0 + nt * 8 + tq * 2 + (e & 1) >= ctx) s[nt][e] = -INFINITY;
    }
    float mx = fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1]));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_new = fmaxf(m_run, mx);
    const float msc = (m_new == -INFINITY) ? 0.f : m_new * p.scale_log2;
    const float corr = (m_run == -INFINITY) ? 0.f : exp2f(m_run * p.scale_log2 - msc);
    m_run = m_new;
    uint32_t pf[1][4];
    const float p0 = exp2f(s[0][0] * p.scale_log2 - msc), p1 = exp2f(s[0][1] * p.scale_log2 - msc);
    const float p2 = exp2f(s[1][0] * p.scale_log2 - msc), p3 = exp2f(s[1][1] * p.scale_log2 - msc);
    l_run = l_run * corr + (p0 + p1 + p2 + p3);
    pf[0][0] = pack_bf16(p0, p1);
    pf[0][1] = 0u;  // padding rows g+8
    pf[0][2] = pack_bf16(p2, p3);
    pf[0][3] = 0u;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      o[i][0] *= corr;
      o[i][1] *= corr;
    }
    pv_tile<1>(pf, smem_u32(stage_v(st)), 0, o);
    __syncwarp();
  }
  cp_async_wait<0>();
  __syncthreads();  // every warp is done with its tiles: the merge scratch may overwrite warp 0's

  // ---- merge the 4 warps through shared memory
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
  {
    float* mine = sPart + (warp * 8 + g) * kDecPartStride;
#pragma unroll
    for (int dt = 0; dt < 16; ++dt) {
      mine[dt * 8 + tq * 2] = o[dt][0];
      mine[dt * 8 + tq * 2 + 1] = o[dt][1];
    }
    if (tq == 0) {
      mine[kHD] = m_run;
      mine[kHD + 1] = l_run;
    }
  }
  __syncthreads();
  // thread -> (row r < G, 8 consecutive d): 128 threads cover 8 rows x 16 chunks
  const int r = tid >> 4, dch = tid & 15;
  float acc[8], m_tot = -INFINITY, l_tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  if (r < G) {
    for (int w = 0; w < 4; ++w) m_tot = fmaxf(m_tot, sPart[(w * 8 + r) * kDecPartStride + kHD]);
    for (int w = 0; w < 4; ++w) {
      const float* pw = sPart + (w * 8 + r) * kDecPartStride;
      const float mw = pw[kHD];
      const float sc = (mw == -INFINITY) ? 0.f : exp2f((mw - m_tot) * p.scale_log2);
      l_tot += pw[kHD + 1] * sc;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += pw[dch * 8 + i] * sc;
    }
  }
  const int head = kvh * G + r;
  if (p.splits == 1) {
    if (r < G) {
      const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
      uint4 ov = make_uint4(pack_bf16(acc[0] * inv, acc[1] * inv), pack_bf16(acc[2] * inv, acc[3] * inv),
                            pack_bf16(acc[4] * inv, acc[5] * inv), pack_bf16(acc[6] * inv, acc[7] * inv));
      *reinterpret_cast<uint4*>(p.out + (long long)b * p.ldo + head * kHD + dch * 8) = ov;
    }
    return;
  }
  float* wsb = p.ws + (((long long)(b * p.Hkv + kvh) * p.splits) * G) * kDecPartStride;
  if (r < G) {
    float* dst = wsb + ((long long)split * G + r) * kDecPartStride;
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[dch * 8 + i] = acc[i];
    if (dch == 0) {
      dst[kHD] = m_tot;
      dst[kHD + 1] = l_tot;
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int prev = atomicAdd(&p.counters[b * p.Hkv + kvh], 1u);
    s_is_last = (prev == (unsigned int)p.splits - 1) ? 1 : 0;
    if (s_is_last) p.counters[b * p.Hkv + kvh] = 0u;  // self-reset for the next launch / graph replay
  }
  __syncthreads();
  if (!s_is_last) return;
  __threadfence();
  if (r < G) {
    float mt = -INFINITY, lt = 0.f;
    for (int sp = 0; sp < p.splits; ++sp) mt = fmaxf(mt, __ldcg(wsb + ((long long)sp * G + r) * kDecPartStride + kHD));
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int sp = 0; sp < p.splits; ++sp) {
      const float* ps = wsb + ((long long)sp * G + r) * kDecPartStride;
      const float ms = __ldcg(ps + kHD);
      const float sc = (ms == -INFINITY) ? 0.f : exp2f((ms - mt) * p.scale_log2);
      lt += __ldcg(ps + kHD + 1) * sc;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += __ldcg(ps + dch * 8 + i) * sc;
    }
    const float inv = lt > 0.f ? 1.f / lt : 0.f;
    uint4 ov = make_uint4(pack_bf16(acc[0] * inv, acc[1] * inv), pack_bf16(acc[2] * inv, acc[3] * inv),
                          pack_bf16(acc[4] * inv, acc[5] * inv), pack_bf16(acc[6] * inv, acc[7] * inv));
    *reinterpret_cast<uint4*>(p.out + (long long)b * p.ldo + head * kHD + dch * 8) = ov;
  }
}

}  // namespace omc

using namespace omc;

namespace omc {
int launch_fa_sm100(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* out,
                    long long ldo, const int32_t* cu, int num_seqs, int max_seqlen, long long total_rows, int Hq, int Hkv,
                    int head_dim, int causal, float scale_log2, cudaStream_t stream);
}

namespace omc { void set_fa_prof(void* ptr); void set_fa_version(int v); }
/* diagnostic: device buffer of 12 x 8 uint64 that CTA (0,0,0) of the tcgen05 attention kernel fills with per-warp clock
 * accumulators (softmax warps: wait S, TMEM load, max/rescale, exp, store+signal; MMA warp: wait P); NULL = off */
extern "C" int omc_attention_set_prof(void* dev_buf) {
  omc::set_fa_prof(dev_buf);
  return OMC_OK;
}
static int g_attn_impl = -1;  // 0 = tcgen05 kernel for full tiles + mma.sync tail, 1 = mma.sync kernel for everything
extern "C" int omc_attention_set_impl(int impl) {
  if (impl < 0 || impl > 2)
    return set_error(OMC_ERR_ARG, "omc_attention_set_impl: 0 (tcgen05, default), 1 (mma.sync) or 2 (tcgen05 with 128-key tiles)");
  g_attn_impl = impl == 1 ? 1 : 0;
  omc::set_fa_version(impl == 2 ? 1 : 2);
  return OMC_OK;
}

extern "C" int omc_attention_fwd_hd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                                    void* out, long long ldo, const int32_t* cu_seqlens, int num_seqs, int max_seqlen,
                                    long long total_rows, int Hq, int Hkv, int head_dim, int causal, float scale, void* stream);

extern "C" int omc_attention_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                                 void* out, long long ldo, const int32_t* cu_seqlens, int num_seqs, int max_seqlen,
                                 long long total_rows, int Hq, int Hkv, int causal, float scale, void* stream) {
  return omc_attention_fwd_hd(q, ldq, k, ldk, v, ldv, out, ldo, cu_seqlens, num_seqs, max_seqlen, total_rows, Hq, Hkv, 128, causal,
                              scale, stream);
}

extern "C" int omc_attention_fwd_hd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                                    void* out, long long ldo, const int32_t* cu_seqlens, int num_seqs, int max_seqlen,
                                    long long total_rows, int Hq, int Hkv, int head_dim, int causal, float scale, void* stream) {
  if (num_seqs <= 0 || max_seqlen <= 0) return OMC_OK;
  if (head_dim != 128 && head_dim != 64) return set_error(OMC_ERR_SHAPE, "omc_attention_fwd: head_dim must be 128 or 64");
  if (Hq <= 0 || Hkv <= 0 || Hq % Hkv != 0) return set_error(OMC_ERR_SHAPE, "omc_attention_fwd: Hq must be a multiple of Hkv");
  if ((ldq | ldk | ldv | ldo) % 8 != 0) return set_error(OMC_ERR_ALIGN, "omc_attention_fwd: row strides must be multiples of 8");
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[cur_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttSmem);
    if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  AttnParams p;
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.out = (bf16*)out;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
  p.cu = cu_seqlens; p.Hq = Hq; p.Hkv = Hkv; p.causal = causal;
  p.scale_log2 = scale * kLog2e;
  p.tail_mode = 0;
  if (g_attn_impl < 0) {
    const char* e = getenv("OMCHAT_B200_ATTN_LEGACY");
    g_attn_impl = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  if (g_attn_impl == 0 && total_rows > 0)  // tcgen05 / TMEM kernel (attention_sm100.cu)
    return launch_fa_sm100(q, ldq, k, ldk, v, ldv, out, ldo, cu_seqlens, num_seqs, max_seqlen, total_rows, Hq, Hkv,
                           head_dim, causal, p.scale_log2, (cudaStream_t)stream);
  if (head_dim != 128) return set_error(OMC_ERR_SHAPE, "omc_attention_fwd: the mma.sync baseline kernel is head_dim 128 only");
  dim3 grid((max_seqlen + kAttM - 1) / kAttM, Hq, num_seqs);
  attention_fwd_kernel<<<grid, kAttThreads, kAttSmem, (cudaStream_t)stream>>>(p);
  return check_launch("attention_fwd");
}

extern "C" int omc_decode_attn_splits(int B, int Hkv, int max_ctx) {
  if (B <= 0 || Hkv <= 0 || max_ctx <= 0) return 1;
  const int tiles = (max_ctx + kDecTile - 1) / kDecTile;
  int by_fill = (2 * num_sms()) / (B * Hkv);  // two 100 KB CTAs per SM: fill ONE wave (rounding up made 1.3 waves at batch 32)
  int by_work = (tiles + 7) / 8;  // >= 2 tiles per warp
  int s = by_fill < by_work ? by_fill : by_work;
  if (s < 1) s = 1;
  if (s > 64) s = 64;
  return s;
}

extern "C" long long omc_decode_attn_workspace_bytes(int B, int Hq, int Hkv, int splits) {
  // fp32 partials [B, Hkv, splits, G, 130] followed by B*Hkv uint32 counters (must be zeroed once by the caller)
  const int G = Hq / Hkv;
  return (long long)B * Hkv * splits * G * kDecPartStride * 4 + (long long)B * Hkv * 4;
}

extern "C" int omc_paged_decode_attn(const void* qkv, long long ldq, const float* inv_freq, void* kv_pool,
                                     const int32_t* block_table, int max_pages, int page_size, const int32_t* ctx_lens,
                                     int B, int Hq, int Hkv, int splits, float scale, void* out, long long ldo,
                                     void* workspace, void* stream) {
  if (B <= 0) return OMC_OK;
  if (Hq % Hkv != 0 || Hq / Hkv > 8) return set_error(OMC_ERR_SHAPE, "omc_paged_decode_attn: group size must be <= 8");
  if (page_size % kDecTile != 0) return set_error(OMC_ERR_SHAPE, "omc_paged_decode_attn: page_size must be a multiple of 16");
  if (splits < 1) return set_error(OMC_ERR_ARG, "omc_paged_decode_attn: splits < 1");
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[cur_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(paged_decode_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDecSmem);
    if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  DecParams p;
  p.qkv = (const bf16*)qkv; p.ldq = ldq; p.pool = (bf16*)kv_pool; p.block_table = block_table;
  p.max_pages = max_pages; p.page_size = page_size; p.ctx_lens = ctx_lens; p.inv_freq = inv_freq;
  p.Hq = Hq; p.Hkv = Hkv; p.G = Hq / Hkv; p.splits = splits; p.scale_log2 = scale * kLog2e;
  p.out = (bf16*)out; p.ldo = ldo;
  const int G = Hq / Hkv;
  p.ws = (float*)workspace;
  p.counters = reinterpret_cast<unsigned int*>((float*)workspace + (long long)B * Hkv * splits * G * kDecPartStride);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(splits, Hkv, B);
  cfg.blockDim = dim3(kDecThreads);
  cfg.dynamicSmemBytes = kDecSmem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;  // a no-op unless the previous kernel triggers early
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, paged_decode_attn_kernel, p);
  if (le != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(le));
  return check_launch("paged_decode_attn");
}
