// Flash attention forward for sm_100a on the 5th-generation tensor cores (tcgen05 + TMEM), head_dim 128 (and 64 in the v2
// kernel: the InternViT-300M tower), bf16.
//
// Serves the same call sites as the mma.sync kernel of attention.cu (InternAttention._flash_attn
// intern_vit_6b/modeling_intern_vit.py:157-172 — non-causal, 1025 tokens, 25 heads — and the causal GQA attention of the
// Qwen2 prefill, transformers modeling_qwen2.py:161-184,206-246) for every FULL 128-row query tile of a packed var-len
// batch. A ragged last query tile is computed like a full one (its extra rows read whatever rows follow in the packed
// buffer, or TMA zero fill) and only the rows inside the sequence are stored; attention.cu keeps the mma.sync kernel as the
// A/B baseline.
//
// One CTA = one (sequence, head, 256-query block) = two 128-row query tiles that ping-pong on the tensor pipe:
//   warp 0        TMA producer: Q tiles once, then K_j / V_j tiles (128 keys x 128 dims, two 64-column 128B-swizzled boxes
//                 each) into 2-stage rings
//   warp 1        MMA issuer (one elected thread): S_t = Q_t K_j^T (SS, K-major operands, fp32 S in TMEM) and
//                 O_t += P_t V_j (TS: P is read from TENSOR MEMORY, V is the MN-major B operand straight from the TMA tile),
//                 ordered S0 S1 | PV0 S0' PV1 S1' | ... so the pipe works on one tile while the other is in softmax
//   warps 4-7     softmax of tile 0, warps 8-11 softmax of tile 1: thread = query row = TMEM lane. Two passes over S in
//                 TMEM (row max, then exp2 / row sum), P written back as packed bf16 over the first 64 columns of S, the O
//                 accumulator rescaled in TMEM only when the running max grew by more than 2^8 (exact: the final
//                 normalisation uses the same reference max)
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); P_t aliases S_t[0,64).
// The last K/V tile of a sequence uses a narrower MMA (N, K rounded up to 16 keys) so a 1025-key sequence does not pay for
// a ninth full tile.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "omc_internal.h"
#include "ptx.cuh"

namespace omc {

int make_tmap_2d(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows);

typedef __nv_bfloat16 bf16;

constexpr int kFaThreads = 384;
constexpr int kFaTile = 128;                    // query rows per tile = keys per K/V tile = head_dim
constexpr int kFaTileBytes = kFaTile * 128 * 2;  // 32 KB: two [128 rows x 64 cols] 128B-swizzled boxes
constexpr int kFaHalfBytes = kFaTileBytes / 2;
constexpr int kFaStages = 2;
constexpr int kFaSmem = 2 * kFaTileBytes + 2 * kFaStages * kFaTileBytes + 1024 + 1024;  // Q0 Q1 | K ring | V ring | bars | align
constexpr float kFaRescaleLog2 = 8.0f;

struct FaParams {
  const int32_t* cu;
  bf16* out;
  long long ldo;
  int Hq, Hkv, causal;
  float scale_log2;
  unsigned long long* prof;  // optional [12 warps][8] clock accumulators of CTA (0,0,0) (omc_attention_set_prof)
};
#define FA_CLK(var) const long long var = prof_on ? clock64() : 0

// packed fp32x2 arithmetic (sm_100 FFMA2 / FADD2): halves the issue slots of the softmax scale and row-sum
__device__ __forceinline__ uint64_t f2_pack(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kFaThreads, 1)
fa_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const FaParams p) {
  const int seq = blockIdx.z, head = blockIdx.y;
  const int row0 = p.cu[seq], len = p.cu[seq + 1] - row0;
  const int n_tiles = (len + kFaTile - 1) / kFaTile;  // the last tile may be ragged: its rows >= len are computed on
  const int nblk = (n_tiles + 1) >> 1;                // whatever the TMA box brings in and never stored
  const int blk = p.causal ? ((int)gridDim.x - 1 - (int)blockIdx.x) : (int)blockIdx.x;  // heavy causal blocks first
  if (blk >= nblk) return;  // uniform for the CTA, before any barrier / allocation
  const int ntile = min(2, n_tiles - blk * 2);
  const int q0 = blk * 2 * kFaTile;
  const int kvh = head / (p.Hq / p.Hkv);
  int kv_len[2], nkv[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    kv_len[t] = p.causal ? min(len, q0 + (t + 1) * kFaTile) : len;
    nkv[t] = (t < ntile) ? (kv_len[t] + kFaTile - 1) / kFaTile : 0;
  }
  const int nkv_max = max(nkv[0], nkv[1]);

  extern __shared__ uint8_t fa_smem_raw[];
  const uint32_t raw_addr = smem_u32(fa_smem_raw);
  uint8_t* smem = fa_smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;                                   // [2][32 KB]
  uint8_t* sK = smem + 2 * kFaTileBytes;                // [stages][32 KB]
  uint8_t* sV = sK + kFaStages * kFaTileBytes;          // [stages][32 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kFaStages * kFaTileBytes);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // [2]
  uint64_t* k_empty = bars + 3;            // [2]
  uint64_t* v_full = bars + 5;             // [2]
  uint64_t* v_empty = bars + 7;            // [2]
  uint64_t* s_full = bars + 9;             // [2 tiles]
  uint64_t* p_ready = bars + 11;           // [2 tiles]
  uint64_t* o_final = bars + 13;           // [2 tiles]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool prof_on = p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  long long pa[6] = {0, 0, 0, 0, 0, 0};
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    if (elect_one()) {
      mbar_init(q_full, 1);
      for (int s = 0; s < kFaStages; ++s) {
        mbar_init(&k_full[s], 1);
        mbar_init(&k_empty[s], 1);
        mbar_init(&v_full[s], 1);
        mbar_init(&v_empty[s], 1);
      }
      for (int t = 0; t < 2; ++t) {
        mbar_init(&s_full[t], 1);
        mbar_init(&p_ready[t], 4);  // one arrival per softmax warp
        mbar_init(&o_final[t], 1);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================================ TMA producer ============================================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, (uint32_t)(ntile * kFaTileBytes));
      for (int t = 0; t < ntile; ++t)
        for (int h = 0; h < 2; ++h)
          tma_load_2d(sQ + t * kFaTileBytes + h * kFaHalfBytes, &tmQ, q_full, head * 128 + h * 64,
                      row0 + q0 + t * kFaTile, kEvictNormal);
      for (int j = 0; j < nkv_max; ++j) {
        const int s = j & 1;
        const uint32_t ph = (uint32_t)(j >> 1) & 1u;
        const int krow = row0 + j * kFaTile;
        mbar_wait(&k_empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&k_full[s], (uint32_t)kFaTileBytes);
        for (int h = 0; h < 2; ++h)
          tma_load_2d(sK + s * kFaTileBytes + h * kFaHalfBytes, &tmK, &k_full[s], kvh * 128 + h * 64, krow, kEvictLast);
        mbar_wait(&v_empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&v_full[s], (uint32_t)kFaTileBytes);
        for (int h = 0; h < 2; ++h)
          tma_load_2d(sV + s * kFaTileBytes + h * kFaHalfBytes, &tmV, &v_full[s], kvh * 128 + h * 64, krow, kEvictLast);
      }
    }
  } else if (warp == 1) {
    // ============================================ MMA issuer ============================================
    if (elect_one()) {
      auto keys_padded = [&](int t, int j) {  // keys of tile j that query tile t multiplies, rounded up to the MMA granule
        const int n = min(kFaTile, kv_len[t] - j * kFaTile);
        return (n + 15) & ~15;
      };
      auto issue_s = [&](int t, int j) {
        const uint32_t idesc = make_idesc_bf16_major(128, keys_padded(t, j), 0, 0);
        const uint32_t qa = smem_u32(sQ + t * kFaTileBytes), ka = smem_u32(sK + (j & 1) * kFaTileBytes);
        const uint32_t d = tmem_base + (uint32_t)(t * 128);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // 16 head dims per instruction: half = kk / 4, 32 B steps inside the swizzle atom
          const uint64_t da = make_sw128_kmajor_desc(qa + (kk >> 2) * kFaHalfBytes) + (uint64_t)(2 * (kk & 3));
          const uint64_t db = make_sw128_kmajor_desc(ka + (kk >> 2) * kFaHalfBytes) + (uint64_t)(2 * (kk & 3));
          umma_bf16<1>(d, da, db, idesc, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t]);
      };
      auto issue_pv = [&](int t, int j) {
        constexpr uint32_t idesc = make_idesc_bf16_major(128, 128, 0, 1);  // B = V is MN-major (head dim contiguous)
        const uint32_t va = smem_u32(sV + (j & 1) * kFaTileBytes);
        const uint32_t d = tmem_base + 256u + (uint32_t)(t * 128), pa = tmem_base + (uint32_t)(t * 128);
        const int ksteps = keys_padded(t, j) >> 4;
        for (int kk = 0; kk < ksteps; ++kk) {  // 16 keys per instruction = 8 packed P columns, 16 V rows (2 KB)
          const uint64_t db = make_sw128_mnmajor_desc(va + kk * 2048, kFaHalfBytes, 1024);
          umma_bf16_ts(d, pa + (uint32_t)(kk * 8), db, idesc, (j > 0 || kk > 0) ? 1u : 0u);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < ntile; ++t) issue_s(t, 0);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < nkv_max; ++j) {
        bool v_waited = false, k_waited = false;
        for (int t = 0; t < ntile; ++t) {
          if (j < nkv[t]) {
            FA_CLK(c0);
            mbar_wait(&p_ready[t], (uint32_t)j & 1u);
            FA_CLK(c1);
            pa[0] += c1 - c0;
            if (!v_waited) {
              mbar_wait(&v_full[j & 1], (uint32_t)(j >> 1) & 1u);
              v_waited = true;
            }
            tc_fence_after();
            issue_pv(t, j);
            if (j == nkv[t] - 1) umma_commit(&o_final[t]);
          }
          if (j + 1 < nkv[t]) {
            if (!k_waited) {
              mbar_wait(&k_full[(j + 1) & 1], (uint32_t)((j + 1) >> 1) & 1u);
              k_waited = true;
              tc_fence_after();
            }
            issue_s(t, j + 1);
          }
        }
        umma_commit(&v_empty[j & 1]);
        if (j + 1 < nkv_max) umma_commit(&k_empty[(j + 1) & 1]);
      }
      if (prof_on) p.prof[1 * 8 + 0] = (unsigned long long)pa[0];
    }
  } else if (warp >= 4) {
    // ============================================ softmax / epilogue ============================================
    const int t = (warp - 4) >> 2;
    if (t < ntile) {
      const int quarter = warp & 3;
      const int r = quarter * 32 + lane;       // row of the tile = TMEM lane
      const int qi = q0 + t * kFaTile + r;     // query index inside the sequence
      const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
      const uint32_t tS = tmem_base + lane_addr + (uint32_t)(t * 128);
      const uint32_t tO = tmem_base + lane_addr + 256u + (uint32_t)(t * 128);
      const float sl2 = p.scale_log2;
      float m_ref = -INFINITY, l_sum = 0.f;
      for (int j = 0; j < nkv[t]; ++j) {
        FA_CLK(c0);
        mbar_wait(&s_full[t], (uint32_t)j & 1u);
        tc_fence_after();
        FA_CLK(c1);
        pa[0] += c1 - c0;
        const int kbase = j * kFaTile;
        const int nvalid = min(kFaTile, kv_len[t] - kbase);
        const int ncols = (nvalid + 15) & ~15;
        int lim = nvalid - 1;                       // columns 0..lim of this tile are visible to this row
        if (p.causal) lim = min(lim, qi - kbase);
        const bool masked = !__all_sync(0xffffffffu, lim >= ncols - 1);
        if (!masked && ncols == kFaTile) {
          // ===== fast path (full, unmasked tile): ONE pass, the whole S row in registers =====
          uint32_t v[128];
          tmem_ld32(tS, v);
          tmem_ld32(tS + 32u, v + 32);
          tmem_ld32(tS + 64u, v + 64);
          tmem_ld32(tS + 96u, v + 96);
          tmem_ld_wait();
          FA_CLK(c2);
          pa[1] += c2 - c1;
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 128; i += 8) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              mx4[c] = max3(mx4[c], __uint_as_float(v[i + 2 * c]), __uint_as_float(v[i + 2 * c + 1]));
          }
          const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
          if (j == 0) {
            m_ref = mx;
          } else {
            const bool grow = (mx - m_ref) * sl2 > kFaRescaleLog2;
            if (__any_sync(0xffffffffu, grow)) {
              const float m_use = grow ? mx : m_ref;
              const float alpha = ex2((m_ref - m_use) * sl2);
              l_sum *= alpha;
              m_ref = m_use;
#pragma unroll 1
              for (int c0 = 0; c0 < 128; c0 += 32) {
                uint32_t o[32];
                tmem_ld32(tO + (uint32_t)c0, o);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                tmem_st32(tO + (uint32_t)c0, o);
              }
            }
          }
          const float neg_m = -m_ref * sl2;
          FA_CLK(c3);
          pa[2] += c3 - c2;
          const uint64_t sl2_2 = f2_pack(sl2, sl2), negm_2 = f2_pack(neg_m, neg_m);
          uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};
          uint32_t pk[64];
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            const uint64_t x = f2_fma(f2_pack(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sl2_2, negm_2);
            float x0, x1;
            f2_unpack(x, x0, x1);
            const float e0 = ex2(x0), e1 = ex2(x1);
            acc2[i & 3] = f2_add(acc2[i & 3], f2_pack(e0, e1));
            pk[i] = pack_bf16(e0, e1);
          }
          FA_CLK(c4);
          pa[3] += c4 - c3;
          tmem_st32(tS, pk);
          tmem_st32(tS + 32u, pk + 32);
          {
            float a0, a1, b0, b1;
            f2_unpack(f2_add(f2_add(acc2[0], acc2[1]), f2_add(acc2[2], acc2[3])), a0, a1);
            (void)b0; (void)b1;
            l_sum += a0 + a1;
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_ready[t]);
          FA_CLK(c5);
          pa[4] += c5 - c4;
          continue;
        }
        // ===== general path (ragged last K/V tile, causal diagonal): two chunked passes with masking =====
        // ---- pass 1: row maximum
        float mx = -INFINITY;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t v[32];
          if (ncols - c0 >= 32) {
            tmem_ld32(tS + (uint32_t)c0, v);
          } else {
            tmem_ld16(tS + (uint32_t)c0, v);
#pragma unroll
            for (int i = 16; i < 32; ++i) v[i] = 0xff800000u;  // -inf
          }
          tmem_ld_wait();
          if (masked) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c0 + i > lim) v[i] = 0xff800000u;
          }
          float a = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), b = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
#pragma unroll
          for (int i = 4; i < 32; i += 4) {
            a = fmaxf(a, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
            b = fmaxf(b, fmaxf(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])));
          }
          mx = fmaxf(mx, fmaxf(a, b));
        }
        // ---- running reference maximum; O is rescaled only when it grew by more than 2^8
        if (j == 0) {
          m_ref = mx;
        } else {
          const bool grow = (mx - m_ref) * sl2 > kFaRescaleLog2;
          if (__any_sync(0xffffffffu, grow)) {
            const float m_use = grow ? mx : m_ref;
            const float alpha = ex2((m_ref - m_use) * sl2);
            l_sum *= alpha;
            m_ref = m_use;
#pragma unroll 1
            for (int c0 = 0; c0 < 128; c0 += 32) {
              uint32_t v[32];
              tmem_ld32(tO + (uint32_t)c0, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
              tmem_st32(tO + (uint32_t)c0, v);
            }
          }
        }
        const float neg_m = -m_ref * sl2;
        // ---- pass 2: p = exp2(s * scale_log2 - m), row sum, packed bf16 P over the first columns of S
        float l0 = 0.f, l1 = 0.f;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t v[32];
          const bool wide = ncols - c0 >= 32;
          if (wide) {
            tmem_ld32(tS + (uint32_t)c0, v);
          } else {
            tmem_ld16(tS + (uint32_t)c0, v);
#pragma unroll
            for (int i = 16; i < 32; ++i) v[i] = 0xff800000u;
          }
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float e0 = ex2(fmaf(__uint_as_float(v[2 * i]), sl2, neg_m));
            float e1 = ex2(fmaf(__uint_as_float(v[2 * i + 1]), sl2, neg_m));
            if (masked) {
              if (c0 + 2 * i > lim) e0 = 0.f;
              if (c0 + 2 * i + 1 > lim) e1 = 0.f;
            }
            l0 += e0;
            l1 += e1;
            pk[i] = pack_bf16(e0, e1);
          }
          if (wide) tmem_st16(tS + (uint32_t)(c0 >> 1), pk);
          else tmem_st8(tS + (uint32_t)(c0 >> 1), pk);
        }
        l_sum += l0 + l1;
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[t]);
      }
      // ---- epilogue: O / l -> bf16 -> global (each thread owns one output row: 256 contiguous bytes)
      if (prof_on && lane == 0)
        for (int i = 0; i < 5; ++i) p.prof[warp * 8 + i] = (unsigned long long)pa[i];
      mbar_wait(&o_final[t], 0);
      tc_fence_after();
      const float inv = 1.0f / l_sum;
      bf16* orow = p.out + (long long)(row0 + qi) * p.ldo + head * 128;
      const bool row_ok = qi < len;  // rows of a ragged last tile beyond the sequence are not stored
#pragma unroll 1
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tO + (uint32_t)c0, v);
        tmem_ld_wait();
        if (!row_ok) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
          o.y = pack_bf16(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
          o.z = pack_bf16(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
          o.w = pack_bf16(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c0 + 8 * i) = o;
        }
      }
    }
  }
  // ============================================ teardown ============================================
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}


// =====================================================================================================================
// v2: 64-key K/V tiles, S double-buffered in TMEM. The MMA issuer computes S_t[j+1] into the other S buffer BEFORE it
// waits for P_t[j], so the softmax warps find their next S tile ready the moment they have written P (v1 made them wait for
// their own P.V and S MMAs: ~1600 of 3800 cycles per 128-key step, tools/attn_check.py clocks). TMEM: S_t,b at columns
// t*128 + b*64 (P_t,b = packed bf16 over its first 32 columns), O_t at 256 + t*128. Since S[j+1] is now issued ahead of
// P.V[j-1]'s completion, the O accumulator gets its own barrier: o_done[t] completes after every P.V and is waited once per
// step (at its end, when the previous P.V has long finished), which also orders the rare lazy rescale of O.
constexpr uint32_t kFaPolyMask = 0x52;             // pairs (of every 8) whose exp2 runs on the FMA pipe: {1, 4, 6} = 3/8
constexpr int kKv2 = 64;                          // keys per K/V tile
constexpr int kKv2Half = kKv2 * 64 * 2;           // 8 KB: one [64 keys x 64 dims] 128B-swizzled box
constexpr int kFa2Stages = 4;
// HD = head_dim: 128 (InternViT-6B, Qwen2) = two 64-column boxes per tile, or 64 (InternViT-300M) = one. Everything that is a
// function of it - tile bytes, boxes per tile, k-steps of S = Q K^T, the N of O += P V, the O columns - hangs off this struct.
template <int HD>
struct Fa2Cfg {
  static_assert(HD == 64 || HD == 128, "head_dim 64 or 128");
  static constexpr int kBoxes = HD / 64;
  static constexpr int kQBytes = kFaTile * HD * 2;   // 32 / 16 KB
  static constexpr int kKvBytes = kKv2 * HD * 2;     // 16 / 8 KB
  static constexpr int kSmem = 2 * kQBytes + 2 * kFa2Stages * kKvBytes + 1024 + 1024;
};

template <int HD>
__global__ void __launch_bounds__(kFaThreads, 1)
fa_fwd_sm100_v2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const FaParams p) {
  using C = Fa2Cfg<HD>;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int row0 = p.cu[seq], len = p.cu[seq + 1] - row0;
  const int n_tiles = (len + kFaTile - 1) / kFaTile;
  const int nblk = (n_tiles + 1) >> 1;
  const int blk = p.causal ? ((int)gridDim.x - 1 - (int)blockIdx.x) : (int)blockIdx.x;
  if (blk >= nblk) return;
  const int ntile = min(2, n_tiles - blk * 2);
  const int q0 = blk * 2 * kFaTile;
  const int kvh = head / (p.Hq / p.Hkv);
  int kv_len[2], nkv[2];
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    kv_len[t] = p.causal ? min(len, q0 + (t + 1) * kFaTile) : len;
    nkv[t] = (t < ntile) ? (kv_len[t] + kKv2 - 1) / kKv2 : 0;
  }
  const int nkv_max = max(nkv[0], nkv[1]);

  extern __shared__ uint8_t fa_smem_raw[];
  const uint32_t raw_addr = smem_u32(fa_smem_raw);
  uint8_t* smem = fa_smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 2 * C::kQBytes;
  uint8_t* sV = sK + kFa2Stages * C::kKvBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kFa2Stages * C::kKvBytes);
  uint64_t* q_full = bars;                  // 1
  uint64_t* k_full = bars + 1;              // [4]
  uint64_t* k_empty = bars + 5;             // [4]
  uint64_t* v_full = bars + 9;              // [4]
  uint64_t* v_empty = bars + 13;            // [4]
  uint64_t* s_full = bars + 17;             // [tile][buffer]
  uint64_t* p_ready = bars + 21;            // [tile][buffer]
  uint64_t* o_done = bars + 25;             // [tile]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool prof_on = p.prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
  long long pa[6] = {0, 0, 0, 0, 0, 0};
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 1) {
    if (elect_one()) {
      mbar_init(q_full, 1);
      for (int s = 0; s < kFa2Stages; ++s) {
        mbar_init(&k_full[s], 1);
        mbar_init(&k_empty[s], 1);
        mbar_init(&v_full[s], 1);
        mbar_init(&v_empty[s], 1);
      }
      for (int i = 0; i < 4; ++i) {
        mbar_init(&s_full[i], 1);
        mbar_init(&p_ready[i], 4);
      }
      mbar_init(&o_done[0], 1);
      mbar_init(&o_done[1], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================================ TMA producer ============================================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, (uint32_t)(ntile * C::kQBytes));
      for (int t = 0; t < ntile; ++t)
        for (int h = 0; h < C::kBoxes; ++h)
          tma_load_2d(sQ + t * C::kQBytes + h * kFaHalfBytes, &tmQ, q_full, head * HD + h * 64,
                      row0 + q0 + t * kFaTile, kEvictNormal);
      for (int j = 0; j < nkv_max; ++j) {
        const int s = j % kFa2Stages;
        const uint32_t ph = (uint32_t)(j / kFa2Stages) & 1u;
        const int krow = row0 + j * kKv2;
        mbar_wait(&k_empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&k_full[s], (uint32_t)C::kKvBytes);
        for (int h = 0; h < C::kBoxes; ++h)
          tma_load_2d(sK + s * C::kKvBytes + h * kKv2Half, &tmK, &k_full[s], kvh * HD + h * 64, krow, kEvictLast);
        mbar_wait(&v_empty[s], ph ^ 1u);
        mbar_arrive_expect_tx(&v_full[s], (uint32_t)C::kKvBytes);
        for (int h = 0; h < C::kBoxes; ++h)
          tma_load_2d(sV + s * C::kKvBytes + h * kKv2Half, &tmV, &v_full[s], kvh * HD + h * 64, krow, kEvictLast);
      }
    }
  } else if (warp == 1) {
    // ============================================ MMA issuer ============================================
    if (elect_one()) {
      auto keys_padded = [&](int t, int j) {
        const int n = min(kKv2, kv_len[t] - j * kKv2);
        return (n + 15) & ~15;
      };
      auto issue_s = [&](int t, int j) {
        const uint32_t idesc = make_idesc_bf16_major(128, keys_padded(t, j), 0, 0);
        const uint32_t qa = smem_u32(sQ + t * C::kQBytes), ka = smem_u32(sK + (j % kFa2Stages) * C::kKvBytes);
        const uint32_t d = tmem_base + (uint32_t)(t * 128 + (j & 1) * 64);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {  // 16 head dims per instruction
          const uint64_t da = make_sw128_kmajor_desc(qa + (kk >> 2) * kFaHalfBytes) + (uint64_t)(2 * (kk & 3));
          const uint64_t db = make_sw128_kmajor_desc(ka + (kk >> 2) * kKv2Half) + (uint64_t)(2 * (kk & 3));
          umma_bf16<1>(d, da, db, idesc, kk > 0 ? 1u : 0u);
        }
        umma_commit(&s_full[t * 2 + (j & 1)]);
      };
      auto issue_pv = [&](int t, int j) {
        constexpr uint32_t idesc = make_idesc_bf16_major(128, HD, 0, 1);  // N = head dims of O; B = V is MN-major
        const uint32_t va = smem_u32(sV + (j % kFa2Stages) * C::kKvBytes);
        const uint32_t d = tmem_base + 256u + (uint32_t)(t * 128), pt = tmem_base + (uint32_t)(t * 128 + (j & 1) * 64);
        const int ksteps = keys_padded(t, j) >> 4;
        for (int kk = 0; kk < ksteps; ++kk) {
          const uint64_t db = make_sw128_mnmajor_desc(va + kk * 2048, kKv2Half, 1024);
          umma_bf16_ts(d, pt + (uint32_t)(kk * 8), db, idesc, (j > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(&o_done[t]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      for (int t = 0; t < ntile; ++t) issue_s(t, 0);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < nkv_max; ++j) {
        if (j + 1 < nkv_max) {  // S of the NEXT step first: it only needs K[j+1] and the S buffer P.V[j-1] has drained
          const int s1 = (j + 1) % kFa2Stages;
          mbar_wait(&k_full[s1], (uint32_t)((j + 1) / kFa2Stages) & 1u);
          tc_fence_after();
          for (int t = 0; t < ntile; ++t)
            if (j + 1 < nkv[t]) issue_s(t, j + 1);
          umma_commit(&k_empty[s1]);
        }
        mbar_wait(&v_full[j % kFa2Stages], (uint32_t)(j / kFa2Stages) & 1u);
        for (int t = 0; t < ntile; ++t) {
          if (j < nkv[t]) {
            FA_CLK(c0);
            mbar_wait(&p_ready[t * 2 + (j & 1)], (uint32_t)(j >> 1) & 1u);
            FA_CLK(c1);
            pa[0] += c1 - c0;
            tc_fence_after();
            issue_pv(t, j);
          }
        }
        umma_commit(&v_empty[j % kFa2Stages]);
      }
      if (prof_on) p.prof[1 * 8 + 0] = (unsigned long long)pa[0];
    }
  } else if (warp >= 4) {
    // ============================================ softmax / epilogue ============================================
    const int t = (warp - 4) >> 2;
    if (t < ntile) {
      const int quarter = warp & 3;
      const int r = quarter * 32 + lane;
      const int qi = q0 + t * kFaTile + r;
      const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
      const uint32_t tS0 = tmem_base + lane_addr + (uint32_t)(t * 128);
      const uint32_t tO = tmem_base + lane_addr + 256u + (uint32_t)(t * 128);
      const float sl2 = p.scale_log2;
      float m_ref = -INFINITY, l_sum = 0.f;
      for (int j = 0; j < nkv[t]; ++j) {
        FA_CLK(c0);
        mbar_wait(&s_full[t * 2 + (j & 1)], (uint32_t)(j >> 1) & 1u);
        tc_fence_after();
        FA_CLK(c1);
        pa[0] += c1 - c0;
        const uint32_t tS = tS0 + (uint32_t)((j & 1) * 64);
        const int kbase = j * kKv2;
        const int nvalid = min(kKv2, kv_len[t] - kbase);
        const int ncols = (nvalid + 15) & ~15;
        int lim = nvalid - 1;
        if (p.causal) lim = min(lim, qi - kbase);
        const bool masked = !__all_sync(0xffffffffu, lim >= ncols - 1);
        bool rescale = false;
        float alpha = 1.f;
        auto update_ref = [&](float mx) {  // running reference maximum; O is rescaled only when it grew by more than 2^8
          if (j == 0) {
            m_ref = mx;
          } else {
            const bool grow = (mx - m_ref) * sl2 > kFaRescaleLog2;
            if (__any_sync(0xffffffffu, grow)) {
              const float m_use = grow ? mx : m_ref;
              alpha = ex2((m_ref - m_use) * sl2);
              l_sum *= alpha;
              m_ref = m_use;
              rescale = true;
            }
          }
        };
        if (!masked && ncols == kKv2) {
          // ===== fast path: the 64 scores of this row in registers, one pass =====
          uint32_t v[64];
          tmem_ld32(tS, v);
          tmem_ld32(tS + 32u, v + 32);
          tmem_ld_wait();
          FA_CLK(c2);
          pa[1] += c2 - c1;
          float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
          for (int i = 0; i < 64; i += 8) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              mx4[c] = max3(mx4[c], __uint_as_float(v[i + 2 * c]), __uint_as_float(v[i + 2 * c + 1]));
          }
          update_ref(fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])));
          const float neg_m = -m_ref * sl2;
          FA_CLK(c3);
          pa[2] += c3 - c2;
          const uint64_t sl2_2 = f2_pack(sl2, sl2), negm_2 = f2_pack(neg_m, neg_m);
          uint64_t acc2[4] = {0ull, 0ull, 0ull, 0ull};
          uint32_t pk[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const uint64_t x = f2_fma(f2_pack(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sl2_2, negm_2);
            float x0, x1, e0, e1;
            f2_unpack(x, x0, x1);
            if (kFaPolyMask & (1u << (i & 7))) {
              // exp2 on the FMA pipe for 3 of 8 pairs (the MUFU unit, 16 ex2/clk/SM, is the co-bottleneck of the tile):
              // n = round(x) via the 1.5*2^23 trick, 2^f for f = x - n in [-0.5, 0.5] by a degree-4 polynomial (rel. error
              // 4e-5, far below the bf16 rounding of P), then n is added into the exponent field.
              const uint64_t xc = f2_pack(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
              const uint64_t tt = f2_add(xc, f2_pack(12582912.f, 12582912.f));
              const uint64_t rr = f2_add(tt, f2_pack(-12582912.f, -12582912.f));
              const uint64_t ff = f2_fma(rr, f2_pack(-1.f, -1.f), xc);
              uint64_t pp = f2_fma(f2_pack(0.00961812911f, 0.00961812911f), ff, f2_pack(0.0555041087f, 0.0555041087f));
              pp = f2_fma(pp, ff, f2_pack(0.240226507f, 0.240226507f));
              pp = f2_fma(pp, ff, f2_pack(0.693147181f, 0.693147181f));
              pp = f2_fma(pp, ff, f2_pack(1.f, 1.f));
              float t0, t1, p0, p1;
              f2_unpack(tt, t0, t1);
              f2_unpack(pp, p0, p1);
              e0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
              e1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
            } else {
              e0 = ex2(x0);
              e1 = ex2(x1);
            }
            acc2[i & 3] = f2_add(acc2[i & 3], f2_pack(e0, e1));
            pk[i] = pack_bf16(e0, e1);
          }
          FA_CLK(c4);
          pa[3] += c4 - c3;
          tmem_st32(tS, pk);
          float a0, a1;
          f2_unpack(f2_add(f2_add(acc2[0], acc2[1]), f2_add(acc2[2], acc2[3])), a0, a1);
          l_sum += a0 + a1;
        } else {
          // ===== general path (ragged last K/V tile, causal diagonal): 16-column chunks with masking =====
          float mx = -INFINITY;
          for (int c0 = 0; c0 < ncols; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tS + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c0 + i <= lim) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
          update_ref(mx);
          const float neg_m = -m_ref * sl2;
          float l0 = 0.f, l1 = 0.f;
          for (int c0 = 0; c0 < ncols; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tS + (uint32_t)c0, v);
            tmem_ld_wait();
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float e0 = ex2(fmaf(__uint_as_float(v[2 * i]), sl2, neg_m));
              float e1 = ex2(fmaf(__uint_as_float(v[2 * i + 1]), sl2, neg_m));
              if (c0 + 2 * i > lim) e0 = 0.f;
              if (c0 + 2 * i + 1 > lim) e1 = 0.f;
              l0 += e0;
              l1 += e1;
              pk[i] = pack_bf16(e0, e1);
            }
            tmem_st8(tS + (uint32_t)(c0 >> 1), pk);
          }
          l_sum += l0 + l1;
        }
        FA_CLK(c5);
        tmem_st_wait();
        if (j > 0) {
          // the previous P.V of this tile (issued a whole softmax ago) must have landed before O is touched or added to
          mbar_wait(&o_done[t], (uint32_t)(j - 1) & 1u);
          if (rescale) {
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < HD; c0 += 32) {
              uint32_t o[32];
              tmem_ld32(tO + (uint32_t)c0, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st32(tO + (uint32_t)c0, o);
            }
            tmem_st_wait();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[t * 2 + (j & 1)]);
        FA_CLK(c6);
        pa[4] += c6 - c5;
      }
      if (prof_on && lane == 0)
        for (int i = 0; i < 5; ++i) p.prof[warp * 8 + i] = (unsigned long long)pa[i];
      // ---- epilogue
      mbar_wait(&o_done[t], (uint32_t)(nkv[t] - 1) & 1u);
      tc_fence_after();
      const float inv = 1.0f / l_sum;
      bf16* orow = p.out + (long long)(row0 + qi) * p.ldo + head * HD;
      const bool row_ok = qi < len;
#pragma unroll 1
      for (int c0 = 0; c0 < HD; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tO + (uint32_t)c0, v);
        tmem_ld_wait();
        if (!row_ok) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
          o.y = pack_bf16(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
          o.z = pack_bf16(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
          o.w = pack_bf16(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c0 + 8 * i) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

static unsigned long long* g_fa_prof = nullptr;
void set_fa_prof(void* ptr) { g_fa_prof = static_cast<unsigned long long*>(ptr); }
static int g_fa_version = 2;  // 2: 64-key tiles, double-buffered S (default); 1: 128-key tiles (kept for A/B)
void set_fa_version(int v) { g_fa_version = v; }

// host launcher (called by omc_attention_fwd in attention.cu): full 128-row query tiles of every sequence
int launch_fa_sm100(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* out,
                    long long ldo, const int32_t* cu, int num_seqs, int max_seqlen, long long total_rows, int Hq, int Hkv,
                    int head_dim, int causal, float scale_log2, cudaStream_t stream) {
  if ((reinterpret_cast<uintptr_t>(out) & 15) || (ldo % 8) != 0)
    return set_error(OMC_ERR_ALIGN, "omc_attention_fwd: output must be 16-byte aligned with a row stride multiple of 8");
  if (head_dim != 128 && !(head_dim == 64 && g_fa_version == 2))
    return set_error(OMC_ERR_SHAPE, "omc_attention_fwd: head_dim must be 128, or 64 on the default (64-key tile) kernel");
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_tmap_2d(&tmQ, q, total_rows, (long long)Hq * head_dim, ldq, kFaTile);
  if (rc) return rc;
  const int kv_box = g_fa_version == 2 ? kKv2 : kFaTile;
  rc = make_tmap_2d(&tmK, k, total_rows, (long long)Hkv * head_dim, ldk, kv_box);
  if (rc) return rc;
  rc = make_tmap_2d(&tmV, v, total_rows, (long long)Hkv * head_dim, ldv, kv_box);
  if (rc) return rc;
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[cur_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(fa_fwd_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFaSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fa_fwd_sm100_v2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fa2Cfg<128>::kSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(fa_fwd_sm100_v2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fa2Cfg<64>::kSmem);
    if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  FaParams p;
  p.cu = cu; p.out = static_cast<bf16*>(out); p.ldo = ldo; p.Hq = Hq; p.Hkv = Hkv; p.causal = causal;
  p.scale_log2 = scale_log2;
  p.prof = g_fa_prof;
  dim3 grid(((max_seqlen + kFaTile - 1) / kFaTile + 1) / 2, Hq, num_seqs);
  if (head_dim == 64) fa_fwd_sm100_v2_kernel<64><<<grid, kFaThreads, Fa2Cfg<64>::kSmem, stream>>>(tmQ, tmK, tmV, p);
  else if (g_fa_version == 2) fa_fwd_sm100_v2_kernel<128><<<grid, kFaThreads, Fa2Cfg<128>::kSmem, stream>>>(tmQ, tmK, tmV, p);
  else fa_fwd_sm100_kernel<<<grid, kFaThreads, kFaSmem, stream>>>(tmQ, tmK, tmV, p);
  return check_launch("fa_fwd_sm100");
}

}  // namespace omc
