// Weight-streaming bf16 GEMM for batched decode steps (M <= 64 rows), sm_100a:  Y[M,N] = epi(X[M,K] * W[N,K]^T).
//
// Successor of gemm_skinny.cu for the decoder's batched step (batches 5..64; Qwen2 q/k/v, o, gate/up, down, lm_head:
// transformers modeling_qwen2.py:46-48,219-221,245,258-263,470-472). Same tcgen05 formulation - operands SWAPPED,
// D[n, m] = sum_k W[n, k] X[m, k], A = a 128-row weight tile (every byte useful), B = X padded to 16/32/64 rows,
// accumulator = 128 TMEM lanes x Mpad columns - but built around what bounded the skinny kernel at batch 32
// (profiles/r01_launches_decode_b32_session6.csv: 0.57 of the HBM roofline):
//
//  * PACKED WEIGHTS. A [128 x 64] TMA box over a row-major matrix is 128 separate 128-byte segments (128 DRAM pages): gate/up
//    streamed at 5.1 TB/s where contiguous stages reach 7.4. Weights are immutable, so they are re-laid once at load time
//    (omc_pack_weight) as [n_tile][k_block] tiles of 128 x 64 bf16 = 16 KB, each stored as the very shared-memory image the
//    MMA wants (K-major, 128-byte swizzle: 16-byte chunk c of row r at r*128 + ((c ^ (r & 7)) << 4)). A stage is then ONE
//    16 KB cp.async.bulk of contiguous memory, and a CTA's whole share of a matrix is one contiguous byte range.
//  * STREAM-K. One CTA per SM, each owning a contiguous range of the n_tiles x num_kb k-blocks of the problem: with at
//    least as many tiles as SMs the k-blocks are dealt out evenly whatever the tile count (296 tiles of gate/up, 1188 of
//    lm_head, 37 of a TP shard: no wave quantisation, at most one cut tile per CTA); with fewer tiles (28 of o_proj /
//    down_proj, 36 of q|k|v) every tile is cut into the same number of equal parts. A cut tile is summed by its owner (the
//    CTA holding k-block 0) from the fp32 partials the others park in a global workspace slot (one per CTA; flag set with
//    release, all flags polled in parallel with acquire, partials added in CTA order, flags cleared by the owner), so
//    results are bit-identical from launch to launch.
//  * PROGRAMMATIC DEPENDENT LAUNCH. Every launch carries cudaLaunchAttributeProgrammaticStreamSerialization and triggers its
//    dependents at once. A CTA needs ~100 KB of shared memory (5 stages of 20 KB cover the HBM latency-bandwidth product of
//    one SM several times over), so the NEXT kernel's CTA moves into the free half of every SM while this kernel still
//    runs, sets up its barriers / TMEM, fills its ring with weights (immutable: no dependency) and only then executes
//    griddepcontrol.wait before touching activations. The ~15 us of launch, set-up and first-load latency each of the ~200
//    kernels of a step used to pay is hidden under its predecessor, and 12 MB of the next matrix is already on chip when
//    the predecessor retires.
//  * RMSNORM FOLDED INTO THE GEMMS (Qwen2RMSNorm feeding q/k/v, gate/up and lm_head). The norm weight is folded into the
//    packed matrix (W'[n,k] = W[n,k] * g[k]), the GEMM runs on the raw residual stream and the epilogue multiplies row m
//    by rstd[m] = rsqrt(sum_k h[m,k]^2 / K + eps). The sum of squares comes from the epilogue that PRODUCED h (EPI_RES of
//    o_proj / down_proj: per 128-column tile partials, summed in tile order by the consumer), so the two row passes per
//    layer disappear. (The reference rounds the normalised row to 16 bit before it multiplies by g; here the rounding
//    happens in W' instead - covered by the stated tolerance, see tests/test_kernels_gpu.py::test_gemm_stream.)
//  * TENSOR-PARALLEL ALL-REDUCE INSIDE THE EPILOGUE (row-parallel o_proj / down_proj under tp_size 2..8, one process per
//    GPU). The owner CTA of an output tile pushes its partial tile straight into every peer's exchange buffer over NVLink
//    (buffers mapped with CUDA IPC: omc_peer_alloc / omc_peer_open) as FLAG-IN-DATA packets - 8 bytes = two bf16 values + a
//    32-bit generation tag, the unit NVLink delivers atomically - then polls the packets the peers write into ITS buffer until
//    they carry this use's tag, and sums the tp_size partials in rank order in fp32 (its own one rounded to bf16 like the
//    ones that travelled), so every GPU ends up with the same bits of the new residual stream; the residual add and the sums
//    of squares for the next folded RMSNorm happen in the same epilogue, and no NCCL kernel (nor a stand-alone add / norm
//    pass) sits between two GEMMs of a decode step. No fence, no separate flag: the first version (fp32 tile + release.sys
//    flag) spent 5-9 us per exchange waiting for 16 KB of peer stores to be acknowledged before the flag could go out.
//    The tag is a per-(channel, tile) use counter every rank keeps in its own buffer (all ranks run the same launch
//    sequence, so the counters agree). One buffer per op kind ("channel": o_proj, down_proj): a producer can only come back
//    to a slot after it has consumed, through the other channel, data its peer sent AFTER reading that slot.
// Epilogues: bias, erf-GELU, +residual, SwiGLU on interleaved gate/up rows (adjacent lanes), fp32 output.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "omc_internal.h"
#include "ptx.cuh"

namespace omc {

int make_tmap_2d(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows);

constexpr int kStThreads = 192;
constexpr int kStBK = 64;
constexpr int kStWBytes = 128 * kStBK * 2;  // one packed tile = one stage of weights: 16 KB
constexpr int kStSlotFloats = 64 * 128;     // workspace slot of one CTA: fp32 partial tile [m][128 lanes]
// ring depth chosen so that TWO CTAs (this kernel's and the next kernel's) fit one SM: ring + fp32 staging tile [mpad][128]
__host__ __device__ constexpr int st_stages(int mpad) { return mpad <= 16 ? 5 : (mpad <= 32 ? 4 : 3); }
__host__ __device__ constexpr int st_smem(int mpad) {
  return st_stages(mpad) * (kStWBytes + mpad * 128) + mpad * 128 * 4 + 1024 + 2048;
}

struct StreamParams {
  int M, N, K, n_tiles, num_kb;
  const uint8_t* wp;  // packed weights
  __nv_bfloat16* out;
  float* out_f32;
  long long ldo;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* res;
  long long ldr;
  int epi;
  int splitk;           // > 0: every tile is cut into splitk equal parts (grid = n_tiles * splitk); 0: even split of all k-blocks
  const float* ssq_in;  // [parts][64] partial sums of squares of the input rows (nullptr: no row scale)
  int ssq_parts;
  float inv_norm_dim, eps;
  float* ssq_out;  // [n_tiles][64] partial sums of squares of the bf16 outputs (EPI_RES), or nullptr
  float* ws;       // [grid][64*128] fp32 partial tiles
  unsigned int* flags;  // [grid]
  unsigned long long* prof;  // optional [grid][8] %globaltimer stamps of this launch (tools/prof_stream.py), else nullptr
  // tensor-parallel exchange (tp_size > 1): xbuf[r] = rank r's exchange buffer as mapped into THIS process
  int* dbg;  // optional pinned host record for the spin-wait watchdog
  int tp_rank, tp_size, tp_channel, tp_fence_all;
  uint8_t* xbuf[8];
};

// exchange buffer layout: [channel 2][src rank 8][tile 64] payload slots of 64 x 128 fp32, then the flags [2][8][64] u32
constexpr int kXTiles = 64;
constexpr size_t kXPayload = (size_t)kStSlotFloats * 4;
constexpr size_t kXFlagsOffset = 2ull * 8 * kXTiles * kXPayload;
constexpr size_t kXBytes = kXFlagsOffset + 2ull * 8 * kXTiles * 4;
__host__ __device__ inline size_t x_slot(int channel, int src, int tile) { return ((size_t)(channel * 8 + src) * kXTiles + tile); }

__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long st_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define ST_STAMP(i) do { if (p.prof != nullptr) p.prof[(size_t)blockIdx.x * 16 + (i)] = st_now(); } while (0)

__device__ __forceinline__ void st_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(kEvictFirst)
      : "memory");
}
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// first k-block (global index tile * num_kb + kb) of CTA j's range
__device__ __forceinline__ long long st_range_start(const StreamParams& p, long long total, unsigned int j, unsigned int grid) {
  if (p.splitk > 0) {
    const unsigned int tile = j / (unsigned int)p.splitk, part = j - tile * (unsigned int)p.splitk;
    return (long long)tile * p.num_kb + (long long)p.num_kb * part / p.splitk;
  }
  return total * j / grid;
}
// Spin-wait guard: a partner that never shows up (a rank that died, a protocol bug) must not hang the GPU - after ~10 s of
// polling the kernel traps, which surfaces as a CUDA error in the host process instead of a dead box.
struct StWatchdog {
  unsigned long long t0 = 0;
  unsigned int n = 0;
  // dbg: optional pinned host record {kind, a, b, c} written before the trap (readable after the context died)
  __device__ __forceinline__ void tick(int* dbg, int kind, int a, int b, int c) {
    if ((++n & 0xFFFFu) == 0u) {
      const unsigned long long now = st_now();
      if (t0 == 0) t0 = now;
      else if (now - t0 > 10000000000ull) {
        if (dbg != nullptr && atomicCAS(dbg, 0, kind) == 0) {
          dbg[1] = a; dbg[2] = b; dbg[3] = c; dbg[4] = (int)blockIdx.x; dbg[5] = (int)gridDim.x;
          __threadfence_system();
        }
        __trap();
      }
    }
  }
};
__device__ __forceinline__ float st_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float st_silu(float x) { return x / (1.0f + __expf(-x)); }

// Sum of the K parts 1..splitk-1 of one (m, 8-feature) chunk, read from the staging tiles of the other CTAs of the cluster
// over distributed shared memory. All loads of a group of four peers are issued before the first add (an add behind
// every load would serialise the ~0.3 us DSMEM round trips: 5 peers x 2 loads x 4 chunks used to cost ~10 us per epilogue).
__device__ __forceinline__ void st_dsmem_add8(uint32_t local_addr, int splitk, float4& a, float4& b4) {
#pragma unroll
  for (int base = 1; base < 8; base += 4) {
    if (base >= splitk) break;
    float4 x[4], y[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = base + q;
      if (r < splitk) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(r));
        asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(x[q].x), "=f"(x[q].y), "=f"(x[q].z), "=f"(x[q].w) : "r"(remote));
        asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(y[q].x), "=f"(y[q].y), "=f"(y[q].z), "=f"(y[q].w) : "r"(remote + 16u));
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (base + q < splitk) {  // rank order: bit-identical from launch to launch
        a.x += x[q].x; a.y += x[q].y; a.z += x[q].z; a.w += x[q].w;
        b4.x += y[q].x; b4.y += y[q].y; b4.z += y[q].z; b4.w += y[q].w;
      }
    }
  }
}

// Sum over ALL splitk staging tiles of the cluster (own one included, through its cluster address) in rank order.
__device__ __forceinline__ void st_dsmem_sum8(uint32_t local_addr, int splitk, float4& a, float4& b4) {
  a = make_float4(0.f, 0.f, 0.f, 0.f);
  b4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int base = 0; base < 8; base += 4) {
    if (base >= splitk) break;
    float4 x[4], y[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int r = base + q;
      if (r < splitk) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(r));
        asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(x[q].x), "=f"(x[q].y), "=f"(x[q].z), "=f"(x[q].w) : "r"(remote));
        asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(y[q].x), "=f"(y[q].y), "=f"(y[q].z), "=f"(y[q].w) : "r"(remote + 16u));
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (base + q < splitk) {
        a.x += x[q].x; a.y += x[q].y; a.z += x[q].z; a.w += x[q].w;
        b4.x += y[q].x; b4.y += y[q].y; b4.z += y[q].z; b4.w += y[q].w;
      }
    }
  }
}

template <int MPAD>
__global__ void __launch_bounds__(kStThreads, 2)
gemm_stream_kernel(const __grid_constant__ CUtensorMap tmX, const StreamParams p) {
  constexpr int kXBytes = MPAD * 128;
  constexpr int kStage = kStWBytes + kXBytes;
  constexpr int S = st_stages(MPAD);
  constexpr int kTmemCols = 2 * MPAD < 32 ? 32 : 2 * MPAD;
  extern __shared__ uint8_t st_smem_raw[];
  const uint32_t raw_addr = smem_u32(st_smem_raw);
  uint8_t* smem = st_smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  float* stage_f = reinterpret_cast<float*>(smem + S * kStage);  // [MPAD][128] fp32 epilogue staging tile
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S * kStage + MPAD * 128 * 4);
  uint64_t* empty_bar = full_bar + S;
  uint64_t* tfull_bar = empty_bar + S;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_rstd = reinterpret_cast<float*>(tmem_slot + 2);  // [64]
  const bool cluster_mode = p.splitk > 1;  // the splitk K parts of a tile = the CTAs of one thread-block cluster
  // cluster mode on one GPU: phase 2 of the epilogue is SHARED by the CTAs of the cluster - CTA r finalises the rows
  // m = r, r + splitk, ... of the tile (sum of all parts over DSMEM, epilogue math, stores, sums of squares), so the
  // hand-off to the next kernel waits for 1 / splitk of the work. (Under tensor parallelism the tile's owner does it all: it
  // is the one talking to the peers.)
  const bool dist = cluster_mode && p.tp_size <= 1;
  const int row_step = dist ? p.splitk : 1;
  const int row_base = dist ? (int)(blockIdx.x % (unsigned int)p.splitk) : 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total = (long long)p.n_tiles * p.num_kb;
  const long long g0 = st_range_start(p, total, blockIdx.x, gridDim.x);
  const long long g1 = blockIdx.x + 1 == gridDim.x ? total : st_range_start(p, total, blockIdx.x + 1, gridDim.x);

  griddep_launch();  // dependents may be scheduled as soon as SM slots free up; they wait for our completion themselves
  if (threadIdx.x == 0) ST_STAMP(0);
  if (warp == 0 && elect_one()) tma_prefetch_desc(&tmX);
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < S; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int b = 0; b < 2; ++b) {
        mbar_init(&tfull_bar[b], 1);
        mbar_init(&tempty_bar[b], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<1>(tmem_slot, kTmemCols);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== producer: weights run ahead of the dependency, activations wait for it =====================
    if (elect_one()) {
      const uint8_t* wsrc = p.wp + (size_t)g0 * kStWBytes;
      const int n_items = (int)(g1 - g0);
      const int pre = n_items < S ? n_items : S;
      for (int i = 0; i < pre; ++i) {
        mbar_arrive_expect_tx(&full_bar[i], (uint32_t)kStage);
        st_bulk_g2s(smem + i * kStage, wsrc + (size_t)i * kStWBytes, kStWBytes, &full_bar[i]);
      }
      ST_STAMP(1);
      griddep_wait();
      ST_STAMP(2);
      for (int i = 0; i < pre; ++i) {
        const int kb = (int)((g0 + i) % p.num_kb);
        tma_load_2d(smem + i * kStage + kStWBytes, &tmX, &full_bar[i], kb * kStBK, 0, kEvictLast);
      }
      for (int i = pre; i < n_items; ++i) {
        const uint32_t s = (uint32_t)i % S, ph = ((uint32_t)i / S) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* st = smem + s * kStage;
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)kStage);
        st_bulk_g2s(st, wsrc + (size_t)i * kStWBytes, kStWBytes, &full_bar[s]);
        const int kb = (int)((g0 + i) % p.num_kb);
        tma_load_2d(st + kStWBytes, &tmX, &full_bar[s], kb * kStBK, 0, kEvictLast);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: one segment (= this CTA's k-blocks of one tile) per TMEM buffer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, MPAD);
      uint32_t it = 0, seg = 0;
      long long g = g0;
      while (g < g1) {
        const long long tile = g / p.num_kb;
        const long long tile_end = (tile + 1) * p.num_kb;
        const long long seg_end = tile_end < g1 ? tile_end : g1;
        const uint32_t buf = seg & 1u;
        if (seg >= 2) mbar_wait(&tempty_bar[buf], ((seg >> 1) - 1u) & 1u);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + buf * (uint32_t)MPAD;
        bool first = true;
        for (; g < seg_end; ++g, ++it) {
          const uint32_t s = it % S, ph = (it / S) & 1u;
          mbar_wait(&full_bar[s], ph);
          if (it == 0) ST_STAMP(3);
          tc_fence_after();
          const uint32_t sw = smem_u32(smem + s * kStage);
          const uint64_t da = make_sw128_kmajor_desc(sw), db = make_sw128_kmajor_desc(sw + kStWBytes);
#pragma unroll
          for (int k = 0; k < kStBK / 16; ++k)
            umma_bf16<1>(d_addr, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (first && k == 0) ? 0u : 1u);
          first = false;
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tfull_bar[buf]);
        ++seg;
      }
      ST_STAMP(4);
    }
  } else {
    // ===================== epilogue warps 2..5 =====================
    // Phase 1: thread = output feature n of the tile (TMEM lane), registers = the M rows -> fp32 staging tile [m][128] in
    // shared memory. Phase 2 (tile owner): thread = (row m, 8 consecutive features): 16-byte loads of bias / residual,
    // 16-byte stores, the sums of squares reduced over the 16 lanes that share a row. In cluster mode the owner adds the
    // other K parts of the tile straight out of their CTAs' staging tiles (distributed shared memory), in rank order.
    const int quarter = warp & 3;
    const int et = (warp - 2) * 32 + lane;  // 0..127 among the epilogue threads
    const int tl = quarter * 32 + lane;     // TMEM lane = row of the tile
    griddep_wait();                         // everything below reads or writes activations
    if (p.ssq_in != nullptr) {
      if (et < 64) {
        float s = 0.f;
        if (et < p.M)
          for (int q = 0; q < p.ssq_parts; ++q) s += __ldcg(p.ssq_in + q * 64 + et);
        s_rstd[et] = rsqrtf(s * p.inv_norm_dim + p.eps);
      }
      epi_barrier();
    }
    float* my_slot = p.ws + (size_t)blockIdx.x * kStSlotFloats;
    constexpr int kIters = MPAD * 16 / 128;  // (m, 8-feature chunk) items per thread in phase 2
    uint32_t seg = 0;
    long long g = g0;
    while (g < g1) {
      const long long tile = g / p.num_kb;
      const long long tile_end = (tile + 1) * p.num_kb;
      const long long seg_end = tile_end < g1 ? tile_end : g1;
      const bool owner = (g == tile * p.num_kb);
      const bool whole = owner && seg_end == tile_end;
      const uint32_t buf = seg & 1u;
      // residual chunks of phase 2 fetched NOW, while the weights stream (out may alias res)
      uint4 resq[kIters];
      if ((owner || dist) && p.epi == EPI_RES) {
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
          const int c = i * 128 + et, m = row_base + row_step * (c >> 4), n = (int)tile * 128 + (c & 15) * 8;
          resq[i] = (m < p.M && m < MPAD && n < p.N) ? __ldcg(reinterpret_cast<const uint4*>(p.res + (long long)m * p.ldr + n))
                                         : make_uint4(0, 0, 0, 0);
        }
      }
      float acc[MPAD];
      mbar_wait(&tfull_bar[buf], (seg >> 1) & 1u);
      tc_fence_after();
      {
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + buf * (uint32_t)MPAD;
#pragma unroll
        for (int c = 0; c < MPAD; c += 16) {
          uint32_t v[16];
          tmem_ld16(t_addr + (uint32_t)c, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[c + i] = __uint_as_float(v[i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[buf]);  // the MMA warp may reuse this accumulator
      if (et == 0 && seg_end == g1) ST_STAMP(5);
      if (!cluster_mode && !owner) {
        // stream-K: tail of a tile whose head lives in the previous CTA: park the partial in global memory, publish
#pragma unroll
        for (int m = 0; m < MPAD; ++m)
          if (m < p.M) my_slot[m * 128 + tl] = acc[m];
        __threadfence();
        epi_barrier();
        if (et == 0) st_release_u32(p.flags + blockIdx.x, 1u);
      } else {
        if (!cluster_mode && !whole) {
          // stream-K: head of a tile that continues in the following CTA(s): add their partials in CTA order
          unsigned int nparts = 0;
          for (unsigned int j = blockIdx.x + 1; j < gridDim.x && st_range_start(p, total, j, gridDim.x) < tile_end; ++j) ++nparts;
          for (unsigned int i = (unsigned int)et; i < nparts; i += 128u) {
            StWatchdog wd;
            while (ld_acquire_u32(p.flags + blockIdx.x + 1 + i) == 0u) wd.tick(p.dbg, 1, (int)tile, (int)i, 0);
          }
          epi_barrier();
          for (unsigned int i = 0; i < nparts; ++i) {
            const float* slot = p.ws + (size_t)(blockIdx.x + 1 + i) * kStSlotFloats;
#pragma unroll
            for (int m = 0; m < MPAD; ++m)
              if (m < p.M) acc[m] += __ldcg(slot + m * 128 + tl);
          }
          epi_barrier();
          // consumed: ready for the next launch (which writes only after its own griddepcontrol.wait)
          for (unsigned int i = (unsigned int)et; i < nparts; i += 128u) p.flags[blockIdx.x + 1 + i] = 0u;
        }
        // ---- phase 1: registers -> staging tile (lane-contiguous: conflict-free)
#pragma unroll
        for (int m = 0; m < MPAD; ++m) stage_f[m * 128 + tl] = acc[m];
        if (cluster_mode) {
          __syncwarp();
          cluster_sync_all();  // every thread of every CTA of the cluster: all K parts of the tile are staged
        } else {
          epi_barrier();
        }
        if (et == 0) ST_STAMP(6);
        unsigned int xtag = 0;
        if (owner && p.tp_size > 1) {
          // ---- tensor-parallel exchange: local sum of the K parts -> bf16 + tag packets into every peer's buffer over NVLink
          const uint32_t stage_addr0 = smem_u32(stage_f);
          const size_t slot_me = x_slot(p.tp_channel, p.tp_rank, (int)tile);
          // this use's tag = the slot's use counter + 1 (kept in this rank's own buffer; equal on all ranks)
          xtag = __ldcg(reinterpret_cast<const unsigned int*>(p.xbuf[p.tp_rank] + kXFlagsOffset) + slot_me) + 1u;
#pragma unroll
          for (int i = 0; i < kIters; ++i) {
            const int c = i * 128 + et, m = c >> 4, n8 = (c & 15) * 8;
            float4 a = *reinterpret_cast<const float4*>(stage_f + m * 128 + n8);
            float4 b4 = *reinterpret_cast<const float4*>(stage_f + m * 128 + n8 + 4);
            if (cluster_mode) st_dsmem_add8(stage_addr0 + (uint32_t)(m * 128 + n8) * 4u, p.splitk, a, b4);
            const uint32_t p0 = pack_bf16(a.x, a.y), p1 = pack_bf16(a.z, a.w), p2 = pack_bf16(b4.x, b4.y), p3 = pack_bf16(b4.z, b4.w);
            // own contribution = the same bf16-rounded values the peers receive (each chunk is touched by this thread only)
            const float2 r0 = unpack_bf16(p0), r1 = unpack_bf16(p1), r2 = unpack_bf16(p2), r3 = unpack_bf16(p3);
            *reinterpret_cast<float4*>(stage_f + m * 128 + n8) = make_float4(r0.x, r0.y, r1.x, r1.y);
            *reinterpret_cast<float4*>(stage_f + m * 128 + n8 + 4) = make_float4(r2.x, r2.y, r3.x, r3.y);
            if (m < p.M) {
              for (int r = 0; r < p.tp_size; ++r) {
                if (r == p.tp_rank) continue;
                uint4* dst = reinterpret_cast<uint4*>(p.xbuf[r] + slot_me * kXPayload + (size_t)(m * 128 + n8) * 4);
                asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(p0), "r"(xtag), "r"(p1), "r"(xtag)
                             : "memory");
                asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 1), "r"(p2), "r"(xtag), "r"(p3), "r"(xtag)
                             : "memory");
              }
            }
          }
          if (et == 0) ST_STAMP(8);
        }
        if (cluster_mode && p.tp_size > 1) {
          // the K parts of this GPU are summed: the other CTAs of the cluster may retire while the exchange is in flight
          __syncwarp();
          cluster_sync_all();
        }
        if (owner || dist) {
          // ---- phase 2
          const uint32_t stage_addr = smem_u32(stage_f);
#pragma unroll
          for (int i = 0; i < kIters; ++i) {
            const int c = i * 128 + et, n8 = (c & 15) * 8;
            // a warp covers two rows (lanes 0-15 / 16-31): leave the loop only when BOTH are past the tile (warp-uniform: the
            // sums of squares below shuffle with a full mask); a lane whose own row is past it computes on row 0 and stores nothing
            if (row_base + row_step * ((i * 128 + (et & ~31)) >> 4) >= MPAD) break;
            const int m_raw = row_base + row_step * (c >> 4);
            const int m = m_raw < MPAD ? m_raw : 0;
            const int n = (int)tile * 128 + n8;
            const bool ok = m_raw < p.M && n < p.N;  // N % 8 == 0: a chunk is entirely inside or outside
            float v[8];
            if (dist) {
              float4 a, b4;
              st_dsmem_sum8(stage_addr + (uint32_t)(m * 128 + n8) * 4u, p.splitk, a, b4);
              v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b4.x; v[5] = b4.y; v[6] = b4.z; v[7] = b4.w;
            } else {
              const float4 a = *reinterpret_cast<const float4*>(stage_f + m * 128 + n8);
              const float4 b4 = *reinterpret_cast<const float4*>(stage_f + m * 128 + n8 + 4);
              v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b4.x; v[5] = b4.y; v[6] = b4.z; v[7] = b4.w;
            }
            if (p.tp_size > 1) {
              // sum of the tp_size partial tiles in RANK order (own one from the staging tile): identical bits on every GPU
              float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
              for (int r = 0; r < p.tp_size; ++r) {
                float4 a, b4;
                if (r == p.tp_rank) {
                  a = make_float4(v[0], v[1], v[2], v[3]);
                  b4 = make_float4(v[4], v[5], v[6], v[7]);
                } else if (m_raw < p.M) {
                  // a peer's chunk = 4 packets {bf16 x 2, tag}: poll until all four carry this use's tag
                  const uint4* src = reinterpret_cast<const uint4*>(p.xbuf[p.tp_rank] + x_slot(p.tp_channel, r, (int)tile) * kXPayload +
                                                                    (size_t)(m * 128 + n8) * 4);
                  uint4 q0, q1;
                  StWatchdog wd;
                  while (true) {
                    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w) : "l"(src) : "memory");
                    asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "l"(src + 1) : "memory");
                    if (q0.y == xtag && q0.w == xtag && q1.y == xtag && q1.w == xtag) break;
                    wd.tick(p.dbg, 2 + 16 * p.tp_rank, (int)tile, r, p.tp_channel);
                  }
                  const float2 r0 = unpack_bf16(q0.x), r1 = unpack_bf16(q0.z), r2 = unpack_bf16(q1.x), r3 = unpack_bf16(q1.z);
                  a = make_float4(r0.x, r0.y, r1.x, r1.y);
                  b4 = make_float4(r2.x, r2.y, r3.x, r3.y);
                } else {
                  a = make_float4(0.f, 0.f, 0.f, 0.f);
                  b4 = a;
                }
                s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w; s[4] += b4.x; s[5] += b4.y; s[6] += b4.z; s[7] += b4.w;
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = s[j];
            }
            if (p.ssq_in != nullptr) {
              const float r = s_rstd[m < 64 ? m : 0];
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] *= r;
            }
            if (p.bias != nullptr && ok) {
              const uint4 bq = __ldg(reinterpret_cast<const uint4*>(p.bias + n));
              const float2 b0 = unpack_bf16(bq.x), b1 = unpack_bf16(bq.y), b2 = unpack_bf16(bq.z), b3 = unpack_bf16(bq.w);
              v[0] += b0.x; v[1] += b0.y; v[2] += b1.x; v[3] += b1.y; v[4] += b2.x; v[5] += b2.y; v[6] += b3.x; v[7] += b3.y;
            }
            if (p.epi == EPI_SWIGLU) {
              // features 2i (gate), 2i+1 (up) are adjacent: 8 features -> 4 outputs
              if (ok) {
                const uint32_t o0 = pack_bf16(st_silu(v[0]) * v[1], st_silu(v[2]) * v[3]);
                const uint32_t o1 = pack_bf16(st_silu(v[4]) * v[5], st_silu(v[6]) * v[7]);
                *reinterpret_cast<uint2*>(p.out + (long long)m * p.ldo + (n >> 1)) = make_uint2(o0, o1);
              }
              continue;
            }
            if (p.epi == EPI_GELU) {
#pragma unroll
              for (int j = 0; j < 8; ++j) v[j] = st_gelu(v[j]);
            } else if (p.epi == EPI_RES) {
              const float2 r0 = unpack_bf16(resq[i].x), r1 = unpack_bf16(resq[i].y), r2 = unpack_bf16(resq[i].z),
                           r3 = unpack_bf16(resq[i].w);
              v[0] += r0.x; v[1] += r0.y; v[2] += r1.x; v[3] += r1.y; v[4] += r2.x; v[5] += r2.y; v[6] += r3.x; v[7] += r3.y;
            }
            if (p.out_f32 != nullptr) {
              if (ok) {
                float* dst = p.out_f32 + (long long)m * p.ldo + n;
                *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4], v[5], v[6], v[7]);
              }
            } else {
              const uint4 o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
              if (ok) *reinterpret_cast<uint4*>(p.out + (long long)m * p.ldo + n) = o;
              if (p.ssq_out != nullptr) {
                // sum of squares of the bf16 values actually written, over the 16 lanes (= 128 features) sharing row m
                const float2 q0 = unpack_bf16(o.x), q1 = unpack_bf16(o.y), q2 = unpack_bf16(o.z), q3 = unpack_bf16(o.w);
                float sq = ok ? (q0.x * q0.x + q0.y * q0.y + q1.x * q1.x + q1.y * q1.y + q2.x * q2.x + q2.y * q2.y +
                                 q3.x * q3.x + q3.y * q3.y) : 0.f;
#pragma unroll
                for (int d = 8; d > 0; d >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, d);
                if ((lane & 15) == 0 && m_raw < p.M) p.ssq_out[(long long)tile * 64 + m] = sq;
              }
            }
          }
        }
        if (owner && p.tp_size > 1) {
          if (et == 0) ST_STAMP(9);
          // every thread holds its copy of the tag: advance this slot's use counter for the next launch on this channel
          epi_barrier();
          if (et == 0)
            reinterpret_cast<unsigned int*>(p.xbuf[p.tp_rank] + kXFlagsOffset)[x_slot(p.tp_channel, p.tp_rank, (int)tile)] = xtag;
        }
        if (cluster_mode && p.tp_size <= 1) {
          // the other parts may only retire (and give up their staging tiles) once the owner has read them
          __syncwarp();
          cluster_sync_all();
        } else if (!cluster_mode) {
          epi_barrier();  // the staging tile is reused by the next segment
        }
      }
      g = seg_end;
      ++seg;
    }
    if (et == 0) ST_STAMP(7);
  }
  if (cluster_mode && warp < 2) {
    // warps 0 / 1 take part in the two cluster barriers of the epilogue (barrier.cluster counts every thread)
    __syncwarp();
    cluster_sync_all();
    cluster_sync_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------ packing
// One thread per 16-byte chunk of the packed image.
__global__ void pack_weight_kernel(const __nv_bfloat16* __restrict__ W, long long ldw, int N, int K, int num_kb,
                                   const __nv_bfloat16* __restrict__ col_scale, uint8_t* __restrict__ packed, long long chunks) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < chunks; i += (long long)gridDim.x * blockDim.x) {
    const long long tile_id = i >> 10;  // 1024 chunks per 16 KB tile
    const int within = (int)(i & 1023);
    const int r = within >> 3, c = within & 7;
    const long long nt = tile_id / num_kb;
    const int kb = (int)(tile_id - nt * num_kb);
    const long long row = nt * 128 + r;
    const int k0 = kb * kStBK + c * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (row < N && k0 < K) {  // K % 8 == 0: a chunk is entirely inside or outside
      v = *reinterpret_cast<const uint4*>(W + row * ldw + k0);
      if (col_scale != nullptr) {
        const uint4 sc = *reinterpret_cast<const uint4*>(col_scale + k0);
        const uint32_t* a = reinterpret_cast<const uint32_t*>(&v);
        const uint32_t* b = reinterpret_cast<const uint32_t*>(&sc);
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 x = unpack_bf16(a[q]), y = unpack_bf16(b[q]);
          o[q] = pack_bf16(x.x * y.x, x.y * y.y);
        }
        v = make_uint4(o[0], o[1], o[2], o[3]);
      }
    }
    *reinterpret_cast<uint4*>(packed + tile_id * kStWBytes + r * 128 + ((c ^ (r & 7)) << 4)) = v;
  }
}

// Residual rows -> per-row sum of squares (one part), for producers that are not a gemm_stream EPI_RES epilogue
// (embedding rows of the first layer, the all-reduced residual stream under tensor parallelism).
__global__ void __launch_bounds__(256) row_ssq_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, int C,
                                                      float* __restrict__ ssq, int parts) {
  griddep_launch();
  griddep_wait();
  const int m = blockIdx.x;
  const __nv_bfloat16* row = x + (long long)m * ldx;
  float s = 0.f;
  for (int i = threadIdx.x * 8; i < C; i += blockDim.x * 8) {
    const uint4 v = *reinterpret_cast<const uint4*>(row + i);
    const uint32_t* a = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 f = unpack_bf16(a[q]);
      s += f.x * f.x + f.y * f.y;
    }
  }
  s = warp_sum(s);
  __shared__ float sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += sh[w];
    ssq[m] = t;
    for (int q = 1; q < parts; ++q) ssq[q * 64 + m] = 0.f;
  }
}

// How many clusters of `splitk` CTAs the device holds at once (2 CTAs per SM by shared memory). Cached per device.
template <int MPAD>
static int st_max_clusters(int splitk) {
  static int cache_dev[kMaxDevices][9];  // 0 = not probed yet, else clusters + 1
  if (splitk < 1 || splitk > 8) return 0;
  int* cache = cache_dev[cur_device()];
  if (cache[splitk] > 0) return cache[splitk] - 1;
  constexpr int smem = st_smem(MPAD);
  cudaFuncSetAttribute(gemm_stream_kernel<MPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(splitk * 64);
  cfg.blockDim = dim3(kStThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = (unsigned)splitk;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_stream_kernel<MPAD>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  cache[splitk] = n + 1;
  return n;
}

template <int MPAD>
static int launch_stream(const CUtensorMap& tmX, const StreamParams& p, int grid, bool pdl, cudaStream_t st) {
  constexpr int smem = st_smem(MPAD);
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[cur_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_stream_kernel<MPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kStThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  int na = 0;
  if (pdl) {
    attrs[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (p.splitk > 1) {  // the K parts of one tile form a cluster (blockIdx = tile * splitk + part)
    attrs[na].id = cudaLaunchAttributeClusterDimension;
    attrs[na].val.clusterDim.x = (unsigned)p.splitk;
    attrs[na].val.clusterDim.y = 1;
    attrs[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attrs;
  cfg.numAttrs = na;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_stream_kernel<MPAD>, tmX, p);
  if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
  return check_launch("gemm_stream");
}

}  // namespace omc

using namespace omc;

extern "C" long long omc_packed_weight_bytes(int N, int K) {
  if (N <= 0 || K <= 0) return -1;
  return (long long)((N + 127) / 128) * ((K + kStBK - 1) / kStBK) * kStWBytes;
}

extern "C" int omc_pack_weight(const void* W, long long ldw, int N, int K, const void* col_scale, void* packed, void* stream) {
  if (N <= 0 || K <= 0 || W == nullptr || packed == nullptr) return set_error(OMC_ERR_ARG, "omc_pack_weight: bad argument");
  if (K % 8 != 0 || ldw % 8 != 0 || (reinterpret_cast<uintptr_t>(W) & 15) || (reinterpret_cast<uintptr_t>(packed) & 1023))
    return set_error(OMC_ERR_ALIGN, "omc_pack_weight: K, ldw multiples of 8, W 16-byte and packed 1024-byte aligned");
  const int num_kb = (K + kStBK - 1) / kStBK;
  const long long chunks = omc_packed_weight_bytes(N, K) / 16;
  long long blocks = (chunks + 255) / 256;
  if (blocks > (long long)num_sms() * 16) blocks = (long long)num_sms() * 16;
  pack_weight_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)W, ldw, N, K, num_kb,
                                                                     (const __nv_bfloat16*)col_scale, (uint8_t*)packed, chunks);
  return check_launch("pack_weight");
}

// profiling hook (tools/prof_stream.py): consecutive launches stamp consecutive [296][8] blocks of `buf`
static int* g_dbg = nullptr;
static unsigned long long* g_prof_buf = nullptr;
static int g_prof_max = 0, g_prof_next = 0;

static int st_max_grid() {
  static int ctas_per_sm = -1;
  if (ctas_per_sm < 0) {
    const char* e = getenv("OMCHAT_B200_STREAM_CTAS_PER_SM");  // 1 (default): the next kernel's CTA co-resides; 2 for A/B
    ctas_per_sm = (e != nullptr && e[0] == '2') ? 2 : 1;
  }
  return ctas_per_sm * num_sms();
}

extern "C" int omc_gemm_stream_set_debug(void* pinned_host_record) {
  g_dbg = static_cast<int*>(pinned_host_record);
  return OMC_OK;
}

extern "C" int omc_gemm_stream_set_prof(void* buf, int max_launches) {
  g_prof_buf = static_cast<unsigned long long*>(buf);
  g_prof_max = buf != nullptr ? max_launches : 0;
  g_prof_next = 0;
  return OMC_OK;
}

extern "C" long long omc_gemm_stream_xchg_bytes(void) { return (long long)kXBytes; }

extern "C" long long omc_gemm_stream_workspace_bytes(void) {
  const long long grid = 2LL * num_sms();
  return grid * kStSlotFloats * 4 + grid * 4 + 256;
}

extern "C" int omc_row_ssq(const void* x, long long ldx, int rows, int C, float* ssq, int parts, int pdl, void* stream) {
  if (rows <= 0) return OMC_OK;
  if (rows > 64 || C % 8 != 0 || ldx % 8 != 0 || parts < 1) return set_error(OMC_ERR_SHAPE, "omc_row_ssq: rows <= 64, C % 8 == 0");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(rows);
  cfg.blockDim = dim3(256);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, row_ssq_kernel, (const __nv_bfloat16*)x, ldx, C, ssq, parts);
  if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
  return check_launch("row_ssq");
}

extern "C" int omc_gemm_stream(const void* X, long long ldx, int M, const void* Wp, int N, int K, void* out, long long ldo,
                               int out_is_f32, const void* bias, const void* res, long long ldr, int epi,
                               const float* ssq_in, int ssq_parts, int norm_dim, float eps, float* ssq_out, void* workspace,
                               int pdl, const omc_tp_xchg* tp, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(OMC_ERR_SHAPE, "omc_gemm_stream: empty problem");
  if (M > 64) return set_error(OMC_ERR_SHAPE, "omc_gemm_stream: M must be <= 64 (use omc_gemm_bf16)");
  if (K % 8 != 0 || N % 8 != 0 || ldo % 8 != 0 || (res != nullptr && ldr % 8 != 0))
    return set_error(OMC_ERR_SHAPE, "omc_gemm_stream: K, N and the row strides must be multiples of 8");
  if (epi < EPI_NONE || epi > EPI_SWIGLU) return set_error(OMC_ERR_ARG, "omc_gemm_stream: unknown epilogue");
  if (epi == EPI_RES && res == nullptr) return set_error(OMC_ERR_ARG, "omc_gemm_stream: EPI_RES needs a residual");
  if (out_is_f32 && epi != EPI_NONE) return set_error(OMC_ERR_ARG, "omc_gemm_stream: fp32 output only with EPI_NONE");
  if (epi == EPI_SWIGLU && (bias != nullptr || N % 2 != 0))
    return set_error(OMC_ERR_ARG, "omc_gemm_stream: SwiGLU epilogue takes no bias and an even N");
  if (ssq_out != nullptr && (epi == EPI_SWIGLU || out_is_f32))
    return set_error(OMC_ERR_ARG, "omc_gemm_stream: sum-of-squares output needs a bf16 row output");
  if (ssq_in != nullptr && (ssq_parts < 1 || norm_dim < 1)) return set_error(OMC_ERR_ARG, "omc_gemm_stream: bad ssq_in description");
  if (workspace == nullptr || (reinterpret_cast<uintptr_t>(Wp) & 1023))
    return set_error(OMC_ERR_ARG, "omc_gemm_stream: workspace missing or packed weights not 1024-byte aligned");
  const int n_tiles = (N + 127) / 128, num_kb = (K + kStBK - 1) / kStBK;
  const int mpad = M <= 16 ? 16 : (M <= 32 ? 32 : 64);
  const long long total = (long long)n_tiles * num_kb;
  long long grid = st_max_grid();
  int splitk = 0;
  if (n_tiles < grid) {
    // few tiles: cut every tile into the same number of equal parts, each at least 4 k-blocks long
    splitk = (int)(grid / n_tiles);
    if (splitk > 8) splitk = 8;  // portable cluster size
    if (splitk > num_kb / 4) splitk = num_kb / 4;
    if (splitk < 1) splitk = 1;
    // largest split whose clusters all fit the device at once (a cluster lives inside one GPC; two CTAs fit one SM, so
    // this counts the slots a co-resident predecessor still holds: its clusters then start as those CTAs retire)
    while (splitk > 1) {
      const int fit = mpad == 16 ? st_max_clusters<16>(splitk) : mpad == 32 ? st_max_clusters<32>(splitk) : st_max_clusters<64>(splitk);
      if (fit >= n_tiles) break;
      --splitk;
    }
    grid = (long long)n_tiles * splitk;
  }
  if (grid > total) grid = total;
  CUtensorMap tmX;
  int rc = make_tmap_2d(&tmX, X, M, K, ldx, mpad);
  if (rc) return rc;
  StreamParams p{};
  p.M = M; p.N = N; p.K = K; p.n_tiles = n_tiles; p.num_kb = num_kb;
  p.wp = static_cast<const uint8_t*>(Wp);
  p.out = out_is_f32 ? nullptr : static_cast<__nv_bfloat16*>(out);
  p.out_f32 = out_is_f32 ? static_cast<float*>(out) : nullptr;
  p.ldo = ldo;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.res = static_cast<const __nv_bfloat16*>(res);
  p.ldr = ldr;
  p.epi = epi;
  p.splitk = splitk;
  p.tp_rank = 0; p.tp_size = 1; p.tp_channel = 0;
  if (tp != nullptr && tp->size > 1) {
    if (tp->size > 8 || tp->rank < 0 || tp->rank >= tp->size || tp->channel < 0 || tp->channel > 1)
      return set_error(OMC_ERR_ARG, "omc_gemm_stream: bad tensor-parallel exchange description");
    if (epi != EPI_RES && epi != EPI_NONE) return set_error(OMC_ERR_ARG, "omc_gemm_stream: the exchange needs EPI_RES / EPI_NONE");
    if (out_is_f32 || n_tiles > kXTiles) return set_error(OMC_ERR_SHAPE, "omc_gemm_stream: exchange needs bf16 output, N <= 8192");
    if (n_tiles >= grid) return set_error(OMC_ERR_SHAPE, "omc_gemm_stream: exchange needs fewer tiles than SMs");
    p.tp_rank = tp->rank; p.tp_size = tp->size; p.tp_channel = tp->channel; p.tp_fence_all = tp->reserved & 1;
    for (int r = 0; r < tp->size; ++r) {
      if (tp->bufs[r] == nullptr) return set_error(OMC_ERR_ARG, "omc_gemm_stream: null exchange buffer");
      p.xbuf[r] = static_cast<uint8_t*>(tp->bufs[r]);
    }
  }
  p.ssq_in = ssq_in; p.ssq_parts = ssq_parts; p.inv_norm_dim = norm_dim > 0 ? 1.0f / (float)norm_dim : 0.f; p.eps = eps;
  p.ssq_out = ssq_out;
  const long long max_grid = 2LL * num_sms();
  p.ws = static_cast<float*>(workspace);
  p.flags = reinterpret_cast<unsigned int*>(static_cast<float*>(workspace) + max_grid * kStSlotFloats);
  p.dbg = g_dbg;
  p.prof = (g_prof_buf != nullptr && g_prof_next < g_prof_max) ? g_prof_buf + (size_t)(g_prof_next++) * max_grid * 16 : nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mpad == 16) return launch_stream<16>(tmX, p, (int)grid, pdl != 0, st);
  if (mpad == 32) return launch_stream<32>(tmX, p, (int)grid, pdl != 0, st);
  return launch_stream<64>(tmX, p, (int)grid, pdl != 0, st);
}
