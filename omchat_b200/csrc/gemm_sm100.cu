// Dense bf16 GEMM for sm_100a:  Y[M,N] = X[M,K] * W[N,K]^T  (+ fused epilogue), fp32 accumulation in TMEM.
//
// This one kernel serves every nn.Linear on the OmChat hot path (reference call sites:
// intern_vit_6b/modeling_intern_vit.py:124,136,183-190 qkv/proj/fc1/fc2; multimodal_projector/builder.py:57-61;
// transformers qwen2 q/k/v/o/gate/up/down; patch-embed conv as im2col GEMM modeling_intern_vit.py:73-75,92).
//
// Structure (persistent, warp-specialised, one CTA or one CTA pair per SM / SM pair):
//   warp 0      : TMA producer  - cp.async.bulk.tensor 128B-swizzled K-major tiles of X and W into a smem ring
//   warp 1      : MMA issuer    - tcgen05.mma (cta_group::1 M=128, or cta_group::2 M=256) into a double-buffered
//                                 TMEM accumulator; tcgen05.commit releases smem slots / publishes accumulators
//   warps 2..5  : epilogue      - tcgen05.ld TMEM->registers, bias / GELU / layer-scale+residual / SwiGLU, bf16 stores
// Three mbarrier pipelines: smem full/empty, TMEM full/empty, and a static persistent tile schedule.
#include <functional>
#include <mutex>
#include <unordered_map>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "omc_internal.h"
#include "ptx.cuh"

namespace omc {

struct GemmParams {
  int M, N, K;
  __nv_bfloat16* out;
  long long ldo;
  const __nv_bfloat16* bias;   // [N] or null
  const __nv_bfloat16* scale;  // [N] (layer-scale) or null
  const __nv_bfloat16* res;    // [M, ldr] or null
  long long ldr;
  float* out_f32;  // if non-null, EPI_NONE writes fp32 here instead of bf16
  int epi;
  int num_m_tiles, num_n_tiles;
  // folded RMSNorm (the norm weight lives in W; see omc_gemm_bf16_norm): row scale rstd[m] = rsqrt(sum_parts / norm_dim + eps)
  const float* ssq_in;   // [ssq_in_parts][ssq_in_ld] partial sums of squares of X's rows, or null
  long long ssq_in_ld;
  int ssq_in_parts;
  float inv_norm_dim, eps;
  float* ssq_out;        // [num_n_tiles][ssq_out_ld]: sums of squares of the bf16 rows written, per N tile, or null
  long long ssq_out_ld;
  // grouped (mixture-of-experts) mode, CG = 1 only: X's rows are sorted by expert with every expert's segment padded to whole
  // 128-row tiles; grp_tile[mt] = expert whose weights [N, K] (rows grp * N ... of the stacked W) tile mt multiplies, < 0 = skip
  const int32_t* grp_tile;
};

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kGemmThreads = 192;
constexpr int kSmemBudget = 227 * 1024;

template <int BN, int CG>
struct GemmCfg {
  static constexpr int kBLoadRows = BN / CG;
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = kBLoadRows * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStagesRaw = (kSmemBudget - 2048) / kStageBytes;
  static constexpr int kStages = kStagesRaw > 8 ? 8 : kStagesRaw;
  static constexpr int kSmemBytes = kStages * kStageBytes + 2048;  // +1024 align slack, +1024 barriers
  static constexpr int kAccStride = 256;                           // TMEM columns between the two accumulators
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

__device__ __forceinline__ void tile_coords(int t, int num_m, int num_n, int& m, int& n) {
  constexpr int G = 8;  // m-tiles per group: keeps G X-tiles and a window of W-tiles hot in L2
  int per_group = G * num_n;
  int g = t / per_group;
  int r = t - g * per_group;
  int gsize = num_m - g * G;
  gsize = gsize > G ? G : gsize;
  n = r / gsize;
  m = g * G + (r - n * gsize);
}

template <int BN, int CG>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmParams p) {
  using Cfg = GemmCfg<BN, CG>;
  constexpr int STAGES = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzled tiles need 1024-byte alignment; identical in both CTAs of a pair.
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* tiles = smem;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = (cta_rank == 0);
  const int num_kb = (p.K + kBK - 1) / kBK;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int cluster_id = (CG == 2) ? (blockIdx.x >> 1) : blockIdx.x;
  const int num_clusters = (CG == 2) ? (gridDim.x >> 1) : gridDim.x;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(&tfull_bar[a], 1);
        mbar_init(&tempty_bar[a], 4 * CG);  // one arrival per epilogue warp (of each CTA in the pair)
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<CG>(tmem_slot, 512);
    tmem_relinquish<CG>();
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Programmatic dependent launch (the grouped GEMMs of a mixture-of-experts decode step are launched that way): everything
  // above - barrier init, the TMEM allocation, the descriptor prefetch - overlapped the predecessor's tail; operands, the
  // tile -> expert table and the epilogue inputs are only touched from here on. No-ops in an ordinary launch.
  pdl_launch();
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      uint32_t it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int mt, nt;
        tile_coords(t, p.num_m_tiles, p.num_n_tiles, mt, nt);
        const int grp = p.grp_tile != nullptr ? __ldg(p.grp_tile + mt) : 0;
        if (grp < 0) continue;
        const int row_a = (mt * CG + (int)cta_rank) * kBM;
        const int row_b = grp * p.N + nt * BN + (int)cta_rank * Cfg::kBLoadRows;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
          mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* sa = tiles + s * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          if constexpr (CG == 1) {
            mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes);
            tma_load_2d(sa, &tmA, &full_bar[s], kb * kBK, row_a, kEvictNormal);
            tma_load_2d(sb, &tmB, &full_bar[s], kb * kBK, row_b, kEvictNormal);
          } else {
            if (leader) mbar_arrive_expect_tx(&full_bar[s], Cfg::kStageBytes * 2);
            tma_load_2d_2sm(sa, &tmA, &full_bar[s], kb * kBK, row_a, kEvictNormal);
            tma_load_2d_2sm(sb, &tmB, &full_bar[s], kb * kBK, row_b, kEvictNormal);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(kBM * CG, BN);
      uint32_t it = 0, lt = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        if (p.grp_tile != nullptr) {
          int mt, nt;
          tile_coords(t, p.num_m_tiles, p.num_n_tiles, mt, nt);
          if (__ldg(p.grp_tile + mt) < 0) continue;
        }
        const uint32_t acc = lt & 1u, acc_ph = (lt >> 1) & 1u;
        ++lt;
        mbar_wait_cluster(&tempty_bar[acc], acc_ph ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * Cfg::kAccStride;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
          mbar_wait_cluster(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(tiles + s * Cfg::kStageBytes);
          const uint64_t da = make_sw128_kmajor_desc(sa);
          const uint64_t db = make_sw128_kmajor_desc(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 elements (32 B) along K inside the 128B swizzle atom: +2 in 16-byte units
            umma_bf16<CG>(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (CG == 1) umma_commit(&empty_bar[s]); else umma_commit_2sm(&empty_bar[s], 0x3);
        }
        if constexpr (CG == 1) umma_commit(&tfull_bar[acc]); else umma_commit_2sm(&tfull_bar[acc], 0x3);
      }
    }
  } else {
    // ===================== epilogue warps (TMEM lane quarter = warp % 4) =====================
    const int quarter = warp & 3;
    const int lane = lane_id();
    uint32_t lt = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      int mt, nt;
      tile_coords(t, p.num_m_tiles, p.num_n_tiles, mt, nt);
      if (p.grp_tile != nullptr && __ldg(p.grp_tile + mt) < 0) continue;
      const uint32_t acc = lt & 1u, acc_ph = (lt >> 1) & 1u;
      ++lt;
      const long long row = (long long)(mt * CG + (int)cta_rank) * kBM + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      // folded RMSNorm: this row's scale, fetched while the MMAs of the tile are still running
      float rstd = 1.0f;
      if (p.ssq_in != nullptr) {
        float ss = 0.f;
        if (row_ok)
          for (int q = 0; q < p.ssq_in_parts; ++q) ss += __ldcg(p.ssq_in + q * p.ssq_in_ld + row);
        rstd = rsqrtf(ss * p.inv_norm_dim + p.eps);
      }
      float ssq_acc = 0.f;
      mbar_wait_cluster(&tfull_bar[acc], acc_ph);
      tc_fence_after();
      const uint32_t t_addr = tmem_base + acc * Cfg::kAccStride + ((uint32_t)(quarter * 32) << 16);

      if (p.epi == EPI_SWIGLU) {
        // W rows alternate gate_i, up_i: accumulator columns (2j, 2j+1) -> output column j
        const int n_out = p.N >> 1;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld32(t_addr + c, v);
          tmem_ld_wait();
          const int col = (nt * BN + c) >> 1;
          if (col >= n_out) break;
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float a0 = silu(__uint_as_float(v[4 * j]) * rstd) * (__uint_as_float(v[4 * j + 1]) * rstd);
            float a1 = silu(__uint_as_float(v[4 * j + 2]) * rstd) * (__uint_as_float(v[4 * j + 3]) * rstd);
            o[j] = pack_bf16(a0, a1);
          }
          if (row_ok) {
            __nv_bfloat16* dst = p.out + row * p.ldo + col;
            if (col + 8 <= n_out) *reinterpret_cast<uint4*>(dst) = make_uint4(o[0], o[1], o[2], o[3]);
            if (col + 16 <= n_out) *reinterpret_cast<uint4*>(dst + 8) = make_uint4(o[4], o[5], o[6], o[7]);
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld32(t_addr + c, v);
          tmem_ld_wait();
          const int col = nt * BN + c;
          if (col >= p.N) break;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * rstd;
          if (p.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (col + 8 * j + 8 <= p.N) {
                uint4 b = *reinterpret_cast<const uint4*>(p.bias + col + 8 * j);
                float2 b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y), b2 = unpack_bf16(b.z), b3 = unpack_bf16(b.w);
                f[8 * j + 0] += b0.x; f[8 * j + 1] += b0.y; f[8 * j + 2] += b1.x; f[8 * j + 3] += b1.y;
                f[8 * j + 4] += b2.x; f[8 * j + 5] += b2.y; f[8 * j + 6] += b3.x; f[8 * j + 7] += b3.y;
              }
            }
          }
          if (p.epi == EPI_GELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
          } else if (p.epi == EPI_RES) {
            if (p.scale != nullptr) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (col + 8 * j + 8 <= p.N) {
                  uint4 b = *reinterpret_cast<const uint4*>(p.scale + col + 8 * j);
                  float2 b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y), b2 = unpack_bf16(b.z), b3 = unpack_bf16(b.w);
                  f[8 * j + 0] *= b0.x; f[8 * j + 1] *= b0.y; f[8 * j + 2] *= b1.x; f[8 * j + 3] *= b1.y;
                  f[8 * j + 4] *= b2.x; f[8 * j + 5] *= b2.y; f[8 * j + 6] *= b3.x; f[8 * j + 7] *= b3.y;
                }
              }
            }
            if (row_ok) {
              const __nv_bfloat16* rp = p.res + row * p.ldr + col;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (col + 8 * j + 8 <= p.N) {
                  uint4 b = *reinterpret_cast<const uint4*>(rp + 8 * j);
                  float2 b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y), b2 = unpack_bf16(b.z), b3 = unpack_bf16(b.w);
                  f[8 * j + 0] += b0.x; f[8 * j + 1] += b0.y; f[8 * j + 2] += b1.x; f[8 * j + 3] += b1.y;
                  f[8 * j + 4] += b2.x; f[8 * j + 5] += b2.y; f[8 * j + 6] += b3.x; f[8 * j + 7] += b3.y;
                }
              }
            }
          }
          if (row_ok) {
            if (p.out_f32 != nullptr) {
              float* dst = p.out_f32 + row * p.ldo + col;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (col + 4 * j + 4 <= p.N)
                  *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else {
              __nv_bfloat16* dst = p.out + row * p.ldo + col;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                if (col + 8 * j + 8 <= p.N) {
                  uint4 o = make_uint4(pack_bf16(f[8 * j + 0], f[8 * j + 1]), pack_bf16(f[8 * j + 2], f[8 * j + 3]),
                                       pack_bf16(f[8 * j + 4], f[8 * j + 5]), pack_bf16(f[8 * j + 6], f[8 * j + 7]));
                  *reinterpret_cast<uint4*>(dst + 8 * j) = o;
                  if (p.ssq_out != nullptr) {  // of the bf16 values actually stored
                    const float2 q0 = unpack_bf16(o.x), q1 = unpack_bf16(o.y), q2 = unpack_bf16(o.z), q3 = unpack_bf16(o.w);
                    ssq_acc += q0.x * q0.x + q0.y * q0.y + q1.x * q1.x + q1.y * q1.y + q2.x * q2.x + q2.y * q2.y +
                               q3.x * q3.x + q3.y * q3.y;
                  }
                }
              }
            }
          }
        }
      }
      if (p.ssq_out != nullptr && row_ok) p.ssq_out[(long long)nt * p.ssq_out_ld + row] = ssq_acc;
      // accumulator drained -> hand the TMEM buffer back to the MMA issuer (in the leader CTA)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 1) mbar_arrive(&tempty_bar[acc]); else mbar_arrive_cluster(&tempty_bar[acc], 0);
      }
    }
  }

  // ===================== teardown =====================
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// 2-D bf16 row-major [rows, cols] (leading dim ld elements), box = [box_rows, 64 cols], 128-byte swizzle.
// A tensor map is a pure function of (base, rows, cols, ld, box_rows): encoded once and kept (weights and the activation
// buffers of a steady-state loop come back with the same key every call - a prefill issues ~900 of these).
namespace {
struct TmapKey {
  const void* base;
  long long rows, cols, ld;
  int box_rows, device;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows && device == o.device;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = std::hash<const void*>()(k.base);
    auto mix = [&h](size_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix((size_t)k.rows); mix((size_t)k.cols); mix((size_t)k.ld); mix((size_t)k.box_rows); mix((size_t)k.device);
    return h;
  }
};
std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
constexpr size_t kTmapCacheMax = 16384;
}  // namespace

int make_tmap_2d(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return set_error(OMC_ERR_DRIVER, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15))
    return set_error(OMC_ERR_ALIGN, "GEMM operand must be 16-byte aligned with a leading dim multiple of 8");
  const TmapKey key{base, rows, cols, ld, box_rows, cur_device()};
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *tm = it->second;
      return OMC_OK;
    }
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(OMC_ERR_DRIVER, "cuTensorMapEncodeTiled failed");
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() >= kTmapCacheMax) g_tmap_cache.clear();
    g_tmap_cache.emplace(key, *tm);
  }
  return OMC_OK;
}

static int g_num_sms[kMaxDevices] = {};
int num_sms() {
  const int dev = cur_device();
  if (g_num_sms[dev] == 0) {
    cudaDeviceGetAttribute(&g_num_sms[dev], cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms[dev] <= 0) g_num_sms[dev] = 148;
  }
  return g_num_sms[dev];
}

int moe_pdl_enabled();  // moe.cu

template <int BN, int CG>
static int launch_gemm(const void* A, long long lda, const void* W, long long ldw, GemmParams p, int max_ctas,
                       cudaStream_t stream, int groups = 1) {
  using Cfg = GemmCfg<BN, CG>;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d(&tmA, A, p.M, p.K, lda, kBM);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, W, (long long)p.N * groups, p.K, ldw, Cfg::kBLoadRows);
  if (rc) return rc;
  p.num_m_tiles = (p.M + kBM * CG - 1) / (kBM * CG);
  p.num_n_tiles = (p.N + BN - 1) / BN;
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[cur_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  int sms = max_ctas > 0 ? max_ctas : num_sms();
  int clusters = sms / CG;
  int tiles = p.num_m_tiles * p.num_n_tiles;
  if (clusters > tiles) clusters = tiles;
  if (clusters < 1) clusters = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CG);
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = CG;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = (p.grp_tile != nullptr && moe_pdl_enabled()) ? 2 : 1;  // grouped (expert) GEMMs: programmatic dependents
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_bf16_kernel<BN, CG>, tmA, tmB, p);
  if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
  return OMC_OK;
}

}  // namespace omc

using namespace omc;

static int gemm_impl(const void* X, long long ldx, const void* W, long long ldw, void* out, long long ldo, int M, int N, int K,
                     const void* bias, const void* scale, const void* res, long long ldr, int epi, int out_is_f32,
                     int tile_cfg, const omc_gemm_norm* nf, void* stream);

extern "C" int omc_gemm_bf16(const void* X, long long ldx, const void* W, long long ldw, void* out, long long ldo,
                             int M, int N, int K, const void* bias, const void* scale, const void* res,
                             long long ldr, int epi, int out_is_f32, int tile_cfg, void* stream) {
  return gemm_impl(X, ldx, W, ldw, out, ldo, M, N, K, bias, scale, res, ldr, epi, out_is_f32, tile_cfg, nullptr, stream);
}

extern "C" int omc_gemm_bf16_norm(const void* X, long long ldx, const void* W, long long ldw, void* out, long long ldo,
                                  int M, int N, int K, const void* bias, const void* scale, const void* res,
                                  long long ldr, int epi, int out_is_f32, int tile_cfg, omc_gemm_norm* nf, void* stream) {
  if (nf == nullptr) return set_error(OMC_ERR_ARG, "omc_gemm_bf16_norm: null norm description");
  return gemm_impl(X, ldx, W, ldw, out, ldo, M, N, K, bias, scale, res, ldr, epi, out_is_f32, tile_cfg, nf, stream);
}

// Grouped GEMM for the routed experts of a mixture-of-experts MLP (transformers Qwen2MoeExperts.forward,
// modeling_qwen2_moe.py:307-331): out[r, :] = epi( X[r, :] . W[e(r)]^T ) where the rows of X are sorted by expert, every expert's
// segment starts on a 128-row tile and tile_expert[r / 128] names the expert (< 0: the tile holds no rows, skipped without
// touching the weights). W is the experts' matrices stacked, [n_experts * N, K]. One launch, no host knowledge of the routing.
extern "C" int omc_gemm_bf16_grouped(const void* X, long long ldx, int M_max, const void* W, long long ldw, int n_experts, int N,
                                     int K, const int32_t* tile_expert, int active_tiles_hint, void* out, long long ldo, int epi,
                                     void* stream) {
  if (X == nullptr || W == nullptr || tile_expert == nullptr || out == nullptr)
    return set_error(OMC_ERR_ARG, "omc_gemm_bf16_grouped: null argument");
  if (M_max <= 0 || M_max % kBM != 0 || n_experts <= 0 || N <= 0 || K <= 0 || N % 128 != 0 || K % 8 != 0)
    return set_error(OMC_ERR_SHAPE, "omc_gemm_bf16_grouped: M_max must be whole 128-row tiles, N a multiple of 128, K of 8");
  if (epi != EPI_NONE && epi != EPI_SWIGLU) return set_error(OMC_ERR_ARG, "omc_gemm_bf16_grouped: epilogue NONE or SWIGLU");
  GemmParams p{};
  p.M = M_max; p.N = N; p.K = K;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.ldo = ldo;
  p.epi = epi;
  p.grp_tile = tile_expert;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // 256-wide tiles unless even the worst case (active_tiles_hint 128-row tiles active) leaves SMs idle: decode steps stream the weights of
  // a few experts, and more, narrower CTAs put more of them in flight
  const int m_tiles = (active_tiles_hint > 0 && active_tiles_hint < M_max / kBM) ? active_tiles_hint : M_max / kBM;
  const long long tiles256 = (long long)m_tiles * ((N + 255) / 256);
  if (N % 256 == 0 && tiles256 >= 2LL * num_sms()) return launch_gemm<256, 1>(X, ldx, W, ldw, p, 0, st, n_experts);
  return launch_gemm<128, 1>(X, ldx, W, ldw, p, 0, st, n_experts);
}

static int gemm_impl(const void* X, long long ldx, const void* W, long long ldw, void* out, long long ldo, int M, int N, int K,
                     const void* bias, const void* scale, const void* res, long long ldr, int epi, int out_is_f32,
                     int tile_cfg, const omc_gemm_norm* nf_c, void* stream) {
  omc_gemm_norm* nf = const_cast<omc_gemm_norm*>(nf_c);
  if (M <= 0 || N <= 0 || K <= 0) return set_error(OMC_ERR_SHAPE, "omc_gemm_bf16: empty problem");
  if (N % 8 != 0 || K % 8 != 0) return set_error(OMC_ERR_SHAPE, "omc_gemm_bf16: N and K must be multiples of 8");
  if (epi < EPI_NONE || epi > EPI_SWIGLU) return set_error(OMC_ERR_ARG, "omc_gemm_bf16: unknown epilogue");
  if (epi == EPI_RES && res == nullptr) return set_error(OMC_ERR_ARG, "omc_gemm_bf16: EPI_RES needs a residual");
  if (out_is_f32 && epi != EPI_NONE) return set_error(OMC_ERR_ARG, "omc_gemm_bf16: fp32 output only with EPI_NONE");
  if (epi == EPI_SWIGLU && (bias != nullptr || N % 16 != 0))
    return set_error(OMC_ERR_ARG, "omc_gemm_bf16: SwiGLU epilogue takes no bias and N % 16 == 0");
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.out = out_is_f32 ? nullptr : static_cast<__nv_bfloat16*>(out);
  p.out_f32 = out_is_f32 ? static_cast<float*>(out) : nullptr;
  p.ldo = ldo;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.scale = static_cast<const __nv_bfloat16*>(scale);
  p.res = static_cast<const __nv_bfloat16*>(res);
  p.ldr = ldr;
  p.epi = epi;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int bn = tile_cfg & 0xFFFF, cg = (tile_cfg >> 16) & 0xF, max_ctas = (tile_cfg >> 20) & 0xFFF;
  if (bn == 0) {
    // auto: 2-CTA pairs with the widest N tile that divides N - except that a 256-wide tile with a partly empty last column
    // tile beats the narrower exact fit when the waste is small (M = 8200: N = 9600 1397 vs 1314 TFLOP/s at 1.3 % waste,
    // N = 3200 1239 vs 1220 at 4 %; profiles/r01_gemm_microbench.jsonl)
    cg = 2;
    const int pad256 = (N + 255) / 256 * 256 - N;
    if (N % 256 == 0 || (N >= 2048 && pad256 * 20 <= N)) bn = 256;
    else bn = (N % 160 == 0) ? 160 : (N % 192 == 0) ? 192 : (N % 128 == 0) ? 128 : 256;
  }
  if (cg == 0) cg = 2;
  if (nf != nullptr) {
    if (nf->ssq_in != nullptr) {
      if (nf->ssq_in_parts < 1 || nf->norm_dim < 1 || nf->ssq_in_ld < M)
        return set_error(OMC_ERR_ARG, "omc_gemm_bf16_norm: bad ssq_in description");
      p.ssq_in = nf->ssq_in; p.ssq_in_parts = nf->ssq_in_parts; p.ssq_in_ld = nf->ssq_in_ld;
      p.inv_norm_dim = 1.0f / (float)nf->norm_dim; p.eps = nf->eps;
    }
    if (nf->ssq_out != nullptr) {
      if (out_is_f32 || epi == EPI_SWIGLU || nf->ssq_out_ld < M)
        return set_error(OMC_ERR_ARG, "omc_gemm_bf16_norm: sums of squares need a bf16 row output");
      p.ssq_out = nf->ssq_out; p.ssq_out_ld = nf->ssq_out_ld;
      nf->ssq_out_parts = (N + bn - 1) / bn;  // one partial per N tile of the configuration that runs
      if (nf->ssq_out_parts > nf->ssq_out_max_parts)
        return set_error(OMC_ERR_ARG, "omc_gemm_bf16_norm: ssq_out buffer holds too few parts");
    }
  }
#define OMC_GEMM_CASE(BN_, CG_) \
  if (bn == BN_ && cg == CG_) return launch_gemm<BN_, CG_>(X, ldx, W, ldw, p, max_ctas, st);
  OMC_GEMM_CASE(256, 1)
  OMC_GEMM_CASE(128, 1)
  OMC_GEMM_CASE(256, 2)
  OMC_GEMM_CASE(192, 2)
  OMC_GEMM_CASE(160, 2)
  OMC_GEMM_CASE(128, 2)
#undef OMC_GEMM_CASE
  return set_error(OMC_ERR_ARG, "omc_gemm_bf16: unsupported tile configuration");
}
