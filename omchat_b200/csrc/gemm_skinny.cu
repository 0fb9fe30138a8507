// Skinny bf16 GEMM for sm_100a:  Y[M,N] = X[M,K] * W[N,K]^T with M <= 64 (decode batches 5..64, HBM-bound weight streaming).
//
// The 128-row tcgen05 tile of gemm_sm100.cu would spend 3/4 of every X tile on zero padding and, worse, has only N/256
// tiles to spread over 148 SMs (o_proj / down_proj of Qwen2-7B: 14 CTAs pulling 135 MB). Here the operands are SWAPPED:
//   D[n, m] = sum_k W[n, k] X[m, k]      A = W tile (128 weight rows x 64 k, all useful bytes), B = X (M padded to 16/32/64)
// so one CTA streams a [128 x Krange] slab of W through an 8-stage TMA ring at full width and the accumulator is a
// 128-lane x Mpad-column TMEM tile (lane = output feature n, column = batch row m). Small N (few 128-row tiles) is split
// along K over the CTAs of a THREAD-BLOCK CLUSTER (<= 8): every CTA parks its fp32 partial tile in its own shared memory
// (the drained TMA ring), the cluster barrier publishes it, and the leader CTA pulls the peers' tiles through distributed
// shared memory, sums them in rank order and applies the epilogue. (The first version met in a global workspace with
// red.global.add + an atomic ticket + a re-zeroing pass: ~8 us of dependent L2 round trips per kernel, which made the
// 25-33 MB o_proj / qkv GEMMs of a batch-32 decode step take 20-30 us instead of the ~5 us their bytes need.)
// Epilogues as in gemm_sm100.cu: bias, erf-GELU, scale*x+residual, SwiGLU on interleaved gate/up rows (adjacent LANES here),
// fp32 output. Call sites: Qwen2 q/k/v/o/gate/up/down/lm_head of a batched decode step (transformers modeling_qwen2.py:46-48,
// 219-221,245,470-472) when the batch is too large for the GEMV kernels.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "omc_internal.h"
#include "ptx.cuh"

namespace omc {

int make_tmap_2d(CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld, int box_rows);

struct SkinnyParams {
  int M, N, K, splitk, num_kb;
  __nv_bfloat16* out;
  float* out_f32;
  long long ldo;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* scale;
  const __nv_bfloat16* res;
  long long ldr;
  int epi;
};

constexpr int kSkThreads = 192;
constexpr int kSkBK = 64;
constexpr int kSkWBytes = 128 * kSkBK * 2;  // 16 KB
// Ring depth 8 (160 KB at M <= 32). Measured: 11 stages (220 KB) change nothing, and neither did replacing the global
// split-K reduction by the cluster one - at batch 32 the kernels are bound by their fixed costs (launch, TMEM / barrier
// set-up, first-load latency, epilogue: ~15-20 us for the 25-33 MB o_proj / qkv GEMMs) and by DRAM efficiency: a
// [128 rows x 128 B] TMA box touches 128 DRAM pages, 5.0-5.4 TB/s on gate/up and lm_head where the persistent decode
// kernel's 14 KB contiguous stages reach 7.4. The cure is one persistent kernel per step.
__host__ __device__ constexpr int sk_stages(int mpad) { return mpad > 0 ? 8 : 8; }

__device__ __forceinline__ float sk_gelu(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float sk_silu(float x) { return x / (1.0f + __expf(-x)); }

template <int MPAD>
__global__ void __launch_bounds__(kSkThreads, 1)
gemm_skinny_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, const SkinnyParams p) {
  constexpr int kXBytes = MPAD * kSkBK * 2;
  constexpr int kStageBytes = kSkWBytes + kXBytes;
  constexpr int kSkStages = sk_stages(MPAD);
  constexpr int kTmemCols = MPAD < 32 ? 32 : MPAD;
  extern __shared__ uint8_t sk_smem_raw[];
  const uint32_t raw_addr = smem_u32(sk_smem_raw);
  uint8_t* smem = sk_smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kSkStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kSkStages;
  uint64_t* tfull_bar = empty_bar + kSkStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  int* s_last = reinterpret_cast<int*>(tmem_slot + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = blockIdx.x / p.splitk, ks = blockIdx.x - nt * p.splitk;
  const int kb0 = (int)((long long)p.num_kb * ks / p.splitk), kb1 = (int)((long long)p.num_kb * (ks + 1) / p.splitk);

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int s = 0; s < kSkStages; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(tfull_bar, 1);
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
    if (elect_one()) {
      uint32_t it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const uint32_t s = it % kSkStages, ph = (it / kSkStages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        uint8_t* sw = smem + s * kStageBytes;
        mbar_arrive_expect_tx(&full_bar[s], (uint32_t)kStageBytes);
        tma_load_2d(sw, &tmW, &full_bar[s], kb * kSkBK, nt * 128, kEvictFirst);   // weights: streamed once
        tma_load_2d(sw + kSkWBytes, &tmX, &full_bar[s], kb * kSkBK, 0, kEvictLast);  // activations: re-read by every CTA
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(128, MPAD);
      uint32_t it = 0;
      for (int kb = kb0; kb < kb1; ++kb, ++it) {
        const uint32_t s = it % kSkStages, ph = (it / kSkStages) & 1u;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sw = smem_u32(smem + s * kStageBytes);
        const uint64_t da = make_sw128_kmajor_desc(sw), db = make_sw128_kmajor_desc(sw + kSkWBytes);
#pragma unroll
        for (int k = 0; k < kSkBK / 16; ++k)
          umma_bf16<1>(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (it | (uint32_t)k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[s]);
      }
      umma_commit(tfull_bar);
    }
  }
  // ===================== epilogue warps (2..5): lane = output feature n, registers = the M batch rows =====================
  const bool epi_warp = warp >= 2;
  const int quarter = warp & 3;
  const int et = (warp - 2) * 32 + lane;  // 0..127 in the epilogue warps
  const int n = nt * 128 + quarter * 32 + lane;
  const bool n_ok = epi_warp && n < p.N;
  float acc[MPAD];
  // residual of the rows this thread will write (EPI_RES, leader CTA only): fetched NOW, while the weights stream. In the
  // epilogue loop a load of res[m+1] may not be hoisted above the store of out[m] (the decoder calls this in place,
  // out == res), so 32 dependent L2 round trips - ~20 us - used to sit at the end of every o_proj / down_proj kernel.
  float resv[MPAD];
  if (epi_warp && p.epi == EPI_RES && ks == 0) {
#pragma unroll
    for (int m = 0; m < MPAD; ++m)
      resv[m] = (m < p.M && n < p.N) ? __bfloat162float(p.res[(long long)m * p.ldr + n]) : 0.f;
  }
  if (epi_warp) {
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
#pragma unroll
    for (int c = 0; c < MPAD; c += 16) {
      uint32_t v[16];
      tmem_ld16(t_addr + (uint32_t)c, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[c + i] = __uint_as_float(v[i]);
    }
  }
  bool do_epilogue = epi_warp;
  if (p.splitk > 1) {
    // split-K inside a cluster: partial tile -> own shared memory [m][128] (the ring is drained: tfull_bar fired after the
    // last MMA read it), cluster barrier, the leader (K slice 0) pulls the peers' tiles over DSMEM in rank order
    float* mine = reinterpret_cast<float*>(smem);
    if (epi_warp && ks != 0) {
#pragma unroll
      for (int m = 0; m < MPAD; ++m)
        if (m < p.M) mine[m * 128 + et] = acc[m];
    }
    __syncwarp();
    cluster_sync_all();  // every thread of every CTA of the cluster (partials visible cluster-wide)
    if (epi_warp && ks == 0) {
      const uint32_t my_addr = smem_u32(mine) + (uint32_t)et * 4u;
      for (int r = 1; r < p.splitk; ++r) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(my_addr), "r"(r));
#pragma unroll
        for (int m = 0; m < MPAD; ++m)
          if (m < p.M) {
            float v;
            asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote + (uint32_t)m * 512u) : "memory");
            acc[m] += v;
          }
      }
    }
    // second phase: the peers may only exit (and give up their shared memory) once the leader has read them; the leader
    // arrives as soon as its reads are done and waits at the end of the kernel
    __syncwarp();
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    do_epilogue = epi_warp && ks == 0;
  }
  if (do_epilogue) {
    const float bias_n = (p.bias != nullptr && n_ok) ? __bfloat162float(p.bias[n]) : 0.f;
    const float scale_n = (p.scale != nullptr && n_ok) ? __bfloat162float(p.scale[n]) : 1.f;
#pragma unroll
    for (int m = 0; m < MPAD; ++m) {
      if (m >= p.M) break;  // uniform
      float v = acc[m] + bias_n;
      if (p.epi == EPI_SWIGLU) {
        const float up = __shfl_down_sync(0xffffffffu, v, 1);  // rows 2i (gate), 2i+1 (up) sit in adjacent lanes
        if (n_ok && (lane & 1) == 0) p.out[(long long)m * p.ldo + (n >> 1)] = __float2bfloat16(sk_silu(v) * up);
        continue;
      }
      if (p.epi == EPI_GELU) v = sk_gelu(v);
      else if (p.epi == EPI_RES && n_ok) v = resv[m] + scale_n * v;
      if (n_ok) {
        if (p.out_f32 != nullptr) p.out_f32[(long long)m * p.ldo + n] = v;
        else p.out[(long long)m * p.ldo + n] = __float2bfloat16(v);
      }
    }
  }
  if (p.splitk > 1) {
    __syncwarp();
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<1>(tmem_base, kTmemCols);
  }
}

// How many clusters of `splitk` CTAs of this kernel the device can hold at once (a cluster lives inside one GPC, so a
// cluster size that does not divide the GPC's SM count strands SMs): the K split is only worth it if all tiles run in
// one wave. Cached per (MPAD, splitk); 0 = unknown / not launchable.
template <int MPAD>
static int max_clusters(int splitk) {
  static int cache_dev[kMaxDevices][9];  // 0 = not probed yet, else clusters + 1
  if (splitk < 1 || splitk > 8) return 0;
  int* cache = cache_dev[cur_device()];
  if (cache[splitk] > 0) return cache[splitk] - 1;
  constexpr int smem = sk_stages(MPAD) * (kSkWBytes + MPAD * kSkBK * 2) + 2048;
  cudaFuncSetAttribute(gemm_skinny_kernel<MPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(splitk * 64);
  cfg.blockDim = dim3(kSkThreads);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = (unsigned)splitk;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_skinny_kernel<MPAD>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  cache[splitk] = n + 1;
  return n;
}

template <int MPAD>
static int launch_skinny(const CUtensorMap& tmW, const CUtensorMap& tmX, const SkinnyParams& p, int ctas, cudaStream_t st) {
  constexpr int smem = sk_stages(MPAD) * (kSkWBytes + MPAD * kSkBK * 2) + 2048;
  static bool attr_set_dev[kMaxDevices] = {};
  bool& attr_set = attr_set_dev[cur_device()];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_skinny_kernel<MPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(kSkThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;  // the K slices of one 128-row tile form a cluster (blockIdx = nt * splitk + ks)
  attrs[0].val.clusterDim.x = (unsigned)p.splitk;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_skinny_kernel<MPAD>, tmW, tmX, p);
  if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
  return check_launch("gemm_skinny");
}

}  // namespace omc

using namespace omc;

extern "C" long long omc_gemm_skinny_workspace_bytes(int max_n) {
  if (max_n <= 0) return -1;
  // fp32 partials [64][max_n] + one ticket per 128-row tile
  return 64LL * max_n * 4 + ((max_n + 127) / 128) * 4LL + 256;
}

extern "C" int omc_gemm_skinny_bf16(const void* X, long long ldx, const void* W, long long ldw, void* out, long long ldo,
                                    int M, int N, int K, const void* bias, const void* scale, const void* res,
                                    long long ldr, int epi, int out_is_f32, void* workspace, long long workspace_bytes,
                                    void* stream) {
  if (M <= 0 || N <= 0 || K <= 0) return set_error(OMC_ERR_SHAPE, "omc_gemm_skinny_bf16: empty problem");
  if (M > 64) return set_error(OMC_ERR_SHAPE, "omc_gemm_skinny_bf16: M must be <= 64 (use omc_gemm_bf16)");
  if (N % 8 != 0 || K % 8 != 0) return set_error(OMC_ERR_SHAPE, "omc_gemm_skinny_bf16: N and K must be multiples of 8");
  if (epi < EPI_NONE || epi > EPI_SWIGLU) return set_error(OMC_ERR_ARG, "omc_gemm_skinny_bf16: unknown epilogue");
  if (epi == EPI_RES && res == nullptr) return set_error(OMC_ERR_ARG, "omc_gemm_skinny_bf16: EPI_RES needs a residual");
  if (out_is_f32 && epi != EPI_NONE) return set_error(OMC_ERR_ARG, "omc_gemm_skinny_bf16: fp32 output only with EPI_NONE");
  if (epi == EPI_SWIGLU && (bias != nullptr || N % 2 != 0))
    return set_error(OMC_ERR_ARG, "omc_gemm_skinny_bf16: SwiGLU epilogue takes no bias and an even N");
  const int n_tiles = (N + 127) / 128, num_kb = (K + kSkBK - 1) / kSkBK;
  const int mpad = M <= 16 ? 16 : (M <= 32 ? 32 : 64);
  int splitk = 1;
  if (n_tiles < num_sms()) {
    splitk = num_sms() / n_tiles;
    if (splitk > 8) splitk = 8;
    if (splitk > num_kb / 4) splitk = num_kb / 4;
    if (splitk < 1) splitk = 1;
    // largest split whose clusters all fit the device in one wave
    while (splitk > 1) {
      const int fit = mpad == 16 ? max_clusters<16>(splitk) : mpad == 32 ? max_clusters<32>(splitk) : max_clusters<64>(splitk);
      if (fit >= n_tiles) break;
      --splitk;
    }
  }
  (void)workspace;        // (the split-K partials used to meet in a global workspace; they now stay inside the cluster)
  (void)workspace_bytes;
  CUtensorMap tmW, tmX;
  int rc = make_tmap_2d(&tmW, W, N, K, ldw, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&tmX, X, M, K, ldx, mpad);
  if (rc) return rc;
  SkinnyParams p{};
  p.M = M; p.N = N; p.K = K; p.splitk = splitk; p.num_kb = num_kb;
  p.out = out_is_f32 ? nullptr : static_cast<__nv_bfloat16*>(out);
  p.out_f32 = out_is_f32 ? static_cast<float*>(out) : nullptr;
  p.ldo = ldo;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.scale = static_cast<const __nv_bfloat16*>(scale);
  p.res = static_cast<const __nv_bfloat16*>(res);
  p.ldr = ldr;
  p.epi = epi;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int ctas = n_tiles * splitk;
  if (mpad == 16) return launch_skinny<16>(tmW, tmX, p, ctas, st);
  if (mpad == 32) return launch_skinny<32>(tmW, tmX, p, ctas, st);
  return launch_skinny<64>(tmW, tmX, p, ctas, st);
}
