// Decode-time linear layers: out[B,N] = epi( norm(x)[B,K] . W[N,K]^T ), B <= 8, HBM-bound weight streaming.
// Each warp owns a group of 4 weight rows and walks K with 128-bit ld.global.nc loads (8 independent 16-byte loads in
// flight per lane), the activation row(s) sit in shared memory (RMS-normalised on the way in when norm_w is given),
// partial dot products are reduced with warp shuffles. Epilogues: bias, residual add, SwiGLU over interleaved rows.
// Reference call sites: transformers models/qwen2/modeling_qwen2.py:46-48 (MLP), :219-221,245 (q/k/v/o),
// :258-263 (RMSNorm fused into the consumer), :470-472 (lm_head).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "omc_internal.h"
#include "ptx.cuh"

namespace omc {

typedef __nv_bfloat16 bf16;
constexpr int kGemvThreads = 256;
constexpr int kGemvWarps = kGemvThreads / 32;
constexpr int kRows = 4;  // weight rows per warp step

struct GemvParams {
  const bf16* x; long long ldx;
  const bf16* W; long long ldw;
  void* out; long long ldo;
  int N, K;
  const bf16* norm_w; float eps;
  const bf16* bias;
  const bf16* res; long long ldr;
  int epi, out_f32;
  int num_groups;  // row groups of kRows
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

__device__ __forceinline__ void fma8(float& acc, uint4 w, const float* xv) {
  float2 a = unpack_bf16(w.x), b = unpack_bf16(w.y), c = unpack_bf16(w.z), d = unpack_bf16(w.w);
  acc = fmaf(a.x, xv[0], acc); acc = fmaf(a.y, xv[1], acc);
  acc = fmaf(b.x, xv[2], acc); acc = fmaf(b.y, xv[3], acc);
  acc = fmaf(c.x, xv[4], acc); acc = fmaf(c.y, xv[5], acc);
  acc = fmaf(d.x, xv[6], acc); acc = fmaf(d.y, xv[7], acc);
}

// weight row index of slot j (0..3) of row-group grp. SwiGLU: W rows alternate gate_i, up_i, so a group of 4 rows holds
// two (gate, up) pairs = output columns 2*grp and 2*grp+1.
__device__ __forceinline__ int group_row(int grp, int j, int) { return grp * kRows + j; }

template <int NB>
__global__ void __launch_bounds__(kGemvThreads) gemv_bf16_kernel(const GemvParams p) {
  extern __shared__ __align__(16) uint8_t gemv_smem[];
  bf16* sx = reinterpret_cast<bf16*>(gemv_smem);  // [NB][K]
  __shared__ float red[kGemvWarps];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.K, nvec = K >> 3;

  // ---- stage (optionally RMS-normalised) activations
  for (int b = 0; b < NB; ++b) {
    const uint4* xr = reinterpret_cast<const uint4*>(p.x + (long long)b * p.ldx);
    uint4* sr = reinterpret_cast<uint4*>(sx + (long long)b * K);
    if (p.norm_w == nullptr) {
      for (int i = tid; i < nvec; i += kGemvThreads) sr[i] = xr[i];
    } else {
      float ss = 0.f;
      for (int i = tid; i < nvec; i += kGemvThreads) {
        uint4 v = xr[i];
        sr[i] = v;
        float2 a = unpack_bf16(v.x), bb = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
        ss += a.x * a.x + a.y * a.y + bb.x * bb.x + bb.y * bb.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
      }
      ss = warp_sum(ss);
      __syncthreads();
      if (lane == 0) red[warp] = ss;
      __syncthreads();
      float tot = 0.f;
#pragma unroll
      for (int i = 0; i < kGemvWarps; ++i) tot += red[i];
      const float rstd = rsqrtf(tot / (float)K + p.eps);
      const uint4* wv = reinterpret_cast<const uint4*>(p.norm_w);
      for (int i = tid; i < nvec; i += kGemvThreads) {
        uint4 v = sr[i], g = wv[i];
        uint32_t xi[4] = {v.x, v.y, v.z, v.w}, gi[4] = {g.x, g.y, g.z, g.w}, oo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float2 a = unpack_bf16(xi[q]), w2 = unpack_bf16(gi[q]);
          float2 n = unpack_bf16(pack_bf16(a.x * rstd, a.y * rstd));
          oo[q] = pack_bf16(n.x * w2.x, n.y * w2.y);
        }
        sr[i] = make_uint4(oo[0], oo[1], oo[2], oo[3]);
      }
    }
  }
  __syncthreads();

  // ---- stream weights: warp-per-row-group, grid-strided
  const int gwarp = blockIdx.x * kGemvWarps + warp, nwarps = gridDim.x * kGemvWarps;
  for (int grp = gwarp; grp < p.num_groups; grp += nwarps) {
    const uint4* wr[kRows];
#pragma unroll
    for (int j = 0; j < kRows; ++j) {
      int row = group_row(grp, j, p.epi);
      if (row >= p.N) row = p.N - 1;
      wr[j] = reinterpret_cast<const uint4*>(p.W + (long long)row * p.ldw);
    }
    float acc[kRows][NB];
#pragma unroll
    for (int j = 0; j < kRows; ++j)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[j][b] = 0.f;

    int i = lane;
    for (; i + 32 < nvec; i += 64) {  // two K-chunks x four rows = 8 loads in flight
      uint4 w0[kRows], w1[kRows];
#pragma unroll
      for (int j = 0; j < kRows; ++j) w0[j] = ld_nc_v4(wr[j] + i);
#pragma unroll
      for (int j = 0; j < kRows; ++j) w1[j] = ld_nc_v4(wr[j] + i + 32);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float xv[8];
        uint4 xs = *reinterpret_cast<const uint4*>(sx + (long long)b * K + (long long)i * 8);
        float2 a = unpack_bf16(xs.x), bb = unpack_bf16(xs.y), c = unpack_bf16(xs.z), d = unpack_bf16(xs.w);
        xv[0] = a.x; xv[1] = a.y; xv[2] = bb.x; xv[3] = bb.y; xv[4] = c.x; xv[5] = c.y; xv[6] = d.x; xv[7] = d.y;
#pragma unroll
        for (int j = 0; j < kRows; ++j) fma8(acc[j][b], w0[j], xv);
        xs = *reinterpret_cast<const uint4*>(sx + (long long)b * K + (long long)(i + 32) * 8);
        a = unpack_bf16(xs.x); bb = unpack_bf16(xs.y); c = unpack_bf16(xs.z); d = unpack_bf16(xs.w);
        xv[0] = a.x; xv[1] = a.y; xv[2] = bb.x; xv[3] = bb.y; xv[4] = c.x; xv[5] = c.y; xv[6] = d.x; xv[7] = d.y;
#pragma unroll
        for (int j = 0; j < kRows; ++j) fma8(acc[j][b], w1[j], xv);
      }
    }
    for (; i < nvec; i += 32) {
      uint4 w0[kRows];
#pragma unroll
      for (int j = 0; j < kRows; ++j) w0[j] = ld_nc_v4(wr[j] + i);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float xv[8];
        uint4 xs = *reinterpret_cast<const uint4*>(sx + (long long)b * K + (long long)i * 8);
        float2 a = unpack_bf16(xs.x), bb = unpack_bf16(xs.y), c = unpack_bf16(xs.z), d = unpack_bf16(xs.w);
        xv[0] = a.x; xv[1] = a.y; xv[2] = bb.x; xv[3] = bb.y; xv[4] = c.x; xv[5] = c.y; xv[6] = d.x; xv[7] = d.y;
#pragma unroll
        for (int j = 0; j < kRows; ++j) fma8(acc[j][b], w0[j], xv);
      }
    }
#pragma unroll
    for (int j = 0; j < kRows; ++j)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[j][b] = warp_sum(acc[j][b]);

    if (lane == 0) {
      if (p.epi == EPI_SWIGLU) {
        const int col0 = grp * 2;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            if (col0 + c < (p.N >> 1)) {
              const float val = silu_f(acc[2 * c][b]) * acc[2 * c + 1][b];
              static_cast<bf16*>(p.out)[(long long)b * p.ldo + col0 + c] = __float2bfloat16(val);
            }
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < kRows; ++j) {
          const int row = grp * kRows + j;
          if (row < p.N) {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              float val = acc[j][b];
              if (p.bias) val += __bfloat162float(p.bias[row]);
              if (p.epi == EPI_RES) val += __bfloat162float(p.res[(long long)b * p.ldr + row]);
              if (p.out_f32) static_cast<float*>(p.out)[(long long)b * p.ldo + row] = val;
              else static_cast<bf16*>(p.out)[(long long)b * p.ldo + row] = __float2bfloat16(val);
            }
          }
        }
      }
    }
  }
}

template <int NB>
static int launch_gemv(const GemvParams& p, cudaStream_t st) {
  const size_t smem = (size_t)NB * p.K * 2;
  static size_t attr_smem_dev[kMaxDevices] = {};
  size_t& attr_smem = attr_smem_dev[cur_device()];
  if (smem > 48 * 1024 && smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(gemv_bf16_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
    attr_smem = smem;
  }
  // CTAs per SM bounded by shared memory; every warp gets at least one row group when possible
  int per_sm = (int)((200 * 1024) / (smem + 1024));
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  int grid = num_sms() * per_sm;
  int need = (p.num_groups + kGemvWarps - 1) / kGemvWarps;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  gemv_bf16_kernel<NB><<<grid, kGemvThreads, smem, st>>>(p);
  return check_launch("gemv");
}

}  // namespace omc

using namespace omc;

extern "C" int omc_gemv_bf16(const void* x, long long ldx, const void* W, long long ldw, void* out, long long ldo, int B,
                             int N, int K, const void* norm_w, float eps, const void* bias, const void* res,
                             long long ldr, int epi, int out_is_f32, void* stream) {
  if (B < 1 || B > 8) return set_error(OMC_ERR_SHAPE, "omc_gemv_bf16: batch must be 1..8 (use omc_gemm_bf16 beyond)");
  if (N <= 0 || K <= 0 || K % 8 != 0) return set_error(OMC_ERR_SHAPE, "omc_gemv_bf16: K must be a positive multiple of 8");
  if (ldx % 8 != 0 || ldw % 8 != 0) return set_error(OMC_ERR_ALIGN, "omc_gemv_bf16: leading dims must be multiples of 8");
  if ((size_t)B * K * 2 > 200 * 1024) return set_error(OMC_ERR_SHAPE, "omc_gemv_bf16: B*K too large for shared memory");
  if (epi != EPI_NONE && epi != EPI_RES && epi != EPI_SWIGLU) return set_error(OMC_ERR_ARG, "omc_gemv_bf16: unknown epilogue");
  if (epi == EPI_RES && res == nullptr) return set_error(OMC_ERR_ARG, "omc_gemv_bf16: EPI_RES needs a residual");
  if (epi == EPI_SWIGLU && (N % 4 != 0 || bias != nullptr || out_is_f32))
    return set_error(OMC_ERR_ARG, "omc_gemv_bf16: SwiGLU needs N % 4 == 0, no bias, bf16 output");
  GemvParams p;
  p.x = (const bf16*)x; p.ldx = ldx; p.W = (const bf16*)W; p.ldw = ldw; p.out = out; p.ldo = ldo;
  p.N = N; p.K = K; p.norm_w = (const bf16*)norm_w; p.eps = eps; p.bias = (const bf16*)bias;
  p.res = (const bf16*)res; p.ldr = ldr; p.epi = epi; p.out_f32 = out_is_f32;
  p.num_groups = (N + kRows - 1) / kRows;
  cudaStream_t st = (cudaStream_t)stream;
  switch (B) {
    case 1: return launch_gemv<1>(p, st);
    case 2: return launch_gemv<2>(p, st);
    case 3: return launch_gemv<3>(p, st);
    case 4: return launch_gemv<4>(p, st);
    case 5: return launch_gemv<5>(p, st);
    case 6: return launch_gemv<6>(p, st);
    case 7: return launch_gemv<7>(p, st);
    default: return launch_gemv<8>(p, st);
  }
}
