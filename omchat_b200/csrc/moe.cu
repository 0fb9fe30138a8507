// Routing side of the Qwen2-MoE sparse MLP block (transformers modeling_qwen2_moe.py:295-374, the language model behind the
// reference's omchat/model/language_model/omchat_qwen2_moe.py): router softmax + top-k, the sort of the (token, expert) pairs
// into expert-contiguous 128-row tiles for the grouped tcgen05 GEMM (omc_gemm_bf16_grouped), and the weighted combine of the
// expert outputs with the sigmoid-gated shared expert and the residual stream. All of it is HBM/L2-bound row work:
//   omc_moe_route    one CTA per token: (optional RMSNorm of the row,) E + 1 dot products against the L2-resident router /
//                    shared-gate rows spread over 8 warps, fp32 softmax, k rounds of warp arg-max, per-expert histogram
//   omc_moe_select   prefill sizes: the logits come from the tcgen05 GEMM (router rows stacked with the gate row); one warp per
//                    token does the softmax / top-k / histogram
//   omc_moe_plan     one CTA: padded segment starts, the tile -> expert table, counters reset for the next call
//   omc_moe_scatter  one CTA per token: the normed row is copied to its k slots (slot = segment start + atomic cursor)
//   omc_moe_combine  one CTA per token: h += sum_j w_j * y[slot_j] + sigmoid_gate * shared_y, fp32 accumulate, one rounding
// No host synchronisation anywhere: the block can be captured in a CUDA graph; the host sizes buffers for the worst case
// (T * k / 128 + E tiles).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "omc_internal.h"
#include "ptx.cuh"

namespace omc {

typedef __nv_bfloat16 bf16;
constexpr int kMoeMaxExperts = 128;  // 4 router logits per lane
constexpr int kMoeMaxTopK = 8;
constexpr int kMoeTile = 128;        // rows per GEMM tile = padding unit of an expert's segment

__device__ __forceinline__ float dot8(uint4 a, uint4 b) {
  float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), a2 = unpack_bf16(a.z), a3 = unpack_bf16(a.w);
  float2 b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y), b2 = unpack_bf16(b.z), b3 = unpack_bf16(b.w);
  return a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y + a2.x * b2.x + a2.y * b2.y + a3.x * b3.x + a3.y * b3.y;
}

// One warp: fp32 softmax over the E router logits of token t (logit[E] = the shared expert's gate logit), k rounds of warp
// arg-max (ties -> lowest expert index), optional renormalisation, histogram.
__device__ __forceinline__ void route_select(const float* logit_row, int E, int top_k, int norm_topk, bool has_gate, int t,
                                             int32_t* __restrict__ topk_ids, float* __restrict__ topk_w,
                                             float* __restrict__ shared_gate, int32_t* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  if (has_gate && lane == 0) shared_gate[t] = 1.f / (1.f + __expf(-logit_row[E]));
  float logit[kMoeMaxExperts / 32];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kMoeMaxExperts / 32; ++j) {
    logit[j] = (j * 32 + lane < E) ? logit_row[j * 32 + lane] : -INFINITY;
    mx = fmaxf(mx, logit[j]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float prob[kMoeMaxExperts / 32], sum = 0.f;
#pragma unroll
  for (int j = 0; j < kMoeMaxExperts / 32; ++j) {
    prob[j] = (j * 32 + lane < E) ? expf(logit[j] - mx) : 0.f;
    sum += prob[j];
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  float sel_w[kMoeMaxTopK];
  int sel_e[kMoeMaxTopK];
  float wsum = 0.f;
#pragma unroll
  for (int r = 0; r < kMoeMaxTopK; ++r) {
    if (r >= top_k) break;
    float bv = -1.f;
    int be = 0x7fffffff;
#pragma unroll
    for (int j = 0; j < kMoeMaxExperts / 32; ++j) {
      const int e = j * 32 + lane;
      if (e < E && prob[j] > bv) {
        bv = prob[j];
        be = e;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oe = __shfl_xor_sync(0xffffffffu, be, o);
      if (ov > bv || (ov == bv && oe < be)) {
        bv = ov;
        be = oe;
      }
    }
    if (be >= E) {  // only with NaN logits (no comparison succeeds): stay inside the expert table, the output is NaN either way
      be = r % E;
      bv = 0.f;
    }
    sel_w[r] = bv * inv;
    sel_e[r] = be;
    wsum += sel_w[r];
    if ((be & 31) == lane) {
#pragma unroll
      for (int j = 0; j < kMoeMaxExperts / 32; ++j)
        if (j == (be >> 5)) prob[j] = -2.f;  // taken
    }
  }
  if (lane == 0) {
    const float rn = norm_topk ? 1.f / wsum : 1.f;
#pragma unroll
    for (int r = 0; r < kMoeMaxTopK; ++r) {
      if (r >= top_k) break;
      topk_ids[(long long)t * top_k + r] = sel_e[r];
      topk_w[(long long)t * top_k + r] = sel_w[r] * rn;
      atomicAdd(counts + sel_e[r], 1);
    }
  }
}

// Prefill-sized T: TOK tokens per CTA (one warp each for the row staging / RMSNorm and for the selection), so that every router
// row fetched from L2 is used for TOK dot products - a CTA per token re-reads the whole router matrix (E x C x 2 B = 245 KB)
// per token: 2 GB of L2 traffic per layer at T = 8192, 48 us per 1024 tokens.
template <int TOK>
__global__ void __launch_bounds__(TOK * 32) moe_route_rows_kernel(const bf16* __restrict__ x, long long ldx, int T, int C,
                                                                  const bf16* __restrict__ norm_w, float eps,
                                                                  bf16* __restrict__ xn_out, long long ldn,
                                                                  const bf16* __restrict__ router_w,
                                                                  const bf16* __restrict__ shared_gate_w, int E, int top_k,
                                                                  int norm_topk, int32_t* __restrict__ topk_ids,
                                                                  float* __restrict__ topk_w, float* __restrict__ shared_gate,
                                                                  int32_t* __restrict__ counts) {
  pdl_launch();
  pdl_wait();
  extern __shared__ uint4 s_rows[];  // [TOK][C / 8]
  __shared__ float s_logit[TOK][kMoeMaxExperts + 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  for (int t0 = blockIdx.x * TOK; t0 < T; t0 += gridDim.x * TOK) {
    const int t = t0 + warp;
    uint4* srow = s_rows + warp * nvec;
    if (t < T) {
      const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)t * ldx);
      if (norm_w != nullptr) {
        float ss = 0.f;
        for (int i = lane; i < nvec; i += 32) {
          const uint4 v = xr[i];
          srow[i] = v;
          float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
          ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
        }
        ss = warp_sum(ss);
        const float rstd = rsqrtf(ss / (float)C + eps);
        uint4* on = reinterpret_cast<uint4*>(xn_out + (long long)t * ldn);
        for (int i = lane; i < nvec; i += 32) {
          const uint4 v = srow[i], ww = __ldg(reinterpret_cast<const uint4*>(norm_w) + i);
          const uint32_t xi[4] = {v.x, v.y, v.z, v.w}, wi[4] = {ww.x, ww.y, ww.z, ww.w};
          uint32_t oo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 a = unpack_bf16(xi[q]), g = unpack_bf16(wi[q]);
            const float2 n = unpack_bf16(pack_bf16(a.x * rstd, a.y * rstd));
            oo[q] = pack_bf16(n.x * g.x, n.y * g.y);
          }
          const uint4 o = make_uint4(oo[0], oo[1], oo[2], oo[3]);
          srow[i] = o;
          on[i] = o;
        }
      } else {
        for (int i = lane; i < nvec; i += 32) srow[i] = xr[i];
      }
    } else {
      for (int i = lane; i < nvec; i += 32) srow[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();
    const int n_rows = E + (shared_gate_w != nullptr ? 1 : 0);
    for (int e = warp; e < n_rows; e += TOK) {
      const uint4* wr = reinterpret_cast<const uint4*>(e < E ? router_w + (long long)e * C : shared_gate_w);
      float acc[TOK];
#pragma unroll
      for (int j = 0; j < TOK; ++j) acc[j] = 0.f;
      for (int i = lane; i < nvec; i += 32) {
        const uint4 wv = __ldg(wr + i);
#pragma unroll
        for (int j = 0; j < TOK; ++j) acc[j] += dot8(s_rows[j * nvec + i], wv);
      }
#pragma unroll
      for (int j = 0; j < TOK; ++j) {
        const float a = warp_sum(acc[j]);
        if (lane == 0) s_logit[j][e] = a;
      }
    }
    __syncthreads();
    if (t < T) route_select(s_logit[warp], E, top_k, norm_topk, shared_gate_w != nullptr, t, topk_ids, topk_w, shared_gate, counts);
    __syncthreads();
  }
}

// Qwen2MoeTopKRouter.forward (modeling_qwen2_moe.py:343-352) + the shared expert's sigmoid gate (:371). One CTA per token:
// the row is staged in shared memory (optionally RMS-normalised on the way in - the decode step hands over the raw residual
// row and gets the normed row back in xn_out), the E + 1 dot products are spread over the 8 warps (independent 128-bit loads of
// the L2-resident router rows in flight per lane), warp 0 finishes with the fp32 softmax and k rounds of arg-max.
// Used for decode steps (T <= 32) with 1024 threads: a handful of CTAs must pull the whole router matrix themselves and only
// more warps put more loads in flight (T = 1: 31 us with 8 warps). Larger T: moe_route_rows_kernel.
template <int kRouteThreads>
__global__ void __launch_bounds__(kRouteThreads) moe_route_kernel(const bf16* __restrict__ x, long long ldx, int T, int C,
                                                                  const bf16* __restrict__ norm_w, float eps,
                                                                  bf16* __restrict__ xn_out, long long ldn,
                                                                  const bf16* __restrict__ router_w,
                                                                  const bf16* __restrict__ shared_gate_w, int E, int top_k,
                                                                  int norm_topk, int32_t* __restrict__ topk_ids,
                                                                  float* __restrict__ topk_w, float* __restrict__ shared_gate,
                                                                  int32_t* __restrict__ counts) {
  pdl_launch();
  pdl_wait();
  extern __shared__ uint4 s_row[];  // [C / 8]
  __shared__ float s_logit[kMoeMaxExperts + 1];
  __shared__ float s_red[kRouteThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  for (int t = blockIdx.x; t < T; t += gridDim.x) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)t * ldx);
    if (norm_w != nullptr) {
      // Qwen2MoeRMSNorm (:70-75) exactly as omc_rmsnorm does it: bf16(x * rstd) * w
      float ss = 0.f;
      for (int i = threadIdx.x; i < nvec; i += kRouteThreads) {
        const uint4 v = xr[i];
        s_row[i] = v;
        float2 a = unpack_bf16(v.x), b = unpack_bf16(v.y), c = unpack_bf16(v.z), d = unpack_bf16(v.w);
        ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
      }
      ss = warp_sum(ss);
      if (lane == 0) s_red[warp] = ss;
      __syncthreads();
      float tot = 0.f;
#pragma unroll
      for (int i = 0; i < kRouteThreads / 32; ++i) tot += s_red[i];
      const float rstd = rsqrtf(tot / (float)C + eps);
      uint4* on = reinterpret_cast<uint4*>(xn_out + (long long)t * ldn);
      for (int i = threadIdx.x; i < nvec; i += kRouteThreads) {
        const uint4 v = s_row[i], ww = reinterpret_cast<const uint4*>(norm_w)[i];
        const uint32_t xi[4] = {v.x, v.y, v.z, v.w}, wi[4] = {ww.x, ww.y, ww.z, ww.w};
        uint32_t oo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 a = unpack_bf16(xi[q]), g = unpack_bf16(wi[q]);
          const float2 n = unpack_bf16(pack_bf16(a.x * rstd, a.y * rstd));
          oo[q] = pack_bf16(n.x * g.x, n.y * g.y);
        }
        const uint4 o = make_uint4(oo[0], oo[1], oo[2], oo[3]);
        s_row[i] = o;
        on[i] = o;
      }
    } else {
      for (int i = threadIdx.x; i < nvec; i += kRouteThreads) s_row[i] = xr[i];
    }
    __syncthreads();
    // rows 0..E-1: router, row E: the shared expert's gate
    const int n_rows = E + (shared_gate_w != nullptr ? 1 : 0);
    for (int e = warp; e < n_rows; e += kRouteThreads / 32) {
      const uint4* wr = reinterpret_cast<const uint4*>(e < E ? router_w + (long long)e * C : shared_gate_w);
      float acc = 0.f;
#pragma unroll 4
      for (int i = lane; i < nvec; i += 32) acc += dot8(s_row[i], __ldg(wr + i));
      acc = warp_sum(acc);
      if (lane == 0) s_logit[e] = acc;
    }
    __syncthreads();
    if (warp == 0)
      route_select(s_logit, E, top_k, norm_topk, shared_gate_w != nullptr, t, topk_ids, topk_w, shared_gate, counts);
    __syncthreads();
  }
}

// Prefill sizes: the E + 1 logits per token come out of the tcgen05 GEMM ([T, C] x [router rows | shared-gate row | zero
// padding]^T, fp32 out); this kernel is only the selection - one warp per token.
__global__ void __launch_bounds__(256) moe_select_kernel(const float* __restrict__ logits, long long ld, int T, int E, int top_k,
                                                         int norm_topk, int has_gate, int32_t* __restrict__ topk_ids,
                                                         float* __restrict__ topk_w, float* __restrict__ shared_gate,
                                                         int32_t* __restrict__ counts) {
  pdl_launch();
  pdl_wait();
  const int warps = blockDim.x >> 5;
  for (int t = blockIdx.x * warps + (threadIdx.x >> 5); t < T; t += gridDim.x * warps)
    route_select(logits + (long long)t * ld, E, top_k, norm_topk, has_gate != 0, t, topk_ids, topk_w, shared_gate, counts);
}

__global__ void __launch_bounds__(128) moe_plan_kernel(int32_t* __restrict__ counts, int E, int max_tiles,
                                                       int32_t* __restrict__ seg_start, int32_t* __restrict__ cursor,
                                                       int32_t* __restrict__ tile_expert) {
  pdl_launch();
  pdl_wait();
  __shared__ int s_start[kMoeMaxExperts + 1];
  if (threadIdx.x == 0) {
    int at = 0;
    for (int e = 0; e < E; ++e) {
      s_start[e] = at;
      at += (counts[e] + kMoeTile - 1) / kMoeTile;  // in tiles
    }
    s_start[E] = at;
  }
  __syncthreads();
  for (int mt = threadIdx.x; mt < max_tiles; mt += blockDim.x) {
    int owner = -1;
    if (mt < s_start[E]) {
      int lo = 0, hi = E;  // last e with s_start[e] <= mt (empty experts share a start with their successor: take the last)
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_start[mid] <= mt) lo = mid; else hi = mid;
      }
      owner = lo;
    }
    tile_expert[mt] = owner;
  }
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    seg_start[e] = s_start[e] * kMoeTile;
    cursor[e] = 0;
    counts[e] = 0;  // ready for the next omc_moe_route
  }
}

__global__ void __launch_bounds__(128) moe_scatter_kernel(const bf16* __restrict__ x, long long ldx, int T, int C,
                                                          const int32_t* __restrict__ topk_ids, int top_k,
                                                          const int32_t* __restrict__ seg_start, int32_t* __restrict__ cursor,
                                                          bf16* __restrict__ xperm, long long ldp, int32_t* __restrict__ slot_of) {
  pdl_launch();
  pdl_wait();
  __shared__ int s_slot[kMoeMaxTopK];
  const int nvec = C >> 3;
  for (int t = blockIdx.x; t < T; t += gridDim.x) {
    if (threadIdx.x < top_k) {
      const int e = topk_ids[(long long)t * top_k + threadIdx.x];
      const int slot = seg_start[e] + atomicAdd(cursor + e, 1);
      s_slot[threadIdx.x] = slot;
      slot_of[(long long)t * top_k + threadIdx.x] = slot;
    }
    __syncthreads();
    const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)t * ldx);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      const uint4 v = xr[i];
      for (int r = 0; r < top_k; ++r) reinterpret_cast<uint4*>(xperm + (long long)s_slot[r] * ldp)[i] = v;
    }
    __syncthreads();
  }
}

// Decode steps with at most 16 (token, expert) pairs: plan + scatter in ONE single-CTA launch.
__global__ void __launch_bounds__(128) moe_plan_scatter_small_kernel(int32_t* __restrict__ counts, int E, int max_tiles,
                                                                     int32_t* __restrict__ seg_start, int32_t* __restrict__ cursor,
                                                                     int32_t* __restrict__ tile_expert, const bf16* __restrict__ x,
                                                                     long long ldx, int T, int C,
                                                                     const int32_t* __restrict__ topk_ids, int top_k,
                                                                     bf16* __restrict__ xperm, long long ldp,
                                                                     int32_t* __restrict__ slot_of) {
  pdl_launch();
  pdl_wait();
  __shared__ int s_start[kMoeMaxExperts + 1];
  __shared__ int s_slot[16];
  if (threadIdx.x == 0) {
    int at = 0;
    for (int e = 0; e < E; ++e) {
      s_start[e] = at;
      at += (counts[e] + kMoeTile - 1) / kMoeTile;
    }
    s_start[E] = at;
    // slots in (token, rank) order: with <= 64 pairs a serial pass is cheaper than atomics + a second launch
    for (int p = 0; p < T * top_k; ++p) {
      const int e = topk_ids[p];
      int n = 0;  // pairs before this one that went to the same expert
      for (int q = 0; q < p; ++q) n += (topk_ids[q] == e);
      s_slot[p] = s_start[e] * kMoeTile + n;
      slot_of[p] = s_slot[p];
    }
  }
  __syncthreads();
  for (int mt = threadIdx.x; mt < max_tiles; mt += blockDim.x) {
    int owner = -1;
    if (mt < s_start[E]) {
      int lo = 0, hi = E;
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_start[mid] <= mt) lo = mid; else hi = mid;
      }
      owner = lo;
    }
    tile_expert[mt] = owner;
  }
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    seg_start[e] = s_start[e] * kMoeTile;
    cursor[e] = counts[e];  // what omc_moe_scatter would have left behind
    counts[e] = 0;
  }
  const int nvec = C >> 3;
  for (int p = 0; p < T * top_k; ++p) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)(p / top_k) * ldx);
    uint4* dst = reinterpret_cast<uint4*>(xperm + (long long)s_slot[p] * ldp);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = xr[i];
  }
}

// Qwen2MoeSparseMoeBlock.forward :363-374 + the decoder layer's residual add (:420): fp32 accumulate, one rounding.
__global__ void __launch_bounds__(128) moe_combine_kernel(bf16* __restrict__ h, long long ldh, int T, int C,
                                                          const bf16* __restrict__ yperm, long long ldy,
                                                          const int32_t* __restrict__ slot_of, const float* __restrict__ topk_w,
                                                          int top_k, const bf16* __restrict__ shared_y, long long lds,
                                                          const float* __restrict__ shared_gate, float* __restrict__ ssq_out,
                                                          int ssq_parts) {
  pdl_launch();  // the next layer's qkv GEMM (a programmatic dependent) may start its weight prefetch
  pdl_wait();
  __shared__ int s_slot[kMoeMaxTopK];
  __shared__ float s_w[kMoeMaxTopK];
  __shared__ float s_red[4];
  const int nvec = C >> 3;
  for (int t = blockIdx.x; t < T; t += gridDim.x) {
    if (threadIdx.x < top_k) {
      s_slot[threadIdx.x] = slot_of[(long long)t * top_k + threadIdx.x];
      s_w[threadIdx.x] = topk_w[(long long)t * top_k + threadIdx.x];
    }
    __syncthreads();
    const float sg = shared_y != nullptr ? shared_gate[t] : 0.f;
    uint4* hr = reinterpret_cast<uint4*>(h + (long long)t * ldh);
    float ssq = 0.f;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      float moe[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int r = 0; r < top_k; ++r) {
        const uint4 v = reinterpret_cast<const uint4*>(yperm + (long long)s_slot[r] * ldy)[i];
        const uint32_t vi[4] = {v.x, v.y, v.z, v.w};
        const float w = s_w[r];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 a = unpack_bf16(vi[q]);
          moe[2 * q] += w * a.x;
          moe[2 * q + 1] += w * a.y;
        }
      }
      if (shared_y != nullptr) {
        const uint4 v = reinterpret_cast<const uint4*>(shared_y + (long long)t * lds)[i];
        const uint32_t vi[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float2 a = unpack_bf16(vi[q]);
          moe[2 * q] += sg * a.x;
          moe[2 * q + 1] += sg * a.y;
        }
      }
      const uint4 hv = hr[i];
      const uint32_t hi[4] = {hv.x, hv.y, hv.z, hv.w};
      uint32_t o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 a = unpack_bf16(hi[q]);
        o[q] = pack_bf16(a.x + moe[2 * q], a.y + moe[2 * q + 1]);
        const float2 r = unpack_bf16(o[q]);  // of the bf16 values actually stored, like the GEMM epilogues' partials
        ssq += r.x * r.x + r.y * r.y;
      }
      hr[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (ssq_out != nullptr) {
      // the row's sum of squares for the folded RMSNorm of the next weight-streaming GEMM (layout of omc_row_ssq: [parts][64])
      ssq = warp_sum(ssq);
      if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = ssq;
      __syncthreads();
      if (threadIdx.x == 0) {
        ssq_out[t] = s_red[0] + s_red[1] + s_red[2] + s_red[3];
        for (int q = 1; q < ssq_parts; ++q) ssq_out[q * 64 + t] = 0.f;
      }
    }
    __syncthreads();
  }
}

// The routing kernels run between weight-streaming GEMMs in the decode step and CAN be launched as programmatic dependents
// (their scheduling overlaps the predecessor's tail; every kernel above starts with pdl_launch + pdl_wait): omc_moe_set_pdl(1).
// Off by default - measured at Qwen1.5-MoE-A2.7B sizes: batch-1 step 2.02 ms with, 1.93 ms without (the early-scheduled
// dependents take SM slots from the predecessor's tail), batch 32 6.74 vs 6.84 ms, prefill 37.5 vs 38.1 ms: no clear gain.
static int g_moe_pdl = 0;
template <typename... KArgs, typename... Args>
static cudaError_t moe_launch(void (*kernel)(KArgs...), int grid, int block, int smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = g_moe_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
int moe_pdl_enabled() { return g_moe_pdl; }

}  // namespace omc

using namespace omc;

extern "C" int omc_moe_set_pdl(int on) {
  g_moe_pdl = on ? 1 : 0;
  return OMC_OK;
}

extern "C" int omc_moe_max_tiles(int T, int top_k, int n_experts) {
  if (T <= 0 || top_k <= 0 || n_experts <= 0) return 0;
  return (int)(((long long)T * top_k) / kMoeTile) + n_experts;  // sum_e ceil(c_e / 128) <= floor(sum c_e / 128) + E
}

extern "C" int omc_moe_route(const void* x, long long ldx, int T, int C, const void* norm_w, float eps, void* xn_out,
                             long long ldn, const void* router_w, const void* shared_gate_w, int n_experts, int top_k,
                             int norm_topk, int32_t* topk_ids, float* topk_w, float* shared_gate, int32_t* counts, void* stream) {
  if (T <= 0) return OMC_OK;
  if (x == nullptr || router_w == nullptr || topk_ids == nullptr || topk_w == nullptr || counts == nullptr ||
      (shared_gate_w != nullptr && shared_gate == nullptr) || (norm_w != nullptr && xn_out == nullptr))
    return set_error(OMC_ERR_ARG, "omc_moe_route: null argument");
  if (n_experts < 1 || n_experts > kMoeMaxExperts || top_k < 1 || top_k > kMoeMaxTopK || top_k > n_experts)
    return set_error(OMC_ERR_SHAPE, "omc_moe_route: at most 128 experts and top-8 routing");
  if (C <= 0 || C % 8 != 0 || C > 8192 || ldx % 8 != 0 || (norm_w != nullptr && ldn % 8 != 0))
    return set_error(OMC_ERR_SHAPE, "omc_moe_route: C must be a multiple of 8 and <= 8192, leading dims multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  if (T <= 32) {
    const int grid = T;
    moe_launch(moe_route_kernel<1024>, grid, 1024, (C / 8) * 16, st,
        (const bf16*)x, ldx, T, C, (const bf16*)norm_w, eps, (bf16*)xn_out, ldn, (const bf16*)router_w,
        (const bf16*)shared_gate_w, n_experts, top_k, norm_topk, topk_ids, topk_w, shared_gate, counts);
  } else {
    constexpr int TOK = 8;
    const int smem = TOK * (C / 8) * 16;  // 32 KB at C = 2048, 128 KB at C = 8192
    static bool attr_set_dev[kMaxDevices] = {};
    bool& attr_set = attr_set_dev[cur_device()];
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(moe_route_rows_kernel<TOK>, cudaFuncAttributeMaxDynamicSharedMemorySize, TOK * 1024 * 16);
      if (e != cudaSuccess) return set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
      attr_set = true;
    }
    const int need = (T + TOK - 1) / TOK;
    const int grid = need < num_sms() * 4 ? need : num_sms() * 4;
    moe_launch(moe_route_rows_kernel<TOK>, grid, TOK * 32, smem, st,
        (const bf16*)x, ldx, T, C, (const bf16*)norm_w, eps, (bf16*)xn_out, ldn, (const bf16*)router_w,
        (const bf16*)shared_gate_w, n_experts, top_k, norm_topk, topk_ids, topk_w, shared_gate, counts);
  }
  return check_launch("moe_route");
}

extern "C" int omc_moe_select(const float* logits, long long ld, int T, int n_experts, int top_k, int norm_topk, int has_gate,
                              int32_t* topk_ids, float* topk_w, float* shared_gate, int32_t* counts, void* stream) {
  if (T <= 0) return OMC_OK;
  if (logits == nullptr || topk_ids == nullptr || topk_w == nullptr || counts == nullptr || (has_gate && shared_gate == nullptr))
    return set_error(OMC_ERR_ARG, "omc_moe_select: null argument");
  if (n_experts < 1 || n_experts > kMoeMaxExperts || top_k < 1 || top_k > kMoeMaxTopK || top_k > n_experts ||
      ld < n_experts + (has_gate ? 1 : 0))
    return set_error(OMC_ERR_SHAPE, "omc_moe_select: at most 128 experts, top-8 routing, ld >= experts (+ 1 with a gate column)");
  const int need = (T + 7) / 8;
  const int grid = need < num_sms() * 8 ? need : num_sms() * 8;
  moe_launch(moe_select_kernel, grid, 256, 0, (cudaStream_t)stream, logits, ld, T, n_experts, top_k, norm_topk, has_gate, topk_ids, topk_w,
                                                            shared_gate, counts);
  return check_launch("moe_select");
}

extern "C" int omc_moe_plan(int32_t* counts, int n_experts, int max_tiles, int32_t* seg_start, int32_t* cursor,
                            int32_t* tile_expert, void* stream) {
  if (counts == nullptr || seg_start == nullptr || cursor == nullptr || tile_expert == nullptr)
    return set_error(OMC_ERR_ARG, "omc_moe_plan: null argument");
  if (n_experts < 1 || n_experts > kMoeMaxExperts || max_tiles < 1) return set_error(OMC_ERR_SHAPE, "omc_moe_plan: bad sizes");
  moe_launch(moe_plan_kernel, 1, 128, 0, (cudaStream_t)stream, counts, n_experts, max_tiles, seg_start, cursor, tile_expert);
  return check_launch("moe_plan");
}

extern "C" int omc_moe_scatter(const void* x, long long ldx, int T, int C, const int32_t* topk_ids, int top_k,
                               const int32_t* seg_start, int32_t* cursor, void* xperm, long long ldp, int32_t* slot_of,
                               void* stream) {
  if (T <= 0) return OMC_OK;
  if (x == nullptr || topk_ids == nullptr || seg_start == nullptr || cursor == nullptr || xperm == nullptr || slot_of == nullptr)
    return set_error(OMC_ERR_ARG, "omc_moe_scatter: null argument");
  if (C % 8 != 0 || ldx % 8 != 0 || ldp % 8 != 0 || top_k < 1 || top_k > kMoeMaxTopK)
    return set_error(OMC_ERR_SHAPE, "omc_moe_scatter: C / leading dims must be multiples of 8, top_k <= 8");
  const int grid = T < num_sms() * 16 ? T : num_sms() * 16;
  moe_launch(moe_scatter_kernel, grid, 128, 0, (cudaStream_t)stream, (const bf16*)x, ldx, T, C, topk_ids, top_k, seg_start, cursor,
                                                             (bf16*)xperm, ldp, slot_of);
  return check_launch("moe_scatter");
}

extern "C" int omc_moe_plan_scatter(int32_t* counts, int n_experts, int max_tiles, int32_t* seg_start, int32_t* cursor,
                                    int32_t* tile_expert, const void* x, long long ldx, int T, int C, const int32_t* topk_ids,
                                    int top_k, void* xperm, long long ldp, int32_t* slot_of, void* stream) {
  if (T * top_k > 16) {
    const int rc = omc_moe_plan(counts, n_experts, max_tiles, seg_start, cursor, tile_expert, stream);
    if (rc != OMC_OK) return rc;
    return omc_moe_scatter(x, ldx, T, C, topk_ids, top_k, seg_start, cursor, xperm, ldp, slot_of, stream);
  }
  if (T <= 0) return OMC_OK;
  if (counts == nullptr || seg_start == nullptr || cursor == nullptr || tile_expert == nullptr || x == nullptr ||
      topk_ids == nullptr || xperm == nullptr || slot_of == nullptr)
    return set_error(OMC_ERR_ARG, "omc_moe_plan_scatter: null argument");
  if (n_experts < 1 || n_experts > kMoeMaxExperts || max_tiles < 1 || C % 8 != 0 || ldx % 8 != 0 || ldp % 8 != 0 || top_k < 1 ||
      top_k > kMoeMaxTopK)
    return set_error(OMC_ERR_SHAPE, "omc_moe_plan_scatter: bad sizes");
  moe_launch(moe_plan_scatter_small_kernel, 1, 128, 0, (cudaStream_t)stream, counts, n_experts, max_tiles, seg_start, cursor, tile_expert,
                                                                      (const bf16*)x, ldx, T, C, topk_ids, top_k, (bf16*)xperm, ldp,
                                                                      slot_of);
  return check_launch("moe_plan_scatter");
}

extern "C" int omc_moe_combine(void* h, long long ldh, int T, int C, const void* yperm, long long ldy, const int32_t* slot_of,
                               const float* topk_w, int top_k, const void* shared_y, long long lds, const float* shared_gate,
                               float* ssq_out, int ssq_parts, void* stream) {
  if (T <= 0) return OMC_OK;
  if (h == nullptr || yperm == nullptr || slot_of == nullptr || topk_w == nullptr || (shared_y != nullptr && shared_gate == nullptr))
    return set_error(OMC_ERR_ARG, "omc_moe_combine: null argument");
  if (C % 8 != 0 || ldh % 8 != 0 || ldy % 8 != 0 || lds % 8 != 0 || top_k < 1 || top_k > kMoeMaxTopK)
    return set_error(OMC_ERR_SHAPE, "omc_moe_combine: C / leading dims must be multiples of 8, top_k <= 8");
  if (ssq_out != nullptr && (T > 64 || ssq_parts < 1))
    return set_error(OMC_ERR_SHAPE, "omc_moe_combine: sums of squares are for decode steps of <= 64 rows (omc_row_ssq layout)");
  const int grid = T < num_sms() * 16 ? T : num_sms() * 16;
  moe_combine_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((bf16*)h, ldh, T, C, (const bf16*)yperm, ldy, slot_of, topk_w, top_k,
                                                             (const bf16*)shared_y, lds, shared_gate, ssq_out, ssq_parts);
  return check_launch("moe_combine");
}
