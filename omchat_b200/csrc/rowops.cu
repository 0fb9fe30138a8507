// HBM-bound row / integer kernels of the OmChat hot path: RMSNorm, patch im2col, embedding assembly,
// CLS-drop + pixel-shuffle gather, embed lookup, image-token splice, RoPE + paged KV append, argmax.
// All are coalesced 128-bit streaming kernels; none of them is GEMM-shaped.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "omc_internal.h"
#include "ptx.cuh"

namespace omc {

typedef __nv_bfloat16 bf16;

// ------------------------------------------------------------------------------------------------ RMSNorm
// out = w * bf16(x * rsqrt(mean(x^2) + eps))      [intern_vit_6b/modeling_intern_vit.py:39-44]
// One 128-thread CTA per row (grid-strided); a row of <= 4096 channels lives in registers between the two passes.
constexpr int kNormThreads = 128;
constexpr int kNormMaxVec = 4;  // 4 * 128 threads * 8 elements = 4096 channels

// segs = 2: every row holds TWO independent C-wide segments side by side (q | k of the ViT's packed qkv rows), normalised
// with w and w2 in one launch (InternAttention q_norm / k_norm, modeling_intern_vit.py:143-146).
__global__ void __launch_bounds__(kNormThreads) rmsnorm_kernel(const bf16* __restrict__ x, long long ldx,
                                                               const bf16* __restrict__ w, const bf16* __restrict__ w2,
                                                               bf16* __restrict__ out, long long ldo, int rows, int C,
                                                               float eps, int segs) {
  __shared__ float red[kNormThreads / 32];
  const int nvec = C >> 3;
  for (int vrow = blockIdx.x; vrow < rows * segs; vrow += gridDim.x) {
    const int row = vrow / segs, seg = vrow - row * segs;
    const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)row * ldx + (long long)seg * C);
    uint4 v[kNormMaxVec];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < kNormMaxVec; ++j) {
      int i = threadIdx.x + j * kNormThreads;
      if (i < nvec) {
        v[j] = xr[i];
        float2 a = unpack_bf16(v[j].x), b = unpack_bf16(v[j].y), c = unpack_bf16(v[j].z), d = unpack_bf16(v[j].w);
        ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
      }
    }
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < kNormThreads / 32; ++i) tot += red[i];
    __syncthreads();
    const float rstd = rsqrtf(tot / (float)C + eps);
    uint4* orow = reinterpret_cast<uint4*>(out + (long long)row * ldo + (long long)seg * C);
    const uint4* wv = reinterpret_cast<const uint4*>(seg == 0 ? w : w2);
#pragma unroll
    for (int j = 0; j < kNormMaxVec; ++j) {
      int i = threadIdx.x + j * kNormThreads;
      if (i < nvec) {
        uint4 ww = wv[i];
        uint32_t xi[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
        uint32_t wi[4] = {ww.x, ww.y, ww.z, ww.w};
        uint32_t oo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float2 a = unpack_bf16(xi[q]);
          float2 g = unpack_bf16(wi[q]);
          // normalised value is rounded to bf16 BEFORE the weight multiply, as the reference does
          float2 n = unpack_bf16(pack_bf16(a.x * rstd, a.y * rstd));
          oo[q] = pack_bf16(n.x * g.x, n.y * g.y);
        }
        orow[i] = make_uint4(oo[0], oo[1], oo[2], oo[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// out = bf16((x - mean) * rsqrt(var + eps) * w + b), statistics and the affine map in fp32, ONE rounding (torch.nn.LayerNorm on
// bf16 rows) - the norm_type = 'layer_norm' option of the InternViT-300M tower (intern_vit_300m/modeling_intern_vit.py:61-64,
// 209-210). Two-pass variance from the registers the row already sits in (no E[x^2] - mean^2 cancellation).
__device__ __forceinline__ float block_sum_128(float v, float* red) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < kNormThreads / 32; ++i) tot += red[i];
  __syncthreads();
  return tot;
}

__global__ void __launch_bounds__(kNormThreads) layernorm_kernel(const bf16* __restrict__ x, long long ldx,
                                                                 const bf16* __restrict__ w, const bf16* __restrict__ b,
                                                                 bf16* __restrict__ out, long long ldo, int rows, int C,
                                                                 float eps) {
  __shared__ float red[kNormThreads / 32];
  const int nvec = C >> 3;
  const float inv_c = 1.f / (float)C;
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
    float f[kNormMaxVec][8];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kNormMaxVec; ++j) {
      int i = threadIdx.x + j * kNormThreads;
      if (i < nvec) {
        uint4 v = xr[i];
        uint32_t xi[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float2 a = unpack_bf16(xi[q]);
          f[j][2 * q] = a.x;
          f[j][2 * q + 1] = a.y;
          s += a.x + a.y;
        }
      }
    }
    const float mean = block_sum_128(s, red) * inv_c;
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < kNormMaxVec; ++j) {
      int i = threadIdx.x + j * kNormThreads;
      if (i < nvec) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float dlt = f[j][q] - mean;
          ss += dlt * dlt;
        }
      }
    }
    const float rstd = rsqrtf(block_sum_128(ss, red) * inv_c + eps);
    uint4* orow = reinterpret_cast<uint4*>(out + (long long)row * ldo);
    const uint4* wv = reinterpret_cast<const uint4*>(w);
    const uint4* bv = reinterpret_cast<const uint4*>(b);
#pragma unroll
    for (int j = 0; j < kNormMaxVec; ++j) {
      int i = threadIdx.x + j * kNormThreads;
      if (i < nvec) {
        uint4 ww = wv[i];
        uint4 bb = b != nullptr ? bv[i] : make_uint4(0u, 0u, 0u, 0u);
        uint32_t wi[4] = {ww.x, ww.y, ww.z, ww.w};
        uint32_t bi[4] = {bb.x, bb.y, bb.z, bb.w};
        uint32_t oo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float2 g = unpack_bf16(wi[q]);
          float2 c = unpack_bf16(bi[q]);
          oo[q] = pack_bf16((f[j][2 * q] - mean) * rstd * g.x + c.x, (f[j][2 * q + 1] - mean) * rstd * g.y + c.y);
        }
        orow[i] = make_uint4(oo[0], oo[1], oo[2], oo[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ patch im2col
template <typename T>
__global__ void im2col_kernel(const T* __restrict__ pix, bf16* __restrict__ cols, long long ldc, int B, int H, int W) {
  const int gh = H / 14, gw = W / 14;
  const long long total = (long long)B * gh * gw * ldc;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(idx % ldc);
    const long long row = idx / ldc;
    float val = 0.f;
    if (col < 588) {
      const int c = col / 196, r = col % 196, ky = r / 14, kx = r % 14;
      const int b = (int)(row / (gh * gw)), pr = (int)(row % (gh * gw)), py = pr / gw, px = pr % gw;
      val = (float)pix[(((long long)b * 3 + c) * H + (py * 14 + ky)) * W + (px * 14 + kx)];
    }
    cols[idx] = __float2bfloat16(val);
  }
}

// ------------------------------------------------------------------------------------------------ CLS + pos-embed
__global__ void vit_assemble_kernel(const bf16* __restrict__ patch, const bf16* __restrict__ cls,
                                    const bf16* __restrict__ pos, bf16* __restrict__ hidden, int B, int P, int C) {
  const int nvec = C >> 3;
  const long long total = (long long)B * (P + 1) * nvec;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int vcol = (int)(idx % nvec);
    const long long row = idx / nvec;
    const int tok = (int)(row % (P + 1));
    const int b = (int)(row / (P + 1));
    uint4 a = (tok == 0) ? reinterpret_cast<const uint4*>(cls)[vcol]
                         : reinterpret_cast<const uint4*>(patch + ((long long)b * P + tok - 1) * C)[vcol];
    uint4 p = reinterpret_cast<const uint4*>(pos + (long long)tok * C)[vcol];
    uint32_t ai[4] = {a.x, a.y, a.z, a.w}, pi[4] = {p.x, p.y, p.z, p.w}, oo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float2 x = unpack_bf16(ai[q]), y = unpack_bf16(pi[q]);
      oo[q] = pack_bf16(x.x + y.x, x.y + y.y);
    }
    reinterpret_cast<uint4*>(hidden + row * C)[vcol] = make_uint4(oo[0], oo[1], oo[2], oo[3]);
  }
}

// ------------------------------------------------------------------------------------------------ CLS drop + pixel shuffle
// out[b, (i/d)*(G/d) + j/d, (i%d)*(d*C) + (j%d)*C + c] = hidden[b, 1 + i*G + j, c]
__global__ void select_pixel_shuffle_kernel(const bf16* __restrict__ hidden, bf16* __restrict__ out, int B, int G,
                                            int C, int d) {
  const int nvec = C >> 3;
  const long long total = (long long)B * G * G * nvec;
  const int Go = G / d;
  const long long Co = (long long)C * d * d;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int vcol = (int)(idx % nvec);
    const long long t = idx / nvec;
    const int j = (int)(t % G), i = (int)((t / G) % G), b = (int)(t / ((long long)G * G));
    uint4 v = reinterpret_cast<const uint4*>(hidden + ((long long)b * (G * G + 1) + 1 + (long long)i * G + j) * C)[vcol];
    const long long orow = (long long)b * Go * Go + (long long)(i / d) * Go + (j / d);
    const long long ocol = (long long)(i % d) * d * C + (long long)(j % d) * C + (long long)vcol * 8;
    *reinterpret_cast<uint4*>(out + orow * Co + ocol) = v;
  }
}

// ------------------------------------------------------------------------------------------------ embed lookup
// ids outside [0, vocab) (vocab > 0) produce a zero row instead of an out-of-bounds read (nn.Embedding raises there).
__global__ void embed_lookup_kernel(const int64_t* __restrict__ ids, int T, const bf16* __restrict__ table, int C,
                                    bf16* __restrict__ out, long long ldo, int vocab) {
  const int nvec = C >> 3;
  for (int t = blockIdx.x; t < T; t += gridDim.x) {
    const long long id = ids[t];
    const bool ok = vocab <= 0 || (id >= 0 && id < vocab);
    const uint4* src = reinterpret_cast<const uint4*>(table + (ok ? id : 0) * (long long)C);
    uint4* dst = reinterpret_cast<uint4*>(out + (long long)t * ldo);
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) dst[i] = ok ? src[i] : make_uint4(0, 0, 0, 0);
  }
}

// ------------------------------------------------------------------------------------------------ splice
// Plan (single CTA): gcount[i] = number of image placeholders strictly before packed token i (gcount[S_total] = all),
// img_base[s] = index of the first image block consumed by sequence s (a sequence without placeholders still
// consumes one block, omchat_arch.py:122-129), out_offsets[s] = packed start row of spliced sequence s.
constexpr int kPlanThreads = 1024;
__global__ void __launch_bounds__(kPlanThreads) splice_plan_kernel(const int64_t* __restrict__ ids,
                                                                   const int32_t* __restrict__ seq_offsets, int n_seq,
                                                                   int S_total, long long image_token, int L,
                                                                   int max_len, int32_t* __restrict__ gcount,
                                                                   int32_t* __restrict__ img_base,
                                                                   int32_t* __restrict__ out_offsets) {
  __shared__ int warp_tot[kPlanThreads / 32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < S_total; base += kPlanThreads) {
    const int i = base + threadIdx.x;
    const int flag = (i < S_total && ids[i] == image_token) ? 1 : 0;
    int incl = flag;  // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, incl, o);
      if ((int)(threadIdx.x & 31) >= o) incl += n;
    }
    if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = incl;
    __syncthreads();
    int wbase = 0;
    for (int wi = 0; wi < (int)(threadIdx.x >> 5); ++wi) wbase += warp_tot[wi];
    const int c = carry;
    if (i < S_total) gcount[i] = c + wbase + incl - flag;
    __syncthreads();
    if (threadIdx.x == kPlanThreads - 1) carry = c + wbase + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    gcount[S_total] = carry;
    int ib = 0, oo = 0;
    for (int s = 0; s < n_seq; ++s) {
      const int a = seq_offsets[s], b = seq_offsets[s + 1];
      const int ga = gcount[a];
      const int gb = (b < S_total) ? gcount[b] : carry;
      const int k = gb - ga;
      img_base[s] = ib - ga;  // image index of a placeholder at packed i in seq s = img_base[s] + gcount[i]
      ib += (k == 0) ? 1 : k;
      int len = (b - a) + k * (L - 1);
      if (max_len > 0 && len > max_len) len = max_len;
      out_offsets[s] = oo;
      oo += len;
    }
    img_base[n_seq] = ib;
    out_offsets[n_seq] = oo;
  }
}

// Gather: one warp per OUTPUT row; binary-search the source token whose destination span covers the row.
__global__ void splice_gather_kernel(const int64_t* __restrict__ ids, const int32_t* __restrict__ seq_offsets,
                                     int n_seq, long long image_token, const bf16* __restrict__ table,
                                     const bf16* __restrict__ feats, int n_img, int L, int C,
                                     const int32_t* __restrict__ gcount, const int32_t* __restrict__ img_base,
                                     const int32_t* __restrict__ out_offsets, bf16* __restrict__ embeds,
                                     int32_t* __restrict__ pos_ids, int32_t* __restrict__ seq_ids, int T_capacity,
                                     int vocab) {
  const int warps_per_cta = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int T_total = out_offsets[n_seq];
  const int nvec = C >> 3;
  for (int t = blockIdx.x * warps_per_cta + (threadIdx.x >> 5); t < T_total && t < T_capacity;
       t += gridDim.x * warps_per_cta) {
    int lo = 0, hi = n_seq - 1;  // sequence: largest s with out_offsets[s] <= t
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (out_offsets[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const int s = lo;
    const int dt = t - out_offsets[s];
    const int a = seq_offsets[s], b = seq_offsets[s + 1];
    const int ga = gcount[a];
    int l2 = a, h2 = b - 1;  // source token: largest i in [a,b) with dest(i) <= dt
    while (l2 < h2) {
      int mid = (l2 + h2 + 1) >> 1;
      int dmid = (mid - a) + (gcount[mid] - ga) * (L - 1);
      if (dmid <= dt) l2 = mid; else h2 = mid - 1;
    }
    const int i = l2;
    const int r = dt - ((i - a) + (gcount[i] - ga) * (L - 1));
    const long long id = ids[i];
    const uint4* src = nullptr;
    if (id == image_token) {
      const int img = img_base[s] + gcount[i];
      if (img < n_img) src = reinterpret_cast<const uint4*>(feats + ((long long)img * L + r) * C);
    } else if (vocab <= 0 || (id >= 0 && id < vocab)) {
      src = reinterpret_cast<const uint4*>(table + id * (long long)C);
    }  // else: id outside the vocabulary -> zero row, never an out-of-bounds read
    uint4* dst = reinterpret_cast<uint4*>(embeds + (long long)t * C);
    for (int v = lane; v < nvec; v += 32) dst[v] = src ? src[v] : make_uint4(0, 0, 0, 0);
    if (lane == 0) {
      pos_ids[t] = dt;
      seq_ids[t] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------ RoPE + paged KV append
// One CTA per token. q/k heads are rotated in place (pairs (i, i+64), fp32 angles = pos * inv_freq[i]); rotated K and
// V are appended to the paged pool.  [transformers qwen2 modeling_qwen2.py:102-113,124-146,227]
__global__ void __launch_bounds__(256) rope_kv_store_kernel(bf16* __restrict__ qkv, long long ld,
                                                            const int32_t* __restrict__ pos,
                                                            const int32_t* __restrict__ seq_ids, int T, int Hq, int Hkv,
                                                            const float* __restrict__ inv_freq,
                                                            bf16* __restrict__ pool,
                                                            const int32_t* __restrict__ block_table, int max_pages,
                                                            int page_size) {
  __shared__ float cs[64], sn[64];
  for (int t = blockIdx.x; t < T; t += gridDim.x) {
    const int p = pos[t];
    __syncthreads();
    if (threadIdx.x < 64) {
      float s, c;
      sincosf((float)p * inv_freq[threadIdx.x], &s, &c);
      cs[threadIdx.x] = c;
      sn[threadIdx.x] = s;
    }
    __syncthreads();
    bf16* row = qkv + (long long)t * ld;
    const int sq = seq_ids ? seq_ids[t] : 0;
    const int page = block_table[(long long)sq * max_pages + p / page_size];
    const int slot = p % page_size;
    bf16* kdst = pool + (((long long)page * 2 + 0) * Hkv * page_size + slot) * 128;
    bf16* vdst = pool + (((long long)page * 2 + 1) * Hkv * page_size + slot) * 128;
    const int nrot = (Hq + Hkv) * 64;
    for (int e = threadIdx.x; e < nrot; e += blockDim.x) {
      const int h = e >> 6, i = e & 63;
      bf16* hp = row + h * 128;
      const float x0 = __bfloat162float(hp[i]), x1 = __bfloat162float(hp[i + 64]);
      const bf16 y0 = __float2bfloat16(x0 * cs[i] - x1 * sn[i]);
      const bf16 y1 = __float2bfloat16(x1 * cs[i] + x0 * sn[i]);
      hp[i] = y0;
      hp[i + 64] = y1;
      if (h >= Hq) {
        bf16* kd = kdst + (long long)(h - Hq) * page_size * 128;
        kd[i] = y0;
        kd[i + 64] = y1;
      }
    }
    const bf16* vsrc = row + (Hq + Hkv) * 128;
    for (int e = threadIdx.x; e < Hkv * 16; e += blockDim.x) {
      const int h = e >> 4, v = e & 15;
      reinterpret_cast<uint4*>(vdst + (long long)h * page_size * 128)[v] =
          reinterpret_cast<const uint4*>(vsrc + h * 128)[v];
    }
  }
}

// ------------------------------------------------------------------------------------------------ row sums of squares
// One warp per row: ssq[m] = sum_k x[m,k]^2 (fp32), for the folded RMSNorm of the GEMM that consumes x (omc_gemm_bf16_norm).
__global__ void __launch_bounds__(256) row_ssq_rows_kernel(const bf16* __restrict__ x, long long ldx, long long rows, int C,
                                                           float* __restrict__ ssq) {
  const int lane = threadIdx.x & 31;
  for (long long m = (long long)blockIdx.x * 8 + (threadIdx.x >> 5); m < rows; m += (long long)gridDim.x * 8) {
    const bf16* row = x + m * ldx;
    float s = 0.f;
    for (int i = lane * 8; i < C; i += 32 * 8) {
      const uint4 v = *reinterpret_cast<const uint4*>(row + i);
      const uint32_t* a = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float2 f = unpack_bf16(a[q]);
        s += f.x * f.x + f.y * f.y;
      }
    }
    s = warp_sum(s);
    if (lane == 0) ssq[m] = s;
  }
}

// ------------------------------------------------------------------------------------------------ argmax
__device__ __forceinline__ void argmax_merge(float& bv, int& bi, float v, int i) {
  if (v > bv || (v == bv && i < bi)) {
    bv = v;
    bi = i;
  }
}
constexpr int kArgmaxChunks = 64;
__global__ void __launch_bounds__(256) argmax_partial_kernel(const float* __restrict__ logits, long long ldl, int V,
                                                             float* __restrict__ pv, int* __restrict__ pi) {
  const int b = blockIdx.y;
  const float* row = logits + (long long)b * ldl;
  const int per = (V + kArgmaxChunks - 1) / kArgmaxChunks;
  const int lo = blockIdx.x * per, hi = min(V, lo + per);
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) argmax_merge(bv, bi, row[i], i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    argmax_merge(bv, bi, ov, oi);
  }
  __shared__ float sv[8];
  __shared__ int si[8];
  if ((threadIdx.x & 31) == 0) {
    sv[threadIdx.x >> 5] = bv;
    si[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) argmax_merge(bv, bi, sv[w], si[w]);
    pv[b * kArgmaxChunks + blockIdx.x] = bv;
    pi[b * kArgmaxChunks + blockIdx.x] = bi;
  }
}
__global__ void argmax_final_kernel(const float* __restrict__ pv, const int* __restrict__ pi, int64_t* __restrict__ next) {
  const int b = blockIdx.x;
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < kArgmaxChunks; i += 32) argmax_merge(bv, bi, pv[b * kArgmaxChunks + i], pi[b * kArgmaxChunks + i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    argmax_merge(bv, bi, ov, oi);
  }
  if (threadIdx.x == 0) next[b] = bi == 0x7fffffff ? 0 : (int64_t)bi;  // all-NaN row: index 0 like torch.argmax, never the sentinel
}

static inline int grid_for(long long work_items, int threads, int per_sm = 8) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = (long long)num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace omc

using namespace omc;

extern "C" int omc_row_ssq_rows(const void* x, long long ldx, long long rows, int C, float* ssq, void* stream) {
  if (rows <= 0) return OMC_OK;
  if (C % 8 != 0 || ldx % 8 != 0) return set_error(OMC_ERR_SHAPE, "omc_row_ssq_rows: C, ldx must be multiples of 8");
  row_ssq_rows_kernel<<<grid_for(rows * 32, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, rows, C, ssq);
  return check_launch("row_ssq_rows");
}

extern "C" int omc_rmsnorm(const void* x, long long ldx, const void* w, void* out, long long ldo, int rows, int C,
                           float eps, void* stream) {
  if (rows <= 0) return OMC_OK;
  if (C <= 0 || C % 8 != 0 || C > kNormThreads * kNormMaxVec * 8)
    return set_error(OMC_ERR_SHAPE, "omc_rmsnorm: C must be a multiple of 8 and <= 4096");
  if (ldx % 8 != 0 || ldo % 8 != 0) return set_error(OMC_ERR_ALIGN, "omc_rmsnorm: leading dims must be multiples of 8");
  int grid = rows < num_sms() * 16 ? rows : num_sms() * 16;
  rmsnorm_kernel<<<grid, kNormThreads, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (const bf16*)w, nullptr, (bf16*)out,
                                                                   ldo, rows, C, eps, 1);
  return check_launch("rmsnorm");
}

extern "C" int omc_layernorm(const void* x, long long ldx, const void* w, const void* b, void* out, long long ldo, int rows,
                             int C, float eps, void* stream) {
  if (rows <= 0) return OMC_OK;
  if (x == nullptr || w == nullptr || out == nullptr) return set_error(OMC_ERR_ARG, "omc_layernorm: null argument");
  if (C <= 0 || C % 8 != 0 || C > kNormThreads * kNormMaxVec * 8)
    return set_error(OMC_ERR_SHAPE, "omc_layernorm: C must be a multiple of 8 and <= 4096");
  if (ldx % 8 != 0 || ldo % 8 != 0) return set_error(OMC_ERR_ALIGN, "omc_layernorm: leading dims must be multiples of 8");
  int grid = rows < num_sms() * 16 ? rows : num_sms() * 16;
  layernorm_kernel<<<grid, kNormThreads, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (const bf16*)w, (const bf16*)b,
                                                                     (bf16*)out, ldo, rows, C, eps);
  return check_launch("layernorm");
}

extern "C" int omc_rmsnorm_pair(void* x, long long ldx, const void* w_a, const void* w_b, int rows, int C, float eps,
                                void* stream) {
  if (rows <= 0) return OMC_OK;
  if (C <= 0 || C % 8 != 0 || C > kNormThreads * kNormMaxVec * 8)
    return set_error(OMC_ERR_SHAPE, "omc_rmsnorm_pair: C must be a multiple of 8 and <= 4096");
  if (ldx % 8 != 0 || ldx < 2LL * C) return set_error(OMC_ERR_ALIGN, "omc_rmsnorm_pair: ldx must be a multiple of 8 and >= 2 C");
  const long long v = 2LL * rows;
  int grid = v < (long long)num_sms() * 16 ? (int)v : num_sms() * 16;
  rmsnorm_kernel<<<grid, kNormThreads, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (const bf16*)w_a, (const bf16*)w_b,
                                                                   (bf16*)x, ldx, rows, C, eps, 2);
  return check_launch("rmsnorm_pair");
}

extern "C" int omc_vit_im2col(const void* pixels, int pixels_are_f32, void* cols, long long ldc, int B, int H, int W,
                              void* stream) {
  if (B <= 0) return OMC_OK;
  if (H % 14 != 0 || W % 14 != 0 || ldc < 588) return set_error(OMC_ERR_SHAPE, "omc_vit_im2col: H, W must be multiples of 14 and ldc >= 588");
  long long total = (long long)B * (H / 14) * (W / 14) * ldc;
  int grid = grid_for(total, 256);
  if (pixels_are_f32)
    im2col_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)pixels, (bf16*)cols, ldc, B, H, W);
  else
    im2col_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)pixels, (bf16*)cols, ldc, B, H, W);
  return check_launch("im2col");
}

extern "C" int omc_vit_assemble(const void* patch, const void* cls, const void* pos, void* hidden, int B, int P, int C,
                                void* stream) {
  if (B <= 0) return OMC_OK;
  if (C % 8 != 0) return set_error(OMC_ERR_SHAPE, "omc_vit_assemble: C % 8 != 0");
  long long total = (long long)B * (P + 1) * (C / 8);
  vit_assemble_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)patch, (const bf16*)cls,
                                                                              (const bf16*)pos, (bf16*)hidden, B, P, C);
  return check_launch("vit_assemble");
}

extern "C" int omc_select_pixel_shuffle(const void* hidden, void* out, int B, int G, int C, int down, void* stream) {
  if (B <= 0) return OMC_OK;
  if (down < 1 || G % down != 0 || C % 8 != 0) return set_error(OMC_ERR_SHAPE, "omc_select_pixel_shuffle: bad G/C/down");
  long long total = (long long)B * G * G * (C / 8);
  select_pixel_shuffle_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)hidden, (bf16*)out,
                                                                                      B, G, C, down);
  return check_launch("select_pixel_shuffle");
}

extern "C" int omc_embed_lookup(const int64_t* ids, int T, const void* table, int C, void* out, long long ldo,
                                int vocab, void* stream) {
  if (T <= 0) return OMC_OK;
  if (C % 8 != 0 || ldo % 8 != 0) return set_error(OMC_ERR_SHAPE, "omc_embed_lookup: C, ldo must be multiples of 8");
  int grid = T < num_sms() * 8 ? T : num_sms() * 8;
  embed_lookup_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(ids, T, (const bf16*)table, C, (bf16*)out, ldo, vocab);
  return check_launch("embed_lookup");
}

extern "C" int omc_splice(const int64_t* ids, const int32_t* seq_offsets, int n_seq, int S_total, long long image_token,
                          const void* table, const void* feats, int n_img, int L, int C, int max_len, void* embeds,
                          int32_t* pos_ids, int32_t* seq_ids, int32_t* out_offsets, int32_t* workspace, int T_capacity,
                          int vocab, void* stream) {
  if (n_seq <= 0 || S_total <= 0) return set_error(OMC_ERR_SHAPE, "omc_splice: empty input");
  if (C % 8 != 0) return set_error(OMC_ERR_SHAPE, "omc_splice: C % 8 != 0");
  if (L < 1) return set_error(OMC_ERR_SHAPE, "omc_splice: L < 1");
  int32_t* gcount = workspace;                  // [S_total + 1]
  int32_t* img_base = workspace + S_total + 1;  // [n_seq + 1]
  splice_plan_kernel<<<1, kPlanThreads, 0, (cudaStream_t)stream>>>(ids, seq_offsets, n_seq, S_total, image_token, L,
                                                                   max_len, gcount, img_base, out_offsets);
  int rc = check_launch("splice_plan");
  if (rc) return rc;
  int grid = grid_for((long long)T_capacity * 32, 256);
  splice_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(ids, seq_offsets, n_seq, image_token, (const bf16*)table,
                                                               (const bf16*)feats, n_img, L, C, gcount, img_base,
                                                               out_offsets, (bf16*)embeds, pos_ids, seq_ids, T_capacity, vocab);
  return check_launch("splice_gather");
}

extern "C" int omc_rope_kv_store(void* qkv, long long ldqkv, const int32_t* pos, const int32_t* seq_ids, int T, int Hq,
                                 int Hkv, const float* inv_freq, void* kv_pool, const int32_t* block_table,
                                 int max_pages, int page_size, void* stream) {
  if (T <= 0) return OMC_OK;
  if (page_size <= 0 || max_pages <= 0) return set_error(OMC_ERR_ARG, "omc_rope_kv_store: bad paging parameters");
  int grid = T < num_sms() * 8 ? T : num_sms() * 8;
  rope_kv_store_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((bf16*)qkv, ldqkv, pos, seq_ids, T, Hq, Hkv, inv_freq,
                                                                (bf16*)kv_pool, block_table, max_pages, page_size);
  return check_launch("rope_kv_store");
}

extern "C" int omc_argmax(const float* logits, long long ldl, int B, int V, int64_t* next, float* workspace,
                          void* stream) {
  if (B <= 0 || V <= 0) return set_error(OMC_ERR_SHAPE, "omc_argmax: empty input");
  float* pv = workspace;                                          // [B * 64]
  int* pi = reinterpret_cast<int*>(workspace + (long long)B * kArgmaxChunks);  // [B * 64]
  argmax_partial_kernel<<<dim3(kArgmaxChunks, B), 256, 0, (cudaStream_t)stream>>>(logits, ldl, V, pv, pi);
  int rc = check_launch("argmax_partial");
  if (rc) return rc;
  argmax_final_kernel<<<B, 32, 0, (cudaStream_t)stream>>>(pv, pi, next);
  return check_launch("argmax_final");
}
