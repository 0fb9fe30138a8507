// GPU any-resolution image preprocessing: the step in front of the vision tower.
//   reference: process_anyres_image omchat/mm_utils.py:119-158 = PIL Image.resize (BICUBIC) of the original to the best grid
//   resolution, paste on a black canvas (resize_and_pad_image :43-73), split into 448x448 patches (divide_to_patches :76-94),
//   plus the whole image resized to 448x448, each crop through CLIPImageProcessor (rescale 1/255, ImageNet normalise;
//   internVIT_encoder.py:26-29).
// The resampling arithmetic is Pillow's (src/libImaging/Resample.c, 8 bits per channel): per output column/row a window of
// 22-bit fixed-point weights (computed by the HOST exactly like precompute_coeffs + normalize_coeffs_8bpc), accumulator
// seeded with 1 << 21, arithmetic shift by 22, clip to 0..255, horizontal pass first with a uint8 intermediate — integer
// work, bit-exact against Pillow. Normalisation is a 3 x 256 fp32 look-up table built by the host with the reference's own
// formulas, so the float output is exact by construction. These kernels are HBM-bound byte shuffles: one thread per output
// pixel (3 channels), coalesced along x; nothing here is GEMM-shaped.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "omc_internal.h"

namespace omc {

constexpr int kPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// out[y, xx, c] = clip8(2^21 + sum_k src[y, xmin(xx) + k, c] * coef[xx, k])                (vertical = 0)
// out[yy, x, c] = clip8(2^21 + sum_k src[ymin(yy) + k, x, c] * coef[yy, k])                (vertical = 1)
__global__ void resample_u8_kernel(const uint8_t* __restrict__ src, int src_h, int src_w, uint8_t* __restrict__ dst,
                                   int dst_h, int dst_w, const int32_t* __restrict__ coefs,
                                   const int32_t* __restrict__ bounds, int ksize, int vertical) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= dst_w || y >= dst_h) return;
  const int o = vertical ? y : x;
  const int lo = __ldg(bounds + 2 * o), n = __ldg(bounds + 2 * o + 1);
  const int32_t* k = coefs + (long long)o * ksize;
  int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
  if (vertical) {
    const uint8_t* p = src + ((long long)lo * src_w + x) * 3;
    const long long pitch = (long long)src_w * 3;
    for (int i = 0; i < n; ++i, p += pitch) {
      const int w = __ldg(k + i);
      a0 += p[0] * w;
      a1 += p[1] * w;
      a2 += p[2] * w;
    }
  } else {
    const uint8_t* p = src + ((long long)y * src_w + lo) * 3;
    for (int i = 0; i < n; ++i, p += 3) {
      const int w = __ldg(k + i);
      a0 += p[0] * w;
      a1 += p[1] * w;
      a2 += p[2] * w;
    }
  }
  uint8_t* q = dst + ((long long)y * dst_w + x) * 3;
  q[0] = clip8(a0);
  q[1] = clip8(a1);
  q[2] = clip8(a2);
}

// crops [1 + (target_h / crop) * (target_w / crop), 3, crop, crop]: crop 0 = the whole image resized to crop x crop, then the
// patches of the black canvas on which the aspect-preserving resize sits at (paste_x, paste_y), row-major.
template <typename OutT>
__global__ void anyres_pack_kernel(const uint8_t* __restrict__ thumb, const uint8_t* __restrict__ resized, int new_w,
                                   int new_h, int target_w, int paste_x, int paste_y, int crop,
                                   const float* __restrict__ lut, OutT* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, n = blockIdx.z;
  if (x >= crop) return;
  uint8_t r = 0, g = 0, b = 0;
  if (n == 0) {
    const uint8_t* p = thumb + ((long long)y * crop + x) * 3;
    r = p[0]; g = p[1]; b = p[2];
  } else {
    const int per_row = target_w / crop;
    const int cy = ((n - 1) / per_row) * crop + y - paste_y, cx = ((n - 1) % per_row) * crop + x - paste_x;
    if (cy >= 0 && cy < new_h && cx >= 0 && cx < new_w) {
      const uint8_t* p = resized + ((long long)cy * new_w + cx) * 3;
      r = p[0]; g = p[1]; b = p[2];
    }
  }
  const long long plane = (long long)crop * crop;
  OutT* o = out + (long long)n * 3 * plane + (long long)y * crop + x;
  o[0] = (OutT)__ldg(lut + r);
  o[plane] = (OutT)__ldg(lut + 256 + g);
  o[2 * plane] = (OutT)__ldg(lut + 512 + b);
}

}  // namespace omc

using namespace omc;

extern "C" int omc_resample_u8(const void* src, int src_h, int src_w, void* dst, int dst_h, int dst_w, const int32_t* coefs,
                               const int32_t* bounds, int ksize, int vertical, void* stream) {
  if (src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0 || ksize <= 0)
    return set_error(OMC_ERR_SHAPE, "omc_resample_u8: empty image or window");
  if (vertical ? (src_w != dst_w) : (src_h != dst_h))
    return set_error(OMC_ERR_SHAPE, "omc_resample_u8: the pass changes one dimension only");
  if (dst_h > 65535) return set_error(OMC_ERR_SHAPE, "omc_resample_u8: more than 65535 rows");
  dim3 grid((dst_w + 127) / 128, dst_h);
  resample_u8_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const uint8_t*)src, src_h, src_w, (uint8_t*)dst, dst_h, dst_w,
                                                             coefs, bounds, ksize, vertical);
  return check_launch("resample_u8");
}

extern "C" int omc_anyres_pack(const void* thumb, const void* resized, int new_w, int new_h, int target_w, int target_h,
                               int paste_x, int paste_y, int crop, const float* lut, void* out, int out_is_bf16, void* stream) {
  if (crop <= 0 || target_w % crop != 0 || target_h % crop != 0 || new_w > target_w || new_h > target_h || new_w <= 0 ||
      new_h <= 0)
    return set_error(OMC_ERR_SHAPE, "omc_anyres_pack: the canvas must be a multiple of the crop and contain the resized image");
  const int n = 1 + (target_w / crop) * (target_h / crop);
  dim3 grid((crop + 127) / 128, crop, n);
  if (out_is_bf16)
    anyres_pack_kernel<__nv_bfloat16><<<grid, 128, 0, (cudaStream_t)stream>>>(
        (const uint8_t*)thumb, (const uint8_t*)resized, new_w, new_h, target_w, paste_x, paste_y, crop, lut, (__nv_bfloat16*)out);
  else
    anyres_pack_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>((const uint8_t*)thumb, (const uint8_t*)resized, new_w,
                                                                      new_h, target_w, paste_x, paste_y, crop, lut, (float*)out);
  return check_launch("anyres_pack");
}
