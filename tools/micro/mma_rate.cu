// Microbenchmark: cycles per tcgen05.mma (cta_group::1, M=128, K=16, bf16) for the operand forms the attention kernel uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I omchat_b200/csrc -o gpurun_out/mma_rate tools/micro/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace omc;

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int variant, int iters) {
  extern __shared__ uint8_t raw[];
  const uint32_t a0 = smem_u32(raw);
  uint8_t* smem = raw + (((a0 + 1023u) & ~1023u) - a0);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (warp == 0) {
    if (elect_one()) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc<1>(&slot, 512);
    tmem_relinquish<1>();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t sa = smem_u32(smem), sb = sa + 32768;
    const int N = (variant == 1) ? 64 : 128;
    const bool ts = (variant == 2 || variant == 4), mn = (variant == 2 || variant == 3);
    const uint32_t idesc = make_idesc_bf16_major(128, N, 0, mn ? 1 : 0);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint64_t da = make_sw128_kmajor_desc(sa + (kk >> 2) * 16384) + (uint64_t)(2 * (kk & 3));
        const uint64_t db = mn ? make_sw128_mnmajor_desc(sb + kk * 2048, 16384, 1024)
                               : make_sw128_kmajor_desc(sb + (kk >> 2) * 16384) + (uint64_t)(2 * (kk & 3));
        if (ts) umma_bf16_ts(tm + 256, tm + kk * 8, db, idesc, 1u);
        else umma_bf16<1>(tm + 256, da, db, idesc, 1u);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[variant] = (t1 - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(tm, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const char* names[5] = {"SS  N=128 B K-major ", "SS  N=64  B K-major ", "TS  N=128 B MN-major", "SS  N=128 B MN-major", "TS  N=128 B K-major "};
  const int iters = 512;
  for (int rep = 0; rep < 2; ++rep)
    for (int v = 0; v < 5; ++v) {
      mma_rate_kernel<<<1, 128, 100 * 1024>>>(d, v, iters);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("variant %d: %s\n", v, cudaGetErrorString(e)); return 1; }
    }
  long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  for (int v = 0; v < 5; ++v) printf("%s: %.1f cycles per MMA (ideal %d)\n", names[v], (double)h[v] / (iters * 8), v == 1 ? 32 : 64);
  return 0;
}
