// C-ABI plumbing: error reporting, version, device queries.
#include <cuda_runtime.h>
#include <string.h>

#include "omc_internal.h"

namespace omc {

static thread_local char g_err[512] = "";

int set_error(int code, const char* msg) {
  strncpy(g_err, msg ? msg : "", sizeof(g_err) - 1);
  g_err[sizeof(g_err) - 1] = 0;
  return code;
}

}  // namespace omc

extern "C" const char* omc_last_error(void) { return omc::g_err; }
extern "C" int omc_version(void) { return 100; }
extern "C" int omc_num_sms(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    return omc::set_error(OMC_ERR_CUDA, "no CUDA device");
  }
  return n;
}

// ---- peer memory (CUDA IPC) for the tensor-parallel persistent decode kernel
extern "C" int omc_peer_alloc(long long bytes, void** ptr, void* handle64) {
  if (bytes <= 0 || ptr == nullptr || handle64 == nullptr) return omc::set_error(OMC_ERR_ARG, "omc_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    if (p) cudaFree(p);
    cudaGetLastError();
    return omc::set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return OMC_OK;
}
extern "C" int omc_peer_open(const void* handle64, void** ptr) {
  if (handle64 == nullptr || ptr == nullptr) return omc::set_error(OMC_ERR_ARG, "omc_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return omc::set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
  }
  *ptr = p;
  return OMC_OK;
}
extern "C" int omc_peer_close(void* ptr) {
  if (ptr == nullptr) return OMC_OK;
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return omc::set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
  }
  return OMC_OK;
}
extern "C" int omc_peer_free(void* ptr) {
  if (ptr == nullptr) return OMC_OK;
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return omc::set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
  }
  return OMC_OK;
}
