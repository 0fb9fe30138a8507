// Internal shared declarations for the omchat_b200 C-ABI library (error codes, epilogue ids, helpers).
#pragma once
#include <cuda_runtime.h>

#include "../../include/omchat_b200.h"

namespace omc {

enum { EPI_NONE = OMC_EPI_NONE, EPI_GELU = OMC_EPI_GELU, EPI_RES = OMC_EPI_RES, EPI_SWIGLU = OMC_EPI_SWIGLU };

// records the message for omc_last_error() and returns `code`
int set_error(int code, const char* msg);
int num_sms();

// Ordinal of the calling thread's current device, clamped to [0, kMaxDevices): index of the per-device caches below.
// cudaFuncSetAttribute and the SM count are per device, so "done once" flags must be kept per ordinal (a process may
// drive several GPUs: OmChatQwen2ForCausalLM(cfg, device="cuda:1")).
constexpr int kMaxDevices = 16;
inline int cur_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    dev = 0;
  }
  return dev < 0 ? 0 : (dev >= kMaxDevices ? kMaxDevices - 1 : dev);
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error(OMC_ERR_CUDA, cudaGetErrorString(e));
    (void)what;
    return OMC_ERR_CUDA;
  }
  return OMC_OK;
}

}  // namespace omc
