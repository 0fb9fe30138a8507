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
