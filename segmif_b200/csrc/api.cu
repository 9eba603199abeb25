// Error plumbing, ABI version and device check for libsegmif_b200.so.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace segmif {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return SEGMIF_ERR_CUDA;
  }
  return SEGMIF_OK;
}

}  // namespace segmif

extern "C" int segmif_abi_version(void) { return SEGMIF_ABI_VERSION; }

extern "C" const char* segmif_last_error(void) { return segmif::g_err; }

extern "C" int segmif_init(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    segmif::set_error("segmif_init: no CUDA device visible (%s)", cudaGetErrorString(e));
    return SEGMIF_ERR_CUDA;
  }
  if (device < 0 || device >= n) {
    segmif::set_error("segmif_init: device %d out of range (0..%d)", device, n - 1);
    return SEGMIF_ERR_INVALID;
  }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    segmif::set_error("segmif_init: cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    return SEGMIF_ERR_CUDA;
  }
  if (prop.major != 10) {
    segmif::set_error("segmif_init: device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major,
                      prop.minor);
    return SEGMIF_ERR_DEVICE;
  }
  return SEGMIF_OK;
}
