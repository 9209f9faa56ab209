// Library-level entry points: version, error reporting, device check.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace skge {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
  const char *base = strrchr(file, '/');
  set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), base ? base + 1 : file,
            line, what);
  return -(int)e;
}

}  // namespace skge

extern "C" {

int skge_version(void) { return 100; }

const char *skge_last_error(void) { return skge::g_err; }

int skge_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    skge::set_error("no CUDA device: %s (libskge_b200 has no CPU fallback)", cudaGetErrorString(e));
    cudaGetLastError();
    return SKGE_ENODEVICE;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    skge::set_error("device %d is sm_%d%d; libskge_b200 is built for sm_100a only", dev, major, minor);
    return SKGE_ENODEVICE;
  }
  return 0;
}

}  // extern "C"
