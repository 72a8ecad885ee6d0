// Library-level entry points of libmoda_b200.so: version, error text, device probe.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace moda {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

}  // namespace moda

extern "C" const char* moda_last_error() { return moda::g_err; }

extern "C" const char* moda_version() { return "moda_b200 0.1 (sm_100a)"; }

// 0 when the current device can run this library (compute capability 10.x), else an error code.
extern "C" int moda_device_check() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { moda::set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return (int)e; }
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) { moda::set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return (int)e; }
  if (p.major != 10) {
    moda::set_error("device %s is sm_%d%d; libmoda_b200 is built for sm_100a only", p.name, p.major, p.minor);
    return -2;
  }
  return 0;
}
