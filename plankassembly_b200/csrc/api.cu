// C-ABI plumbing: version, device check, thread-local error string.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void pa_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int pa_abi_version(void) { return PA_ABI_VERSION; }
extern "C" const char* pa_last_error(void) { return g_err; }

extern "C" int pa_device_ok(void) {
  int dev = 0;
  cudaDeviceProp p;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
    pa_set_error("no CUDA device");
    return 0;
  }
  if (p.major != 10) {
    pa_set_error("plank_b200 targets sm_100a only; found sm_%d%d", p.major, p.minor);
    return 0;
  }
  return 1;
}

// SM count of the current device, cached per device (grids are sized in multiples of it; B200 = 148)
int pa_num_sms() {
  static int cache[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMs;
  if (cache[dev] == 0) {
    int n = 0;
    cache[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : kNumSMs;
  }
  return cache[dev];
}
