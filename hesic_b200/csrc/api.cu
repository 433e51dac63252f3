// Library-level entry points: version, error reporting, device check, launch counter.
#include <stdarg.h>

#include "common.cuh"

namespace hesic {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace hesic

extern "C" int hesic_abi_version(void) { return HESIC_ABI_VERSION; }
extern "C" const char *hesic_last_error(void) { return hesic::g_err; }

extern "C" int64_t hesic_launch_count(int reset) {
  return reset ? hesic::g_launches.exchange(0) : hesic::g_launches.load();
}

extern "C" int hesic_device_check(char *name, int name_len) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return hesic::cuda_fail(e, "cudaGetDevice");
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) return hesic::cuda_fail(e, "cudaGetDeviceProperties");
  if (name && name_len > 0) snprintf(name, name_len, "%s", p.name);
  if (p.major != 10) {
    hesic::set_error("hesic_b200 is built for sm_100a only; device '%s' is sm_%d%d", p.name, p.major, p.minor);
    return HESIC_E_CUDA;
  }
  return HESIC_OK;
}
