// segb200 -- C-ABI plumbing: version, thread-local error message, launch check.
#include "common.cuh"
#include "../../include/segb200.h"

#include <stdarg.h>

namespace segb200 {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error((int)e, "%s: %s", what, cudaGetErrorString(e));
  return 0;
}

}  // namespace segb200

extern "C" int segb200_version(void) { return SEGB200_VERSION; }
extern "C" const char* segb200_last_error(void) { return segb200::g_err; }
