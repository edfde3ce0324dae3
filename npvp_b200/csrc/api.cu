// Error plumbing and bookkeeping shared by every extern "C" entry point.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";
int g_launch_count = 0;

void npvp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int npvp_check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    npvp_set_error("%s: %s", what, cudaGetErrorString(e));
    return NPVP_ERR_CUDA;
  }
  return NPVP_OK;
}

extern "C" const char* npvp_last_error(void) { return g_err; }
extern "C" int npvp_version(void) { return 100; }
extern "C" int64_t npvp_launch_count(void) { return g_launch_count; }
extern "C" void npvp_reset_launch_count(void) { g_launch_count = 0; }
