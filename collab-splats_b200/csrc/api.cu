// Library-level entry points of the C-ABI (see include/rade_b200.h).
#include "common.cuh"

static thread_local int g_last_cuda_error = 0;

extern "C" void rs_set_last_cuda_error(int code) { g_last_cuda_error = code; }

// number of kernel launches issued through this library (bench.py reports it as gpu_launches)
static unsigned long long g_launches = 0;
extern "C" void rs_count_launches(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
extern "C" unsigned long long rs_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int rs_version(void) { return 100; }  // 0.1.0

extern "C" int rs_last_cuda_error(void) { return g_last_cuda_error; }

extern "C" const char* rs_error_string(int status) {
  switch (status) {
    case RS_OK: return "ok";
    case RS_ERR_BAD_ARG: return "bad argument (null pointer, negative size or unsupported shape)";
    case RS_ERR_LAUNCH: return cudaGetErrorString((cudaError_t)g_last_cuda_error);
    case RS_ERR_UNSUPPORTED: return "unsupported configuration (see include/rade_b200.h)";
    default: return "unknown status";
  }
}
