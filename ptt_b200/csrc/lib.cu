// Version / error text of libptt_b200.so (include/ptt_b200.h).
#include <atomic>

#include "common.cuh"

static std::atomic<unsigned long long> g_launches{0};

void ptt_count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

extern "C" unsigned long long ptt_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" const char* ptt_version(void) { return "ptt_b200 0.1.0 sm_100a"; }

extern "C" const char* ptt_error_string(int code) {
  switch (code) {
    case PTT_OK: return "ok";
    case PTT_ERR_INVALID_ARGUMENT: return "invalid argument (null pointer, negative size, or sizes that contradict each other)";
    case PTT_ERR_UNSUPPORTED: return "shape outside what this kernel family covers";
    case PTT_ERR_WORKSPACE: return "workspace missing or too small (query the matching *_workspace_bytes)";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown ptt_b200 error";
}
