// Version / error text of libptt_b200.so (include/ptt_b200.h).
#include <atomic>

#include "common.cuh"

static std::atomic<unsigned long long> g_launches{0};

void ptt_count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

extern "C" unsigned long long ptt_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ---- the fault word (tc_common.cuh: bounded barrier waits).  One host-mapped, portable word per process, allocated the
// first time a tensor-core kernel family is configured (warm-up, never inside a stream capture); with unified
// addressing the host pointer is the device pointer on every GPU.
static std::atomic<unsigned int*> g_fault_word{nullptr};

unsigned int* ptt_fault_word() {
  unsigned int* w = g_fault_word.load(std::memory_order_acquire);
  if (w != nullptr) return w;
  void* p = nullptr;
  if (cudaHostAlloc(&p, 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  *static_cast<volatile unsigned int*>(p) = 0u;
  unsigned int* expected = nullptr;
  if (!g_fault_word.compare_exchange_strong(expected, static_cast<unsigned int*>(p), std::memory_order_acq_rel)) {
    cudaFreeHost(p);               // another thread won the race
    return expected;
  }
  return static_cast<unsigned int*>(p);
}

bool ptt_fault_pending() {
  const unsigned int* w = g_fault_word.load(std::memory_order_acquire);
  return w != nullptr && *reinterpret_cast<const volatile unsigned int*>(w) != 0u;
}

extern "C" int ptt_fault_status(void) { return ptt_fault_pending() ? PTT_ERR_DEVICE_FAULT : PTT_OK; }

extern "C" void ptt_fault_clear(void) {
  unsigned int* w = g_fault_word.load(std::memory_order_acquire);
  if (w != nullptr) *reinterpret_cast<volatile unsigned int*>(w) = 0u;
}

extern "C" const char* ptt_version(void) { return "ptt_b200 0.2.0 sm_100a"; }

extern "C" const char* ptt_error_string(int code) {
  switch (code) {
    case PTT_OK: return "ok";
    case PTT_ERR_INVALID_ARGUMENT: return "invalid argument (null pointer, negative size, or sizes that contradict each other)";
    case PTT_ERR_UNSUPPORTED: return "shape outside what this kernel family covers";
    case PTT_ERR_WORKSPACE: return "workspace missing or too small (query the matching *_workspace_bytes)";
    case PTT_ERR_DEVICE_FAULT: return "a kernel of this library gave up a bounded barrier wait earlier (its results and everything "
                                      "enqueued after it are invalid): synchronise, then ptt_fault_clear()";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown ptt_b200 error";
}
