// Shared helpers for the sm_100a kernels of libptt_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ptt_b200.h"

#define PTT_CHECK_ARG(cond)                      \
  do {                                           \
    if (!(cond)) return PTT_ERR_INVALID_ARGUMENT; \
  } while (0)

// The library's fault word (lib.cu): raised by a kernel whose bounded barrier wait expired (tc_common.cuh).
unsigned int* ptt_fault_word();     // host-mapped, device-visible; nullptr if it cannot be allocated
bool ptt_fault_pending();           // plain host read, no CUDA call

// Returns the launch error (if any) of the kernels enqueued so far by this call, without syncing; a device fault
// reported by an EARLIER kernel is sticky (PTT_ERR_DEVICE_FAULT until ptt_fault_clear()).
static inline int ptt_launch_status() {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  return ptt_fault_pending() ? PTT_ERR_DEVICE_FAULT : PTT_OK;
}

// Kernel-launch counter behind ptt_launch_count() (bench.py reports it as `gpu_launches`).
void ptt_count_launches(int n);
#define PTT_LAUNCHED() ptt_count_launches(1)

static inline cudaStream_t as_stream(ptt_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int round_up(int x, int a) { return (x + a - 1) / a * a; }
static inline long long llmin_(long long a, long long b) { return a < b ? a : b; }

// Function attributes (opt-in shared memory, cluster occupancy) belong to a device context: launch helpers keep their
// "configured" state per device so that one process may drive several GPUs.
constexpr int PTT_MAX_DEVICES = 64;
static inline int ptt_current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= PTT_MAX_DEVICES) dev = 0;
  return dev;
}

// a*a + b*b + c*c exactly as nvcc's default contraction emits it for upstream pointnet2_ops
// (oracle/probe_contraction.sh): FMUL on the middle term, then two FFMAs.  Written with intrinsics
// so the result does not depend on how the compiler feels about this translation unit.
__device__ __forceinline__ float sq3(float a, float b, float c) {
  return __fmaf_rn(c, c, __fmaf_rn(a, a, __fmul_rn(b, b)));
}

// layer_utils.square_distance (layer_utils.py:26): ((dx*dx + dy*dy) + dz*dz), every step rounded.
__device__ __forceinline__ float sq3_nofma(float a, float b, float c) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
}

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
