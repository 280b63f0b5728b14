/*
 * ptt_b200_tuning.h -- test / tuning surface of libptt_b200.so.  NOT part of the drop-in boundary.
 *
 * Everything here is either a PROCESS-WIDE switch (not thread-safe, affects every later call on every
 * stream) or an entry point that runs one kernel variant the product dispatcher would not pick.  The
 * parity tests use them to run every variant of a kernel family against the oracle
 * (tests/test_gpu_ops.py, tests/test_gpu_modules.py) and the tools/ scripts to record timelines.
 * Product code (the Python files of ptt_b200) never calls them; ptt_b200.h's contract holds as long as they are left alone.
 */
#ifndef PTT_B200_TUNING_H_
#define PTT_B200_TUNING_H_

#include "ptt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* 1 = every dense contraction runs on the exact-fp32 CUDA-core path (the on-device cross-check of the tcgen05 path). */
PTT_API void ptt_debug_force_ffma(int on);

/* Transformer pair-row passes: 0 = automatic (CTA pairs at d_model >= 256, else single CTAs); 1, 2, 4 = single-CTA
 * MMAs with the weight ring multicast over a cluster of that size; -2 = CTA pairs; -1 = the unfused generic path. */
PTT_API void ptt_debug_set_cluster(int cluster_size);

/* Fused SA kernel: clock64 timeline buffer picked up by the next launches (NULL = off). */
PTT_API void ptt_debug_sa_timeline(long long* device_buf);

/* One pair-row pass of the transformer block with a timeline buffer (>= 6000 int64).  mode 0 / 1 / 2 = pass 1 / 2 / 3. */
PTT_API int ptt_debug_tr_pass_timeline(const float* xyz, const int* knn, int B, int n, int k, int dm, const float* wd0,
                                       int ldw0, const void* wimg, const float* bias, float* out, long long* dbg,
                                       int flags, void* stream, int mode, const float* aux, const float* big);

/* FPS with an explicit (threads per cloud, points per thread) instead of the dispatcher's choice; threads * ppt >= N. */
PTT_API int ptt_fps_variant(const float* xyz, int B, int N, int npoint, int* idx, float* new_xyz, int threads, int ppt,
                            ptt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PTT_B200_TUNING_H_ */
