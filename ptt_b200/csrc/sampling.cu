// Furthest-point sampling and point gather for sm_100a.
//
// Replaces pointnet2_ops `_ext.furthest_point_sampling`, `_ext.furthest_point_sampling_with_dist`,
// `_ext.gather_points(_grad)` (reference call sites pointnet2_utils.py:48,78,112,118).
//
// FPS is a serial chain of npoint-1 dependent arg-max steps over one small cloud (12 KB at N=1024):
// latency-bound, neither HBM- nor tensor-bound.  Design: one CTA per cloud, the cloud's points and
// their running min-distances live in REGISTERS (PPT points per thread), the arg-max is two
// `redux.sync` per warp plus one shared-memory exchange and ONE __syncthreads per step
// (double-buffered slots), and the winner's coordinates come from a shared-memory float4 tile.
//
// Tie-breaking is part of the contract (real clouds hold exact duplicates).
// Upstream's kernel runs bs = min(512, 2^floor(log2 N)) threads; thread t scans k = t, t+bs, ...
// with strict '>' and the shared-memory tree keeps the LOWER slot on ties.  The winner among equal
// values is therefore the point minimising  key(k) = (bitreverse_log2(bs)(k mod bs), k div bs)
// lexicographically.  That closed form lets this kernel use any thread mapping: it reduces
// (value, key) pairs, max value first, min key second.  tests/ checks it against the oracle, which
// simulates the upstream tree literally.
#include "common.cuh"
#include "ptt_b200_tuning.h"

namespace {

__host__ __device__ __forceinline__ int upstream_block_log2(int n) {
  int l = 0;
  while ((2 << l) <= n && (2 << l) <= 512) ++l;
  return l;  // bs = 1 << l
}

__device__ __forceinline__ unsigned tie_key(int k, int L) {
  const unsigned low = (unsigned)k & ((1u << L) - 1u);
  const unsigned rev = L ? (__brev(low) >> (32 - L)) : 0u;
  return (rev << 20) | ((unsigned)k >> L);
}

__device__ __forceinline__ int tie_key_decode(unsigned key, int L) {
  const unsigned rev = key >> 20;
  const unsigned low = L ? (__brev(rev) >> (32 - L)) : 0u;
  return (int)(((key & 0xFFFFFu) << L) | low);
}

// (value, inverted key) packed so that a signed 64-bit max picks max value, then min key.
// Values are >= +0 or exactly -1.0f, so their bit patterns order correctly as signed ints.
__device__ __forceinline__ long long pack_candidate(int value_bits, unsigned inv_key) {
  return (long long)(((unsigned long long)(unsigned)value_bits << 32) | inv_key);
}

template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS) fps_kernel(const float* __restrict__ xyz, int N, int M,
                                                       int* __restrict__ idx_out,
                                                       float* __restrict__ new_xyz_out) {
  constexpr int WARPS = THREADS / 32;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_pts = reinterpret_cast<float4*>(smem_raw);            // N points
  int* s_idx = reinterpret_cast<int*>(s_pts + N);                 // M selected indices
  __shared__ __align__(16) long long s_red[2][WARPS > 1 ? WARPS : 1];

  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const float* P = xyz + (size_t)b * N * 3;
  const int L = upstream_block_log2(N);

  float px[PPT], py[PPT], pz[PPT], mind[PPT];
  unsigned inv[PPT];
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int k = tid + i * THREADS;
    if (k < N) {
      const float x = P[3 * k + 0], y = P[3 * k + 1], z = P[3 * k + 2];
      px[i] = x; py[i] = y; pz[i] = z;
      s_pts[k] = make_float4(x, y, z, 0.f);
      // upstream: `if (mag <= 1e-3) continue;` -- a float compared with a double literal
      const bool valid = !((double)sq3(x, y, z) <= 1e-3);
      mind[i] = valid ? 1e10f : -1.0f;  // fminf(d, -1) stays -1: skipped points never update
      inv[i] = ~tie_key(k, L);
    } else {
      px[i] = py[i] = pz[i] = 0.f;
      mind[i] = -1.0f;
      inv[i] = 0u;
    }
  }
  if (tid == 0) s_idx[0] = 0;
  __syncthreads();

  int old = 0;
  for (int j = 1; j < M; ++j) {
    const float4 o = s_pts[old];
    float mx = -1.0f;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const float d = sq3(px[i] - o.x, py[i] - o.y, pz[i] - o.z);
      const float m = fminf(d, mind[i]);
      mind[i] = m;
      mx = fmaxf(mx, m);
    }
    unsigned binv = 0u;
#pragma unroll
    for (int i = 0; i < PPT; ++i) binv = (mind[i] == mx) ? max(binv, inv[i]) : binv;

    const int vb = __float_as_int(mx);
    const int wv = __reduce_max_sync(0xffffffffu, vb);
    const unsigned wi = __reduce_max_sync(0xffffffffu, vb == wv ? binv : 0u);
    long long best;
    if (WARPS > 1) {
      const int par = j & 1;
      if ((tid & 31) == 0) s_red[par][tid >> 5] = pack_candidate(wv, wi);
      __syncthreads();
      best = s_red[par][0];
#pragma unroll
      for (int w = 1; w < WARPS; ++w) {
        const long long c = s_red[par][w];
        best = c > best ? c : best;
      }
    } else {
      best = pack_candidate(wv, wi);
    }
    old = tie_key_decode(~(unsigned)(best & 0xffffffffLL), L);
    if (tid == 0) s_idx[j] = old;
  }
  __syncthreads();

  int* out = idx_out + (size_t)b * M;
  for (int j = tid; j < M; j += THREADS) out[j] = s_idx[j];
  if (new_xyz_out != nullptr) {
    float* nx = new_xyz_out + (size_t)b * M * 3;
    for (int e = tid; e < M * 3; e += THREADS) {
      const int j = e / 3, c = e - 3 * j;
      const float4 p = s_pts[s_idx[j]];
      nx[e] = c == 0 ? p.x : (c == 1 ? p.y : p.z);
    }
  }
}

// Any N: min-distances in global scratch, points re-read through L1/L2.  `dist` != nullptr selects
// the furthest_point_sampling_with_dist flavour (row `old` of a precomputed (N,N) matrix, no skip).
__global__ void __launch_bounds__(512) fps_generic_kernel(const float* __restrict__ xyz,
                                                           const float* __restrict__ dist, int N, int M,
                                                           float* __restrict__ temp, int* __restrict__ idx_out,
                                                           float* __restrict__ new_xyz_out) {
  constexpr int THREADS = 512, WARPS = 16;
  __shared__ __align__(16) long long s_red[2][WARPS];
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const float* P = xyz ? xyz + (size_t)b * N * 3 : nullptr;
  const float* D = dist ? dist + (size_t)b * N * N : nullptr;
  float* T = temp + (size_t)b * N;
  int* out = idx_out + (size_t)b * M;
  const int L = upstream_block_log2(N);

  for (int k = tid; k < N; k += THREADS) {
    bool valid = true;
    if (P) valid = !((double)sq3(P[3 * k], P[3 * k + 1], P[3 * k + 2]) <= 1e-3);
    T[k] = valid ? 1e10f : -1.0f;
  }
  if (tid == 0) out[0] = 0;
  __syncthreads();

  int old = 0;
  for (int j = 1; j < M; ++j) {
    float ox = 0.f, oy = 0.f, oz = 0.f;
    if (P) { ox = P[3 * old]; oy = P[3 * old + 1]; oz = P[3 * old + 2]; }
    long long mine = pack_candidate(__float_as_int(-1.0f), ~0u);  // (-1, k = 0)
    for (int k = tid; k < N; k += THREADS) {
      const float d = P ? sq3(P[3 * k] - ox, P[3 * k + 1] - oy, P[3 * k + 2] - oz) : D[(size_t)old * N + k];
      const float m = fminf(d, T[k]);
      T[k] = m;
      const long long c = pack_candidate(__float_as_int(m), ~tie_key(k, L));
      mine = c > mine ? c : mine;
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const long long c = __shfl_xor_sync(0xffffffffu, mine, s);
      mine = c > mine ? c : mine;
    }
    const int par = j & 1;
    if ((tid & 31) == 0) s_red[par][tid >> 5] = mine;
    __syncthreads();
    long long best = s_red[par][0];
#pragma unroll
    for (int w = 1; w < WARPS; ++w) {
      const long long c = s_red[par][w];
      best = c > best ? c : best;
    }
    old = tie_key_decode(~(unsigned)(best & 0xffffffffLL), L);
    if (tid == 0) out[j] = old;
  }
  __syncthreads();
  if (new_xyz_out != nullptr && P) {
    float* nx = new_xyz_out + (size_t)b * M * 3;
    for (int e = tid; e < M * 3; e += THREADS) nx[e] = P[3 * out[e / 3] + e % 3];
  }
}

// out[b,c,j] = points[b,c,idx[b,j]]      grid (ceil(M/256), ceil(C/8), B)
__global__ void gather_points_kernel(const float* __restrict__ points, const int* __restrict__ idx, int C,
                                     int N, int M, float* __restrict__ out) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int i = idx[(size_t)b * M + j];
  const int c0 = blockIdx.y * 8;
#pragma unroll
  for (int c = c0; c < c0 + 8; ++c)
    if (c < C) out[((size_t)b * C + c) * M + j] = __ldg(points + ((size_t)b * C + c) * N + i);
}

__global__ void gather_points_grad_kernel(const float* __restrict__ grad_out, const int* __restrict__ idx,
                                          int C, int N, int M, float* __restrict__ grad_points) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= M) return;
  const int i = idx[(size_t)b * M + j];
  const int c0 = blockIdx.y * 8;
#pragma unroll
  for (int c = c0; c < c0 + 8; ++c)
    if (c < C) atomicAdd(grad_points + ((size_t)b * C + c) * N + i, grad_out[((size_t)b * C + c) * M + j]);
}

template <int THREADS, int PPT>
int launch_fps(const float* xyz, int B, int N, int M, int* idx, float* new_xyz, cudaStream_t st) {
  const size_t smem = (size_t)N * sizeof(float4) + (size_t)M * sizeof(int);
  auto kern = fps_kernel<THREADS, PPT>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<B, THREADS, smem, st>>>(xyz, N, M, idx, new_xyz); PTT_LAUNCHED();
  return ptt_launch_status();
}

}  // namespace

extern "C" size_t ptt_furthest_point_sampling_workspace_bytes(int B, int N, int npoint) {
  (void)npoint;
  return N > 8192 && B > 0 ? (size_t)B * N * sizeof(float) : 0;
}

// Tuning hook (declared in include/ptt_b200_tuning.h): run FPS with an explicit (threads, points/thread).
extern "C" __attribute__((visibility("default"))) int ptt_fps_variant(const float* xyz, int B, int N, int npoint, int* idx, float* new_xyz,
                               int threads, int ppt, ptt_stream_t stream) {
  cudaStream_t st = as_stream(stream);
  if ((long long)threads * ppt < N) return PTT_ERR_UNSUPPORTED;
#define PTT_FPS_CASE(T, P) \
  if (threads == T && ppt == P) return launch_fps<T, P>(xyz, B, N, npoint, idx, new_xyz, st);
  PTT_FPS_CASE(32, 1) PTT_FPS_CASE(32, 2) PTT_FPS_CASE(32, 4) PTT_FPS_CASE(32, 8) PTT_FPS_CASE(32, 16)
  PTT_FPS_CASE(64, 2) PTT_FPS_CASE(64, 4) PTT_FPS_CASE(64, 8) PTT_FPS_CASE(64, 16)
  PTT_FPS_CASE(128, 2) PTT_FPS_CASE(128, 4) PTT_FPS_CASE(128, 8) PTT_FPS_CASE(128, 16)
  PTT_FPS_CASE(256, 2) PTT_FPS_CASE(256, 4) PTT_FPS_CASE(256, 8) PTT_FPS_CASE(256, 16)
  PTT_FPS_CASE(512, 2) PTT_FPS_CASE(512, 4) PTT_FPS_CASE(512, 8) PTT_FPS_CASE(512, 16)
#undef PTT_FPS_CASE
  return PTT_ERR_UNSUPPORTED;
}

extern "C" int ptt_furthest_point_sampling(const float* xyz, int B, int N, int npoint, int* idx,
                                           float* new_xyz, void* workspace, size_t workspace_bytes,
                                           ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && N >= 1 && npoint >= 0);
  if (B == 0 || npoint == 0) return PTT_OK;
  PTT_CHECK_ARG(xyz != nullptr && idx != nullptr);
  cudaStream_t st = as_stream(stream);
  if (N <= 8192) {
    // smallest CTA whose registers hold the cloud with <= 16 points per thread; the step latency is
    // dominated by the cross-warp exchange, so fewer warps win until the per-thread scan grows.
    int threads, ppt;
    if (N <= 32) { threads = 32; ppt = 1; }
    else if (N <= 64) { threads = 32; ppt = 2; }
    else if (N <= 128) { threads = 32; ppt = 4; }
    else if (N <= 256) { threads = 64; ppt = 4; }
    else if (N <= 512) { threads = 128; ppt = 4; }
    else if (N <= 1024) { threads = 128; ppt = 8; }
    else if (N <= 2048) { threads = 256; ppt = 8; }
    else if (N <= 4096) { threads = 512; ppt = 8; }
    else { threads = 512; ppt = 16; }
    return ptt_fps_variant(xyz, B, N, npoint, idx, new_xyz, threads, ppt, stream);
  }
  const size_t need = ptt_furthest_point_sampling_workspace_bytes(B, N, npoint);
  if (workspace == nullptr || workspace_bytes < need) return PTT_ERR_WORKSPACE;
  fps_generic_kernel<<<B, 512, 0, st>>>(xyz, nullptr, N, npoint, (float*)workspace, idx, new_xyz); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" size_t ptt_furthest_point_sampling_with_dist_workspace_bytes(int B, int N, int npoint) {
  (void)npoint;
  return B > 0 ? (size_t)B * N * sizeof(float) : 0;
}

extern "C" int ptt_furthest_point_sampling_with_dist(const float* dist, int B, int N, int npoint, int* idx,
                                                     void* workspace, size_t workspace_bytes,
                                                     ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && N >= 1 && npoint >= 0);
  if (B == 0 || npoint == 0) return PTT_OK;
  PTT_CHECK_ARG(dist != nullptr && idx != nullptr);
  if (workspace == nullptr || workspace_bytes < (size_t)B * N * sizeof(float)) return PTT_ERR_WORKSPACE;
  fps_generic_kernel<<<B, 512, 0, as_stream(stream)>>>(nullptr, dist, N, npoint, (float*)workspace, idx, nullptr); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_gather_points(const float* points, const int* idx, int B, int C, int N, int M, float* out,
                                 ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && C >= 0 && N >= 1 && M >= 0);
  if (B == 0 || C == 0 || M == 0) return PTT_OK;
  PTT_CHECK_ARG(points && idx && out);
  dim3 grid(ceil_div(M, 256), ceil_div(C, 8), B);
  gather_points_kernel<<<grid, 256, 0, as_stream(stream)>>>(points, idx, C, N, M, out); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_gather_points_grad(const float* grad_out, const int* idx, int B, int C, int N, int M,
                                      float* grad_points, ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && C >= 0 && N >= 1 && M >= 0);
  if (B == 0 || C == 0) return PTT_OK;
  PTT_CHECK_ARG(grad_points != nullptr);
  cudaError_t e = cudaMemsetAsync(grad_points, 0, (size_t)B * C * N * sizeof(float), as_stream(stream));
  if (e != cudaSuccess) return (int)e;
  if (M == 0) return PTT_OK;
  PTT_CHECK_ARG(grad_out && idx);
  dim3 grid(ceil_div(M, 256), ceil_div(C, 8), B);
  gather_points_grad_kernel<<<grid, 256, 0, as_stream(stream)>>>(grad_out, idx, C, N, M, grad_points); PTT_LAUNCHED();
  return ptt_launch_status();
}
