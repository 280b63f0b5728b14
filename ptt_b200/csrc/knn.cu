// k-nearest-neighbour selection for the transformer block (variants.py:150-151), sm_100a.
//
// The reference builds the full (B,n,n) distance matrix and ARGSORTS every row to keep 16 columns.
// Here a warp owns one query point: the n candidate distances sit in registers (n/32 per lane) and
// k rounds of a two-step `redux.sync` arg-min (value, then lowest index among equals) pick the
// neighbours in ascending (distance, index) order -- a stable sort's order; the reference's
// argsort is unstable, so its order among exact ties is unspecified.
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int KNN_WARPS = 4;

template <int NPL>  // candidates per lane: n <= 32 * NPL
__global__ void __launch_bounds__(KNN_WARPS * 32) knn_kernel(const float* __restrict__ xyz, int n, int k,
                                                              int* __restrict__ knn_idx) {
  extern __shared__ float s_xyz[];  // [3][n]
  const int b = blockIdx.y;
  const float* P = xyz + (size_t)b * n * 3;
  float* sx = s_xyz;
  float* sy = s_xyz + n;
  float* sz = s_xyz + 2 * n;
  for (int e = threadIdx.x; e < 3 * n; e += blockDim.x) {
    const int p = e / 3, c = e - 3 * p;
    (c == 0 ? sx : (c == 1 ? sy : sz))[p] = P[e];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = blockIdx.x * KNN_WARPS + warp; i < n; i += gridDim.x * KNN_WARPS) {
    const float qx = sx[i], qy = sy[i], qz = sz[i];
    float d[NPL];
#pragma unroll
    for (int t = 0; t < NPL; ++t) {
      const int j = lane + 32 * t;
      d[t] = j < n ? sq3_nofma(qx - sx[j], qy - sy[j], qz - sz[j]) : CUDART_INF_F;
      if (!(d[t] == d[t])) d[t] = CUDART_INF_F;  // NaN never selected before finite candidates
    }
    int* row = knn_idx + ((size_t)b * n + i) * k;
    unsigned used = 0u;  // bit t: candidate lane + 32*t already emitted
    for (int r = 0; r < k; ++r) {
      float best = CUDART_INF_F;
      int bt = NPL;  // NPL = "nothing left in this lane"
#pragma unroll
      for (int t = 0; t < NPL; ++t) {
        const bool open = !((used >> t) & 1u) && (lane + 32 * t < n);
        if (open && (bt == NPL || d[t] < best)) { best = d[t]; bt = t; }  // strict '<': lowest index wins ties
      }
      // non-negative floats (and +inf) order like their bit patterns
      const unsigned vb = bt < NPL ? __float_as_uint(best) : 0xffffffffu;
      const unsigned wv = __reduce_min_sync(0xffffffffu, vb);
      const unsigned cand = (vb == wv && bt < NPL) ? (unsigned)(lane + 32 * bt) : 0xffffffffu;
      const unsigned wj = __reduce_min_sync(0xffffffffu, cand);
      if (wj != 0xffffffffu && (int)(wj & 31u) == lane) used |= 1u << (wj >> 5);
      if (lane == 0) row[r] = wj == 0xffffffffu ? 0 : (int)wj;
    }
  }
}

template <int NPL>
int launch_knn(const float* xyz, int B, int n, int k, int* out, cudaStream_t st) {
  const size_t smem = (size_t)3 * n * sizeof(float);
  dim3 grid(min(ceil_div(n, KNN_WARPS), 64), B);
  knn_kernel<NPL><<<grid, KNN_WARPS * 32, smem, st>>>(xyz, n, k, out); PTT_LAUNCHED();
  return ptt_launch_status();
}

}  // namespace

extern "C" int ptt_knn(const float* xyz, int B, int n, int k, int* knn_idx, ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && n >= 0 && k >= 0);
  if (B == 0 || n == 0 || k == 0) return PTT_OK;
  PTT_CHECK_ARG(xyz && knn_idx && k <= n);
  cudaStream_t st = as_stream(stream);
  if (n <= 32) return launch_knn<1>(xyz, B, n, k, knn_idx, st);
  if (n <= 64) return launch_knn<2>(xyz, B, n, k, knn_idx, st);
  if (n <= 128) return launch_knn<4>(xyz, B, n, k, knn_idx, st);
  if (n <= 256) return launch_knn<8>(xyz, B, n, k, knn_idx, st);
  if (n <= 512) return launch_knn<16>(xyz, B, n, k, knn_idx, st);
  if (n <= 1024) return launch_knn<32>(xyz, B, n, k, knn_idx, st);
  return PTT_ERR_UNSUPPORTED;
}
