// Ball query for sm_100a.  Replaces pointnet2_ops `_ext.ball_query` (pointnet2_utils.py:287).
//
// Upstream runs ONE THREAD per centre, serially scanning N points from global memory.  Here a WARP
// owns a centre: the cloud's xyz tile is staged once per CTA in shared memory (SoA, conflict-free),
// the 32 lanes test 32 consecutive points per step, and `ballot` + `popc` of the lower lanes gives
// every hit its rank in INDEX ORDER, so the result is exactly upstream's "first nsample hits in
// ascending k, tail padded with the first hit, zeros when empty".  The scan stops as soon as
// nsample hits exist.  Bytes: B*(12N + 12M + 4*M*nsample) algorithmic; xyz is re-staged by the
// ceil(M/64) CTAs of a cloud out of L2.
#include "common.cuh"

namespace {

constexpr int BQ_WARPS = 8;
constexpr int BQ_CENTRES_PER_CTA = 32;      // 4 centres per warp: enough CTAs to fill the machine at M = 64..512
constexpr int BQ_MAX_SMEM_POINTS = 16384;  // 192 KB of SoA floats

template <bool STAGED>
__global__ void __launch_bounds__(BQ_WARPS * 32) ball_query_kernel(const float* __restrict__ new_xyz,
                                                                    const float* __restrict__ xyz, int N, int M,
                                                                    float radius2, int ns,
                                                                    int* __restrict__ idx_out) {
  extern __shared__ float s_xyz[];  // [3][N] when STAGED
  const int b = blockIdx.y;
  const float* P = xyz + (size_t)b * N * 3;
  float* sx = s_xyz;
  float* sy = s_xyz + N;
  float* sz = s_xyz + 2 * N;
  if (STAGED) {
    for (int e = threadIdx.x; e < 3 * N; e += blockDim.x) {
      const float v = P[e];
      const int k = e / 3, c = e - 3 * k;
      (c == 0 ? sx : (c == 1 ? sy : sz))[k] = v;
    }
    __syncthreads();
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const int j_end = min(M, (int)(blockIdx.x + 1) * BQ_CENTRES_PER_CTA);
  for (int j = blockIdx.x * BQ_CENTRES_PER_CTA + warp; j < j_end; j += BQ_WARPS) {
    const float* q = new_xyz + ((size_t)b * M + j) * 3;
    const float nx = __ldg(q), ny = __ldg(q + 1), nz = __ldg(q + 2);
    int* row = idx_out + ((size_t)b * M + j) * ns;
    int cnt = 0, first = 0;
    // four 32-point chunks per step: their 12 shared-memory loads and 4 ballots are independent, so the scan is not a
    // chain of load -> test -> ballot latencies; hits are still committed chunk by chunk, in index order
    for (int base = 0; base < N && cnt < ns; base += 128) {
      unsigned ballots[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = base + u * 32 + lane;
        bool hit = false;
        if (k < N) {
          float x, y, z;
          if (STAGED) { x = sx[k]; y = sy[k]; z = sz[k]; }
          else { x = __ldg(P + 3 * k); y = __ldg(P + 3 * k + 1); z = __ldg(P + 3 * k + 2); }
          hit = sq3(nx - x, ny - y, nz - z) < radius2;
        }
        ballots[u] = __ballot_sync(0xffffffffu, hit);
      }
      if ((ballots[0] | ballots[1] | ballots[2] | ballots[3]) == 0u) continue;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned ballot = ballots[u];
        if (ballot == 0u || cnt >= ns) continue;
        if (cnt == 0) first = base + u * 32 + __ffs(ballot) - 1;
        const int pos = cnt + __popc(ballot & lt_mask);
        if (((ballot >> lane) & 1u) && pos < ns) row[pos] = base + u * 32 + lane;
        cnt += __popc(ballot);
      }
    }
    cnt = min(cnt, ns);
    for (int l = cnt + lane; l < ns; l += 32) row[l] = first;  // pad with the first hit; 0 if none
  }
}

}  // namespace

extern "C" int ptt_ball_query(const float* new_xyz, const float* xyz, int B, int N, int M, float radius,
                              int nsample, int* idx, ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && N >= 1 && M >= 0 && nsample >= 0);
  if (B == 0 || M == 0 || nsample == 0) return PTT_OK;
  PTT_CHECK_ARG(new_xyz && xyz && idx);
  const float radius2 = radius * radius;  // fp32 product, as upstream
  dim3 grid(ceil_div(M, BQ_CENTRES_PER_CTA), B);
  cudaStream_t st = as_stream(stream);
  if (N <= BQ_MAX_SMEM_POINTS) {
    const size_t smem = (size_t)3 * N * sizeof(float);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(ball_query_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    ball_query_kernel<true><<<grid, BQ_WARPS * 32, smem, st>>>(new_xyz, xyz, N, M, radius2, nsample, idx); PTT_LAUNCHED();
  } else {
    ball_query_kernel<false><<<grid, BQ_WARPS * 32, 0, st>>>(new_xyz, xyz, N, M, radius2, nsample, idx); PTT_LAUNCHED();
  }
  return ptt_launch_status();
}
