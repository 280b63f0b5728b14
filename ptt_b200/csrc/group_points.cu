// Neighbour grouping (+ gradient) and layout transposes for sm_100a.
// Replaces pointnet2_ops `_ext.group_points(_grad)` (pointnet2_utils.py:237,257).
//
// group_points is pure HBM traffic: it writes the nsample-times-duplicated tensor (B,C,M,K).  Each
// thread owns one (centre, sample) slot, reads its index once (coalesced) and streams 8 channels,
// so the big output is written fully coalesced and the gathers hit a 4 KB row that lives in L1/L2.
// The fused eval path (sa_mlp.cu) never materialises this tensor; this kernel serves the drop-in
// `_ext` API and the training path.
#include "common.cuh"

namespace {

// grid (ceil(M*K/256), ceil(C/8), B)
__global__ void group_points_kernel(const float* __restrict__ points, const int* __restrict__ idx, int C, int N,
                                    int MK, float* __restrict__ out) {
  const int b = blockIdx.z;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= MK) return;
  const int i = idx[(size_t)b * MK + l];
  const int c0 = blockIdx.y * 8;
  float v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) v[u] = (c0 + u < C) ? __ldg(points + ((size_t)b * C + c0 + u) * N + i) : 0.f;
#pragma unroll
  for (int u = 0; u < 8; ++u)
    if (c0 + u < C) __stcs(out + ((size_t)b * C + c0 + u) * MK + l, v[u]);  // streaming store: written once
}

__global__ void group_points_grad_kernel(const float* __restrict__ grad_out, const int* __restrict__ idx, int C,
                                         int N, int MK, float* __restrict__ grad_points) {
  const int b = blockIdx.z;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= MK) return;
  const int i = idx[(size_t)b * MK + l];
  const int c0 = blockIdx.y * 8;
#pragma unroll
  for (int u = 0; u < 8; ++u)
    if (c0 + u < C) atomicAdd(grad_points + ((size_t)b * C + c0 + u) * N + i, __ldcs(grad_out + ((size_t)b * C + c0 + u) * MK + l));
}

// (B,C,N) -> (B,N,ld), 32x32 tiles through shared memory; columns C..ld-1 are zero-filled.
__global__ void cm_to_pm_kernel(const float* __restrict__ src, int C, int N, float* __restrict__ dst, int ld) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, n = n0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && n < N) ? src[((size_t)b * C + c) * N + n] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int n = n0 + r, c = c0 + threadIdx.x;
    if (n < N && c < ld) dst[((size_t)b * N + n) * ld + c] = tile[threadIdx.x][r];
  }
}

// (B,N,ld)[:, :, 0:C] -> (B,C,N)
__global__ void pm_to_cm_kernel(const float* __restrict__ src, int ld, int C, int N, float* __restrict__ dst) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int n = n0 + r, c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (c < C && n < N) ? src[((size_t)b * N + n) * ld + c] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r, n = n0 + threadIdx.x;
    if (c < C && n < N) dst[((size_t)b * C + c) * N + n] = tile[threadIdx.x][r];
  }
}

}  // namespace

extern "C" int ptt_group_points(const float* points, const int* idx, int B, int C, int N, int M, int K,
                                float* out, ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && C >= 0 && N >= 1 && M >= 0 && K >= 0);
  if (B == 0 || C == 0 || M == 0 || K == 0) return PTT_OK;
  PTT_CHECK_ARG(points && idx && out);
  dim3 grid(ceil_div(M * K, 256), ceil_div(C, 8), B);
  group_points_kernel<<<grid, 256, 0, as_stream(stream)>>>(points, idx, C, N, M * K, out); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_group_points_grad(const float* grad_out, const int* idx, int B, int C, int N, int M, int K,
                                     float* grad_points, ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && C >= 0 && N >= 1 && M >= 0 && K >= 0);
  if (B == 0 || C == 0) return PTT_OK;
  PTT_CHECK_ARG(grad_points != nullptr);
  cudaError_t e = cudaMemsetAsync(grad_points, 0, (size_t)B * C * N * sizeof(float), as_stream(stream));
  if (e != cudaSuccess) return (int)e;
  if (M == 0 || K == 0) return PTT_OK;
  PTT_CHECK_ARG(grad_out && idx);
  dim3 grid(ceil_div(M * K, 256), ceil_div(C, 8), B);
  group_points_grad_kernel<<<grid, 256, 0, as_stream(stream)>>>(grad_out, idx, C, N, M * K, grad_points); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_cm_to_pm(const float* src_cm, int B, int C, int N, float* dst_pm, int ld, ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && C >= 0 && N >= 0 && ld >= C);
  if (B == 0 || N == 0 || ld == 0) return PTT_OK;
  PTT_CHECK_ARG(dst_pm && (src_cm || C == 0));
  dim3 grid(ceil_div(N, 32), ceil_div(ld, 32), B);
  cm_to_pm_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(src_cm, C, N, dst_pm, ld); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_pm_to_cm(const float* src_pm, int ld, int B, int C, int N, float* dst_cm, ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && C >= 0 && N >= 0 && ld >= C);
  if (B == 0 || N == 0 || C == 0) return PTT_OK;
  PTT_CHECK_ARG(src_pm && dst_cm);
  dim3 grid(ceil_div(N, 32), ceil_div(C, 32), B);
  pm_to_cm_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(src_pm, ld, C, N, dst_cm); PTT_LAUNCHED();
  return ptt_launch_status();
}
