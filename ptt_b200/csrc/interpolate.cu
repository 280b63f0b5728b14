// three_nn / three_interpolate (+grad) for sm_100a.  Replaces pointnet2_ops `_ext.three_nn`,
// `_ext.three_interpolate(_grad)` (pointnet2_utils.py:145,182,204).  Wrapped by the reference but
// never called by any PTT model; provided so the `_ext` surface is complete.
#include "common.cuh"

namespace {

constexpr int NN_TILE = 1024;

// one thread per unknown point; known points stream through shared memory in tiles
__global__ void three_nn_kernel(const float* __restrict__ unknown, const float* __restrict__ known, int n, int m,
                                float* __restrict__ dist2, int* __restrict__ idx) {
  __shared__ float s_k[NN_TILE * 3];
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (j < n) {
    const float* u = unknown + ((size_t)b * n + j) * 3;
    ux = u[0]; uy = u[1]; uz = u[2];
  }
  // upstream keeps the three best as doubles initialised to 1e40 and compares the float distance
  double best1 = 1e40, best2 = 1e40, best3 = 1e40;
  int besti1 = 0, besti2 = 0, besti3 = 0;
  for (int base = 0; base < m; base += NN_TILE) {
    const int cnt = min(NN_TILE, m - base);
    __syncthreads();
    for (int e = threadIdx.x; e < cnt * 3; e += blockDim.x) s_k[e] = known[((size_t)b * m + base) * 3 + e];
    __syncthreads();
    if (j < n) {
      for (int t = 0; t < cnt; ++t) {
        const float d = sq3(ux - s_k[3 * t], uy - s_k[3 * t + 1], uz - s_k[3 * t + 2]);
        const int k = base + t;
        if (d < best1) {
          best3 = best2; besti3 = besti2; best2 = best1; besti2 = besti1; best1 = d; besti1 = k;
        } else if (d < best2) {
          best3 = best2; besti3 = besti2; best2 = d; besti2 = k;
        } else if (d < best3) {
          best3 = d; besti3 = k;
        }
      }
    }
  }
  if (j < n) {
    float* dd = dist2 + ((size_t)b * n + j) * 3;
    int* ii = idx + ((size_t)b * n + j) * 3;
    dd[0] = (float)best1; dd[1] = (float)best2; dd[2] = (float)best3;
    ii[0] = besti1; ii[1] = besti2; ii[2] = besti3;
  }
}

// grid (ceil(n/256), c, B)
__global__ void three_interpolate_kernel(const float* __restrict__ points, const int* __restrict__ idx,
                                         const float* __restrict__ weight, int c, int m, int n,
                                         float* __restrict__ out) {
  const int b = blockIdx.z, l = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float* w = weight + ((size_t)b * n + j) * 3;
  const int* ii = idx + ((size_t)b * n + j) * 3;
  const float* src = points + ((size_t)b * c + l) * m;
  out[((size_t)b * c + l) * n + j] =
      __fmaf_rn(__ldg(src + ii[2]), w[2], __fmaf_rn(__ldg(src + ii[0]), w[0], __fmul_rn(__ldg(src + ii[1]), w[1])));
}

__global__ void three_interpolate_grad_kernel(const float* __restrict__ grad_out, const int* __restrict__ idx,
                                              const float* __restrict__ weight, int c, int n, int m,
                                              float* __restrict__ grad_points) {
  const int b = blockIdx.z, l = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const float g = grad_out[((size_t)b * c + l) * n + j];
  const float* w = weight + ((size_t)b * n + j) * 3;
  const int* ii = idx + ((size_t)b * n + j) * 3;
  float* dst = grad_points + ((size_t)b * c + l) * m;
  atomicAdd(dst + ii[0], g * w[0]);
  atomicAdd(dst + ii[1], g * w[1]);
  atomicAdd(dst + ii[2], g * w[2]);
}

}  // namespace

extern "C" int ptt_three_nn(const float* unknown, const float* known, int B, int n, int m, float* dist2, int* idx,
                            ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && n >= 0 && m >= 0);
  if (B == 0 || n == 0) return PTT_OK;
  PTT_CHECK_ARG(unknown && dist2 && idx && (known || m == 0));
  dim3 grid(ceil_div(n, 256), B);
  three_nn_kernel<<<grid, 256, 0, as_stream(stream)>>>(unknown, known, n, m, dist2, idx); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_three_interpolate(const float* points, const int* idx, const float* weight, int B, int c, int m,
                                     int n, float* out, ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && c >= 0 && m >= 1 && n >= 0);
  if (B == 0 || c == 0 || n == 0) return PTT_OK;
  PTT_CHECK_ARG(points && idx && weight && out);
  dim3 grid(ceil_div(n, 256), c, B);
  three_interpolate_kernel<<<grid, 256, 0, as_stream(stream)>>>(points, idx, weight, c, m, n, out); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_three_interpolate_grad(const float* grad_out, const int* idx, const float* weight, int B, int c,
                                          int n, int m, float* grad_points, ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && c >= 0 && m >= 1 && n >= 0);
  if (B == 0 || c == 0) return PTT_OK;
  PTT_CHECK_ARG(grad_points != nullptr);
  cudaError_t e = cudaMemsetAsync(grad_points, 0, (size_t)B * c * m * sizeof(float), as_stream(stream));
  if (e != cudaSuccess) return (int)e;
  if (n == 0) return PTT_OK;
  PTT_CHECK_ARG(grad_out && idx && weight);
  dim3 grid(ceil_div(n, 256), c, B);
  three_interpolate_grad_kernel<<<grid, 256, 0, as_stream(stream)>>>(grad_out, idx, weight, c, n, m, grad_points); PTT_LAUNCHED();
  return ptt_launch_status();
}
