// fp32 row-block contraction (CUDA-core FFMA path) + the nn.Linear entry points.
//
// This is the exact-fp32 path: 64x64 output tile per CTA, 16-deep K slabs staged in shared memory,
// 4x4 register tile per thread, fused scale/shift/ReLU/residual epilogue.  It serves (a) the
// ptt_linear_* C entry points, (b) every dense contraction of the SA-MLP and transformer paths until
// the tcgen05 kernels take them over, and (c) the on-device cross-check for those kernels.
#include "gemm.cuh"
#include "ptt_b200_tuning.h"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

template <bool VEC_A>
__global__ void __launch_bounds__(THREADS) gemm_ffma_kernel(PttGemmArgs p) {
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads; thread owns rows ty*4.., cols tx*4..
  const int row0 = blockIdx.x * BM, col0 = blockIdx.y * BN;

  // A loader: thread -> (row = tid / 4, k quad = tid % 4)
  const int ar = tid >> 2, ak = (tid & 3) * 4;
  const int grow = row0 + ar;
  const float* arow = nullptr;
  if (grow < p.R) {
    const long long src = p.a_rows ? (long long)p.a_rows[grow] : (long long)grow;
    arow = p.x + src * p.ldx;
  }
  // B loader: thread -> (k = tid / 16, col quad = tid % 16)
  const int bk = tid >> 4, bc = (tid & 15) * 4;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float4 ra, rb;
  auto fetch = [&](int k0) {
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    if (arow != nullptr) {
      const int k = k0 + ak;
      if (VEC_A && k + 3 < p.K) {
        ra = __ldg(reinterpret_cast<const float4*>(arow + k));
      } else {
        if (k < p.K) ra.x = __ldg(arow + k);
        if (k + 1 < p.K) ra.y = __ldg(arow + k + 1);
        if (k + 2 < p.K) ra.z = __ldg(arow + k + 2);
        if (k + 3 < p.K) ra.w = __ldg(arow + k + 3);
      }
    }
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    const int kb = k0 + bk, c = col0 + bc;
    if (kb < p.K && c < p.ldw) rb = __ldg(reinterpret_cast<const float4*>(p.wt + (size_t)kb * p.ldw + c));
  };
  auto stash = [&](int buf) {
    As[buf][ak + 0][ar] = ra.x;
    As[buf][ak + 1][ar] = ra.y;
    As[buf][ak + 2][ar] = ra.z;
    As[buf][ak + 3][ar] = ra.w;
    *reinterpret_cast<float4*>(&Bs[buf][bk][bc]) = rb;
  };

  const int nslab = (p.K + BK - 1) / BK;
  fetch(0);
  stash(0);
  __syncthreads();
  for (int s = 0; s < nslab; ++s) {
    const int buf = s & 1;
    if (s + 1 < nslab) fetch((s + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (s + 1 < nslab) {
      stash(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row0 + ty * 4 + i;
    if (r >= p.R) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + tx * 4 + j;
      if (c >= p.N) continue;
      float v = acc[i][j];
      if (p.scale) v *= __ldg(p.scale + c);
      if (p.shift) v += __ldg(p.shift + c);
      if (p.relu) v = fmaxf(v, 0.f);
      if (p.residual) v += __ldg(p.residual + (size_t)r * p.ldr + c);
      p.y[(size_t)r * p.ldy + c] = v;
    }
  }
}

// nn.Linear weight (Cout, K) [+ bias (Cout)] -> columns col0..col0+Cout-1 of the transposed image
// (K + 1 rows of ldw floats; row K is the bias).  The image is zero-filled by the caller first.
__global__ void linear_pack_kernel(const float* __restrict__ w, const float* __restrict__ bias, int K, int Cout,
                                   int ldw, int col0, float* __restrict__ params) {
  const int total = (K + 1) * Cout;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int k = e / Cout, c = e - k * Cout;
    params[(size_t)k * ldw + col0 + c] = k < K ? w[(size_t)c * K + k] : (bias ? bias[c] : 0.f);
  }
}

}  // namespace

static int g_force_ffma = 0;
// Test hook (declared in include/ptt_b200_tuning.h): 1 = run every contraction on the CUDA-core path.
extern "C" __attribute__((visibility("default"))) void ptt_debug_force_ffma(int on) { g_force_ffma = on; }

int ptt_gemm_launch(const PttGemmArgs& a, cudaStream_t st) {
  if (a.R <= 0 || a.N <= 0) return PTT_OK;
  if (a.batch > 1) {      // batched problems exist on the tensor-core path only
    const bool ok = a.wimg != nullptr && ptt_tc_gemm_supported(a) && a.x_bstride % 4 == 0;
    return ok ? ptt_tc_gemm_launch(a, a.wimg, st) : PTT_ERR_UNSUPPORTED;
  }
  if (a.wimg != nullptr && (!g_force_ffma || a.a_ka != nullptr) && ptt_tc_gemm_supported(a)) return ptt_tc_gemm_launch(a, a.wimg, st);
  if (a.a_ka != nullptr) return PTT_ERR_UNSUPPORTED;      // the operand transform exists on the tensor-core path only
  return ptt_gemm_launch_ffma(a, st);
}

int ptt_gemm_launch_ffma(const PttGemmArgs& a, cudaStream_t st) {
  if (a.R <= 0 || a.N <= 0) return PTT_OK;
  dim3 grid(ceil_div(a.R, BM), ceil_div(a.N, BN));
  const bool vec = (a.ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15u) == 0);
  if (vec) {
    gemm_ffma_kernel<true><<<grid, THREADS, 0, st>>>(a);
  } else {
    gemm_ffma_kernel<false><<<grid, THREADS, 0, st>>>(a);
  }
  PTT_LAUNCHED();
  return ptt_launch_status();
}

int ptt_linear_pack_cols(const float* weight, const float* bias, int K, int Cout, int ldw, int col0, float* params,
                         cudaStream_t st) {
  const int total = (K + 1) * Cout;
  linear_pack_kernel<<<min(ceil_div(total, 256), 1024), 256, 0, st>>>(weight, bias, K, Cout, ldw, col0, params); PTT_LAUNCHED();
  return ptt_launch_status();
}

int ptt_linear_pack_launch(const float* weight, const float* bias, int K, int Cout, float* params, cudaStream_t st) {
  return ptt_linear_pack_all(weight, K, 1, bias, K, Cout, params, st);
}

extern "C" size_t ptt_linear_params_floats(int K, int Cout) {
  if (K < 0 || Cout <= 0) return 0;
  return (size_t)(K + 1) * ptt_linear_ldw(Cout) + ptt_tc_weight_floats(K, Cout);
}

extern "C" int ptt_linear_pack(const float* weight, const float* bias, int K, int Cout, float* params,
                               ptt_stream_t stream) {
  PTT_CHECK_ARG(K >= 1 && Cout >= 1 && weight && params);
  return ptt_linear_pack_launch(weight, bias, K, Cout, params, as_stream(stream));
}

extern "C" int ptt_linear_pack_strided(const float* weight, long long ld_c, long long ld_k, const float* bias, int K, int Cout,
                                       float* params, ptt_stream_t stream) {
  PTT_CHECK_ARG(K >= 1 && Cout >= 1 && weight && params);
  return ptt_linear_pack_all(weight, ld_c, ld_k, bias, K, Cout, params, as_stream(stream));
}

extern "C" int ptt_linear_pack_batch(const PttPackDesc* descs_device, int count, ptt_stream_t stream) {
  PTT_CHECK_ARG(count >= 0 && count <= 65535 && (count == 0 || descs_device != nullptr));
  return ptt_linear_pack_batch_launch(descs_device, count, as_stream(stream));
}

extern "C" int ptt_linear_fwd(const float* x, int ldx, int R, int K, const float* params, int Cout, int relu,
                              const float* residual, int ldr, float* y, int ldy, ptt_stream_t stream) {
  PTT_CHECK_ARG(R >= 0 && K >= 1 && Cout >= 1 && ldx >= K && ldy >= Cout);
  if (R == 0) return PTT_OK;
  PTT_CHECK_ARG(x && params && y && (residual == nullptr || ldr >= Cout));
  PttGemmArgs a;
  a.x = x; a.ldx = ldx; a.R = R; a.K = K;
  a.wt = params; a.ldw = ptt_linear_ldw(Cout); a.N = Cout;
  a.shift = params + (size_t)K * a.ldw;
  a.wimg = params + (size_t)(K + 1) * a.ldw;
  a.relu = relu;
  a.residual = residual; a.ldr = ldr;
  a.y = y; a.ldy = ldy;
  return ptt_gemm_launch(a, as_stream(stream));
}
