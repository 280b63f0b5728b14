// Training-mode kernels of the hot path (SURVEY.md 8(e) / BASELINE configs[3]): what surrounds the tensor-core
// contractions (tc_gemm.cu forward / input gradient, tc_wgrad.cu weight gradient) when BatchNorm runs on batch
// statistics and gradients flow.  Activations are PAIR-ROW matrices (rows = (centre, sample) pairs, channels contiguous),
// and a layer's normalised output is never stored: its consumer applies relu(ka * y + kb) per channel while it loads the
// PRE-BatchNorm output y (ka = gamma * rstd, kb = beta - mean * ka).
//
//   reference                                                       here
//   QueryAndGroup + cat (pointnet2_utils.py:320-380)                sa_group_rows(_grad)
//   Conv2d 1x1 -> BatchNorm2d(train) -> ReLU (pytorch_utils.py)     ws_gemm (statistics in its epilogue; else tc_gemm +
//                                                                   col_stats) + bn_train_finalize (two-phase)
//   F.max_pool2d over nsample (pointnet2_modules.py:85)             bn_relu_maxpool (+ the first-max index for backward)
//   autograd of BatchNorm / ReLU / max_pool                         bn_relu_bwd_reduce + bn_relu_bwd_apply (dense),
//                                                                   bn_relu_bwd_pooled_* (below the max: arg-max rows only)
#include "gemm.cuh"

namespace {

constexpr int TO_THREADS = 256;

// ---------------------------------------------------------------------------------------------- column reductions
// Rows [r0, r1) of this CTA, 4 consecutive columns per thread, TO_THREADS / (C / 4) rows per pass; register partial sums,
// shared-memory combine over the row slots, one double atomicAdd per column and CTA.
template <int NV, class F>
__device__ __forceinline__ void column_reduce(long long R, int C, double* out /* [NV][C] */, F&& row_values) {
  __shared__ float s_part[TO_THREADS * 4 * 2];
  const int tpr = C / 4;                                   // threads per row
  const int slots = TO_THREADS / tpr;                      // rows per pass
  const int slot = threadIdx.x / tpr, c4 = threadIdx.x - slot * tpr;
  const long long chunk = (R + gridDim.x - 1) / gridDim.x;
  const long long r0 = (long long)blockIdx.x * chunk, r1 = r0 + chunk < R ? r0 + chunk : R;
  float acc[NV][4];
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int u = 0; u < 4; ++u) acc[v][u] = 0.f;
  if (slot < slots) {
#pragma unroll 4
    for (long long r = r0 + slot; r < r1; r += slots) row_values(r, c4 * 4, acc);
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u) s_part[threadIdx.x * 4 + u] = acc[v][u];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += TO_THREADS) {
      float s = 0.f;
      for (int sl = 0; sl < slots; ++sl) s += s_part[(sl * tpr + c / 4) * 4 + (c & 3)];
      atomicAdd(out + (size_t)v * C + c, (double)s);
    }
  }
}

__global__ void __launch_bounds__(TO_THREADS) col_stats_kernel(const float* __restrict__ y, int ldy, long long R, int C,
                                                                double* __restrict__ sums) {
  column_reduce<2>(R, C, sums, [&](long long r, int c, float (&acc)[2][4]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(y + r * ldy + c));
    acc[0][0] += v.x; acc[0][1] += v.y; acc[0][2] += v.z; acc[0][3] += v.w;
    acc[1][0] = fmaf(v.x, v.x, acc[1][0]); acc[1][1] = fmaf(v.y, v.y, acc[1][1]);
    acc[1][2] = fmaf(v.z, v.z, acc[1][2]); acc[1][3] = fmaf(v.w, v.w, acc[1][3]);
  });
}

// BatchNorm2d.forward(training) bookkeeping: batch mean / biased variance -> (ka, kb, mean, rstd); running statistics
// updated with the unbiased variance and `momentum` (torch semantics)
__global__ void bn_train_finalize_kernel(const double* __restrict__ sums, long long R, int C, const float* __restrict__ gamma,
                                         const float* __restrict__ beta, float eps, float momentum,
                                         float* __restrict__ running_mean, float* __restrict__ running_var,
                                         float* __restrict__ ka, float* __restrict__ kb, float* __restrict__ mean_out,
                                         float* __restrict__ rstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sums[c] / (double)R;
  double var = sums[C + c] / (double)R - mean * mean;
  var = var > 0.0 ? var : 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  const float a = g * rstd;
  ka[c] = a;
  kb[c] = b - (float)mean * a;
  mean_out[c] = (float)mean;
  rstd_out[c] = rstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = R > 1 ? var * (double)R / (double)(R - 1) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// ---------------------------------------------------------------------------------------------- grouping
// X0[(b, j, s), :] = [ (xyz[idx] - new_xyz[j]) (/ radius) | feats[idx, 0:C] | 0 pad ]   -- the reference's channel order
template <typename I>   // I = unsigned when rows * ld / 4 < 2^32 (the 64-bit divisions are emulated: ~3x the instructions)
__global__ void __launch_bounds__(TO_THREADS) sa_group_rows_kernel2(const float* __restrict__ xyz, const float* __restrict__ feats,
                                                                     int ldf, const float* __restrict__ new_xyz,
                                                                     const int* __restrict__ idx, int N, int M, int ns, int C,
                                                                     float radius, int normalize, long long rows,
                                                                     float* __restrict__ out, int ld) {
  // one thread per float4 of an output row (ld % 4 == 0): consecutive threads write consecutive 16-byte pieces
  const I q4 = (I)(ld / 4);
  const I total = (I)rows * q4;
  for (I e = (I)blockIdx.x * TO_THREADS + threadIdx.x; e < total; e += (I)gridDim.x * TO_THREADS) {
    const I r = e / q4;
    const int c0 = (int)(e - r * q4) * 4;
    const I cj = r / (I)ns;                               // (b, j)
    const I b = cj / (I)M;
    const size_t src = (size_t)b * N + __ldg(idx + r);
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int c = c0 + u;
      if (c < 3) {
        float v = __fsub_rn(__ldg(xyz + src * 3 + c), __ldg(new_xyz + (size_t)cj * 3 + c));
        if (normalize) v = __fdiv_rn(v, radius);
        o[u] = v;
      } else {
        o[u] = c < 3 + C ? __ldg(feats + src * ldf + (c - 3)) : 0.f;
      }
    }
    *reinterpret_cast<float4*>(out + (size_t)r * ld + c0) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// scatter of dX0: d_feats[b, idx, c] += dX0[r, 3 + c]; d_xyz[b, idx] += dX0[r, 0:3] (/ radius); d_new_xyz[b, j] -= the same
__global__ void __launch_bounds__(TO_THREADS) sa_group_rows_grad_kernel(const float* __restrict__ dx0, int ld,
                                                                         const int* __restrict__ idx, int N, int M, int ns,
                                                                         int C, float radius, int normalize, long long rows,
                                                                         float* __restrict__ d_feats, int ldf,
                                                                         float* __restrict__ d_xyz, float* __restrict__ d_new_xyz) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (TO_THREADS / 32) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (TO_THREADS / 32);
  for (long long r = warp; r < rows; r += nwarps) {
    const long long cj = r / ns;
    const long long b = cj / M;
    const int i = __ldg(idx + r);
    const float* g = dx0 + r * ld;
    if (d_feats != nullptr) {
      float* f = d_feats + (b * N + i) * (long long)ldf;
      for (int c = lane; c < C; c += 32) atomicAdd(f + c, __ldg(g + 3 + c));
    }
    if (d_xyz != nullptr && lane < 3) {
      float v = __ldg(g + lane);
      if (normalize) v = __fdiv_rn(v, radius);
      atomicAdd(d_xyz + (b * N + i) * 3 + lane, v);
      if (d_new_xyz != nullptr) atomicAdd(d_new_xyz + cj * 3 + lane, -v);
    }
  }
}

// ---------------------------------------------------------------------------------------------- pooling
// out[g, c] = max_s relu(ka[c] * y[g*ns + s, c] + kb[c]); arg[g, c] = the FIRST s attaining it (max_pool2d's choice)
__global__ void __launch_bounds__(TO_THREADS) bn_relu_maxpool_kernel(const float* __restrict__ y, int ldy, long long groups, int ns,
                                                                      int C, const float* __restrict__ ka, const float* __restrict__ kb,
                                                                      float* __restrict__ out, int ldo, int* __restrict__ arg) {
  const long long total = groups * (C / 4);
  for (long long e = (long long)blockIdx.x * TO_THREADS + threadIdx.x; e < total; e += (long long)gridDim.x * TO_THREADS) {
    const long long g = e / (C / 4);
    const int c = (int)(e - g * (C / 4)) * 4;
    const float4 a = __ldg(reinterpret_cast<const float4*>(ka + c)), b = __ldg(reinterpret_cast<const float4*>(kb + c));
    float best[4] = {-1.f, -1.f, -1.f, -1.f};
    int bi[4] = {0, 0, 0, 0};
    for (int s = 0; s < ns; ++s) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(y + (g * ns + s) * ldy + c));
      const float z[4] = {fmaxf(fmaf(v.x, a.x, b.x), 0.f), fmaxf(fmaf(v.y, a.y, b.y), 0.f), fmaxf(fmaf(v.z, a.z, b.z), 0.f),
                          fmaxf(fmaf(v.w, a.w, b.w), 0.f)};
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (z[u] > best[u]) { best[u] = z[u]; bi[u] = s; }
    }
    *reinterpret_cast<float4*>(out + g * ldo + c) = make_float4(best[0], best[1], best[2], best[3]);
    *reinterpret_cast<int4*>(arg + g * C + c) = make_int4(bi[0], bi[1], bi[2], bi[3]);
  }
}

// ---------------------------------------------------------------------------------------------- BatchNorm + ReLU backward
// m = dz * [ka * y + kb > 0];   s1 = sum_r m (= d beta),  s2 = sum_r m * yhat (= d gamma),  yhat = (y - mean) * rstd.
// dz is dense (R, ldz), or POOLED: dz (groups, ldz) reaches only row arg[g, c] of its group (max_pool2d backward).
struct BnBwd {
  const float* dz; int ldz;
  const int* arg; int ns;                 // arg != nullptr: pooled
  const float* y; int ldy;
  const float *ka, *kb, *mean, *rstd;
  long long R; int C;
};

struct BnChan {                           // the per-channel terms of 4 consecutive channels
  float a[4], b[4], mu[4], rs[4];
};

__device__ __forceinline__ BnChan bn_chan(const BnBwd& p, int c) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p.ka + c)), b = __ldg(reinterpret_cast<const float4*>(p.kb + c));
  const float4 mu = __ldg(reinterpret_cast<const float4*>(p.mean + c)), rs = __ldg(reinterpret_cast<const float4*>(p.rstd + c));
  BnChan ch = {{a.x, a.y, a.z, a.w}, {b.x, b.y, b.z, b.w}, {mu.x, mu.y, mu.z, mu.w}, {rs.x, rs.y, rs.z, rs.w}};
  return ch;
}

__device__ __forceinline__ void bn_bwd_m(const BnBwd& p, const BnChan& ch, long long r, int c, float (&m)[4], float (&yh)[4]) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p.y + r * p.ldy + c));
  const float yv[4] = {v.x, v.y, v.z, v.w};
  float d[4];
  if (p.arg != nullptr) {
    long long g;
    int s;
    if (p.R <= 0x7fffffffLL) {              // 32-bit division (the 64-bit one is emulated)
      const unsigned gu = (unsigned)r / (unsigned)p.ns;
      g = gu;
      s = (int)((unsigned)r - gu * (unsigned)p.ns);
    } else {
      g = r / p.ns;
      s = (int)(r - g * p.ns);
    }
    const int4 ai = __ldg(reinterpret_cast<const int4*>(p.arg + g * p.C + c));
    const float4 dv = __ldg(reinterpret_cast<const float4*>(p.dz + g * p.ldz + c));
    d[0] = ai.x == s ? dv.x : 0.f; d[1] = ai.y == s ? dv.y : 0.f; d[2] = ai.z == s ? dv.z : 0.f; d[3] = ai.w == s ? dv.w : 0.f;
  } else {
    const float4 dv = __ldg(reinterpret_cast<const float4*>(p.dz + r * p.ldz + c));
    d[0] = dv.x; d[1] = dv.y; d[2] = dv.z; d[3] = dv.w;
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    m[u] = fmaf(yv[u], ch.a[u], ch.b[u]) > 0.f ? d[u] : 0.f;
    yh[u] = (yv[u] - ch.mu[u]) * ch.rs[u];
  }
}

__global__ void __launch_bounds__(TO_THREADS) bn_relu_bwd_reduce_kernel(const BnBwd p, double* __restrict__ sums) {
  const BnChan ch = bn_chan(p, (threadIdx.x % (p.C / 4)) * 4);   // column_reduce gives a thread the same 4 channels every row
  column_reduce<2>(p.R, p.C, sums, [&](long long r, int c, float (&acc)[2][4]) {
    float m[4], yh[4];
    bn_bwd_m(p, ch, r, c, m, yh);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc[0][u] += m[u];
      acc[1][u] = fmaf(m[u], yh[u], acc[1][u]);
    }
  });
}

// dy = gamma * rstd * (m - s1 / R - yhat * s2 / R)
__global__ void __launch_bounds__(TO_THREADS) bn_relu_bwd_apply_kernel(const BnBwd p, const double* __restrict__ sums,
                                                                        const float* __restrict__ gamma, float* __restrict__ dy,
                                                                        int ld_dy, float* __restrict__ dparam) {
  if (dparam != nullptr && blockIdx.x == 0)                 // (2, C) fp32 copy of the sums: d beta, d gamma as the optimiser wants them
    for (int c = threadIdx.x; c < 2 * p.C; c += TO_THREADS) dparam[c] = (float)sums[c];
  const int tpr = p.C / 4;                                  // threads per row
  const double inv_r = 1.0 / (double)p.R;
  auto coefficients = [&](int c, float (&gs)[4], float (&s1)[4], float (&s2)[4]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      gs[u] = (gamma ? __ldg(gamma + c + u) : 1.f) * __ldg(p.rstd + c + u);
      s1[u] = (float)(sums[c + u] * inv_r);
      s2[u] = (float)(sums[p.C + c + u] * inv_r);
    }
  };
  if (TO_THREADS % tpr == 0) {
    // a thread keeps its 4 channels over all its rows: every per-channel term is loaded once
    const int c = (threadIdx.x % tpr) * 4;
    const int rpp = TO_THREADS / tpr;                       // rows per CTA and pass
    const BnChan ch = bn_chan(p, c);
    float gs[4], s1[4], s2[4];
    coefficients(c, gs, s1, s2);
#pragma unroll 4
    for (long long r = (long long)blockIdx.x * rpp + threadIdx.x / tpr; r < p.R; r += (long long)gridDim.x * rpp) {
      float m[4], yh[4];
      bn_bwd_m(p, ch, r, c, m, yh);
      *reinterpret_cast<float4*>(dy + r * ld_dy + c) =
          make_float4(gs[0] * (m[0] - s1[0] - yh[0] * s2[0]), gs[1] * (m[1] - s1[1] - yh[1] * s2[1]),
                      gs[2] * (m[2] - s1[2] - yh[2] * s2[2]), gs[3] * (m[3] - s1[3] - yh[3] * s2[3]));
    }
    return;
  }
  const long long total = p.R * tpr;
  for (long long e = (long long)blockIdx.x * TO_THREADS + threadIdx.x; e < total; e += (long long)gridDim.x * TO_THREADS) {
    const long long r = e / tpr;
    const int c = (int)(e - r * tpr) * 4;
    const BnChan ch = bn_chan(p, c);
    float m[4], yh[4], gs[4], s1[4], s2[4];
    coefficients(c, gs, s1, s2);
    bn_bwd_m(p, ch, r, c, m, yh);
    *reinterpret_cast<float4*>(dy + r * ld_dy + c) =
        make_float4(gs[0] * (m[0] - s1[0] - yh[0] * s2[0]), gs[1] * (m[1] - s1[1] - yh[1] * s2[1]),
                    gs[2] * (m[2] - s1[2] - yh[2] * s2[2]), gs[3] * (m[3] - s1[3] - yh[3] * s2[3]));
  }
}

// POOLED form (the layer below the max over nsample): only the arg-max row of a (group, channel) carries gradient, so the
// two sums need one y value per (group, channel) -- a gather of groups x C values instead of a pass over R x C --
// and the apply pass walks a group's ns rows with its arg / dz loaded once (no per-row division, no per-row compare loads).
__global__ void __launch_bounds__(TO_THREADS) bn_relu_bwd_pooled_reduce_kernel(const BnBwd p, double* __restrict__ sums) {
  const BnChan ch = bn_chan(p, (threadIdx.x % (p.C / 4)) * 4);
  column_reduce<2>(p.R / p.ns, p.C, sums, [&](long long g, int c, float (&acc)[2][4]) {
    const int4 ai = __ldg(reinterpret_cast<const int4*>(p.arg + g * p.C + c));
    const float4 dv = __ldg(reinterpret_cast<const float4*>(p.dz + g * p.ldz + c));
    const int a4[4] = {ai.x, ai.y, ai.z, ai.w};
    const float d4[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float yv = __ldg(p.y + (g * p.ns + a4[u]) * p.ldy + c + u);
      const float m = fmaf(yv, ch.a[u], ch.b[u]) > 0.f ? d4[u] : 0.f;
      acc[0][u] += m;
      acc[1][u] = fmaf(m, (yv - ch.mu[u]) * ch.rs[u], acc[1][u]);
    }
  });
}

// requires TO_THREADS % (C / 4) == 0: a thread keeps its 4 channels, TO_THREADS / (C / 4) groups per CTA and pass
__global__ void __launch_bounds__(TO_THREADS) bn_relu_bwd_pooled_apply_kernel(const BnBwd p, const double* __restrict__ sums,
                                                                               const float* __restrict__ gamma,
                                                                               float* __restrict__ dy, int ld_dy,
                                                                               float* __restrict__ dparam) {
  if (dparam != nullptr && blockIdx.x == 0)
    for (int c = threadIdx.x; c < 2 * p.C; c += TO_THREADS) dparam[c] = (float)sums[c];
  const int tpr = p.C / 4;
  const int c = (threadIdx.x % tpr) * 4;
  const int gpp = TO_THREADS / tpr;
  const double inv_r = 1.0 / (double)p.R;
  const BnChan ch = bn_chan(p, c);
  float gs[4], s1[4], s2[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    gs[u] = (gamma ? __ldg(gamma + c + u) : 1.f) * ch.rs[u];
    s1[u] = (float)(sums[c + u] * inv_r);
    s2[u] = (float)(sums[p.C + c + u] * inv_r);
  }
  const long long groups = p.R / p.ns;
  for (long long g = (long long)blockIdx.x * gpp + threadIdx.x / tpr; g < groups; g += (long long)gridDim.x * gpp) {
    const int4 ai = __ldg(reinterpret_cast<const int4*>(p.arg + g * p.C + c));
    const float4 dv = __ldg(reinterpret_cast<const float4*>(p.dz + g * p.ldz + c));
    const float* yp = p.y + g * p.ns * p.ldy + c;
    float* op = dy + g * p.ns * ld_dy + c;
#pragma unroll 4
    for (int s = 0; s < p.ns; ++s) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(yp + (long long)s * p.ldy));
      const float yv[4] = {v.x, v.y, v.z, v.w};
      const float d4[4] = {ai.x == s ? dv.x : 0.f, ai.y == s ? dv.y : 0.f, ai.z == s ? dv.z : 0.f, ai.w == s ? dv.w : 0.f};
      float o[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float m = fmaf(yv[u], ch.a[u], ch.b[u]) > 0.f ? d4[u] : 0.f;
        o[u] = gs[u] * (m - s1[u] - (yv[u] - ch.mu[u]) * ch.rs[u] * s2[u]);
      }
      *reinterpret_cast<float4*>(op + (long long)s * ld_dy) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

int grid_rows(long long work, int per_block) {
  long long b = (work + per_block - 1) / per_block;
  return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

}  // namespace

extern "C" int ptt_col_stats(const float* y, int ldy, long long R, int C, double* sums, ptt_stream_t stream);

extern "C" int ptt_linear_fwd_ex(const float* x, int ldx, int R, int K, const float* a_ka, const float* a_kb,
                                 const float* params, int Cout, int relu, const float* residual, int ldr, float* y, int ldy,
                                 ptt_stream_t stream) {
  PTT_CHECK_ARG(R >= 0 && K >= 1 && Cout >= 1 && ldx >= K && ldy >= Cout && ((a_ka == nullptr) == (a_kb == nullptr)));
  if (R == 0) return PTT_OK;
  PTT_CHECK_ARG(x && params && y && (residual == nullptr || ldr >= Cout));
  PttGemmArgs a;
  a.x = x; a.ldx = ldx; a.R = R; a.K = K; a.a_ka = a_ka; a.a_kb = a_kb;
  a.wt = params; a.ldw = ptt_linear_ldw(Cout); a.N = Cout;
  a.shift = params + (size_t)K * a.ldw;
  a.wimg = params + (size_t)(K + 1) * a.ldw;
  a.relu = relu;
  a.residual = residual; a.ldr = ldr;
  a.y = y; a.ldy = ldy;
  return ptt_gemm_launch(a, as_stream(stream));
}

// y = f(x) . W^T [+ bias]; sums_or_null (2, Cout) double <- column sums of y and y^2 (phase 1 of the two-phase BatchNorm).
// Bias-free layers over many rows run on the weight-stationary persistent kernel, whose transposed epilogue produces the
// statistics for free; everything else is ptt_linear_fwd_ex followed by the column reduction.
extern "C" int ptt_linear_fwd_stats(const float* x, int ldx, long long R, int K, const float* a_ka, const float* a_kb,
                                    const float* params, int Cout, int has_bias, float* y, int ldy, double* sums_or_null,
                                    ptt_stream_t stream) {
  PTT_CHECK_ARG(R >= 0 && R <= 0x7fffffffLL && K >= 1 && Cout >= 1 && ldx >= K && ldy >= Cout &&
                ((a_ka == nullptr) == (a_kb == nullptr)));
  cudaStream_t st = as_stream(stream);
  if (R == 0) {
    if (sums_or_null != nullptr) {
      cudaError_t e = cudaMemsetAsync(sums_or_null, 0, (size_t)2 * Cout * sizeof(double), st);
      if (e != cudaSuccess) return (int)e;
    }
    return PTT_OK;
  }
  PTT_CHECK_ARG(x && params && y);
  const int ldw = ptt_linear_ldw(Cout);
  if (!has_bias && ptt_ws_gemm_supported(x, ldx, R, K, Cout))
    return ptt_ws_gemm_launch(x, ldx, R, K, a_ka, a_kb, params + (size_t)(K + 1) * ldw, Cout, y, ldy, sums_or_null, st);
  int rc = ptt_linear_fwd_ex(x, ldx, (int)R, K, a_ka, a_kb, params, Cout, 0, nullptr, 0, y, ldy, stream);
  if (rc != PTT_OK || sums_or_null == nullptr) return rc;
  return ptt_col_stats(y, ldy, R, Cout, sums_or_null, stream);
}

extern "C" int ptt_linear_wgrad(const float* dy, int ldy, const float* x, int ldx, const float* x_ka, const float* x_kb,
                                long long R, int M, int N, float* dw, int ldw, float* dbias_or_null, ptt_stream_t stream) {
  PTT_CHECK_ARG(R >= 0 && M >= 1 && N >= 1 && ldy >= M && ldx >= N && ldw >= N && ((x_ka == nullptr) == (x_kb == nullptr)));
  if (R == 0) return PTT_OK;
  PTT_CHECK_ARG(dy && x && dw);
  return ptt_tc_wgrad_launch(dy, ldy, x, ldx, x_ka, x_kb, R, M, N, dw, ldw, dbias_or_null, as_stream(stream));
}

extern "C" int ptt_col_stats(const float* y, int ldy, long long R, int C, double* sums, ptt_stream_t stream) {
  PTT_CHECK_ARG(R >= 0 && C >= 4 && C % 4 == 0 && C <= 4 * TO_THREADS && ldy >= C && ldy % 4 == 0 && sums);
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)2 * C * sizeof(double), st);
  if (e != cudaSuccess) return (int)e;
  if (R == 0) return PTT_OK;
  PTT_CHECK_ARG(y != nullptr);
  col_stats_kernel<<<grid_rows(R, 64), TO_THREADS, 0, st>>>(y, ldy, R, C, sums); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_bn_train_finalize(const double* sums, long long R, int C, const float* gamma, const float* beta, float eps,
                                     float momentum, float* running_mean, float* running_var, float* ka, float* kb,
                                     float* mean, float* rstd, ptt_stream_t stream) {
  PTT_CHECK_ARG(R >= 1 && C >= 1 && sums && ka && kb && mean && rstd);
  bn_train_finalize_kernel<<<ceil_div(C, 128), 128, 0, as_stream(stream)>>>(sums, R, C, gamma, beta, eps, momentum, running_mean,
                                                                            running_var, ka, kb, mean, rstd); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_sa_group_rows(const float* xyz, const float* feats, int ldf, const float* new_xyz, const int* idx, int B,
                                 int N, int M, int ns, int C, float radius, int normalize_xyz, float* rows_out, int ld,
                                 ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && N >= 1 && M >= 0 && ns >= 1 && C >= 0 && ld >= C + 3 && ld % 4 == 0 && (C == 0 || ldf >= C));
  const long long rows = (long long)B * M * ns;
  if (rows == 0) return PTT_OK;
  PTT_CHECK_ARG(xyz && new_xyz && idx && rows_out && (C == 0 || feats));
  const int grid = grid_rows(rows * (ld / 4), TO_THREADS);
  if (rows * (ld / 4) + (long long)grid * TO_THREADS < 0xffffffffLL)
    sa_group_rows_kernel2<unsigned><<<grid, TO_THREADS, 0, as_stream(stream)>>>(xyz, feats, ldf, new_xyz, idx, N, M, ns, C, radius,
                                                                              normalize_xyz, rows, rows_out, ld);
  else
    sa_group_rows_kernel2<unsigned long long><<<grid, TO_THREADS, 0, as_stream(stream)>>>(xyz, feats, ldf, new_xyz, idx, N, M, ns, C,
                                                                                        radius, normalize_xyz, rows, rows_out, ld);
  PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_sa_group_rows_grad(const float* d_rows, int ld, const int* idx, int B, int N, int M, int ns, int C,
                                      float radius, int normalize_xyz, float* d_feats, int ldf, float* d_xyz, float* d_new_xyz,
                                      ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && N >= 1 && M >= 0 && ns >= 1 && C >= 0 && ld >= C + 3 && (d_feats == nullptr || ldf >= C));
  const long long rows = (long long)B * M * ns;
  if (rows == 0 || (d_feats == nullptr && d_xyz == nullptr)) return PTT_OK;
  PTT_CHECK_ARG(d_rows && idx);
  sa_group_rows_grad_kernel<<<grid_rows(rows, 8), TO_THREADS, 0, as_stream(stream)>>>(d_rows, ld, idx, N, M, ns, C, radius,
                                                                                     normalize_xyz, rows, d_feats, ldf, d_xyz,
                                                                                     d_new_xyz); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_bn_relu_maxpool(const float* y, int ldy, long long groups, int ns, int C, const float* ka, const float* kb,
                                   float* out, int ldo, int* argmax, ptt_stream_t stream) {
  PTT_CHECK_ARG(groups >= 0 && ns >= 1 && C >= 4 && C % 4 == 0 && ldy >= C && ldy % 4 == 0 && ldo >= C && ldo % 4 == 0);
  if (groups == 0) return PTT_OK;
  PTT_CHECK_ARG(y && ka && kb && out && argmax);
  bn_relu_maxpool_kernel<<<grid_rows(groups * (C / 4), TO_THREADS), TO_THREADS, 0, as_stream(stream)>>>(y, ldy, groups, ns, C, ka, kb,
                                                                                                       out, ldo, argmax);
  PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_bn_relu_bwd(const float* dz, int ldz, const int* argmax_or_null, int ns, const float* y, int ldy, long long R,
                               int C, const float* ka, const float* kb, const float* mean, const float* rstd,
                               const float* gamma, double* sums, float* dy, int ld_dy, float* dparam_or_null,
                               ptt_stream_t stream) {
  PTT_CHECK_ARG(R >= 1 && C >= 4 && C % 4 == 0 && C <= 4 * TO_THREADS && ldy >= C && ldy % 4 == 0 && ldz >= C && ldz % 4 == 0 &&
                ld_dy >= C && ld_dy % 4 == 0 && ns >= 1);
  PTT_CHECK_ARG(dz && y && ka && kb && mean && rstd && sums && dy);
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)2 * C * sizeof(double), st);
  if (e != cudaSuccess) return (int)e;
  BnBwd p;
  p.dz = dz; p.ldz = ldz; p.arg = argmax_or_null; p.ns = ns; p.y = y; p.ldy = ldy;
  p.ka = ka; p.kb = kb; p.mean = mean; p.rstd = rstd; p.R = R; p.C = C;
  if (argmax_or_null != nullptr && R % ns == 0 && TO_THREADS % (C / 4) == 0) {
    const long long groups = R / ns;
    bn_relu_bwd_pooled_reduce_kernel<<<grid_rows(groups, 16), TO_THREADS, 0, st>>>(p, sums); PTT_LAUNCHED();
    bn_relu_bwd_pooled_apply_kernel<<<grid_rows(groups * (C / 4), TO_THREADS), TO_THREADS, 0, st>>>(p, sums, gamma, dy, ld_dy,
                                                                                                  dparam_or_null); PTT_LAUNCHED();
    return ptt_launch_status();
  }
  bn_relu_bwd_reduce_kernel<<<grid_rows(R, 64), TO_THREADS, 0, st>>>(p, sums); PTT_LAUNCHED();
  bn_relu_bwd_apply_kernel<<<grid_rows(R * (C / 4), TO_THREADS), TO_THREADS, 0, st>>>(p, sums, gamma, dy, ld_dy, dparam_or_null); PTT_LAUNCHED();
  return ptt_launch_status();
}
