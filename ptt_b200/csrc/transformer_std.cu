// TransformerBlockSTD (transformer_block/variants.py:12-40): dense n x n dot-product attention per cloud --
// the literal "QKV / softmax / AV" block of the registry (not selected by the shipped YAMLs).
//
//   x = fc1(f); q, k, v = Wq x, Wk x, Wv x;  attn = softmax(q k^T / sqrt(d), over keys);
//   res = attn (v + fc_delta(xyz));  out = fc2(res) + f;  returns (out, attn (B, n, n))
//
// Both attention contractions run on the tensor cores as BATCHED row-block contractions (tc_gemm.cu, one batch element
// per cloud, fp16 hi/lo split -> fp32-class accuracy): the per-cloud K and (V + pos)^T operands are re-packed into the
// UMMA weight-image layout on the fly.  The row softmax is a warp-per-row kernel.
#include <math_constants.h>

#include "gemm.cuh"

namespace {

struct StdLayout {
  int dp, dm, ld;
  size_t fc1, qkv, delta0, delta2, fc2, total;   // linear images: transposed fp32 weight + bias row + tcgen05 image
};

bool std_layout(int dp, int dm, StdLayout* L) {
  if (dp < 1 || dm < 1) return false;
  L->dp = dp; L->dm = dm; L->ld = round_up(dm, 4);
  size_t off = 0;
  auto take = [&](int K, int Cout) {
    size_t o = off;
    off += align_up((size_t)(K + 1) * round_up(Cout, 4), 4) + ptt_tc_weight_floats(K, Cout);
    return o;
  };
  L->fc1 = take(dp, dm);
  L->qkv = take(dm, 3 * L->ld);
  L->delta0 = take(3, dm);
  L->delta2 = take(dm, dm);
  L->fc2 = take(dm, dp);
  L->total = off;
  return true;
}

struct StdWorkspace {
  int ld, npad;
  size_t x, qkv, h0, vp, s, res, kimg, vimg, kimg_bytes, vimg_bytes, total;   // float offsets; images in floats too
};

void std_workspace(int B, int n, const StdLayout& L, StdWorkspace* W) {
  const size_t tokens = (size_t)B * n;
  W->ld = L.ld;
  W->npad = round_up(n, 4);
  size_t off = 0;
  auto take = [&](size_t cnt) { size_t o = off; off += align_up(cnt, 64); return o; };
  W->x = take(tokens * W->ld);
  W->qkv = take(tokens * 3 * W->ld);
  W->h0 = take(tokens * W->ld);
  W->vp = take(tokens * W->ld);
  W->res = take(tokens * W->ld);
  W->s = take(tokens * W->npad);
  W->kimg_bytes = ptt_tc_weight_halves(L.dm, n) * 2;          // per cloud: (Cout = n keys, K = dm)
  W->vimg_bytes = ptt_tc_weight_halves(n, L.dm) * 2;          // per cloud: (Cout = dm channels, K = n keys)
  W->kimg = take((size_t)B * W->kimg_bytes / 4);
  W->vimg = take((size_t)B * W->vimg_bytes / 4);
  W->total = off;
}

// h0[t, c] = relu(Wd0[c, :] . xyz_t + bd0[c])        (fc_delta.0 on the absolute coordinates)
__global__ void __launch_bounds__(256) std_delta0_kernel(const float* __restrict__ xyz, const float* __restrict__ img, int dm,
                                                          int ldw, long long tokens, float* __restrict__ h, int ldh) {
  const long long total = tokens * dm;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long t = e / dm;
    const int c = (int)(e - t * dm);
    const float* p = xyz + t * 3;
    float v = __ldg(img + 3 * ldw + c);
    v = fmaf(__ldg(p), __ldg(img + c), v);
    v = fmaf(__ldg(p + 1), __ldg(img + ldw + c), v);
    v = fmaf(__ldg(p + 2), __ldg(img + 2 * ldw + c), v);
    h[(size_t)t * ldh + c] = fmaxf(v, 0.f);
  }
}

// one warp per (cloud, query) row: p = softmax(s / divisor) over the n keys; written in place (padding columns = 0)
// and, when requested, to the dense (B, n, n) attention output
__global__ void __launch_bounds__(256) std_softmax_kernel(float* __restrict__ s, int n, int npad, float divisor, long long rows,
                                                           float* __restrict__ attn) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = warp; row < rows; row += nwarps) {
    float* p = s + (size_t)row * npad;
    float m = -CUDART_INF_F;
    for (int j = lane; j < n; j += 32) m = fmaxf(m, p[j] / divisor);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) sum += expf(p[j] / divisor - m);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    for (int j = lane; j < npad; j += 32) {
      const float v = j < n ? expf(p[j] / divisor - m) / sum : 0.f;
      p[j] = v;
      if (attn && j < n) attn[(size_t)row * n + j] = v;
    }
  }
}

}  // namespace

extern "C" size_t ptt_transformer_std_params_floats(int d_points, int d_model) {
  StdLayout L;
  return std_layout(d_points, d_model, &L) ? L.total : 0;
}

extern "C" int ptt_transformer_std_pack_params(int d_points, int d_model, const float* fc1_w, const float* fc1_b,
                                               const float* fc2_w, const float* fc2_b, const float* delta0_w,
                                               const float* delta0_b, const float* delta2_w, const float* delta2_b,
                                               const float* wq, const float* wk, const float* wv, float* params,
                                               ptt_stream_t stream) {
  StdLayout L;
  PTT_CHECK_ARG(std_layout(d_points, d_model, &L) && params);
  PTT_CHECK_ARG(fc1_w && fc2_w && delta0_w && delta2_w && wq && wk && wv);
  cudaStream_t st = as_stream(stream);
  const int dp = d_points, dm = d_model, ld = L.ld;
  cudaError_t e = cudaMemsetAsync(params, 0, L.total * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  int rc;
  if ((rc = ptt_linear_pack_cols(fc1_w, fc1_b, dp, dm, ld, 0, params + L.fc1, st))) return rc;
  if ((rc = ptt_linear_pack_cols(wq, nullptr, dm, dm, 3 * ld, 0, params + L.qkv, st))) return rc;
  if ((rc = ptt_linear_pack_cols(wk, nullptr, dm, dm, 3 * ld, ld, params + L.qkv, st))) return rc;
  if ((rc = ptt_linear_pack_cols(wv, nullptr, dm, dm, 3 * ld, 2 * ld, params + L.qkv, st))) return rc;
  if ((rc = ptt_linear_pack_cols(delta0_w, delta0_b, 3, dm, ld, 0, params + L.delta0, st))) return rc;
  if ((rc = ptt_linear_pack_cols(delta2_w, delta2_b, dm, dm, ld, 0, params + L.delta2, st))) return rc;
  if ((rc = ptt_linear_pack_cols(fc2_w, fc2_b, dm, dp, round_up(dp, 4), 0, params + L.fc2, st))) return rc;
  auto tc_pack = [&](size_t img, int K, int Cout) {
    const int ldw = round_up(Cout, 4);
    return ptt_tc_pack_weight(params + img, 1, ldw, Cout, K, params + img + (size_t)(K + 1) * ldw, st);
  };
  if ((rc = tc_pack(L.fc1, dp, dm))) return rc;
  if ((rc = tc_pack(L.qkv, dm, 3 * ld))) return rc;
  if ((rc = tc_pack(L.delta2, dm, dm))) return rc;
  if ((rc = tc_pack(L.fc2, dm, dp))) return rc;
  return ptt_launch_status();
}

extern "C" size_t ptt_transformer_std_workspace_bytes(int B, int n, int d_points, int d_model) {
  StdLayout L;
  StdWorkspace W;
  if (B <= 0 || n <= 0 || !std_layout(d_points, d_model, &L)) return 0;
  std_workspace(B, n, L, &W);
  return W.total * sizeof(float);
}

extern "C" int ptt_transformer_std_fwd(const float* xyz, const float* features, int B, int n, int d_points, int d_model,
                                       const float* params, float* out, float* attn_or_null, void* workspace,
                                       size_t workspace_bytes, ptt_stream_t stream) {
  StdLayout L;
  PTT_CHECK_ARG(B >= 0 && n >= 1 && std_layout(d_points, d_model, &L));
  if (B == 0) return PTT_OK;
  PTT_CHECK_ARG(xyz && features && params && out);
  StdWorkspace W;
  std_workspace(B, n, L, &W);
  if (workspace == nullptr || workspace_bytes < W.total * sizeof(float)) return PTT_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 15u) != 0) return PTT_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  float* ws = static_cast<float*>(workspace);
  const int dp = d_points, dm = d_model, ld = W.ld, ldq = 3 * W.ld, npad = W.npad;
  const long long tokens = (long long)B * n;
  float *x = ws + W.x, *qkv = ws + W.qkv, *h0 = ws + W.h0, *vp = ws + W.vp, *s = ws + W.s, *res = ws + W.res;
  void* kimg = ws + W.kimg;
  void* vimg = ws + W.vimg;
  int rc;

  auto linear = [&](const float* in, int ldin, long long R, int K, size_t img, int Cout, int ldw, bool bias, int relu,
                    const float* residual, int ldr, float* y, int ldy) {
    PttGemmArgs g;
    g.x = in; g.ldx = ldin; g.R = (int)R; g.K = K;
    g.wt = params + img; g.ldw = ldw; g.N = Cout;
    g.shift = bias ? params + img + (size_t)K * ldw : nullptr;
    g.wimg = params + img + (size_t)(K + 1) * ldw;
    g.relu = relu; g.residual = residual; g.ldr = ldr; g.y = y; g.ldy = ldy;
    return ptt_gemm_launch(g, st);
  };
  if ((rc = linear(features, dp, tokens, dp, L.fc1, dm, ld, true, 0, nullptr, 0, x, ld))) return rc;
  if ((rc = linear(x, ld, tokens, dm, L.qkv, 3 * ld, 3 * ld, false, 0, nullptr, 0, qkv, ldq))) return rc;
  {
    const long long total = tokens * dm;
    std_delta0_kernel<<<(unsigned)llmin_((total + 255) / 256, 148LL * 32), 256, 0, st>>>(xyz, params + L.delta0, dm, ld, tokens, h0,
                                                                                           ld); PTT_LAUNCHED();
  }
  // vp = v + fc_delta.2(h0)
  if ((rc = linear(h0, ld, tokens, dm, L.delta2, dm, ld, true, 0, qkv + 2 * ld, ldq, vp, ld))) return rc;

  // scores: per cloud S = q . k^T  (keys as the "weight": Cout = n, K = dm, src(c, kk) = qkv[(b*n + c)*ldq + ld + kk])
  if ((rc = ptt_tc_pack_weight(qkv + ld, ldq, 1, n, dm, kimg, st, nullptr, B, (long long)n * ldq, W.kimg_bytes))) return rc;
  {
    PttGemmArgs g;
    g.x = qkv; g.ldx = ldq; g.R = n; g.K = dm; g.N = n; g.wimg = kimg;
    g.y = s; g.ldy = npad;
    g.batch = B; g.x_bstride = (long long)n * ldq; g.y_bstride = (long long)n * npad; g.wimg_bstride = W.kimg_bytes;
    if ((rc = ptt_gemm_launch(g, st))) return rc;
  }
  {
    const long long blocks = llmin_((tokens + 7) / 8, 148LL * 16);
    std_softmax_kernel<<<(unsigned)blocks, 256, 0, st>>>(s, n, npad, sqrtf((float)dm), tokens, attn_or_null); PTT_LAUNCHED();
  }
  // res: per cloud P . (v + pos)  ((v+pos)^T as the "weight": Cout = dm, K = n, src(c, kk) = vp[(b*n + kk)*ld + c])
  if ((rc = ptt_tc_pack_weight(vp, 1, ld, dm, n, vimg, st, nullptr, B, (long long)n * ld, W.vimg_bytes))) return rc;
  {
    PttGemmArgs g;
    g.x = s; g.ldx = npad; g.R = n; g.K = n; g.N = dm; g.wimg = vimg;
    g.y = res; g.ldy = ld;
    g.batch = B; g.x_bstride = (long long)n * npad; g.y_bstride = (long long)n * ld; g.wimg_bstride = W.vimg_bytes;
    if ((rc = ptt_gemm_launch(g, st))) return rc;
  }
  return linear(res, ld, tokens, dm, L.fc2, dp, round_up(dp, 4), true, 0, features, dp, out, dp);
}
