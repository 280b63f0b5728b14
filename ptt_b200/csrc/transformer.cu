// Point-Track-Transformer block (kNN vector attention) for sm_100a, forward.
//   TransformerBlock.forward          transformer_block/variants.py:149-165
//   TransformerBlockOffset.forward    transformer_block/variants.py:319-334  (variant 1: fc2(x - res))
//
//   x = fc1(f); q,k,v = Wq x, Wk x, Wv x                              (rows: B*n tokens)
//   pos_ij = fc_delta(xyz_i - xyz_j),  j in knn(i)                    (rows: B*n*k pairs)
//   a_ij   = fc_gamma(q_i - k_j + pos_ij);  p = softmax_j(a / sqrt(d_model))   per channel
//   res_i  = sum_j p_ij * (v_j + pos_ij);   out = fc2(res) + f
//
// Stage 1 (this file, v1): token- and pair-major activations in workspace, the dense contractions
// through gemm.cuh, the gathers / softmax / weighted sum in fused element kernels.  The (B,n,n)
// distance matrix, its argsort and the three (B,n,k,d) gathered tensors of the reference are never built.
#include "gemm.cuh"
#include "tr_fused.cuh"

namespace {

struct TrLayout {
  int dp, dm;
  // linear images (wt rows + bias row), in floats from the start of the params block
  size_t fc1, qkv, delta0, delta2, gamma0, gamma2, fc2;
  // fused path (tr_fused.cu): fc_gamma.0 is linear, so its q / k / pos parts are pre-multiplied at pack time:
  //   qkg   : x -> [Wg0.Wq x | Wg0.Wk x | Wv x]      (replaces the q, k, v projection)
  //   wprime: (Wg0.Wd2) as tcgen05 image, cprime = Wg0.bd2 + bg0
  //   scratch: the three d x d products in fp32 (row-major), kept for inspection
  //   qkg1  : f -> [Wg0.Wq.W1 f | Wg0.Wk.W1 f | Wv.W1 f] + folded biases: fc1 is linear and x = fc1(f) is only used through
  //           q, k, v (variants.py:152-153), so the token projection is ONE d_points -> 3 d_model contraction
  //   scratch2: temporaries of that fold (two dm x dp products, two dm vectors)
  size_t qkg, wprime, cprime, scratch, qkg1, scratch2, total;
};

bool tr_layout(int dp, int dm, TrLayout* L) {
  if (dp < 1 || dm < 1) return false;
  L->dp = dp; L->dm = dm;
  size_t off = 0;
  // each image: transposed fp32 weight (K rows) + bias row, then the fp16 hi/lo tensor-core image
  auto take = [&](int K, int Cout) {
    size_t o = off;
    off += align_up((size_t)(K + 1) * round_up(Cout, 4), 4) + ptt_tc_weight_floats(K, Cout);
    return o;
  };
  L->fc1 = take(dp, dm);
  L->qkv = take(dm, 3 * round_up(dm, 4));
  L->delta0 = take(3, dm);
  L->delta2 = take(dm, dm);
  L->gamma0 = take(dm, dm);
  L->gamma2 = take(dm, dm);
  L->fc2 = take(dm, dp);
  L->qkg = take(dm, 3 * round_up(dm, 4));
  L->wprime = off; off += ptt_tc_weight_floats(dm, dm);
  L->cprime = off; off += round_up(dm, 4);
  L->scratch = off; off += (size_t)3 * dm * dm;
  L->qkg1 = take(dp, 3 * round_up(dm, 4));
  L->scratch2 = off; off += (size_t)2 * dm * round_up(dp, 4) + 2 * round_up(dm, 4);
  L->total = off;
  return true;
}

// cprime[c] = sum_m Wg0[c, m] * bd2[m] + bg0[c]
__global__ void tr_cprime_kernel(const float* __restrict__ wg0, const float* __restrict__ bd2, const float* __restrict__ bg0, int dm,
                                 float* __restrict__ cprime) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= dm) return;
  float acc = bg0 ? bg0[c] : 0.f;
  if (bd2)
    for (int m = 0; m < dm; ++m) acc = fmaf(wg0[(size_t)c * dm + m], bd2[m], acc);
  cprime[c] = acc;
}

// out[c] = sum_m w[c * ldw + m] * v[m]   (v == nullptr: out = 0); one warp per output row
__global__ void tr_matvec_kernel(const float* __restrict__ w, int ldw, const float* __restrict__ v, int rows, int K,
                                 float* __restrict__ out) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= rows) return;
  float acc = 0.f;
  if (v)
    for (int m = lane; m < K; m += 32) acc = fmaf(w[(size_t)c * ldw + m], v[m], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[c] = acc;
}

// h[(b,i,j), c] = relu(Wd0[c,:] . (xyz_i - xyz_knn(i,j)) + bd0[c])      delta0 image: 3 rows wt + bias row
__global__ void __launch_bounds__(256) tr_delta0_kernel(const float* __restrict__ xyz, const int* __restrict__ knn,
                                                         const float* __restrict__ img, int n, int k, int dm, int ldw,
                                                         long long pairs, float* __restrict__ h, int ldh) {
  const long long total = pairs * dm;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long pr = e / dm;
    const int c = (int)(e - pr * dm);
    const long long tok = pr / k;            // b * n + i
    const long long b = tok / n;
    const int j = __ldg(knn + pr);
    const float* pi = xyz + tok * 3;
    const float* pj = xyz + (b * n + j) * 3;
    const float dx = __ldg(pi) - __ldg(pj), dy = __ldg(pi + 1) - __ldg(pj + 1), dz = __ldg(pi + 2) - __ldg(pj + 2);
    float v = __ldg(img + 3 * ldw + c);
    v = fmaf(dx, __ldg(img + c), v);
    v = fmaf(dy, __ldg(img + ldw + c), v);
    v = fmaf(dz, __ldg(img + 2 * ldw + c), v);
    h[(size_t)pr * ldh + c] = fmaxf(v, 0.f);
  }
}

// a[(b,i,j), c] = q[(b,i), c] - kk[(b,knn), c] + pos[(b,i,j), c]
__global__ void __launch_bounds__(256) tr_attn_in_kernel(const float* __restrict__ qkv, int ldq, int koff,
                                                          const int* __restrict__ knn, const float* __restrict__ pos,
                                                          int n, int k, int dm, long long pairs, float* __restrict__ a,
                                                          int ld) {
  const long long total = pairs * dm;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long pr = e / dm;
    const int c = (int)(e - pr * dm);
    const long long tok = pr / k;
    const long long b = tok / n;
    const int j = __ldg(knn + pr);
    const float q = __ldg(qkv + tok * ldq + c);
    const float kk = __ldg(qkv + (b * n + j) * ldq + koff + c);
    a[(size_t)pr * ld + c] = (q - kk) + __ldg(pos + (size_t)pr * ld + c);
  }
}

// per (token, channel): p = softmax_j(logit / sqrt(dm)); res = sum_j p * (v[knn] + pos)
__global__ void __launch_bounds__(256) tr_softmax_agg_kernel(const float* __restrict__ logit, const float* __restrict__ pos,
                                                              int ld, const float* __restrict__ qkv, int ldq, int voff,
                                                              const int* __restrict__ knn, int n, int k, int dm,
                                                              float divisor, long long tokens, const float* __restrict__ x_sub,
                                                              float* __restrict__ res, int ldres, float* __restrict__ attn) {
  const long long total = tokens * dm;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long tok = e / dm;
    const int c = (int)(e - tok * dm);
    const long long b = tok / n;
    const float* lg = logit + (size_t)tok * k * ld + c;
    float m = -INFINITY;
    for (int j = 0; j < k; ++j) m = fmaxf(m, lg[(size_t)j * ld] / divisor);
    float s = 0.f;
    for (int j = 0; j < k; ++j) s += expf(lg[(size_t)j * ld] / divisor - m);
    float acc = 0.f;
    for (int j = 0; j < k; ++j) {
      const float p = expf(lg[(size_t)j * ld] / divisor - m) / s;
      const int nb = __ldg(knn + tok * k + j);
      const float v = __ldg(qkv + (b * n + nb) * ldq + voff + c);
      acc = fmaf(p, v + __ldg(pos + ((size_t)tok * k + j) * ld + c), acc);
      if (attn) attn[((size_t)tok * k + j) * dm + c] = p;
    }
    if (x_sub) acc = __ldg(x_sub + tok * (long long)ldres + c) - acc;  // Offset variant: fc2(x - res)
    res[(size_t)tok * ldres + c] = acc;
  }
}

struct TrWorkspace {
  int ld;      // row stride of dm-wide activations
  int ldq;     // row stride of the qkv block (3 * ld)
  size_t knn, x, qkv, h, pos, a, res, total;  // float offsets
};

void tr_workspace(int B, int n, int k, const TrLayout& L, TrWorkspace* W) {
  const size_t tokens = (size_t)B * n, pairs = tokens * k;
  W->ld = round_up(L.dm, 4);
  W->ldq = 3 * W->ld;
  size_t off = 0;
  auto take = [&](size_t cnt) { size_t o = off; off += align_up(cnt, 64); return o; };
  W->knn = take(pairs);
  W->x = take(tokens * W->ld);
  W->qkv = take(tokens * W->ldq);
  W->res = take(tokens * W->ld);
  W->h = take(pairs * W->ld);
  W->pos = take(pairs * W->ld);
  W->a = take(pairs * W->ld);
  W->total = off;
}

inline unsigned grid_for(long long total) { return (unsigned)llmin_((total + 255) / 256, 148LL * 32); }

}  // namespace

extern "C" size_t ptt_transformer_params_floats(int d_points, int d_model) {
  TrLayout L;
  return tr_layout(d_points, d_model, &L) ? L.total : 0;
}

extern "C" int ptt_transformer_pack_params(int d_points, int d_model, const float* fc1_w, const float* fc1_b,
                                           const float* fc2_w, const float* fc2_b, const float* delta0_w,
                                           const float* delta0_b, const float* delta2_w, const float* delta2_b,
                                           const float* gamma0_w, const float* gamma0_b, const float* gamma2_w,
                                           const float* gamma2_b, const float* wq, const float* wk, const float* wv,
                                           float* params, ptt_stream_t stream) {
  TrLayout L;
  PTT_CHECK_ARG(tr_layout(d_points, d_model, &L) && params);
  PTT_CHECK_ARG(fc1_w && fc2_w && delta0_w && delta2_w && gamma0_w && gamma2_w && wq && wk && wv);
  cudaStream_t st = as_stream(stream);
  const int dp = d_points, dm = d_model, ld = round_up(dm, 4);
  cudaError_t e = cudaMemsetAsync(params, 0, L.total * sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  int rc;
  if ((rc = ptt_linear_pack_cols(fc1_w, fc1_b, dp, dm, ld, 0, params + L.fc1, st))) return rc;
  if ((rc = ptt_linear_pack_cols(wq, nullptr, dm, dm, 3 * ld, 0, params + L.qkv, st))) return rc;
  if ((rc = ptt_linear_pack_cols(wk, nullptr, dm, dm, 3 * ld, ld, params + L.qkv, st))) return rc;
  if ((rc = ptt_linear_pack_cols(wv, nullptr, dm, dm, 3 * ld, 2 * ld, params + L.qkv, st))) return rc;
  if ((rc = ptt_linear_pack_cols(delta0_w, delta0_b, 3, dm, ld, 0, params + L.delta0, st))) return rc;
  if ((rc = ptt_linear_pack_cols(delta2_w, delta2_b, dm, dm, ld, 0, params + L.delta2, st))) return rc;
  if ((rc = ptt_linear_pack_cols(gamma0_w, gamma0_b, dm, dm, ld, 0, params + L.gamma0, st))) return rc;
  if ((rc = ptt_linear_pack_cols(gamma2_w, gamma2_b, dm, dm, ld, 0, params + L.gamma2, st))) return rc;
  if ((rc = ptt_linear_pack_cols(fc2_w, fc2_b, dm, dp, round_up(dp, 4), 0, params + L.fc2, st))) return rc;
  // tensor-core images, built from the transposed fp32 images just written (src(c,k) = wt[k*ldw + c])
  auto tc_pack = [&](size_t img, int K, int Cout) {
    const int ldw = round_up(Cout, 4);
    return ptt_tc_pack_weight(params + img, 1, ldw, Cout, K, params + img + (size_t)(K + 1) * ldw, st);
  };
  if ((rc = tc_pack(L.fc1, dp, dm))) return rc;
  if ((rc = tc_pack(L.qkv, dm, 3 * ld))) return rc;
  if ((rc = tc_pack(L.delta2, dm, dm))) return rc;
  if ((rc = tc_pack(L.gamma0, dm, dm))) return rc;
  if ((rc = tc_pack(L.gamma2, dm, dm))) return rc;
  if ((rc = tc_pack(L.fc2, dm, dp))) return rc;
  if (tr_fused_supported(1, 1, dm)) {
    // products P[c, k] = sum_m Wg0[c, m] * W[m, k]: a row-block contraction with x = Wg0 (rows c) and the (m, k) matrix
    // W used directly as the "transposed weight" image
    float* sc = params + L.scratch;
    auto product = [&](const float* w, float* dst) {
      PttGemmArgs g;
      g.x = gamma0_w; g.ldx = dm; g.R = dm; g.K = dm;
      g.wt = w; g.ldw = dm; g.N = dm;
      g.y = dst; g.ldy = dm;
      return ptt_gemm_launch_ffma(g, st);
    };
    if ((rc = product(delta2_w, sc))) return rc;
    if ((rc = product(wq, sc + (size_t)dm * dm))) return rc;
    if ((rc = product(wk, sc + (size_t)2 * dm * dm))) return rc;
    if ((rc = ptt_tc_pack_weight(sc, dm, 1, dm, dm, params + L.wprime, st))) return rc;
    tr_cprime_kernel<<<ceil_div(dm, 128), 128, 0, st>>>(gamma0_w, delta2_b, gamma0_b, dm, params + L.cprime); PTT_LAUNCHED();
    if ((rc = ptt_linear_pack_cols(sc + (size_t)dm * dm, nullptr, dm, dm, 3 * ld, 0, params + L.qkg, st))) return rc;
    if ((rc = ptt_linear_pack_cols(sc + (size_t)2 * dm * dm, nullptr, dm, dm, 3 * ld, ld, params + L.qkg, st))) return rc;
    if ((rc = ptt_linear_pack_cols(wv, nullptr, dm, dm, 3 * ld, 2 * ld, params + L.qkg, st))) return rc;
    if ((rc = tc_pack(L.qkg, dm, 3 * ld))) return rc;
    // fold fc1 into the token projection: P = W . W1 (dm x dp), bias = W . b1, for W in {Wg0.Wq, Wg0.Wk, Wv}
    const int ldp = round_up(dp, 4);
    float* T = params + L.scratch2;                 // fc1_w copied to 16-byte aligned rows of ldp floats (any d_points)
    float* P = T + (size_t)dm * ldp;
    float* bv = P + (size_t)dm * ldp;
    e = cudaMemcpy2DAsync(T, (size_t)ldp * sizeof(float), fc1_w, (size_t)dp * sizeof(float), (size_t)dp * sizeof(float), dm,
                          cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return (int)e;
    const float* mats[3] = {sc + (size_t)dm * dm, sc + (size_t)2 * dm * dm, wv};          // Wg0.Wq, Wg0.Wk, Wv (dm x dm, row-major)
    for (int i = 0; i < 3; ++i) {
      PttGemmArgs g;                                // P (dm x dp, row stride dp) = mats[i] (dm x dm) . fc1_w (dm x dp)
      g.x = mats[i]; g.ldx = dm; g.R = dm; g.K = dm;
      g.wt = T; g.ldw = ldp; g.N = dp;
      g.y = P; g.ldy = dp;
      if ((rc = ptt_gemm_launch_ffma(g, st))) return rc;
      tr_matvec_kernel<<<ceil_div(dm * 32, 128), 128, 0, st>>>(mats[i], dm, fc1_b, dm, dm, bv); PTT_LAUNCHED();
      if ((rc = ptt_linear_pack_cols(P, bv, dp, dm, 3 * ld, i * ld, params + L.qkg1, st))) return rc;
    }
    if ((rc = tc_pack(L.qkg1, dp, 3 * ld))) return rc;
  }
  return ptt_launch_status();
}

extern "C" size_t ptt_transformer_block_workspace_bytes(int B, int n, int k, int d_points, int d_model) {
  TrLayout L;
  TrWorkspace W;
  if (B <= 0 || n <= 0 || k <= 0 || !tr_layout(d_points, d_model, &L)) return 0;
  tr_workspace(B, n, k, L, &W);
  return W.total * sizeof(float);
}

// Float offsets of the activations the block leaves in its workspace (the training path's backward reads g and pos + v
// from it): h_out[0..6] = knn, x, qkv, res, g (= relu(fc_gamma.0(.)), pairs x ld), pos + v (pairs x ld), ld
extern "C" int ptt_transformer_block_workspace_layout(int B, int n, int k, int d_points, int d_model, size_t* h_out) {
  TrLayout L;
  TrWorkspace W;
  PTT_CHECK_ARG(B > 0 && n > 0 && k > 0 && h_out && tr_layout(d_points, d_model, &L));
  tr_workspace(B, n, k, L, &W);
  h_out[0] = W.knn; h_out[1] = W.x; h_out[2] = W.qkv; h_out[3] = W.res; h_out[4] = W.h; h_out[5] = W.pos; h_out[6] = (size_t)W.ld;
  return PTT_OK;
}

namespace {
struct TrExtra {
  const float* q_features = nullptr;   // CrossAttentionBlock: the queries come from these (B, n, d_points) features
  float divisor = 0.f;                 // softmax temperature; 0 = sqrt(d_model)
  const float* pair_scalar = nullptr;  // (B, n, k) and (d_model): pre-activation of fc_gamma.0 += pair_scalar * pair_vec
  const float* pair_vec = nullptr;
};
}  // namespace

// variant: bit 0 = Offset (fc2(x - res)), bit 1 = raw (out (B, n, d_model) = the aggregated res; no fc2, no residual)
static int tr_block_impl(const float* xyz, const float* features, int B, int n, int k, int d_points, int d_model,
                         int variant_flags, const float* params, const int* knn_idx_or_null, float* out, float* attn_or_null,
                         void* workspace, size_t workspace_bytes, ptt_stream_t stream, const TrExtra& ex) {
  TrLayout L;
  PTT_CHECK_ARG(B >= 0 && n >= 1 && k >= 1 && k <= n && tr_layout(d_points, d_model, &L));
  PTT_CHECK_ARG(variant_flags >= 0 && variant_flags <= 3);
  PTT_CHECK_ARG((ex.pair_scalar == nullptr) == (ex.pair_vec == nullptr));
  const int variant = variant_flags & 1;
  const bool raw = (variant_flags & 2) != 0;
  if (B == 0) return PTT_OK;
  PTT_CHECK_ARG(xyz && features && params && out);
  TrWorkspace W;
  tr_workspace(B, n, k, L, &W);
  if (workspace == nullptr || workspace_bytes < W.total * sizeof(float)) return PTT_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 31u) != 0) return PTT_ERR_WORKSPACE;   // 256-bit loads / stores of workspace rows
  cudaStream_t st = as_stream(stream);
  float* ws = static_cast<float*>(workspace);
  const int dp = d_points, dm = d_model, ld = W.ld, ldq = W.ldq;
  const long long tokens = (long long)B * n, pairs = tokens * k;
  int rc;

  const int* knn = knn_idx_or_null;
  if (knn == nullptr) {
    int* own = reinterpret_cast<int*>(ws + W.knn);
    if ((rc = ptt_knn(xyz, B, n, k, own, stream))) return rc;
    knn = own;
  }
  float* x = ws + W.x;
  float* qkv = ws + W.qkv;
  float* h = ws + W.h;
  float* pos = ws + W.pos;
  float* a = ws + W.a;
  float* res = ws + W.res;

  auto linear = [&](const float* in, int ldin, long long R, int K, size_t img, int Cout, int ldw, bool bias, int relu,
                    const float* residual, int ldr, float* y, int ldy) {
    PttGemmArgs g;
    g.x = in; g.ldx = ldin; g.R = (int)R; g.K = K;
    g.wt = params + img; g.ldw = ldw; g.N = Cout;
    g.shift = bias ? params + img + (size_t)K * ldw : nullptr;
    g.wimg = params + img + (size_t)(K + 1) * ldw;
    g.relu = relu;
    g.residual = residual; g.ldr = ldr;
    g.y = y; g.ldy = ldy;
    return ptt_gemm_launch(g, st);
  };

  const bool fused = tr_fused_supported(n, k, dm);
  // fused path, plain variant: fc1 is folded into the projection (x itself is never needed); Offset variant: x is kept
  // for fc2(x - res).  fused: [Wg0.Wq x | Wg0.Wk x | Wv x]; generic path: [q | k | v]
  if (ex.pair_scalar != nullptr && !fused) return PTT_ERR_UNSUPPORTED;   // the rank-1 term exists in the tcgen05 passes only
  const bool folded = fused && variant == 0;
  if (folded) {
    if ((rc = linear(features, dp, tokens, dp, L.qkg1, 3 * ld, 3 * ld, true, 0, nullptr, 0, qkv, ldq))) return rc;
    // queries from other features: redo the first ld columns (the weight images are output-block major)
    if (ex.q_features && (rc = linear(ex.q_features, dp, tokens, dp, L.qkg1, ld, 3 * ld, true, 0, nullptr, 0, qkv, ldq))) return rc;
  } else {
    if ((rc = linear(features, dp, tokens, dp, L.fc1, dm, ld, true, 0, nullptr, 0, x, ld))) return rc;
    if ((rc = linear(x, ld, tokens, dm, fused ? L.qkg : L.qkv, 3 * ld, 3 * ld, false, 0, nullptr, 0, qkv, ldq))) return rc;
    if (ex.q_features) {
      if (variant == 1) return PTT_ERR_UNSUPPORTED;      // `res` doubles as the scratch for fc1(q_features); Offset needs x
      if ((rc = linear(ex.q_features, dp, tokens, dp, L.fc1, dm, ld, true, 0, nullptr, 0, res, ld))) return rc;
      if ((rc = linear(res, ld, tokens, dm, fused ? L.qkg : L.qkv, ld, 3 * ld, false, 0, nullptr, 0, qkv, ldq))) return rc;
    }
  }
  const float divisor = ex.divisor > 0.f ? ex.divisor : sqrtf((float)dm);
  float* res_out = raw ? out : res;            // raw: the aggregation writes the caller's (B*n, dm) buffer directly
  const int ld_res = raw ? dm : ld;

  if (fused) {
    // ---- pair-row passes on the tensor cores with generated A operands and fused reductions (tr_fused.cu)
    auto img_w = [&](size_t img) { return params + img + (size_t)(dm + 1) * ld; };   // tcgen05 image behind the fp32 one
    auto img_b = [&](size_t img) { return params + img + (size_t)dm * ld; };         // bias row
    TrPassArgs t;
    t.n = n; t.k = k; t.dm = dm; t.pairs = pairs; t.xyz = xyz; t.knn = knn;
    t.qkv = qkv; t.ldq = ldq; t.koff = ld; t.voff = 2 * ld; t.divisor = divisor;
    // pass 1: pos = fc_delta.2(relu(fc_delta.0(xyz_i - xyz_j))); what is stored is pos_ij + v_j, the only form in which
    // pos is used again (by the aggregation of pass 3)
    TrPassArgs p1 = t;
    p1.wd0 = params + L.delta0; p1.ldw0 = ld;
    p1.wimg = img_w(L.delta2); p1.bias = img_b(L.delta2); p1.relu = 0; p1.out = pos; p1.ldo = ld;
    p1.koff = 2 * ld; p1.qk_mode = 1;
    if ((rc = tr_fused_launch(p1, TR_PROD_DELTA0, TR_EPI_STORE_QK, st))) return rc;
    // pass 2: g = relu(fc_gamma.0(q_i - k_j + pos)) = relu((Wg0.Wd2) h_ij + (Wg0 q)_i - (Wg0 k)_j + (Wg0 bd2 + bg0)):
    // the A operand is the same generated h as in pass 1 (no loads); the per-token q / k parts are added in the epilogue
    TrPassArgs p2 = t;
    p2.wd0 = params + L.delta0; p2.ldw0 = ld;
    p2.wimg = params + L.wprime; p2.bias = params + L.cprime; p2.relu = 1; p2.out = h; p2.ldo = ld;
    p2.row_scalar = ex.pair_scalar; p2.col_vec = ex.pair_vec;
    if ((rc = tr_fused_launch(p2, TR_PROD_DELTA0, TR_EPI_STORE_QK, st))) return rc;
    // pass 3: logits = fc_gamma.2(g); softmax over the k neighbours; res = sum p * (v + pos)
    TrPassArgs p3 = t;
    p3.a_src = h; p3.lda = ld; p3.pos = pos; p3.wimg = img_w(L.gamma2); p3.bias = img_b(L.gamma2);
    p3.out = res_out; p3.ldo = ld_res; p3.attn = attn_or_null; p3.pos_has_v = 1;
    if (variant == 1) { p3.x_sub = x; p3.ldx = ld; }
    if ((rc = tr_fused_launch(p3, TR_PROD_PLAIN, TR_EPI_SOFTMAX, st))) return rc;
    if (raw) return PTT_OK;
    return linear(res, ld, tokens, dm, L.fc2, dp, round_up(dp, 4), true, 0, features, dp, out, dp);
  }
  tr_delta0_kernel<<<grid_for(pairs * dm), 256, 0, st>>>(xyz, knn, params + L.delta0, n, k, dm, ld, pairs, h, ld); PTT_LAUNCHED();
  if ((rc = linear(h, ld, pairs, dm, L.delta2, dm, ld, true, 0, nullptr, 0, pos, ld))) return rc;
  tr_attn_in_kernel<<<grid_for(pairs * dm), 256, 0, st>>>(qkv, ldq, ld, knn, pos, n, k, dm, pairs, a, ld); PTT_LAUNCHED();
  if ((rc = linear(a, ld, pairs, dm, L.gamma0, dm, ld, true, 1, nullptr, 0, h, ld))) return rc;
  if ((rc = linear(h, ld, pairs, dm, L.gamma2, dm, ld, true, 0, nullptr, 0, a, ld))) return rc;
  tr_softmax_agg_kernel<<<grid_for(tokens * dm), 256, 0, st>>>(a, pos, ld, qkv, ldq, 2 * ld, knn, n, k, dm,
                                                                divisor, tokens, variant == 1 ? x : nullptr, res_out,
                                                                ld_res, attn_or_null); PTT_LAUNCHED();
  if ((rc = ptt_launch_status())) return rc;
  if (raw) return PTT_OK;
  return linear(res, ld, tokens, dm, L.fc2, dp, round_up(dp, 4), true, 0, features, dp, out, dp);
}

extern "C" int ptt_transformer_block_fwd(const float* xyz, const float* features, int B, int n, int k, int d_points,
                                         int d_model, int variant, const float* params, const int* knn_idx_or_null,
                                         float* out, float* attn_or_null, void* workspace, size_t workspace_bytes,
                                         ptt_stream_t stream) {
  PTT_CHECK_ARG(variant == 0 || variant == 1);
  return tr_block_impl(xyz, features, B, n, k, d_points, d_model, variant, params, knn_idx_or_null, out, attn_or_null,
                       workspace, workspace_bytes, stream, TrExtra());
}

extern "C" int ptt_transformer_block_fwd_ex(const float* xyz, const float* features, const float* q_features_or_null, int B,
                                            int n, int k, int d_points, int d_model, int variant_flags, float divisor_or_0,
                                            const float* params, const int* knn_idx_or_null,
                                            const float* pair_scalar_or_null, const float* pair_vec_or_null, float* out,
                                            float* attn_or_null, void* workspace, size_t workspace_bytes,
                                            ptt_stream_t stream) {
  TrExtra ex;
  ex.q_features = q_features_or_null;
  ex.divisor = divisor_or_0;
  ex.pair_scalar = pair_scalar_or_null;
  ex.pair_vec = pair_vec_or_null;
  return tr_block_impl(xyz, features, B, n, k, d_points, d_model, variant_flags, params, knn_idx_or_null, out, attn_or_null,
                       workspace, workspace_bytes, stream, ex);
}

// ------------------------------------------------------------------------------------------------------------------
// Small token-level kernels of the secondary registry blocks
// ------------------------------------------------------------------------------------------------------------------
namespace {

// one warp per token: sim[(b, i), j] = cos(q_i, k_{knn(i, j)})   (F.cosine_similarity, eps 1e-8; variants.py:78-79)
__global__ void __launch_bounds__(256) pair_cosine_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ kk,
                                                           int ldk, const int* __restrict__ knn, int n, int k, int d,
                                                           long long tokens, float* __restrict__ sim) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const float eps = 1e-8f;
  for (long long tok = warp; tok < tokens; tok += nwarps) {
    const long long b = tok / n;
    const float* qr = q + tok * ldq;
    float qq = 0.f;
    for (int c = lane; c < d; c += 32) { const float v = __ldg(qr + c); qq = fmaf(v, v, qq); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
    const float inv_q = 1.f / fmaxf(sqrtf(qq), eps);
    for (int j = 0; j < k; ++j) {
      const float* kr = kk + (b * n + __ldg(knn + tok * k + j)) * ldk;
      float dot = 0.f, k2 = 0.f;
      for (int c = lane; c < d; c += 32) {
        const float v = __ldg(kr + c);
        dot = fmaf(v, __ldg(qr + c), dot);
        k2 = fmaf(v, v, k2);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        dot += __shfl_xor_sync(0xffffffffu, dot, o);
        k2 += __shfl_xor_sync(0xffffffffu, k2, o);
      }
      if (lane == 0) sim[tok * k + j] = (dot * inv_q) * (1.f / fmaxf(sqrtf(k2), eps));
    }
  }
}

// one warp per row: y = LayerNorm(x) * gamma + beta (+ residual); biased variance, two passes over registers-free rows
__global__ void __launch_bounds__(256) layer_norm_kernel(const float* __restrict__ x, int ldx, long long R, int C,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float eps, const float* __restrict__ residual, int ldr,
                                                          float* __restrict__ y, int ldy) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = warp; r < R; r += nwarps) {
    const float* xr = x + r * ldx;
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s += __ldg(xr + c);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / (float)C;
    float v = 0.f;
    for (int c = lane; c < C; c += 32) { const float dlt = __ldg(xr + c) - mean; v = fmaf(dlt, dlt, v); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v / (float)C + eps);
    for (int c = lane; c < C; c += 32) {
      float o = (__ldg(xr + c) - mean) * rstd;
      o = o * (gamma ? __ldg(gamma + c) : 1.f) + (beta ? __ldg(beta + c) : 0.f);
      if (residual) o += __ldg(residual + r * ldr + c);
      y[r * ldy + c] = o;
    }
  }
}

// one thread per (cloud, channel): p = softmax over the n tokens of logits[b, :, c] / divisor; out = p * other  (variants.py:118-121)
__global__ void __launch_bounds__(128) token_softmax_gate_kernel(const float* __restrict__ logits, int ldl,
                                                                  const float* __restrict__ other, int ldo, int n, int C,
                                                                  float divisor, long long total, float* __restrict__ out,
                                                                  int ldy, float* __restrict__ attn) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / C;
    const int c = (int)(e - b * C);
    const float* lg = logits + b * n * ldl + c;
    float m = -INFINITY;
    for (int i = 0; i < n; ++i) m = fmaxf(m, lg[(size_t)i * ldl] / divisor);
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += expf(lg[(size_t)i * ldl] / divisor - m);
    for (int i = 0; i < n; ++i) {
      const float p = expf(lg[(size_t)i * ldl] / divisor - m) / s;
      out[(b * n + i) * ldy + c] = p * __ldg(other + (b * n + i) * ldo + c);
      if (attn) attn[(b * n + i) * C + c] = p;
    }
  }
}

}  // namespace

extern "C" int ptt_pair_cosine(const float* q, int ldq, const float* kmat, int ldk, const int* knn, int B, int n, int k,
                               int d, float* sim, ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && n >= 1 && k >= 1 && d >= 1 && ldq >= d && ldk >= d);
  if (B == 0) return PTT_OK;
  PTT_CHECK_ARG(q && kmat && knn && sim);
  const long long tokens = (long long)B * n;
  pair_cosine_kernel<<<(unsigned)llmin_((tokens + 7) / 8, 148LL * 8), 256, 0, as_stream(stream)>>>(q, ldq, kmat, ldk, knn, n, k, d,
                                                                                                    tokens, sim); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_layer_norm_fwd(const float* x, int ldx, int R, int C, const float* gamma, const float* beta, float eps,
                                  const float* residual_or_null, int ldr, float* y, int ldy, ptt_stream_t stream) {
  PTT_CHECK_ARG(R >= 0 && C >= 1 && ldx >= C && ldy >= C && (residual_or_null == nullptr || ldr >= C));
  if (R == 0) return PTT_OK;
  PTT_CHECK_ARG(x && y);
  layer_norm_kernel<<<(unsigned)llmin_(((long long)R + 7) / 8, 148LL * 8), 256, 0, as_stream(stream)>>>(
      x, ldx, R, C, gamma, beta, eps, residual_or_null, ldr, y, ldy); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_token_softmax_gate(const float* logits, int ldl, const float* other, int ldo, int B, int n, int C,
                                      float divisor, float* out, int ldy, float* attn_or_null, ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && n >= 1 && C >= 1 && ldl >= C && ldo >= C && ldy >= C && divisor > 0.f);
  if (B == 0) return PTT_OK;
  PTT_CHECK_ARG(logits && other && out);
  const long long total = (long long)B * C;
  token_softmax_gate_kernel<<<(unsigned)llmin_((total + 127) / 128, 148LL * 8), 128, 0, as_stream(stream)>>>(
      logits, ldl, other, ldo, n, C, divisor, total, out, ldy, attn_or_null); PTT_LAUNCHED();
  return ptt_launch_status();
}
