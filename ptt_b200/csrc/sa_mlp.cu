// Set-abstraction layer body for sm_100a, eval mode:
//   QueryAndGroup (pointnet2_utils.py:320-380) + SharedMLP (pytorch_utils.py:12-36) + max over nsample
//   (pointnet2_modules.py:83-88), with BatchNorm folded into a per-channel scale/shift.
//
// Data layout: activations are POINT-MAJOR (one contiguous row of channels per point / per (centre,
// sample) pair), so a neighbour gather is a contiguous row copy and every layer is a row-block
// contraction.  The reference's channel-major (B,C,M,ns) tensors are never built.
//
// Stage 1 (this file, v1): rows are materialised once in workspace ([feats | rel_xyz] per pair),
// the layers run through the row-block contraction of gemm.cuh, a final kernel takes the max over
// each centre's nsample rows and writes point-major and/or channel-major output.
#include "gemm.cuh"
#include "sa_fused.cuh"

namespace {

constexpr int SA_MAX_LAYERS = 4;

struct SaLayout {
  int n_layers;
  int dims[SA_MAX_LAYERS + 1];
  int ldw[SA_MAX_LAYERS];
  size_t wt[SA_MAX_LAYERS], scale[SA_MAX_LAYERS], shift[SA_MAX_LAYERS], wimg[SA_MAX_LAYERS];
  // fused path (sa_fused.cu), present when fused_dims: per-point layer-1 image, scaled xyz columns, scale-folded W2 / W3
  bool fused_dims;
  size_t g1img, wxs, w2s, w3s;
  size_t total;
};

bool sa_layout(int C, int n_layers, const int* h_dims, SaLayout* L) {
  if (C < 0 || n_layers < 1 || n_layers > SA_MAX_LAYERS || h_dims == nullptr) return false;
  if (h_dims[0] != C + 3) return false;
  L->n_layers = n_layers;
  size_t off = 0;
  for (int l = 0; l <= n_layers; ++l) {
    if (h_dims[l] < 1) return false;
    L->dims[l] = h_dims[l];
  }
  for (int l = 0; l < n_layers; ++l) {
    L->ldw[l] = round_up(L->dims[l + 1], 4);
    L->wt[l] = off;
    off += (size_t)L->dims[l] * L->ldw[l];
    L->scale[l] = off;
    off += L->ldw[l];
    L->shift[l] = off;
    off += L->ldw[l];
    L->wimg[l] = off;      // fp16 hi/lo tensor-core image of the same (permuted) weight
    off += ptt_tc_weight_floats(L->dims[l], L->dims[l + 1]);
  }
  L->fused_dims = n_layers == 3 && sa_fused_supported(L->dims[1], L->dims[2], L->dims[3], 32);
  L->g1img = L->wxs = L->w2s = L->w3s = 0;
  if (L->fused_dims) {
    L->g1img = off; off += C > 0 ? ptt_tc_weight_floats(C, L->dims[1]) : 0;
    L->wxs = off;   off += (size_t)3 * L->dims[1];
    L->w2s = off;   off += ptt_tc_weight_floats(L->dims[1], L->dims[2]);
    L->w3s = off;   off += ptt_tc_weight_floats(L->dims[2], L->dims[3]);
  }
  L->total = off;
  return true;
}

// wxs[d, c] = scale1[c] * wt0[(C + d), c]     (the xyz rows of the permuted layer-0 image)
__global__ void sa_pack_wx_kernel(const float* __restrict__ wt0, int ldw, int C, int d1, const float* __restrict__ scale,
                                  float* __restrict__ wxs) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 3 * d1; e += gridDim.x * blockDim.x) {
    const int d = e / d1, c = e - d * d1;
    wxs[e] = scale[c] * wt0[(size_t)(C + d) * ldw + c];
  }
}

// conv weight (Cout, Cin) with Cin ordered [xyz(3) | feats(C)]  ->  wt (Cin x ldw) with rows ordered
// [feats(C) | xyz(3)] when permute != 0 (layer 0), else unchanged order.
__global__ void sa_pack_weight_kernel(const float* __restrict__ w, int Cin, int Cout, int ldw, int C_feat, int permute,
                                      float* __restrict__ wt) {
  const int total = Cin * ldw;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int k = e / ldw, c = e - k * ldw;
    int src_k = k;
    if (permute) src_k = k < C_feat ? k + 3 : k - C_feat;
    wt[e] = c < Cout ? w[(size_t)c * Cin + src_k] : 0.f;
  }
}

__global__ void sa_pack_vec_kernel(const float* __restrict__ scale, const float* __restrict__ shift, int Cout, int ldw,
                                   float* __restrict__ dscale, float* __restrict__ dshift) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ldw; c += gridDim.x * blockDim.x) {
    dscale[c] = c < Cout ? (scale ? scale[c] : 1.f) : 0.f;
    dshift[c] = c < Cout ? (shift ? shift[c] : 0.f) : 0.f;
  }
}

// One warp per (centre, sample) pair: X[row] = [ feats[b, idx, 0:C] | rel_xyz(3) | 0 pad ].
__global__ void __launch_bounds__(256) sa_group_rows_kernel(const float* __restrict__ xyz, const float* __restrict__ feats,
                                                             int ldf, const float* __restrict__ new_xyz,
                                                             const int* __restrict__ idx, int N, int M, int ns, int C,
                                                             float radius, int normalize, long long rows,
                                                             float* __restrict__ X, int ldx,
                                                             const float* __restrict__ pair_scalar) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long row = warp; row < rows; row += nwarps) {
    const long long cj = row / ns;           // b * M + j
    const int b = (int)(cj / M);
    const int i = idx ? __ldg(idx + row) : (int)(row % ns);
    const float* src = feats ? feats + ((size_t)b * N + i) * ldf : nullptr;
    float* dst = X + (size_t)row * ldx;
    for (int c = lane; c < C; c += 32) dst[c] = __ldg(src + c);
    if (lane < ldx - C) {
      float v = 0.f;
      if (pair_scalar != nullptr) {
        if (lane == 0) v = __ldg(pair_scalar + row);
      } else if (lane < 3) {
        v = __fsub_rn(__ldg(xyz + ((size_t)b * N + i) * 3 + lane), __ldg(new_xyz + (size_t)cj * 3 + lane));
        if (normalize) v = __fdiv_rn(v, radius);
      }
      dst[C + lane] = v;
    }
  }
}

// out_pm[(b*M+j), c] = max_s H[((b*M+j)*ns + s), c]
__global__ void __launch_bounds__(256) sa_group_max_kernel(const float* __restrict__ H, int ldh, int ns, int Cout,
                                                            long long centres, float* __restrict__ out_pm, int ld_out) {
  const long long total = centres * Cout;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long cj = e / Cout;
    const int c = (int)(e - cj * Cout);
    const float* h = H + (size_t)cj * ns * ldh + c;
    float m = h[0];
    for (int s = 1; s < ns; ++s) m = fmaxf(m, h[(size_t)s * ldh]);
    out_pm[(size_t)cj * ld_out + c] = m;
  }
}

}  // namespace

extern "C" size_t ptt_sa_params_floats(int C, int n_layers, const int* h_dims) {
  SaLayout L;
  return sa_layout(C, n_layers, h_dims, &L) ? L.total : 0;
}

extern "C" int ptt_sa_pack_params(int C, int n_layers, const int* h_dims, const float* const* h_weights,
                                  const float* const* h_scale, const float* const* h_shift, float* params,
                                  ptt_stream_t stream) {
  SaLayout L;
  PTT_CHECK_ARG(sa_layout(C, n_layers, h_dims, &L) && h_weights && params);
  cudaStream_t st = as_stream(stream);
  for (int l = 0; l < n_layers; ++l) {
    PTT_CHECK_ARG(h_weights[l] != nullptr);
    const int total = L.dims[l] * L.ldw[l];
    sa_pack_weight_kernel<<<min(ceil_div(total, 256), 1024), 256, 0, st>>>(h_weights[l], L.dims[l], L.dims[l + 1],
                                                                             L.ldw[l], C, l == 0, params + L.wt[l]); PTT_LAUNCHED();
    sa_pack_vec_kernel<<<ceil_div(L.ldw[l], 256), 256, 0, st>>>(h_scale ? h_scale[l] : nullptr,
                                                                 h_shift ? h_shift[l] : nullptr, L.dims[l + 1], L.ldw[l],
                                                                 params + L.scale[l], params + L.shift[l]); PTT_LAUNCHED();
    int rc = ptt_tc_pack_weight(params + L.wt[l], 1, L.ldw[l], L.dims[l + 1], L.dims[l], params + L.wimg[l], st);
    if (rc != PTT_OK) return rc;
  }
  if (L.fused_dims) {
    int rc;
    if (C > 0 && (rc = ptt_tc_pack_weight(params + L.wt[0], 1, L.ldw[0], L.dims[1], C, params + L.g1img, st))) return rc;
    sa_pack_wx_kernel<<<ceil_div(3 * L.dims[1], 256), 256, 0, st>>>(params + L.wt[0], L.ldw[0], C, L.dims[1],
                                                                    params + L.scale[0], params + L.wxs); PTT_LAUNCHED();
    if ((rc = ptt_tc_pack_weight(params + L.wt[1], 1, L.ldw[1], L.dims[2], L.dims[1], params + L.w2s, st, params + L.scale[1]))) return rc;
    if ((rc = ptt_tc_pack_weight(params + L.wt[2], 1, L.ldw[2], L.dims[3], L.dims[2], params + L.w3s, st, params + L.scale[2]))) return rc;
  }
  return ptt_launch_status();
}

namespace {
struct SaWorkspace {
  bool fused;
  int ldx, ldh;
  size_t x_off, ha_off, hb_off, pm_off, g_off, total;  // in floats
};
bool sa_workspace(int B, int N, int M, int ns, int C, const SaLayout& L, SaWorkspace* W) {
  const size_t rows = (size_t)B * M * ns;
  int cmax = 0;
  for (int l = 1; l <= L.n_layers; ++l) cmax = max(cmax, L.dims[l]);
  W->ldx = round_up(C + 3, 4);
  W->ldh = round_up(cmax, 4);
  W->fused = L.fused_dims && sa_fused_supported(L.dims[1], L.dims[2], L.dims[3], ns);
  size_t off = 0;
  if (W->fused) {
    W->x_off = W->ha_off = W->hb_off = 0;
    W->g_off = off;  off += align_up(C > 0 ? (size_t)B * N * L.dims[1] : 0, 64);
    W->pm_off = off; off += align_up((size_t)B * M * W->ldh, 64);
    W->total = off + 64;
    return true;
  }
  W->g_off = 0;
  W->x_off = off;  off += align_up(rows * W->ldx, 64);
  W->ha_off = off; off += align_up(rows * W->ldh, 64);
  W->hb_off = off; off += align_up(rows * W->ldh, 64);
  W->pm_off = off; off += align_up((size_t)B * M * W->ldh, 64);
  W->total = off;
  return true;
}
}  // namespace

extern "C" size_t ptt_sa_mlp_workspace_bytes(int B, int N, int M, int ns, int C, int n_layers, const int* h_dims) {
  SaLayout L;
  SaWorkspace W;
  if (B <= 0 || N <= 0 || M <= 0 || ns <= 0 || !sa_layout(C, n_layers, h_dims, &L)) return 0;
  sa_workspace(B, N, M, ns, C, L, &W);
  return W.total * sizeof(float);
}

// pair_scalar != nullptr: the relative part of pair (j, s) is (pair_scalar[b, j, s], 0, 0) and xyz / new_xyz are unused;
// idx == nullptr (then ns == N): every centre's group is all N points in order.
static int sa_mlp_fwd_impl(const float* xyz, const float* feats, int ldf, const float* new_xyz, const int* idx,
                           const float* pair_scalar, int B, int N, int M, int ns, int C, float radius, int normalize_xyz,
                           int n_layers, const int* h_dims, const float* params, float* out_pm, int ld_out, float* out_cm,
                           void* workspace, size_t workspace_bytes, ptt_stream_t stream) {
  SaLayout L;
  PTT_CHECK_ARG(B >= 0 && N >= 1 && M >= 0 && ns >= 1 && sa_layout(C, n_layers, h_dims, &L));
  if (B == 0 || M == 0) return PTT_OK;
  PTT_CHECK_ARG(params && (out_pm || out_cm));
  PTT_CHECK_ARG(pair_scalar != nullptr || (xyz && new_xyz));
  PTT_CHECK_ARG(idx != nullptr || ns == N);
  PTT_CHECK_ARG(C == 0 || (feats != nullptr && ldf >= C));
  const int Cout = L.dims[n_layers];
  PTT_CHECK_ARG(out_pm == nullptr || ld_out >= Cout);
  SaWorkspace W;
  sa_workspace(B, N, M, ns, C, L, &W);
  if (workspace == nullptr || workspace_bytes < W.total * sizeof(float)) return PTT_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 15u) != 0) return PTT_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  float* ws = static_cast<float*>(workspace);

  if (W.fused && (out_pm == nullptr || (ld_out % 4 == 0 && (reinterpret_cast<uintptr_t>(out_pm) & 15u) == 0))) {
    // ---- fused path: per-point layer-1 contraction, then ONE kernel for gather + layers 1-3 + max
    int rc;
    float* gprime = nullptr;
    if (C > 0) {
      gprime = ws + W.g_off;
      PttGemmArgs g;
      g.x = feats; g.ldx = ldf; g.R = B * N; g.K = C;
      g.wt = params + L.wt[0]; g.ldw = L.ldw[0]; g.N = L.dims[1];
      g.wimg = params + L.g1img;
      g.scale = params + L.scale[0]; g.shift = params + L.shift[0]; g.relu = 0;
      g.y = gprime; g.ldy = L.dims[1];
      if ((rc = ptt_gemm_launch(g, st))) return rc;
    }
    float* pm = out_pm ? out_pm : ws + W.pm_off;
    const int ld_pm = out_pm ? ld_out : W.ldh;
    SaFusedArgs f;
    f.xyz = xyz; f.new_xyz = new_xyz; f.idx = idx; f.pair_scalar = pair_scalar;
    f.gprime = gprime; f.shift1 = params + L.shift[0]; f.wx = params + L.wxs;
    f.w2img = params + L.w2s; f.w3img = params + L.w3s;
    f.shift2 = params + L.shift[1]; f.shift3 = params + L.shift[2];
    f.B = B; f.N = N; f.M = M; f.ns = ns; f.radius = radius; f.normalize = normalize_xyz;
    f.out_pm = pm; f.ld_out = ld_pm; f.rows = (long long)B * M * ns;
    if ((rc = sa_fused_launch(f, L.dims[1], L.dims[2], L.dims[3], st))) return rc;
    if (out_cm) rc = ptt_pm_to_cm(pm, ld_pm, B, L.dims[3], M, out_cm, stream);
    return rc;
  }
  if (W.fused) return PTT_ERR_INVALID_ARGUMENT;   // out_pm must be 16-byte aligned with ld_out % 4 == 0

  float* X = ws + W.x_off;
  float* H[2] = {ws + W.ha_off, ws + W.hb_off};
  const long long rows = (long long)B * M * ns;

  {
    const long long blocks = llmin_((rows + 7) / 8, 148LL * 16);
    sa_group_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(xyz, feats, ldf, new_xyz, idx, N, M, ns, C, radius,
                                                           normalize_xyz, rows, X, W.ldx, pair_scalar); PTT_LAUNCHED();
  }
  const float* in = X;
  int ld_in = W.ldx;
  for (int l = 0; l < n_layers; ++l) {
    PttGemmArgs a;
    a.x = in; a.ldx = ld_in; a.R = (int)rows; a.K = L.dims[l];
    a.wt = params + L.wt[l]; a.ldw = L.ldw[l]; a.N = L.dims[l + 1];
    a.wimg = params + L.wimg[l];
    a.scale = params + L.scale[l]; a.shift = params + L.shift[l]; a.relu = 1;
    a.y = H[l & 1]; a.ldy = W.ldh;
    int rc = ptt_gemm_launch(a, st);
    if (rc != PTT_OK) return rc;
    in = a.y;
    ld_in = W.ldh;
  }
  float* pm = out_pm ? out_pm : ws + W.pm_off;
  const int ld_pm = out_pm ? ld_out : W.ldh;
  {
    const long long total = (long long)B * M * Cout;
    const long long blocks = llmin_((total + 255) / 256, 148LL * 16);
    sa_group_max_kernel<<<(unsigned)blocks, 256, 0, st>>>(in, ld_in, ns, Cout, (long long)B * M, pm, ld_pm); PTT_LAUNCHED();
  }
  int rc = ptt_launch_status();
  if (rc != PTT_OK) return rc;
  if (out_cm) rc = ptt_pm_to_cm(pm, ld_pm, B, Cout, M, out_cm, stream);
  return rc;
}

extern "C" int ptt_sa_mlp_fwd(const float* xyz, const float* feats, int ldf, const float* new_xyz, const int* idx,
                              int B, int N, int M, int ns, int C, float radius, int normalize_xyz, int n_layers,
                              const int* h_dims, const float* params, float* out_pm, int ld_out, float* out_cm,
                              void* workspace, size_t workspace_bytes, ptt_stream_t stream) {
  PTT_CHECK_ARG(B == 0 || M == 0 || (xyz && new_xyz && idx));
  return sa_mlp_fwd_impl(xyz, feats, ldf, new_xyz, idx, nullptr, B, N, M, ns, C, radius, normalize_xyz, n_layers, h_dims,
                         params, out_pm, ld_out, out_cm, workspace, workspace_bytes, stream);
}

// ------------------------------------------------------------------------------------------------------------------
// CosineSimAug fusion (similarity_modules/p2b_xcoor.py:25-43): for every (search point s, template point t) the row
//   [cos(template_feats_t, search_feats_s) | template_xyz_t | template_feats_t]      (1 + 3 + f channels)
// goes through the SharedMLP and the result is max-pooled over the n1 templates.  Only ONE of the 1 + 3 + f input
// channels depends on s, so this is a set-abstraction layer whose "points" are the templates (features [xyz_t | feats_t],
// C = 3 + f), whose every group is all n1 templates, and whose relative part is (sim, 0, 0): the same fused kernel.  The
// packed parameters are those of ptt_sa_pack_params for C = 3 + f with the layer-0 weight columns ordered
// [w_sim, 0, 0 | W0[:, 1:]] (the binding builds it).
// ------------------------------------------------------------------------------------------------------------------
namespace {

// inv[r] = 1 / max(|row r|, eps): one warp per feature row (F.cosine_similarity's eps = 1e-8)
__global__ void __launch_bounds__(256) row_inv_norm_kernel(const float* __restrict__ X, int ld, int f, long long rows,
                                                            float* __restrict__ inv) {
  const int lane = threadIdx.x & 31;
  const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = warp; r < rows; r += nwarps) {
    const float* x = X + r * ld;
    float ss = 0.f;
    for (int c = lane; c < f; c += 32) { const float v = __ldg(x + c); ss = fmaf(v, v, ss); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) inv[r] = 1.f / fmaxf(sqrtf(ss), 1e-8f);
  }
}

// sim[b, s, t] = (S[b, s, :] . T[b, t, :]) * inv_s[b, s] * inv_t[b, t]: a 32 (search) x 64 (template) output tile per CTA,
// 16-deep feature slabs staged in shared memory, 2 x 4 outputs per thread (exact fp32, CUDA cores: 2 MFLOP per frame)
constexpr int CS_TS = 32, CS_TT = 64, CS_K = 16;
__global__ void __launch_bounds__(256) cosine_sim_kernel(const float* __restrict__ S, int lds, const float* __restrict__ T,
                                                          int ldt, const float* __restrict__ inv_s,
                                                          const float* __restrict__ inv_t, int n1, int n2, int f,
                                                          float* __restrict__ sim) {
  __shared__ float As[CS_K][CS_TS + 1];
  __shared__ float Bs[CS_K][CS_TT + 1];
  const int b = blockIdx.z, s0 = blockIdx.x * CS_TS, t0 = blockIdx.y * CS_TT;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;      // thread -> rows ty*2.., columns tx*4..
  const float* Sb = S + (size_t)b * n2 * lds;
  const float* Tb = T + (size_t)b * n1 * ldt;
  float acc[2][4] = {};
  for (int k0 = 0; k0 < f; k0 += CS_K) {
    for (int e = tid; e < CS_TS * CS_K; e += 256) {
      const int r = e / CS_K, k = e % CS_K;
      As[k][r] = (s0 + r < n2 && k0 + k < f) ? __ldg(Sb + (size_t)(s0 + r) * lds + k0 + k) : 0.f;
    }
    for (int e = tid; e < CS_TT * CS_K; e += 256) {
      const int r = e / CS_K, k = e % CS_K;
      Bs[k][r] = (t0 + r < n1 && k0 + k < f) ? __ldg(Tb + (size_t)(t0 + r) * ldt + k0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < CS_K; ++k) {
      const float a0 = As[k][ty * 2], a1 = As[k][ty * 2 + 1];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float bv = Bs[k][tx * 4 + j];
        acc[0][j] = fmaf(a0, bv, acc[0][j]);
        acc[1][j] = fmaf(a1, bv, acc[1][j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int sr = s0 + ty * 2 + i;
    if (sr >= n2) continue;
    const float is = __ldg(inv_s + (size_t)b * n2 + sr);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int tc = t0 + tx * 4 + j;
      if (tc < n1) sim[((size_t)b * n2 + sr) * n1 + tc] = (acc[i][j] * is) * __ldg(inv_t + (size_t)b * n1 + tc);
    }
  }
}

// X[b * n1 + t] = [template_xyz_t (3) | template_feats_t (f) | 0 pad]
__global__ void __launch_bounds__(256) cosine_rows_kernel(const float* __restrict__ txyz, const float* __restrict__ T, int ldt,
                                                           int f, long long rows, float* __restrict__ X, int ldx) {
  const long long total = rows * ldx;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / ldx;
    const int c = (int)(e - r * ldx);
    float v = 0.f;
    if (c < 3) v = __ldg(txyz + r * 3 + c);
    else if (c < 3 + f) v = __ldg(T + r * ldt + (c - 3));
    X[e] = v;
  }
}

struct CosWorkspace {
  int ldx;
  size_t sim, x, inv_s, inv_t, sa, total;   // float offsets
};
bool cos_workspace(int B, int n1, int n2, int f, int n_layers, const int* h_dims, CosWorkspace* W) {
  W->ldx = round_up(3 + f, 4);
  size_t off = 0;
  W->sim = off; off += align_up((size_t)B * n2 * n1, 64);
  W->x = off;   off += align_up((size_t)B * n1 * W->ldx, 64);
  W->inv_s = off; off += align_up((size_t)B * n2, 64);
  W->inv_t = off; off += align_up((size_t)B * n1, 64);
  W->sa = off;
  const size_t sa_bytes = ptt_sa_mlp_workspace_bytes(B, n1, n2, n1, 3 + f, n_layers, h_dims);
  if (sa_bytes == 0) return false;
  off += sa_bytes / sizeof(float) + 64;
  W->total = off;
  return true;
}

}  // namespace

extern "C" size_t ptt_cosine_fusion_workspace_bytes(int B, int n1, int n2, int f, int n_layers, const int* h_dims) {
  CosWorkspace W;
  if (B <= 0 || n1 <= 0 || n2 <= 0 || f <= 0 || !cos_workspace(B, n1, n2, f, n_layers, h_dims, &W)) return 0;
  return W.total * sizeof(float);
}

extern "C" int ptt_cosine_fusion_fwd(const float* search_feats, int lds, const float* template_feats, int ldt,
                                     const float* template_xyz, int B, int n1, int n2, int f, int n_layers,
                                     const int* h_dims, const float* params, float* out_pm, int ld_out, void* workspace,
                                     size_t workspace_bytes, ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && n1 >= 1 && n2 >= 0 && f >= 1 && lds >= f && ldt >= f && h_dims != nullptr && h_dims[0] == f + 6);
  if (B == 0 || n2 == 0) return PTT_OK;
  PTT_CHECK_ARG(search_feats && template_feats && template_xyz && params && out_pm);
  CosWorkspace W;
  PTT_CHECK_ARG(cos_workspace(B, n1, n2, f, n_layers, h_dims, &W));
  if (workspace == nullptr || workspace_bytes < W.total * sizeof(float)) return PTT_ERR_WORKSPACE;
  if ((reinterpret_cast<uintptr_t>(workspace) & 15u) != 0) return PTT_ERR_WORKSPACE;
  cudaStream_t st = as_stream(stream);
  float* ws = static_cast<float*>(workspace);
  float* sim = ws + W.sim;
  float* X = ws + W.x;
  {
    const long long rows = (long long)B * n2, trows = (long long)B * n1;
    float* inv_s = ws + W.inv_s;
    float* inv_t = ws + W.inv_t;
    row_inv_norm_kernel<<<(unsigned)llmin_((rows + 7) / 8, 148LL * 8), 256, 0, st>>>(search_feats, lds, f, rows, inv_s); PTT_LAUNCHED();
    row_inv_norm_kernel<<<(unsigned)llmin_((trows + 7) / 8, 148LL * 8), 256, 0, st>>>(template_feats, ldt, f, trows, inv_t); PTT_LAUNCHED();
    cosine_sim_kernel<<<dim3(ceil_div(n2, CS_TS), ceil_div(n1, CS_TT), B), 256, 0, st>>>(search_feats, lds, template_feats, ldt,
                                                                                         inv_s, inv_t, n1, n2, f, sim); PTT_LAUNCHED();
    cosine_rows_kernel<<<(unsigned)llmin_((trows * W.ldx + 255) / 256, 148LL * 8), 256, 0, st>>>(template_xyz, template_feats,
                                                                                                 ldt, f, trows, X, W.ldx); PTT_LAUNCHED();
  }
  return sa_mlp_fwd_impl(nullptr, X, W.ldx, nullptr, nullptr, sim, B, n1, n2, n1, 3 + f, 1.f, 0, n_layers, h_dims, params,
                         out_pm, ld_out, nullptr, ws + W.sa, workspace_bytes - W.sa * sizeof(float), stream);
}
