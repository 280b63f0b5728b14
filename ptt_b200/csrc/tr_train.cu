// Backward-pass element kernels of the kNN vector-attention block (transformer_block/variants.py:149-165), the pieces
// between the tensor-core contractions of ptt_b200/train_ops.py::_TransformerTrain:
//
//   forward (fused, tr_fused.cu)      res_i = sum_j p_ij (v_j + pos_ij),  p = softmax_j(fc_gamma(q_i - k_j + pos_ij) / sqrt(d))
//   tr_softmax_bwd                    d(v+pos)_ij = p_ij dres_i ;  dlogit_ij = p_ij (dres_i (v+pos)_ij - sum_j' p_ij' dres_i (v+pos)_ij') / sqrt(d)
//   tr_pair_inputs                    recomputes what the fused forward never stores: h1 = relu(fc_delta.0(xyz_i - xyz_j)),
//                                     the 3-vector itself, and the attention input a = q_i - k_j + pos_ij
//   tr_mask_positive                  dy <- dy * [ref > 0]   (ReLU backward against a stored activation)
//   tr_pair_scatter                   dpos = da + d(v+pos);  dq_i = sum_j da_ij;  dk_j -= da_ij;  dv_j += d(v+pos)_ij  (atomics,
//                                     like upstream's group_points_grad)
#include "common.cuh"
#include "gemm.cuh"
#include "tr_fused.cuh"

namespace {

constexpr int TT = 256;
inline unsigned tt_grid(long long total) { return (unsigned)llmin_((total + TT - 1) / TT, 148LL * 16); }

__global__ void __launch_bounds__(TT) tr_softmax_bwd_kernel(const float* __restrict__ dres, int ldr, const float* __restrict__ attn,
                                                             const float* __restrict__ vp, int ld, long long tokens, int k, int dm,
                                                             float inv_div, float* __restrict__ dlogit, float* __restrict__ dvp) {
  const long long total = tokens * dm;
  for (long long e = (long long)blockIdx.x * TT + threadIdx.x; e < total; e += (long long)gridDim.x * TT) {
    const long long tok = e / dm;
    const int c = (int)(e - tok * dm);
    const float dr = __ldg(dres + tok * ldr + c);
    float dot = 0.f;
    for (int j = 0; j < k; ++j) {
      const long long pr = tok * k + j;
      dot = fmaf(__ldg(attn + pr * dm + c), dr * __ldg(vp + pr * ld + c), dot);
    }
    for (int j = 0; j < k; ++j) {
      const long long pr = tok * k + j;
      const float p = __ldg(attn + pr * dm + c);
      dvp[pr * ld + c] = p * dr;
      dlogit[pr * ld + c] = p * (dr * __ldg(vp + pr * ld + c) - dot) * inv_div;
    }
  }
}

// h1 (pairs, ld) = relu(Wd0 . delta + bd0), delta rows (pairs, 4) = (xyz_i - xyz_j, 0), a_in (pairs, ld) = q_i - k_j + vp_ij - v_j
__global__ void __launch_bounds__(TT) tr_pair_inputs_kernel(const float* __restrict__ xyz, const int* __restrict__ knn, int n, int k,
                                                             int dm, const float* __restrict__ wd0 /* (dm,3) */,
                                                             const float* __restrict__ bd0, const float* __restrict__ q,
                                                             const float* __restrict__ kk, const float* __restrict__ v, int ldt,
                                                             const float* __restrict__ vp, int ld, long long pairs,
                                                             float* __restrict__ h1, float* __restrict__ delta,
                                                             float* __restrict__ a_in) {
  const long long total = pairs * dm;
  for (long long e = (long long)blockIdx.x * TT + threadIdx.x; e < total; e += (long long)gridDim.x * TT) {
    const long long pr = e / dm;
    const int c = (int)(e - pr * dm);
    const long long tok = pr / k;
    const long long b = tok / n;
    const long long nb = b * n + __ldg(knn + pr);
    const float dx = __ldg(xyz + tok * 3) - __ldg(xyz + nb * 3), dy = __ldg(xyz + tok * 3 + 1) - __ldg(xyz + nb * 3 + 1),
                dz = __ldg(xyz + tok * 3 + 2) - __ldg(xyz + nb * 3 + 2);
    float hv = bd0 ? __ldg(bd0 + c) : 0.f;
    hv = fmaf(dx, __ldg(wd0 + c * 3), hv);
    hv = fmaf(dy, __ldg(wd0 + c * 3 + 1), hv);
    hv = fmaf(dz, __ldg(wd0 + c * 3 + 2), hv);
    h1[pr * ld + c] = fmaxf(hv, 0.f);
    if (c < 4) delta[pr * 4 + c] = c == 0 ? dx : (c == 1 ? dy : (c == 2 ? dz : 0.f));
    a_in[pr * ld + c] = (__ldg(q + tok * ldt + c) - __ldg(kk + nb * ldt + c)) + (__ldg(vp + pr * ld + c) - __ldg(v + nb * ldt + c));
  }
}

// the same, 4 channels per thread (dm, ld, ldt multiples of 4; 16-byte aligned bases): 16-byte accesses, 32-bit index math
__global__ void __launch_bounds__(TT) tr_pair_inputs_v4_kernel(const float* __restrict__ xyz, const int* __restrict__ knn, int n, int k,
                                                                int dm, const float* __restrict__ wd0, const float* __restrict__ bd0,
                                                                const float* __restrict__ q, const float* __restrict__ kk,
                                                                const float* __restrict__ v, int ldt, const float* __restrict__ vp,
                                                                int ld, unsigned pairs, float* __restrict__ h1,
                                                                float* __restrict__ delta, float* __restrict__ a_in) {
  const unsigned q4 = (unsigned)dm / 4;
  const unsigned total = pairs * q4;
  for (unsigned e = blockIdx.x * TT + threadIdx.x; e < total; e += gridDim.x * TT) {
    const unsigned pr = e / q4;
    const int c = (int)(e - pr * q4) * 4;
    const unsigned tok = pr / (unsigned)k;
    const unsigned nb = (tok / (unsigned)n) * (unsigned)n + (unsigned)__ldg(knn + pr);
    const float dx = __ldg(xyz + (size_t)tok * 3) - __ldg(xyz + (size_t)nb * 3), dy = __ldg(xyz + (size_t)tok * 3 + 1) - __ldg(xyz + (size_t)nb * 3 + 1),
                dz = __ldg(xyz + (size_t)tok * 3 + 2) - __ldg(xyz + (size_t)nb * 3 + 2);
    float hv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float t = bd0 ? __ldg(bd0 + c + u) : 0.f;
      t = fmaf(dx, __ldg(wd0 + (c + u) * 3), t);
      t = fmaf(dy, __ldg(wd0 + (c + u) * 3 + 1), t);
      t = fmaf(dz, __ldg(wd0 + (c + u) * 3 + 2), t);
      hv[u] = fmaxf(t, 0.f);
    }
    *reinterpret_cast<float4*>(h1 + (size_t)pr * ld + c) = make_float4(hv[0], hv[1], hv[2], hv[3]);
    if (c == 0) *reinterpret_cast<float4*>(delta + (size_t)pr * 4) = make_float4(dx, dy, dz, 0.f);
    const float4 qv = __ldg(reinterpret_cast<const float4*>(q + (size_t)tok * ldt + c));
    const float4 kv = __ldg(reinterpret_cast<const float4*>(kk + (size_t)nb * ldt + c));
    const float4 pv = __ldg(reinterpret_cast<const float4*>(vp + (size_t)pr * ld + c));
    const float4 vv = __ldg(reinterpret_cast<const float4*>(v + (size_t)nb * ldt + c));
    *reinterpret_cast<float4*>(a_in + (size_t)pr * ld + c) =
        make_float4((qv.x - kv.x) + (pv.x - vv.x), (qv.y - kv.y) + (pv.y - vv.y), (qv.z - kv.z) + (pv.z - vv.z), (qv.w - kv.w) + (pv.w - vv.w));
  }
}

__global__ void __launch_bounds__(TT) tr_mask_positive_kernel(float* __restrict__ dy, const float* __restrict__ ref, long long total4) {
  for (long long e = (long long)blockIdx.x * TT + threadIdx.x; e < total4; e += (long long)gridDim.x * TT) {
    float4 d = reinterpret_cast<float4*>(dy)[e];
    const float4 r = __ldg(reinterpret_cast<const float4*>(ref) + e);
    d.x = r.x > 0.f ? d.x : 0.f; d.y = r.y > 0.f ? d.y : 0.f; d.z = r.z > 0.f ? d.z : 0.f; d.w = r.w > 0.f ? d.w : 0.f;
    reinterpret_cast<float4*>(dy)[e] = d;
  }
}

// one thread per (token, channel): da (pairs, ld) becomes dpos = da + dvp in place; dq (tokens, ldt) written; dk / dv (tokens,
// ldt) accumulated with atomics (zero them first)
__global__ void __launch_bounds__(TT) tr_pair_scatter_kernel(float* __restrict__ da, const float* __restrict__ dvp, int ld,
                                                              const int* __restrict__ knn, int n, int k, int dm, long long tokens,
                                                              float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv,
                                                              int ldt) {
  const long long total = tokens * dm;
  for (long long e = (long long)blockIdx.x * TT + threadIdx.x; e < total; e += (long long)gridDim.x * TT) {
    const long long tok = e / dm;
    const int c = (int)(e - tok * dm);
    const long long b = tok / n;
    float sq = 0.f;
    for (int j = 0; j < k; ++j) {
      const long long pr = tok * k + j;
      const long long nb = b * n + __ldg(knn + pr);
      const float a = da[pr * ld + c], p = __ldg(dvp + pr * ld + c);
      sq += a;
      atomicAdd(dk + nb * ldt + c, -a);
      atomicAdd(dv + nb * ldt + c, p);
      da[pr * ld + c] = a + p;
    }
    dq[tok * ldt + c] = sq;
  }
}

// 4 channels per thread; the scatter uses 16-byte vector reductions (red.global.add.v4.f32)
__global__ void __launch_bounds__(TT) tr_pair_scatter_v4_kernel(float* __restrict__ da, const float* __restrict__ dvp, int ld,
                                                                 const int* __restrict__ knn, int n, int k, int dm, unsigned tokens,
                                                                 float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv,
                                                                 int ldt) {
  const unsigned q4 = (unsigned)dm / 4;
  const unsigned total = tokens * q4;
  for (unsigned e = blockIdx.x * TT + threadIdx.x; e < total; e += gridDim.x * TT) {
    const unsigned tok = e / q4;
    const int c = (int)(e - tok * q4) * 4;
    const unsigned b0 = (tok / (unsigned)n) * (unsigned)n;
    float4 sq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int j = 0; j < k; ++j) {
      const size_t pr = (size_t)tok * k + j;
      const size_t nb = b0 + (unsigned)__ldg(knn + pr);
      float4* ap = reinterpret_cast<float4*>(da + pr * ld + c);
      const float4 a = *ap, p = __ldg(reinterpret_cast<const float4*>(dvp + pr * ld + c));
      sq.x += a.x; sq.y += a.y; sq.z += a.z; sq.w += a.w;
      atomicAdd(reinterpret_cast<float4*>(dk + nb * ldt + c), make_float4(-a.x, -a.y, -a.z, -a.w));
      atomicAdd(reinterpret_cast<float4*>(dv + nb * ldt + c), p);
      *ap = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    }
    *reinterpret_cast<float4*>(dq + (size_t)tok * ldt + c) = sq;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" int ptt_tr_softmax_bwd(const float* dres, int ldr, const float* attn, const float* vp, int ld, long long tokens, int k,
                                  int dm, float divisor, float* dlogit, float* dvp, ptt_stream_t stream) {
  PTT_CHECK_ARG(tokens >= 0 && k >= 1 && dm >= 1 && ld >= dm && ldr >= dm && divisor > 0.f);
  if (tokens == 0) return PTT_OK;
  PTT_CHECK_ARG(dres && attn && vp && dlogit && dvp);
  tr_softmax_bwd_kernel<<<tt_grid(tokens * dm), TT, 0, as_stream(stream)>>>(dres, ldr, attn, vp, ld, tokens, k, dm, 1.f / divisor,
                                                                           dlogit, dvp); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_tr_pair_inputs(const float* xyz, const int* knn, int B, int n, int k, int dm, const float* delta0_w,
                                  const float* delta0_b, const float* q, const float* kk, const float* v, int ldt,
                                  const float* vp, int ld, float* h1, float* delta, float* a_in, ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && n >= 1 && k >= 1 && dm >= 4 && ld >= dm && ldt >= dm);
  if (B == 0) return PTT_OK;
  PTT_CHECK_ARG(xyz && knn && delta0_w && q && kk && v && vp && h1 && delta && a_in);
  const long long pairs = (long long)B * n * k;
  if (dm % 4 == 0 && ld % 4 == 0 && ldt % 4 == 0 && pairs * (dm / 4) < 0x7fffffffLL && aligned16(q) && aligned16(kk) && aligned16(v) &&
      aligned16(vp) && aligned16(h1) && aligned16(delta) && aligned16(a_in)) {
    tr_pair_inputs_v4_kernel<<<tt_grid(pairs * (dm / 4)), TT, 0, as_stream(stream)>>>(xyz, knn, n, k, dm, delta0_w, delta0_b, q, kk, v,
                                                                                     ldt, vp, ld, (unsigned)pairs, h1, delta, a_in);
    PTT_LAUNCHED();
    return ptt_launch_status();
  }
  tr_pair_inputs_kernel<<<tt_grid(pairs * dm), TT, 0, as_stream(stream)>>>(xyz, knn, n, k, dm, delta0_w, delta0_b, q, kk, v, ldt, vp,
                                                                          ld, pairs, h1, delta, a_in); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_tr_mask_positive(float* dy, const float* ref, long long count, ptt_stream_t stream) {
  PTT_CHECK_ARG(count >= 0 && count % 4 == 0);
  if (count == 0) return PTT_OK;
  PTT_CHECK_ARG(dy && ref);
  tr_mask_positive_kernel<<<tt_grid(count / 4), TT, 0, as_stream(stream)>>>(dy, ref, count / 4); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_tr_pair_scatter(float* da, const float* dvp, int ld, const int* knn, int B, int n, int k, int dm, float* dq,
                                   float* dk, float* dv, int ldt, ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && n >= 1 && k >= 1 && dm >= 1 && ld >= dm && ldt >= dm);
  if (B == 0) return PTT_OK;
  PTT_CHECK_ARG(da && dvp && knn && dq && dk && dv);
  const long long tokens = (long long)B * n;
  if (dm % 4 == 0 && ld % 4 == 0 && ldt % 4 == 0 && tokens * k * (dm / 4) < 0x7fffffffLL && aligned16(da) && aligned16(dvp) && aligned16(dq) &&
      aligned16(dk) && aligned16(dv)) {
    tr_pair_scatter_v4_kernel<<<tt_grid(tokens * (dm / 4)), TT, 0, as_stream(stream)>>>(da, dvp, ld, knn, n, k, dm, (unsigned)tokens, dq,
                                                                                       dk, dv, ldt);
    PTT_LAUNCHED();
    return ptt_launch_status();
  }
  tr_pair_scatter_kernel<<<tt_grid(tokens * dm), TT, 0, as_stream(stream)>>>(da, dvp, ld, knn, n, k, dm, tokens, dq, dk, dv, ldt);
  PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_tr_rows_linear(const float* x, int ldx, long long R, int d, const float* params, const float* mask_ref,
                                  float* out, ptt_stream_t stream) {
  PTT_CHECK_ARG(R >= 0 && d >= 1 && ldx >= d);
  if (R == 0) return PTT_OK;
  PTT_CHECK_ARG(x && params && out);
  if ((d != 256 && d != 512) || (ldx % 4) || (reinterpret_cast<uintptr_t>(x) & 15u)) return PTT_ERR_UNSUPPORTED;
  TrPassArgs t;
  t.n = 1; t.k = 1; t.dm = d; t.pairs = R;
  t.a_src = x; t.lda = ldx;
  t.wimg = params + (size_t)(d + 1) * ptt_linear_ldw(d);      // the fp16 hi/lo image inside the packed nn.Linear block
  t.bias = nullptr;
  t.out = out; t.ldo = d;
  t.mask_ref = mask_ref;
  return tr_fused_launch(t, TR_PROD_PLAIN, TR_EPI_STORE, as_stream(stream));
}
