// Ball query for sm_100a.  Replaces pointnet2_ops `_ext.ball_query` (pointnet2_utils.py:287).
//
// Semantics (upstream): per centre scan the cloud IN INDEX ORDER, keep the first nsample points
// with d2 < radius^2 (fp32, the distance contracted exactly as nvcc contracts upstream's expression: sq3), pad the tail
// with the first hit, leave the row zero when there is none.
//
// Kernel: a thread scans a SEGMENT of the cloud for one centre, in index order, over a SHARED-MEMORY copy of the cloud.
// The 32 lanes of a warp are 32 different centres scanning the same segment, so every point read is one broadcast
// LDS.128 (no bank conflicts, one wavefront), the scan needs no cross-lane traffic, and a thread's hit list is in index
// order by construction.  S = 1, 2 or 4 warps share a centre group (segment s = [s*len, (s+1)*len)): their lists are
// concatenated in segment order and cut at nsample, which is exactly the serial first-nsample rule; the split only buys
// parallelism (a 1024-point scan is 4 x 256 points) -- one thread per centre, upstream's mapping, leaves 9 warps per SM
// at batch 48 and a 50 us dependent chain.  The points of a step are fetched eight at a time before they are tested (the
// compiler does not hoist loads across the predicated appends by itself).  Rows are assembled in shared memory and
// leave the CTA as ONE contiguous, coalesced block.  A warp stops as soon as all its lists are full.
//
// ptt_ball_query_nested answers the THREE queries of a backbone branch in one launch: PointnetSAModuleVotes samples
// layers 2-3 with 'sequence' = arange(npoint) (pointnet2_modules.py:70-71), so the centres of level l are the first
// M_l FPS samples and the cloud of level l >= 1 is the first M_{l-1} of them -- all three levels only depend on the
// FPS output, and one grid covers them (the CTAs of a level stage that level's cloud).
// Bytes: B * sum_l (12 N_l + 12 M_l + 4 M_l ns_l) algorithmic; a cloud is re-staged by the ceil(M/128) CTAs that share
// it, out of L2.
#include "common.cuh"

namespace {

constexpr int BQ_THREADS = 256;             // 8 warps = 8 / S centre groups of 32
constexpr int BQ_MAX_LEVELS = 4;
constexpr int BQ_MAX_SMEM = 200 * 1024;

struct BqLevel {
  const float* src;        // (B, >= N, 3) cloud, batch stride src_bs floats
  const float* ctr;        // (B, >= M, 3) centres, batch stride ctr_bs floats
  int* out;                // (B, M, ns) contiguous
  long long src_bs, ctr_bs;
  int N, M, ns;
  float r2;
  int S;                   // segments (warps) per centre group: 1, 2 or 4
  int cta_end;             // CTAs [previous cta_end, cta_end) of blockIdx.x belong to this level
};

struct BqArgs {
  BqLevel lv[BQ_MAX_LEVELS];
  int levels;
};

__host__ __device__ inline int bq_stride(int ns) { return (ns + 1) | 1; }   // ns slots + a dump slot; odd: lanes at equal fill levels hit different banks

// STAGED: the level's cloud fits in shared memory (float4 per point); otherwise points come through L1 / L2.
template <bool STAGED>
__global__ void __launch_bounds__(BQ_THREADS) ball_query_kernel(const __grid_constant__ BqArgs a) {
  extern __shared__ __align__(16) unsigned char bq_smem[];
  int l = 0, cta0 = 0;
  while (l + 1 < a.levels && (int)blockIdx.x >= a.lv[l].cta_end) {
    cta0 = a.lv[l].cta_end;
    ++l;
  }
  const BqLevel& L = a.lv[l];
  const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = L.N, ns = L.ns, S = L.S;
  const int cpc = (BQ_THREADS / 32 / S) * 32;                   // centres per CTA
  const float* P = L.src + (size_t)b * L.src_bs;
  const int stride = bq_stride(ns);
  // smem: [points float4[N]] | lists int[256][stride] | final rows int[cpc][ns] | counts int[256]
  float4* s_pts = reinterpret_cast<float4*>(bq_smem);
  int* s_list = reinterpret_cast<int*>(bq_smem + (STAGED ? (size_t)N * sizeof(float4) : 0));
  int* s_final = s_list + BQ_THREADS * stride;
  int* s_cnt = s_final + cpc * ns;

  if (STAGED) {
    for (int k = tid; k < N; k += BQ_THREADS) s_pts[k] = make_float4(__ldg(P + 3 * k), __ldg(P + 3 * k + 1), __ldg(P + 3 * k + 2), 0.f);
  }
  const int group = warp / S, seg = warp - group * S;           // warp-uniform
  const int cl = group * 32 + lane;                             // centre within the CTA
  const int j0 = ((int)blockIdx.x - cta0) * cpc;
  const int ncent = min(cpc, L.M - j0);
  const bool active = cl < ncent;
  float nx = 0.f, ny = 0.f, nz = 0.f;
  if (active) {
    const float* q = L.ctr + (size_t)b * L.ctr_bs + (size_t)(j0 + cl) * 3;
    nx = __ldg(q); ny = __ldg(q + 1); nz = __ldg(q + 2);
  }
  if (STAGED) __syncthreads();

  int* row = s_list + tid * stride;
  const float r2 = L.r2;
  int cnt = active ? 0 : ns;                                    // lanes without a centre count as full
  const int seg_len = ((N + S - 1) / S + 7) & ~7;
  int k = seg * seg_len;
  const int k_end = min(N, k + seg_len);
  auto fetch = [&](int kk) -> float4 {
    if (STAGED) return s_pts[kk];
    return make_float4(__ldg(P + 3 * kk), __ldg(P + 3 * kk + 1), __ldg(P + 3 * kk + 2), 0.f);
  };
  // the distance test does not depend on the fill level: the only loop-carried dependency is the increment of cnt
  // (hits beyond nsample land in the row's dump slot)
  auto test = [&](const float4& p, int kk) {
    if (sq3(nx - p.x, ny - p.y, nz - p.z) < r2) {
      row[min(cnt, ns)] = kk;
      ++cnt;
    }
  };
  for (; k + 8 <= k_end; k += 8) {
    float4 p[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) p[u] = fetch(k + u);
#pragma unroll
    for (int u = 0; u < 8; ++u) test(p[u], k + u);
    if (__all_sync(0xffffffffu, cnt >= ns)) { k = k_end; break; }
  }
  for (; k < k_end; ++k) test(fetch(k), k);
  cnt = min(cnt, ns);
  s_cnt[tid] = cnt;
  __syncthreads();

  // concatenate the S lists of a centre in segment order, cut at ns
  if (active) {
    int off = 0, total = 0;
    for (int s2 = 0; s2 < S; ++s2) {
      const int c2 = s_cnt[(group * S + s2) * 32 + lane];
      off += s2 < seg ? c2 : 0;
      total += c2;
    }
    int* fin = s_final + cl * ns;
    for (int i = 0; i < cnt && off + i < ns; ++i) fin[off + i] = row[i];
    if (seg == 0) {
      // pad with the first hit (the first entry of the first non-empty list); an empty row stays zero
      int first = 0;
      for (int s2 = S - 1; s2 >= 0; --s2) {
        const int t2 = (group * S + s2) * 32 + lane;
        if (s_cnt[t2] > 0) first = s_list[t2 * stride];
      }
      for (int i = min(total, ns); i < ns; ++i) fin[i] = first;
    }
  }
  __syncthreads();
  // rows j0 .. j0+ncent-1 are one contiguous block of the output
  int* out = L.out + ((size_t)b * L.M + j0) * ns;
  const int total = ncent * ns;
  for (int e = tid; e < total; e += BQ_THREADS) out[e] = s_final[e];
}

int bq_launch(BqArgs& a, int B, cudaStream_t st) {
  int ctas = 0;
  size_t smem_staged = 0, smem_rows = 0;
  for (int l = 0; l < a.levels; ++l) {
    BqLevel& L = a.lv[l];
    L.S = L.N >= 512 ? 4 : (L.N >= 128 ? 2 : 1);
    const int cpc = BQ_THREADS / L.S;
    ctas += ceil_div(L.M, cpc);
    L.cta_end = ctas;
    const size_t rows = ((size_t)BQ_THREADS * bq_stride(L.ns) + (size_t)cpc * L.ns + BQ_THREADS) * sizeof(int);
    smem_rows = rows > smem_rows ? rows : smem_rows;
    const size_t staged = (size_t)L.N * sizeof(float4) + rows;
    smem_staged = staged > smem_staged ? staged : smem_staged;
  }
  if (smem_rows > BQ_MAX_SMEM) return PTT_ERR_UNSUPPORTED;      // nsample beyond ~100: not a PointNet++ configuration
  const bool staged = smem_staged <= BQ_MAX_SMEM;
  const size_t smem = staged ? smem_staged : smem_rows;
  auto kern = staged ? ball_query_kernel<true> : ball_query_kernel<false>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<dim3(ctas, B), BQ_THREADS, smem, st>>>(a); PTT_LAUNCHED();
  return ptt_launch_status();
}

}  // namespace

extern "C" int ptt_ball_query(const float* new_xyz, const float* xyz, int B, int N, int M, float radius,
                              int nsample, int* idx, ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && N >= 1 && M >= 0 && nsample >= 0);
  if (B == 0 || M == 0 || nsample == 0) return PTT_OK;
  PTT_CHECK_ARG(new_xyz && xyz && idx);
  BqArgs a = {};
  a.levels = 1;
  BqLevel& L = a.lv[0];
  L.src = xyz; L.src_bs = (long long)N * 3; L.N = N;
  L.ctr = new_xyz; L.ctr_bs = (long long)M * 3; L.M = M;
  L.r2 = radius * radius;  // fp32 product, as upstream
  L.ns = nsample; L.out = idx;
  return bq_launch(a, B, as_stream(stream));
}

extern "C" int ptt_ball_query_nested(const float* xyz, const float* samples, int B, int N, int levels, const int* h_M,
                                     const float* h_radius, const int* h_nsample, int* const* h_idx,
                                     ptt_stream_t stream) {
  PTT_CHECK_ARG(B >= 0 && N >= 1 && levels >= 1 && levels <= BQ_MAX_LEVELS && h_M && h_radius && h_nsample && h_idx);
  for (int l = 0; l < levels; ++l) {
    PTT_CHECK_ARG(h_M[l] >= 1 && h_nsample[l] >= 1 && h_idx[l] != nullptr);
    PTT_CHECK_ARG(l == 0 ? true : h_M[l] <= h_M[l - 1]);         // nested prefixes
  }
  if (B == 0) return PTT_OK;
  PTT_CHECK_ARG(xyz && samples);
  BqArgs a = {};
  a.levels = levels;
  const long long sbs = (long long)h_M[0] * 3;
  for (int l = 0; l < levels; ++l) {
    BqLevel& L = a.lv[l];
    L.src = l == 0 ? xyz : samples;
    L.src_bs = l == 0 ? (long long)N * 3 : sbs;
    L.N = l == 0 ? N : h_M[l - 1];
    L.ctr = samples; L.ctr_bs = sbs; L.M = h_M[l];
    L.r2 = h_radius[l] * h_radius[l];
    L.ns = h_nsample[l]; L.out = h_idx[l];
  }
  return bq_launch(a, B, as_stream(stream));
}
