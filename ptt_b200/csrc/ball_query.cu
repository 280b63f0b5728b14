// Ball query for sm_100a.  Replaces pointnet2_ops `_ext.ball_query` (pointnet2_utils.py:287).
//
// Semantics (upstream): per centre scan the cloud IN INDEX ORDER, keep the first nsample points
// with d2 < radius^2 (fp32, the distance contracted exactly as nvcc contracts upstream's expression: sq3), pad the tail
// with the first hit, leave the row zero when there is none.
//
// Kernel: ONE THREAD per centre, as upstream, but over a SHARED-MEMORY copy of the cloud.  Every lane of a warp reads
// the same point at the same time, so the read is one broadcast LDS.128 (no bank conflicts, one wavefront), the scan
// needs no cross-lane traffic at all, and the hit list is appended in index order by construction.  The rows are
// collected in shared memory (odd stride: lanes at the same fill level hit different banks) and leave the CTA as ONE
// contiguous, coalesced block.  A warp stops as soon as all its centres are full.  Compared with the round-1 kernel
// (a warp per centre, 32 points per step, ballot + popc ranks) this executes half the instructions per (centre, point)
// test and stages the cloud with plain vector copies instead of a divide per element.
//
// ptt_ball_query_nested answers the THREE queries of a backbone branch in one launch: PointnetSAModuleVotes samples
// layers 2-3 with 'sequence' = arange(npoint) (pointnet2_modules.py:70-71), so the centres of level l are the first
// M_l FPS samples and the cloud of level l >= 1 is the first M_{l-1} of them -- all three levels only depend on the
// FPS output, and one grid covers them (the CTAs of a level stage that level's cloud).
// Bytes: B * sum_l (12 N_l + 12 M_l + 4 M_l ns_l) algorithmic; a cloud is re-staged by the ceil(M/128) CTAs that share
// it, out of L2.
#include "common.cuh"

namespace {

constexpr int BQ_THREADS = 128;             // centres per CTA
constexpr int BQ_MAX_LEVELS = 4;
constexpr int BQ_MAX_SMEM = 200 * 1024;

struct BqLevel {
  const float* src;        // (B, >= N, 3) cloud, batch stride src_bs floats
  const float* ctr;        // (B, >= M, 3) centres, batch stride ctr_bs floats
  int* out;                // (B, M, ns) contiguous
  long long src_bs, ctr_bs;
  int N, M, ns;
  float r2;
  int cta_end;             // CTAs [previous cta_end, cta_end) of blockIdx.x belong to this level
};

struct BqArgs {
  BqLevel lv[BQ_MAX_LEVELS];
  int levels;
};

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
  const int b = blockIdx.y, tid = threadIdx.x;
  const int N = L.N, ns = L.ns;
  const float* P = L.src + (size_t)b * L.src_bs;
  float4* s_pts = reinterpret_cast<float4*>(bq_smem);
  const int stride = (ns + 1) | 1;                              // ns slots + one dump slot; odd: conflict-free appends at equal fill levels
  int* s_rows = reinterpret_cast<int*>(bq_smem + (STAGED ? (size_t)N * sizeof(float4) : 0));

  if (STAGED) {
    for (int k = tid; k < N; k += BQ_THREADS) s_pts[k] = make_float4(__ldg(P + 3 * k), __ldg(P + 3 * k + 1), __ldg(P + 3 * k + 2), 0.f);
  }
  const int j0 = ((int)blockIdx.x - cta0) * BQ_THREADS;
  const int ncent = min(BQ_THREADS, L.M - j0);
  const bool active = tid < ncent;
  float nx = 0.f, ny = 0.f, nz = 0.f;
  if (active) {
    const float* q = L.ctr + (size_t)b * L.ctr_bs + (size_t)(j0 + tid) * 3;
    nx = __ldg(q); ny = __ldg(q + 1); nz = __ldg(q + 2);
  }
  if (STAGED) __syncthreads();

  int* row = s_rows + tid * stride;
  const float r2 = L.r2;
  int cnt = active ? 0 : ns;                                    // lanes without a centre count as full
  auto test = [&](int k) {
    float4 p;
    if (STAGED) p = s_pts[k];
    else p = make_float4(__ldg(P + 3 * k), __ldg(P + 3 * k + 1), __ldg(P + 3 * k + 2), 0.f);
    // the distance test does not depend on the fill level: the only loop-carried dependency is the increment of cnt
    // (hits beyond nsample land in the row's dump slot)
    if (sq3(nx - p.x, ny - p.y, nz - p.z) < r2) {
      row[min(cnt, ns)] = k;
      ++cnt;
    }
  };
  int k = 0;
  for (; k + 8 <= N; k += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) test(k + u);
    if (__all_sync(0xffffffffu, cnt >= ns)) { k = N; break; }
  }
  for (; k < N; ++k) test(k);

  cnt = min(cnt, ns);
  if (active) {
    const int first = cnt > 0 ? row[0] : 0;                     // pad with the first hit; an empty row stays zero
    for (int s = cnt; s < ns; ++s) row[s] = first;
  }
  __syncthreads();
  // rows j0 .. j0+ncent-1 are one contiguous block of the output
  int* out = L.out + ((size_t)b * L.M + j0) * ns;
  const int total = ncent * ns;
  for (int e = tid; e < total; e += BQ_THREADS) {
    const int r = e / ns;
    out[e] = s_rows[r * stride + (e - r * ns)];
  }
}

int bq_launch(BqArgs& a, int B, cudaStream_t st) {
  int ctas = 0;
  size_t smem_staged = 0, smem_rows = 0;
  for (int l = 0; l < a.levels; ++l) {
    BqLevel& L = a.lv[l];
    ctas += ceil_div(L.M, BQ_THREADS);
    L.cta_end = ctas;
    const size_t rows = (size_t)BQ_THREADS * ((L.ns + 1) | 1) * sizeof(int);
    smem_rows = rows > smem_rows ? rows : smem_rows;
    const size_t staged = (size_t)L.N * sizeof(float4) + rows;
    smem_staged = staged > smem_staged ? staged : smem_staged;
  }
  if (smem_rows > BQ_MAX_SMEM) return PTT_ERR_UNSUPPORTED;      // nsample beyond ~390: not a PointNet++ configuration
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
