// Weight-gradient contraction for sm_100a (tcgen05 + TMEM): the "TN" GEMM of a linear layer's backward pass,
//
//   dW[m, n] += sum_r  dY[r, m] * X[r, n]          (m < M output channels, n < N input channels, r < R rows)
//
// where the reduction runs over the (huge) row dimension -- R = B * M_centres * nsample pair-rows of a set-abstraction
// layer or B * n * k pair-rows of the transformer block (up to 786 k) -- and the result is small (<= 512 x 512).
// Both operands are row-major activations, i.e. the contraction index r is the SLOW dimension of both: in UMMA terms
// both operands are MN-MAJOR.  The producers therefore stage 64-row blocks exactly as they lie in memory -- one
// 128-byte swizzled shared-memory line = 64 consecutive channels of one row -- and the instruction descriptor's
// a_major / b_major bits tell the tensor core to read them transposed.  No transposition pass, no scattered 2-byte
// stores.  fp32-class accuracy through the same fp16 hi/lo split as tc_gemm.cu (three MMAs per K = 16 step).
//
// Grid: ((M tile, N tile) pairs, row chunks [split-K]).  A CTA reduces ITS rows into a TMEM accumulator
// (128 lanes x NT*64 columns) and adds it to dW with fp32 atomics (the order of those additions is not fixed, like the
// atomicAdd scatter of upstream's group_points_grad; the differences are ~1e-7 relative).
//   warps 0-15 producers: fp32 rows of dY and X -> (optional per-channel affine + ReLU on X) -> fp16 hi/lo -> smem;
//              warps 0-7 afterwards run the epilogue (TMEM -> registers -> atomicAdd)
//   warp  16   TMEM allocation + tcgen05.mma issue
// The X operand can be given as the PRE-activation of the producing layer: x = relu(ka[n] * y[r, n] + kb[n]) is applied
// while the rows are staged (BatchNorm + ReLU of the training path), so normalised activations are never stored.
#include "gemm.cuh"
#include "tc_common.cuh"
#include <cmath>

namespace {

constexpr int WG_PROD = 512;             // 16 producer warps: the staging is load-latency bound, more warps = more loads in flight
constexpr int WG_THREADS = WG_PROD + 32;  // + the MMA warp

struct WgArgs {
  const float* dy; int ldy;              // (R, >= M)
  const float* x; int ldx;               // (R, >= N)
  const float* x_ka; const float* x_kb;  // optional per-channel affine (+ ReLU) applied to x on load; both or neither
  float* dw; int ldw;                    // (M, ldw >= N), accumulated into
  float* dbias;                          // optional (M): += column sums of dY (the bias gradient of the same layer)
  long long R;
  int M, N;
  long long rows_per_cta;                // multiple of the k-block height
  int tiles_m;
  int vec4;                              // dw rows are 16-byte aligned and ldw % 4 == 0: vector reductions
};

// MN-major SWIZZLE_128B descriptor: 64-channel chunks LBO apart, 8-row groups SBO = 1 KB apart
__device__ __forceinline__ uint64_t wg_desc(uint32_t smem_addr, uint32_t chunk_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(chunk_bytes >> 4) << 16;             // LBO
  d |= (uint64_t)(1024 >> 4) << 32;                    // SBO
  d |= (uint64_t)1 << 46;                              // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}

__host__ __device__ constexpr uint32_t wg_idesc(int M, int N) {
  return tc::idesc_f16<false>(M, N) | (1u << 15) | (1u << 16);   // a_major = b_major = MN
}

// N tile = NT * 64 columns (NT = 1..5), M tile = MT * 128 channels (MT = 1, 2: two accumulators share the staged X rows,
// which halves the re-reads of X when M >= 256), KBR = rows per k-block (64, or 32 where two stages of 64 do not fit)
template <int NT, int MT, int KBR>
struct WgCfg {
  static constexpr uint32_t CHUNK = KBR * 128;                           // 64 channels x KBR rows
  static constexpr uint32_t A_HALF = 2 * MT * CHUNK;
  static constexpr uint32_t B_HALF = NT * CHUNK;
  static constexpr uint32_t STAGE = 2 * A_HALF + 2 * B_HALF;
  static constexpr int STAGES = (3 * STAGE + 2048 <= 232448) ? 3 : 2;
  static constexpr uint32_t SMEM = STAGES * STAGE + 1024 + 256;
  static constexpr int ACC_COLS = MT * NT * 64;
  static constexpr int TMEM_COLS = ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512));
  static_assert(ACC_COLS <= 512 && STAGES * STAGE + 2048 <= 232448, "tile does not fit TMEM / shared memory");
};

template <int NT, int MT, int KBR>
__global__ void __launch_bounds__(WG_THREADS, 1) tc_wgrad_kernel(const __grid_constant__ WgArgs a) {
  using Cfg = WgCfg<NT, MT, KBR>;
  constexpr int BN = NT * 64, BM = MT * 128;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ctrl = smem + Cfg::STAGES * Cfg::STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(ctrl);            // [STAGES]
  uint64_t* empty = full + Cfg::STAGES;                           // [STAGES]
  uint64_t* accum_full = empty + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // blockIdx.x = the (M tile, N tile) pair, blockIdx.y = the row chunk: the CTAs that read the SAME rows are neighbours in
  // launch order, run at the same time and share those rows through L2 (with the row chunk fastest, the re-reads of a
  // 512 x 512 layer -- dY twice, X four times -- came from HBM: 1.2 GB instead of 0.4 GB)
  const int tile_m = blockIdx.x % a.tiles_m, tile_n = blockIdx.x / a.tiles_m;
  const long long r_begin = (long long)blockIdx.y * a.rows_per_cta;
  const long long r_end = r_begin + a.rows_per_cta < a.R ? r_begin + a.rows_per_cta : a.R;
  const int m0 = tile_m * BM, n0 = tile_n * BN;
  const int KB = r_end > r_begin ? (int)((r_end - r_begin + KBR - 1) / KBR) : 0;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      tc::mbar_init(&full[s], WG_PROD / 32);          // one arrival per producer warp
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(accum_full, 1);
    tc::mbar_init_fence();
  }
  if (warp == WG_PROD / 32) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < WG_PROD / 32) {
    // ------------------------------------------------------------ producers
    // The k-block is staged in two halves of KBR / 2 rows, software-pipelined: the global loads of the following halves are
    // in flight while half h is converted and stored, so the load latency is paid once per CTA, not once per k-block.
    int stage = 0;
    uint32_t phase = 0;
    const bool xform = a.x_ka != nullptr;
    constexpr int HR = KBR / 2;
    constexpr int AQ = BM / 4, BQ = BN / 4;             // float4 per dY / X row
    constexpr int NA = (HR * AQ + WG_PROD - 1) / WG_PROD, NB = (HR * BQ + WG_PROD - 1) / WG_PROD;
    static_assert(WG_PROD % AQ == 0, "a producer thread keeps its dY columns over all its rows");
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);       // column sums of this thread's 4 dY channels (bias gradient)
    auto load_half = [&](int hb, float4 (&va)[NA], float4 (&vb)[NB]) {
      const long long rb = r_begin + (long long)hb * HR;
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        const int e = tid + WG_PROD * i;
        const int row = e / AQ, c4 = e - row * AQ;
        const long long r = rb + row;
        const int m = m0 + c4 * 4;
        va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < HR * AQ && r < r_end && m < a.M) {
          const float* p = a.dy + r * a.ldy + m;
          if (m + 3 < a.M) va[i] = __ldg(reinterpret_cast<const float4*>(p));
          else { va[i].x = __ldg(p); if (m + 1 < a.M) va[i].y = __ldg(p + 1); if (m + 2 < a.M) va[i].z = __ldg(p + 2); }
        }
      }
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const int e = tid + WG_PROD * i;
        const int row = e / BQ, c4 = e - row * BQ;
        const long long r = rb + row;
        const int n = n0 + c4 * 4;
        vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < HR * BQ && r < r_end && n < a.N) {
          const float* p = a.x + r * a.ldx + n;
          if (n + 3 < a.N) vb[i] = __ldg(reinterpret_cast<const float4*>(p));
          else { vb[i].x = __ldg(p); if (n + 1 < a.N) vb[i].y = __ldg(p + 1); if (n + 2 < a.N) vb[i].z = __ldg(p + 2); }
        }
      }
    };
    auto store_half = [&](int half, float4 (&va)[NA], float4 (&vb)[NB]) {
      uint8_t* a_hi = smem + stage * Cfg::STAGE;
      uint8_t* a_lo = a_hi + Cfg::A_HALF;
      uint8_t* b_hi = a_lo + Cfg::A_HALF;
      uint8_t* b_lo = b_hi + Cfg::B_HALF;
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        const int e = tid + WG_PROD * i;
        if (e < HR * AQ) {
          const int row = half * HR + e / AQ, c4 = e % AQ;
          bsum.x += va[i].x; bsum.y += va[i].y; bsum.z += va[i].z; bsum.w += va[i].w;
          uint2 ph, pl;
          tc::split_f16x2(va[i].x, va[i].y, ph.x, pl.x);
          tc::split_f16x2(va[i].z, va[i].w, ph.y, pl.y);
          const uint32_t off = (uint32_t)(c4 >> 4) * Cfg::CHUNK + tc::sw128_offset(row, (c4 & 15) >> 1) + ((c4 & 1) << 3);
          *reinterpret_cast<uint2*>(a_hi + off) = ph;
          *reinterpret_cast<uint2*>(a_lo + off) = pl;
        }
      }
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const int e = tid + WG_PROD * i;
        if (e < HR * BQ) {
          const int row = half * HR + e / BQ, c4 = e % BQ;
          const int n = n0 + c4 * 4;
          if (xform) {
            float* f = reinterpret_cast<float*>(&vb[i]);
#pragma unroll
            for (int u = 0; u < 4; ++u) f[u] = n + u < a.N ? fmaxf(fmaf(f[u], __ldg(a.x_ka + n + u), __ldg(a.x_kb + n + u)), 0.f) : 0.f;
          }
          uint2 ph, pl;
          tc::split_f16x2(vb[i].x, vb[i].y, ph.x, pl.x);
          tc::split_f16x2(vb[i].z, vb[i].w, ph.y, pl.y);
          const uint32_t off = (uint32_t)(c4 >> 4) * Cfg::CHUNK + tc::sw128_offset(row, (c4 & 15) >> 1) + ((c4 & 1) << 3);
          *reinterpret_cast<uint2*>(b_hi + off) = ph;
          *reinterpret_cast<uint2*>(b_lo + off) = pl;
        }
      }
    };
    constexpr int KD = (NA + NB) <= 3 ? 2 : 1;            // k-blocks held in registers (see below)
    // KD k-blocks (2 * KD halves) live in registers: narrow tiles move few bytes per half, so they keep two k-blocks of
    // loads in flight to cover the HBM latency; wide tiles already have > 64 KB outstanding with one.  (An L2 bulk
    // prefetch several k-blocks ahead was measured too: no gain, the register-level loads already cover the latency.)
    float4 va[KD][2][NA], vb[KD][2][NB];
#pragma unroll
    for (int d = 0; d < KD; ++d)
      if (d < KB) {
        load_half(2 * d, va[d][0], vb[d][0]);
        load_half(2 * d + 1, va[d][1], vb[d][1]);          // rows beyond r_end load as zeros
      }
    for (int kb0 = 0; kb0 < KB; kb0 += KD) {
#pragma unroll
      for (int d = 0; d < KD; ++d) {
        const int kb = kb0 + d;
        if (kb < KB) {
          const bool more = kb + KD < KB;
          tc::mbar_wait(&empty[stage], phase ^ 1);
          store_half(0, va[d][0], vb[d][0]);
          if (more) load_half(2 * (kb + KD), va[d][0], vb[d][0]);
          store_half(1, va[d][1], vb[d][1]);
          if (more) load_half(2 * (kb + KD) + 1, va[d][1], vb[d][1]);
          tc::fence_proxy_async_smem();                     // every writer publishes its stores to the async proxy ...
          tc::mbar_arrive_warp(&full[stage]);               // ... and one lane per warp signals: 16 arrivals per k-block, not 512
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
    if (a.dbias != nullptr && tile_n == 0) {          // rows beyond r_end and channels beyond M were staged as zeros
      const int m = m0 + (tid % AQ) * 4;
      const float bs[4] = {bsum.x, bsum.y, bsum.z, bsum.w};
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (m + u < a.M) atomicAdd(a.dbias + m + u, bs[u]);
    }
    // ------------------------------------------------------------ epilogue: warp w -> lane quarter w % 4, column half w / 4
    if (KB > 0 && warp < 8) {
      tc::mbar_wait(accum_full, 0);
      tc::tc_fence_after();
      const int quarter = warp & 3, chalf = warp >> 2;
      constexpr int CH = Cfg::ACC_COLS / 2;            // accumulator columns per warp (both M halves laid side by side)
#pragma unroll 1
      for (int c0 = chalf * CH; c0 < chalf * CH + CH; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
        const int mt = c0 / BN, nc = c0 - mt * BN;     // which 128-channel half, column inside the N tile
        const int m = m0 + mt * 128 + quarter * 32 + lane;
        if (m < a.M) {
          // 16-byte vector reductions (red.global.add.v4.f32): a quarter of the L2 atomic operations.  ldw % 4 == 0, so a
          // group that straddles N only adds the zeros staged for the padding columns.
          float* row = a.dw + (size_t)m * a.ldw + n0 + nc;
          if (a.vec4) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (n0 + nc + j < a.N) atomicAdd(reinterpret_cast<float4*>(row + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + nc + j < a.N) atomicAdd(row + j, v[j]);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    constexpr int N1 = BN > 256 ? 256 : BN, N2 = BN - N1;
    constexpr uint32_t IDESC1 = wg_idesc(128, N1);
    constexpr uint32_t IDESC2 = wg_idesc(128, N2 > 0 ? N2 : 16);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < KB; ++kb) {
      tc::mbar_wait(&full[stage], phase);
      tc::tc_fence_after();
      const uint32_t base = tc::smem_u32(smem + stage * Cfg::STAGE);
      const uint64_t db_hi = wg_desc(base + 2 * Cfg::A_HALF, Cfg::CHUNK), db_lo = wg_desc(base + 2 * Cfg::A_HALF + Cfg::B_HALF, Cfg::CHUNK);
#pragma unroll
      for (int k = 0; k < KBR / 16; ++k) {
        const uint64_t adv = (uint64_t)(k * (2048 >> 4));       // 16 rows = two 8-row groups = 2 KB
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const uint64_t da_hi = wg_desc(base + mt * 2 * Cfg::CHUNK, Cfg::CHUNK);
          const uint64_t da_lo = wg_desc(base + Cfg::A_HALF + mt * 2 * Cfg::CHUNK, Cfg::CHUNK);
          const uint32_t d = tmem_base + (uint32_t)(mt * BN);
          tc::mma_f16_w(d, da_hi + adv, db_hi + adv, IDESC1, (kb | k) != 0);
          tc::mma_f16_w(d, da_lo + adv, db_hi + adv, IDESC1, 1);
          tc::mma_f16_w(d, da_hi + adv, db_lo + adv, IDESC1, 1);
          if (N2 > 0) {                                         // columns 256 .. BN-1: the chunks after the fourth (MT == 1 only)
            const uint64_t nb = (uint64_t)((4 * Cfg::CHUNK) >> 4);
            tc::mma_f16_w(d + 256, da_hi + adv, db_hi + adv + nb, IDESC2, (kb | k) != 0);
            tc::mma_f16_w(d + 256, da_lo + adv, db_hi + adv + nb, IDESC2, 1);
            tc::mma_f16_w(d + 256, da_hi + adv, db_lo + adv + nb, IDESC2, 1);
          }
        }
      }
      tc::mma_commit_w(&empty[stage]);
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
    if (KB > 0) tc::mma_commit_w(accum_full);
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == WG_PROD / 32) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int NT, int MT, int KBR>
int wg_launch(WgArgs a, int ntiles_n, cudaStream_t st) {
  using Cfg = WgCfg<NT, MT, KBR>;
  auto kern = tc_wgrad_kernel<NT, MT, KBR>;
  static bool configured[PTT_MAX_DEVICES] = {};
  const int dev = ptt_current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return (int)e;
    if (int rc = tc::tc_bind_fault(ptt_fault_word())) return rc;
    configured[dev] = true;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles_m = ceil_div(a.M, MT * 128);
  const int tiles_mn = tiles_m * ntiles_n;
  const long long kblocks = (a.R + KBR - 1) / KBR;
  int nsplit = (int)llmin_(kblocks, (long long)((2 * sms + tiles_mn - 1) / tiles_mn));   // ~2 waves of CTAs
  // ... unless the rows are few: every CTA adds its whole tile to dW with atomics, so the split costs nsplit x M x N reductions
  // (~1.6e11 floats/s with 16-byte vector reductions) against (k-blocks / nsplit) x ~1.5 us of staging per CTA; the sum is
  // smallest at nsplit^2 = k-blocks x 2.4e5 / (tiles x tile size).  (A 6144-row, 512 x 256 gradient: 48 -> 13 CTAs per tile.)
  {
    const double tile_elems = (double)tiles_mn * (MT * 128) * (NT * 64);
    const int best = (int)(sqrt((double)kblocks * 2.4e5 / tile_elems) + 0.5);
    if (best < nsplit) nsplit = best;
  }
  if (nsplit < 1) nsplit = 1;
  a.rows_per_cta = ((kblocks + nsplit - 1) / nsplit) * KBR;
  nsplit = (int)((a.R + a.rows_per_cta - 1) / a.rows_per_cta);
  a.tiles_m = tiles_m;
  dim3 grid(tiles_mn, nsplit);
  kern<<<grid, WG_THREADS, Cfg::SMEM, st>>>(a); PTT_LAUNCHED();
  return ptt_launch_status();
}

}  // namespace

// dW (M, ldw)[:, 0:N] += dY (R, ldy)[:, 0:M]^T . f(X (R, ldx)[:, 0:N]);  f = identity, or relu(ka * x + kb) per column;
// dbias (M, optional) += column sums of dY
int ptt_tc_wgrad_launch(const float* dy, int ldy, const float* x, int ldx, const float* x_ka, const float* x_kb,
                        long long R, int M, int N, float* dw, int ldw, float* dbias, cudaStream_t st) {
  if (R <= 0 || M <= 0 || N <= 0) return PTT_OK;
  if ((ldy % 4) || (ldx % 4) || (reinterpret_cast<uintptr_t>(dy) & 15u) || (reinterpret_cast<uintptr_t>(x) & 15u))
    return PTT_ERR_UNSUPPORTED;                       // rows are read as float4
  WgArgs a;
  a.dy = dy; a.ldy = ldy; a.x = x; a.ldx = ldx; a.x_ka = x_ka; a.x_kb = x_kb;
  a.vec4 = (ldw % 4 == 0) && (reinterpret_cast<uintptr_t>(dw) & 15u) == 0;
  a.dw = dw; a.ldw = ldw; a.dbias = dbias; a.R = R; a.M = M; a.N = N; a.rows_per_cta = 0;
  // Tile shape by HBM traffic per row (the kernel is bandwidth-bound): a CTA reads its rows of dY once per N tile and its
  // rows of X once per M tile.  One accumulator (M tile 128, N tile <= 320) or two (M tile 256, N tile <= 256).
  const int chunks = ceil_div(N, 64);
  const int nt1_tiles = ceil_div(chunks, 5), nt2_tiles = ceil_div(chunks, 4);
  const long long traffic1 = (long long)M * nt1_tiles + (long long)N * ceil_div(M, 128);
  const long long traffic2 = (long long)M * nt2_tiles + (long long)N * ceil_div(M, 256);
  if (M > 128 && traffic2 < traffic1) {
    const int nt = ceil_div(chunks, nt2_tiles);
    switch (nt) {
      case 1: return wg_launch<1, 2, 64>(a, nt2_tiles, st);
      case 2: return wg_launch<2, 2, 64>(a, nt2_tiles, st);
      default: break;                                   // wider N tiles: two stages of two accumulators do not fit shared memory
    }
  }
  const int nt = ceil_div(chunks, nt1_tiles);
  switch (nt) {
    case 1: return wg_launch<1, 1, 64>(a, nt1_tiles, st);
    case 2: return wg_launch<2, 1, 64>(a, nt1_tiles, st);
    case 3: return wg_launch<3, 1, 64>(a, nt1_tiles, st);
    case 4: return wg_launch<4, 1, 64>(a, nt1_tiles, st);
    default: return wg_launch<5, 1, 64>(a, nt1_tiles, st);
  }
}
