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
// Grid: (row chunks [split-K], 128-channel tiles of M, tiles of N).  A CTA reduces ITS rows into a TMEM accumulator
// (128 lanes x NT*64 columns) and adds it to dW with fp32 atomics (the order of those additions is not fixed, like the
// atomicAdd scatter of upstream's group_points_grad; the differences are ~1e-7 relative).
//   warps 0-7  producers: fp32 rows of dY and X -> (optional per-channel affine + ReLU on X) -> fp16 hi/lo -> smem;
//              afterwards the epilogue (TMEM -> registers -> atomicAdd)
//   warp  8    TMEM allocation + tcgen05.mma issue
// The X operand can be given as the PRE-activation of the producing layer: x = relu(ka[n] * y[r, n] + kb[n]) is applied
// while the rows are staged (BatchNorm + ReLU of the training path), so normalised activations are never stored.
#include "gemm.cuh"
#include "tc_common.cuh"

namespace {

constexpr int WG_THREADS = 288;          // 8 producer warps + MMA warp
constexpr int WG_KB = 64;                // rows per k-block
constexpr uint32_t WG_CHUNK = 8192;      // one 64-channel chunk of a k-block: 64 rows x 128 bytes

struct WgArgs {
  const float* dy; int ldy;              // (R, >= M)
  const float* x; int ldx;               // (R, >= N)
  const float* x_ka; const float* x_kb;  // optional per-channel affine (+ ReLU) applied to x on load; both or neither
  float* dw; int ldw;                    // (M, ldw >= N), accumulated into
  long long R;
  int M, N;
  long long rows_per_cta;                // multiple of 64
};

// MN-major SWIZZLE_128B descriptor: 64-channel chunks LBO = 8 KB apart, 8-row groups SBO = 1 KB apart
__device__ __forceinline__ uint64_t wg_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(WG_CHUNK >> 4) << 16;                // LBO
  d |= (uint64_t)(1024 >> 4) << 32;                    // SBO
  d |= (uint64_t)1 << 46;                              // descriptor version 1 (Blackwell)
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}

__host__ __device__ constexpr uint32_t wg_idesc(int M, int N) {
  return tc::idesc_f16<false>(M, N) | (1u << 15) | (1u << 16);   // a_major = b_major = MN
}

template <int NT>                        // N tile = NT * 64 columns (NT = 1..5)
struct WgCfg {
  static constexpr uint32_t A_HALF = 2 * WG_CHUNK;                       // 128 channels
  static constexpr uint32_t B_HALF = NT * WG_CHUNK;
  static constexpr uint32_t STAGE = 2 * A_HALF + 2 * B_HALF;
  static constexpr int STAGES = 2;
  static constexpr uint32_t SMEM = STAGES * STAGE + 1024 + 256;
  static constexpr int TMEM_COLS = NT <= 1 ? 64 : (NT <= 2 ? 128 : (NT <= 4 ? 256 : 512));
};

template <int NT>
__global__ void __launch_bounds__(WG_THREADS, 1) tc_wgrad_kernel(const __grid_constant__ WgArgs a) {
  using Cfg = WgCfg<NT>;
  constexpr int BN = NT * 64;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* ctrl = smem + Cfg::STAGES * Cfg::STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(ctrl);            // [STAGES]
  uint64_t* empty = full + Cfg::STAGES;                           // [STAGES]
  uint64_t* accum_full = empty + Cfg::STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long r_begin = (long long)blockIdx.x * a.rows_per_cta;
  const long long r_end = r_begin + a.rows_per_cta < a.R ? r_begin + a.rows_per_cta : a.R;
  const int m0 = blockIdx.y * 128, n0 = blockIdx.z * BN;
  const int KB = r_end > r_begin ? (int)((r_end - r_begin + WG_KB - 1) / WG_KB) : 0;

  if (tid == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      tc::mbar_init(&full[s], 256);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(accum_full, 1);
    tc::mbar_init_fence();
  }
  if (warp == 8) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------ producers
    int stage = 0;
    uint32_t phase = 0;
    const bool xform = a.x_ka != nullptr;
    for (int kb = 0; kb < KB; ++kb) {
      const long long rb = r_begin + (long long)kb * WG_KB;
      // issue this k-block's global loads before waiting for the stage: dY 64 x 128, X 64 x BN (float4 granules)
      float4 va[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = (tid >> 5) + 8 * i, c4 = tid & 31;
        const long long r = rb + row;
        const int m = m0 + c4 * 4;
        va[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < r_end && m < a.M) {
          const float* p = a.dy + r * a.ldy + m;
          if (m + 3 < a.M) va[i] = __ldg(reinterpret_cast<const float4*>(p));
          else { va[i].x = __ldg(p); if (m + 1 < a.M) va[i].y = __ldg(p + 1); if (m + 2 < a.M) va[i].z = __ldg(p + 2); }
        }
      }
      constexpr int BQ = BN / 4;                       // float4 per X row
      constexpr int NB = (64 * BQ + 255) / 256;        // float4 per thread
      float4 vb[NB];
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const int e = tid + 256 * i;
        const int row = e / BQ, c4 = e - row * BQ;
        const long long r = rb + row;
        const int n = n0 + c4 * 4;
        vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < 64 * BQ && r < r_end && n < a.N) {
          const float* p = a.x + r * a.ldx + n;
          if (n + 3 < a.N) vb[i] = __ldg(reinterpret_cast<const float4*>(p));
          else { vb[i].x = __ldg(p); if (n + 1 < a.N) vb[i].y = __ldg(p + 1); if (n + 2 < a.N) vb[i].z = __ldg(p + 2); }
          if (xform) {
            float* f = reinterpret_cast<float*>(&vb[i]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (n + u < a.N) f[u] = fmaxf(fmaf(f[u], __ldg(a.x_ka + n + u), __ldg(a.x_kb + n + u)), 0.f);
          }
        }
      }
      tc::mbar_wait(&empty[stage], phase ^ 1);
      uint8_t* a_hi = smem + stage * Cfg::STAGE;
      uint8_t* a_lo = a_hi + Cfg::A_HALF;
      uint8_t* b_hi = a_lo + Cfg::A_HALF;
      uint8_t* b_lo = b_hi + Cfg::B_HALF;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = (tid >> 5) + 8 * i, c4 = tid & 31;
        uint2 ph, pl;
        tc::split_f16x2(va[i].x, va[i].y, ph.x, pl.x);
        tc::split_f16x2(va[i].z, va[i].w, ph.y, pl.y);
        const uint32_t off = (uint32_t)(c4 >> 4) * WG_CHUNK + tc::sw128_offset(row, (c4 & 15) >> 1) + ((c4 & 1) << 3);
        *reinterpret_cast<uint2*>(a_hi + off) = ph;
        *reinterpret_cast<uint2*>(a_lo + off) = pl;
      }
#pragma unroll
      for (int i = 0; i < NB; ++i) {
        const int e = tid + 256 * i;
        if (e < 64 * BQ) {
          const int row = e / BQ, c4 = e - row * BQ;
          uint2 ph, pl;
          tc::split_f16x2(vb[i].x, vb[i].y, ph.x, pl.x);
          tc::split_f16x2(vb[i].z, vb[i].w, ph.y, pl.y);
          const uint32_t off = (uint32_t)(c4 >> 4) * WG_CHUNK + tc::sw128_offset(row, (c4 & 15) >> 1) + ((c4 & 1) << 3);
          *reinterpret_cast<uint2*>(b_hi + off) = ph;
          *reinterpret_cast<uint2*>(b_lo + off) = pl;
        }
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&full[stage]);
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
    // ------------------------------------------------------------ epilogue: warp w -> lane quarter w % 4, column half w / 4
    if (KB > 0) {
      tc::mbar_wait(accum_full, 0);
      tc::tc_fence_after();
      const int quarter = warp & 3, chalf = warp >> 2;
      const int m = m0 + quarter * 32 + lane;
      constexpr int CH = BN / 2;                       // NT*32 columns per warp
#pragma unroll 1
      for (int c0 = chalf * CH; c0 < chalf * CH + CH; c0 += 32) {
        float v[32];
        tc::tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
        if (m < a.M) {
          float* row = a.dw + (size_t)m * a.ldw + n0 + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + j < a.N) atomicAdd(row + j, v[j]);
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
      const uint64_t da_hi = wg_desc(base), da_lo = wg_desc(base + Cfg::A_HALF);
      const uint64_t db_hi = wg_desc(base + 2 * Cfg::A_HALF), db_lo = wg_desc(base + 2 * Cfg::A_HALF + Cfg::B_HALF);
#pragma unroll
      for (int k = 0; k < WG_KB / 16; ++k) {
        const uint64_t adv = (uint64_t)(k * (2048 >> 4));       // 16 rows = two 8-row groups = 2 KB
        tc::mma_f16_w(tmem_base, da_hi + adv, db_hi + adv, IDESC1, (kb | k) != 0);
        tc::mma_f16_w(tmem_base, da_lo + adv, db_hi + adv, IDESC1, 1);
        tc::mma_f16_w(tmem_base, da_hi + adv, db_lo + adv, IDESC1, 1);
        if (N2 > 0) {                                           // columns 256 .. BN-1: the chunks after the fourth
          const uint64_t nb = (uint64_t)((4 * WG_CHUNK) >> 4);
          tc::mma_f16_w(tmem_base + 256, da_hi + adv, db_hi + adv + nb, IDESC2, (kb | k) != 0);
          tc::mma_f16_w(tmem_base + 256, da_lo + adv, db_hi + adv + nb, IDESC2, 1);
          tc::mma_f16_w(tmem_base + 256, da_hi + adv, db_lo + adv + nb, IDESC2, 1);
        }
      }
      tc::mma_commit_w(&empty[stage]);
      if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
    }
    if (KB > 0) tc::mma_commit_w(accum_full);
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int NT>
int wg_launch(const WgArgs& a, int nsplit, int ntiles_n, cudaStream_t st) {
  using Cfg = WgCfg<NT>;
  auto kern = tc_wgrad_kernel<NT>;
  static bool configured[PTT_MAX_DEVICES] = {};
  const int dev = ptt_current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM);
    if (e != cudaSuccess) return (int)e;
    if (int rc = tc::tc_bind_fault(ptt_fault_word())) return rc;
    configured[dev] = true;
  }
  dim3 grid(nsplit, ceil_div(a.M, 128), ntiles_n);
  kern<<<grid, WG_THREADS, Cfg::SMEM, st>>>(a); PTT_LAUNCHED();
  return ptt_launch_status();
}

}  // namespace

// dW (M, ldw)[:, 0:N] += dY (R, ldy)[:, 0:M]^T . f(X (R, ldx)[:, 0:N]);  f = identity, or relu(ka * x + kb) per column
int ptt_tc_wgrad_launch(const float* dy, int ldy, const float* x, int ldx, const float* x_ka, const float* x_kb,
                        long long R, int M, int N, float* dw, int ldw, cudaStream_t st) {
  if (R <= 0 || M <= 0 || N <= 0) return PTT_OK;
  if ((ldy % 4) || (ldx % 4) || (reinterpret_cast<uintptr_t>(dy) & 15u) || (reinterpret_cast<uintptr_t>(x) & 15u))
    return PTT_ERR_UNSUPPORTED;                       // rows are read as float4
  WgArgs a;
  a.dy = dy; a.ldy = ldy; a.x = x; a.ldx = ldx; a.x_ka = x_ka; a.x_kb = x_kb;
  a.dw = dw; a.ldw = ldw; a.R = R; a.M = M; a.N = N;
  // N tiles of NT * 64 columns, NT <= 5 (320 columns: the 259 / 260-wide first layers fit one tile)
  const int chunks = ceil_div(N, 64);
  const int ntiles = ceil_div(chunks, 5);
  const int nt = ceil_div(chunks, ntiles);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ptt_current_device());
  const int tiles_mn = ceil_div(M, 128) * ntiles;
  long long kblocks = (R + WG_KB - 1) / WG_KB;
  int nsplit = (int)llmin_(kblocks, (long long)((2 * sms + tiles_mn - 1) / tiles_mn));   // ~2 waves of CTAs
  if (nsplit < 1) nsplit = 1;
  a.rows_per_cta = ((kblocks + nsplit - 1) / nsplit) * WG_KB;
  nsplit = (int)((R + a.rows_per_cta - 1) / a.rows_per_cta);
  switch (nt) {
    case 1: return wg_launch<1>(a, nsplit, ntiles, st);
    case 2: return wg_launch<2>(a, nsplit, ntiles, st);
    case 3: return wg_launch<3>(a, nsplit, ntiles, st);
    case 4: return wg_launch<4>(a, nsplit, ntiles, st);
    default: return wg_launch<5>(a, nsplit, ntiles, st);
  }
}
