// Weight-stationary, persistent row-block contraction with fused column statistics (sm_100a, tcgen05 + TMEM):
//
//   y[r, 0:N] = f(x[r, 0:K]) . W[0:N, 0:K]^T,   f = identity or relu(ka[k] * x + kb[k])     (r < R up to ~10^6 rows)
//   sums[0][c] += sum_r y[r, c],  sums[1][c] += sum_r y[r, c]^2                           (optional, double)
//
// This is the shape of every contraction of the TRAINING path of the set-abstraction layers (Conv2d 1x1 without bias over
// B * M * nsample pair-rows, pytorch_utils.py:25-36; and its input gradient with the transposed weight): K, N <= 260 / 256
// but hundreds of thousands of rows, i.e. HBM-bound.  tc_gemm.cu is built for the opposite regime (a few waves of CTAs):
// it re-loads the weight image and pays its prologue once per 128 rows -- 6144 times at R = 786 k.  Here
//   * one persistent CTA per SM keeps the WHOLE weight image (fp16 hi / lo, <= 160 KB) in shared memory,
//   * the producers stream 128-row tiles of x through a ring (next k-block prefetched in registers, across tiles),
//   * the accumulators are issued TRANSPOSED (D^T = W . X^T: output channels in the TMEM lanes) and double-buffered, so
//     the epilogue of tile t overlaps the MMAs of tile t + 1; a lane is ONE output channel and its registers walk the
//     rows: every store is a coalesced 128-byte row segment and the BatchNorm statistics of the training path are two
//     FMAs per element in the lane's own registers -- the separate column-reduction pass over y disappears.
// fp32-class accuracy through the fp16 hi / lo split (three MMAs per K = 16 step), as everywhere else.
//   warps 0-7   epilogue (warp w: TMEM lane quarter w % 4 = 32 channels; rows [64 * (w / 4), +64) of the tile)
//   warps 8-15  producers (fp32 rows -> optional affine + ReLU -> fp16 hi / lo -> UMMA K-major SW128 tile)
//   warp  16    loads the weight image once (TMA engine bulk copies)
//   warp  17    TMEM allocation + tcgen05.mma issue
#include "gemm.cuh"
#include "tc_common.cuh"

namespace {

constexpr int WS_THREADS = 576;
constexpr int WS_TM = 128;                      // rows per tile
constexpr uint32_t WS_KBLOCK = 32768;           // A k-block: 128 rows x 64 k, hi (16 KB) + lo (16 KB)
constexpr uint32_t WS_WBLOCK = 32768;           // weight block: 128 channels x 64 k, hi + lo
constexpr uint32_t WS_SMEM_MAX = 232448 - 2048;

struct WsArgs {
  const float* x; int ldx;
  const float* ka; const float* kb;             // optional operand transform (both or neither)
  const uint8_t* wimg;                          // ptt_tc_pack_weight image: blocks (nb, kb, hi|lo) of 8 KB
  int n_wblocks;                                // 64-channel blocks present in the image
  float* y; int ldy;
  double* sums;                                 // (2, N) or nullptr
  long long R;
  int K, N, KB, num_tiles;
};

template <int NCH>                               // 128-channel accumulator halves: N <= NCH * 128
__global__ void __launch_bounds__(WS_THREADS, 1) ws_gemm_kernel(const __grid_constant__ WsArgs a, const int nsa) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
  const int KB = a.KB;
  uint8_t* w_smem = smem;                                          // [KB][NCH] blocks of 32 KB
  uint8_t* a_ring = w_smem + (size_t)KB * NCH * WS_WBLOCK;         // [nsa] stages of 32 KB
  uint8_t* ctrl = a_ring + (size_t)nsa * WS_KBLOCK;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(ctrl);            // [4]
  uint64_t* a_empty = a_full + 4;                                  // [4]
  uint64_t* d_full = a_empty + 4;                                  // [2]
  uint64_t* d_free = d_full + 2;                                   // [2]
  uint64_t* w_full = d_free + 2;                                   // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int iters = (a.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
  auto tile_of = [&](int it) { return (long long)blockIdx.x + (long long)it * gridDim.x; };

  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      tc::mbar_init(&a_full[s], 256);
      tc::mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&d_full[s], 1);
      tc::mbar_init(&d_free[s], 256);
    }
    tc::mbar_init(w_full, 1);
    tc::mbar_init_fence();
  }
  // channel blocks the image does not hold (N padded up to NCH * 128) must read as zero
  {
    const int have = a.n_wblocks, want = NCH * 2;
    if (have < want) {
      for (int kb = 0; kb < KB; ++kb)
        for (int j = have; j < want; ++j) {
          // block j of k-block kb: hi rows at [(j / 2) half][j % 2 * 8 KB], lo 16 KB further
          uint8_t* base = w_smem + ((size_t)kb * NCH + (j >> 1)) * WS_WBLOCK + (size_t)(j & 1) * 8192;
          for (int e = tid * 16; e < 8192; e += WS_THREADS * 16) {
            *reinterpret_cast<uint4*>(base + e) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(base + 16384 + e) = make_uint4(0, 0, 0, 0);
          }
        }
      tc::fence_proxy_async_smem();
    }
  }
  if (warp == 17) tc::tmem_alloc(tmem_slot, NCH == 1 ? 256 : 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // =================================================================== epilogue (transposed accumulators)
    const int quarter = warp & 3, rhalf = warp >> 2;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    float s1[NCH], s2[NCH];
#pragma unroll
    for (int h = 0; h < NCH; ++h) s1[h] = s2[h] = 0.f;
    for (int it = 0; it < iters; ++it) {
      const int buf = it & 1;
      const long long r0 = tile_of(it) * WS_TM + rhalf * 64;
      tc::mbar_wait(&d_full[buf], (it >> 1) & 1);
      tc::tc_fence_after();
#pragma unroll
      for (int h = 0; h < NCH; ++h) {
        const int ch = h * 128 + quarter * 32 + lane;
        const bool ch_ok = ch < a.N;
#pragma unroll 1
        for (int b = 0; b < 2; ++b) {
          float v[32];
          tc::tmem_ld32(lane_addr + (uint32_t)(buf * NCH * 128 + h * 128 + rhalf * 64 + b * 32), v);
          if (h == NCH - 1 && b == 1) {                         // accumulator drained
            tc::tc_fence_before();
            tc::mbar_arrive(&d_free[buf]);
          }
          const long long rb = r0 + b * 32;
          if (rb >= a.R || !ch_ok) continue;
          float* op = a.y + rb * a.ldy + ch;
          if (a.R - rb >= 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              op[(long long)j * a.ldy] = v[j];
              s1[h] += v[j];
              s2[h] = fmaf(v[j], v[j], s2[h]);
            }
          } else {                                              // the ragged last block of the problem
            const int nrow = (int)(a.R - rb);
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nrow) {
                op[(long long)j * a.ldy] = v[j];
                s1[h] += v[j];
                s2[h] = fmaf(v[j], v[j], s2[h]);
              }
          }
        }
      }
    }
    if (a.sums != nullptr) {
#pragma unroll
      for (int h = 0; h < NCH; ++h) {
        const int ch = h * 128 + quarter * 32 + lane;
        if (ch < a.N) {
          atomicAdd(a.sums + ch, (double)s1[h]);
          atomicAdd(a.sums + a.N + ch, (double)s2[h]);
        }
      }
    }
  } else if (warp < 16) {
    // =================================================================== producers
    const int pt = tid - 256;
    const int c4 = pt & 15, rsub = pt >> 4;              // float4 column within the 64-wide k-block, row sub-index
    const bool xform = a.ka != nullptr;
    int stage = 0;
    uint32_t phase = 0;
    const int total_blocks = iters * KB;
    auto load_block = [&](int gb, float4 (&v)[8]) {        // gb = it * KB + kb (may run one past the end: zeros)
      const int it = gb / KB, kb = gb - it * KB;
      const long long r0 = tile_of(it) * WS_TM;
      const int k = kb * 64 + c4 * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long r = r0 + i * 16 + rsub;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gb < total_blocks && r < a.R && k < a.K) {
          const float* p = a.x + r * a.ldx + k;
          if (k + 3 < a.K) v[i] = __ldg(reinterpret_cast<const float4*>(p));
          else { v[i].x = __ldg(p); if (k + 1 < a.K) v[i].y = __ldg(p + 1); if (k + 2 < a.K) v[i].z = __ldg(p + 2); }
        }
      }
    };
    float4 cur[8], nxt[8];
    load_block(0, cur);
    for (int gb = 0; gb < total_blocks; ++gb) {
      load_block(gb + 1, nxt);
      const int it = gb / KB, kb = gb - it * KB;
      if (xform) {
        const long long r0 = tile_of(it) * WS_TM;
        const int k = kb * 64 + c4 * 4;
        float ka[4], kc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          ka[u] = k + u < a.K ? __ldg(a.ka + k + u) : 0.f;
          kc[u] = k + u < a.K ? __ldg(a.kb + k + u) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = r0 + i * 16 + rsub < a.R;          // rows beyond R stay zero (they would enter the statistics)
          cur[i].x = ok ? fmaxf(fmaf(cur[i].x, ka[0], kc[0]), 0.f) : 0.f;
          cur[i].y = ok ? fmaxf(fmaf(cur[i].y, ka[1], kc[1]), 0.f) : 0.f;
          cur[i].z = ok ? fmaxf(fmaf(cur[i].z, ka[2], kc[2]), 0.f) : 0.f;
          cur[i].w = ok ? fmaxf(fmaf(cur[i].w, ka[3], kc[3]), 0.f) : 0.f;
        }
      }
      tc::mbar_wait(&a_empty[stage], phase ^ 1);
      uint8_t* a_hi = a_ring + (size_t)stage * WS_KBLOCK;
      uint8_t* a_lo = a_hi + 16384;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 16 + rsub;
        uint2 ph, pl;
        tc::split_f16x2(cur[i].x, cur[i].y, ph.x, pl.x);
        tc::split_f16x2(cur[i].z, cur[i].w, ph.y, pl.y);
        const uint32_t off = tc::sw128_offset(r, c4 >> 1) + ((c4 & 1) << 3);
        *reinterpret_cast<uint2*>(a_hi + off) = ph;
        *reinterpret_cast<uint2*>(a_lo + off) = pl;
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(&a_full[stage]);
      if (++stage == nsa) { stage = 0; phase ^= 1; }
#pragma unroll
      for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
    }
  } else if (warp == 16) {
    // =================================================================== weight loader: the whole image, once
    if (lane == 0) {
      const int have = a.n_wblocks < NCH * 2 ? a.n_wblocks : NCH * 2;
      tc::mbar_arrive_expect_tx(w_full, (uint32_t)KB * (uint32_t)have * 16384u);
      for (int kb = 0; kb < KB; ++kb)
        for (int j = 0; j < have; ++j) {
          const uint8_t* src = a.wimg + ((size_t)j * KB + kb) * 16384;                  // (nb = j, kb): hi 8 KB, lo 8 KB
          uint8_t* dst = w_smem + ((size_t)kb * NCH + (j >> 1)) * WS_WBLOCK + (size_t)(j & 1) * 8192;
          tc::bulk_g2s(dst, src, 8192, w_full);
          tc::bulk_g2s(dst + 16384, src + 8192, 8192, w_full);
        }
    }
  } else {
    // =================================================================== MMA issuer (whole warp, one elected lane issues)
    constexpr uint32_t IDESC = tc::idesc_f16<false>(128, WS_TM);     // D^T item: M = 128 channels, N = 128 rows
    const uint32_t a_addr = tc::smem_u32(a_ring), w_addr = tc::smem_u32(w_smem);
    tc::mbar_wait(w_full, 0);
    tc::tc_fence_after();
    int sa = 0;
    uint32_t pa = 0;
    for (int it = 0; it < iters; ++it) {
      const int buf = it & 1;
      tc::mbar_wait(&d_free[buf], (uint32_t)((it >> 1) & 1) ^ 1u);   // the epilogue has drained this buffer (tile it - 2)
      tc::tc_fence_after();
#pragma unroll 1
      for (int kb = 0; kb < KB; ++kb) {
        tc::mbar_wait(&a_full[sa], pa);
        tc::tc_fence_after();
        const uint64_t dx_hi = tc::smem_desc_sw128(a_addr + sa * WS_KBLOCK);
        const uint64_t dx_lo = tc::smem_desc_sw128(a_addr + sa * WS_KBLOCK + 16384);
#pragma unroll
        for (int h = 0; h < NCH; ++h) {
          const uint32_t wb = w_addr + (uint32_t)((kb * NCH + h) * WS_WBLOCK);
          const uint64_t dw_hi = tc::smem_desc_sw128(wb), dw_lo = tc::smem_desc_sw128(wb + 16384);
          const uint32_t d = tmem_base + (uint32_t)(buf * NCH * 128 + h * 128);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t adv = (uint64_t)(ks * 2);
            tc::mma_f16_w(d, dw_hi + adv, dx_hi + adv, IDESC, (kb | ks) != 0);
            tc::mma_f16_w(d, dw_hi + adv, dx_lo + adv, IDESC, 1);
            tc::mma_f16_w(d, dw_lo + adv, dx_hi + adv, IDESC, 1);
          }
        }
        tc::mma_commit_w(&a_empty[sa]);
        if (++sa == nsa) { sa = 0; pa ^= 1; }
      }
      tc::mma_commit_w(&d_full[buf]);
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, NCH == 1 ? 256 : 512);
  }
}

template <int NCH>
int ws_launch(const WsArgs& a, int nsa, size_t smem, cudaStream_t st) {
  auto kern = ws_gemm_kernel<NCH>;
  static bool configured[PTT_MAX_DEVICES] = {};
  const int dev = ptt_current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_MAX);
    if (e != cudaSuccess) return (int)e;
    if (int rc = tc::tc_bind_fault(ptt_fault_word())) return rc;
    configured[dev] = true;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = a.num_tiles < sms ? a.num_tiles : sms;
  kern<<<grid, WS_THREADS, smem, st>>>(a, nsa); PTT_LAUNCHED();
  return ptt_launch_status();
}

}  // namespace

// Does the weight-stationary kernel cover this problem?  (weight image + >= 2 ring stages in shared memory, N <= 256,
// float4-readable rows, enough rows for persistence to matter)
bool ptt_ws_gemm_supported(const float* x, int ldx, long long R, int K, int N) {
  if (R < 4096 || K < 1 || N < 1 || N > 256) return false;
  if ((ldx % 4) != 0 || (reinterpret_cast<uintptr_t>(x) & 15u) != 0 || round_up(K, 4) > ldx) return false;
  const int KB = ceil_div(K, 64), NCH = N > 128 ? 2 : 1;
  return (size_t)KB * NCH * WS_WBLOCK + 2 * (size_t)WS_KBLOCK + 1024 + 256 <= WS_SMEM_MAX;
}

int ptt_ws_gemm_launch(const float* x, int ldx, long long R, int K, const float* ka, const float* kb, const void* wimg, int N,
                       float* y, int ldy, double* sums, cudaStream_t st) {
  if (R <= 0) return PTT_OK;
  WsArgs a;
  a.x = x; a.ldx = ldx; a.ka = ka; a.kb = kb; a.wimg = static_cast<const uint8_t*>(wimg);
  a.n_wblocks = ceil_div(N, 64);
  a.y = y; a.ldy = ldy; a.sums = sums; a.R = R; a.K = K; a.N = N;
  a.KB = ceil_div(K, 64);
  a.num_tiles = (int)((R + WS_TM - 1) / WS_TM);
  const int NCH = N > 128 ? 2 : 1;
  const size_t wbytes = (size_t)a.KB * NCH * WS_WBLOCK;
  int nsa = (int)((WS_SMEM_MAX - wbytes - 1024 - 256) / WS_KBLOCK);
  nsa = nsa > 4 ? 4 : nsa;
  if (nsa < 2) return PTT_ERR_UNSUPPORTED;
  const size_t smem = wbytes + (size_t)nsa * WS_KBLOCK + 1024 + 256;
  if (sums != nullptr) {
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)2 * N * sizeof(double), st);
    if (e != cudaSuccess) return (int)e;
  }
  return NCH == 1 ? ws_launch<1>(a, nsa, smem, st) : ws_launch<2>(a, nsa, smem, st);
}
