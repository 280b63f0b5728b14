// Tensor-core row-block contraction for sm_100a (tcgen05 + TMEM), fp32 in / fp32 out at fp32-class accuracy.
//
//   y[r, 0:N] = act( scale[c] * (x[row(r), 0:K] . W[c, 0:K]) + shift[c] ) (+ residual[r, c])
//
// fp32-within-1e-4 on 16-bit tensor-core inputs: every operand is split into two fp16 halves (x = hi + lo,
// ~22 significant bits) and each K-step issues three MMAs into the same TMEM accumulator,
//   hi.hi + lo.hi + hi.lo        (lo.lo ~ 2^-22 relative is dropped),
// i.e. a K' = 3K fp16 contraction with an fp32 accumulator.  Weights are split once at pack time into the
// UMMA canonical K-major SWIZZLE_128B image (so the TMA engine streams them with plain bulk copies);
// activations are split on the fly by the producer warps while they stage the A tile in shared memory.
//
// CTA = 128 output rows x BN output columns; warp roles:
//   warps 0-7  A producers (global fp32 rows, optional row gather -> fp16 hi/lo -> swizzled smem); warps 0-3 then run
//              the epilogue (TMEM -> registers -> scale/shift/ReLU/residual -> global)
//   warp  8    weight loader: cp.async.bulk (TMA engine) of pre-swizzled 8 KB blocks, mbarrier complete_tx
//   warp  9    TMEM allocation + tcgen05.mma / tcgen05.commit issue (convergent warp, one elected lane)
// K is consumed in 64-wide blocks through a 2-4 stage ring of (A hi, A lo, B hi, B lo) tiles.
#include "gemm.cuh"
#include "tc_common.cuh"

namespace {

constexpr int TBM = 128;       // rows per CTA
constexpr int TBK = 64;        // K per stage = one 128-byte swizzle row of fp16
constexpr int TC_THREADS = 320;       // 8 producer warps (the first 4 also run the epilogue) + loader + MMA
constexpr uint32_t A_HALF_BYTES = TBM * TBK * 2;      // 16 KB
constexpr uint32_t W_BLOCK_BYTES = 64 * TBK * 2;      // 8 KB: 64 weight rows x 64 k

template <int BN>
struct TcCfg {
  static constexpr uint32_t B_HALF_BYTES = BN * TBK * 2;
  static constexpr uint32_t STAGE_BYTES = 2 * A_HALF_BYTES + 2 * B_HALF_BYTES;
  static constexpr int STAGES = BN == 256 ? 2 : (BN == 128 ? 3 : 4);
  static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 4096 /*barriers, tables*/;
};

struct TcParams {
  PttGemmArgs g;
  const __half* wimg;   // packed weight image
  int n_wblocks;        // 64-row weight blocks available (Np / 64)
  int k_blocks;         // Kp / 64
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(const __grid_constant__ TcParams p) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
  uint8_t* ctrl = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_a = reinterpret_cast<uint64_t*>(ctrl);           // [STAGES]
  uint64_t* full_b = full_a + STAGES;                              // [STAGES]
  uint64_t* empty = full_b + STAGES;                               // [STAGES]
  uint64_t* accum_full = empty + STAGES;                           // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);
  int* s_row = reinterpret_cast<int*>(ctrl + 256);                 // [128] source row per tile row, -1 = none
  float* s_scale = reinterpret_cast<float*>(ctrl + 1024);          // [BN] epilogue scale of this CTA's columns
  float* s_shift = reinterpret_cast<float*>(ctrl + 2048);          // [BN] epilogue shift

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // blockIdx.x = row tile * n_tiles + column tile: the CTAs that read the same 128 rows of x are neighbours in launch order
  // and share them through L2 (row tile fastest put them a whole wave apart: x came from HBM once per column tile)
  const int n_tiles = (p.g.N + BN - 1) / BN;
  const int ntile = blockIdx.x % n_tiles;
  const int row0 = (blockIdx.x / n_tiles) * TBM;
  const PttGemmArgs& g = p.g;
  const int KB = p.k_blocks;
  const int bz = blockIdx.z;                                       // batch element
  const float* gx = g.x + (long long)bz * g.x_bstride;
  float* gy = g.y + (long long)bz * g.y_bstride;
  const float* gres = g.residual ? g.residual + (long long)bz * g.res_bstride : nullptr;
  const uint8_t* gw = reinterpret_cast<const uint8_t*>(p.wimg) + (size_t)bz * g.wimg_bstride;

  if (tid < TBM) {
    const int r = row0 + tid;
    s_row[tid] = r < g.R ? (g.a_rows ? g.a_rows[r] : r) : -1;
  }
  for (int c = tid; c < BN; c += TC_THREADS) {
    const int col = ntile * BN + c;
    s_scale[c] = (g.scale && col < g.N) ? __ldg(g.scale + col) : 1.f;
    s_shift[c] = (g.shift && col < g.N) ? __ldg(g.shift + col) : 0.f;
  }
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(&full_a[s], 256);
      tc::mbar_init(&full_b[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(accum_full, 1);
    tc::mbar_init_fence();
  }
  if (warp == 9) tc::tmem_alloc(tmem_slot, BN);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------ A producers
    // The 8 row segments of k-block kb+1 are requested before k-block kb is converted, so the global-load latency is
    // exposed once per CTA instead of once per k-block; full k-blocks take a branch-free path.
    const int c4 = tid & 15;          // float4 column within the 64-wide k block
    const int rsub = tid >> 4;        // 0..15
    int stage = 0;
    uint32_t phase = 0;
    auto load_block = [&](int kb, float4 (&v)[8]) {
      const int k = kb * TBK + c4 * 4;
      if ((kb + 1) * TBK <= g.K) {                                   // uniform: the whole k-block lies inside K
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int src = s_row[i * 16 + rsub];
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (src >= 0) v[i] = __ldg(reinterpret_cast<const float4*>(gx + (size_t)src * g.ldx + k));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int src = s_row[i * 16 + rsub];
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (src >= 0 && k < g.K) {
            const float* ptr = gx + (size_t)src * g.ldx + k;
            if (k + 3 < g.K) {
              v[i] = __ldg(reinterpret_cast<const float4*>(ptr));
            } else {
              v[i].x = __ldg(ptr);
              if (k + 1 < g.K) v[i].y = __ldg(ptr + 1);
              if (k + 2 < g.K) v[i].z = __ldg(ptr + 2);
            }
          }
        }
      }
    };
    // optional A transform (training path): x <- relu(ka[k] * x + kb[k]) for the four k columns this thread stages
    const bool xform = g.a_ka != nullptr;
    auto affine_relu = [&](int kb, float4 (&v)[8]) {
      const int k = kb * TBK + c4 * 4;
      float ka[4], kc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ka[u] = k + u < g.K ? __ldg(g.a_ka + k + u) : 0.f;
        kc[u] = k + u < g.K ? __ldg(g.a_kb + k + u) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        v[i].x = fmaxf(fmaf(v[i].x, ka[0], kc[0]), 0.f);
        v[i].y = fmaxf(fmaf(v[i].y, ka[1], kc[1]), 0.f);
        v[i].z = fmaxf(fmaf(v[i].z, ka[2], kc[2]), 0.f);
        v[i].w = fmaxf(fmaf(v[i].w, ka[3], kc[3]), 0.f);
      }
    };
    float4 cur[8], nxt[8];
    load_block(0, cur);
    for (int kb = 0; kb < KB; ++kb) {
      if (kb + 1 < KB) load_block(kb + 1, nxt);
      if (xform) affine_relu(kb, cur);
      tc::mbar_wait(&empty[stage], phase ^ 1);
      uint8_t* a_hi = smem + stage * Cfg::STAGE_BYTES;
      uint8_t* a_lo = a_hi + A_HALF_BYTES;
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
      tc::mbar_arrive(&full_a[stage]);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
#pragma unroll
      for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
    }
  }
  if (warp < 8) {
    // ------------------------------------------------------------ epilogue: warp w drains TMEM lane quarter w % 4, column
    // half w / 4 (all eight producer warps take part)
    tc::mbar_wait(accum_full, 0);
    tc::tc_fence_after();
    const int quarter = warp & 3, chalf = warp >> 2;
    const int r = row0 + quarter * 32 + lane;
    const bool row_ok = r < g.R;
    const bool vec_y = (g.ldy % 4 == 0) && ((reinterpret_cast<uintptr_t>(gy) & 15u) == 0);
    const bool vec_r = gres && (g.ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(gres) & 15u) == 0);
    constexpr int CH = BN >= 64 ? BN / 2 : BN;          // columns per warp
#pragma unroll 1
    for (int c0 = chalf * CH; c0 < chalf * CH + CH; c0 += 32) {
      const int col0 = ntile * BN + c0;
      if (col0 >= g.N) break;       // warp-uniform
      float v[32];
      tc::tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c0, v);
      if (!row_ok) continue;
      float* yrow = gy + (size_t)r * g.ldy + col0;
      const float* rrow = gres ? gres + (size_t)r * g.ldr + col0 : nullptr;
      if (col0 + 32 <= g.N && vec_y && (rrow == nullptr || vec_r)) {
        // whole 32-column block inside N, vector accesses: straight-line
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 sc = *reinterpret_cast<const float4*>(s_scale + c0 + j);
          const float4 sh = *reinterpret_cast<const float4*>(s_shift + c0 + j);
          float4 o = make_float4(fmaf(v[j + 0], sc.x, sh.x), fmaf(v[j + 1], sc.y, sh.y), fmaf(v[j + 2], sc.z, sh.z),
                                 fmaf(v[j + 3], sc.w, sh.w));
          if (g.relu) o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
          if (rrow) {
            const float4 rr = __ldg(reinterpret_cast<const float4*>(rrow + j));
            o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
          }
          *reinterpret_cast<float4*>(yrow + j) = o;
        }
        continue;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float o[4];
        const float4 sc = *reinterpret_cast<const float4*>(s_scale + c0 + j);
        const float4 sh = *reinterpret_cast<const float4*>(s_shift + c0 + j);
        o[0] = fmaf(v[j + 0], sc.x, sh.x);
        o[1] = fmaf(v[j + 1], sc.y, sh.y);
        o[2] = fmaf(v[j + 2], sc.z, sh.z);
        o[3] = fmaf(v[j + 3], sc.w, sh.w);
        if (g.relu) {
#pragma unroll
          for (int u = 0; u < 4; ++u) o[u] = fmaxf(o[u], 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (col0 + j + u < g.N) yrow[j + u] = o[u] + (rrow ? __ldg(rrow + j + u) : 0.f);
      }
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------ weight loader (TMA engine)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      constexpr int SUB = BN / 64;
      int nsub = p.n_wblocks - ntile * SUB;
      nsub = nsub > SUB ? SUB : nsub;
      for (int kb = 0; kb < KB; ++kb) {
        tc::mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* b_hi = smem + stage * Cfg::STAGE_BYTES + 2 * A_HALF_BYTES;
        uint8_t* b_lo = b_hi + Cfg::B_HALF_BYTES;
        tc::mbar_arrive_expect_tx(&full_b[stage], (uint32_t)nsub * 2u * W_BLOCK_BYTES);
        for (int j = 0; j < nsub; ++j) {
          const size_t blk = ((size_t)(ntile * SUB + j) * KB + kb) * 2;
          const uint8_t* src = gw + blk * W_BLOCK_BYTES;
          tc::bulk_g2s(b_hi + j * W_BLOCK_BYTES, src, W_BLOCK_BYTES, &full_b[stage]);
          tc::bulk_g2s(b_lo + j * W_BLOCK_BYTES, src + W_BLOCK_BYTES, W_BLOCK_BYTES, &full_b[stage]);
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    {
      constexpr uint32_t IDESC = tc::idesc_f16<false>(TBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < KB; ++kb) {
        tc::mbar_wait(&full_a[stage], phase);
        tc::mbar_wait(&full_b[stage], phase);
        tc::tc_fence_after();
        const uint32_t a_hi = tc::smem_u32(smem + stage * Cfg::STAGE_BYTES);
        const uint64_t da_hi = tc::smem_desc_sw128(a_hi);
        const uint64_t da_lo = tc::smem_desc_sw128(a_hi + A_HALF_BYTES);
        const uint64_t db_hi = tc::smem_desc_sw128(a_hi + 2 * A_HALF_BYTES);
        const uint64_t db_lo = tc::smem_desc_sw128(a_hi + 2 * A_HALF_BYTES + Cfg::B_HALF_BYTES);
#pragma unroll
        for (int k = 0; k < TBK / 16; ++k) {
          const uint64_t adv = (uint64_t)(k * 2);   // 32 bytes per UMMA_K, in 16-byte units of the address field
          tc::mma_f16_w(tmem_base, da_hi + adv, db_hi + adv, IDESC, (kb | k) != 0);
          tc::mma_f16_w(tmem_base, da_lo + adv, db_hi + adv, IDESC, 1);
          tc::mma_f16_w(tmem_base, da_hi + adv, db_lo + adv, IDESC, 1);
        }
        tc::mma_commit_w(&empty[stage]);     // frees the stage once these MMAs have read it
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      tc::mma_commit_w(accum_full);
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, BN);
  }
}

// nn.Linear-style weight (Cout, K) row-major fp32  ->  fp16 hi/lo image of 8 KB blocks:
//   block (nb, kb, half) at ((nb * KB + kb) * 2 + half) * 8 KB; inside, row r (0..63) / 16-byte chunk c at
//   sw128_offset(r, c); rows >= Cout and k >= K are zero.
//   src(c, k) = w[c * ld_c + k * ld_k]   (so a transposed (K, ldw) image can be the source too)
// wt (optional, batch 1 only): the CUDA-core image of the same layer is written by the same launch -- wt[k * ldw + c] =
// src(c, k) for k < K, the bias row at k = K, zeros in the padding columns (one launch per layer instead of memset + two)
__global__ void tc_pack_weight_kernel(const float* __restrict__ w, long long ld_c, long long ld_k, int Cout, int K, int NB,
                                      int KB, const float* __restrict__ row_scale, __half* __restrict__ img,
                                      long long w_bstride, size_t img_bstride, float* __restrict__ wt, int ldw,
                                      const float* __restrict__ bias) {
  w += (long long)blockIdx.y * w_bstride;                                                  // batch element
  img = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(img) + (size_t)blockIdx.y * img_bstride);
  if (wt != nullptr) {
    const int total_t = (K + 1) * ldw;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total_t; e += gridDim.x * blockDim.x) {
      const int k = e / ldw, c = e - k * ldw;
      wt[e] = c < Cout ? (k < K ? w[(long long)c * ld_c + (long long)k * ld_k] : (bias ? bias[c] : 0.f)) : 0.f;
    }
  }
  const long long total = (long long)NB * KB * 64 * 8;   // (block, row, chunk)
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e & 7);
    const int r = (int)((e >> 3) & 63);
    const long long blk = e >> 9;
    const int kb = (int)(blk % KB), nb = (int)(blk / KB);
    const int n = nb * 64 + r;
    __half hi[8], lo[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = kb * 64 + c * 8 + u;
      float x = (n < Cout && k < K) ? w[(long long)n * ld_c + (long long)k * ld_k] : 0.f;
      if (row_scale != nullptr && n < Cout) x *= row_scale[n];     // fold a per-output-channel scale into the weight
      tc::split_f16(x, hi[u], lo[u]);
    }
    uint8_t* base = reinterpret_cast<uint8_t*>(img) + (size_t)blk * 2 * W_BLOCK_BYTES + tc::sw128_offset(r, c);
    *reinterpret_cast<uint4*>(base) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(base + W_BLOCK_BYTES) = *reinterpret_cast<const uint4*>(lo);
  }
}

// Many layers in one launch: blockIdx.y = layer, its descriptor read from a device table (the training step repacks ~70
// layers per step; one 5 us launch instead of 70)
__global__ void tc_pack_batch_kernel(const PttPackDesc* __restrict__ descs) {
  const PttPackDesc d = descs[blockIdx.y];
  const float* __restrict__ w = d.weight;
  const int K = d.K, Cout = d.Cout, ldw = (Cout + 3) & ~3;       // = ptt_linear_ldw(Cout)
  const int NB = (Cout + 63) / 64, KB = (K + 63) / 64;
  float* wt = d.params;
  __half* img = reinterpret_cast<__half*>(d.params + (size_t)(K + 1) * ldw);
  const int total_t = (K + 1) * ldw;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total_t; e += gridDim.x * blockDim.x) {
    const int k = e / ldw, c = e - k * ldw;
    wt[e] = c < Cout ? (k < K ? w[(long long)c * d.ld_c + (long long)k * d.ld_k] : (d.bias ? d.bias[c] : 0.f)) : 0.f;
  }
  const int total = NB * KB * 512;   // (block, row, chunk)
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int c = e & 7, r = (e >> 3) & 63, blk = e >> 9;
    const int kb = blk % KB, nb = blk / KB;
    const int n = nb * 64 + r;
    __half hi[8], lo[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = kb * 64 + c * 8 + u;
      const float x = (n < Cout && k < K) ? w[(long long)n * d.ld_c + (long long)k * d.ld_k] : 0.f;
      tc::split_f16(x, hi[u], lo[u]);
    }
    uint8_t* base = reinterpret_cast<uint8_t*>(img) + (size_t)blk * 2 * W_BLOCK_BYTES + tc::sw128_offset(r, c);
    *reinterpret_cast<uint4*>(base) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(base + W_BLOCK_BYTES) = *reinterpret_cast<const uint4*>(lo);
  }
}

template <int BN>
int tc_launch(const TcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  auto kern = tc_gemm_kernel<BN>;
  static bool configured[PTT_MAX_DEVICES] = {};   // per device; idempotent, racing threads set the same value
  const int dev = ptt_current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    if (int rc = tc::tc_bind_fault(ptt_fault_word())) return rc;
    configured[dev] = true;
  }
  dim3 grid(ceil_div(p.g.R, TBM) * ceil_div(p.g.N, BN), 1, p.g.batch > 0 ? p.g.batch : 1);
  kern<<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(p); PTT_LAUNCHED();
  return ptt_launch_status();
}

}  // namespace

size_t ptt_tc_weight_halves(int K, int Cout) {
  return (size_t)ceil_div(Cout, 64) * ceil_div(K, 64) * 2 * (W_BLOCK_BYTES / 2);
}

int ptt_tc_pack_weight(const float* w, long long ld_c, long long ld_k, int Cout, int K, void* img, cudaStream_t st,
                       const float* row_scale, int batch, long long w_bstride, size_t img_bstride) {
  const int NB = ceil_div(Cout, 64), KB = ceil_div(K, 64);
  const long long total = (long long)NB * KB * 512;
  dim3 grid((unsigned)llmin_((total + 255) / 256, 2048), batch > 0 ? batch : 1);
  tc_pack_weight_kernel<<<grid, 256, 0, st>>>(w, ld_c, ld_k, Cout, K, NB, KB, row_scale, static_cast<__half*>(img), w_bstride,
                                              img_bstride, nullptr, 0, nullptr); PTT_LAUNCHED();
  return ptt_launch_status();
}

int ptt_linear_pack_batch_launch(const PttPackDesc* descs_device, int count, cudaStream_t st) {
  if (count <= 0) return PTT_OK;
  tc_pack_batch_kernel<<<dim3(32, (unsigned)count), 256, 0, st>>>(descs_device); PTT_LAUNCHED();
  return ptt_launch_status();
}

// the whole packed nn.Linear image (transposed fp32 weight + bias row + fp16 hi/lo image) in one launch
int ptt_linear_pack_all(const float* w, long long ld_c, long long ld_k, const float* bias, int K, int Cout, float* params,
                        cudaStream_t st) {
  const int NB = ceil_div(Cout, 64), KB = ceil_div(K, 64), ldw = ptt_linear_ldw(Cout);
  const long long total = (long long)NB * KB * 512;
  dim3 grid((unsigned)llmin_((total + 255) / 256, 2048), 1);
  tc_pack_weight_kernel<<<grid, 256, 0, st>>>(w, ld_c, ld_k, Cout, K, NB, KB, nullptr,
                                              reinterpret_cast<__half*>(params + (size_t)(K + 1) * ldw), 0, 0, params, ldw, bias);
  PTT_LAUNCHED();
  return ptt_launch_status();
}

bool ptt_tc_gemm_supported(const PttGemmArgs& a) {
  return a.R > 0 && a.N > 0 && a.K > 0 && (a.ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15u) == 0) &&
         round_up(a.K, 4) <= a.ldx;
}

int ptt_tc_gemm_launch(const PttGemmArgs& a, const void* wimg, cudaStream_t st) {
  TcParams p;
  p.g = a;
  p.wimg = static_cast<const __half*>(wimg);
  p.n_wblocks = ceil_div(a.N, 64);
  p.k_blocks = ceil_div(a.K, 64);
  // Tile width by a small cost model: these problems are a few waves of short pipelines, so what matters is the number
  // of waves times (fixed per-CTA latency + k-blocks x the slower of producer and MMA issue per k-block), in cycles.
  const long long row_tiles = ceil_div(a.R, TBM);
  const int batch = a.batch > 0 ? a.batch : 1;
  const int bn_max = a.N <= 64 ? 64 : (a.N <= 128 ? 128 : 256);
  int bn = bn_max;
  long long best = -1;
  for (int cand = bn_max; cand >= 64; cand >>= 1) {
    const long long ctas = row_tiles * ceil_div(a.N, cand) * batch;
    const long long waves = (ctas + 147) / 148;
    const long long per_kb = 12LL * (cand / 2 > 70 ? cand / 2 : 70);
    const long long t = waves * (6000 + (long long)p.k_blocks * (per_kb > 800 ? per_kb : 800));
    if (best < 0 || t < best) { best = t; bn = cand; }
  }
  if (bn == 64) return tc_launch<64>(p, st);
  if (bn == 128) return tc_launch<128>(p, st);
  return tc_launch<256>(p, st);
}
