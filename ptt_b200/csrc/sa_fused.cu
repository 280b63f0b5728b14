// Fused set-abstraction layer body for sm_100a: neighbour gather + 3-layer shared MLP + max over nsample in ONE
// persistent tcgen05 kernel; the grouped (B, C, M, ns) tensor and both hidden activations never touch HBM.
//
//   reference: QueryAndGroup (pointnet2_utils.py:320-380) -> SharedMLP (pytorch_utils.py:12-36) -> max_pool2d
//              (pointnet2_modules.py:83-88), eval mode, BatchNorm folded.
//
// Layer 1 is separable: the grouped input row of pair (centre j, neighbour i) is [feats[i] | (xyz[i]-c_j)/r], so
//   scale1 * (W1 . row) + shift1 = G'[i] + Wx' . rel,      G' = scale1 * (feats . W1f^T) + shift1   (per POINT, not per
//   pair: M*ns/N = 16x fewer rows, one small tcgen05 contraction beforehand),   Wx' = scale1 * W1x  (3 columns).
// So this kernel gathers G' rows (D1 floats) instead of feature rows, finishes layer 1 on CUDA cores (3 FMA + ReLU per
// element), and runs layers 2 and 3 on the tensor cores with the fp16 hi/lo split of tc_gemm.cu (3 MMAs per K-step).
//
// One CTA per SM, 128 pair-rows (128/ns centres) per tile, static round-robin over tiles.  Warp roles:
//   warps 0-7   epilogue (warp w: TMEM lane quarter w%4, column half w/4):
//                 D2 (TMEM) -> +shift, ReLU -> fp16 hi/lo -> HB (smem, UMMA K-major SW128);
//                 D3 (TMEM) -> +shift, ReLU -> max over each centre's ns rows (redux.sync) -> out (B, M, D3)
//   warps 8-15  producers: per-row (point index, rel xyz) -> gather G' -> layer 1 -> fp16 hi/lo -> HA (smem)
//   warp  16    weight loader: W2 / W3 as 16 KB (64 out-channels x 64 k, hi+lo) blocks through a 5-stage ring (TMA engine)
//   warp  17    TMEM allocation + the single thread issuing tcgen05.mma / tcgen05.commit
// BatchNorm scales of layers 2 and 3 are folded into the packed weights, so both epilogues are acc + shift.
#include "sa_fused.cuh"
#include "tc_common.cuh"

namespace {

constexpr int SF_THREADS = 576;            // 8 epilogue + 8 producer warps + loader + MMA
constexpr int SF_TM = 128;
constexpr int SF_NS = 5;                       // ring stages
constexpr uint32_t SF_ITEM_BYTES = 16384;      // one weight block: 64 rows x 64 k, hi (8 KB) + lo (8 KB)
constexpr uint32_t SF_KBLOCK_BYTES = 32768;    // one activation k-block: 128 rows x 64 k, hi (16 KB) + lo (16 KB)

template <int D1, int D2, int D3>
struct SfCfg {
  static constexpr int KB1 = D1 / 64, KB2 = D2 / 64, NB2 = D2 / 64, NB3 = D3 / 64;
  static constexpr uint32_t HA_BYTES = KB1 * SF_KBLOCK_BYTES;
  static constexpr uint32_t HB_BYTES = KB2 * SF_KBLOCK_BYTES;
  static constexpr uint32_t RING_BYTES = SF_NS * SF_ITEM_BYTES;
  static constexpr uint32_t CTRL_BYTES = 256 + 128 * 16 + (D2 + D3) * 4;
  static constexpr uint32_t SMEM_BYTES = HA_BYTES + HB_BYTES + RING_BYTES + CTRL_BYTES + 1024;
  static constexpr uint32_t D2_COL = 0, D3_COL = 128, TMEM_COLS = 512;
};

__device__ __forceinline__ void producer_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack2(__half a, __half b) {
  return (uint32_t)__half_as_ushort(a) | ((uint32_t)__half_as_ushort(b) << 16);
}

template <int D1, int D2, int D3>
__global__ void __launch_bounds__(SF_THREADS, 1) sa_fused_kernel(const __grid_constant__ SaFusedArgs a, const int num_tiles) {
  using Cfg = SfCfg<D1, D2, D3>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // stays a shared-space pointer
  uint8_t* HA = smem;
  uint8_t* HB = HA + Cfg::HA_BYTES;
  uint8_t* ring = HB + Cfg::HB_BYTES;
  uint8_t* ctrl = ring + Cfg::RING_BYTES;
  uint64_t* full_b = reinterpret_cast<uint64_t*>(ctrl);   // [SF_NS]
  uint64_t* empty = full_b + SF_NS;                        // [SF_NS]
  uint64_t* ha_full = empty + SF_NS;
  uint64_t* ha_free = ha_full + 1;
  uint64_t* d2_full = ha_free + 1;
  uint64_t* hb_full = d2_full + 1;
  uint64_t* d3_full = hb_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d3_full + 1);
  float4* s_info = reinterpret_cast<float4*>(ctrl + 256);                 // [128] {rel.xyz, point row as int bits}
  float* s_sh2 = reinterpret_cast<float*>(ctrl + 256 + 128 * 16);         // [D2]
  float* s_sh3 = s_sh2 + D2;                                              // [D3]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int c = tid; c < D2; c += SF_THREADS) s_sh2[c] = __ldg(a.shift2 + c);
  for (int c = tid; c < D3; c += SF_THREADS) s_sh3[c] = __ldg(a.shift3 + c);
  if (tid == 0) {
    for (int s = 0; s < SF_NS; ++s) {
      tc::mbar_init(&full_b[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(ha_full, 256);
    tc::mbar_init(ha_free, 1);
    tc::mbar_init(d2_full, 1);
    tc::mbar_init(hb_full, 256);
    tc::mbar_init(d3_full, 1);
    tc::mbar_init_fence();
  }
  if (warp == 17) tc::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int ns = a.ns;

  if (warp < 8) {
    // =================================================================== epilogue warps
    const int quarter = warp & 3, half = warp >> 2;
    const int r = quarter * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const unsigned gmask = ns >= 32 ? 0xffffffffu : (((1u << ns) - 1u) << (lane & ~(ns - 1)));
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const uint32_t par = it & 1;
      // ---- E2: D2 -> H2 (HB)
      tc::mbar_wait(d2_full, par);
      tc::tc_fence_after();
#pragma unroll 1
      for (int c0 = half * 32; c0 < D2; c0 += 64) {
        float v[32];
        tc::tmem_ld32(lane_addr + Cfg::D2_COL + (uint32_t)c0, v);
        uint8_t* blk = HB + (c0 >> 6) * SF_KBLOCK_BYTES;
        const int chunk0 = (c0 & 63) >> 3;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          __half hi[8], lo[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float h = fmaxf(v[ch * 8 + u] + s_sh2[c0 + ch * 8 + u], 0.f);
            tc::split_f16(h, hi[u], lo[u]);
          }
          const uint32_t off = tc::sw128_offset(r, chunk0 + ch);
          *reinterpret_cast<uint4*>(blk + off) =
              make_uint4(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], hi[5]), pack2(hi[6], hi[7]));
          *reinterpret_cast<uint4*>(blk + 16384 + off) =
              make_uint4(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]), pack2(lo[4], lo[5]), pack2(lo[6], lo[7]));
        }
      }
      tc::tc_fence_before();
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(hb_full);
      // ---- E3: D3 -> max over each centre's rows -> out
      tc::mbar_wait(d3_full, par);
      tc::tc_fence_after();
      const long long R = (long long)tile * SF_TM + r;
      const bool writer = (lane & (ns - 1)) == 0 && R < a.rows;
      float* orow = a.out_pm + (R / ns) * (long long)a.ld_out;
#pragma unroll 1
      for (int c0 = half * 32; c0 < D3; c0 += 64) {
        float v[32];
        tc::tmem_ld32(lane_addr + Cfg::D3_COL + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float y = v[j] + s_sh3[c0 + j];
          y = y > 0.f ? y : 0.f;                               // ReLU; exactly +0 so that uint order == float order
          v[j] = __uint_as_float(__reduce_max_sync(gmask, __float_as_uint(y)));
        }
        if (writer) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      }
      tc::tc_fence_before();
    }
  } else if (warp < 16) {
    // =================================================================== producers: layer 1 on CUDA cores
    const int pt = tid - 256;            // 0..255
    const int q = pt & 15;               // float4 column (and q + 16 when D1 == 128)
    const int rsub = pt >> 4;            // 0..15
    constexpr int NQ = D1 / 64;          // float4 columns per thread
    float wx[NQ][3][4], g0[NQ][4];
#pragma unroll
    for (int h = 0; h < NQ; ++h)
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = (q + 16 * h) * 4 + u;
#pragma unroll
        for (int d = 0; d < 3; ++d) wx[h][d][u] = __ldg(a.wx + d * D1 + c);
        g0[h][u] = a.gprime ? 0.f : __ldg(a.shift1 + c);
      }
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      // per-row neighbour info
      producer_bar();                                   // everyone is done reading s_info of the previous tile
      if (pt < SF_TM) {
        const long long R = (long long)tile * SF_TM + pt;
        float4 info = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
        if (R < a.rows) {
          const long long cj = R / ns;
          const long long b = cj / a.M;
          const long long src = b * a.N + __ldg(a.idx + R);
          const float* p = a.xyz + src * 3;
          const float* c = a.new_xyz + cj * 3;
          float dx = __fsub_rn(__ldg(p), __ldg(c)), dy = __fsub_rn(__ldg(p + 1), __ldg(c + 1)), dz = __fsub_rn(__ldg(p + 2), __ldg(c + 2));
          if (a.normalize) { dx = __fdiv_rn(dx, a.radius); dy = __fdiv_rn(dy, a.radius); dz = __fdiv_rn(dz, a.radius); }
          info = make_float4(dx, dy, dz, __int_as_float((int)src));
        }
        s_info[pt] = info;
      }
      producer_bar();
      tc::mbar_wait(ha_free, (uint32_t)(it & 1) ^ 1u);  // GEMM2 of the previous tile has consumed HA
#pragma unroll 4
      for (int i = 0; i < 8; ++i) {
        const int r = i * 16 + rsub;
        const float4 info = s_info[r];
        const int src = __float_as_int(info.w);
#pragma unroll
        for (int h = 0; h < NQ; ++h) {
          float4 g = make_float4(g0[h][0], g0[h][1], g0[h][2], g0[h][3]);
          if (a.gprime != nullptr && src >= 0) g = __ldg(reinterpret_cast<const float4*>(a.gprime + (size_t)src * D1 + (q + 16 * h) * 4));
          float y[4] = {g.x, g.y, g.z, g.w};
          __half hi[4], lo[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float t = fmaf(wx[h][0][u], info.x, y[u]);
            t = fmaf(wx[h][1][u], info.y, t);
            t = fmaf(wx[h][2][u], info.z, t);
            t = src >= 0 ? fmaxf(t, 0.f) : 0.f;
            tc::split_f16(t, hi[u], lo[u]);
          }
          uint8_t* blk = HA + h * SF_KBLOCK_BYTES;
          const uint32_t off = tc::sw128_offset(r, q >> 1) + ((q & 1) << 3);
          *reinterpret_cast<uint2*>(blk + off) = make_uint2(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]));
          *reinterpret_cast<uint2*>(blk + 16384 + off) = make_uint2(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]));
        }
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(ha_full);
    }
  } else if (warp == 16) {
    // =================================================================== weight loader
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint8_t* w2 = static_cast<const uint8_t*>(a.w2img);
      const uint8_t* w3 = static_cast<const uint8_t*>(a.w3img);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
#pragma unroll 1
        for (int item = 0; item < Cfg::KB1 * Cfg::NB2 + Cfg::KB2 * Cfg::NB3; ++item) {
          const uint8_t* src;
          if (item < Cfg::KB1 * Cfg::NB2) {
            const int kb = item / Cfg::NB2, nb = item % Cfg::NB2;
            src = w2 + (size_t)(nb * Cfg::KB1 + kb) * SF_ITEM_BYTES;
          } else {
            const int j = item - Cfg::KB1 * Cfg::NB2;
            const int kb = j / Cfg::NB3, nb = j % Cfg::NB3;
            src = w3 + (size_t)(nb * Cfg::KB2 + kb) * SF_ITEM_BYTES;
          }
          tc::mbar_wait(&empty[stage], phase ^ 1);
          tc::mbar_arrive_expect_tx(&full_b[stage], SF_ITEM_BYTES);
          tc::bulk_g2s(ring + stage * SF_ITEM_BYTES, src, SF_ITEM_BYTES, &full_b[stage]);
          if (++stage == SF_NS) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // =================================================================== MMA issuer (whole warp, one elected lane issues)
    {
      constexpr uint32_t IDESC = tc::idesc_f16<false>(SF_TM, 64);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      const uint32_t ha_addr = tc::smem_u32(HA), hb_addr = tc::smem_u32(HB), ring_addr = tc::smem_u32(ring);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const uint32_t par = it & 1;
        // ---- layer 2: D2 = H1 . W2'^T
        tc::mbar_wait(ha_full, par);
        tc::tc_fence_after();
#pragma unroll 1
        for (int kb = 0; kb < Cfg::KB1; ++kb) {
#pragma unroll 1
          for (int nb = 0; nb < Cfg::NB2; ++nb) {
            tc::mbar_wait(&full_b[stage], phase);
            tc::tc_fence_after();
            const uint64_t da_hi = tc::smem_desc_sw128(ha_addr + kb * SF_KBLOCK_BYTES);
            const uint64_t da_lo = tc::smem_desc_sw128(ha_addr + kb * SF_KBLOCK_BYTES + 16384);
            const uint64_t db_hi = tc::smem_desc_sw128(ring_addr + stage * SF_ITEM_BYTES);
            const uint64_t db_lo = tc::smem_desc_sw128(ring_addr + stage * SF_ITEM_BYTES + 8192);
            const uint32_t d = tmem_base + Cfg::D2_COL + nb * 64;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = (uint64_t)(k * 2);
              tc::mma_f16_w(d, da_hi + adv, db_hi + adv, IDESC, (kb | k) != 0);
              tc::mma_f16_w(d, da_lo + adv, db_hi + adv, IDESC, 1);
              tc::mma_f16_w(d, da_hi + adv, db_lo + adv, IDESC, 1);
            }
            tc::mma_commit_w(&empty[stage]);
            if (++stage == SF_NS) { stage = 0; phase ^= 1; }
          }
        }
        tc::mma_commit_w(ha_free);
        tc::mma_commit_w(d2_full);
        // ---- layer 3: D3 = H2 . W3'^T
        tc::mbar_wait(hb_full, par);
        tc::tc_fence_after();
#pragma unroll 1
        for (int kb = 0; kb < Cfg::KB2; ++kb) {
#pragma unroll 1
          for (int nb = 0; nb < Cfg::NB3; ++nb) {
            tc::mbar_wait(&full_b[stage], phase);
            tc::tc_fence_after();
            const uint64_t da_hi = tc::smem_desc_sw128(hb_addr + kb * SF_KBLOCK_BYTES);
            const uint64_t da_lo = tc::smem_desc_sw128(hb_addr + kb * SF_KBLOCK_BYTES + 16384);
            const uint64_t db_hi = tc::smem_desc_sw128(ring_addr + stage * SF_ITEM_BYTES);
            const uint64_t db_lo = tc::smem_desc_sw128(ring_addr + stage * SF_ITEM_BYTES + 8192);
            const uint32_t d = tmem_base + Cfg::D3_COL + nb * 64;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = (uint64_t)(k * 2);
              tc::mma_f16_w(d, da_hi + adv, db_hi + adv, IDESC, (kb | k) != 0);
              tc::mma_f16_w(d, da_lo + adv, db_hi + adv, IDESC, 1);
              tc::mma_f16_w(d, da_hi + adv, db_lo + adv, IDESC, 1);
            }
            tc::mma_commit_w(&empty[stage]);
            if (++stage == SF_NS) { stage = 0; phase ^= 1; }
          }
        }
        tc::mma_commit_w(d3_full);
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int D1, int D2, int D3>
int sf_launch(const SaFusedArgs& a, cudaStream_t st) {
  using Cfg = SfCfg<D1, D2, D3>;
  auto kern = sa_fused_kernel<D1, D2, D3>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  const long long tiles = (a.rows + SF_TM - 1) / SF_TM;
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)llmin_(tiles, sms);
  kern<<<grid, SF_THREADS, Cfg::SMEM_BYTES, st>>>(a, (int)tiles); PTT_LAUNCHED();
  return ptt_launch_status();
}

}  // namespace

bool sa_fused_supported(int d1, int d2, int d3, int ns) {
  const bool dims = (d1 == 64 && d2 == 64 && d3 == 128) || (d1 == 128 && d2 == 128 && d3 == 256);
  const bool group = ns >= 1 && ns <= 32 && (ns & (ns - 1)) == 0;
  return dims && group;
}

int sa_fused_launch(const SaFusedArgs& a, int d1, int d2, int d3, cudaStream_t st) {
  if (a.rows <= 0) return PTT_OK;
  if (d1 == 64 && d2 == 64 && d3 == 128) return sf_launch<64, 64, 128>(a, st);
  if (d1 == 128 && d2 == 128 && d3 == 256) return sf_launch<128, 128, 256>(a, st);
  return PTT_ERR_UNSUPPORTED;
}
