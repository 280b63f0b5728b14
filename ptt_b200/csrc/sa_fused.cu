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
//   warps 0-7   epilogue (warp w: TMEM lane quarter w%4, column half / channel block w/4):
//                 E2: D2 (TMEM, rows in lanes) -> +shift, ReLU -> fp16 hi/lo -> H2 ring slot (smem, UMMA K-major SW128);
//                 E3: D3^T (TMEM, CHANNELS in lanes: layer 3 is issued transposed) -> max over each centre's ns rows
//                     along the thread's registers -> +shift, ReLU -> out (B, M, D3), 32 consecutive channels per store
//   warps 8-15  producers: per-row (point index, rel xyz) -> gather G' -> layer 1 -> fp16 hi/lo -> H1 ring slot
//   warp  16    weight loader: W2 / W3 as (k-block, <=128 out-channels) items through a 3-stage ring (TMA engine)
//   warp  17    TMEM allocation + tcgen05.mma issue (whole warp convergent, one elected lane)
// H1 / H2 are handed over one k-block at a time through small rings (SfCfg); the single-k-block stack keeps two tiles
// in flight (SfCfg::DEEP).  BatchNorm scales of layers 2 and 3 are folded into the packed weights, so both epilogues are
// acc + shift.  The same kernel serves CosineSimAug (pair_scalar, implicit groups of all N points: sa_mlp.cu).
#include "sa_fused.cuh"
#include "ptt_b200_tuning.h"
#include "tc_common.cuh"

namespace {

constexpr int SF_THREADS = 576;            // 8 epilogue + 8 producer warps + loader + MMA
constexpr int SF_TM = 128;
constexpr int SF_NS = 3;                       // weight ring stages
constexpr uint32_t SF_STAGE_BYTES = 32768;     // one weight item: <=128 rows x 64 k, hi + lo
constexpr uint32_t SF_KBLOCK_BYTES = 32768;    // one activation k-block: 128 rows x 64 k, hi (16 KB) + lo (16 KB)

// H1 and H2 live in RINGS of k-block slots (NRA / NRB of them) handed over one k-block at a time: producers -> layer 2 and
// epilogue -> layer 3.  Up to 128-wide hidden layers a ring holds the whole activation (slot == k-block); at 256 it holds
// two k-blocks, which is what lets the (256, 256, 256) stack of the box head fit next to the weight ring.
template <int D1, int D2, int D3>
struct SfCfg {
  static constexpr int KB1 = D1 / 64, KB2 = D2 / 64;
  // DEEP (single-k-block layers, i.e. the 64-wide SA1 stack): TWO tiles in flight -- H1 / H2 slots and both accumulators
  // exist per tile parity, the MMA warp issues layer 2 of tile t+1 BEFORE layer 3 of tile t and the epilogue converts
  // H2(t+1) BEFORE it reduces D3(t), so none of the per-tile hand-offs (producer -> MMA -> epilogue -> MMA -> epilogue)
  // sits on the critical path any more.  Wider stacks have neither the shared memory nor the TMEM columns for that.
  static constexpr bool DEEP = KB1 == 1 && KB2 == 1 && D2 <= 64 && D3 <= 128;
  static constexpr int NRA = DEEP ? 2 : (KB1 < 2 ? KB1 : 2), NRB = DEEP ? 2 : (KB2 < 2 ? KB2 : 2);
  static constexpr int NI2 = D2 < 128 ? D2 : 128, NI3 = D3 < 128 ? D3 : 128;   // MMA N / rows per weight item
  static constexpr int ITEMS2 = D2 / NI2, ITEMS3 = D3 / NI3;                   // items per k-block
  static constexpr bool SH2_SMEM = D2 < 256;                                    // shift2 staged in shared memory when it fits
  static constexpr uint32_t HA_BYTES = NRA * SF_KBLOCK_BYTES;
  static constexpr uint32_t HB_BYTES = NRB * SF_KBLOCK_BYTES;
  static constexpr uint32_t RING_BYTES = SF_NS * SF_STAGE_BYTES;
  static constexpr uint32_t CTRL_BYTES = 256 + 128 * 16 + (SH2_SMEM ? D2 * 4 : 0);
  static constexpr uint32_t SMEM_BYTES = HA_BYTES + HB_BYTES + RING_BYTES + CTRL_BYTES;   // 231,680 B at (256,256,256); limit 232,448
  static constexpr uint32_t D2_COL = 0, D3_COL = D2 < 128 ? 128 : D2, TMEM_COLS = 512;   // D3^T: ITEMS3 blocks of 128 columns
  static constexpr uint32_t D2_STRIDE = DEEP ? 64 : 0, D3_STRIDE = DEEP ? 128 : 0;       // per tile parity (DEEP)
  static_assert(D3_COL + ITEMS3 * 128 + D3_STRIDE <= 512 && D2_COL + D2 + D2_STRIDE <= D3_COL, "accumulators exceed TMEM");
};

__device__ __forceinline__ void producer_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// weight item (kb, ni) of a (Dout, K) image made of 8 KB (64 rows x 64 k) hi / lo blocks -> one ring stage laid out
// [hi rows 0..NI) | lo rows 0..NI)]
template <int NI>
__device__ __forceinline__ void sf_load_item(uint8_t* dst, const uint8_t* img, int KB, int kb, int ni, uint64_t* bar) {
  tc::mbar_arrive_expect_tx(bar, NI * 256);
  if (NI == 128) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint8_t* src = img + (size_t)((ni * 2 + j) * KB + kb) * 16384;
      tc::bulk_g2s(dst + j * 8192, src, 8192, bar);
      tc::bulk_g2s(dst + 16384 + j * 8192, src + 8192, 8192, bar);
    }
  } else {
    tc::bulk_g2s(dst, img + (size_t)(ni * KB + kb) * 16384, 16384, bar);
  }
}

// E3 of one 32-row block held by one thread (its output channel): max over every group of NSW consecutive rows, then
// shift + ReLU and the store to (centre, channel).  dst = the block's first centre; groups wider than the block (ns > 32)
// are combined through the zero-initialised output with atomicMax (values are >= +0, so unsigned order == float order).
template <int NSW>
__device__ __forceinline__ void sf_store_groups(float (&v)[32], long long R0, long long rows, float shift, float* dst, int ld_out,
                                                bool atomic) {
#pragma unroll
  for (int o = 1; o < NSW; o <<= 1)
#pragma unroll
    for (int j = 0; j < 32; j += 2 * o) v[j] = fmaxf(v[j], v[j + o]);
#pragma unroll
  for (int g = 0; g < 32 / NSW; ++g) {
    if (R0 + g * NSW < rows) {
      const float y = fmaxf(v[g * NSW] + shift, 0.f);
      float* p = dst + (long long)g * ld_out;
      if (atomic) atomicMax(reinterpret_cast<unsigned int*>(p), __float_as_uint(y));
      else *p = y;
    }
  }
}

template <int D1, int D2, int D3>
__global__ void __launch_bounds__(SF_THREADS, 1) sa_fused_kernel(const __grid_constant__ SaFusedArgs a, const int num_tiles) {
  using Cfg = SfCfg<D1, D2, D3>;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* HA = smem;
  uint8_t* HB = HA + Cfg::HA_BYTES;
  uint8_t* ring = HB + Cfg::HB_BYTES;
  uint8_t* ctrl = ring + Cfg::RING_BYTES;
  uint64_t* full_b = reinterpret_cast<uint64_t*>(ctrl);   // [SF_NS]
  uint64_t* empty = full_b + SF_NS;                        // [SF_NS]
  uint64_t* ha_full = empty + SF_NS;                       // [NRA] producers -> MMA
  uint64_t* ha_free = ha_full + Cfg::NRA;                  // [NRA] MMA -> producers
  uint64_t* hb_full = ha_free + Cfg::NRA;                  // [NRB] epilogue -> MMA
  uint64_t* hb_free = hb_full + Cfg::NRB;                  // [NRB] MMA -> epilogue
  uint64_t* d2_full = hb_free + Cfg::NRB;                  // [2] (one per tile parity; DEEP uses both)
  uint64_t* d3_full = d2_full + 2;                         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d3_full + 2);
  float4* s_info = reinterpret_cast<float4*>(ctrl + 256);                 // [128] {rel.xyz, point row as int bits}
  float* s_sh2 = reinterpret_cast<float*>(ctrl + 256 + 128 * 16);         // [D2] (shift3 stays in global: E3 is off the critical path)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if ((tc::smem_u32(smem) & 1023u) != 0) __trap();          // the UMMA tiles need 1024-byte alignment
  if (Cfg::SH2_SMEM)
    for (int c = tid; c < D2; c += SF_THREADS) s_sh2[c] = __ldg(a.shift2 + c);

  if (tid == 0) {
    for (int s = 0; s < SF_NS; ++s) {
      tc::mbar_init(&full_b[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < Cfg::NRA; ++s) {
      tc::mbar_init(&ha_full[s], 8);
      tc::mbar_init(&ha_free[s], 1);
    }
    for (int s = 0; s < Cfg::NRB; ++s) {
      tc::mbar_init(&hb_full[s], 8);
      tc::mbar_init(&hb_free[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&d2_full[s], 1);
      tc::mbar_init(&d3_full[s], 1);
    }
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
    // E3 ownership: this warp's 32 lanes are 32 consecutive output channels of one 128-channel block of layer 3
    static_assert(Cfg::NI3 == 128 && (Cfg::ITEMS3 == 1 || Cfg::ITEMS3 == 2), "layer 3 is issued transposed in 128-channel blocks");
    constexpr int E3_BLOCKS = Cfg::ITEMS3 == 2 ? 4 : 2;    // 32-row column groups handled by this warp per tile
    const int e3_mb = Cfg::ITEMS3 == 2 ? half : 0;
    const int e3_ch = e3_mb * 128 + quarter * 32 + lane;
    const float e3_shift = __ldg(a.shift3 + e3_ch);
    const int nsw = ns < 32 ? ns : 32;
    const int log2ns = 31 - __clz(ns);
    int sb = 0;                 // H2 ring slot
    uint32_t pb = 0;            // its phase
    // buffer / barrier parity of local tile `it`: DEEP alternates two accumulator sets, otherwise there is one
    auto e2 = [&](int it, int tile) {
      const uint32_t p = Cfg::DEEP ? (uint32_t)(it & 1) : 0u, par = Cfg::DEEP ? (uint32_t)((it >> 1) & 1) : (uint32_t)(it & 1);
      const bool erec = a.dbg != nullptr && blockIdx.x == 0 && tid == 0 && it < 30;
      (void)tile;
      // ---- E2: D2 -> H2, one 64-column k-block at a time so that layer 3 can start on the first one
      tc::mbar_wait(&d2_full[p], par);
      tc::tc_fence_after();
      if (erec) a.dbg[1000 + it * 4 + 0] = clock64();
#pragma unroll 1
      for (int kb = 0; kb < Cfg::KB2; ++kb) {
        const int c0 = kb * 64 + half * 32;
        float v[32];
        tc::tmem_ld32(lane_addr + Cfg::D2_COL + p * Cfg::D2_STRIDE + (uint32_t)c0, v);
        tc::mbar_wait(&hb_free[sb], pb ^ 1);            // layer 3 has consumed the k-block that used this slot
        uint8_t* blk = HB + sb * SF_KBLOCK_BYTES;
        const int chunk0 = half * 4;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float4 s0, s1;
          if (Cfg::SH2_SMEM) {
            s0 = *reinterpret_cast<const float4*>(s_sh2 + c0 + ch * 8);
            s1 = *reinterpret_cast<const float4*>(s_sh2 + c0 + ch * 8 + 4);
          } else {
            s0 = __ldg(reinterpret_cast<const float4*>(a.shift2 + c0 + ch * 8));
            s1 = __ldg(reinterpret_cast<const float4*>(a.shift2 + c0 + ch * 8 + 4));
          }
          const float sh[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
          float h[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) h[u] = fmaxf(v[ch * 8 + u] + sh[u], 0.f);
          uint4 hi, lo;
          tc::split_f16x2(h[0], h[1], hi.x, lo.x);
          tc::split_f16x2(h[2], h[3], hi.y, lo.y);
          tc::split_f16x2(h[4], h[5], hi.z, lo.z);
          tc::split_f16x2(h[6], h[7], hi.w, lo.w);
          const uint32_t off = tc::sw128_offset(r, chunk0 + ch);
          *reinterpret_cast<uint4*>(blk + off) = hi;
          *reinterpret_cast<uint4*>(blk + 16384 + off) = lo;
        }
        tc::tc_fence_before();
        tc::fence_proxy_async_smem();
        tc::mbar_arrive_warp(&hb_full[sb]);
        if (++sb == Cfg::NRB) { sb = 0; pb ^= 1; }
      }
      if (erec) a.dbg[1000 + it * 4 + 1] = clock64();
    };
    auto e3 = [&](int it, int tile) {
      const uint32_t p = Cfg::DEEP ? (uint32_t)(it & 1) : 0u, par = Cfg::DEEP ? (uint32_t)((it >> 1) & 1) : (uint32_t)(it & 1);
      const bool erec = a.dbg != nullptr && blockIdx.x == 0 && tid == 0 && it < 30;
      // ---- E3: D3 -> max over each centre's rows -> out
      tc::mbar_wait(&d3_full[p], par);
      tc::tc_fence_after();
      if (erec) a.dbg[1000 + it * 4 + 2] = clock64();
      // Layer 3 is computed TRANSPOSED (D3^T = W3' . H2^T: output channels in the TMEM lanes, the tile's 128 pair-rows in
      // the columns), so the max over a centre's ns rows runs along the registers of ONE thread -- no shuffles, no warp
      // reductions -- and the 32 lanes of a warp write 32 consecutive channels of a centre (one coalesced 128-byte store).
      // shift + ReLU commute with the max and are applied to the survivors only.
#pragma unroll 1
      for (int blk = 0; blk < E3_BLOCKS; ++blk) {
        const int cg = Cfg::ITEMS3 == 2 ? blk : half * 2 + blk;        // 32-row column group of the tile
        float v[32];
        tc::tmem_ld32(lane_addr + Cfg::D3_COL + p * Cfg::D3_STRIDE + (uint32_t)(e3_mb * SF_TM + cg * 32), v);
        const long long R0 = (long long)tile * SF_TM + cg * 32;          // first pair-row of this block
        float* dst = a.out_pm + (R0 >> log2ns) * (long long)a.ld_out + e3_ch;
        switch (nsw) {                                                   // uniform: straight-line code per group width
          case 32: sf_store_groups<32>(v, R0, a.rows, e3_shift, dst, a.ld_out, ns > 32); break;
          case 16: sf_store_groups<16>(v, R0, a.rows, e3_shift, dst, a.ld_out, false); break;
          case 8: sf_store_groups<8>(v, R0, a.rows, e3_shift, dst, a.ld_out, false); break;
          case 4: sf_store_groups<4>(v, R0, a.rows, e3_shift, dst, a.ld_out, false); break;
          case 2: sf_store_groups<2>(v, R0, a.rows, e3_shift, dst, a.ld_out, false); break;
          default: sf_store_groups<1>(v, R0, a.rows, e3_shift, dst, a.ld_out, false); break;
        }
      }
      tc::tc_fence_before();
      if (erec) a.dbg[1000 + it * 4 + 3] = clock64();
    };
    const int stride = gridDim.x;
    if (Cfg::DEEP) {
      if ((int)blockIdx.x < num_tiles) e2(0, blockIdx.x);
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += stride, ++it) {
        if (tile + stride < num_tiles) e2(it + 1, tile + stride);
        e3(it, tile);
      }
    } else {
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += stride, ++it) {
        e2(it, tile);
        e3(it, tile);
      }
    }
  } else if (warp < 16) {
    // =================================================================== producers: layer 1 on CUDA cores
    const int pt = tid - 256;            // 0..255
    const int q = pt & 15;               // float4 column of the k-block
    const int rsub = pt >> 4;            // 0..15
    // Per-row neighbour info is software-pipelined: the dependent global loads (ball-query index -> point) of tile
    // t+1 are issued before H1 of tile t is written, so their latency never sits on the producers' critical path.
    auto row_index = [&](int tile) -> int {          // ball-query result of this thread's row, -1 beyond the end
      const long long R = (long long)tile * SF_TM + pt;
      if (!(pt < SF_TM && tile < num_tiles && R < a.rows)) return -1;
      return a.idx ? __ldg(a.idx + R) : (int)(R & (ns - 1));     // ns is a power of two
    };
    // The loads of a row's point and centre are issued one tile ahead and their values stay RAW in registers until the
    // top of the next tile: any arithmetic on them here would make the warp wait for the loads it has just issued.
    struct RowRaw { float px, py, pz, cx, cy, cz; int src; };
    const int plog2ns = 31 - __clz(ns);
    const float inv_radius = 1.f / a.radius;
    auto row_raw = [&](int tile, int i) -> RowRaw {
      RowRaw w = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, -1};
      if (i >= 0) {
        const long long R = (long long)tile * SF_TM + pt;
        const int cj = (int)(R >> plog2ns);                 // ns is a power of two here; 32-bit index arithmetic
        const int src = (cj / a.M) * a.N + i;
        w.src = src;
        if (a.pair_scalar != nullptr) {
          w.px = __ldg(a.pair_scalar + R);
        } else {
          const float* p = a.xyz + (long long)src * 3;
          const float* c = a.new_xyz + (long long)cj * 3;
          w.px = __ldg(p); w.py = __ldg(p + 1); w.pz = __ldg(p + 2);
          w.cx = __ldg(c); w.cy = __ldg(c + 1); w.cz = __ldg(c + 2);
        }
      }
      return w;
    };
    auto row_info = [&](const RowRaw& w) -> float4 {
      if (w.src < 0) return make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
      if (a.pair_scalar != nullptr) return make_float4(w.px, 0.f, 0.f, __int_as_float(w.src));
      float dx = __fsub_rn(w.px, w.cx), dy = __fsub_rn(w.py, w.cy), dz = __fsub_rn(w.pz, w.cz);
      if (a.normalize) {
        // x / radius as a branch-free reciprocal + one Newton correction (correctly rounded for normal operands): the IEEE
        // division's slow-path branches cost ~1.4 k cycles per tile here (3 divisions, ~4 warps per scheduler)
        auto div_r = [&](float x) { const float q = x * inv_radius; return fmaf(fmaf(-q, a.radius, x), inv_radius, q); };
        dx = div_r(dx); dy = div_r(dy); dz = div_r(dz);
      }
      return make_float4(dx, dy, dz, __int_as_float(w.src));
    };
    // layer-1 xyz weights of this thread's channels: kept in registers for up to two k-blocks, re-read (L1) beyond
    constexpr bool HOIST = Cfg::KB1 <= 1;   // beyond one k-block the registers go to the gather prefetch instead
    constexpr int HK = HOIST ? Cfg::KB1 : 1;
    float wxr[HK][3][4], g0r[HK][4];
    if (HOIST) {
#pragma unroll
      for (int h = 0; h < HK; ++h)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = h * 64 + q * 4 + u;
#pragma unroll
          for (int d = 0; d < 3; ++d) wxr[h][d][u] = __ldg(a.wx + d * D1 + c);
          g0r[h][u] = a.gprime ? 0.f : __ldg(a.shift1 + c);
        }
    }
    const int stride = gridDim.x;
    int idx_next = row_index(blockIdx.x + stride);              // index for tile t+1
    RowRaw raw_cur = row_raw(blockIdx.x, row_index(blockIdx.x));
    int it = 0;
    int sa = 0;                 // H1 ring slot
    uint32_t pa = 0;            // its phase
    for (int tile = blockIdx.x; tile < num_tiles; tile += stride, ++it) {
      const bool prec0 = a.dbg != nullptr && blockIdx.x == 0 && pt == 0 && it < 30;
      if (prec0) a.dbg[3000 + it * 4 + 0] = clock64();
      producer_bar();                                   // everyone is done reading s_info of the previous tile
      if (prec0) a.dbg[3000 + it * 4 + 1] = clock64();
      if (pt < SF_TM) s_info[pt] = row_info(raw_cur);
      if (prec0) a.dbg[3000 + it * 4 + 3] = clock64();
      producer_bar();
      if (prec0) a.dbg[3000 + it * 4 + 2] = clock64();
      // issue the loads for the next two tiles now; they are consumed at the top of the next tile
      const RowRaw raw_nxt = row_raw(tile + stride, idx_next);
      const int idx_nn = row_index(tile + 2 * stride);
      const bool prec = a.dbg != nullptr && blockIdx.x == 0 && pt == 0 && it < 30;
      if (prec) a.dbg[2000 + it * 3 + 0] = clock64();
      // The G' rows of a k-block (8 rows per thread, dependent on s_info) are all in flight at once, and those of
      // k-block h+1 are requested before k-block h is converted: one exposed gather latency per tile instead of 2 KB1.
      auto gather = [&](int h, float4 (&dst)[8]) {
        const int c = h * 64 + q * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int src = __float_as_int(s_info[i * 16 + rsub].w);
          dst[i] = (a.gprime != nullptr && src >= 0) ? __ldg(reinterpret_cast<const float4*>(a.gprime + (size_t)src * D1 + c))
                                                     : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      float4 gcur[8], gnxt[8];
      gather(0, gcur);
#pragma unroll
      for (int h = 0; h < Cfg::KB1; ++h) {
        const int c = h * 64 + q * 4;                   // this thread's 4 channels of k-block h
        float wx[3][4], g0[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int d = 0; d < 3; ++d) wx[d][u] = HOIST ? wxr[h < HK ? h : 0][d][u] : __ldg(a.wx + d * D1 + c + u);
          g0[u] = HOIST ? g0r[h < HK ? h : 0][u] : (a.gprime ? 0.f : __ldg(a.shift1 + c + u));
        }
        if (h + 1 < Cfg::KB1) gather(h + 1, gnxt);
        tc::mbar_wait(&ha_free[sa], pa ^ 1);            // layer 2 has consumed the k-block that used this slot
        if (prec && h == 0) a.dbg[2000 + it * 3 + 1] = clock64();
        uint8_t* blk = HA + sa * SF_KBLOCK_BYTES;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 16 + rsub;
          const float4 info = s_info[r];
          const int src = __float_as_int(info.w);
          float y[4] = {gcur[i].x + g0[0], gcur[i].y + g0[1], gcur[i].z + g0[2], gcur[i].w + g0[3]};   // g0 = 0 with G'
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float t = fmaf(wx[0][u], info.x, y[u]);
            t = fmaf(wx[1][u], info.y, t);
            t = fmaf(wx[2][u], info.z, t);
            y[u] = src >= 0 ? fmaxf(t, 0.f) : 0.f;
          }
          uint2 hi, lo;
          tc::split_f16x2(y[0], y[1], hi.x, lo.x);
          tc::split_f16x2(y[2], y[3], hi.y, lo.y);
          const uint32_t off = tc::sw128_offset(r, q >> 1) + ((q & 1) << 3);
          *reinterpret_cast<uint2*>(blk + off) = hi;
          *reinterpret_cast<uint2*>(blk + 16384 + off) = lo;
        }
        tc::fence_proxy_async_smem();
        tc::mbar_arrive_warp(&ha_full[sa]);
        if (++sa == Cfg::NRA) { sa = 0; pa ^= 1; }
#pragma unroll
        for (int i = 0; i < 8; ++i) gcur[i] = gnxt[i];
      }
      if (prec) a.dbg[2000 + it * 3 + 2] = clock64();
      raw_cur = raw_nxt;
      idx_next = idx_nn;
    }
  } else if (warp == 16) {
    // =================================================================== weight loader
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint8_t* w2 = static_cast<const uint8_t*>(a.w2img);
      const uint8_t* w3 = static_cast<const uint8_t*>(a.w3img);
      auto load_layer = [&](int layer) {               // all weight items of one layer of one tile, in consumption order
        const int n_items = layer == 2 ? Cfg::KB1 * Cfg::ITEMS2 : Cfg::KB2 * Cfg::ITEMS3;
#pragma unroll 1
        for (int item = 0; item < n_items; ++item) {
          tc::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* dst = ring + stage * SF_STAGE_BYTES;
          if (layer == 2) sf_load_item<Cfg::NI2>(dst, w2, Cfg::KB1, item / Cfg::ITEMS2, item % Cfg::ITEMS2, &full_b[stage]);
          else sf_load_item<Cfg::NI3>(dst, w3, Cfg::KB2, item / Cfg::ITEMS3, item % Cfg::ITEMS3, &full_b[stage]);
          if (++stage == SF_NS) { stage = 0; phase ^= 1; }
        }
      };
      const int stride = gridDim.x;
      if (Cfg::DEEP) {                                  // MMA order: G2(0) | G2(t+1), G3(t) | ...
        if ((int)blockIdx.x < num_tiles) load_layer(2);
        for (int tile = blockIdx.x; tile < num_tiles; tile += stride) {
          if (tile + stride < num_tiles) load_layer(2);
          load_layer(3);
        }
      } else {
        for (int tile = blockIdx.x; tile < num_tiles; tile += stride) {
          load_layer(2);
          load_layer(3);
        }
      }
    }
  } else {
    // =================================================================== MMA issuer (whole warp, one elected lane issues)
    {
      constexpr uint32_t IDESC2 = tc::idesc_f16<false>(SF_TM, Cfg::NI2);
      constexpr uint32_t IDESC3 = tc::idesc_f16<false>(Cfg::NI3, SF_TM);      // layer 3 transposed: M = channels, N = rows
      int stage = 0;
      uint32_t phase = 0;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      const uint32_t ha_addr = tc::smem_u32(HA), hb_addr = tc::smem_u32(HB), ring_addr = tc::smem_u32(ring);
      const bool rec = a.dbg != nullptr && blockIdx.x == 0 && lane == 0;
      int nrec = 0;
      auto stamp = [&](int tag) { if (rec && nrec < 300) { a.dbg[2 * nrec] = tag; a.dbg[2 * nrec + 1] = clock64(); ++nrec; } };
      auto g2 = [&](int it) {
        const uint32_t p = Cfg::DEEP ? (uint32_t)(it & 1) : 0u;
        // ---- layer 2: D2 = H1 . W2'^T, k-block by k-block as the producers hand H1 over
        stamp(1);
#pragma unroll 1
        for (int kb = 0; kb < Cfg::KB1; ++kb) {
          tc::mbar_wait(&ha_full[sa], pa);
          tc::tc_fence_after();
          if (kb == 0) stamp(2);
#pragma unroll 1
          for (int ni = 0; ni < Cfg::ITEMS2; ++ni) {
            tc::mbar_wait(&full_b[stage], phase);
            tc::tc_fence_after();
            const uint64_t da_hi = tc::smem_desc_sw128(ha_addr + sa * SF_KBLOCK_BYTES);
            const uint64_t da_lo = tc::smem_desc_sw128(ha_addr + sa * SF_KBLOCK_BYTES + 16384);
            const uint64_t db_hi = tc::smem_desc_sw128(ring_addr + stage * SF_STAGE_BYTES);
            const uint64_t db_lo = tc::smem_desc_sw128(ring_addr + stage * SF_STAGE_BYTES + Cfg::NI2 * 128);
            const uint32_t d = tmem_base + Cfg::D2_COL + p * Cfg::D2_STRIDE + ni * Cfg::NI2;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = (uint64_t)(k * 2);
              tc::mma_f16_w_fill(d, da_hi + adv, db_hi + adv, IDESC2, (kb | k) != 0);   // A_hi stays in the collector ...
              tc::mma_f16_w_lastuse(d, da_hi + adv, db_lo + adv, IDESC2, 1);           // ... for the second MMA
              tc::mma_f16_w(d, da_lo + adv, db_hi + adv, IDESC2, 1);
            }
            tc::mma_commit_w(&empty[stage]);
            if (++stage == SF_NS) { stage = 0; phase ^= 1; }
          }
          tc::mma_commit_w(&ha_free[sa]);
          if (++sa == Cfg::NRA) { sa = 0; pa ^= 1; }
        }
        tc::mma_commit_w(&d2_full[p]);
        stamp(3);
      };
      auto g3 = [&](int it) {
        const uint32_t p = Cfg::DEEP ? (uint32_t)(it & 1) : 0u;
        // ---- layer 3: D3 = H2 . W3'^T, k-block by k-block as the epilogue hands H2 over
#pragma unroll 1
        for (int kb = 0; kb < Cfg::KB2; ++kb) {
          tc::mbar_wait(&hb_full[sb], pb);
          tc::tc_fence_after();
          if (kb == 0) stamp(4);
#pragma unroll 1
          for (int ni = 0; ni < Cfg::ITEMS3; ++ni) {
            tc::mbar_wait(&full_b[stage], phase);
            tc::tc_fence_after();
            // transposed: A = the weight item (M = 128 output channels), B = the H2 k-block (N = 128 pair-rows)
            const uint64_t da_hi = tc::smem_desc_sw128(ring_addr + stage * SF_STAGE_BYTES);
            const uint64_t da_lo = tc::smem_desc_sw128(ring_addr + stage * SF_STAGE_BYTES + Cfg::NI3 * 128);
            const uint64_t db_hi = tc::smem_desc_sw128(hb_addr + sb * SF_KBLOCK_BYTES);
            const uint64_t db_lo = tc::smem_desc_sw128(hb_addr + sb * SF_KBLOCK_BYTES + 16384);
            const uint32_t d = tmem_base + Cfg::D3_COL + p * Cfg::D3_STRIDE + ni * SF_TM;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint64_t adv = (uint64_t)(k * 2);
              tc::mma_f16_w_fill(d, da_hi + adv, db_hi + adv, IDESC3, (kb | k) != 0);   // A_hi stays in the collector ...
              tc::mma_f16_w_lastuse(d, da_hi + adv, db_lo + adv, IDESC3, 1);           // ... for the second MMA
              tc::mma_f16_w(d, da_lo + adv, db_hi + adv, IDESC3, 1);
            }
            tc::mma_commit_w(&empty[stage]);
            if (++stage == SF_NS) { stage = 0; phase ^= 1; }
          }
          tc::mma_commit_w(&hb_free[sb]);
          if (++sb == Cfg::NRB) { sb = 0; pb ^= 1; }
        }
        tc::mma_commit_w(&d3_full[p]);
        stamp(5);
      };
      const int stride = gridDim.x;
      if (Cfg::DEEP) {
        if ((int)blockIdx.x < num_tiles) g2(0);
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += stride, ++it) {
          if (tile + stride < num_tiles) g2(it + 1);
          g3(it);
        }
      } else {
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += stride, ++it) {
          g2(it);
          g3(it);
        }
      }
      if (rec) a.dbg[2 * nrec] = -1;
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
  static bool configured[PTT_MAX_DEVICES] = {};
  const int dev = ptt_current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    if (int rc = tc::tc_bind_fault(ptt_fault_word())) return rc;
    configured[dev] = true;
  }
  const long long tiles = (a.rows + SF_TM - 1) / SF_TM;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = (int)llmin_(tiles, sms);
  kern<<<grid, SF_THREADS, Cfg::SMEM_BYTES, st>>>(a, (int)tiles); PTT_LAUNCHED();
  return ptt_launch_status();
}

}  // namespace

long long* g_sa_dbg = nullptr;   // tuning hook: timeline buffer picked up by the next launches (ptt_debug_sa_timeline)
extern "C" __attribute__((visibility("default"))) void ptt_debug_sa_timeline(long long* buf) { g_sa_dbg = buf; }

bool sa_fused_supported(int d1, int d2, int d3, int ns) {
  const bool dims = (d1 == 64 && d2 == 64 && d3 == 128) || (d1 == 128 && d2 == 128 && d3 == 256) ||
                    (d1 == 256 && d2 == 256 && d3 == 256);
  const bool group = ns >= 1 && ns <= SF_TM && (ns & (ns - 1)) == 0;   // whole groups per 128-row tile
  return dims && group;
}

int sa_fused_launch(const SaFusedArgs& a_in, int d1, int d2, int d3, cudaStream_t st) {
  SaFusedArgs a = a_in;
  a.dbg = g_sa_dbg;
  if (a.rows <= 0) return PTT_OK;
  if (a.ns > 32) {   // groups wider than a warp are combined with atomicMax: the output starts at +0
    cudaError_t e = cudaMemset2DAsync(a.out_pm, (size_t)a.ld_out * sizeof(float), 0, (size_t)d3 * sizeof(float),
                                      (size_t)(a.rows / a.ns), st);
    if (e != cudaSuccess) return (int)e;
  }
  if (d1 == 64 && d2 == 64 && d3 == 128) return sf_launch<64, 64, 128>(a, st);
  if (d1 == 128 && d2 == 128 && d3 == 256) return sf_launch<128, 128, 256>(a, st);
  if (d1 == 256 && d2 == 256 && d3 == 256) return sf_launch<256, 256, 256>(a, st);
  return PTT_ERR_UNSUPPORTED;
}
