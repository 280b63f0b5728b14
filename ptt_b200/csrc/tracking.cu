// Per-frame pre / post-processing of the tracking loop on the device, for T independent
// tracklets at once.  The reference does this on the host, per tracklet and per frame, in numpy
// (tools/eval_utils/eval_tracking_utils.py:140-274, ptt/datasets/kitti/kitti_tracking_utils.py:192-367) with a
// host <-> device round trip around every model call; here the crop, the resampling and the box update are three small
// kernels in the same CUDA graph as the model, so the box a frame produces feeds the next frame's crop without leaving
// the GPU.
//
//   track_crop_kernel        crop_center_pc (:300-340) / get_model (:219-237): two-stage axis-aligned crop around a box
//                            (world frame, then box frame), ORDER-PRESERVING compaction (block prefix sums), points
//                            written in the box frame; up to two sources are concatenated (template = first + previous)
//   track_regularize_kernel  regularize_pc(istrain=False) (:342-367): np.random.seed(1) + np.random.randint(0, n, size)
//                            reproduced from the MT19937 stream of seed 1 (masked rejection of 32-bit outputs: the
//                            j-th index is the j-th accepted output, found with a block prefix sum over the stream)
//   track_update_kernel      get_box_by_offset (:192-216) incl. its np.random.uniform(-1, 1) clamps
//
// Arithmetic is the reference's, restated in oracle/tracking_ref.py: float32 clouds, float64 box state, every in-place
// numpy assignment rounds to float32 once, comparisons are exact float64 comparisons, 3-term dot products are evaluated
// left to right without FMA (intrinsics below).  Box state per tracklet = 15 doubles: center | R row-major | wlh.
#include "common.cuh"

namespace {

constexpr int TK_THREADS = 256;
constexpr int TK_BOX = 15;

__device__ __forceinline__ double dot3(const double* r, double x, double y, double z) {
  return __dadd_rn(__dadd_rn(__dmul_rn(r[0], x), __dmul_rn(r[1], y)), __dmul_rn(r[2], z));
}

// inclusive-exclusive block scan of one flag per thread: returns this thread's rank among the set flags of the block and
// the block total (every thread gets it); s_w: TK_THREADS / 32 ints of shared memory
__device__ __forceinline__ int block_rank(bool flag, int* s_w, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned b = __ballot_sync(0xffffffffu, flag);
  const int in_warp = __popc(b & ((1u << lane) - 1u));
  __syncthreads();                       // s_w of the previous call has been read by everyone
  if (lane == 0) s_w[warp] = __popc(b);
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < TK_THREADS / 32; ++w) {
    const int c = s_w[w];
    base += w < warp ? c : 0;
    tot += c;
  }
  total = tot;
  return base + in_warp;
}

struct CropSrc {
  const float* pts;      // (T, cap, 3)
  const int* cnt;        // (T)
  const double* box;     // (T, 15); unused when precropped
  int cap;
  int precropped;        // 1: the points are already cropped and in the box frame (the cached first-frame part of the template)
};

struct CropArgs {
  CropSrc src[2];
  int nsrc;
  double offset, scale;
  int search;            // 1: second-stage margin offset + 0.6 * wlh[1] (a gt box exists, :323); 0: offset (:333)
  float* out;            // (T, cap_out, 3)
  int* out_cnt;          // (T)
  int cap_out;
};

__global__ void __launch_bounds__(TK_THREADS) track_crop_kernel(const __grid_constant__ CropArgs a) {
  __shared__ double s_b[21];             // maxi[3] mini[3] maxi2[3] mini2[3] Rt[9]
  __shared__ double s_trans[3];
  __shared__ int s_w[TK_THREADS / 32];
  const int t = blockIdx.x, tid = threadIdx.x;
  float* out = a.out + (size_t)t * a.cap_out * 3;
  int written = 0;
  for (int s = 0; s < a.nsrc; ++s) {
    const CropSrc& S = a.src[s];
    const float* P = S.pts + (size_t)t * S.cap * 3;
    const int n = min(S.cnt[t], S.cap);
    if (S.precropped) {
      for (int e = tid; e < n * 3; e += TK_THREADS) out[(size_t)written * 3 + e] = P[e];
      written += n;
      continue;
    }
    __syncthreads();
    if (tid == 0) {
      const double* B = S.box + (size_t)t * TK_BOX;
      const double *c = B, *R = B + 3, *wlh = B + 12;
      // Box.corners() (:170-189) of the box with wlh * (4 * scale), +- 2 * offset
      const double sc1 = 4 * a.scale, off1 = 2 * a.offset;
      const double w = wlh[0] * sc1, l = wlh[1] * sc1, h = wlh[2] * sc1;
      const double sx[8] = {1, 1, 1, 1, -1, -1, -1, -1}, sy[8] = {1, -1, -1, 1, 1, -1, -1, 1}, sz[8] = {1, 1, -1, -1, 1, 1, -1, -1};
      for (int i = 0; i < 3; ++i) {
        double mx = -INFINITY, mn = INFINITY;
        for (int k = 0; k < 8; ++k) {
          const double v = __dadd_rn(dot3(R + 3 * i, (l / 2) * sx[k], (w / 2) * sy[k], (h / 2) * sz[k]), c[i]);
          mx = fmax(mx, v);
          mn = fmin(mn, v);
        }
        s_b[i] = mx + off1;
        s_b[3 + i] = mn - off1;
        s_trans[i] = -c[i];
      }
      // second stage: the box at the origin with identity orientation, wlh * scale, +- the margin
      const double off2 = a.search ? __dadd_rn(a.offset, __dmul_rn(wlh[1], 0.6)) : a.offset;
      const double half[3] = {(wlh[1] * a.scale) / 2, (wlh[0] * a.scale) / 2, (wlh[2] * a.scale) / 2};
      for (int i = 0; i < 3; ++i) {
        s_b[6 + i] = half[i] + off2;
        s_b[9 + i] = -half[i] - off2;
      }
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) s_b[12 + 3 * i + j] = R[3 * j + i];      // R^T
    }
    __syncthreads();
    for (int k0 = 0; k0 < n; k0 += TK_THREADS) {
      const int k = k0 + tid;
      bool keep = false;
      float r[3] = {0.f, 0.f, 0.f};
      if (k < n) {
        const float p[3] = {P[3 * k], P[3 * k + 1], P[3 * k + 2]};
        keep = true;
#pragma unroll
        for (int i = 0; i < 3; ++i) keep = keep && (double)p[i] > s_b[3 + i] && (double)p[i] < s_b[i];
        float q[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) q[i] = (float)__dadd_rn((double)p[i], s_trans[i]);               // PointCloud.translate
#pragma unroll
        for (int i = 0; i < 3; ++i) r[i] = (float)dot3(s_b + 12 + 3 * i, (double)q[0], (double)q[1], (double)q[2]);   // .rotate
#pragma unroll
        for (int i = 0; i < 3; ++i) keep = keep && (double)r[i] > s_b[9 + i] && (double)r[i] < s_b[6 + i];
      }
      int total;
      const int rank = block_rank(keep, s_w, total);
      if (keep && written + rank < a.cap_out) {
        float* o = out + (size_t)(written + rank) * 3;
        o[0] = r[0]; o[1] = r[1]; o[2] = r[2];
      }
      written += total;
    }
  }
  if (tid == 0) a.out_cnt[t] = min(written, a.cap_out);
}

// pts (T, cap, 3) with cnt (T) valid rows -> out (T, size, 3); mt_pos (T) = position in the seed-1 MT19937 stream
__global__ void __launch_bounds__(TK_THREADS) track_regularize_kernel(const float* __restrict__ pts, const int* __restrict__ cnt,
                                                                       int cap, int size, const unsigned* __restrict__ mt,
                                                                       int mt_len, int* __restrict__ mt_pos,
                                                                       float* __restrict__ out) {
  __shared__ int s_w[TK_THREADS / 32];
  const int t = blockIdx.x, tid = threadIdx.x;
  const float* P = pts + (size_t)t * cap * 3;
  float* O = out + (size_t)t * size * 3;
  const int n = min(cnt[t], cap);
  if (n <= 2) {                                               // <= 2 points survived the crop: zeros (:359-360)
    for (int e = tid; e < size * 3; e += TK_THREADS) O[e] = 0.f;
    return;
  }
  if (n == size) {                                            // no resampling, no reseed (:349)
    for (int e = tid; e < size * 3; e += TK_THREADS) O[e] = P[e];
    return;
  }
  const unsigned rng = (unsigned)(n - 1);
  unsigned mask = rng;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  int got = 0;
  for (int p0 = 0; p0 < mt_len && got < size; p0 += TK_THREADS) {
    const int p = p0 + tid;
    unsigned v = 0;
    bool acc = false;
    if (p < mt_len) {
      v = mt[p] & mask;
      acc = v <= rng;
    }
    int total;
    const int rank = block_rank(acc, s_w, total);
    const int j = got + rank;
    if (acc && j < size) {
      const float* src = P + (size_t)v * 3;
      float* o = O + (size_t)j * 3;
      o[0] = src[0]; o[1] = src[1]; o[2] = src[2];
      if (j == size - 1) mt_pos[t] = p + 1;                   // the stream position after the last accepted output
    }
    got += total;
  }
}

__global__ void track_update_kernel(const float* __restrict__ best, int ldb, double* __restrict__ state, int T, int use_z,
                                    const unsigned* __restrict__ mt, int mt_len, int* __restrict__ mt_pos,
                                    double* __restrict__ results, int max_frames, int* __restrict__ frame_idx) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < T) {
    double* B = state + (size_t)t * TK_BOX;
    const float* e = best + (size_t)t * ldb;
    float o0 = e[0], o1 = e[1];
    const float o2 = e[2];
    int pos = mt_pos[t];
    auto uniform = [&]() -> float {                           // np.random.uniform(-1, 1), stored into a float32 row
      const unsigned a = mt[min(pos, mt_len - 2)] >> 5, b = mt[min(pos + 1, mt_len - 1)] >> 6;
      pos += 2;
      return (float)(-1.0 + 2.0 * (((double)a * 67108864.0 + (double)b) / 9007199254740992.0));
    };
    if ((double)o0 > B[12]) o0 = uniform();                   // offset[0] > wlh[0]                     (:208-209)
    if ((double)o1 > fmin(B[13], 2.0)) o1 = uniform();        // offset[1] > min(wlh[1], 2)             (:210-211)
    mt_pos[t] = pos;
    const double oz = use_z ? (double)o2 : 0.0;
    // offset[-1] * np.pi / 180 in float32 (NumPy >= 2: Python floats are weak)
    const double th = (double)__fdiv_rn(__fmul_rn(e[3], 3.14159274101257324f), 180.f);
    const double c = cos(th), s = sin(th);
    double R[9];
    for (int i = 0; i < 9; ++i) R[i] = B[3 + i];
    for (int i = 0; i < 3; ++i) {
      B[i] = __dadd_rn(B[i], dot3(R + 3 * i, (double)o0, (double)o1, oz));
      B[3 + 3 * i + 0] = R[3 * i] * c + R[3 * i + 1] * s;      // R . Rz(theta)
      B[3 + 3 * i + 1] = R[3 * i + 1] * c - R[3 * i] * s;
    }
    const int f = min(*frame_idx, max_frames - 1);
    double* out = results + ((size_t)f * T + t) * TK_BOX;
    for (int i = 0; i < TK_BOX; ++i) out[i] = B[i];
  }
}

// the frame counter is bumped by its own single-thread launch: no ordering exists between the blocks of the update grid
__global__ void track_bump_kernel(int* frame_idx, int max_frames) {
  if (threadIdx.x == 0 && blockIdx.x == 0 && *frame_idx < max_frames) *frame_idx += 1;
}

}  // namespace

// First n 32-bit outputs of MT19937 seeded like np.random.seed(seed) (init_genrand), into HOST memory.
extern "C" int ptt_mt19937_stream(unsigned seed, int n, unsigned* h_out) {
  PTT_CHECK_ARG(n >= 0 && (n == 0 || h_out != nullptr));
  static const int NN = 624, MM = 397;
  unsigned mt[NN];
  mt[0] = seed;
  for (int i = 1; i < NN; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (unsigned)i;
  int idx = NN;
  for (int k = 0; k < n; ++k) {
    if (idx >= NN) {
      for (int i = 0; i < NN; ++i) {
        const unsigned y = (mt[i] & 0x80000000u) | (mt[(i + 1) % NN] & 0x7fffffffu);
        mt[i] = mt[(i + MM) % NN] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    unsigned y = mt[idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    h_out[k] = y;
  }
  return PTT_OK;
}

extern "C" int ptt_track_crop(int T, int n_sources, const float* const* h_points, const int* const* h_counts,
                              const double* const* h_boxes, const int* h_caps, const int* h_precropped, double offset,
                              double scale, int search, float* out, int cap_out, int* out_counts, ptt_stream_t stream) {
  PTT_CHECK_ARG(T >= 0 && n_sources >= 1 && n_sources <= 2 && h_points && h_counts && h_boxes && h_caps && h_precropped);
  if (T == 0) return PTT_OK;
  PTT_CHECK_ARG(out && out_counts && cap_out >= 1);
  CropArgs a = {};
  a.nsrc = n_sources;
  for (int s = 0; s < n_sources; ++s) {
    PTT_CHECK_ARG(h_points[s] && h_counts[s] && h_caps[s] >= 1 && (h_precropped[s] || h_boxes[s]));
    a.src[s].pts = h_points[s]; a.src[s].cnt = h_counts[s]; a.src[s].box = h_boxes[s];
    a.src[s].cap = h_caps[s]; a.src[s].precropped = h_precropped[s];
  }
  a.offset = offset; a.scale = scale; a.search = search;
  a.out = out; a.out_cnt = out_counts; a.cap_out = cap_out;
  track_crop_kernel<<<T, TK_THREADS, 0, as_stream(stream)>>>(a); PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_track_regularize(const float* points, const int* counts, int T, int cap, int size,
                                    const unsigned* mt_stream, int mt_len, int* mt_pos, float* out, ptt_stream_t stream) {
  PTT_CHECK_ARG(T >= 0 && cap >= 1 && size >= 1 && mt_len >= 4 * size);
  if (T == 0) return PTT_OK;
  PTT_CHECK_ARG(points && counts && mt_stream && mt_pos && out);
  track_regularize_kernel<<<T, TK_THREADS, 0, as_stream(stream)>>>(points, counts, cap, size, mt_stream, mt_len, mt_pos, out);
  PTT_LAUNCHED();
  return ptt_launch_status();
}

extern "C" int ptt_track_update(const float* best_box, int ld_best, double* box_state, int T, int use_z,
                                const unsigned* mt_stream, int mt_len, int* mt_pos, double* results, int max_frames,
                                int* frame_idx, ptt_stream_t stream) {
  PTT_CHECK_ARG(T >= 0 && ld_best >= 4 && max_frames >= 1 && mt_len >= 8);
  if (T == 0) return PTT_OK;
  PTT_CHECK_ARG(best_box && box_state && mt_stream && mt_pos && results && frame_idx);
  cudaStream_t st = as_stream(stream);
  track_update_kernel<<<ceil_div(T, 128), 128, 0, st>>>(best_box, ld_best, box_state, T, use_z, mt_stream, mt_len, mt_pos,
                                                        results, max_frames, frame_idx); PTT_LAUNCHED();
  track_bump_kernel<<<1, 32, 0, st>>>(frame_idx, max_frames); PTT_LAUNCHED();
  return ptt_launch_status();
}
