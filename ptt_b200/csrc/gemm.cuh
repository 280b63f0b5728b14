// Internal (not exported) dense row-block contraction used by the SA-MLP and transformer paths:
//   y[r, 0:N] = act( scale[c] * (x[row(r), 0:K] . wt[0:K, c]) + shift[c] ) (+ residual[r, c])
// x is point-major (one row per point / per (centre, sample) pair), wt is the packed TRANSPOSED
// weight (K rows of ldw floats, ldw % 4 == 0, zero padded).
#pragma once

#include "common.cuh"

struct PttGemmArgs {
  const float* x = nullptr;
  int ldx = 0;
  const int* a_rows = nullptr;  // optional: output row r reads x row a_rows[r]
  // optional per-K-channel affine + ReLU applied to the A operand while it is staged: x <- relu(a_ka[k] * x + a_kb[k]).
  // The training path feeds a layer with the PRE-BatchNorm output of the previous one this way (tcgen05 path only).
  const float* a_ka = nullptr;
  const float* a_kb = nullptr;
  int R = 0, K = 0;
  const float* wt = nullptr;   // fp32 transposed weight (K rows of ldw floats): the CUDA-core path
  int ldw = 0;
  const void* wimg = nullptr;  // fp16 hi/lo UMMA image of the same weight (ptt_tc_pack_weight): the tcgen05 path
  int N = 0;
  const float* scale = nullptr;  // nullptr = 1
  const float* shift = nullptr;  // nullptr = 0   (bias or folded BatchNorm shift)
  int relu = 0;
  const float* residual = nullptr;
  int ldr = 0;
  float* y = nullptr;
  int ldy = 0;
  // batched form (tcgen05 path only): `batch` independent problems; x / y / residual advance by their strides (floats)
  // and the weight image by wimg_bstride (bytes) per batch element
  int batch = 1;
  long long x_bstride = 0, y_bstride = 0, res_bstride = 0;
  size_t wimg_bstride = 0;
};

// Runs on the tcgen05 path when `wimg` is set and the A operand qualifies (16-byte aligned rows), else on the
// CUDA-core path.  Both produce fp32-class results (tests compare them against each other on the device).
int ptt_gemm_launch(const PttGemmArgs& a, cudaStream_t st);
int ptt_gemm_launch_ffma(const PttGemmArgs& a, cudaStream_t st);

// tcgen05 path (tc_gemm.cu)
size_t ptt_tc_weight_halves(int K, int Cout);   // size of the fp16 image, in 2-byte units (multiple of 8192)
// src(c, k) = w[c * ld_c + k * ld_k]: (Cout, K) row-major weight -> ld_c = K, ld_k = 1; transposed (K, ldw) image -> ld_c = 1, ld_k = ldw
// row_scale (optional, Cout floats): the image holds diag(row_scale) . W
int ptt_tc_pack_weight(const float* w, long long ld_c, long long ld_k, int Cout, int K, void* img, cudaStream_t st,
                       const float* row_scale = nullptr, int batch = 1, long long w_bstride = 0, size_t img_bstride = 0);
// nn.Linear image of ptt_linear_pack (transposed fp32 weight, bias row, fp16 image) from a strided source, one launch
int ptt_linear_pack_all(const float* w, long long ld_c, long long ld_k, const float* bias, int K, int Cout, float* params,
                        cudaStream_t st);
// the same for `count` layers described by a DEVICE table (include/ptt_b200.h: PttPackDesc), one launch
int ptt_linear_pack_batch_launch(const PttPackDesc* descs_device, int count, cudaStream_t st);
bool ptt_tc_gemm_supported(const PttGemmArgs& a);
int ptt_tc_gemm_launch(const PttGemmArgs& a, const void* wimg, cudaStream_t st);
// floats occupied by the tcgen05 image of a (Cout, K) weight
static inline size_t ptt_tc_weight_floats(int K, int Cout) { return ptt_tc_weight_halves(K, Cout) / 2; }

// Weight gradient (tc_wgrad.cu): dW (M, ldw)[:, 0:N] += dY (R, ldy)[:, 0:M]^T . f(X (R, ldx)[:, 0:N]), f = identity or
// relu(ka[n] * x + kb[n]); rows are read as float4 (ldy, ldx multiples of 4, 16-byte aligned bases).  dbias (optional, M
// floats) += column sums of dY: the bias gradient of the same layer, taken from the rows while they are staged.
int ptt_tc_wgrad_launch(const float* dy, int ldy, const float* x, int ldx, const float* x_ka, const float* x_kb,
                        long long R, int M, int N, float* dw, int ldw, float* dbias, cudaStream_t st);

// Weight-stationary persistent contraction with fused column statistics (ws_gemm.cu): y = f(x) . W^T (no bias), optional
// sums (2, N) double <- column sums of y and y^2.  wimg = the ptt_tc_pack_weight image of W (N, K).
bool ptt_ws_gemm_supported(const float* x, int ldx, long long R, int K, int N);
int ptt_ws_gemm_launch(const float* x, int ldx, long long R, int K, const float* ka, const float* kb, const void* wimg, int N,
                       float* y, int ldy, double* sums, cudaStream_t st);

// nn.Linear (Cout,K) weight [+bias] -> transposed image (K+1 rows x ldw, last row = bias), zero padded
int ptt_linear_pack_launch(const float* weight, const float* bias, int K, int Cout, float* params, cudaStream_t st);
// same, into columns [col0, col0+Cout) of an image whose rows are ldw wide (caller zero-fills the image)
int ptt_linear_pack_cols(const float* weight, const float* bias, int K, int Cout, int ldw, int col0, float* params,
                         cudaStream_t st);

// packed linear image: wt (K x ldw) followed by bias (ldw); ldw = round_up(Cout, 4)
static inline int ptt_linear_ldw(int cout) { return round_up(cout, 4); }
