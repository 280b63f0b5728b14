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
  int R = 0, K = 0;
  const float* wt = nullptr;
  int ldw = 0;
  int N = 0;
  const float* scale = nullptr;  // nullptr = 1
  const float* shift = nullptr;  // nullptr = 0   (bias or folded BatchNorm shift)
  int relu = 0;
  const float* residual = nullptr;
  int ldr = 0;
  float* y = nullptr;
  int ldy = 0;
};

int ptt_gemm_launch(const PttGemmArgs& a, cudaStream_t st);

// nn.Linear (Cout,K) weight [+bias] -> transposed image (K+1 rows x ldw, last row = bias), zero padded
int ptt_linear_pack_launch(const float* weight, const float* bias, int K, int Cout, float* params, cudaStream_t st);
// same, into columns [col0, col0+Cout) of an image whose rows are ldw wide (caller zero-fills the image)
int ptt_linear_pack_cols(const float* weight, const float* bias, int K, int Cout, int ldw, int col0, float* params,
                         cudaStream_t st);

// packed linear image: wt (K x ldw) followed by bias (ldw); ldw = round_up(Cout, 4)
static inline int ptt_linear_ldw(int cout) { return round_up(cout, 4); }
