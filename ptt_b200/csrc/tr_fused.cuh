// Internal interface of the transformer pair-row passes (tr_fused.cu); see that file for the design.
#pragma once

#include "common.cuh"

enum TrProducer { TR_PROD_PLAIN = 0, TR_PROD_DELTA0 = 1, TR_PROD_QKPOS = 2 };
// STORE_QK: out = relu(acc + bias + QG[token] - KG[neighbour])  (the q/k part of fc_gamma.0, pre-multiplied per token)
enum TrEpilogue { TR_EPI_STORE = 0, TR_EPI_SOFTMAX = 1, TR_EPI_STORE_QK = 2 };

struct TrPassArgs {
  int n = 0, k = 0, dm = 0;      // tokens per cloud, neighbours per token, d_model
  long long pairs = 0;           // B * n * k rows
  const float* xyz = nullptr;    // (B, n, 3)
  const int* knn = nullptr;      // (B, n, k) neighbour index inside the cloud
  // producer sources
  const float* a_src = nullptr;  // PLAIN: (pairs, lda) fp32 rows
  int lda = 0;
  const float* wd0 = nullptr;    // DELTA0: fc_delta.0 image, rows 0..2 = weight^T (3, ldw0), row 3 = bias
  int ldw0 = 0;
  const float* qkv = nullptr;    // QKPOS producer / SOFTMAX epilogue: (B*n, ldq) with q at column 0, k at koff, v at voff
  int ldq = 0, koff = 0, voff = 0;
  const float* pos = nullptr;    // (pairs, dm): QKPOS producer input, SOFTMAX epilogue input
  // contraction
  const void* wimg = nullptr;    // tcgen05 image of the (dm, dm) weight
  const float* bias = nullptr;   // (dm)
  int relu = 0;
  int qk_mode = 0;               // STORE_QK epilogue: 0 = + qkv[token, 0:] - qkv[nbr, koff:] ; 1 = + qkv[nbr, koff:] only
  int pos_has_v = 0;             // SOFTMAX epilogue: `pos` already holds pos + v[nbr] (pass 1 ran with qk_mode 1)
  // outputs
  float* out = nullptr;          // STORE: (pairs, ldo);  SOFTMAX: res (B*n, ldo)
  int ldo = 0;
  float* attn = nullptr;         // SOFTMAX, optional: (pairs, dm) softmax weights
  float divisor = 1.f;           // sqrt(d_model)
  long long* dbg = nullptr;      // optional timeline buffer (clock64 stamps of CTA 0; tuning only)
  int dbg_flags = 0;             // tuning only: 1 = skip the epilogue's global stores, 2 = loader re-reads one weight item
  const float* x_sub = nullptr;  // SOFTMAX, Offset variant: res = x_sub - res   (B*n, ldx)
  int ldx = 0;
  // STORE_QK, optional rank-1 term: out += row_scalar[pair] * col_vec[channel]  (before the ReLU) -- the similarity
  // column of TransformerBlockCosine's fc_sim pushed through fc_gamma.0
  const float* row_scalar = nullptr;   // (pairs)
  const float* col_vec = nullptr;      // (dm)
  // STORE (transposed accumulators), optional: out <- out * [mask_ref > 0], mask_ref (pairs, dm) -- the ReLU backward of
  // an input-gradient contraction against the stored activation (training path)
  const float* mask_ref = nullptr;
};

bool tr_fused_supported(int n, int k, int dm);
int tr_fused_launch(const TrPassArgs& a, int producer, int epilogue, cudaStream_t st);
