// Internal interface of the fused set-abstraction kernel (sa_fused.cu); see that file for the design.
#pragma once

#include "common.cuh"

struct SaFusedArgs {
  const float* xyz = nullptr;      // (B, N, 3)
  const float* new_xyz = nullptr;  // (B, M, 3)
  const int* idx = nullptr;        // (B, M, ns) ball-query result; nullptr = every centre's group is all N points in order (ns == N)
  const float* pair_scalar = nullptr;  // optional (B, M, ns): the pair's relative part is (pair_scalar, 0, 0) instead of
                                       // (xyz[i] - new_xyz[j]) (/ radius) -- the similarity column of CosineSimAug
  const float* gprime = nullptr;   // (B*N, D1) = scale1 * (feats . W1f^T) + shift1, or nullptr when the layer has no features
  const float* shift1 = nullptr;   // (D1) used when gprime == nullptr
  const float* wx = nullptr;       // (3, D1): scale1 * W1[:, xyz columns]
  const void* w2img = nullptr;     // tcgen05 images of diag(scale2).W2 and diag(scale3).W3
  const void* w3img = nullptr;
  const float* shift2 = nullptr;
  const float* shift3 = nullptr;
  int B = 0, N = 0, M = 0, ns = 0;
  float radius = 1.f;
  int normalize = 0;
  float* out_pm = nullptr;         // (B*M, ld_out), D3 valid columns
  int ld_out = 0;
  long long rows = 0;              // B * M * ns
  long long* dbg = nullptr;        // optional clock64 timeline of CTA 0 (tuning only)
};

bool sa_fused_supported(int d1, int d2, int d3, int ns);
int sa_fused_launch(const SaFusedArgs& a, int d1, int d2, int d3, cudaStream_t st);
