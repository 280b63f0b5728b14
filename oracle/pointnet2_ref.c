/*
 * oracle/pointnet2_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the native ops the reference calls through
 * `pointnet2_ops._ext` (reference: ptt/models/backbones_3d/pointnet2/pointnet2_utils.py:24,
 * call sites :48,:78,:112,:118,:145,:182,:204,:237,:257,:287).
 *
 * PARITY UNPINNED: the arithmetic lives in the third-party, un-vendored, un-pinned pip
 * dependency `pointnet2_ops` (erikwijmans/Pointnet2_PyTorch, pointnet2_ops_lib, v3.0.0,
 * requirements.txt:3) which is absent from /root/reference and has no CPU path; the
 * reference holds no tests or golden vectors for it.  What is restated here is the
 * PUBLISHED algorithm of that package with these recorded decisions:
 *
 *   (i)   distance arithmetic = what nvcc's default -fmad=true emits for the upstream
 *         expression `a*a + b*b + c*c`, observed with nvcc 12.9 for sm_75 and sm_100a
 *         (oracle/probe_contraction.sh):   fmaf(c,c, fmaf(a,a, b*b)).
 *   (ii)  FPS skips points with (double)(x*x+y*y+z*z) <= 1e-3 and never updates them.
 *   (iii) FPS arg-max: per-thread strict '>' in index order, then the shared-memory tree
 *         reduction of a block of bs = min(512, 2^floor(log2 N)) threads where ties keep
 *         the lower slot.  The tree is SIMULATED literally here (no closed form), so the
 *         CUDA kernel's closed-form tie-break key is checked against it.
 *   (iv)  ball query: strict d2 < radius*radius (fp32), index-ordered first-nsample,
 *         first hit pre-fills all slots, empty row stays 0.
 *   (v)   int32 outputs.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  Build: `make -C oracle` (gcc -O2 -ffp-contract=off).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

static int opt_n_threads(int work_size) {
  /* upstream cuda_utils.h: clamp(2^floor(log2 w), 1, 512) */
  int p = 1;
  while ((p << 1) <= work_size && (p << 1) <= 512) p <<= 1;
  return p < 1 ? 1 : p;
}

/* a*a + b*b + c*c as nvcc contracts it: mul on the middle term, then two FMAs */
static inline float sq3(float a, float b, float c) {
  return fmaf(c, c, fmaf(a, a, b * b));
}

/* shared tree reduction of the upstream FPS kernels: slot t absorbs slot t+s when strictly larger */
static void tree_argmax(float* v, int* vi, int bs) {
  for (int s = bs / 2; s >= 1; s >>= 1) {
    for (int t = 0; t < s; ++t) {
      float v1 = v[t], v2 = v[t + s];
      int i1 = vi[t], i2 = vi[t + s];
      v[t] = v1 > v2 ? v1 : v2; /* max(v1, v2) */
      vi[t] = v2 > v1 ? i2 : i1;
    }
  }
}

/* _ext.furthest_point_sampling(xyz[B,N,3], npoint) -> idx[B,npoint]   (pointnet2_utils.py:78) */
ORACLE_API void ref_furthest_point_sampling(int B, int N, int M, const float* xyz, int* idx) {
  if (M <= 0) return;
  const int bs = opt_n_threads(N);
  float* temp = (float*)malloc(sizeof(float) * (size_t)N);
  float* sv = (float*)malloc(sizeof(float) * (size_t)bs);
  int* si = (int*)malloc(sizeof(int) * (size_t)bs);
  for (int b = 0; b < B; ++b) {
    const float* P = xyz + (size_t)b * N * 3;
    int* out = idx + (size_t)b * M;
    for (int k = 0; k < N; ++k) temp[k] = 1e10f;
    int old = 0;
    out[0] = 0;
    for (int j = 1; j < M; ++j) {
      const float x1 = P[old * 3 + 0], y1 = P[old * 3 + 1], z1 = P[old * 3 + 2];
      for (int t = 0; t < bs; ++t) {
        float best = -1.0f;
        int besti = 0;
        for (int k = t; k < N; k += bs) {
          const float x2 = P[k * 3 + 0], y2 = P[k * 3 + 1], z2 = P[k * 3 + 2];
          const float mag = sq3(x2, y2, z2);
          if ((double)mag <= 1e-3) continue;
          const float d = sq3(x2 - x1, y2 - y1, z2 - z1);
          const float d2 = fminf(d, temp[k]);
          temp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        sv[t] = best;
        si[t] = besti;
      }
      tree_argmax(sv, si, bs);
      old = si[0];
      out[j] = old;
    }
  }
  free(temp);
  free(sv);
  free(si);
}

/* _ext.furthest_point_sampling_with_dist(dist[B,N,N], npoint)  (pointnet2_utils.py:48; dead in the
 * reference -- 'ffps' -- and not exported by the pinned upstream; semantics = same loop on a
 * precomputed distance row, no origin skip). */
ORACLE_API void ref_furthest_point_sampling_with_dist(int B, int N, int M, const float* dist, int* idx) {
  if (M <= 0) return;
  const int bs = opt_n_threads(N);
  float* temp = (float*)malloc(sizeof(float) * (size_t)N);
  float* sv = (float*)malloc(sizeof(float) * (size_t)bs);
  int* si = (int*)malloc(sizeof(int) * (size_t)bs);
  for (int b = 0; b < B; ++b) {
    const float* D = dist + (size_t)b * N * N;
    int* out = idx + (size_t)b * M;
    for (int k = 0; k < N; ++k) temp[k] = 1e10f;
    int old = 0;
    out[0] = 0;
    for (int j = 1; j < M; ++j) {
      for (int t = 0; t < bs; ++t) {
        float best = -1.0f;
        int besti = 0;
        for (int k = t; k < N; k += bs) {
          const float d2 = fminf(D[(size_t)old * N + k], temp[k]);
          temp[k] = d2;
          besti = d2 > best ? k : besti;
          best = d2 > best ? d2 : best;
        }
        sv[t] = best;
        si[t] = besti;
      }
      tree_argmax(sv, si, bs);
      old = si[0];
      out[j] = old;
    }
  }
  free(temp);
  free(sv);
  free(si);
}

/* _ext.gather_points(points[B,C,N], idx[B,M]) -> out[B,C,M]   (pointnet2_utils.py:112) */
ORACLE_API void ref_gather_points(int B, int C, int N, int M, const float* points, const int* idx, float* out) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < M; ++j)
        out[((size_t)b * C + c) * M + j] = points[((size_t)b * C + c) * N + idx[(size_t)b * M + j]];
}

/* _ext.gather_points_grad(grad_out[B,C,M], idx[B,M], N) -> grad_points[B,C,N]   (pointnet2_utils.py:118) */
ORACLE_API void ref_gather_points_grad(int B, int C, int N, int M, const float* grad_out, const int* idx,
                                       float* grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)B * C * N);
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < M; ++j)
        grad_points[((size_t)b * C + c) * N + idx[(size_t)b * M + j]] += grad_out[((size_t)b * C + c) * M + j];
}

/* _ext.ball_query(new_xyz[B,M,3], xyz[B,N,3], radius, nsample) -> idx[B,M,nsample]   (pointnet2_utils.py:287) */
ORACLE_API void ref_ball_query(int B, int N, int M, float radius, int nsample, const float* new_xyz,
                               const float* xyz, int* idx) {
  const float radius2 = radius * radius;
  memset(idx, 0, sizeof(int) * (size_t)B * M * nsample);
  for (int b = 0; b < B; ++b) {
    const float* P = xyz + (size_t)b * N * 3;
    const float* Q = new_xyz + (size_t)b * M * 3;
    for (int j = 0; j < M; ++j) {
      const float nx = Q[j * 3 + 0], ny = Q[j * 3 + 1], nz = Q[j * 3 + 2];
      int* row = idx + ((size_t)b * M + j) * nsample;
      int cnt = 0;
      for (int k = 0; k < N && cnt < nsample; ++k) {
        const float d2 = sq3(nx - P[k * 3 + 0], ny - P[k * 3 + 1], nz - P[k * 3 + 2]);
        if (d2 < radius2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) row[l] = k;
          row[cnt] = k;
          ++cnt;
        }
      }
    }
  }
}

/* _ext.group_points(points[B,C,N], idx[B,M,K]) -> out[B,C,M,K]   (pointnet2_utils.py:237) */
ORACLE_API void ref_group_points(int B, int C, int N, int M, int K, const float* points, const int* idx,
                                 float* out) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float* src = points + ((size_t)b * C + c) * N;
      float* dst = out + ((size_t)b * C + c) * M * K;
      const int* ii = idx + (size_t)b * M * K;
      for (int l = 0; l < M * K; ++l) dst[l] = src[ii[l]];
    }
}

/* _ext.group_points_grad(grad_out[B,C,M,K], idx[B,M,K], N) -> grad_points[B,C,N]   (pointnet2_utils.py:257)
 * (upstream accumulates with atomicAdd in unspecified order; here index order) */
ORACLE_API void ref_group_points_grad(int B, int C, int N, int M, int K, const float* grad_out, const int* idx,
                                      float* grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)B * C * N);
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float* g = grad_out + ((size_t)b * C + c) * M * K;
      float* dst = grad_points + ((size_t)b * C + c) * N;
      const int* ii = idx + (size_t)b * M * K;
      for (int l = 0; l < M * K; ++l) dst[ii[l]] += g[l];
    }
}

/* _ext.three_nn(unknown[B,n,3], known[B,m,3]) -> dist2[B,n,3], idx[B,n,3]   (pointnet2_utils.py:145; unused by PTT) */
ORACLE_API void ref_three_nn(int B, int n, int m, const float* unknown, const float* known, float* dist2,
                             int* idx) {
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < n; ++j) {
      const float* u = unknown + ((size_t)b * n + j) * 3;
      double best1 = 1e40, best2 = 1e40, best3 = 1e40;
      int besti1 = 0, besti2 = 0, besti3 = 0;
      for (int k = 0; k < m; ++k) {
        const float* p = known + ((size_t)b * m + k) * 3;
        const float d = sq3(u[0] - p[0], u[1] - p[1], u[2] - p[2]);
        if (d < best1) {
          best3 = best2; besti3 = besti2;
          best2 = best1; besti2 = besti1;
          best1 = d; besti1 = k;
        } else if (d < best2) {
          best3 = best2; besti3 = besti2;
          best2 = d; besti2 = k;
        } else if (d < best3) {
          best3 = d; besti3 = k;
        }
      }
      float* dd = dist2 + ((size_t)b * n + j) * 3;
      int* ii = idx + ((size_t)b * n + j) * 3;
      dd[0] = (float)best1; dd[1] = (float)best2; dd[2] = (float)best3;
      ii[0] = besti1; ii[1] = besti2; ii[2] = besti3;
    }
}

/* _ext.three_interpolate(points[B,c,m], idx[B,n,3], weight[B,n,3]) -> out[B,c,n]   (pointnet2_utils.py:182) */
ORACLE_API void ref_three_interpolate(int B, int c, int m, int n, const float* points, const int* idx,
                                      const float* weight, float* out) {
  for (int b = 0; b < B; ++b)
    for (int l = 0; l < c; ++l) {
      const float* src = points + ((size_t)b * c + l) * m;
      for (int j = 0; j < n; ++j) {
        const float* w = weight + ((size_t)b * n + j) * 3;
        const int* ii = idx + ((size_t)b * n + j) * 3;
        /* a*w1 + b*w2 + c*w3 under the same contraction rule */
        out[((size_t)b * c + l) * n + j] = fmaf(src[ii[2]], w[2], fmaf(src[ii[0]], w[0], src[ii[1]] * w[1]));
      }
    }
}

/* _ext.three_interpolate_grad(grad_out[B,c,n], idx, weight, m) -> grad_points[B,c,m]   (pointnet2_utils.py:204) */
ORACLE_API void ref_three_interpolate_grad(int B, int c, int n, int m, const float* grad_out, const int* idx,
                                           const float* weight, float* grad_points) {
  memset(grad_points, 0, sizeof(float) * (size_t)B * c * m);
  for (int b = 0; b < B; ++b)
    for (int l = 0; l < c; ++l) {
      float* dst = grad_points + ((size_t)b * c + l) * m;
      for (int j = 0; j < n; ++j) {
        const float g = grad_out[((size_t)b * c + l) * n + j];
        const float* w = weight + ((size_t)b * n + j) * 3;
        const int* ii = idx + ((size_t)b * n + j) * 3;
        dst[ii[0]] += g * w[0];
        dst[ii[1]] += g * w[1];
        dst[ii[2]] += g * w[2];
      }
    }
}

/* kNN used by TransformerBlock.forward (variants.py:150-151): squared distances exactly as
 * layer_utils.square_distance computes them in fp32 -- ((dx*dx + dy*dy) + dz*dz), every step rounded,
 * no FMA -- then the k smallest per row in (distance, index) order (a STABLE argsort; torch's
 * argsort is unstable, so exact ties may come out in another order in the reference). */
ORACLE_API void ref_knn(int B, int n, int k, const float* xyz, int* knn_idx) {
  float* d = (float*)malloc(sizeof(float) * (size_t)n);
  unsigned char* used = (unsigned char*)malloc((size_t)n);
  for (int b = 0; b < B; ++b) {
    const float* P = xyz + (size_t)b * n * 3;
    for (int i = 0; i < n; ++i) {
      for (int j = 0; j < n; ++j) {
        volatile float dx = P[i * 3 + 0] - P[j * 3 + 0];
        volatile float dy = P[i * 3 + 1] - P[j * 3 + 1];
        volatile float dz = P[i * 3 + 2] - P[j * 3 + 2];
        volatile float xx = dx * dx, yy = dy * dy, zz = dz * dz;
        volatile float s = xx + yy;
        d[j] = s + zz;
        used[j] = 0;
      }
      for (int t = 0; t < k; ++t) {
        int best = -1;
        for (int j = 0; j < n; ++j)
          if (!used[j] && (best < 0 || d[j] < d[best])) best = j;
        if (best < 0) best = 0; /* k > n: unreachable for valid inputs */
        used[best] = 1;
        knn_idx[((size_t)b * n + i) * k + t] = best;
      }
    }
  }
  free(d);
  free(used);
}
