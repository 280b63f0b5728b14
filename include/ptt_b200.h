/*
 * ptt_b200.h -- C ABI of libptt_b200.so: the B200 (sm_100a) point-feature hot path of PTT.
 *
 * This is the drop-in boundary.  The reference (shanjiayao/PTT) reaches its native code through the
 * Python import `pointnet2_ops._ext` (ptt/models/backbones_3d/pointnet2/pointnet2_utils.py:24), a
 * third-party pybind11/ATen extension.  Each `ptt_*` entry point below replaces one `_ext.*`
 * function (cited per function as file:line of the reference call site) or one fused stretch of
 * the nn.Module layer above it.  The Python side binds them with ctypes (ptt_b200/_lib.py); the
 * binding a reference maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named h_*;
 *   - all tensors are contiguous; float = fp32, int = int32; layouts are written as (dims);
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it, never synchronised;
 *   - no device-memory allocation: scratch memory comes in through `workspace` (query its size with
 *     the matching *_workspace_bytes call; 256-byte aligned).  Process-wide state is limited to what
 *     this header names: the launch counter, the fault word (one 64-byte pinned host allocation made
 *     when the first tensor-core kernel family is configured) and the per-device "kernel attributes
 *     configured" bits.  The process-wide TUNING switches live in ptt_b200_tuning.h, not here;
 *   - return value: 0 on success, a cudaError_t (> 0) from the launch, or a PTT_ERR_* (< 0);
 *     nothing ever calls exit() (upstream pointnet2_ops does on launch failure) and no kernel traps:
 *     a kernel that gives up a bounded barrier wait raises the fault word instead (ptt_fault_status);
 *   - inputs are borrowed and never written; outputs never alias inputs.
 */
#ifndef PTT_B200_H_
#define PTT_B200_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PTT_API __attribute__((visibility("default")))
#else
#define PTT_API
#endif

#define PTT_OK 0
#define PTT_ERR_INVALID_ARGUMENT (-1) /* null pointer, negative size, npoint > N ... */
#define PTT_ERR_UNSUPPORTED (-2)      /* shape outside what the kernel family covers  */
#define PTT_ERR_WORKSPACE (-3)        /* workspace missing or too small               */
#define PTT_ERR_DEVICE_FAULT (-4)     /* an earlier kernel gave up a bounded wait     */

typedef void* ptt_stream_t; /* cudaStream_t */

/* "ptt_b200 <version> sm_100a" */
PTT_API const char* ptt_version(void);
/* Text for a return code of any function below (static storage). */
PTT_API const char* ptt_error_string(int code);

/* Number of kernels this library has launched in this process so far (monotonic; diagnostics only). */
PTT_API unsigned long long ptt_launch_count(void);

/* Device-fault word.  The tensor-core kernels hand tiles between warps through mbarriers with BOUNDED waits; a
 * waiter that expires (a protocol bug, never observed in the shipped configurations) raises this word and the kernel
 * drains instead of hanging or trapping, so the CUDA context stays usable.  From then on every entry point
 * returns PTT_ERR_DEVICE_FAULT (sticky); results produced since the last call that returned 0 are invalid.
 * ptt_fault_status: 0 or PTT_ERR_DEVICE_FAULT (a plain host read: meaningful after the stream was synchronised).
 * ptt_fault_clear: re-arm after the caller has synchronised and discarded the affected results. */
PTT_API int ptt_fault_status(void);
PTT_API void ptt_fault_clear(void);

/* ---------------------------------------------------------------------------------------------
 * a1  _ext.furthest_point_sampling(xyz, npoint)                     pointnet2_utils.py:78
 *     xyz (B,N,3) -> idx (B,npoint) int32.  idx[.,0] = 0; points with |p|^2 <= 1e-3 are never
 *     candidates; exact ties resolve as upstream's block tree reduction does.
 *     new_xyz (B,npoint,3), if not NULL, receives xyz[idx] (fuses _ext.gather_points of
 *     pointnet2_modules.py:79-81).  Workspace is only needed for N > 8192 (B*N floats: the clouds
 *     that do not fit one CTA's registers keep their running distances in memory).
 * ------------------------------------------------------------------------------------------- */
PTT_API size_t ptt_furthest_point_sampling_workspace_bytes(int B, int N, int npoint);
PTT_API int ptt_furthest_point_sampling(const float* xyz, int B, int N, int npoint, int* idx, float* new_xyz,
                                void* workspace, size_t workspace_bytes, ptt_stream_t stream);

/*     _ext.furthest_point_sampling_with_dist(dist, npoint)          pointnet2_utils.py:48
 *     dist (B,N,N) -> idx (B,npoint).  ('ffps' sampling; dead in the shipped configs.) */
PTT_API size_t ptt_furthest_point_sampling_with_dist_workspace_bytes(int B, int N, int npoint);
PTT_API int ptt_furthest_point_sampling_with_dist(const float* dist, int B, int N, int npoint, int* idx,
                                          void* workspace, size_t workspace_bytes, ptt_stream_t stream);

/* a2  _ext.gather_points(points, idx)                                pointnet2_utils.py:112
 *     points (B,C,N), idx (B,M) -> out (B,C,M) */
PTT_API int ptt_gather_points(const float* points, const int* idx, int B, int C, int N, int M, float* out,
                      ptt_stream_t stream);
/*     _ext.gather_points_grad(grad_out, idx, N)                      pointnet2_utils.py:118
 *     grad_out (B,C,M), idx (B,M) -> grad_points (B,C,N), zero-filled here then scatter-added */
PTT_API int ptt_gather_points_grad(const float* grad_out, const int* idx, int B, int C, int N, int M,
                           float* grad_points, ptt_stream_t stream);

/* a3  _ext.ball_query(new_xyz, xyz, radius, nsample)                 pointnet2_utils.py:287
 *     new_xyz (B,M,3), xyz (B,N,3) -> idx (B,M,nsample) int32: the first nsample k (ascending) with
 *     d2 < radius^2 (fp32), tail padded with the first hit, all-zero row when there is none. */
PTT_API int ptt_ball_query(const float* new_xyz, const float* xyz, int B, int N, int M, float radius,
                   int nsample, int* idx, ptt_stream_t stream);
/*     The ball queries of ALL set-abstraction layers of a backbone branch in one launch.  Layers after the first
 *     sample with 'sequence' = arange(npoint) (pointnet2_modules.py:70-71; ptt.yaml SAMPLE_METHOD), so with
 *     samples (B,M_0,3) = xyz gathered by the first layer's FPS, level l queries the centres samples[:, :M_l] against
 *     the cloud xyz (B,N,3) for l = 0 and samples[:, :M_{l-1}] for l >= 1 (M_l non-increasing, levels <= 4):
 *     idx[l] (B,M_l,nsample_l), each exactly what ptt_ball_query returns for that pair.  h_* are HOST arrays. */
PTT_API int ptt_ball_query_nested(const float* xyz, const float* samples, int B, int N, int levels, const int* h_M,
                          const float* h_radius, const int* h_nsample, int* const* h_idx, ptt_stream_t stream);

/* a4  _ext.group_points(points, idx)                                 pointnet2_utils.py:237
 *     points (B,C,N), idx (B,M,K) -> out (B,C,M,K) */
PTT_API int ptt_group_points(const float* points, const int* idx, int B, int C, int N, int M, int K,
                     float* out, ptt_stream_t stream);
/*     _ext.group_points_grad(grad_out, idx, N)                       pointnet2_utils.py:257
 *     grad_out (B,C,M,K) -> grad_points (B,C,N), zero-filled here then scatter-added */
PTT_API int ptt_group_points_grad(const float* grad_out, const int* idx, int B, int C, int N, int M, int K,
                          float* grad_points, ptt_stream_t stream);

/*     _ext.three_nn(unknown, known)                                  pointnet2_utils.py:145
 *     unknown (B,n,3), known (B,m,3) -> dist2 (B,n,3) (SQUARED distances), idx (B,n,3) */
PTT_API int ptt_three_nn(const float* unknown, const float* known, int B, int n, int m, float* dist2, int* idx,
                 ptt_stream_t stream);
/*     _ext.three_interpolate(points, idx, weight)                    pointnet2_utils.py:182
 *     points (B,c,m), idx (B,n,3), weight (B,n,3) -> out (B,c,n) */
PTT_API int ptt_three_interpolate(const float* points, const int* idx, const float* weight, int B, int c, int m,
                          int n, float* out, ptt_stream_t stream);
/*     _ext.three_interpolate_grad(grad_out, idx, weight, m)          pointnet2_utils.py:204 */
PTT_API int ptt_three_interpolate_grad(const float* grad_out, const int* idx, const float* weight, int B, int c,
                               int n, int m, float* grad_points, ptt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Layout helpers for the fused path (point-major activations).
 *   channel-major (B,C,N)  <->  point-major (B,N,ld) with ld >= C (padding columns written as 0)
 * ------------------------------------------------------------------------------------------- */
PTT_API int ptt_cm_to_pm(const float* src_cm, int B, int C, int N, float* dst_pm, int ld, ptt_stream_t stream);
PTT_API int ptt_pm_to_cm(const float* src_pm, int ld, int B, int C, int N, float* dst_cm, ptt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a5-a7  QueryAndGroup + SharedMLP + max-pool of one set-abstraction layer, eval mode
 *        pointnet2_utils.py:320-380, pytorch_utils.py:12-36, pointnet2_modules.py:83-88
 *
 *   xyz (B,N,3); feats point-major (B,N,ldf) with C valid channels, or NULL when C == 0;
 *   new_xyz (B,M,3) centres; idx (B,M,ns) from ptt_ball_query.
 *   Row (b,j,s) of the implicit grouped matrix is [ feats[b,idx,0:C] | (xyz[b,idx]-new_xyz[b,j]) (/radius) ]
 *   (the grouped tensor is never materialised).  n_layers (1..4) of y = relu(scale*(W x) + shift)
 *   (BatchNorm folded into scale/shift), then max over the ns rows of each centre.
 *
 *   dims[0..n_layers]: dims[0] = C + 3 is implied by C; dims[l] = output channels of layer l (l >= 1).
 *   params: the packed image produced by ptt_sa_pack_params (device memory).
 *   out_pm (B,M,ld_out) point-major and/or out_cm (B,Cout,M) channel-major (either may be NULL).
 * ------------------------------------------------------------------------------------------- */
/* Packing: weights[l] is the conv weight (dims[l+1], dims[l]) row-major with input channels in the
 * REFERENCE order [xyz(3) | feats(C)] (QueryAndGroup's cat, pointnet2_utils.py:359); scale/shift
 * (dims[l+1]).  h_* arrays of n_layers DEVICE pointers live on the host. */
PTT_API size_t ptt_sa_params_floats(int C, int n_layers, const int* h_dims);
PTT_API int ptt_sa_pack_params(int C, int n_layers, const int* h_dims, const float* const* h_weights,
                       const float* const* h_scale, const float* const* h_shift, float* params,
                       ptt_stream_t stream);
PTT_API size_t ptt_sa_mlp_workspace_bytes(int B, int N, int M, int ns, int C, int n_layers, const int* h_dims);
PTT_API int ptt_sa_mlp_fwd(const float* xyz, const float* feats, int ldf, const float* new_xyz, const int* idx,
                   int B, int N, int M, int ns, int C, float radius, int normalize_xyz, int n_layers,
                   const int* h_dims, const float* params, float* out_pm, int ld_out, float* out_cm,
                   void* workspace, size_t workspace_bytes, ptt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * N1  CosineSimAug fusion (similarity_modules/p2b_xcoor.py:25-43)
 *
 *   search_feats (B,n2,lds>=f), template_feats (B,n1,ldt>=f) point-major, template_xyz (B,n1,3)
 *   -> out_pm (B,n2,ld_out): max over the n1 templates of SharedMLP([cos(t,s) | xyz_t | feats_t]).
 *   The 1x1 Conv1d stack that follows (p2b_xcoor.py:44) is two ptt_linear_fwd calls.
 *   params: ptt_sa_pack_params image for C = 3 + f whose layer-0 weight has its (1 + 3 + f) reference input
 *   columns re-ordered to [w_sim, 0, 0 | W0[:, 1:]] (i.e. dims[0] = f + 6): the fusion is a set-abstraction
 *   layer over the templates with the similarity as the only pair-dependent input (see csrc/sa_mlp.cu).
 * ------------------------------------------------------------------------------------------- */
PTT_API size_t ptt_cosine_fusion_workspace_bytes(int B, int n1, int n2, int f, int n_layers, const int* h_dims);
PTT_API int ptt_cosine_fusion_fwd(const float* search_feats, int lds, const float* template_feats, int ldt,
                          const float* template_xyz, int B, int n1, int n2, int f, int n_layers,
                          const int* h_dims, const float* params, float* out_pm, int ld_out, void* workspace,
                          size_t workspace_bytes, ptt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a9  TransformerBlock.forward (kNN vector attention)           transformer_block/variants.py:149-165
 *
 *   knn: xyz (B,n,3) -> knn_idx (B,n,k): the k nearest (self included) by the fp32 squared distance
 *   of layer_utils.square_distance (layer_utils.py:26), ascending, ties lowest index first.
 * ------------------------------------------------------------------------------------------- */
PTT_API int ptt_knn(const float* xyz, int B, int n, int k, int* knn_idx, ptt_stream_t stream);

/*   y (R,ldy)[:, 0:Cout] = act( x (R,ldx)[:, 0:K] . W^T + bias ) (+ residual (R,ldr))
 *   wt: packed transposed weight from ptt_linear_pack (nn.Linear weight (Cout,K) -> (Kp,Cp) image). */
PTT_API size_t ptt_linear_params_floats(int K, int Cout);
PTT_API int ptt_linear_pack(const float* weight, const float* bias, int K, int Cout, float* params,
                    ptt_stream_t stream);
/* The same image from a STRIDED source: W[c, k] = weight[c * ld_c + k * ld_k] (ptt_linear_pack is ld_c = K, ld_k = 1; the
 * transposed weight of an input-gradient contraction is ld_c = 1, ld_k = its row length -- no transposed copy). */
PTT_API int ptt_linear_pack_strided(const float* weight, long long ld_c, long long ld_k, const float* bias, int K, int Cout,
                            float* params, ptt_stream_t stream);
/* Many layers in ONE launch (a training step repacks every layer after the optimiser update): `descs_device` is a DEVICE
 * array of `count` descriptors; every entry is packed as ptt_linear_pack_strided would pack it. */
typedef struct PttPackDesc {
  const float* weight;       /* W[c, k] = weight[c * ld_c + k * ld_k] */
  long long ld_c, ld_k;
  const float* bias;         /* Cout floats or NULL */
  float* params;             /* ptt_linear_params_floats(K, Cout) floats */
  int K, Cout;
} PttPackDesc;
PTT_API int ptt_linear_pack_batch(const PttPackDesc* descs_device, int count, ptt_stream_t stream);
PTT_API int ptt_linear_fwd(const float* x, int ldx, int R, int K, const float* params, int Cout, int relu,
                   const float* residual, int ldr, float* y, int ldy, ptt_stream_t stream);

/*   Whole block.  features (B,n,d_points) -> out (B,n,d_points); attn (B,n,k,d_model) or NULL.
 *   params: image from ptt_transformer_pack_params; tensors are the reference state_dict entries
 *   fc1.{weight,bias}, fc2.{weight,bias}, fc_delta.{0,2}.{weight,bias}, fc_gamma.{0,2}.{weight,bias},
 *   w_qs.weight, w_ks.weight, w_vs.weight (nn.Linear layout (out,in)).
 *   variant: 0 = TransformerBlock, 1 = TransformerBlockOffset (fc2(x - res), variants.py:333). */
PTT_API size_t ptt_transformer_params_floats(int d_points, int d_model);
PTT_API int ptt_transformer_pack_params(int d_points, int d_model, const float* fc1_w, const float* fc1_b,
                                const float* fc2_w, const float* fc2_b, const float* delta0_w,
                                const float* delta0_b, const float* delta2_w, const float* delta2_b,
                                const float* gamma0_w, const float* gamma0_b, const float* gamma2_w,
                                const float* gamma2_b, const float* wq, const float* wk, const float* wv,
                                float* params, ptt_stream_t stream);
PTT_API size_t ptt_transformer_block_workspace_bytes(int B, int n, int k, int d_points, int d_model);
PTT_API int ptt_transformer_block_fwd(const float* xyz, const float* features, int B, int n, int k, int d_points,
                              int d_model, int variant, const float* params, const int* knn_idx_or_null,
                              float* out, float* attn_or_null, void* workspace, size_t workspace_bytes,
                              ptt_stream_t stream);

/* a10 / N4: the same kNN vector-attention core for the other registered blocks (transformer_block/__init__.py:7-17):
 *   variant_flags   bit 0 = Offset (variants.py:297-334); bit 1 = raw: out (B,n,d_model) is the aggregated `res`,
 *                   no fc2 / residual (TransformerBlockMLP's two-layer fc2 :225-229, MulHeadTransformerLayer's proj +
 *                   LayerNorms multitransformer.py:61-63 and TransformerBlockBackbone :294 finish it themselves)
 *   q_features      CrossAttentionBlock (variants.py:190-208): the queries are w_qs(fc1(q_features)); NULL = features
 *   divisor_or_0    softmax temperature (MulHeadTransformerLayer: sqrt(head_dim)); 0 = sqrt(d_model)
 *   pair_scalar/vec TransformerBlockCosine (variants.py:66-88): fc_gamma.0's pre-activation += pair_scalar[b,i,j] *
 *                   pair_vec[c] (the similarity column of fc_sim pushed through fc_gamma.0); both or neither;
 *                   d_model in {64,128,256,512} only (PTT_ERR_UNSUPPORTED otherwise)                               */
PTT_API int ptt_transformer_block_fwd_ex(const float* xyz, const float* features, const float* q_features_or_null, int B,
                                 int n, int k, int d_points, int d_model, int variant_flags, float divisor_or_0,
                                 const float* params, const int* knn_idx_or_null, const float* pair_scalar_or_null,
                                 const float* pair_vec_or_null, float* out, float* attn_or_null, void* workspace,
                                 size_t workspace_bytes, ptt_stream_t stream);
/* sim (B,n,k) = cos(q[b,i,:], kmat[b,knn[b,i,j],:]) with F.cosine_similarity's eps = 1e-8 (variants.py:78-79) */
PTT_API int ptt_pair_cosine(const float* q, int ldq, const float* kmat, int ldk, const int* knn, int B, int n, int k,
                    int d, float* sim, ptt_stream_t stream);
/* y = LayerNorm_C(x) * gamma + beta (+ residual), rows of C channels (multitransformer.py:31-32,62-63) */
PTT_API int ptt_layer_norm_fwd(const float* x, int ldx, int R, int C, const float* gamma, const float* beta, float eps,
                       const float* residual_or_null, int ldr, float* y, int ldy, ptt_stream_t stream);
/* TransformerBlockALL (variants.py:111-124): out[b,i,c] = softmax_i(logits[b,:,c] / divisor) * other[b,i,c];
 * attn_or_null (B,n,C) receives the softmax */
PTT_API int ptt_token_softmax_gate(const float* logits, int ldl, const float* other, int ldo, int B, int n, int C,
                           float divisor, float* out, int ldy, float* attn_or_null, ptt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * a10 TransformerBlockSTD.forward (dense n x n dot-product attention)   transformer_block/variants.py:12-40
 *   features (B,n,d_points) -> out (B,n,d_points); attn (B,n,n) or NULL.  Parameter tensors are the reference
 *   state_dict entries fc1, fc2, fc_delta.{0,2} (weight + bias) and w_qs, w_ks, w_vs (weight).
 * ------------------------------------------------------------------------------------------- */
PTT_API size_t ptt_transformer_std_params_floats(int d_points, int d_model);
PTT_API int ptt_transformer_std_pack_params(int d_points, int d_model, const float* fc1_w, const float* fc1_b,
                                    const float* fc2_w, const float* fc2_b, const float* delta0_w,
                                    const float* delta0_b, const float* delta2_w, const float* delta2_b,
                                    const float* wq, const float* wk, const float* wv, float* params,
                                    ptt_stream_t stream);
PTT_API size_t ptt_transformer_std_workspace_bytes(int B, int n, int d_points, int d_model);
PTT_API int ptt_transformer_std_fwd(const float* xyz, const float* features, int B, int n, int d_points, int d_model,
                            const float* params, float* out, float* attn_or_null, void* workspace,
                            size_t workspace_bytes, ptt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Training path (BASELINE configs[3]): BatchNorm on batch statistics and gradients.
 *     Activations are pair-row matrices (rows, ld) with ld % 4 == 0 and 16-byte aligned bases.  A layer's normalised
 *     output is never stored: consumers apply relu(ka[c] * y + kb[c]) while loading the PRE-BatchNorm output y.
 *     Reference: Conv2d 1x1 -> BatchNorm2d -> ReLU of pytorch_utils.py:12-36 in train() mode, F.max_pool2d of
 *     pointnet2_modules.py:85, and torch autograd of both; nn.Linear of transformer_block/variants.py.
 * ------------------------------------------------------------------------------------------- */
/* ptt_linear_fwd with the A operand transformed on load: x <- relu(a_ka[k] * x + a_kb[k]) (both NULL: plain). */
PTT_API int ptt_linear_fwd_ex(const float* x, int ldx, int R, int K, const float* a_ka, const float* a_kb, const float* params,
                      int Cout, int relu, const float* residual, int ldr, float* y, int ldy, ptt_stream_t stream);
/* ptt_linear_fwd_ex (no activation, no residual) + the column statistics of its result in one call: sums_or_null (2,Cout)
 * double <- sum_r y and sum_r y*y.  Bias-free layers (has_bias == 0) over >= 4096 rows with K, Cout <= 260 / 256 run on a
 * weight-stationary persistent kernel that produces the statistics in its epilogue (csrc/ws_gemm.cu). */
PTT_API int ptt_linear_fwd_stats(const float* x, int ldx, long long R, int K, const float* a_ka, const float* a_kb,
                         const float* params, int Cout, int has_bias, float* y, int ldy, double* sums_or_null,
                         ptt_stream_t stream);
/* Weight gradient: dw (M,ldw)[:, 0:N] += dy (R,ldy)[:, 0:M]^T . f(x (R,ldx)[:, 0:N]), f = identity or relu(ka*x + kb)
 * per column of x; tcgen05 with MN-major operands, split over the rows, fp32 atomics into dw (zero it first).
 * dbias_or_null (M floats, zero it first) += column sums of dy -- the bias gradient of the same layer, accumulated from
 * the rows of dy while they are staged for the contraction (no second pass over dy). */
PTT_API int ptt_linear_wgrad(const float* dy, int ldy, const float* x, int ldx, const float* x_ka, const float* x_kb,
                     long long R, int M, int N, float* dw, int ldw, float* dbias_or_null, ptt_stream_t stream);
/* sums (2,C) double <- column sums of y and y*y over R rows (phase 1 of the two-phase BatchNorm). */
PTT_API int ptt_col_stats(const float* y, int ldy, long long R, int C, double* sums, ptt_stream_t stream);
/* Phase 2: mean / biased variance from sums -> ka = gamma*rstd, kb = beta - mean*ka, mean, rstd (C floats each);
 * running_mean / running_var (may be NULL) updated as torch does (momentum, unbiased variance). */
PTT_API int ptt_bn_train_finalize(const double* sums, long long R, int C, const float* gamma, const float* beta, float eps,
                          float momentum, float* running_mean, float* running_var, float* ka, float* kb, float* mean,
                          float* rstd, ptt_stream_t stream);
/* QueryAndGroup as a pair-row matrix in the reference's channel order (pointnet2_utils.py:350-361):
 * rows_out[(b,j,s), :] = [ (xyz[idx]-new_xyz[j]) (/radius) | feats[idx, 0:C] | 0 ], feats point-major (B,N,ldf). */
PTT_API int ptt_sa_group_rows(const float* xyz, const float* feats, int ldf, const float* new_xyz, const int* idx, int B, int N,
                      int M, int ns, int C, float radius, int normalize_xyz, float* rows_out, int ld, ptt_stream_t stream);
/* Its backward: scatter-add (atomics, like _ext.group_points_grad) of d_rows into d_feats (B,N,ldf) and, when not
 * NULL, d_xyz (B,N,3) / d_new_xyz (B,M,3).  The outputs are accumulated into (zero them first). */
PTT_API int ptt_sa_group_rows_grad(const float* d_rows, int ld, const int* idx, int B, int N, int M, int ns, int C, float radius,
                           int normalize_xyz, float* d_feats, int ldf, float* d_xyz, float* d_new_xyz, ptt_stream_t stream);
/* out (groups,ldo) = max over the ns rows of each group of relu(ka*y + kb); argmax (groups,C) int32 = first maximum. */
PTT_API int ptt_bn_relu_maxpool(const float* y, int ldy, long long groups, int ns, int C, const float* ka, const float* kb,
                        float* out, int ldo, int* argmax, ptt_stream_t stream);
/* Backward of y -> relu(BatchNorm_train(y)) [-> max over ns when argmax is given, dz then (R/ns, ldz)]:
 * dy (R,ld_dy) = gamma*rstd*(m - s1/R - yhat*s2/R), m = dz*[ka*y+kb > 0]; sums (2,C) double <- (s1 = d beta, s2 = d gamma);
 * dparam_or_null (2,C) float <- the same two rows in fp32 (the parameter gradients, no conversion pass). */
PTT_API int ptt_bn_relu_bwd(const float* dz, int ldz, const int* argmax_or_null, int ns, const float* y, int ldy, long long R,
                    int C, const float* ka, const float* kb, const float* mean, const float* rstd, const float* gamma,
                    double* sums, float* dy, int ld_dy, float* dparam_or_null, ptt_stream_t stream);

/* Training of the kNN vector-attention block (variants.py:149-165): the forward is ptt_transformer_block_fwd with attn
 * requested; these are the element kernels of its backward (ptt_b200/train_ops.py::_TransformerTrain), between the
 * ptt_linear_fwd / ptt_linear_wgrad contractions.  Pair-row matrices (B*n*k, ld), token matrices (B*n, ldt).
 * ptt_transformer_block_workspace_layout: float offsets of what the forward leaves in its workspace,
 * h_out[0..6] = knn, x, qkv, res, g = relu(fc_gamma.0(.)), pos + v, ld. */
PTT_API int ptt_transformer_block_workspace_layout(int B, int n, int k, int d_points, int d_model, size_t* h_out);
/* dvp = attn * dres_i ;  dlogit = attn * (dres_i * vp - sum_j attn * dres_i * vp) / divisor   (softmax over the k rows of a token) */
PTT_API int ptt_tr_softmax_bwd(const float* dres, int ldr, const float* attn, const float* vp, int ld, long long tokens, int k,
                       int dm, float divisor, float* dlogit, float* dvp, ptt_stream_t stream);
/* h1 = relu(fc_delta.0(xyz_i - xyz_j)), delta (pairs,4) = (xyz_i - xyz_j, 0), a_in = q_i - k_j + (vp_ij - v_j);
 * delta0_w (dm,3) / delta0_b (dm) in nn.Linear layout; q, kk, v token matrices (B*n, ldt) */
PTT_API int ptt_tr_pair_inputs(const float* xyz, const int* knn, int B, int n, int k, int dm, const float* delta0_w,
                       const float* delta0_b, const float* q, const float* kk, const float* v, int ldt, const float* vp, int ld,
                       float* h1, float* delta, float* a_in, ptt_stream_t stream);
/* dy <- dy * [ref > 0], count floats (multiple of 4), same layout */
PTT_API int ptt_tr_mask_positive(float* dy, const float* ref, long long count, ptt_stream_t stream);
/* out (R,d) = x (R,ldx)[:, 0:d] . W^T [* (mask_ref > 0)] for a square d x d layer (d = 256 or 512) over many rows: the
 * persistent CTA-pair kernel of the block's forward passes with plain rows as the operand -- the pair-level input-gradient
 * contractions of the backward (params = ptt_linear_pack_strided image of the (d,d) matrix, bias-free; mask_ref (R,d) or
 * NULL = the ReLU backward against the stored activation, fused into the epilogue).  PTT_ERR_UNSUPPORTED for other d. */
PTT_API int ptt_tr_rows_linear(const float* x, int ldx, long long R, int d, const float* params, const float* mask_ref_or_null,
                       float* out, ptt_stream_t stream);
/* da <- da + dvp (= dpos) in place; dq_i = sum_j da_ij; dk[knn] -= da; dv[knn] += dvp (atomics; zero dk / dv first) */
PTT_API int ptt_tr_pair_scatter(float* da, const float* dvp, int ld, const int* knn, int B, int n, int k, int dm, float* dq,
                        float* dk, float* dv, int ldt, ptt_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * N3  The per-frame pre / post-processing of the tracking loop, for T independent tracklets at once
 *     (tools/eval_utils/eval_tracking_utils.py:140-274; ptt/datasets/kitti/kitti_tracking_utils.py:192-367).
 *     Box state of a tracklet = 15 doubles: center(3) | rotation matrix row-major(9) | wlh(3).  Clouds are padded
 *     batches (T,cap,3) float32 with per-tracklet counts.  Arithmetic as restated in oracle/tracking_ref.py.
 * ------------------------------------------------------------------------------------------- */
/* np.random.seed(seed): the first n 32-bit outputs of MT19937 (init_genrand seeding) into HOST memory h_out. */
PTT_API int ptt_mt19937_stream(unsigned seed, int n, unsigned* h_out);
/* crop_center_pc / get_model (kitti_tracking_utils.py:219-237,300-340): every source s (1 or 2; h_* are HOST arrays of
 * n_sources DEVICE pointers / ints) is cropped around its box -- world-frame AABB of the 4*scale box +- 2*offset, then
 * box-frame AABB of the scale box +- margin (search != 0: offset + 0.6*wlh[1], the search area :323; else offset, the
 * template :333) -- and written in the box frame, order preserved, sources concatenated: out (T,cap_out,3),
 * out_counts (T).  h_precropped[s] != 0: source s is copied as it is (an already cropped part of the template). */
PTT_API int ptt_track_crop(int T, int n_sources, const float* const* h_points, const int* const* h_counts,
                   const double* const* h_boxes, const int* h_caps, const int* h_precropped, double offset, double scale,
                   int search, float* out, int cap_out, int* out_counts, ptt_stream_t stream);
/* regularize_pc(istrain=False) (:342-367): n = counts[t] points -> size points: zeros when n <= 2, a copy when
 * n == size, else np.random.seed(1); points[np.random.randint(0, n, size)].  mt_stream = ptt_mt19937_stream(1, ...)
 * in DEVICE memory (mt_len >= 4*size); mt_pos (T) receives the stream position after a resampling (unchanged otherwise). */
PTT_API int ptt_track_regularize(const float* points, const int* counts, int T, int cap, int size, const unsigned* mt_stream,
                         int mt_len, int* mt_pos, float* out, ptt_stream_t stream);
/* post_process (eval_tracking_utils.py:266-274) after the on-device proposal selection: get_box_by_offset (:192-216) of
 * best_box (T,ld_best) = (x, y, z, theta_degrees, ...) float32 applied to box_state (T,15) in place (its two
 * np.random.uniform(-1,1) clamps consume mt_stream at mt_pos); the new state is also stored at
 * results[*frame_idx] ((max_frames,T,15)) and *frame_idx is incremented. */
PTT_API int ptt_track_update(const float* best_box, int ld_best, double* box_state, int T, int use_z, const unsigned* mt_stream,
                     int mt_len, int* mt_pos, double* results, int max_frames, int* frame_idx, ptt_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PTT_B200_H_ */
