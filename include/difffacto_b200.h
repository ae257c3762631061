/*
 * difffacto_b200 -- C ABI of the B200-native DiffFacto sampling hot path.
 *
 * One shared library (difffacto_b200/lib/libdifffacto_b200.so), plain C linkage, raw device
 * pointers + sizes + a CUDA stream; no torch types anywhere.  Every entry point returns an
 * int status (DFB200_OK == 0) instead of the reference's `exit(-1)` on launch failure
 * (reference: pointnet2_ops_lib/pointnet2_ops/_ext-src/include/cuda_utils.h:30-39).
 * Ownership follows the reference: the CALLER owns and allocates every buffer (inputs, outputs,
 * scratch); the library never allocates device memory.  Initialisation the semantics rely on
 * (idx zeros for ball_query, temp = 1e10 for FPS, zeroed grads) is done INSIDE the library, so
 * output buffers may be uninitialised on entry.
 *
 * Each declaration cites the reference interface (file:line under /root/reference) it replaces.
 * All tensors are contiguous, row-major, fp32 / int32, resident on the current device.
 * `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 */
#ifndef DIFFFACTO_B200_H_
#define DIFFFACTO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFB200_OK 0
#define DFB200_ERR_INVALID_ARG 1
#define DFB200_ERR_UNSUPPORTED 2
#define DFB200_ERR_CUDA 3
#define DFB200_ERR_WORKSPACE 4

typedef void* dfb200_stream_t;

/* ABI version of this header (bumped on any signature change). */
int dfb200_abi_version(void);
/* Human-readable message for the last non-OK status returned on this thread. */
const char* dfb200_last_error(void);
/* Number of kernels this library has launched on this process so far (bench.py's gpu_launches). */
unsigned long long dfb200_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * pointnet2_ops  (reference: pointnet2_ops_lib/pointnet2_ops/_ext-src/src/bindings.cpp:6-19)
 * ------------------------------------------------------------------------------------------ */

/* out[b,c,j] = points[b,c,idx[b,j]].  points (b,c,n), idx (b,npoints) -> out (b,c,npoints).
 * Replaces gather_points_kernel_wrapper, sampling.cpp:4-6 / sampling_gpu.cu:22-30. */
int dfb200_gather_points(int b, int c, int n, int npoints, const float* points, const int* idx,
                         float* out, dfb200_stream_t stream);
/* grad_points[b,c,idx[b,j]] += grad_out[b,c,j]; grad_points (b,c,n) is zeroed by the callee.
 * Replaces gather_points_grad_kernel_wrapper, sampling.cpp:7-9 / sampling_gpu.cu:49-57. */
int dfb200_gather_points_grad(int b, int c, int n, int npoints, const float* grad_out,
                              const int* idx, float* grad_points, dfb200_stream_t stream);
/* Iterative furthest point sampling.  dataset (b,n,3) -> idxs (b,m) int32, idxs[:,0] = 0.
 * `temp` (b,n) is optional scratch kept for signature compatibility: if non-NULL it receives the
 * final per-point min squared distances (1e10 for never-updated points) exactly as the
 * reference leaves them.  Bit-exact with the reference kernel incl. the |p|^2 <= 1e-3 skip and
 * its tree-reduction tie rule.
 * Replaces furthest_point_sampling_kernel_wrapper, sampling.cpp:11-13 / sampling_gpu.cu:175-229. */
int dfb200_furthest_point_sampling(int b, int n, int m, const float* dataset, float* temp,
                                   int* idxs, dfb200_stream_t stream);
/* idx[b,j,:] = first `nsample` indices k (ascending) with |new_xyz[b,j]-xyz[b,k]|^2 < radius^2,
 * padded with the first hit; all zeros if the ball is empty.  new_xyz (b,m,3), xyz (b,n,3).
 * Every element of idx is written (no zero-initialisation by the caller is needed, unlike the
 * reference).  For 1024 <= n <= 8192 two kernels are enqueued on `stream` (uniform-grid pass, then
 * an ordered scan of the clouds the first pass handed over); idx is used to pass a -1 marker between
 * them, so it must not be read concurrently on another stream before the call's work has finished.
 * Replaces query_ball_point_kernel_wrapper, ball_query.cpp:4-6 / ball_query_gpu.cu:46-54. */
int dfb200_query_ball_point(int b, int n, int m, float radius, int nsample, const float* new_xyz,
                            const float* xyz, int* idx, dfb200_stream_t stream);
/* out[b,c,j,k] = points[b,c,idx[b,j,k]].  points (b,c,n), idx (b,npoints,nsample).
 * Replaces group_points_kernel_wrapper, group_points.cpp:4-6 / group_points_gpu.cu:30-39. */
int dfb200_group_points(int b, int c, int n, int npoints, int nsample, const float* points,
                        const int* idx, float* out, dfb200_stream_t stream);
/* Scatter-add of grad_out (b,c,npoints,nsample) into grad_points (b,c,n) (zeroed by callee).
 * Replaces group_points_grad_kernel_wrapper, group_points.cpp:8-10 / group_points_gpu.cu:66-75. */
int dfb200_group_points_grad(int b, int c, int n, int npoints, int nsample, const float* grad_out,
                             const int* idx, float* grad_points, dfb200_stream_t stream);
/* Three nearest `known` points of each `unknown` point: SQUARED distances ascending + indices.
 * unknown (b,n,3), known (b,m,3) -> dist2 (b,n,3), idx (b,n,3).
 * Replaces three_nn_kernel_wrapper, interpolate.cpp:4-5 / interpolate_gpu.cu:61-68. */
int dfb200_three_nn(int b, int n, int m, const float* unknown, const float* known, float* dist2,
                    int* idx, dfb200_stream_t stream);
/* out[b,c,j] = sum_q weight[b,j,q] * points[b,c,idx[b,j,q]].  points (b,c,m) -> out (b,c,n).
 * Replaces three_interpolate_kernel_wrapper, interpolate.cpp:6-8 / interpolate_gpu.cu:103-111. */
int dfb200_three_interpolate(int b, int c, int m, int n, const float* points, const int* idx,
                             const float* weight, float* out, dfb200_stream_t stream);
/* grad_points (b,c,m), zeroed by callee.
 * Replaces three_interpolate_grad_kernel_wrapper, interpolate.cpp:9-12 / interpolate_gpu.cu:145-154. */
int dfb200_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out, const int* idx,
                                  const float* weight, float* grad_points, dfb200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Evaluation kernels
 * ------------------------------------------------------------------------------------------ */

/* Chamfer distance, both directions.  xyz1 (b,n,3), xyz2 (b,m,3) -> dist1 (b,n), idx1 (b,n),
 * dist2 (b,m), idx2 (b,m): squared distance to / index of the nearest point of the other cloud
 * (earliest index wins ties).  Runs on `stream` (the reference uses the legacy default stream).
 * Replaces chamfer_cuda_forward, python/difffacto/metrics/chamfer_dist/chamfer.cu:147-171. */
int dfb200_chamfer_forward(int b, int n, const float* xyz1, int m, const float* xyz2, float* dist1,
                           float* dist2, int* idx1, int* idx2, dfb200_stream_t stream);
/* grad_xyz1 (b,n,3), grad_xyz2 (b,m,3), zeroed by callee.
 * Replaces chamfer_cuda_backward, chamfer.cu:203-229. */
int dfb200_chamfer_backward(int b, int n, const float* xyz1, int m, const float* xyz2,
                            const int* idx1, const int* idx2, const float* grad_dist1,
                            const float* grad_dist2, float* grad_xyz1, float* grad_xyz2,
                            dfb200_stream_t stream);

/* Approximate earth mover's distance by the auction algorithm (n == m, n % 1024 == 0, b <= 512,
 * coordinates in [0,1]).  All scratch is caller-allocated with the reference's shapes
 * (python/difffacto/metrics/emd/emd_module.py:46-57); assignment/assignment_inv/price/... are
 * (re)initialised by the callee.  dist (b,n) squared distances of the final assignment (b,n).
 * One persistent kernel runs all `iters` auction rounds (the reference launches 7 kernels per
 * round).  Replaces emd_cuda_forward, python/difffacto/metrics/emd/emd_cuda.cu:228-282. */
int dfb200_emd_forward(int b, int n, const float* xyz1, const float* xyz2, float* dist,
                       int* assignment, float* price, int* assignment_inv, int* bid,
                       float* bid_increments, float* max_increments, int* unass_idx,
                       int* unass_cnt, int* unass_cnt_sum, int* cnt_tmp, int* max_idx, float eps,
                       int iters, dfb200_stream_t stream);
/* grad_xyz1 (b,n,3), zeroed by callee.  Replaces emd_cuda_backward, emd_cuda.cu:302-317. */
int dfb200_emd_backward(int b, int n, const float* xyz1, const float* xyz2, float* grad_xyz,
                        const float* grad_dist, const int* assignment, dfb200_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Cross-diffusion denoiser (TransformerNet) and anchored DDPM
 * ------------------------------------------------------------------------------------------ */

/* Static architecture of NETS['TransformerNet'] as resolved from the config
 * (python/difffacto/models/diffusions/nets/attention.py:318-383; configs/gen_chair.py:50-66). */
typedef struct dfb200_denoiser_cfg {
  int in_channels;  /* 3: channels of x before the concatenations                        */
  int out_channels; /* 3                                                                  */
  int n_heads;      /* 8                                                                  */
  int d_head;       /* 16  (inner_dim = n_heads*d_head must be 128)                       */
  int depth;        /* 5 transformer blocks (single_attn=True: cross-attn + GEGLU FF)     */
  int context_dim;  /* 262: channels of the concatenated ctx list before +n_class +256    */
  int n_class;      /* 4 part tokens (keys/values)                                        */
  int flags;        /* DFB200_NET_* below                                                 */
} dfb200_denoiser_cfg;

#define DFB200_NET_CLASS_COND 1        /* class_cond and not add_class_cond: one-hot in ctx */
#define DFB200_NET_CAT_PARAMS_TO_X 2   /* cat anchors+variances to x (attention.py:400-404) */
#define DFB200_NET_CAT_CLASS_TO_X 4    /* cat one-hot(assignment) to x (attention.py:405-407) */
#define DFB200_NET_MASK_UNREFERENCED 8 /* mask_out_unreferenced_code (attention.py:409)     */
#define DFB200_NET_INCLUDE_STD 16      /* include_std: sqrt(variance) instead of variance   */

/* Precision mode of the denoiser GEMMs. */
#define DFB200_MODE_FP32 0 /* CUDA-core FFMA, fp32 everywhere (reference numerics)          */
#define DFB200_MODE_BF16 1 /* tcgen05 tensor cores: bf16 operands, fp32 accumulate in TMEM  */
#define DFB200_MODE_TF32 2 /* tcgen05 kind::tf32: fp32 operands rounded to 10 mantissa bits, fp32 accumulate; LayerNorm, softmax,
                              erf-GELU, proj_in/out and every bias in fp32 -- the tensor-core mode at reference tolerance (<= 2e-3) */

/* Number of fp32 parameter tensors dfb200_denoiser_pack expects: 12 + 13*depth, in this order
 * (names as in the reference state_dict under `diffusion.model.`):
 *   pre_norm.weight, pre_norm.bias, post_norm.weight, post_norm.bias, proj_in.weight,
 *   proj_in.bias, time_embed.net.0.proj.weight, time_embed.net.0.proj.bias,
 *   time_embed.net.2.weight, time_embed.net.2.bias, proj_out.weight, proj_out.bias,
 *   then for each block i: norm2.weight, norm2.bias, norm3.weight, norm3.bias, attn2.to_q.weight,
 *   attn2.to_k.weight, attn2.to_v.weight, attn2.to_out.0.weight, attn2.to_out.0.bias,
 *   ff.net.0.proj.weight, ff.net.0.proj.bias, ff.net.2.weight, ff.net.2.bias. */
int dfb200_denoiser_num_params(const dfb200_denoiser_cfg* cfg);
/* Bytes of the packed device weight image (fp32 copies + bf16 UMMA operand images). */
size_t dfb200_denoiser_packed_bytes(const dfb200_denoiser_cfg* cfg);
/* Build the packed image in caller-owned device memory from device fp32 parameter tensors. */
int dfb200_denoiser_pack(const dfb200_denoiser_cfg* cfg, const float* const* params, int n_params,
                         void* packed, dfb200_stream_t stream);
/* Scratch bytes for one forward at (B,N) in `mode`. */
size_t dfb200_denoiser_workspace_bytes(const dfb200_denoiser_cfg* cfg, int mode, int B, int N);

/* eps = TransformerNet(x, t, ctx, anchors, variances, valid_id, anchor_assignment).
 *   x, anchors, variances, eps_out : (B,3,N) channel-major (the memory the reference's
 *                                    `anchors.transpose(1,2)` views alias)
 *   t          : (B) fp32, the value the reference feeds to timestep_embedding
 *   ctx        : (B, context_dim, n_class) = torch.cat(ctx_list, dim=1)
 *   anchor_assignment : (B,N) int32 part id of each point;  valid_id : (B,n_class) or NULL
 * Replaces TransformerNet.forward/_forward_attn, attention.py:385-440 (+ blocks :161-306). */
int dfb200_denoiser_forward(const dfb200_denoiser_cfg* cfg, const void* packed, int mode, int B,
                            int N, const float* x, const float* t, const float* ctx,
                            const float* anchors, const float* variances,
                            const int* anchor_assignment, const float* valid_id, float* eps_out,
                            void* workspace, size_t workspace_bytes, dfb200_stream_t stream);

/* Schedule table: device fp32 array [DFB200_SCHED_ROWS][T], each row float32(np.float64 table)
 * exactly as python/difffacto/models/diffusions/anchored_diffusion.py:62-112 builds them. */
#define DFB200_SCHED_SQRT_ALPHAS_CUMPROD 0
#define DFB200_SCHED_SQRT_ONE_MINUS_ALPHAS_CUMPROD 1
#define DFB200_SCHED_SQRT_RECIP_ALPHAS_CUMPROD 2
#define DFB200_SCHED_SQRT_RECIPM1_ALPHAS_CUMPROD 3
#define DFB200_SCHED_POSTERIOR_VARIANCE 4
#define DFB200_SCHED_POSTERIOR_MEAN_COEF1 5
#define DFB200_SCHED_POSTERIOR_MEAN_COEF2 6
#define DFB200_SCHED_POSTERIOR_MEAN_COEF3 7
#define DFB200_SCHED_ROWS 8

/* One reverse step given eps (epsilon / fixed_small / learn_variance / learn_anchor path):
 *   x0   = sqrt_recip[t]*(x_t-a)+a - sqrt_recipm1[t]*sqrt(var)*eps   (anchored_diffusion.py:401-409)
 *   mean = c1[t]*x0 + c2[t]*x_t + c3[t]*a                            (:184-188)
 *   x'   = mean + [t!=0]*sqrt(post_var[t]*var)*noise                 (:313, :476-483)
 * t (B) int32; noise may be NULL only if every t == 0; pred_xstart may be NULL.
 * Replaces AnchoredDiffusion.p_mean_variance/p_sample arithmetic, anchored_diffusion.py:227-484. */
int dfb200_ddpm_step(int B, int N, int T, const float* sched, const int* t, const float* x_t,
                     const float* eps, const float* anchors, const float* variance,
                     const float* noise, float* x_prev, float* pred_xstart,
                     dfb200_stream_t stream);
/* DDIM variant of the reverse step (ddim_sampling=True; anchored_diffusion.py:114-126, :368-374, :480-481):
 *   x_prev = (x0 - a)*sqrt(alphas_cumprod_prev[t]) + a + sqrt(var)*xt_dir_coeff[t]*eps + eta*[t!=0]*sqrt(post_var[t]*var)*noise
 * alphas_cumprod_prev / xt_dir_coeff: device float32[T] tables (float32 of the reference's float64 numpy tables);
 * noise may be NULL (treated as 0).  pred_xstart optional. */
int dfb200_ddim_step(int B, int N, int T, const float* sched, const int* t, const float* x_t, const float* eps,
                     const float* anchors, const float* variance, const float* noise,
                     const float* alphas_cumprod_prev, const float* xt_dir_coeff, float eta, float* x_prev,
                     float* pred_xstart, dfb200_stream_t stream);
/* Classifier-free guidance mix, eps_out = (1 - w) * eps_uncond + w * eps_cond (anchored_diffusion.py:263-266). */
int dfb200_guidance_mix(size_t count, float classifier_weight, const float* eps_uncond, const float* eps_cond,
                        float* eps_out, dfb200_stream_t stream);
/* x_t = sqrt_ac[t]*(x0-a)+a + sqrt_1mac[t]*sqrt(var)*noise.  Replaces q_sample, :148-173. */
int dfb200_q_sample(int B, int N, int T, const float* sched, const int* t, const float* x_start,
                    const float* anchors, const float* variance, const float* noise, float* x_t,
                    dfb200_stream_t stream);
/* Fill `out` (count floats) with N(0,1) samples from the library's counter-based Philox4x32-10
 * stream (seed, offset); the same generator the sampling loop uses when `noise == NULL`. */
int dfb200_philox_normal(float* out, size_t count, uint64_t seed, uint64_t offset,
                         dfb200_stream_t stream);

/* ---- training-side primitives (fp32), composed by difffacto_b200/train_ops.py into the differentiable training
 * forward/backward of TransformerNet (reference nets/attention.py:50-57, 77-94, 161-204, 259-306, 385-440 run under
 * torch autograd; anchored_diffusion.py:760-852 training_losses).  All pointers are device pointers, row-major. ---- */
/* C[M,N] = (beta ? C : 0) + bias[j] + sum_k A(i,k) B(k,j).  a_k_contiguous: A(i,k) = A[i*lda + k], else A[k*lda + i];
 * b_k_contiguous: B(k,j) = B[j*ldb + k] (an nn.Linear weight), else B[k*ldb + j].  split_k > 1 reduces K in `split_k`
 * slices with atomicAdd (C must hold its initial value).  bias may be NULL. */
int dfb200_sgemm(int a_k_contiguous, int b_k_contiguous, int M, int N, int K, const float* A, int lda, const float* B,
                 int ldb, float* C, int ldc, const float* bias, int beta, int split_k, dfb200_stream_t stream);
/* The same contract on the tensor cores: operands rounded to bf16 on the fly, fp32 accumulation in tensor memory
 * (tcgen05, 128x128 tiles).  Used for the large Linear layers of the training path when precision = "bf16". */
int dfb200_gemm_bf16(int a_k_contiguous, int b_k_contiguous, int M, int N, int K, const float* A, int lda, const float* B,
                     int ldb, float* C, int ldc, const float* bias, int beta, int split_k, dfb200_stream_t stream);
/* out[j] += sum_i X[i*ld + j] (bias gradients; `out` accumulates). */
int dfb200_colsum_accumulate(long long M, int N, const float* X, int ld, float* out, dfb200_stream_t stream);
/* nn.LayerNorm(128, eps=1e-5) over M rows; mean/rstd (M) saved for the backward.  backward: dx written, dgamma/dbeta
 * accumulated (atomicAdd) into caller-zeroed (or running) buffers. */
int dfb200_layernorm128_forward(long long M, const float* x, const float* gamma, const float* beta, float* y,
                                float* mean, float* rstd, dfb200_stream_t stream);
int dfb200_layernorm128_backward(long long M, const float* x, const float* gamma, const float* mean, const float* rstd,
                                 const float* dy, float* dx, float* dgamma_accum, float* dbeta_accum,
                                 dfb200_stream_t stream);
/* The same with dx += dres (M x 128, may be NULL): the gradient of the residual connection around the LayerNorm'd branch
 * (attention.py:296-306: x = attn(norm(x)) + x) joins the LayerNorm backward instead of a separate accumulation pass. */
int dfb200_layernorm128_backward_residual(long long M, const float* x, const float* gamma, const float* mean, const float* rstd,
                                          const float* dy, const float* dres, float* dx, float* dgamma_accum, float* dbeta_accum,
                                          dfb200_stream_t stream);
/* GEGLU (attention.py:50-57): h (M, 2H) = [a | g] -> u (M, H) = a * gelu_erf(g); backward dh (M, 2H). */
int dfb200_geglu_forward(long long M, int H, const float* h, float* u, dfb200_stream_t stream);
int dfb200_geglu_backward(long long M, int H, const float* h, const float* du, float* dh, dfb200_stream_t stream);
/* Cross-attention core over the 4 part tokens (attention.py:183-203): q/o (B*N,128) in 8 heads x 16, k/v (B,4,128),
 * valid_id (B,4) or NULL, probs (B*N,8,4) saved for the backward.  backward: dq written, dk/dv (B,4,128) accumulated. */
int dfb200_part_attention_forward(int B, int N, const float* q, const float* k, const float* v, const float* valid_id,
                                  float* o, float* probs, dfb200_stream_t stream);
int dfb200_part_attention_backward(int B, int N, const float* q, const float* k, const float* v, const float* valid_id,
                                   const float* probs, const float* d_o, float* dq, float* dk_accum, float* dv_accum,
                                   dfb200_stream_t stream);
/* timestep_embedding (nets/utils.py:7-24): out (B,256) = [cos(t f) | sin(t f)], f = the 128 frequencies. */
int dfb200_timestep_embedding(int B, const float* t, const float* freqs128, float* out, dfb200_stream_t stream);
/* FeedForward's first half in ONE kernel (attention.py:77-94: Linear(dim, 2H) -> GEGLU -> Dropout): h = x W1^T + b1 (M x 2H, kept for
 * the backward) and u = dropout(h[:, :H] * gelu(h[:, H:])) (M x H) from a tcgen05 kind::tf32 GEMM whose epilogue applies the GEGLU and
 * the dropout mask of dfb200_geglu_dropout_forward (same Philox key: seed / offset / *step), so the pre-activation is never read
 * back.  x: M x K (leading dimension ldx), w1: 2H x K (ldw), both k-contiguous; needs H % 64 == 0, ldx % 4 == ldw % 4 == 0 and
 * 16-byte aligned pointers, otherwise DFB200_ERR_UNSUPPORTED (callers then use dfb200_gemm_bf16 + dfb200_geglu_dropout_forward). */
int dfb200_ff_in_forward(long long M, int H, int K, const float* x, int ldx, const float* w1, int ldw, const float* b1, float p,
                         uint64_t seed, uint64_t offset, const unsigned long long* step, float* h, float* u, dfb200_stream_t stream);
/* FeedForward's GEGLU fused with the Dropout behind it (attention.py:77-94): u = dropout(a * gelu(g)) for h = [a | g] (M, 2H), with
 * the mask dfb200_dropout(_stepped) would draw on u (p = 0: none; step may be NULL).  The backward also accumulates the column
 * sums of dh (M, 2H) into db_accum (2H floats, may be NULL): the bias gradient of the Linear that produced h. */
int dfb200_geglu_dropout_forward(long long M, int H, float p, uint64_t seed, uint64_t offset, const unsigned long long* step,
                                 const float* h, float* u, dfb200_stream_t stream);
int dfb200_geglu_dropout_backward(long long M, int H, float p, uint64_t seed, uint64_t offset, const unsigned long long* step,
                                  const float* h, const float* du, float* dh, float* db_accum, dfb200_stream_t stream);
/* Inverted dropout with a Philox mask keyed by (seed, offset): y = (keep ? x/(1-p) : 0) + residual (residual may be
 * NULL).  The backward pass applies the same call (without residual) to the incoming gradient. */
int dfb200_dropout(size_t count, float p, uint64_t seed, uint64_t offset, const float* x, const float* residual, float* y,
                   dfb200_stream_t stream);
/* The same with the mask additionally keyed by a DEVICE-resident step counter (*step is read by the kernel): a training step
 * captured in a CUDA graph draws a fresh mask on every replay when the graph increments the counter (train_graph.py). */
int dfb200_dropout_stepped(size_t count, float p, uint64_t seed, uint64_t offset, const unsigned long long* step, const float* x,
                           const float* residual, float* y, dfb200_stream_t stream);
/* torch.optim.Adam (no amsgrad; L2 weight decay) for a whole parameter group in ONE launch per 320 tensors.  `tensors` is a HOST array;
 * every pointer in it is a device pointer to `count` fp32 elements.  *step (device) is the 1-based step count of THIS update (the
 * caller increments it before the call - it lives on the device so that the update can be captured in a CUDA graph); the gradients
 * are multiplied by *grad_scale when it is non-NULL (gradient clipping coefficient).  Replaces the optimizer the reference builds
 * from cfg.optimizer (python/difffacto/runner/runner.py:60-66, torch.optim.Adam) on the training hot path. */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  long long count;
} dfb200_adam_tensor_t;
int dfb200_adam_step(int n_tensors, const dfb200_adam_tensor_t* tensors, const long long* step, float lr, float beta1, float beta2, float eps,
                     float weight_decay, const float* grad_scale, dfb200_stream_t stream);
/* Backward of dfb200_q_sample: any of the three outputs may be NULL. */
int dfb200_q_sample_backward(int B, int N, int T, const float* sched, const int* t, const float* variance,
                             const float* noise, const float* grad_x_t, float* grad_x_start, float* grad_anchors,
                             float* grad_variance, dfb200_stream_t stream);

/* ---- training-side encoder primitives (row f3): PointNetV2 (models/encoders/pointnet.py:122-214) and the forward direction
 * of the latent flows (encoders/flow.py:24-45), composed by difffacto_b200/train_ops.py ---- */
/* nn.BatchNorm1d in training mode over the M rows of a row-major (M, C) matrix (+ fused ReLU): batch statistics (saved in
 * mean / rstd for the backward), running-statistics update (momentum; pointers may be NULL).  scratch2C: 2*C floats. */
int dfb200_batchnorm_forward(long long M, int C, int relu, const float* x, const float* gamma, const float* beta, float* y,
                             float* mean, float* rstd, float* running_mean, float* running_var, float momentum,
                             float* scratch2C, dfb200_stream_t stream);
/* y = (x - mean) * rstd * gamma + beta (+ ReLU) with caller-supplied statistics (eval mode). */
int dfb200_batchnorm_apply(long long M, int C, int relu, const float* x, const float* mean, const float* rstd, const float* gamma,
                           const float* beta, float* y, dfb200_stream_t stream);
int dfb200_batchnorm_backward(long long M, int C, int relu, const float* x, const float* y, const float* dy, const float* gamma,
                              const float* mean, const float* rstd, float* dx, float* dgamma, float* dbeta, dfb200_stream_t stream);
int dfb200_relu_backward(size_t count, const float* y, const float* dy, float* dx, dfb200_stream_t stream);
/* Anchor-weighted max-pool of PointNetV2 (pointnet.py:194-198) without the (B,C,N,A) intermediate:
 * out[b,c,a] = max_n x[b,n,c] * w[b,n,a] * scale, arg = winning n.  x (B,N,C), w (B,N,A), A <= 4.  backward: dx (zeroed by the
 * caller) += w * scale * dout at the winners. */
int dfb200_weighted_maxpool_forward(int B, int N, int C, int A, float scale, const float* x, const float* w, float* out, int* arg,
                                    dfb200_stream_t stream);
int dfb200_weighted_maxpool_backward(int B, int N, int C, int A, float scale, const float* w, const int* arg, const float* dout,
                                     float* dx_zeroed, dfb200_stream_t stream);
/* CouplingLayer forward (flow.py:33-37): y1 = x2 * sigmoid(s + 2) + shift, logdet[r] = sum log sigmoid(s + 2); s_t (B, 2d);
 * x2 / y1 are d columns with leading dimensions ldx / ldy.  backward: ds_t (B,2d), dx2 (ld lddx) from dy1 and dlogdet (B). */
int dfb200_coupling_forward(int B, int d, const float* s_t, const float* x2, int ldx, float* y1, int ldy, float* logdet,
                            dfb200_stream_t stream);
int dfb200_coupling_backward(int B, int d, const float* s_t, const float* x2, int ldx, const float* dy1, int ldy,
                             const float* dlogdet, float* ds_t, float* dx2, int lddx, dfb200_stream_t stream);

/* ---- encoder side of sampling (PartEncoder.sample_latents, part_encoders.py:1052-1110): forward-only helpers used with
 * dfb200_sgemm / dfb200_geglu_forward / dfb200_gather_points by difffacto_b200/models/encoders/part_encoders.py ---- */
/* nn.LayerNorm(D, eps 1e-5) over M rows, any D. */
int dfb200_layernorm_forward(long long M, int D, const float* x, const float* gamma, const float* beta, float* y,
                             dfb200_stream_t stream);
int dfb200_relu(size_t count, float* x_inplace, dfb200_stream_t stream);
int dfb200_scale(size_t count, float alpha, const float* x, float* y, dfb200_stream_t stream);        /* y = alpha x */
int dfb200_exp_shift(size_t count, float shift, const float* x, float* y, dfb200_stream_t stream);    /* y = exp(x + shift) */
/* CouplingLayer reverse (encoders/flow.py:24-45): target[:, :d] (leading dimension ld, in place) =
 * (target - s_t[:, d:]) / sigmoid(s_t[:, :d] + 2);  s_t (B, 2d). */
int dfb200_coupling_reverse(int B, int d, const float* s_t, float* target, int ld, dfb200_stream_t stream);
/* Multi-head self-attention among n_tok <= 8 tokens (attention.py:179-204 with context = x): q/k/v/out
 * (Bt, n_tok, heads*d_head), valid (Bt, n_tok) masks keys, may be NULL. */
int dfb200_token_attention(int Bt, int n_tok, int heads, int d_head, const float* q, const float* k, const float* v,
                           const float* valid, float* out, dfb200_stream_t stream);

/* Scratch bytes for dfb200_ddpm_sample_loop / dfb200_sample_loop (one size serves every option set). */
size_t dfb200_ddpm_sample_loop_workspace_bytes(const dfb200_denoiser_cfg* cfg, int mode, int B,
                                               int N, int T);
/* Full reverse process x_T -> x_0 for a batch: T x (denoiser + fused eps->x_{t-1} update).
 *   x          : (B,3,N) in: x_T if x_T_from_noise == 0; N(0,1) noise z that is turned into
 *                sqrt(var)*z + anchors (anchored_diffusion.py:564) if == 1; ignored if == 2 (z is
 *                Philox(seed) draw number T); out: x_0.
 *   noise      : (T,B,3,N) per-step N(0,1) noise, index [T-1-i] for step i counting down from
 *                T-1 (step order of the loop), or NULL to use Philox(seed).
 *   traj       : optional (T / traj_interval, B,3,N) buffer; x after step t is stored for every t>0 with
 *                t % traj_interval == 0 at slot t/traj_interval - 1, and x_T itself in the last slot when
 *                T % traj_interval == 0 (AnchorDiffAE.decode's ret_traj/ret_interval keys, T included,
 *                python/difffacto/models/networks/anchor_gen.py:160-167).
 * Replaces AnchoredDiffusion.p_sample_loop_progressive, anchored_diffusion.py:528-588. */
int dfb200_ddpm_sample_loop(const dfb200_denoiser_cfg* cfg, const void* packed, int mode, int B,
                            int N, int T, const float* sched, float* x, int x_T_from_noise,
                            const float* ctx, const float* anchors, const float* variance,
                            const int* anchor_assignment, const float* valid_id,
                            const float* noise, uint64_t seed, float* traj, int traj_interval,
                            void* workspace, size_t workspace_bytes, dfb200_stream_t stream);

/* Variants of the loop, all inside the same persistent fused kernel in bf16 mode (all-zero / NULL = the call above):
 *   timesteps      strided step lists (DDIM, anchored_diffusion.py:114-126): DEVICE int32[n_timesteps], strictly
 *                  decreasing, execution order, plus the same values on the HOST (the launch sequence is host-driven);
 *   first_step /   run only steps [first_step, first_step + num_steps) of the list (num_steps 0 = to the end): a caller
 *   num_steps      that consumes the loop as a generator (AnchorDiffAE.decode over p_sample_loop_progressive) asks for
 *                  it chunk by chunk; `noise`, `step_sample`, `step_xstart` then index from this call's first step.
 *                  The x_T initialisation, the trajectory's x_T slot and the loop tables belong to the call with
 *                  first_step == 0; later calls on the SAME workspace set tables_ready = 1;
 *   ddim           the DDIM update (:368-374, :480-481) with eta and the two float32[T] DEVICE tables
 *                  float32(alphas_cumprod_prev), float32(sqrt(1 - ac - eta^2 posterior_variance));
 *   guidance       classifier-free guidance (:263-266): a second, zero-context denoiser pass per step,
 *                  eps = (1 - w) eps_uncond + w eps_cond;
 *   step_sample /  optional (num_steps,B,3,N) outputs: `sample` and `pred_xstart` of EVERY executed step (the dict the
 *   step_xstart    reference's generator yields, :587);
 *   step_sample_list / step_xstart_list   the same outputs as HOST arrays of num_steps separate (B,3,N) device buffers (entries
 *                  may be NULL), so that a caller who keeps only some steps does not pin the others' memory. */
typedef struct dfb200_sample_opts {
  const int* timesteps;
  const int* timesteps_host;
  int n_timesteps;
  int first_step, num_steps, tables_ready;
  int ddim;
  float ddim_eta;
  const float* alphas_cumprod_prev;
  const float* xt_dir_coeff;
  int guidance;
  float classifier_weight;
  float* step_sample;
  float* step_xstart;
  float* const* step_sample_list;
  float* const* step_xstart_list;
} dfb200_sample_opts;
/* Steps one persistent launch of the fused kernel covers for these sizes: the natural num_steps of a chunked caller. */
int dfb200_sample_loop_chunk(const dfb200_denoiser_cfg* cfg, int mode, int B, int N, int T);
int dfb200_sample_loop(const dfb200_denoiser_cfg* cfg, const void* packed, int mode, int B, int N, int T,
                       const float* sched, float* x, int x_T_from_noise, const float* ctx, const float* anchors,
                       const float* variance, const int* anchor_assignment, const float* valid_id,
                       const float* noise, uint64_t seed, float* traj, int traj_interval,
                       const dfb200_sample_opts* opts, void* workspace, size_t workspace_bytes,
                       dfb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DIFFFACTO_B200_H_ */
