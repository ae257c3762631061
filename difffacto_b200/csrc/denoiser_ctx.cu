// Denoiser plumbing: config validation, packed weight image, workspace carving, and the per-call
// context path (timestep embedding MLP -> K/V of the 4 part tokens for every block).
// Reference: python/difffacto/models/diffusions/nets/attention.py:385-398 (context assembly),
// :50-57,77-94 (time_embed GEGLU FeedForward), :184-185 (to_k/to_v), nets/utils.py:7-24 (sinusoid).
#include <math.h>
#include <vector>

#include "denoiser.cuh"

namespace dfb200 {

int make_net_dims(const dfb200_denoiser_cfg* cfg, NetDims* d) {
  DFB_REQUIRE(cfg != nullptr, DFB200_ERR_INVALID_ARG, "denoiser: cfg is NULL");
  DFB_REQUIRE(cfg->n_heads * cfg->d_head == D_MODEL, DFB200_ERR_UNSUPPORTED,
              "denoiser: inner_dim n_heads*d_head must be 128 (got %d*%d)", cfg->n_heads, cfg->d_head);
  DFB_REQUIRE(cfg->d_head == 16 && cfg->n_heads == 8, DFB200_ERR_UNSUPPORTED, "denoiser: only 8 heads x 16 supported");
  DFB_REQUIRE(cfg->depth >= 1 && cfg->depth <= MAX_DEPTH, DFB200_ERR_UNSUPPORTED, "denoiser: depth %d out of [1,%d]", cfg->depth, MAX_DEPTH);
  DFB_REQUIRE(cfg->n_class == MAX_TOKENS, DFB200_ERR_UNSUPPORTED, "denoiser: n_class must be 4 (got %d)", cfg->n_class);
  DFB_REQUIRE(cfg->in_channels == 3 && cfg->out_channels == 3, DFB200_ERR_UNSUPPORTED, "denoiser: in/out channels must be 3");
  DFB_REQUIRE(cfg->context_dim >= 1 && cfg->context_dim <= 4096, DFB200_ERR_INVALID_ARG, "denoiser: bad context_dim %d", cfg->context_dim);
  const int need = DFB200_NET_CLASS_COND | DFB200_NET_CAT_PARAMS_TO_X | DFB200_NET_CAT_CLASS_TO_X;
  DFB_REQUIRE((cfg->flags & need) == need, DFB200_ERR_UNSUPPORTED,
              "denoiser: class_cond, cat_params_to_x and cat_class_to_x must all be set (flags=0x%x)", cfg->flags);
  d->c_in = cfg->in_channels + 6 + cfg->n_class;
  d->c_out = cfg->out_channels;
  d->c_ctx_static = cfg->context_dim + cfg->n_class;
  d->c_ctx = d->c_ctx_static + D_TEMB;
  d->depth = cfg->depth;
  d->n_tok = cfg->n_class;
  d->n_heads = cfg->n_heads;
  d->d_head = cfg->d_head;
  d->flags = cfg->flags;
  return DFB200_OK;
}

size_t param_numel(const NetDims& d, bool global, int which) {
  if (global) {
    switch (which) {
      case P_PRE_W: case P_PRE_B: case P_POST_W: case P_POST_B: case P_IN_B: return D_MODEL;
      case P_IN_W: return (size_t)D_MODEL * d.c_in;
      case P_TE0_W: return (size_t)2 * D_TEMB_H * D_TEMB;
      case P_TE0_B: return 2 * D_TEMB_H;
      case P_TE2_W: return (size_t)D_TEMB * D_TEMB_H;
      case P_TE2_B: return D_TEMB;
      case P_OUT_W: return (size_t)d.c_out * D_MODEL;
      case P_OUT_B: return d.c_out;
    }
  } else {
    switch (which) {
      case B_N2_W: case B_N2_B: case B_N3_W: case B_N3_B: case B_BO: case B_B2: return D_MODEL;
      case B_WQ: case B_WO: return (size_t)D_MODEL * D_MODEL;
      case B_WK: case B_WV: return (size_t)D_MODEL * d.c_ctx;
      case B_W1: return (size_t)2 * D_FF * D_MODEL;
      case B_B1: return 2 * D_FF;
      case B_W2: return (size_t)D_MODEL * D_FF;
    }
  }
  return 0;
}

size_t tc_stream_bytes_for(const NetDims& d);    // denoiser_tc.cu
size_t tf32_stream_bytes_for(const NetDims& d);  // denoiser_tf32.cu
int tf32_pack_stream(const PackLayout& L, void* packed, cudaStream_t st);

int make_pack_layout(const dfb200_denoiser_cfg* cfg, PackLayout* L) {
  int rc = make_net_dims(cfg, &L->d);
  if (rc != DFB200_OK) return rc;
  size_t off = 0;
  auto take = [&off](size_t n) { size_t o = off; off += (n + 3) & ~(size_t)3; return o; };  // 16 B aligned
  for (int i = 0; i < N_GLOBAL_PARAMS; ++i) L->g[i] = take(param_numel(L->d, true, i));
  for (int l = 0; l < L->d.depth; ++l)
    for (int i = 0; i < N_BLOCK_PARAMS; ++i) L->blk[l][i] = take(param_numel(L->d, false, i));
  L->freqs = take(D_TEMB / 2);
  L->fp32_floats = off;
  L->tc_stream_off = (off * sizeof(float) + 1023) & ~(size_t)1023;
  L->tc_stream_bytes = tc_stream_bytes_for(L->d);
  L->tf32_stream_off = (L->tc_stream_off + L->tc_stream_bytes + 1023) & ~(size_t)1023;
  L->tf32_stream_bytes = tf32_stream_bytes_for(L->d);
  L->total_bytes = L->tf32_stream_off + L->tf32_stream_bytes;
  return DFB200_OK;
}

Workspace carve_workspace(const NetDims& d, int mode, int B, int N, void* base) {
  Workspace w{};
  size_t off = 0;
  auto take = [&](size_t floats) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += ((floats * sizeof(float)) + 255) & ~(size_t)255;
    return p;
  };
  const size_t M = (size_t)B * N;
  w.temb_h = take((size_t)B * D_TEMB_H);
  w.temb = take((size_t)B * D_TEMB);
  w.kv = take((size_t)B * d.depth * 2 * d.n_tok * D_MODEL);
  if (mode == DFB200_MODE_FP32) {
    w.x = take(M * D_MODEL);
    w.q = take(M * D_MODEL);
    w.u = take(M * D_FF);
  } else if (mode == DFB200_MODE_TF32) {
    w.fold = take((tf32_fold_bytes_for(d, B) + 3) / 4);
  } else {
    w.fold = take((tc_fold_bytes_for(d, B) + 3) / 4);
  }
  w.bytes = off;
  return w;
}

// ---------------------------------------------------------------------------------------------
// time_embed: emb = [cos(t f) | sin(t f)] -> GEGLU(256 -> 2*1024) -> Linear(1024 -> 256)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float gelu_erf(float x) {  // F.gelu default (exact erf form)
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
}

// grid (B, 1024/64): each CTA produces 64 gated hidden units; one warp per unit pair (value, gate)
__global__ void __launch_bounds__(256)
temb_hidden_kernel(const float* __restrict__ t, const float* __restrict__ freqs,
                   const float* __restrict__ w0, const float* __restrict__ b0, float* __restrict__ hidden) {
  __shared__ float emb[D_TEMB];
  const int b = blockIdx.x;
  const float tv = __ldg(t + b);
  for (int i = threadIdx.x; i < D_TEMB / 2; i += blockDim.x) {
    const float a = tv * __ldg(freqs + i);  // args = t[:,None] * freqs[None]   (nets/utils.py:20)
    emb[i] = cosf(a);
    emb[i + D_TEMB / 2] = sinf(a);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int u = warp; u < 64; u += 8) {
    const int i = blockIdx.y * 64 + u;
    const float* wa = w0 + (size_t)i * D_TEMB;
    const float* wg = w0 + (size_t)(i + D_TEMB_H) * D_TEMB;
    float sa = 0.f, sg = 0.f;
#pragma unroll
    for (int k = lane * 4; k < D_TEMB; k += 128) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(wa + k));
      const float4 g = __ldg(reinterpret_cast<const float4*>(wg + k));
      sa += a.x * emb[k] + a.y * emb[k + 1] + a.z * emb[k + 2] + a.w * emb[k + 3];
      sg += g.x * emb[k] + g.y * emb[k + 1] + g.z * emb[k + 2] + g.w * emb[k + 3];
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      sa += __shfl_xor_sync(0xFFFFFFFFu, sa, d);
      sg += __shfl_xor_sync(0xFFFFFFFFu, sg, d);
    }
    if (lane == 0) hidden[(size_t)b * D_TEMB_H + i] = (sa + __ldg(b0 + i)) * gelu_erf(sg + __ldg(b0 + i + D_TEMB_H));
  }
}

// grid (B, 256/32): warp per 4 outputs
__global__ void __launch_bounds__(256)
temb_out_kernel(const float* __restrict__ hidden, const float* __restrict__ w2, const float* __restrict__ b2,
                float* __restrict__ temb) {
  __shared__ float h[D_TEMB_H];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < D_TEMB_H; i += blockDim.x) h[i] = hidden[(size_t)b * D_TEMB_H + i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int u = warp; u < 32; u += 8) {
    const int o = blockIdx.y * 32 + u;
    const float* w = w2 + (size_t)o * D_TEMB_H;
    float s = 0.f;
#pragma unroll
    for (int k = lane * 4; k < D_TEMB_H; k += 128) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(w + k));
      s += a.x * h[k] + a.y * h[k + 1] + a.z * h[k + 2] + a.w * h[k + 3];
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
    if (lane == 0) temb[(size_t)b * D_TEMB + o] = s + __ldg(b2 + o);
  }
}

// grid (B, depth*2): CTA = one of {K,V} of one block for one sample: 128 outputs x 4 tokens, K = c_ctx.
struct KvArgs {
  const float* w[MAX_DEPTH][2];
};
__global__ void __launch_bounds__(256)
context_kv_kernel(KvArgs args, int c_ctx, int c_raw, int n_tok, const float* __restrict__ ctx,
                  const float* __restrict__ temb, float* __restrict__ kv) {
  extern __shared__ float cs[];  // [n_tok][c_ctx]
  const int b = blockIdx.x;
  const int l = blockIdx.y >> 1, which = blockIdx.y & 1;
  const int c_static = c_raw + n_tok;
  for (int i = threadIdx.x; i < n_tok * c_ctx; i += blockDim.x) {
    const int j = i / c_ctx, k = i - j * c_ctx;
    float v;
    if (k < c_raw) v = __ldg(ctx + ((size_t)b * c_raw + k) * n_tok + j);  // ctx is (B, C, n_tok)
    else if (k < c_static) v = (k - c_raw) == j ? 1.f : 0.f;              // torch.eye(n_class) class embed
    else v = temb[(size_t)b * D_TEMB + (k - c_static)];
    cs[i] = v;
  }
  __syncthreads();
  const float* W = args.w[l][which];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* out = kv + (((size_t)b * gridDim.y + blockIdx.y) * n_tok) * D_MODEL;
  for (int o = warp; o < D_MODEL; o += 8) {
    const float* w = W + (size_t)o * c_ctx;
    float s[MAX_TOKENS] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < c_ctx; k += 32) {
      const float wv = __ldg(w + k);
#pragma unroll
      for (int j = 0; j < MAX_TOKENS; ++j) s[j] += wv * cs[j * c_ctx + k];
    }
#pragma unroll
    for (int j = 0; j < MAX_TOKENS; ++j) {
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) s[j] += __shfl_xor_sync(0xFFFFFFFFu, s[j], d);
    }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < MAX_TOKENS; ++j) out[(size_t)j * D_MODEL + o] = s[j];
    }
  }
}

int launch_context_kv(const PackLayout& L, const float* P, int B, const float* t, const float* ctx,
                      Workspace& ws, cudaStream_t st) {
  if (B == 0) return DFB200_OK;
  temb_hidden_kernel<<<dim3(B, D_TEMB_H / 64), 256, 0, st>>>(t, P + L.freqs, P + L.g[P_TE0_W], P + L.g[P_TE0_B], ws.temb_h);
  DFB_LAUNCH_CHECK();
  temb_out_kernel<<<dim3(B, D_TEMB / 32), 256, 0, st>>>(ws.temb_h, P + L.g[P_TE2_W], P + L.g[P_TE2_B], ws.temb);
  DFB_LAUNCH_CHECK();
  KvArgs a{};
  for (int l = 0; l < L.d.depth; ++l) {
    a.w[l][0] = P + L.blk[l][B_WK];
    a.w[l][1] = P + L.blk[l][B_WV];
  }
  const size_t smem = sizeof(float) * L.d.n_tok * L.d.c_ctx;
  DFB_REQUIRE(smem <= 48 * 1024, DFB200_ERR_UNSUPPORTED, "denoiser: context width %d too large", L.d.c_ctx);
  context_kv_kernel<<<dim3(B, L.d.depth * 2), 256, smem, st>>>(a, L.d.c_ctx, L.d.c_ctx_static - L.d.n_tok, L.d.n_tok,
                                                                 ctx, ws.temb, ws.kv);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

// grid (B, depth*2): static half of K/V: 128 outputs x 4 tokens over the first c_static context columns
__global__ void __launch_bounds__(256)
context_kv_static_kernel(KvArgs args, int c_ctx, int c_raw, int n_tok, const float* __restrict__ ctx, float* __restrict__ kv) {
  extern __shared__ float cs[];  // [n_tok][c_static]
  const int b = blockIdx.x;
  const int l = blockIdx.y >> 1, which = blockIdx.y & 1;
  const int c_static = c_raw + n_tok;
  for (int i = threadIdx.x; i < n_tok * c_static; i += blockDim.x) {
    const int j = i / c_static, k = i - j * c_static;
    cs[i] = k < c_raw ? __ldg(ctx + ((size_t)b * c_raw + k) * n_tok + j) : ((k - c_raw) == j ? 1.f : 0.f);
  }
  __syncthreads();
  const float* W = args.w[l][which];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* out = kv + (((size_t)b * gridDim.y + blockIdx.y) * n_tok) * D_MODEL;
  for (int o = warp; o < D_MODEL; o += 8) {
    const float* w = W + (size_t)o * c_ctx;
    float s[MAX_TOKENS] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < c_static; k += 32) {
      const float wv = __ldg(w + k);
#pragma unroll
      for (int j = 0; j < MAX_TOKENS; ++j) s[j] += wv * cs[j * c_static + k];
    }
#pragma unroll
    for (int j = 0; j < MAX_TOKENS; ++j) {
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) s[j] += __shfl_xor_sync(0xFFFFFFFFu, s[j], d);
    }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < MAX_TOKENS; ++j) out[(size_t)j * D_MODEL + o] = s[j];
    }
  }
}

// grid (T, depth*2): time half of K/V: kv_time[t, l, which, :] = W[:, c_static:] . temb[t]
__global__ void __launch_bounds__(256)
context_kv_time_kernel(KvArgs args, int c_ctx, int c_static, const float* __restrict__ temb, float* __restrict__ kv_time) {
  __shared__ float te[D_TEMB];
  const int t = blockIdx.x;
  const int l = blockIdx.y >> 1, which = blockIdx.y & 1;
  for (int i = threadIdx.x; i < D_TEMB; i += blockDim.x) te[i] = temb[(size_t)t * D_TEMB + i];
  __syncthreads();
  const float* W = args.w[l][which];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* out = kv_time + ((size_t)t * gridDim.y + blockIdx.y) * D_MODEL;
  for (int o = warp; o < D_MODEL; o += 8) {
    const float* w = W + (size_t)o * c_ctx + c_static;
    float s = 0.f;
    for (int k = lane; k < D_TEMB; k += 32) s += __ldg(w + k) * te[k];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
    if (lane == 0) out[o] = s;
  }
}

static KvArgs make_kv_args(const PackLayout& L, const float* P) {
  KvArgs a{};
  for (int l = 0; l < L.d.depth; ++l) {
    a.w[l][0] = P + L.blk[l][B_WK];
    a.w[l][1] = P + L.blk[l][B_WV];
  }
  return a;
}

int launch_context_kv_static(const PackLayout& L, const float* P, int B, const float* ctx, float* kv_static, cudaStream_t st) {
  if (B == 0) return DFB200_OK;
  const size_t smem = sizeof(float) * L.d.n_tok * L.d.c_ctx_static;
  DFB_REQUIRE(smem <= 48 * 1024, DFB200_ERR_UNSUPPORTED, "denoiser: context width %d too large", L.d.c_ctx_static);
  context_kv_static_kernel<<<dim3(B, L.d.depth * 2), 256, smem, st>>>(make_kv_args(L, P), L.d.c_ctx, L.d.c_ctx_static - L.d.n_tok,
                                                                        L.d.n_tok, ctx, kv_static);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

int launch_context_kv_time(const PackLayout& L, const float* P, int T, const float* t_values, float* temb_h, float* temb,
                           float* kv_time, cudaStream_t st) {
  if (T == 0) return DFB200_OK;
  temb_hidden_kernel<<<dim3(T, D_TEMB_H / 64), 256, 0, st>>>(t_values, P + L.freqs, P + L.g[P_TE0_W], P + L.g[P_TE0_B], temb_h);
  DFB_LAUNCH_CHECK();
  temb_out_kernel<<<dim3(T, D_TEMB / 32), 256, 0, st>>>(temb_h, P + L.g[P_TE2_W], P + L.g[P_TE2_B], temb);
  DFB_LAUNCH_CHECK();
  context_kv_time_kernel<<<dim3(T, L.d.depth * 2), 256, 0, st>>>(make_kv_args(L, P), L.d.c_ctx, L.d.c_ctx_static, temb, kv_time);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

int tc_pack_stream(const PackLayout& L, void* packed, cudaStream_t st);  // denoiser_tc.cu
int denoiser_forward_tc(const PackLayout& L, const void* packed, int B, int N, const float* x,
                        const float* anchors, const float* variances, const int* assign,
                        const float* valid_id, float* eps_out, Workspace& ws, cudaStream_t st);

}  // namespace dfb200

using namespace dfb200;

extern "C" int dfb200_denoiser_num_params(const dfb200_denoiser_cfg* cfg) {
  NetDims d;
  if (make_net_dims(cfg, &d) != DFB200_OK) return -1;
  return N_GLOBAL_PARAMS + N_BLOCK_PARAMS * d.depth;
}

extern "C" size_t dfb200_denoiser_packed_bytes(const dfb200_denoiser_cfg* cfg) {
  PackLayout L;
  if (make_pack_layout(cfg, &L) != DFB200_OK) return 0;
  return L.total_bytes;
}

extern "C" int dfb200_denoiser_pack(const dfb200_denoiser_cfg* cfg, const float* const* params, int n_params,
                                    void* packed, dfb200_stream_t stream) {
  PackLayout L;
  int rc = make_pack_layout(cfg, &L);
  if (rc != DFB200_OK) return rc;
  DFB_REQUIRE(params != nullptr && packed != nullptr, DFB200_ERR_INVALID_ARG, "denoiser_pack: NULL argument");
  const int n_expected = N_GLOBAL_PARAMS + N_BLOCK_PARAMS * L.d.depth;
  // one optional trailing tensor: the 128 sinusoid frequencies (device fp32), for callers that
  // want them bit-identical to their own host framework's exp()
  DFB_REQUIRE(n_params == n_expected || n_params == n_expected + 1, DFB200_ERR_INVALID_ARG,
              "denoiser_pack: expected %d (+1 optional freqs) parameter tensors, got %d", n_expected, n_params);
  DFB_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 1023) == 0, DFB200_ERR_INVALID_ARG, "denoiser_pack: packed must be 1024 B aligned");
  cudaStream_t st = as_stream(stream);
  float* P = reinterpret_cast<float*>(packed);
  DFB_CUDA(cudaMemsetAsync(packed, 0, L.total_bytes, st));
  int pi = 0;
  for (int i = 0; i < N_GLOBAL_PARAMS; ++i, ++pi)
    DFB_CUDA(cudaMemcpyAsync(P + L.g[i], params[pi], sizeof(float) * param_numel(L.d, true, i), cudaMemcpyDeviceToDevice, st));
  for (int l = 0; l < L.d.depth; ++l)
    for (int i = 0; i < N_BLOCK_PARAMS; ++i, ++pi)
      DFB_CUDA(cudaMemcpyAsync(P + L.blk[l][i], params[pi], sizeof(float) * param_numel(L.d, false, i), cudaMemcpyDeviceToDevice, st));
  // sinusoid frequencies: exp(-ln(10000) * i / 128) evaluated in fp32 on the host, the way the
  // reference builds them on the CPU before moving them to the device (nets/utils.py:17-19)
  if (n_params == n_expected + 1) {
    DFB_CUDA(cudaMemcpyAsync(P + L.freqs, params[n_expected], sizeof(float) * (D_TEMB / 2), cudaMemcpyDeviceToDevice, st));
  } else {
    std::vector<float> fr(D_TEMB / 2);
    for (int i = 0; i < D_TEMB / 2; ++i) fr[i] = expf((-(float)log(10000.0) * (float)i) / (float)(D_TEMB / 2));
    DFB_CUDA(cudaMemcpyAsync(P + L.freqs, fr.data(), sizeof(float) * fr.size(), cudaMemcpyHostToDevice, st));
    DFB_CUDA(cudaStreamSynchronize(st));  // fr is a stack-lifetime host buffer
  }
  rc = tc_pack_stream(L, packed, st);
  if (rc != DFB200_OK) return rc;
  return tf32_pack_stream(L, packed, st);
}

extern "C" size_t dfb200_denoiser_workspace_bytes(const dfb200_denoiser_cfg* cfg, int mode, int B, int N) {
  NetDims d;
  if (make_net_dims(cfg, &d) != DFB200_OK || B < 0 || N < 0) return 0;
  return carve_workspace(d, mode, B, N, nullptr).bytes;
}

extern "C" int dfb200_denoiser_forward(const dfb200_denoiser_cfg* cfg, const void* packed, int mode, int B, int N,
                                       const float* x, const float* t, const float* ctx, const float* anchors,
                                       const float* variances, const int* anchor_assignment, const float* valid_id,
                                       float* eps_out, void* workspace, size_t workspace_bytes, dfb200_stream_t stream) {
  PackLayout L;
  int rc = make_pack_layout(cfg, &L);
  if (rc != DFB200_OK) return rc;
  DFB_REQUIRE(B >= 0 && N >= 0, DFB200_ERR_INVALID_ARG, "denoiser_forward: negative size");
  DFB_REQUIRE(mode == DFB200_MODE_FP32 || mode == DFB200_MODE_BF16 || mode == DFB200_MODE_TF32, DFB200_ERR_INVALID_ARG,
              "denoiser_forward: unknown mode %d", mode);
  if (B == 0 || N == 0) return DFB200_OK;
  Workspace ws = carve_workspace(L.d, mode, B, N, workspace);
  DFB_REQUIRE(workspace != nullptr && workspace_bytes >= ws.bytes, DFB200_ERR_WORKSPACE,
              "denoiser_forward: workspace too small (%zu < %zu)", workspace_bytes, ws.bytes);
  cudaStream_t st = as_stream(stream);
  const float* P = reinterpret_cast<const float*>(packed);
  rc = launch_context_kv(L, P, B, t, ctx, ws, st);
  if (rc != DFB200_OK) return rc;
  const float* valid = (L.d.flags & DFB200_NET_MASK_UNREFERENCED) ? valid_id : nullptr;
  if (mode == DFB200_MODE_FP32)
    return denoiser_forward_fp32(L, P, B, N, x, anchors, variances, anchor_assignment, valid, eps_out, ws, st);
  if (mode == DFB200_MODE_TF32)
    return denoiser_forward_tf32(L, packed, B, N, x, anchors, variances, anchor_assignment, valid, eps_out, ws, st);
  return denoiser_forward_tc(L, packed, B, N, x, anchors, variances, anchor_assignment, valid, eps_out, ws, st);
}
