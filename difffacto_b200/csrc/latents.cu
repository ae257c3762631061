// Encoder side of sampling: the small forward-only kernels that, together with dfb200_sgemm /
// dfb200_geglu_forward (train_ops.cu) and dfb200_gather_points (pointnet2.cu), run PartEncoder.sample_latents
// (reference python/difffacto/models/encoders/part_encoders.py:1052-1110): 4 latent coupling flows in reverse
// (encoders/flow.py:7-78), PartAlignerTransformer (part_encoders.py:20-143; 5 self-attention blocks over the 4 part
// tokens, inner width 256) and prepare_ctx (:1317-1327).  Everything here is per-batch work done once before the
// 1000-step loop: tiny tensors, launch-latency bound; the kernels are simple on purpose.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace dfb200 {

// nn.LayerNorm(D, eps=1e-5) over M rows, any D: warp per row
__global__ void __launch_bounds__(256)
ln_generic_kernel(long long M, int D, const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                  float* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const float* xr = x + row * D;
  float s = 0.f;
  for (int i = lane; i < D; i += 32) s += __ldg(xr + i);
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
  const float mean = s / D;
  float q = 0.f;
  for (int i = lane; i < D; i += 32) { const float t = __ldg(xr + i) - mean; q = fmaf(t, t, q); }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) q += __shfl_xor_sync(0xFFFFFFFFu, q, d);
  const float rstd = rsqrtf(q / D + 1e-5f);
  for (int i = lane; i < D; i += 32) y[row * D + i] = (__ldg(xr + i) - mean) * rstd * __ldg(g + i) + __ldg(b + i);
}

__global__ void __launch_bounds__(256) relu_kernel(long long n, float* __restrict__ x) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = fmaxf(x[i], 0.f);
}
// y = alpha * x
__global__ void __launch_bounds__(256) scale_kernel(long long n, float alpha, const float* __restrict__ x, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = alpha * __ldg(x + i);
}
// y = exp(x + shift)
__global__ void __launch_bounds__(256) exp_shift_kernel(long long n, float shift, const float* __restrict__ x, float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = expf(__ldg(x + i) + shift);
}

// CouplingLayer reverse (flow.py:24-45): scale = sigmoid(s_t[:, :d] + 2), shift = s_t[:, d:], target = (target - shift) / scale.
// target: d columns of a row-major matrix with leading dimension ld (updated in place).
__global__ void __launch_bounds__(256)
coupling_reverse_kernel(int B, int d, const float* __restrict__ s_t, float* __restrict__ target, int ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * d) return;
  const int r = i / d, c = i - r * d;
  const float scale = 1.f / (1.f + expf(-(__ldg(s_t + (size_t)r * 2 * d + c) + 2.f)));
  const float shift = __ldg(s_t + (size_t)r * 2 * d + d + c);
  float* t = target + (size_t)r * ld + c;
  *t = (*t - shift) / scale;
}

// Multi-head attention among a handful of tokens (CrossAttention.forward with context = x, attention.py:179-204):
// q/k/v/out (Bt, n_tok, heads*d_head); valid (Bt, n_tok) masks KEYS (masked_fill(~mask, -finfo.max)); thread per
// (sample, head, query token).
__global__ void __launch_bounds__(128)
token_attention_kernel(int Bt, int n_tok, int heads, int d_head, const float* __restrict__ q, const float* __restrict__ k,
                       const float* __restrict__ v, const float* __restrict__ valid, float* __restrict__ out) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= Bt * heads * n_tok) return;
  const int i = id % n_tok, h = (id / n_tok) % heads, b = id / (n_tok * heads);
  const int inner = heads * d_head;
  const float scale = rsqrtf((float)d_head);
  const float* qi = q + ((size_t)b * n_tok + i) * inner + h * d_head;
  float s[8];
  float mx = -FLT_MAX;
  for (int j = 0; j < n_tok; ++j) {
    const float* kj = k + ((size_t)b * n_tok + j) * inner + h * d_head;
    float d = 0.f;
    for (int c = 0; c < d_head; ++c) d = fmaf(__ldg(qi + c), __ldg(kj + c), d);
    d *= scale;
    if (valid != nullptr && __ldg(valid + b * n_tok + j) == 0.f) d = -FLT_MAX;
    s[j] = d;
    mx = fmaxf(mx, d);
  }
  float sum = 0.f;
  for (int j = 0; j < n_tok; ++j) { s[j] = expf(s[j] - mx); sum += s[j]; }
  const float inv = 1.f / sum;
  float* o = out + ((size_t)b * n_tok + i) * inner + h * d_head;
  for (int c = 0; c < d_head; ++c) {
    float a = 0.f;
    for (int j = 0; j < n_tok; ++j) a = fmaf(s[j] * inv, __ldg(v + ((size_t)b * n_tok + j) * inner + h * d_head + c), a);
    o[c] = a;
  }
}

}  // namespace dfb200

using namespace dfb200;

extern "C" int dfb200_layernorm_forward(long long M, int D, const float* x, const float* gamma, const float* beta, float* y,
                                        dfb200_stream_t stream) {
  if (M <= 0 || D <= 0) return DFB200_OK;
  ln_generic_kernel<<<(unsigned)cdiv(M, 8LL), 256, 0, as_stream(stream)>>>(M, D, x, gamma, beta, y);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}
extern "C" int dfb200_relu(size_t count, float* x, dfb200_stream_t stream) {
  if (count == 0) return DFB200_OK;
  relu_kernel<<<(unsigned)cdiv((long long)count, 256LL), 256, 0, as_stream(stream)>>>((long long)count, x);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}
extern "C" int dfb200_scale(size_t count, float alpha, const float* x, float* y, dfb200_stream_t stream) {
  if (count == 0) return DFB200_OK;
  scale_kernel<<<(unsigned)cdiv((long long)count, 256LL), 256, 0, as_stream(stream)>>>((long long)count, alpha, x, y);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}
extern "C" int dfb200_exp_shift(size_t count, float shift, const float* x, float* y, dfb200_stream_t stream) {
  if (count == 0) return DFB200_OK;
  exp_shift_kernel<<<(unsigned)cdiv((long long)count, 256LL), 256, 0, as_stream(stream)>>>((long long)count, shift, x, y);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}
extern "C" int dfb200_coupling_reverse(int B, int d, const float* s_t, float* target, int ld, dfb200_stream_t stream) {
  DFB_REQUIRE(B >= 0 && d >= 0 && ld >= d, DFB200_ERR_INVALID_ARG, "coupling_reverse: bad sizes B=%d d=%d ld=%d", B, d, ld);
  if (B == 0 || d == 0) return DFB200_OK;
  coupling_reverse_kernel<<<cdiv(B * d, 256), 256, 0, as_stream(stream)>>>(B, d, s_t, target, ld);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}
extern "C" int dfb200_token_attention(int Bt, int n_tok, int heads, int d_head, const float* q, const float* k, const float* v,
                                      const float* valid, float* out, dfb200_stream_t stream) {
  DFB_REQUIRE(Bt >= 0 && n_tok >= 1 && n_tok <= 8 && heads >= 1 && d_head >= 1, DFB200_ERR_INVALID_ARG,
              "token_attention: n_tok must be in [1, 8] (got %d)", n_tok);
  if (Bt == 0) return DFB200_OK;
  token_attention_kernel<<<cdiv(Bt * heads * n_tok, 128), 128, 0, as_stream(stream)>>>(Bt, n_tok, heads, d_head, q, k, v, valid, out);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}
