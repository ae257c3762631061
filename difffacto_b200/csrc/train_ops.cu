#include <mutex>
// Training-side primitives of the cross-diffusion denoiser (fp32, CUDA cores): the differentiable building blocks that
// difffacto_b200/train_ops.py composes - through torch.autograd.Function wrappers - into the training forward/backward of
// TransformerNet (reference: python/difffacto/models/diffusions/nets/attention.py:50-57 GEGLU, :77-94 FeedForward,
// :161-204 CrossAttention, :259-306 BasicTransformerBlock, :385-440 TransformerNet; loss: anchored_diffusion.py:760-852).
// The sampling path never touches this file; it exists so that `training_losses(...).backward()` runs on this repo's own
// kernels.  One strided SGEMM serves every Linear (forward, dgrad, wgrad with split-K).
#include <float.h>
#include <math.h>

#include "denoiser.cuh"
#include "philox.cuh"
#include "geglu_math.cuh"

namespace dfb200 {

// ---------------------------------------------------------------------------------------------
// C[M,N] = (beta ? C : 0) + bias[j] + sum_k A(i,k) B(k,j)
//   A_KC:  A(i,k) = A[i*lda + k]  (k contiguous)      else A(i,k) = A[k*lda + i]  (i contiguous)
//   B_KC:  B(k,j) = B[j*ldb + k]  (k contiguous)      else B(k,j) = B[k*ldb + j]  (j contiguous)
// 64x64 tile, 16-deep k slices, 256 threads x (4x4).  gridDim.z > 1 = split-K (atomicAdd epilogue; C must hold its
// initial value, bias is added by split 0 only).
// ---------------------------------------------------------------------------------------------
constexpr int SG_T = 64, SG_K = 16, SG_LD = SG_T + 4;

template <bool KC>
__device__ __forceinline__ void sg_load(float (*S)[SG_LD], const float* __restrict__ P, int ld, int r0, int rows, int k0, int kend) {
  const int tid = threadIdx.x;
  if (KC) {
    const int r = tid >> 2, k4 = (tid & 3) * 4;
    const int gr = r0 + r, gk = k0 + k4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (gr < rows) {
      const float* p = P + (size_t)gr * ld + gk;
      if (gk + 3 < kend && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (gk + e < kend) v[e] = __ldg(p + e);
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) S[k4 + e][r] = v[e];
  } else {
    const int k = tid >> 4, r4 = (tid & 15) * 4;
    const int gk = k0 + k, gr = r0 + r4;
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gk < kend) {
      const float* p = P + (size_t)gk * ld + gr;
      if (gr + 3 < rows && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        t = __ldg(reinterpret_cast<const float4*>(p));
      } else {
        if (gr < rows) t.x = __ldg(p);
        if (gr + 1 < rows) t.y = __ldg(p + 1);
        if (gr + 2 < rows) t.z = __ldg(p + 2);
        if (gr + 3 < rows) t.w = __ldg(p + 3);
      }
    }
    *reinterpret_cast<float4*>(&S[k][r4]) = t;
  }
}

template <bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256)
sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
             float* __restrict__ C, int ldc, const float* __restrict__ bias, int beta, int k_per_split) {
  __shared__ __align__(16) float As[SG_K][SG_LD];
  __shared__ __align__(16) float Bs[SG_K][SG_LD];
  const int i0 = blockIdx.y * SG_T, j0 = blockIdx.x * SG_T;
  const int kbeg = blockIdx.z * k_per_split, kend = min(K, kbeg + k_per_split);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = kbeg; k0 < kend; k0 += SG_K) {
    sg_load<A_KC>(As, A, lda, i0, M, k0, kend);
    sg_load<B_KC>(Bs, B, ldb, j0, N, k0, kend);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_K; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
    __syncthreads();
  }
  const bool split = gridDim.z > 1;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i >= M) continue;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = j0 + tx * 4 + c;
      if (j >= N) continue;
      float v = acc[r][c];
      if (bias != nullptr && blockIdx.z == 0) v += __ldg(bias + j);
      float* o = C + (size_t)i * ldc + j;
      if (split) atomicAdd(o, v);
      else *o = beta ? *o + v : v;
    }
  }
}

// out[j] += sum_i X[i*ld + j]   (bias gradients)
__global__ void __launch_bounds__(256)
colsum_kernel(long long M, int N, const float* __restrict__ X, int ld, float* __restrict__ out, int rows_per_block) {
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (j < N)
    for (long long r = r0 + w; r < r1; r += 8) s += __ldg(X + r * ld + j);
  __shared__ float part[8][33];
  part[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && j < N) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += part[k][threadIdx.x & 31];
    atomicAdd(out + j, t);
  }
}

// The same for 16-byte aligned rows (N % 4 == 0): a lane owns 4 columns, a warp reads 512 contiguous bytes per row with 4
// independent 128-bit loads in flight (the scalar form keeps one 128-byte load per warp in flight: 14 us for 32 768 x 128, 2.5 us
// of HBM time).
constexpr int CS_ROWS = 128;
__global__ void __launch_bounds__(256)
colsum_vec_kernel(long long M, int N, const float* __restrict__ X, int ld, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * 128 + lane * 4;
  const long long r0 = (long long)blockIdx.y * CS_ROWS, r1 = min(M, r0 + CS_ROWS);
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (j < N) {
    const float* col = X + j;
    long long r = r0 + w;
    for (; r + 24 < r1; r += 32) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(col + r * ld)), b = __ldg(reinterpret_cast<const float4*>(col + (r + 8) * ld));
      const float4 c = __ldg(reinterpret_cast<const float4*>(col + (r + 16) * ld)), d = __ldg(reinterpret_cast<const float4*>(col + (r + 24) * ld));
      s.x += (a.x + b.x) + (c.x + d.x); s.y += (a.y + b.y) + (c.y + d.y); s.z += (a.z + b.z) + (c.z + d.z); s.w += (a.w + b.w) + (c.w + d.w);
    }
    for (; r < r1; r += 8) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(col + r * ld));
      s.x += a.x; s.y += a.y; s.z += a.z; s.w += a.w;
    }
  }
  __shared__ float4 part[8][32];
  part[w][lane] = s;
  __syncthreads();
  if (w == 0 && j < N) {
    float4 t = part[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) { const float4 q = part[k][lane]; t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w; }
    atomicAdd(out + j, t.x); atomicAdd(out + j + 1, t.y); atomicAdd(out + j + 2, t.z); atomicAdd(out + j + 3, t.w);
  }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over 128 features (nn.LayerNorm(128), eps 1e-5): warp per row, lane = 4 features
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ln_fwd_kernel(long long M, const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
              float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const float4 v = __ldg(reinterpret_cast<const float4*>(x + row * D_MODEL) + lane);
  float s = v.x + v.y + v.z + v.w;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
  const float mean = s * (1.f / D_MODEL);
  const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
  float q = dx * dx + dy * dy + dz * dz + dw * dw;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) q += __shfl_xor_sync(0xFFFFFFFFu, q, d);
  const float rstd = rsqrtf(q * (1.f / D_MODEL) + LN_EPS);
  const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + lane), bb = __ldg(reinterpret_cast<const float4*>(b) + lane);
  reinterpret_cast<float4*>(y + row * D_MODEL)[lane] =
      make_float4(dx * rstd * gg.x + bb.x, dy * rstd * gg.y + bb.y, dz * rstd * gg.z + bb.z, dw * rstd * gg.w + bb.w);
  if (lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
}

// dx = rstd * (g.dy - mean(g.dy) - xhat * mean(g.dy.xhat));  dgamma += dy.xhat;  dbeta += dy   (atomics per CTA)
__global__ void __launch_bounds__(256)
ln_bwd_kernel(long long M, const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ mean_in,
              const float* __restrict__ rstd_in, const float* __restrict__ dy, const float* __restrict__ dres, float* __restrict__ dx,
              float* __restrict__ dgamma, float* __restrict__ dbeta, int rows_per_warp) {
  __shared__ float red[2][D_MODEL];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < D_MODEL) { red[0][threadIdx.x] = 0.f; red[1][threadIdx.x] = 0.f; }
  __syncthreads();
  const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + lane);
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag;
  const long long r0 = ((long long)blockIdx.x * 8 + warp) * rows_per_warp;
  for (long long row = r0; row < min(M, r0 + rows_per_warp); ++row) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + row * D_MODEL) + lane);
    const float4 d = __ldg(reinterpret_cast<const float4*>(dy + row * D_MODEL) + lane);
    const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
    // gradient of the residual connection that bypasses this LayerNorm (x also feeds `x + branch(LN(x))`): added here instead of by a
    // separate accumulation pass over the (M, 128) tensor
    const float4 dr = dres != nullptr ? __ldg(reinterpret_cast<const float4*>(dres + row * D_MODEL) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 xh = make_float4((v.x - mean) * rstd, (v.y - mean) * rstd, (v.z - mean) * rstd, (v.w - mean) * rstd);
    const float4 gd = make_float4(gg.x * d.x, gg.y * d.y, gg.z * d.z, gg.w * d.w);
    float m1 = gd.x + gd.y + gd.z + gd.w;
    float m2 = gd.x * xh.x + gd.y * xh.y + gd.z * xh.z + gd.w * xh.w;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) { m1 += __shfl_xor_sync(0xFFFFFFFFu, m1, s); m2 += __shfl_xor_sync(0xFFFFFFFFu, m2, s); }
    m1 *= (1.f / D_MODEL); m2 *= (1.f / D_MODEL);
    reinterpret_cast<float4*>(dx + row * D_MODEL)[lane] =
        make_float4(rstd * (gd.x - m1 - xh.x * m2) + dr.x, rstd * (gd.y - m1 - xh.y * m2) + dr.y, rstd * (gd.z - m1 - xh.z * m2) + dr.z,
                    rstd * (gd.w - m1 - xh.w * m2) + dr.w);
    ag.x += d.x * xh.x; ag.y += d.y * xh.y; ag.z += d.z * xh.z; ag.w += d.w * xh.w;
    ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w;
  }
  atomicAdd(&red[0][lane * 4], ag.x); atomicAdd(&red[0][lane * 4 + 1], ag.y); atomicAdd(&red[0][lane * 4 + 2], ag.z); atomicAdd(&red[0][lane * 4 + 3], ag.w);
  atomicAdd(&red[1][lane * 4], ab.x); atomicAdd(&red[1][lane * 4 + 1], ab.y); atomicAdd(&red[1][lane * 4 + 2], ab.z); atomicAdd(&red[1][lane * 4 + 3], ab.w);
  __syncthreads();
  if (threadIdx.x < D_MODEL) { atomicAdd(dgamma + threadIdx.x, red[0][threadIdx.x]); atomicAdd(dbeta + threadIdx.x, red[1][threadIdx.x]); }
}

// ---------------------------------------------------------------------------------------------
// GEGLU (attention.py:50-57): h = [a | g] (2*H wide), u = a * gelu(g) with the exact-CDF GELU (F.gelu default) to fp32 accuracy
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
geglu_fwd_kernel(long long total, int H, const float* __restrict__ h, float* __restrict__ u) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const long long row = q / H;
  const int c = (int)(q - row * H);
  const float a = __ldg(h + row * 2 * H + c), g = __ldg(h + row * 2 * H + H + c);
  float cdf, pdf;
  normal_cdf_pdf(g, cdf, pdf);
  u[q] = a * (g * cdf);
}
__global__ void __launch_bounds__(256)
geglu_bwd_kernel(long long total, int H, const float* __restrict__ h, const float* __restrict__ du, float* __restrict__ dh) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const long long row = q / H;
  const int c = (int)(q - row * H);
  const float a = __ldg(h + row * 2 * H + c), g = __ldg(h + row * 2 * H + H + c), d = __ldg(du + q);
  float cdf, pdf;
  normal_cdf_pdf(g, cdf, pdf);
  dh[row * 2 * H + c] = d * g * cdf;
  dh[row * 2 * H + H + c] = d * a * (cdf + g * pdf);
}

// GEGLU fused with the Dropout that follows it in FeedForward (attention.py:77-94: Sequential(GEGLU, Dropout, Linear)) and, in the
// backward pass, with the bias gradient of the Linear in front of it (column sums of dh).  Round 2: the unfused chain wrote and
// re-read the (M, H) activation for the dropout and re-read the (M, 2H) gradient for the column sums.  A thread owns 4 consecutive
// hidden units (float4); the dropout mask is the one dfb200_dropout would draw on u (Philox quad = element index / 4), so fused
// and unfused paths are interchangeable.  p == 0: no mask.  Requires H % 4 == 0.
__global__ void __launch_bounds__(256)
geglu_dropout_fwd_kernel(long long quads, int H4, float p, float scale, uint64_t seed, uint64_t offset,
                         const unsigned long long* __restrict__ step, const float4* __restrict__ h, float4* __restrict__ u) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= quads) return;
  const long long row = q / H4;
  const int c = (int)(q - row * H4);
  const float4 a = __ldg(h + row * 2 * H4 + c), g = __ldg(h + row * 2 * H4 + H4 + c);
  float m[4] = {1.f, 1.f, 1.f, 1.f};
  if (p > 0.f) {
    if (step != nullptr) seed += (uint64_t)__ldg(step) * 0x9E3779B97F4A7C15ull;
    dropout_keep4(p, scale, seed, offset, q, m);
  }
  float c0, c1, c2, c3, pd;
  normal_cdf_pdf(g.x, c0, pd); normal_cdf_pdf(g.y, c1, pd); normal_cdf_pdf(g.z, c2, pd); normal_cdf_pdf(g.w, c3, pd);
  u[q] = make_float4(m[0] * a.x * (g.x * c0), m[1] * a.y * (g.y * c1), m[2] * a.z * (g.z * c2), m[3] * a.w * (g.w * c3));
}
// grid (ceil(M / GD_ROWS), H4 / 128); block 256 = 128 quads x 2 row phases.  db_accum (2H floats, may be NULL) += column sums of dh.
constexpr int GD_ROWS = 128;
__global__ void __launch_bounds__(256)
geglu_dropout_bwd_kernel(long long M, int H4, float p, float scale, uint64_t seed, uint64_t offset,
                         const unsigned long long* __restrict__ step, const float4* __restrict__ h, const float4* __restrict__ du,
                         float4* __restrict__ dh, float* __restrict__ db_accum) {
  const int c = blockIdx.y * 128 + (threadIdx.x & 127), rs = threadIdx.x >> 7;
  const long long r0 = (long long)blockIdx.x * GD_ROWS, r1 = min(M, r0 + GD_ROWS);
  if (p > 0.f && step != nullptr) seed += (uint64_t)__ldg(step) * 0x9E3779B97F4A7C15ull;
  float sa[4] = {0.f, 0.f, 0.f, 0.f}, sg[4] = {0.f, 0.f, 0.f, 0.f};
  if (c < H4) {
#pragma unroll 2
    for (long long row = r0 + rs; row < r1; row += 2) {
      const float4 a4 = __ldg(h + row * 2 * H4 + c), g4 = __ldg(h + row * 2 * H4 + H4 + c), d4 = __ldg(du + row * H4 + c);
      float m[4] = {1.f, 1.f, 1.f, 1.f};
      if (p > 0.f) dropout_keep4(p, scale, seed, offset, row * H4 + c, m);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, g[4] = {g4.x, g4.y, g4.z, g4.w}, d[4] = {d4.x * m[0], d4.y * m[1], d4.z * m[2], d4.w * m[3]};
      float da[4], dg[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float cdf, pdf;
        normal_cdf_pdf(g[i], cdf, pdf);
        da[i] = d[i] * g[i] * cdf;
        dg[i] = d[i] * a[i] * (cdf + g[i] * pdf);
        sa[i] += da[i]; sg[i] += dg[i];
      }
      dh[row * 2 * H4 + c] = make_float4(da[0], da[1], da[2], da[3]);
      dh[row * 2 * H4 + H4 + c] = make_float4(dg[0], dg[1], dg[2], dg[3]);
    }
  }
  if (db_accum == nullptr) return;
  __shared__ float part[128][8];
  if (rs == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { part[threadIdx.x & 127][i] = sa[i]; part[threadIdx.x & 127][4 + i] = sg[i]; }
  }
  __syncthreads();
  if (rs == 0 && c < H4) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      atomicAdd(db_accum + 4 * c + i, sa[i] + part[threadIdx.x][i]);
      atomicAdd(db_accum + 4 * H4 + 4 * c + i, sg[i] + part[threadIdx.x][4 + i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Cross-attention core over the 4 part tokens (attention.py:183-203): 8 heads x 16, masked softmax over 4 keys.
// Warp per token, lane = 4 of the 128 dims (head = lane / 4).  q/o (B*N,128), k/v (B,4,128), probs (B*N,8,4).
// ---------------------------------------------------------------------------------------------
constexpr int PA_TOK = 64;   // tokens per CTA (of one sample: N % PA_TOK is handled by clamping), 8 per warp
__global__ void __launch_bounds__(256)
part_attn_fwd_kernel(int N, const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                     const float* __restrict__ valid, float* __restrict__ o, float* __restrict__ probs) {
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float4* kb = reinterpret_cast<const float4*>(k + (size_t)b * MAX_TOKENS * D_MODEL);
  const float4* vb = reinterpret_cast<const float4*>(v + (size_t)b * MAX_TOKENS * D_MODEL);
  float4 k4[MAX_TOKENS], v4[MAX_TOKENS];
  bool ok[MAX_TOKENS];
#pragma unroll
  for (int j = 0; j < MAX_TOKENS; ++j) {
    k4[j] = __ldg(kb + j * 32 + lane); v4[j] = __ldg(vb + j * 32 + lane);
    ok[j] = valid == nullptr || __ldg(valid + b * MAX_TOKENS + j) != 0.f;
  }
  const int p0 = blockIdx.x * PA_TOK;
  for (int p = p0 + warp; p < min(N, p0 + PA_TOK); p += 8) {
    const size_t tok = (size_t)b * N + p;
    const float4 q4 = __ldg(reinterpret_cast<const float4*>(q + tok * D_MODEL) + lane);
    float s[MAX_TOKENS];
#pragma unroll
    for (int j = 0; j < MAX_TOKENS; ++j) {
      float d = q4.x * k4[j].x + q4.y * k4[j].y + q4.z * k4[j].z + q4.w * k4[j].w;
      d += __shfl_xor_sync(0xFFFFFFFFu, d, 1);
      d += __shfl_xor_sync(0xFFFFFFFFu, d, 2);
      s[j] = ok[j] ? d * 0.25f : -FLT_MAX;  // dim_head ** -0.5; masked_fill(~mask, -finfo.max)
    }
    const float mx = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
    float e[MAX_TOKENS], sum = 0.f;
#pragma unroll
    for (int j = 0; j < MAX_TOKENS; ++j) { e[j] = expf(s[j] - mx); sum += e[j]; }
    const float inv = 1.f / sum;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < MAX_TOKENS; ++j) {
      const float pj = e[j] * inv;
      acc.x = fmaf(pj, v4[j].x, acc.x); acc.y = fmaf(pj, v4[j].y, acc.y); acc.z = fmaf(pj, v4[j].z, acc.z); acc.w = fmaf(pj, v4[j].w, acc.w);
      e[j] = pj;
    }
    reinterpret_cast<float4*>(o + tok * D_MODEL)[lane] = acc;
    if ((lane & 3) == 0) reinterpret_cast<float4*>(probs + tok * 32)[lane >> 2] = make_float4(e[0], e[1], e[2], e[3]);
  }
}

// dq per token; dk/dv accumulated over the tokens of the sample (registers -> shared -> one atomicAdd per CTA and element)
__global__ void __launch_bounds__(256)
part_attn_bwd_kernel(int N, const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                     const float* __restrict__ valid, const float* __restrict__ probs, const float* __restrict__ d_o,
                     float* __restrict__ dq, float* __restrict__ dk, float* __restrict__ dv) {
  __shared__ float red[2][MAX_TOKENS][D_MODEL];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 2 * MAX_TOKENS * D_MODEL; i += 256) (&red[0][0][0])[i] = 0.f;
  __syncthreads();
  const float4* kb = reinterpret_cast<const float4*>(k + (size_t)b * MAX_TOKENS * D_MODEL);
  const float4* vb = reinterpret_cast<const float4*>(v + (size_t)b * MAX_TOKENS * D_MODEL);
  float4 k4[MAX_TOKENS], v4[MAX_TOKENS], ak[MAX_TOKENS], av[MAX_TOKENS];
  bool ok[MAX_TOKENS];
#pragma unroll
  for (int j = 0; j < MAX_TOKENS; ++j) {
    k4[j] = __ldg(kb + j * 32 + lane); v4[j] = __ldg(vb + j * 32 + lane);
    ak[j] = make_float4(0.f, 0.f, 0.f, 0.f); av[j] = ak[j];
    ok[j] = valid == nullptr || __ldg(valid + b * MAX_TOKENS + j) != 0.f;
  }
  const int p0 = blockIdx.x * PA_TOK;
  for (int p = p0 + warp; p < min(N, p0 + PA_TOK); p += 8) {
    const size_t tok = (size_t)b * N + p;
    const float4 q4 = __ldg(reinterpret_cast<const float4*>(q + tok * D_MODEL) + lane);
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(d_o + tok * D_MODEL) + lane);
    const float4 pr = __ldg(reinterpret_cast<const float4*>(probs + tok * 32) + (lane >> 2));
    const float pj[MAX_TOKENS] = {pr.x, pr.y, pr.z, pr.w};
    float dp[MAX_TOKENS], dot = 0.f;
#pragma unroll
    for (int j = 0; j < MAX_TOKENS; ++j) {
      float d = g4.x * v4[j].x + g4.y * v4[j].y + g4.z * v4[j].z + g4.w * v4[j].w;
      d += __shfl_xor_sync(0xFFFFFFFFu, d, 1);
      d += __shfl_xor_sync(0xFFFFFFFFu, d, 2);
      dp[j] = d;
      dot = fmaf(pj[j], d, dot);
    }
    float4 dq4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < MAX_TOKENS; ++j) {
      // softmax backward; masked logits receive no gradient (masked_fill), and sim = 0.25 * q.k
      const float ds = ok[j] ? 0.25f * pj[j] * (dp[j] - dot) : 0.f;
      dq4.x = fmaf(ds, k4[j].x, dq4.x); dq4.y = fmaf(ds, k4[j].y, dq4.y); dq4.z = fmaf(ds, k4[j].z, dq4.z); dq4.w = fmaf(ds, k4[j].w, dq4.w);
      ak[j].x = fmaf(ds, q4.x, ak[j].x); ak[j].y = fmaf(ds, q4.y, ak[j].y); ak[j].z = fmaf(ds, q4.z, ak[j].z); ak[j].w = fmaf(ds, q4.w, ak[j].w);
      av[j].x = fmaf(pj[j], g4.x, av[j].x); av[j].y = fmaf(pj[j], g4.y, av[j].y); av[j].z = fmaf(pj[j], g4.z, av[j].z); av[j].w = fmaf(pj[j], g4.w, av[j].w);
    }
    reinterpret_cast<float4*>(dq + tok * D_MODEL)[lane] = dq4;
  }
#pragma unroll
  for (int j = 0; j < MAX_TOKENS; ++j) {
    float* rk = &red[0][j][lane * 4];
    float* rv = &red[1][j][lane * 4];
    atomicAdd(rk, ak[j].x); atomicAdd(rk + 1, ak[j].y); atomicAdd(rk + 2, ak[j].z); atomicAdd(rk + 3, ak[j].w);
    atomicAdd(rv, av[j].x); atomicAdd(rv + 1, av[j].y); atomicAdd(rv + 2, av[j].z); atomicAdd(rv + 3, av[j].w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MAX_TOKENS * D_MODEL; i += 256) {
    atomicAdd(dk + (size_t)b * MAX_TOKENS * D_MODEL + i, (&red[0][0][0])[i]);
    atomicAdd(dv + (size_t)b * MAX_TOKENS * D_MODEL + i, (&red[1][0][0])[i]);
  }
}

// timestep_embedding (nets/utils.py:7-24): out[b] = [cos(t_b f_0..f_127) | sin(t_b f_0..f_127)]
__global__ void __launch_bounds__(128)
timestep_embedding_kernel(const float* __restrict__ t, const float* __restrict__ freqs, float* __restrict__ out) {
  const int b = blockIdx.x, i = threadIdx.x;
  const float a = __ldg(t + b) * __ldg(freqs + i);
  out[(size_t)b * 256 + i] = cosf(a);
  out[(size_t)b * 256 + 128 + i] = sinf(a);
}

// Inverted dropout with a counter-based mask (Philox, one draw per 4 elements): y = (keep ? x / (1 - p) : 0) (+ residual).  The backward
// pass calls the same kernel on the incoming gradient with the same (seed, offset).
__global__ void __launch_bounds__(256)
dropout_kernel(long long total, float p, float scale, uint64_t seed, uint64_t offset, const unsigned long long* __restrict__ step,
               const float* __restrict__ x, const float* __restrict__ residual, float* __restrict__ y) {
  const long long q4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long e = q4 * 4;
  if (e >= total) return;
  if (step != nullptr) seed += (uint64_t)__ldg(step) * 0x9E3779B97F4A7C15ull;  // device-side step counter (CUDA-graph replays)
  uint32_t r[4];
  philox4x32_10((uint64_t)q4, offset, seed, r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (e + i < total) {
      const float u = (float)(r[i] >> 8) * 5.9604644775390625e-8f;  // [0, 1)
      const float r = residual != nullptr ? __ldg(residual + e + i) : 0.f;
      y[e + i] = (u >= p ? __ldg(x + e + i) * scale : 0.f) + r;
    }
  }
}

// q_sample backward (anchored_diffusion.py:169-173): x_t = sa (x0 - a) + a + sb sqrt(var) noise
//   d x0 = sa g;   d a = (1 - sa) g;   d var = sb noise g / (2 sqrt(var))
__global__ void __launch_bounds__(256)
q_sample_bwd_kernel(long long total, int per_sample, int T, const float* __restrict__ sched, const int* __restrict__ t,
                    const float* __restrict__ variance, const float* __restrict__ noise, const float* __restrict__ g,
                    float* __restrict__ dx0, float* __restrict__ da, float* __restrict__ dvar) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const int tt = __ldg(t + (int)(q / per_sample));
  const float sa = __ldg(sched + DFB200_SCHED_SQRT_ALPHAS_CUMPROD * T + tt);
  const float sb = __ldg(sched + DFB200_SCHED_SQRT_ONE_MINUS_ALPHAS_CUMPROD * T + tt);
  const float gg = __ldg(g + q);
  if (dx0 != nullptr) dx0[q] = sa * gg;
  if (da != nullptr) da[q] = (1.f - sa) * gg;
  if (dvar != nullptr) dvar[q] = sb * __ldg(noise + q) * gg * 0.5f * rsqrtf(__ldg(variance + q));
}

}  // namespace dfb200

using namespace dfb200;

extern "C" int dfb200_sgemm(int a_k_contiguous, int b_k_contiguous, int M, int N, int K, const float* A, int lda, const float* B,
                            int ldb, float* C, int ldc, const float* bias, int beta, int split_k, dfb200_stream_t stream) {
  DFB_REQUIRE(M >= 0 && N >= 0 && K >= 0 && split_k >= 1, DFB200_ERR_INVALID_ARG, "sgemm: bad sizes M=%d N=%d K=%d split=%d", M, N, K, split_k);
  if (M == 0 || N == 0) return DFB200_OK;
  int kps = cdiv(cdiv(K, split_k), SG_K) * SG_K;
  if (kps == 0) kps = SG_K;
  const int splits = K == 0 ? 1 : cdiv(K, kps);
  dim3 grid(cdiv(N, SG_T), cdiv(M, SG_T), splits);
  DFB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, DFB200_ERR_INVALID_ARG, "sgemm: grid too large");
  cudaStream_t st = as_stream(stream);
#define SG_LAUNCH(AK, BK) sgemm_kernel<AK, BK><<<grid, 256, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, beta, kps)
  if (a_k_contiguous && b_k_contiguous) SG_LAUNCH(true, true);
  else if (a_k_contiguous) SG_LAUNCH(true, false);
  else if (b_k_contiguous) SG_LAUNCH(false, true);
  else SG_LAUNCH(false, false);
#undef SG_LAUNCH
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_colsum_accumulate(long long M, int N, const float* X, int ld, float* out, dfb200_stream_t stream) {
  if (M <= 0 || N <= 0) return DFB200_OK;
  if (N % 4 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 && M >= 1024) {
    colsum_vec_kernel<<<dim3(cdiv(N, 128), (unsigned)cdiv(M, (long long)CS_ROWS)), 256, 0, as_stream(stream)>>>(M, N, X, ld, out);
    DFB_LAUNCH_CHECK();
    return DFB200_OK;
  }
  const int rpb = 512;
  colsum_kernel<<<dim3(cdiv(N, 32), (unsigned)cdiv(M, (long long)rpb)), 256, 0, as_stream(stream)>>>(M, N, X, ld, out, rpb);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_layernorm128_forward(long long M, const float* x, const float* gamma, const float* beta, float* y,
                                           float* mean, float* rstd, dfb200_stream_t stream) {
  if (M <= 0) return DFB200_OK;
  ln_fwd_kernel<<<(unsigned)cdiv(M, 8LL), 256, 0, as_stream(stream)>>>(M, x, gamma, beta, y, mean, rstd);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_layernorm128_backward(long long M, const float* x, const float* gamma, const float* mean, const float* rstd,
                                            const float* dy, float* dx, float* dgamma_accum, float* dbeta_accum,
                                            dfb200_stream_t stream) {
  if (M <= 0) return DFB200_OK;
  const int rpw = 16;
  ln_bwd_kernel<<<(unsigned)cdiv(M, (long long)(8 * rpw)), 256, 0, as_stream(stream)>>>(M, x, gamma, mean, rstd, dy, nullptr, dx, dgamma_accum,
                                                                                        dbeta_accum, rpw);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_layernorm128_backward_residual(long long M, const float* x, const float* gamma, const float* mean, const float* rstd,
                                                      const float* dy, const float* dres, float* dx, float* dgamma_accum, float* dbeta_accum,
                                                      dfb200_stream_t stream) {
  if (M <= 0) return DFB200_OK;
  const int rpw = 16;
  ln_bwd_kernel<<<(unsigned)cdiv(M, (long long)(8 * rpw)), 256, 0, as_stream(stream)>>>(M, x, gamma, mean, rstd, dy, dres, dx, dgamma_accum,
                                                                                        dbeta_accum, rpw);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_geglu_forward(long long M, int H, const float* h, float* u, dfb200_stream_t stream) {
  const long long total = M * H;
  if (total <= 0) return DFB200_OK;
  geglu_fwd_kernel<<<(unsigned)cdiv(total, 256LL), 256, 0, as_stream(stream)>>>(total, H, h, u);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_geglu_backward(long long M, int H, const float* h, const float* du, float* dh, dfb200_stream_t stream) {
  const long long total = M * H;
  if (total <= 0) return DFB200_OK;
  geglu_bwd_kernel<<<(unsigned)cdiv(total, 256LL), 256, 0, as_stream(stream)>>>(total, H, h, du, dh);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_geglu_dropout_forward(long long M, int H, float p, uint64_t seed, uint64_t offset, const unsigned long long* step,
                                           const float* h, float* u, dfb200_stream_t stream) {
  DFB_REQUIRE(H % 4 == 0 && p >= 0.f && p < 1.f, DFB200_ERR_INVALID_ARG, "geglu_dropout_forward: H %% 4 != 0 or p outside [0, 1) (H=%d)", H);
  const long long quads = M * (H / 4);
  if (quads <= 0) return DFB200_OK;
  geglu_dropout_fwd_kernel<<<(unsigned)cdiv(quads, 256LL), 256, 0, as_stream(stream)>>>(quads, H / 4, p, 1.f / (1.f - p), seed, offset, step,
                                                                                      reinterpret_cast<const float4*>(h),
                                                                                      reinterpret_cast<float4*>(u));
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_geglu_dropout_backward(long long M, int H, float p, uint64_t seed, uint64_t offset, const unsigned long long* step,
                                            const float* h, const float* du, float* dh, float* db_accum, dfb200_stream_t stream) {
  DFB_REQUIRE(H % 4 == 0 && p >= 0.f && p < 1.f, DFB200_ERR_INVALID_ARG, "geglu_dropout_backward: H %% 4 != 0 or p outside [0, 1) (H=%d)", H);
  if (M <= 0 || H <= 0) return DFB200_OK;
  geglu_dropout_bwd_kernel<<<dim3((unsigned)cdiv(M, (long long)GD_ROWS), cdiv(H / 4, 128)), 256, 0, as_stream(stream)>>>(
      M, H / 4, p, 1.f / (1.f - p), seed, offset, step, reinterpret_cast<const float4*>(h), reinterpret_cast<const float4*>(du),
      reinterpret_cast<float4*>(dh), db_accum);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_part_attention_forward(int B, int N, const float* q, const float* k, const float* v, const float* valid_id,
                                             float* o, float* probs, dfb200_stream_t stream) {
  DFB_REQUIRE(B >= 0 && N >= 0 && B <= 65535, DFB200_ERR_INVALID_ARG, "part_attention_forward: bad sizes B=%d N=%d", B, N);
  if (B == 0 || N == 0) return DFB200_OK;
  part_attn_fwd_kernel<<<dim3(cdiv(N, PA_TOK), B), 256, 0, as_stream(stream)>>>(N, q, k, v, valid_id, o, probs);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_part_attention_backward(int B, int N, const float* q, const float* k, const float* v, const float* valid_id,
                                              const float* probs, const float* d_o, float* dq, float* dk_accum, float* dv_accum,
                                              dfb200_stream_t stream) {
  DFB_REQUIRE(B >= 0 && N >= 0 && B <= 65535, DFB200_ERR_INVALID_ARG, "part_attention_backward: bad sizes B=%d N=%d", B, N);
  if (B == 0 || N == 0) return DFB200_OK;
  part_attn_bwd_kernel<<<dim3(cdiv(N, PA_TOK), B), 256, 0, as_stream(stream)>>>(N, q, k, v, valid_id, probs, d_o, dq, dk_accum, dv_accum);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_timestep_embedding(int B, const float* t, const float* freqs128, float* out, dfb200_stream_t stream) {
  if (B <= 0) return DFB200_OK;
  timestep_embedding_kernel<<<B, 128, 0, as_stream(stream)>>>(t, freqs128, out);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_dropout(size_t count, float p, uint64_t seed, uint64_t offset, const float* x, const float* residual, float* y,
                              dfb200_stream_t stream) {
  DFB_REQUIRE(p >= 0.f && p < 1.f, DFB200_ERR_INVALID_ARG, "dropout: p must be in [0, 1)");
  if (count == 0) return DFB200_OK;
  dropout_kernel<<<(unsigned)cdiv((long long)((count + 3) / 4), 256LL), 256, 0, as_stream(stream)>>>((long long)count, p, 1.f / (1.f - p), seed,
                                                                                                     offset, nullptr, x, residual, y);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_dropout_stepped(size_t count, float p, uint64_t seed, uint64_t offset, const unsigned long long* step, const float* x,
                                      const float* residual, float* y, dfb200_stream_t stream) {
  DFB_REQUIRE(p >= 0.f && p < 1.f, DFB200_ERR_INVALID_ARG, "dropout: p must be in [0, 1)");
  if (count == 0) return DFB200_OK;
  dropout_kernel<<<(unsigned)cdiv((long long)((count + 3) / 4), 256LL), 256, 0, as_stream(stream)>>>((long long)count, p, 1.f / (1.f - p), seed,
                                                                                                     offset, step, x, residual, y);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_q_sample_backward(int B, int N, int T, const float* sched, const int* t, const float* variance,
                                        const float* noise, const float* grad_x_t, float* grad_x_start, float* grad_anchors,
                                        float* grad_variance, dfb200_stream_t stream) {
  DFB_REQUIRE(B >= 0 && N >= 0 && T > 0, DFB200_ERR_INVALID_ARG, "q_sample_backward: bad sizes B=%d N=%d T=%d", B, N, T);
  const long long total = (long long)B * 3 * N;
  if (total == 0) return DFB200_OK;
  q_sample_bwd_kernel<<<(unsigned)cdiv(total, 256LL), 256, 0, as_stream(stream)>>>(total, 3 * N, T, sched, t, variance, noise, grad_x_t,
                                                                                   grad_x_start, grad_anchors, grad_variance);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

// =============================================================================================
// Training-side encoder primitives: PointNetV2 (reference models/encoders/pointnet.py:122-214)
// = 1x1 convolutions (dfb200_sgemm / dfb200_gemm_bf16 over the B*N point rows) + BatchNorm1d + ReLU + the anchor-weighted
// max-pool, and the forward direction of the latent coupling flows with their log-determinant (encoders/flow.py:24-45).
// =============================================================================================
namespace dfb200 {

// Column sums over M rows of a row-major (M, C) matrix, optionally of TWO quantities at once:
//   mode 0 (BatchNorm forward statistics):  s1 = sum x,            s2 = sum x^2
//   mode 1 (BatchNorm backward):            s1 = sum g,            s2 = sum g * xhat,   g = dy * (y > 0 if relu)
// One thread per column inside a 32-column x 8-row-slice tile; partial sums land with atomicAdd.
__global__ void __launch_bounds__(256)
bn_colstats_kernel(long long M, int C, int mode, int relu, const float* __restrict__ x, const float* __restrict__ dy,
                   const float* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ rstd,
                   float* __restrict__ s1, float* __restrict__ s2, int rows_per_block) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int w = threadIdx.x >> 5;
  const long long r0 = (long long)blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float a = 0.f, b = 0.f;
  if (c < C) {
    const float mu = mode ? __ldg(mean + c) : 0.f, rs = mode ? __ldg(rstd + c) : 0.f;
    for (long long r = r0 + w; r < r1; r += 8) {
      const float xv = __ldg(x + r * C + c);
      if (mode == 0) {
        a += xv;
        b = fmaf(xv, xv, b);
      } else {
        float g = __ldg(dy + r * C + c);
        if (relu && __ldg(y + r * C + c) <= 0.f) g = 0.f;
        a += g;
        b = fmaf(g, (xv - mu) * rs, b);
      }
    }
  }
  __shared__ float pa[8][33], pb[8][33];
  pa[w][threadIdx.x & 31] = a;
  pb[w][threadIdx.x & 31] = b;
  __syncthreads();
  if (w == 0 && c < C) {
    float ta = 0.f, tb = 0.f;
    for (int k = 0; k < 8; ++k) { ta += pa[k][threadIdx.x & 31]; tb += pb[k][threadIdx.x & 31]; }
    atomicAdd(s1 + c, ta);
    atomicAdd(s2 + c, tb);
  }
}

// mean / rstd from the column sums (biased variance, eps 1e-5) and the running-statistics update of nn.BatchNorm1d
// (momentum 0.1, unbiased variance)
__global__ void bn_finalize_kernel(long long M, int C, const float* __restrict__ s1, const float* __restrict__ s2, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ running_mean, float* __restrict__ running_var, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mu = s1[c] / (float)M;
  const float var = fmaxf(s2[c] / (float)M - mu * mu, 0.f);
  mean[c] = mu;
  rstd[c] = rsqrtf(var + 1e-5f);
  if (running_mean != nullptr) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * ((float)M / (float)max(M - 1, 1LL));
  }
}

// y = (x - mean) * rstd * gamma + beta (+ ReLU)
__global__ void __launch_bounds__(256)
bn_apply_kernel(long long total, int C, int relu, const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ y) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const int c = (int)(q % C);
  float v = (__ldg(x + q) - __ldg(mean + c)) * __ldg(rstd + c) * __ldg(gamma + c) + __ldg(beta + c);
  if (relu) v = fmaxf(v, 0.f);
  y[q] = v;
}

// dx = gamma * rstd * (g - sum_g / M - xhat * sum_gx / M),  g = dy * (y > 0 if relu)
__global__ void __launch_bounds__(256)
bn_dx_kernel(long long total, long long M, int C, int relu, const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ y,
             const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ sum_g,
             const float* __restrict__ sum_gx, float* __restrict__ dx) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const int c = (int)(q % C);
  float g = __ldg(dy + q);
  if (relu && __ldg(y + q) <= 0.f) g = 0.f;
  const float rs = __ldg(rstd + c);
  const float xh = (__ldg(x + q) - __ldg(mean + c)) * rs;
  const float invM = 1.f / (float)M;
  dx[q] = __ldg(gamma + c) * rs * (g - __ldg(sum_g + c) * invM - xh * __ldg(sum_gx + c) * invM);
}

// ReLU backward: dx = dy * (y > 0)
__global__ void __launch_bounds__(256) relu_bwd_kernel(long long n, const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = __ldg(y + i) > 0.f ? __ldg(dy + i) : 0.f;
}

// Anchor-weighted max-pool of PointNetV2 (pointnet.py:194-198) WITHOUT the (B, C, N, A) intermediate:
//   out[b, c, a] = max_n  x[b, n, c] * w[b, n, a] * scale        x (B, N, C) channel-last, w (B, N, A), A <= 4
// thread = (channel, anchor); the point loop reads x coalesced over channels and w broadcast.  arg[b,c,a] = winning n.
__global__ void __launch_bounds__(128)
weighted_maxpool_fwd_kernel(int N, int C, int A, float scale, const float* __restrict__ x, const float* __restrict__ w,
                            float* __restrict__ out, int* __restrict__ arg) {
  const int b = blockIdx.y;
  const int cl = threadIdx.x & 31, a = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  if (c >= C || a >= A) return;
  const float* xb = x + (size_t)b * N * C + c;
  const float* wb = w + (size_t)b * N * A + a;
  float best = -INFINITY;
  int bi = 0;
  for (int n = 0; n < N; ++n) {
    const float v = __ldg(xb + (size_t)n * C) * __ldg(wb + (size_t)n * A) * scale;
    if (v > best) { best = v; bi = n; }  // first maximum, as torch.max
  }
  out[((size_t)b * C + c) * A + a] = best;
  arg[((size_t)b * C + c) * A + a] = bi;
}
// dx[b, arg, c] += w[b, arg, a] * scale * dout[b, c, a]      (dx zeroed by the caller)
__global__ void __launch_bounds__(256)
weighted_maxpool_bwd_kernel(long long total, int N, int C, int A, float scale, const float* __restrict__ w, const int* __restrict__ arg,
                            const float* __restrict__ dout, float* __restrict__ dx) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= total) return;
  const int a = (int)(q % A);
  const int c = (int)((q / A) % C);
  const long long b = q / ((long long)A * C);
  const int n = __ldg(arg + q);
  atomicAdd(dx + ((size_t)b * N + n) * C + c, __ldg(w + ((size_t)b * N + n) * A + a) * scale * __ldg(dout + q));
}

// CouplingLayer forward (flow.py:33-37): scale = sigmoid(s + 2); y1 = x2 * scale + shift; logdet[r] = sum_c log(scale)
// s_t (B, 2d) = [s | shift]; x2 / y1: d columns with leading dimension ld (y1 may alias x2).  One warp per row.
__global__ void __launch_bounds__(256)
coupling_fwd_kernel(int B, int d, const float* __restrict__ s_t, const float* __restrict__ x2, int ldx, float* __restrict__ y1, int ldy,
                    float* __restrict__ logdet) {
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= B) return;
  float ld = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float sc = 1.f / (1.f + expf(-(__ldg(s_t + (size_t)r * 2 * d + c) + 2.f)));
    y1[(size_t)r * ldy + c] = x2[(size_t)r * ldx + c] * sc + __ldg(s_t + (size_t)r * 2 * d + d + c);
    ld += logf(sc);
  }
#pragma unroll
  for (int k = 16; k >= 1; k >>= 1) ld += __shfl_xor_sync(0xFFFFFFFFu, ld, k);
  if (lane == 0) logdet[r] = ld;
}
// backward: given dy1 (B,d; ld ldy) and dlogdet (B): ds_t (B,2d) and dx2 (B,d; ld ldx)
//   d scale = dy1 * x2 + dlogdet / scale;  ds = d scale * scale (1 - scale);  dshift = dy1;  dx2 = dy1 * scale
__global__ void __launch_bounds__(256)
coupling_bwd_kernel(int B, int d, const float* __restrict__ s_t, const float* __restrict__ x2, int ldx, const float* __restrict__ dy1, int ldy,
                    const float* __restrict__ dlogdet, float* __restrict__ ds_t, float* __restrict__ dx2, int lddx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * d) return;
  const int r = i / d, c = i - r * d;
  const float sc = 1.f / (1.f + expf(-(__ldg(s_t + (size_t)r * 2 * d + c) + 2.f)));
  const float g = __ldg(dy1 + (size_t)r * ldy + c);
  const float dsc = g * __ldg(x2 + (size_t)r * ldx + c) + __ldg(dlogdet + r) / sc;
  ds_t[(size_t)r * 2 * d + c] = dsc * sc * (1.f - sc);
  ds_t[(size_t)r * 2 * d + d + c] = g;
  dx2[(size_t)r * lddx + c] = g * sc;
}

}  // namespace dfb200

extern "C" int dfb200_batchnorm_forward(long long M, int C, int relu, const float* x, const float* gamma, const float* beta, float* y,
                                        float* mean, float* rstd, float* running_mean, float* running_var, float momentum,
                                        float* scratch2C, dfb200_stream_t stream) {
  DFB_REQUIRE(M >= 1 && C >= 1, DFB200_ERR_INVALID_ARG, "batchnorm_forward: bad sizes M=%lld C=%d", M, C);
  cudaStream_t st = as_stream(stream);
  DFB_CUDA(cudaMemsetAsync(scratch2C, 0, sizeof(float) * 2 * (size_t)C, st));
  const int rpb = 1024;
  bn_colstats_kernel<<<dim3(cdiv(C, 32), (unsigned)cdiv(M, (long long)rpb)), 256, 0, st>>>(M, C, 0, 0, x, nullptr, nullptr, nullptr, nullptr,
                                                                                          scratch2C, scratch2C + C, rpb);
  DFB_LAUNCH_CHECK();
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>(M, C, scratch2C, scratch2C + C, mean, rstd, running_mean, running_var, momentum);
  DFB_LAUNCH_CHECK();
  const long long total = M * C;
  bn_apply_kernel<<<(unsigned)cdiv(total, 256LL), 256, 0, st>>>(total, C, relu, x, mean, rstd, gamma, beta, y);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

/* eval mode: mean = running_mean, rstd = 1/sqrt(running_var + eps) supplied by the caller */
extern "C" int dfb200_batchnorm_apply(long long M, int C, int relu, const float* x, const float* mean, const float* rstd, const float* gamma,
                                      const float* beta, float* y, dfb200_stream_t stream) {
  const long long total = M * C;
  if (total <= 0) return DFB200_OK;
  bn_apply_kernel<<<(unsigned)cdiv(total, 256LL), 256, 0, as_stream(stream)>>>(total, C, relu, x, mean, rstd, gamma, beta, y);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_batchnorm_backward(long long M, int C, int relu, const float* x, const float* y, const float* dy, const float* gamma,
                                         const float* mean, const float* rstd, float* dx, float* dgamma, float* dbeta, dfb200_stream_t stream) {
  DFB_REQUIRE(M >= 1 && C >= 1, DFB200_ERR_INVALID_ARG, "batchnorm_backward: bad sizes M=%lld C=%d", M, C);
  cudaStream_t st = as_stream(stream);
  DFB_CUDA(cudaMemsetAsync(dgamma, 0, sizeof(float) * (size_t)C, st));
  DFB_CUDA(cudaMemsetAsync(dbeta, 0, sizeof(float) * (size_t)C, st));
  const int rpb = 1024;
  bn_colstats_kernel<<<dim3(cdiv(C, 32), (unsigned)cdiv(M, (long long)rpb)), 256, 0, st>>>(M, C, 1, relu, x, dy, y, mean, rstd, dbeta, dgamma, rpb);
  DFB_LAUNCH_CHECK();
  const long long total = M * C;
  bn_dx_kernel<<<(unsigned)cdiv(total, 256LL), 256, 0, st>>>(total, M, C, relu, x, dy, y, mean, rstd, gamma, dbeta, dgamma, dx);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_relu_backward(size_t count, const float* y, const float* dy, float* dx, dfb200_stream_t stream) {
  if (count == 0) return DFB200_OK;
  relu_bwd_kernel<<<(unsigned)cdiv((long long)count, 256LL), 256, 0, as_stream(stream)>>>((long long)count, y, dy, dx);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_weighted_maxpool_forward(int B, int N, int C, int A, float scale, const float* x, const float* w, float* out, int* arg,
                                               dfb200_stream_t stream) {
  DFB_REQUIRE(B >= 0 && N >= 1 && C >= 1 && A >= 1 && A <= 4 && B <= 65535, DFB200_ERR_INVALID_ARG, "weighted_maxpool_forward: bad sizes");
  if (B == 0) return DFB200_OK;
  weighted_maxpool_fwd_kernel<<<dim3(cdiv(C, 32), B), 128, 0, as_stream(stream)>>>(N, C, A, scale, x, w, out, arg);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_weighted_maxpool_backward(int B, int N, int C, int A, float scale, const float* w, const int* arg, const float* dout,
                                                float* dx_zeroed, dfb200_stream_t stream) {
  const long long total = (long long)B * C * A;
  if (total <= 0) return DFB200_OK;
  weighted_maxpool_bwd_kernel<<<(unsigned)cdiv(total, 256LL), 256, 0, as_stream(stream)>>>(total, N, C, A, scale, w, arg, dout, dx_zeroed);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_coupling_forward(int B, int d, const float* s_t, const float* x2, int ldx, float* y1, int ldy, float* logdet,
                                       dfb200_stream_t stream) {
  if (B <= 0 || d <= 0) return DFB200_OK;
  coupling_fwd_kernel<<<cdiv(B, 8), 256, 0, as_stream(stream)>>>(B, d, s_t, x2, ldx, y1, ldy, logdet);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_coupling_backward(int B, int d, const float* s_t, const float* x2, int ldx, const float* dy1, int ldy,
                                        const float* dlogdet, float* ds_t, float* dx2, int lddx, dfb200_stream_t stream) {
  if (B <= 0 || d <= 0) return DFB200_OK;
  coupling_bwd_kernel<<<cdiv(B * d, 256), 256, 0, as_stream(stream)>>>(B, d, s_t, x2, ldx, dy1, ldy, dlogdet, ds_t, dx2, lddx);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

// ---------------------------------------------------------------------------------------------------------------------------
// Adam for ALL parameter tensors of a group in one launch (torch.optim.Adam semantics, no amsgrad; the reference builds
// torch.optim.Adam from the config, runner.py:60-66).  torch's fused multi-tensor Adam needs 8 launches of ~22 us for the ~130
// small tensors of the denoiser plus 3 for the per-tensor step counters (0.25 ms of a 3 ms step); here the tensor table travels as a
// kernel parameter (up to ADAM_MAX_TENSORS entries per launch), one CTA per 2 048-element chunk finds its tensor by binary search,
// and the step count is ONE device scalar (so the launch is capturable in a CUDA graph).
// ---------------------------------------------------------------------------------------------------------------------------
namespace dfb200 {
constexpr int ADAM_MAX_TENSORS = 320, ADAM_CHUNK = 2048, ADAM_THREADS = 256;
struct AdamBatch {
  dfb200_adam_tensor_t t[ADAM_MAX_TENSORS];
  int first_chunk[ADAM_MAX_TENSORS + 1];
  int n;
};
__global__ void __launch_bounds__(ADAM_THREADS)
adam_kernel(const __grid_constant__ AdamBatch B, const long long* __restrict__ step, float lr, float beta1, float beta2, float eps,
            float weight_decay, const float* __restrict__ grad_scale) {
  __shared__ float s_c[2];
  if (threadIdx.x == 0) {  // bias corrections in double, as torch does on the host for the non-capturable form
    const double t = (double)*step;
    s_c[0] = (float)(1.0 / (1.0 - pow((double)beta1, t)));        // 1 / bias_correction1
    s_c[1] = (float)(1.0 / sqrt(1.0 - pow((double)beta2, t)));    // 1 / sqrt(bias_correction2)
  }
  int lo = 0, hi = B.n;  // last tensor with first_chunk <= blockIdx.x
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (B.first_chunk[mid] <= (int)blockIdx.x) lo = mid; else hi = mid;
  }
  const dfb200_adam_tensor_t T = B.t[lo];
  const long long base = (long long)((int)blockIdx.x - B.first_chunk[lo]) * ADAM_CHUNK;
  const int cnt = (int)min((long long)ADAM_CHUNK, T.count - base);
  float* p = T.param + base;
  const float* g = T.grad + base;
  float* m = T.exp_avg + base;
  float* v = T.exp_avg_sq + base;
  __syncthreads();
  const float step_size = lr * s_c[0], rsb2 = s_c[1];
  const float gs = grad_scale != nullptr ? __ldg(grad_scale) : 1.f;
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg = gg * gs + weight_decay * pp;
    mm = mm + (gg - mm) * (1.f - beta1);            // exp_avg.lerp_(grad, 1 - beta1)
    vv = vv * beta2 + (1.f - beta2) * gg * gg;      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    pp -= step_size * mm / (sqrtf(vv) * rsb2 + eps);
  };
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0;
  if (vec) {
    for (int i = threadIdx.x * 4; i + 3 < cnt; i += ADAM_THREADS * 4) {
      float4 pp = *reinterpret_cast<float4*>(p + i), mm = *reinterpret_cast<float4*>(m + i), vv = *reinterpret_cast<float4*>(v + i);
      const float4 gg = *reinterpret_cast<const float4*>(g + i);
      upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
      *reinterpret_cast<float4*>(p + i) = pp; *reinterpret_cast<float4*>(m + i) = mm; *reinterpret_cast<float4*>(v + i) = vv;
    }
    for (int i = (cnt & ~3) + threadIdx.x; i < cnt; i += ADAM_THREADS) upd(p[i], g[i], m[i], v[i]);
  } else {
    for (int i = threadIdx.x; i < cnt; i += ADAM_THREADS) upd(p[i], g[i], m[i], v[i]);
  }
}
}  // namespace dfb200

extern "C" int dfb200_adam_step(int n_tensors, const dfb200_adam_tensor_t* tensors, const long long* step, float lr, float beta1, float beta2,
                                float eps, float weight_decay, const float* grad_scale, dfb200_stream_t stream) {
  DFB_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || tensors != nullptr) && step != nullptr, DFB200_ERR_INVALID_ARG, "adam_step: bad arguments");
  DFB_REQUIRE(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, DFB200_ERR_INVALID_ARG,
              "adam_step: betas must be in [0, 1) and eps >= 0 (got %g, %g, %g)", (double)beta1, (double)beta2, (double)eps);
  static dfb200::AdamBatch batch;  // 14 KB: built on the host, passed by value
  static std::mutex mu;
  std::lock_guard<std::mutex> lock(mu);
  int done = 0;
  while (done < n_tensors) {
    int n = 0, chunks = 0;
    while (done < n_tensors && n < dfb200::ADAM_MAX_TENSORS) {
      const dfb200_adam_tensor_t& t = tensors[done++];
      DFB_REQUIRE(t.count >= 0 && (t.count == 0 || (t.param && t.grad && t.exp_avg && t.exp_avg_sq)), DFB200_ERR_INVALID_ARG,
                  "adam_step: tensor %d has a null pointer or a negative count", done - 1);
      if (t.count == 0) continue;
      const long long c = (t.count + dfb200::ADAM_CHUNK - 1) / dfb200::ADAM_CHUNK;
      DFB_REQUIRE(chunks + c < 0x7fffffffLL, DFB200_ERR_INVALID_ARG, "adam_step: too many elements in one launch");
      batch.t[n] = t;
      batch.first_chunk[n] = chunks;
      chunks += (int)c;
      ++n;
    }
    if (n == 0) continue;
    batch.first_chunk[n] = chunks;
    batch.n = n;
    dfb200::adam_kernel<<<chunks, dfb200::ADAM_THREADS, 0, as_stream(stream)>>>(batch, step, lr, beta1, beta2, eps, weight_decay, grad_scale);
    DFB_LAUNCH_CHECK();
  }
  return DFB200_OK;
}
