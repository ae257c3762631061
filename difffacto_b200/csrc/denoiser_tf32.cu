// DFB200_MODE_TF32: the cross-diffusion denoiser on the tensor cores at REFERENCE tolerance.
//
// Reference computation: python/difffacto/models/diffusions/nets/attention.py:385-440 (+ blocks :161-306); the reference runs
// it in true fp32 (cuBLAS, TF32 off).  The bf16 kernel (denoiser_tc.cu) rounds every GEMM operand to 8 mantissa bits
// (|eps - reference| ~ 6e-3); this kernel keeps the same fused structure but feeds the tensor cores fp32 containers
// (tcgen05.mma kind::tf32: 10 explicit mantissa bits, round-to-nearest applied when an operand is staged, fp32 accumulation
// in TMEM) and evaluates everything that is not a dense contraction in fp32 on CUDA cores:
//   * one 128-token tile per CTA pass, residual stream x (fp32) in TENSOR MEMORY (128 lanes x 128 columns); the attention-out
//     and FF-out GEMMs accumulate into it;
//   * proj_in (13 -> 128) + pre_norm and post_norm + proj_out (128 -> 3) are fp32 CUDA-core code (they are 0.2 % of the FLOPs);
//   * cross-attention folded per (sample, block) into W_sim (32 x 128) / W_pv (128 x 32) as in the bf16 kernel; the logit
//     bias is added in fp32 by the softmax, bo / b2 are K = 8 MMAs against a ones tile with the bias split hi + lo in tf32;
//   * GEGLU with the reference's exact-CDF GELU to 2.7e-5 (sigmoid of a fitted odd polynomial, ex2 / rcp MUFU forms), bias b1' in fp32;
//   * the hidden chunk (64 value + 64 gate columns) is DOUBLE-BUFFERED in TMEM, so FF-in of chunk c+1 runs on the tensor pipe
//     while all 8 epilogue warps apply GEGLU to chunk c; the gated activations go back to TMEM over the value columns just read;
//   * both feed-forward GEMMs run in TS form (A operand in tensor memory): the LayerNorm-3 output is written to TMEM columns
//     [384,512) and read by all 8 FF-in chunks, FF-out reads the gated activations from the hidden buffer -- 128 KB less
//     shared-memory traffic per chunk than the SS forms of the first version (4-byte operands make this kernel shared-memory bound).
// TMEM map (512 columns): X [0,128)  H0 [128,256)  H1 [256,384)  S [384,416) during attention / LN3 output A [384,512) during FF.
// Weights stream L2 -> smem as pre-packed tf32 UMMA tiles (16 KB quarter / half chunks) through a 6-slot cp.async.bulk ring.
#include <float.h>

#include "denoiser.cuh"
#include "tc_common.cuh"

namespace dfb200 {
using namespace tc;

namespace t32 {
constexpr int SLOT = 18432;          // 16 KB tile + 2 KB slab
constexpr int NSLOT = 6;
constexpr int STATIC_PER_LAYER = 48; // 8 x (4 W1' quarters + 2 W2 halves)
constexpr int PKT_PER_LAYER = 50;    // + the two fold packets
constexpr int FF_CHUNKS = 8;
constexpr int SLAB_OFF = 16384;
__host__ __device__ inline int w1q(int c, int q) { return c == 0 ? q : 4 + 6 * (c - 1) + q; }      // static index of W1'_c, k [32q, 32q+32)
__host__ __device__ inline int w2h(int c, int h) { return c < 7 ? 4 + 6 * c + 4 + h : 46 + h; }    // static index of W2_c, k [32h, 32h+32)
// layer-local packet p (0..49) -> bytes
__host__ __device__ inline int pkt_bytes(int p) {
  if (p == 0) return 16384 + 128;   // W_sim + b_sim (32 fp32)
  if (p == 1) return SLOT;          // W_pv + bo slab
  return p - 2 == 47 ? SLOT : 16384;  // W2_7 second half carries the b2 slab
}
// extras after the static stream (fp32): WinT[13][128] | b_in[128] | pre_g[128] | pre_b[128] | hw[3][128] | hb[4]
constexpr int EX_WINT = 0, EX_BIN = 13 * 128, EX_PREG = EX_BIN + 128, EX_PREB = EX_PREG + 128, EX_HW = EX_PREB + 128, EX_HB = EX_HW + 3 * 128;
constexpr int EX_FLOATS = EX_HB + 4;

constexpr int THREADS = 320;  // warps 0-3: token rows (all phases), 4-7: second column half of the GEGLU, 8: MMA issuer, 9: weight producer
constexpr uint32_t SM_A = 0;                               // 65536  A operand tile (128 x 128 tf32)
constexpr uint32_t SM_U = 65536;                           // 32768  gated activations (128 x 64 tf32); attention probabilities (128 x 32)
constexpr uint32_t SM_RING = 98304;                        // NSLOT x SLOT
constexpr uint32_t SM_ONES = SM_RING + NSLOT * SLOT;       // 4096   ones tile (128 x 8 tf32: k = 0,1 -> 1)
constexpr uint32_t SM_EX = SM_ONES + 4096;                 // extras image
constexpr uint32_t SM_BAR = SM_EX + ((EX_FLOATS * 4 + 127) / 128) * 128;
constexpr uint32_t SM_TMEM = SM_BAR + 256;
constexpr uint32_t SMEM_BYTES = SM_BAR + 512;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");
enum Bar { BAR_A = 0, BAR_S = 1, BAR_X = 2, BAR_ACC = 3 /*[2]*/, BAR_UREADY = 5 /*[2], one per accumulator buffer*/, BAR_WFULL = 7 /*[NSLOT]*/,
           BAR_WEMPTY = 7 + NSLOT, BAR_COUNT = 7 + 2 * NSLOT };
static_assert(BAR_COUNT * 8 <= 256, "barrier block");
}  // namespace t32

size_t tf32_stream_bytes_for(const NetDims& d) {
  return (size_t)d.depth * t32::STATIC_PER_LAYER * t32::SLOT + sizeof(float) * (((size_t)t32::EX_FLOATS + 63) & ~(size_t)63);
}
size_t tf32_fold_bytes_for(const NetDims& d, int B) { return (size_t)B * d.depth * 2 * t32::SLOT; }

// ---- pack kernels -------------------------------------------------------------------------------------------------
// dst: tf32 UMMA tile of R rows x KC k-values (values rounded to nearest tf32); element (r,k) = src[rowmap(r)*ld + k0 + k] * gamma[k0+k]
__global__ void pack_tile32_kernel(uint8_t* __restrict__ dst, int R, int KC, const float* __restrict__ src, int ld, int k0,
                                   const float* __restrict__ gamma, int geglu_chunk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * KC) return;
  const int r = i / KC, k = i - r * KC;
  int row = r;
  if (geglu_chunk >= 0) row = r < 64 ? 64 * geglu_chunk + r : D_FF + 64 * geglu_chunk + (r - 64);
  float v = __ldg(src + (size_t)row * ld + k0 + k);
  if (gamma != nullptr) v *= __ldg(gamma + k0 + k);
  if (geglu_chunk >= 0 && r < 64) v *= 0.5f;  // a * gelu(g) = (a/2) * g * (1 + erf(g / sqrt 2))
  *reinterpret_cast<float*>(dst + tile_off32(R, r, k)) = to_tf32(v);
}
// bias slab: R rows x 16 B; k = 0: tf32(b), k = 1: tf32(b - hi)
__global__ void pack_bias32_kernel(uint8_t* __restrict__ dst, int R, const float* __restrict__ bias) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const float v = __ldg(bias + r);
  const float hi = to_tf32(v);
  *reinterpret_cast<float4*>(dst + r * 16) = make_float4(hi, to_tf32(v - hi), 0.f, 0.f);
}
__global__ void __launch_bounds__(128)
pack_extras32_kernel(float* __restrict__ ex, const float* __restrict__ w_in, const float* __restrict__ b_in, const float* __restrict__ pre_w,
                     const float* __restrict__ pre_b, const float* __restrict__ w_out, const float* __restrict__ b_out,
                     const float* __restrict__ post_w, const float* __restrict__ post_b) {
  const int n = threadIdx.x;
  for (int i = 0; i < 13; ++i) ex[t32::EX_WINT + i * 128 + n] = __ldg(w_in + n * 13 + i);
  ex[t32::EX_BIN + n] = __ldg(b_in + n);
  ex[t32::EX_PREG + n] = __ldg(pre_w + n);
  ex[t32::EX_PREB + n] = __ldg(pre_b + n);
  for (int c = 0; c < 3; ++c) ex[t32::EX_HW + c * 128 + n] = __ldg(w_out + c * D_MODEL + n) * __ldg(post_w + n);
  if (n < 4) {
    float v = 0.f;
    if (n < 3) {
      v = __ldg(b_out + n);
      for (int k = 0; k < D_MODEL; ++k) v = fmaf(__ldg(w_out + n * D_MODEL + k), __ldg(post_b + k), v);
    }
    ex[t32::EX_HB + n] = v;
  }
}

int tf32_pack_stream(const PackLayout& L, void* packed, cudaStream_t st) {
  using namespace t32;
  const float* P = reinterpret_cast<const float*>(packed);
  uint8_t* S = reinterpret_cast<uint8_t*>(packed) + L.tf32_stream_off;
  auto tile = [&](uint8_t* dst, int R, int KC, const float* src, int ld, int k0, const float* gamma, int chunk) {
    pack_tile32_kernel<<<cdiv(R * KC, 256), 256, 0, st>>>(dst, R, KC, src, ld, k0, gamma, chunk);
    count_launch();
  };
  for (int l = 0; l < L.d.depth; ++l) {
    const size_t* o = L.blk[l];
    uint8_t* base = S + (size_t)l * STATIC_PER_LAYER * SLOT;
    for (int c = 0; c < FF_CHUNKS; ++c) {
      for (int q = 0; q < 4; ++q) tile(base + (size_t)w1q(c, q) * SLOT, 128, 32, P + o[B_W1], D_MODEL, 32 * q, P + o[B_N3_W], c);
      for (int h = 0; h < 2; ++h) tile(base + (size_t)w2h(c, h) * SLOT, 128, 32, P + o[B_W2], D_FF, 64 * c + 32 * h, nullptr, -1);
    }
    pack_bias32_kernel<<<1, 128, 0, st>>>(base + (size_t)w2h(FF_CHUNKS - 1, 1) * SLOT + SLAB_OFF, 128, P + o[B_B2]);
    count_launch();
  }
  float* ex = reinterpret_cast<float*>(S + (size_t)L.d.depth * STATIC_PER_LAYER * SLOT);
  pack_extras32_kernel<<<1, 128, 0, st>>>(ex, P + L.g[P_IN_W], P + L.g[P_IN_B], P + L.g[P_PRE_W], P + L.g[P_PRE_B], P + L.g[P_OUT_W],
                                          P + L.g[P_OUT_B], P + L.g[P_POST_W], P + L.g[P_POST_B]);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

// ---- per-forward fold kernel (tf32 tiles): see context_fold_kernel in denoiser_tc.cu for the algebra -------------------------
//   pkt0: W_sim[(h,j)][k] * log2(e) as a 32 x 128 tile | b_sim[(h,j)] * log2(e) (32 fp32)      pkt1: W_pv[c][(h,j)] as a 128 x 32 tile | bo slab
__global__ void __launch_bounds__(256)
context_fold32_kernel(int depth, const float* __restrict__ kv, const float* __restrict__ foldw_all, size_t foldw_stride,
                      const float* __restrict__ bo0, size_t bo_stride, uint8_t* __restrict__ fold_all) {
  constexpr float LOG2E = 1.4426950408889634f;
  __shared__ float K[MAX_TOKENS][D_MODEL], V[MAX_TOKENS][D_MODEL];
  const int b = blockIdx.x, l = blockIdx.y, t = threadIdx.x;
  const float* src = kv + ((size_t)b * depth + l) * 1024;
  for (int i = t; i < 512; i += 256) {
    (&K[0][0])[i] = __ldg(src + i);
    (&V[0][0])[i] = __ldg(src + 512 + i);
  }
  __syncthreads();
  const float* WqG = foldw_all + (size_t)l * foldw_stride;
  const float* bqG = WqG + D_MODEL * D_MODEL;
  const float* WoT = bqG + D_MODEL;
  uint8_t* p0 = fold_all + ((size_t)b * depth + l) * 2 * t32::SLOT;
  uint8_t* p1 = p0 + t32::SLOT;
  const int col = t & 127, half = t >> 7;
  for (int hh = 0; hh < 4; ++hh) {
    const int h = half * 4 + hh;
    float acc[MAX_TOKENS] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      const float w = __ldg(WqG + (16 * h + d) * D_MODEL + col);
#pragma unroll
      for (int j = 0; j < MAX_TOKENS; ++j) acc[j] = fmaf(K[j][16 * h + d], w, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < MAX_TOKENS; ++j) *reinterpret_cast<float*>(p0 + tile_off32(32, h * 4 + j, col)) = to_tf32(LOG2E * acc[j]);
  }
  for (int hh = 0; hh < 4; ++hh) {
    const int h = half * 4 + hh;
    float acc[MAX_TOKENS] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < 16; ++d) {
      const float w = __ldg(WoT + (16 * h + d) * D_MODEL + col);
#pragma unroll
      for (int j = 0; j < MAX_TOKENS; ++j) acc[j] = fmaf(V[j][16 * h + d], w, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < MAX_TOKENS; ++j) *reinterpret_cast<float*>(p1 + tile_off32(128, col, h * 4 + j)) = to_tf32(acc[j]);
  }
  if (t < 32) {
    const int h = t >> 2, j = t & 3;
    float v = 0.f;
    for (int d = 0; d < 16; ++d) v = fmaf(K[j][16 * h + d], __ldg(bqG + 16 * h + d), v);
    reinterpret_cast<float*>(p0 + t32::SLAB_OFF)[t] = LOG2E * v;
  }
  if (t < 128) {
    const float v = __ldg(bo0 + (size_t)l * bo_stride + t);
    const float hi = to_tf32(v);
    *reinterpret_cast<float4*>(p1 + t32::SLAB_OFF + t * 16) = make_float4(hi, to_tf32(v - hi), 0.f, 0.f);
  }
}

// ---- the fused kernel ----------------------------------------------------------------------------------------------
struct Tf32Params {
  const uint8_t* stream; const uint8_t* fold; const float* extras; const float* b1p;
  const float* x; const float* anchors; const float* variances; const int* assign; const float* valid;
  float* eps_out;
  int N, depth, flags;
  long long M;
};

// GEGLU for two columns at fp32 accuracy: returns (a_half + ba) * (g + bg) * 2 Phi(g + bg), rounded to tf32.  The reference's
// F.gelu is x Phi(x) with the exact normal CDF.  Phi(x) = 1 / (1 + exp(-2 q(x))) holds exactly for q = atanh(erf(x / sqrt 2)), an
// odd function; q(x) = x (c0 + c1 x^2 + c2 x^4 + c3 x^6) fitted on |x| <= 12 reproduces the erf GELU to max |error| 2.7e-5
// (a tenth of the tf32 rounding of the result; the inner polynomial stays >= 0.5 for every x, so the form is sign-safe), and
// ex2 / rcp are the ~1e-7 MUFU forms -- unlike tanh.approx (5e-4).  16 instructions and 4 MUFU per column pair; the
// Abramowitz-Stegun erf of the first version took ~30 and bound the kernel (ncu: issue slots 49 %, tensor pipe 29 %).
__device__ __forceinline__ float2 geglu_erf2(float2 a_half, float2 g, float2 ba, float2 bg) {
  constexpr float K = -2.8853900817779268f;  // -2 log2(e)
  constexpr float k0 = K * 0.7974859391733992f, k1 = K * 0.037037304123455336f, k2 = K * -0.0003620539674771765f,
                  k3 = K * 8.837883123174581e-07f;
  g = __fadd2_rn(g, bg);
  a_half = __fadd2_rn(a_half, ba);
  const float2 m = __fmul2_rn(g, g);
  float2 p = __ffma2_rn(m, make_float2(k3, k3), make_float2(k2, k2));
  p = __ffma2_rn(p, m, make_float2(k1, k1));
  p = __ffma2_rn(p, m, make_float2(k0, k0));
  const float2 arg = __fmul2_rn(g, p);  // -2 q(g) log2 e
  float ex, ey;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(arg.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ey) : "f"(arg.y));
  const float2 den = __ffma2_rn(make_float2(ex, ey), make_float2(0.5f, 0.5f), make_float2(0.5f, 0.5f));  // (1 + e) / 2; inf -> rcp = 0
  float rx, ry;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rx) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(ry) : "f"(den.y));
  // (one reciprocal of the product for both columns -- 3 MUFU per pair -- measured slower: 80.2 vs 82.7 shapes/s; the phase is
  //  issue / latency bound, not MUFU bound)
  const float2 u = __fmul2_rn(__fmul2_rn(a_half, g), make_float2(rx, ry));
  return make_float2(to_tf32(u.x), to_tf32(u.y));
}
// D[tmem] += A[tmem] * B[smem]^T, kind::tf32, A operand in TENSOR MEMORY (lane = row, one 32-bit k-value per column, 8 columns per
// K = 8 step).  Measured 96 cycles per M128 x N128 x K8 MMA against 105 in SS form (tools/micro/umma_tf32_ts.cu), and no
// shared-memory read of A.
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void row_stats32(uint32_t taddr, float& mean, float& rstd) {
  float s = 0.f, q = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {
    float h[32];
    tmem_ld32(taddr + cb * 32, h);
    tmem_wait_ld();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      s += h[2 * k]; q = fmaf(h[2 * k], h[2 * k], q);
      s1 += h[2 * k + 1]; q1 = fmaf(h[2 * k + 1], h[2 * k + 1], q1);
    }
  }
  mean = (s + s1) * (1.f / D_MODEL);
  // two-pass-quality variance is not needed: |mean| << std on LayerNorm'ed residual streams; clamp guards cancellation
  const float var = fmaxf((q + q1) * (1.f / D_MODEL) - mean * mean, 0.f);
  rstd = rsqrtf(var + LN_EPS);
}
// LayerNorm (gain / bias folded into the following weights) of the TMEM row -> tf32 A-operand row
__device__ __forceinline__ void row_layernorm_to_tile32(uint32_t taddr, uint8_t* tile, int r) {
  float mean, rstd;
  row_stats32(taddr, mean, rstd);
  const float nm = -mean * rstd;
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {
    float h[32];
    tmem_ld32(taddr + cb * 32, h);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(tile + (cb * 8 + j) * 2048 + r * 16) =
          make_float4(to_tf32(fmaf(h[4 * j], rstd, nm)), to_tf32(fmaf(h[4 * j + 1], rstd, nm)), to_tf32(fmaf(h[4 * j + 2], rstd, nm)),
                      to_tf32(fmaf(h[4 * j + 3], rstd, nm)));
  }
}

// LayerNorm of the TMEM row -> tf32 A-operand row in TENSOR MEMORY (columns [dst, dst + 128)): the FF-in GEMM reads it in TS form
__device__ __forceinline__ void row_layernorm_to_tmem32(uint32_t taddr, uint32_t dst) {
  float mean, rstd;
  row_stats32(taddr, mean, rstd);
  const float nm = -mean * rstd;
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {
    float h[32];
    tmem_ld32(taddr + cb * 32, h);
    tmem_wait_ld();
#pragma unroll
    for (int k = 0; k < 32; ++k) h[k] = to_tf32(fmaf(h[k], rstd, nm));
    tmem_st32(dst + cb * 32, h);
  }
  tmem_wait_st();
}

__global__ void __launch_bounds__(t32::THREADS, 1) denoiser_tf32_kernel(const Tf32Params P) {
  using namespace t32;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_TMEM);
  const float* ex = reinterpret_cast<const float*>(smem + SM_EX);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);

  if (warp == 8 && lane == 0) {
    mbar_init(&bars[BAR_A], 128); mbar_init(&bars[BAR_S], 1); mbar_init(&bars[BAR_X], 1);
    mbar_init(&bars[BAR_ACC], 1); mbar_init(&bars[BAR_ACC + 1], 1);
    mbar_init(&bars[BAR_UREADY], 256);
    mbar_init(&bars[BAR_UREADY + 1], 256);
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&bars[BAR_WFULL + i], 1); mbar_init(&bars[BAR_WEMPTY + i], 1); }
    fence_barrier_init();
  }
  if (tid < 256) {  // ones tile: slab 0 (k 0..3) = {1,1,0,0}, slab 1 (k 4..7) = 0
    const float one = tid < 128 ? 1.f : 0.f;
    *reinterpret_cast<float4*>(smem + SM_ONES + tid * 16) = make_float4(one, one, 0.f, 0.f);
  }
  for (int i = tid; i < EX_FLOATS; i += THREADS) reinterpret_cast<float*>(smem + SM_EX)[i] = __ldg(P.extras + i);
  fence_proxy_async();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tmem != 0u) __trap();
  const long long n_tiles = P.M / 128;

  if (warp < 8) {
    const int H = warp >> 2, r = tid & 127;   // H: column half in the GEGLU phase; rows belong to warps 0-3 elsewhere
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t X = lane_base, S = lane_base + 384;
    uint8_t* a_tile = smem + SM_A;
    uint8_t* u_tile = smem + SM_U;
    uint32_t ph_s = 0, ph_x = 0, ph_acc0 = 0, ph_acc1 = 0;
    int G = 0;  // packets consumed so far (this CTA), to find the fold packet of a layer
#pragma unroll 1
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long tok = tile * 128 + r;
      const long long b = tok / P.N;
      const int p = (int)(tok - b * P.N);
      uint32_t vmask = 0;
      if (H == 0) {
#pragma unroll
        for (int j = 0; j < MAX_TOKENS; ++j)
          if (P.valid == nullptr || __ldg(P.valid + b * MAX_TOKENS + j) != 0.f) vmask |= 1u << j;
        // ---- proj_in (13 -> 128) + pre_norm in fp32: x = LN(W f + b) -> TMEM ----
        float f[9];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          f[c] = __ldg(P.x + (b * 3 + c) * P.N + p);
          f[3 + c] = __ldg(P.anchors + (b * 3 + c) * P.N + p);
          const float v = __ldg(P.variances + (b * 3 + c) * P.N + p);
          f[6 + c] = (P.flags & DFB200_NET_INCLUDE_STD) ? sqrtf(v) : v;
        }
        const int part = __ldg(P.assign + tok);
        float s = 0.f, q = 0.f;
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
          float h[32];
#pragma unroll
          for (int n4 = 0; n4 < 8; ++n4) {
            const int n = cb * 32 + n4 * 4;
            float4 acc = *reinterpret_cast<const float4*>(ex + EX_BIN + n);
            const float4 wc = *reinterpret_cast<const float4*>(ex + EX_WINT + (9 + part) * 128 + n);
            acc.x += wc.x; acc.y += wc.y; acc.z += wc.z; acc.w += wc.w;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
              const float4 w = *reinterpret_cast<const float4*>(ex + EX_WINT + i * 128 + n);
              acc.x = fmaf(w.x, f[i], acc.x); acc.y = fmaf(w.y, f[i], acc.y); acc.z = fmaf(w.z, f[i], acc.z); acc.w = fmaf(w.w, f[i], acc.w);
            }
            h[n4 * 4] = acc.x; h[n4 * 4 + 1] = acc.y; h[n4 * 4 + 2] = acc.z; h[n4 * 4 + 3] = acc.w;
          }
#pragma unroll
          for (int k = 0; k < 32; ++k) { s += h[k]; q = fmaf(h[k], h[k], q); }
          tmem_st32(X + cb * 32, h);
        }
        tmem_wait_st();
        const float mean = s * (1.f / D_MODEL);
        const float rstd = rsqrtf(fmaxf(q * (1.f / D_MODEL) - mean * mean, 0.f) + LN_EPS);
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
          float h[32];
          tmem_ld32(X + cb * 32, h);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 32; ++k) h[k] = fmaf((h[k] - mean) * rstd, ex[EX_PREG + cb * 32 + k], ex[EX_PREB + cb * 32 + k]);
          tmem_st32(X + cb * 32, h);
        }
        tmem_wait_st();
      }
#pragma unroll 1
      for (int l = 0; l < P.depth; ++l) {
        const int G0 = G;
        G += PKT_PER_LAYER;
        if (H == 0) {
          // ---- LN2 -> A ----
          row_layernorm_to_tile32(X, a_tile, r);
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(&bars[BAR_A]);
          // ---- logits (x log2 e) -> + b_sim -> softmax over the 4 part tokens per head -> P tile ----
          mbar_wait(&bars[BAR_S], ph_s);
          ph_s ^= 1;
          tc_fence_after();
          mbar_wait(&bars[BAR_WFULL + G0 % NSLOT], (uint32_t)(G0 / NSLOT) & 1u);  // the fold packet (b_sim) is visible to this thread
          const float* bsim = reinterpret_cast<const float*>(smem + SM_RING + (G0 % NSLOT) * SLOT + SLAB_OFF);
          float sv[32];
          tmem_ld32(S, sv);
          tmem_wait_ld();
          float pr[32];
#pragma unroll
          for (int h = 0; h < 8; ++h) {
            float s0 = (vmask & 1u) ? sv[4 * h] + bsim[4 * h] : -FLT_MAX, s1 = (vmask & 2u) ? sv[4 * h + 1] + bsim[4 * h + 1] : -FLT_MAX;
            float s2 = (vmask & 4u) ? sv[4 * h + 2] + bsim[4 * h + 2] : -FLT_MAX, s3 = (vmask & 8u) ? sv[4 * h + 3] + bsim[4 * h + 3] : -FLT_MAX;
            const float mx = fmaxf(fmaxf(s0, s1), fmaxf(s2, s3));
            s0 = ex2f(s0 - mx); s1 = ex2f(s1 - mx); s2 = ex2f(s2 - mx); s3 = ex2f(s3 - mx);
            const float inv = 1.f / ((s0 + s1) + (s2 + s3));
            pr[4 * h] = to_tf32(s0 * inv); pr[4 * h + 1] = to_tf32(s1 * inv); pr[4 * h + 2] = to_tf32(s2 * inv); pr[4 * h + 3] = to_tf32(s3 * inv);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(u_tile + j * 2048 + r * 16) = make_float4(pr[4 * j], pr[4 * j + 1], pr[4 * j + 2], pr[4 * j + 3]);
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(&bars[BAR_A]);
          // ---- x += P W_pv^T + bo;  LN3 -> A ----
          mbar_wait(&bars[BAR_X], ph_x);
          ph_x ^= 1;
          tc_fence_after();
          row_layernorm_to_tmem32(X, lane_base + 384);  // LN3 output: the TS-form A operand of all 8 FF-in chunks (the logits are dead)
          tc_fence_before();
          mbar_arrive(&bars[BAR_A]);
        }
        // ---- GEGLU feed-forward: 8 chunks of 64 value + 64 gate columns, all 8 warps (column halves) ----
        const float4* b1p = reinterpret_cast<const float4*>(P.b1p + (size_t)l * 2 * D_FF) + H * 8;
#pragma unroll 1
        for (int c = 0; c < FF_CHUNKS; ++c) {
          float ba[32], bg[32];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 va = __ldg(b1p + c * 32 + k), vg = __ldg(b1p + c * 32 + 16 + k);
            ba[4 * k] = va.x; ba[4 * k + 1] = va.y; ba[4 * k + 2] = va.z; ba[4 * k + 3] = va.w;
            bg[4 * k] = vg.x; bg[4 * k + 1] = vg.y; bg[4 * k + 2] = vg.z; bg[4 * k + 3] = vg.w;
          }
          const uint32_t ACC = lane_base + 128 + (c & 1) * 128 + H * 32;
          if (c & 1) { mbar_wait(&bars[BAR_ACC + 1], ph_acc1); ph_acc1 ^= 1; }
          else { mbar_wait(&bars[BAR_ACC], ph_acc0); ph_acc0 ^= 1; }
          tc_fence_after();
          float a[32], gt[32];
          tmem_ld32(ACC, a);
          tmem_ld32(ACC + 64, gt);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float2 u = geglu_erf2(make_float2(a[2 * k], a[2 * k + 1]), make_float2(gt[2 * k], gt[2 * k + 1]),
                                        make_float2(ba[2 * k], ba[2 * k + 1]), make_float2(bg[2 * k], bg[2 * k + 1]));
            a[2 * k] = u.x; a[2 * k + 1] = u.y;
          }
          // the gated activations go back to TENSOR MEMORY over the value columns this thread has just consumed; FF-out reads them
          // in TS form (the first version wrote a 32 KB shared-memory tile per chunk and waited for the previous FF-out to free it)
          tmem_st32(ACC, a);
          tmem_wait_st();
          tc_fence_before();
          // one barrier PER BUFFER: the two column halves (warps 0-3 / 4-7) may be a chunk apart, and on a single barrier 128
          // arrivals for chunk c plus 128 for chunk c+1 from the same half would complete the phase without the other half
          mbar_arrive(&bars[BAR_UREADY + (c & 1)]);
        }
        if (H == 0) {
          mbar_wait(&bars[BAR_X], ph_x);
          ph_x ^= 1;
          tc_fence_after();
        }
      }
      if (H == 0) {
        // ---- post_norm + proj_out (128 -> 3) in fp32 ----
        float mean, rstd;
        row_stats32(X, mean, rstd);
        float e0 = 0.f, e1 = 0.f, e2 = 0.f;
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
          float h[32];
          tmem_ld32(X + cb * 32, h);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float y = (h[k] - mean) * rstd;
            e0 = fmaf(y, ex[EX_HW + cb * 32 + k], e0);
            e1 = fmaf(y, ex[EX_HW + 128 + cb * 32 + k], e1);
            e2 = fmaf(y, ex[EX_HW + 256 + cb * 32 + k], e2);
          }
        }
        P.eps_out[(b * 3 + 0) * P.N + p] = e0 + ex[EX_HB];
        P.eps_out[(b * 3 + 1) * P.N + p] = e1 + ex[EX_HB + 1];
        P.eps_out[(b * 3 + 2) * P.N + p] = e2 + ex[EX_HB + 2];
        tc_fence_before();
      }
    }
    tc_fence_before();
  } else if (warp == 8) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc128 = make_idesc_tf32(128, 128), idesc32 = make_idesc_tf32(128, 32);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t ring = sbase + SM_RING, a_base = sbase + SM_A, u_base = sbase + SM_U;
    const uint64_t ones_desc = make_smem_desc(sbase + SM_ONES, 2048, TILE_SBO);
    uint32_t ph_a = 0, ph_u = 0;  // ph_u: bit b = parity of BAR_UREADY[b]
    int G = 0;
    auto pkt = [&](int g) -> uint32_t {
      mbar_wait(&bars[BAR_WFULL + g % NSLOT], (uint32_t)(g / NSLOT) & 1u);
      return ring + (uint32_t)(g % NSLOT) * SLOT;
    };
    auto wait_a = [&]() { mbar_wait(&bars[BAR_A], ph_a); ph_a ^= 1; tc_fence_after(); };
    // H_c = LN3(x) W1'_c^T -> hidden buffer c & 1: four k-quarter packets of 4 K-steps each
    auto ff_in = [&](int c, int g_first) {
      const uint32_t d = 128 + (c & 1) * 128;
      for (int q = 0; q < 4; ++q) {
        const uint32_t pw = pkt(g_first + q);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32_ts(d, 384 + (4 * q + ks) * 8, make_smem_desc(pw + ks * 4096, 2048, TILE_SBO), idesc128, (q | ks) ? 1u : 0u);
          umma_commit(&bars[BAR_WEMPTY + (g_first + q) % NSLOT]);
          if (q == 3) umma_commit(&bars[BAR_ACC + (c & 1)]);
        }
        __syncwarp();
      }
    };
#pragma unroll 1
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int l = 0; l < P.depth; ++l) {
        const int G0 = G;
        G += PKT_PER_LAYER;
        // ---- logits: S = LN2(x) W_sim^T -> columns [384, 416) ----
        wait_a();
        const uint32_t pf0 = pkt(G0);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 16; ++ks)
            umma_tf32(384, make_smem_desc(a_base + ks * 4096, 2048, TILE_SBO), make_smem_desc(pf0 + ks * 1024, 512, TILE_SBO), idesc32, ks ? 1u : 0u);
          umma_commit(&bars[BAR_S]);
        }
        __syncwarp();
        // ---- x += P W_pv^T + bo ----
        wait_a();
        const uint32_t pf1 = pkt(G0 + 1);
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_tf32(0, make_smem_desc(u_base + ks * 4096, 2048, TILE_SBO), make_smem_desc(pf1 + ks * 4096, 2048, TILE_SBO), idesc128, 1u);
          umma_tf32(0, ones_desc, make_smem_desc(pf1 + SLAB_OFF, 0, TILE_SBO), idesc128, 1u);
          umma_commit(&bars[BAR_X]);
          umma_commit(&bars[BAR_WEMPTY + G0 % NSLOT]);
          umma_commit(&bars[BAR_WEMPTY + (G0 + 1) % NSLOT]);
        }
        __syncwarp();
        // ---- feed-forward: FF-in of chunk c+1 is issued before FF-out of chunk c waits for its gated activations ----
        wait_a();
        ff_in(0, G0 + 2 + w1q(0, 0));
#pragma unroll 1
        for (int c = 0; c < FF_CHUNKS; ++c) {
          if (c + 1 < FF_CHUNKS) ff_in(c + 1, G0 + 2 + w1q(c + 1, 0));
          mbar_wait(&bars[BAR_UREADY + (c & 1)], (ph_u >> (c & 1)) & 1u);
          ph_u ^= 1u << (c & 1);
          tc_fence_after();
          for (int h = 0; h < 2; ++h) {
            const int g = G0 + 2 + w2h(c, h);
            const uint32_t pw = pkt(g);
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                umma_tf32_ts(0, 128 + (c & 1) * 128 + (4 * h + ks) * 8, make_smem_desc(pw + ks * 4096, 2048, TILE_SBO), idesc128, 1u);
              if (c == FF_CHUNKS - 1 && h == 1) umma_tf32(0, ones_desc, make_smem_desc(pw + SLAB_OFF, 0, TILE_SBO), idesc128, 1u);
              umma_commit(&bars[BAR_WEMPTY + g % NSLOT]);
              if (h == 1 && c == FF_CHUNKS - 1) umma_commit(&bars[BAR_X]);
            }
            __syncwarp();
          }
        }
      }
    }
    tc_fence_before();
  } else {
    // =========================== weight producer ===========================
    int G = 0;
#pragma unroll 1
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long b = tile * 128 / P.N;
      for (int lp = 0; lp < P.depth * PKT_PER_LAYER; ++lp, ++G) {
        const int slot = G % NSLOT;
        mbar_wait(&bars[BAR_WEMPTY + slot], ((uint32_t)(G / NSLOT) & 1u) ^ 1u);
        if (elect_one()) {
          const int l = lp / PKT_PER_LAYER, p = lp - l * PKT_PER_LAYER;
          const uint32_t bytes = (uint32_t)pkt_bytes(p);
          const uint8_t* src = p < 2 ? P.fold + (((size_t)b * P.depth + l) * 2 + p) * SLOT
                                     : P.stream + ((size_t)l * STATIC_PER_LAYER + (p - 2)) * SLOT;
          mbar_arrive_expect_tx(&bars[BAR_WFULL + slot], bytes);
          bulk_g2s(smem + SM_RING + slot * SLOT, src, bytes, &bars[BAR_WFULL + slot]);
        }
        __syncwarp();
      }
    }
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

const float* tc_foldw_base(const PackLayout& L, const void* packed, size_t* stride_floats);  // denoiser_tc.cu
const float* tc_b1p_base(const PackLayout& L, const void* packed);

int denoiser_forward_tf32(const PackLayout& L, const void* packed, int B, int N, const float* x, const float* anchors,
                          const float* variances, const int* assign, const float* valid_id, float* eps_out, Workspace& ws,
                          cudaStream_t st) {
  using namespace t32;
  DFB_REQUIRE(N % 128 == 0, DFB200_ERR_UNSUPPORTED, "denoiser (tf32 mode): N must be a multiple of 128 (got %d); use fp32 mode", N);
  static DeviceOnce attr_once;
  if (attr_once.first_time())
    DFB_CUDA(cudaFuncSetAttribute(denoiser_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  size_t fw_stride = 0;
  const float* foldw = tc_foldw_base(L, packed, &fw_stride);
  const float* Pf = reinterpret_cast<const float*>(packed);
  const size_t bo_stride = L.d.depth > 1 ? L.blk[1][B_BO] - L.blk[0][B_BO] : 0;  // the blocks' parameters are laid out uniformly
  context_fold32_kernel<<<dim3(B, L.d.depth), 256, 0, st>>>(L.d.depth, ws.kv, foldw, fw_stride, Pf + L.blk[0][B_BO], bo_stride,
                                                            reinterpret_cast<uint8_t*>(ws.fold));
  DFB_LAUNCH_CHECK();
  Tf32Params p{};
  p.stream = reinterpret_cast<const uint8_t*>(packed) + L.tf32_stream_off;
  p.fold = reinterpret_cast<const uint8_t*>(ws.fold);
  p.extras = reinterpret_cast<const float*>(p.stream + (size_t)L.d.depth * STATIC_PER_LAYER * SLOT);
  p.b1p = tc_b1p_base(L, packed);
  p.x = x; p.anchors = anchors; p.variances = variances; p.assign = assign; p.valid = valid_id; p.eps_out = eps_out;
  p.N = N; p.depth = L.d.depth; p.flags = L.d.flags;
  p.M = (long long)B * N;
  const int n_sm = current_device_sm_count();
  DFB_REQUIRE(n_sm > 0, DFB200_ERR_CUDA, "denoiser (tf32 mode): cannot query the SM count of the current device");
  const long long tiles = p.M / 128;
  denoiser_tf32_kernel<<<(int)(tiles < n_sm ? tiles : n_sm), THREADS, SMEM_BYTES, st>>>(p);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

}  // namespace dfb200
