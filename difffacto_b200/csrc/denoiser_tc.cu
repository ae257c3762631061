// DFB200_MODE_BF16: the whole cross-diffusion denoiser as ONE fused tcgen05 kernel.
//
// Reference computation: python/difffacto/models/diffusions/nets/attention.py:385-440 (+ blocks
// :161-306).  The reference materialises every activation in HBM (512 B/token per tensor, 4 KB/token
// for the GEGLU intermediate, ~180 kernels per step); here a CTA owns two 128-token tiles for the whole
// network and HBM sees only the 13 input features and the 3 output channels per token:
//   * residual stream x (fp32) lives in TENSOR MEMORY: 128 lanes x 128 columns per tile; the out-proj and
//     FF-out GEMMs accumulate straight into it (tcgen05.mma D += A.B), so the residual add is free;
//   * LayerNorm / attention-over-4-part-tokens / GEGLU run on CUDA cores out of TMEM (one thread = one
//     token row) and write the next bf16 A operand into shared memory in the canonical UMMA K-major
//     layout; LN gains/biases and all Linear biases are folded into the packed weights (bias = one
//     extra K=16 MMA against a constant "ones" tile, bf16 hi+lo split so it is fp32-accurate);
//   * weights stream L2 -> smem as pre-packed bf16 UMMA tiles through a 6-slot cp.async.bulk ring fed by
//     a producer warp; both tiles consume each packet, halving L2 traffic per token;
//   * one warp issues all MMAs; the two tiles ping-pong so the epilogue of one overlaps the MMAs of the
//     other; within a tile the FF hidden chunks are double-buffered in TMEM.
// TMEM map (512 columns): X0 [0,128) X1 [128,256) ACC0 [256,384) ACC1 [384,512).
#include <float.h>

#include "denoiser.cuh"
#include "tc_common.cuh"

namespace dfb200 {
using namespace tc;

// ---------------------------------------------------------------------------------------------
// packed bf16 stream: per layer 36 packets, each at a fixed 18 KB stride, in MMA consumption order
//   0: Wq' k[0,64)  + bias slab bq' (2 KB @16384)     1: Wq' k[64,128)
//   2: Wo  k[0,64)  + bias slab bo                    3: Wo  k[64,128)
//   4: W1' chunk 0 (+1 KB bias slab)   5: W1' chunk 1
//   6+2c: W2 chunk c, 7+2c: W1' chunk c+2   (c = 0..13);   34: W2 chunk 14;   35: W2 chunk 15 + b2 slab (2 KB @8192)
// Wq' = Wq.diag(norm2.w), bq' = Wq.norm2.b;  W1' = W1.diag(norm3.w), b1' = b1 + W1.norm3.b
// W1' chunk c: rows [0,32) = value units 32c.., rows [32,64) = gate units 512+32c..  (64 x 128)
// W2 chunk c: 128 rows x k[32c, 32c+32)
// ---------------------------------------------------------------------------------------------
constexpr int SLOT_BYTES = 18432;
constexpr int NSLOT = 6;
constexpr int PKT_PER_LAYER = 36;
constexpr int FF_CHUNKS = 16;
constexpr int BIAS_OFF_W128 = 16384;  // q0 / o0 / W1 packets: bias slab after 16 KB of weights
constexpr int BIAS_OFF_W2 = 8192;

__host__ __device__ inline int pkt_bytes(int p) {
  if (p == 0 || p == 2) return 16384 + 2048;
  if (p == 1 || p == 3) return 16384;
  if (p == 35) return 8192 + 2048;
  if (p == 34) return 8192;
  if (p == 4 || p == 5) return 16384 + 1024;
  return ((p - 6) & 1) ? 16384 + 1024 : 8192;
}
__host__ __device__ inline int pkt_w2(int c) { return c <= 13 ? 6 + 2 * c : 20 + c; }
__host__ __device__ inline int pkt_w1(int c) { return c < 2 ? 4 + c : 7 + 2 * (c - 2); }

size_t tc_stream_bytes_for(const NetDims& d) {
  // stream + folded head (3x128 weights + 4 biases, fp32)
  return (size_t)d.depth * PKT_PER_LAYER * SLOT_BYTES + sizeof(float) * (3 * D_MODEL + 4);
}

// ---- pack kernels (run once per weight update) -----------------------------------------------------
// dst: UMMA tile of R rows x KC k-values; element (r,k) = src[rowmap(r)*ld + k0 + k] * (gamma ? gamma[k0+k] : 1)
__global__ void pack_tile_kernel(uint8_t* __restrict__ dst, int R, int KC, const float* __restrict__ src, int ld, int k0,
                                 const float* __restrict__ gamma, int geglu_chunk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * KC) return;
  const int r = i / KC, k = i - r * KC;
  int row = r;
  if (geglu_chunk >= 0) row = r < 32 ? 32 * geglu_chunk + r : D_FF + 32 * geglu_chunk + (r - 32);
  float v = __ldg(src + (size_t)row * ld + k0 + k);
  if (gamma != nullptr) v *= __ldg(gamma + k0 + k);
  if (geglu_chunk >= 0 && r < 32) v *= 0.5f;  // value rows carry gelu's 0.5:  a*gelu(g) = (a/2)*g*(1+tanh(..))
  *reinterpret_cast<__nv_bfloat16*>(dst + tile_off(R, r, k)) = __float2bfloat16_rn(v);
}
// bias slab: R rows x 8 k-values (16 B per row); k=0: bf16 hi, k=1: bf16 lo of  bias[row] + W[row,:].beta
__global__ void pack_bias_kernel(uint8_t* __restrict__ dst, int R, const float* __restrict__ bias,
                                 const float* __restrict__ W, int ld, const float* __restrict__ beta, int geglu_chunk) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  int row = r;
  if (geglu_chunk >= 0) row = r < 32 ? 32 * geglu_chunk + r : D_FF + 32 * geglu_chunk + (r - 32);
  float v = bias != nullptr ? __ldg(bias + row) : 0.f;
  if (beta != nullptr)
    for (int k = 0; k < ld; ++k) v = fmaf(__ldg(W + (size_t)row * ld + k), __ldg(beta + k), v);
  if (geglu_chunk >= 0 && r < 32) v *= 0.5f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dst + r * 16);
  o[0] = hi;
  o[1] = lo;
#pragma unroll
  for (int k = 2; k < 8; ++k) o[k] = __float2bfloat16_rn(0.f);
}
// folded head: w_out' = w_out.diag(post_norm.w), b_out' = b_out + w_out.post_norm.b   (fp32)
__global__ void pack_head_kernel(float* __restrict__ dst, const float* __restrict__ w_out, const float* __restrict__ b_out,
                                 const float* __restrict__ g, const float* __restrict__ be) {
  const int c = blockIdx.x;
  float acc = 0.f;
  for (int k = threadIdx.x; k < D_MODEL; k += 32) {
    const float w = __ldg(w_out + c * D_MODEL + k);
    dst[c * D_MODEL + k] = w * __ldg(g + k);
    acc = fmaf(w, __ldg(be + k), acc);
  }
  for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
  if (threadIdx.x == 0) dst[3 * D_MODEL + c] = acc + __ldg(b_out + c);
}

int tc_pack_stream(const PackLayout& L, void* packed, cudaStream_t st) {
  const float* P = reinterpret_cast<const float*>(packed);
  uint8_t* S = reinterpret_cast<uint8_t*>(packed) + L.tc_stream_off;
  auto tile = [&](uint8_t* dst, int R, int KC, const float* src, int ld, int k0, const float* gamma, int chunk) {
    pack_tile_kernel<<<cdiv(R * KC, 256), 256, 0, st>>>(dst, R, KC, src, ld, k0, gamma, chunk);
    count_launch();
  };
  auto bias = [&](uint8_t* dst, int R, const float* b, const float* W, int ld, const float* beta, int chunk) {
    pack_bias_kernel<<<cdiv(R, 128), 128, 0, st>>>(dst, R, b, W, ld, beta, chunk);
    count_launch();
  };
  for (int l = 0; l < L.d.depth; ++l) {
    const size_t* o = L.blk[l];
    uint8_t* base = S + (size_t)l * PKT_PER_LAYER * SLOT_BYTES;
    auto pk = [&](int p) { return base + (size_t)p * SLOT_BYTES; };
    tile(pk(0), 128, 64, P + o[B_WQ], D_MODEL, 0, P + o[B_N2_W], -1);
    bias(pk(0) + BIAS_OFF_W128, 128, nullptr, P + o[B_WQ], D_MODEL, P + o[B_N2_B], -1);
    tile(pk(1), 128, 64, P + o[B_WQ], D_MODEL, 64, P + o[B_N2_W], -1);
    tile(pk(2), 128, 64, P + o[B_WO], D_MODEL, 0, nullptr, -1);
    bias(pk(2) + BIAS_OFF_W128, 128, P + o[B_BO], nullptr, 0, nullptr, -1);
    tile(pk(3), 128, 64, P + o[B_WO], D_MODEL, 64, nullptr, -1);
    for (int c = 0; c < FF_CHUNKS; ++c) {
      tile(pk(pkt_w1(c)), 64, 128, P + o[B_W1], D_MODEL, 0, P + o[B_N3_W], c);
      bias(pk(pkt_w1(c)) + BIAS_OFF_W128, 64, P + o[B_B1], P + o[B_W1], D_MODEL, P + o[B_N3_B], c);
      tile(pk(pkt_w2(c)), 128, 32, P + o[B_W2], D_FF, 32 * c, nullptr, -1);
    }
    bias(pk(35) + BIAS_OFF_W2, 128, P + o[B_B2], nullptr, 0, nullptr, -1);
  }
  float* head = reinterpret_cast<float*>(S + (size_t)L.d.depth * PKT_PER_LAYER * SLOT_BYTES);
  pack_head_kernel<<<3, 32, 0, st>>>(head, P + L.g[P_OUT_W], P + L.g[P_OUT_B], P + L.g[P_POST_W], P + L.g[P_POST_B]);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

// ---------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------
constexpr int TC_THREADS = 320;  // warps 0-3: tile 0 epilogue, 4-7: tile 1 epilogue, 8: MMA issuer, 9: weight producer
constexpr uint32_t SM_A = 0;               // 2 x 32768  A operand tiles (128 x 128 bf16)
constexpr uint32_t SM_U = 65536;           // [2][2] x 8192  gated FF activations (128 x 32 bf16)
constexpr uint32_t SM_ONES = 98304;        // 4096  ones tile (128 x 16 bf16: k=0,1 -> 1)
constexpr uint32_t SM_RING = 102400;       // 6 x 18432
constexpr uint32_t SM_KV = 212992;         // [2][2][4][128] fp32
constexpr uint32_t SM_BAR = 221184;        // mbarriers
constexpr uint32_t SM_TMEM = SM_BAR + 256;
constexpr uint32_t TC_SMEM_BYTES = SM_TMEM + 64;

enum Bar { BAR_A = 0 /*[2]*/, BAR_ACC = 2 /*[2][2]*/, BAR_UREADY = 6 /*[2][2]*/, BAR_UFREE = 10 /*[2][2]*/, BAR_X = 14 /*[2]*/,
           BAR_WFULL = 16 /*[6]*/, BAR_WEMPTY = 22 /*[6]*/, BAR_COUNT = 28 };

struct TcParams {
  const uint8_t* stream;
  const float* head;  // folded proj_out: [3][128] weights, then 3 biases
  const float* w_in; const float* b_in; const float* pre_w; const float* pre_b;
  const float* kv;    // [B][depth][2][4][128]
  const float* x; const float* anchors; const float* variances; const int* assign; const float* valid;
  float* eps_out;
  int N, depth, flags;
  long long M;
  long long* dbg;  // optional timeline buffer (DFB200_TC_TIMELINE env): CTA 0 records clock64() at phase boundaries
};

// Packed fp32x2 math (FFMA2 on sm_100): the CUDA-core epilogues are the bottleneck of this kernel (the
// tensor pipe waits on them), so every elementwise chain below processes two columns per instruction.
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// GEGLU for two columns: (a/2 arrives from the MMA) * g * (1 + tanh(g*(c0 + c1 g^2))).
// tanh-form GELU with (c0, c1) refit against the exact erf GELU (max abs error 2.7e-4 over R, below the bf16
// rounding of the result); 5 packed FMA-pipe instructions + 2 MUFU.TANH per column pair (erff costs ~45/column).
__device__ __forceinline__ float2 geglu2(float2 a_half, float2 g) {
  const float2 g2 = __fmul2_rn(g, g);
  const float2 in = __fmul2_rn(g, __ffma2_rn(g2, f2s(0.034700932528f), f2s(0.800156991001f)));
  const float2 t = f2(tanh_approx(in.x), tanh_approx(in.y));
  const float2 ag = __fmul2_rn(a_half, g);
  return __ffma2_rn(ag, t, ag);
}

// timeline instrumentation (off unless a buffer is supplied): slot layout [who][event], who 0 = tile-0 row 0, 1 = MMA lane
#define TL(who, ev)                                                                                  \
  do {                                                                                               \
    if (P.dbg != nullptr && blockIdx.x == 0 && tl_on) P.dbg[(who) * 512 + (ev)] = clock64();         \
  } while (0)

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// one thread = one token row: mean / rstd of the 128-wide fp32 row held in TMEM columns [col, col+128)
__device__ __forceinline__ void row_stats(uint32_t taddr, float& mean, float& rstd) {
  float2 s2 = f2s(0.f), q2 = f2s(0.f);
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {
    float h[32];
    tmem_ld32(taddr + cb * 32, h);
    tmem_wait_ld();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float2 v = f2(h[2 * k], h[2 * k + 1]);
      s2 = __fadd2_rn(s2, v);
      q2 = __ffma2_rn(v, v, q2);
    }
  }
  const float s = s2.x + s2.y, q = q2.x + q2.y;
  mean = s * (1.f / D_MODEL);
  const float var = fmaxf(q * (1.f / D_MODEL) - mean * mean, 0.f);
  rstd = rsqrtf(var + LN_EPS);
}

// normalise the TMEM row (gain/bias are folded into the next weights) and write it as the bf16 A operand row
__device__ __forceinline__ void row_normalize_to_tile(uint32_t taddr, float mean, float rstd, uint8_t* tile, int r) {
  const float2 nm = f2s(-mean * rstd), rs = f2s(rstd);
#pragma unroll
  for (int cb = 0; cb < 4; ++cb) {
    float h[32];
    tmem_ld32(taddr + cb * 32, h);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 y = __ffma2_rn(f2(h[8 * j + 2 * i], h[8 * j + 2 * i + 1]), rs, nm);
        w[i] = pack_bf16(y.x, y.y);
      }
      *reinterpret_cast<uint4*>(tile + (cb * 4 + j) * 2048 + r * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// D[128 x NB] (+)= A[128 x 16*KSTEPS] . B[NB x 16*KSTEPS]^T.  A tiles have 128 rows (k-slab = 2048 B), B tiles NB rows
// (k-slab = NB*16 B).  Fully unrolled so that descriptors are (uniform base + immediate).
template <int NB, int KSTEPS>
__device__ __forceinline__ void umma_gemm(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, uint32_t acc_first) {
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks)
    umma_bf16(d_tmem, make_smem_desc(a_addr + ks * 4096, 2048, TILE_SBO), make_smem_desc(b_addr + ks * (NB * 32), NB * 16, TILE_SBO),
              idesc, ks > 0 ? 1u : acc_first);
}

// Same for both tiles of the CTA against ONE weight tile, K-steps interleaved (T0,k),(T1,k).
template <int NB, int KSTEPS>
__device__ __forceinline__ void umma_gemm2(uint32_t d0, uint32_t d1, uint32_t a0, uint32_t a1, uint32_t b_addr, uint32_t idesc,
                                           uint32_t acc_first) {
#pragma unroll
  for (int ks = 0; ks < KSTEPS; ++ks) {
    const uint64_t bd = make_smem_desc(b_addr + ks * (NB * 32), NB * 16, TILE_SBO);
    umma_bf16(d0, make_smem_desc(a0 + ks * 4096, 2048, TILE_SBO), bd, idesc, ks > 0 ? 1u : acc_first);
    umma_bf16(d1, make_smem_desc(a1 + ks * 4096, 2048, TILE_SBO), bd, idesc, ks > 0 ? 1u : acc_first);
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1) denoiser_tc_kernel(const TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_TMEM);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);  // warp-uniform by construction (helps uniform-datapath codegen)

  // ---- one-time setup ----
  if (warp == 8 && lane == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(&bars[BAR_A + i], 128); mbar_init(&bars[BAR_X + i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&bars[BAR_ACC + i], 1); mbar_init(&bars[BAR_UREADY + i], 128); mbar_init(&bars[BAR_UFREE + i], 1); }
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&bars[BAR_WFULL + i], 1); mbar_init(&bars[BAR_WEMPTY + i], 1); }
    fence_barrier_init();
  }
  if (tid < 256) {  // ones tile: slab 0 (k 0..7) = {1,1,0,...}, slab 1 (k 8..15) = 0
    uint4 v = make_uint4(tid < 128 ? 0x3F803F80u : 0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(smem + SM_ONES + tid * 16) = v;
  }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tmem != 0u) __trap();  // a 512-column allocation owns the whole TMEM: base = lane 0, column 0 (the MMA path relies on it)

  if (warp < 8) {
    // =========================== epilogue warps: one thread per token row ===========================
    const int T = warp >> 2, r = tid & 127;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t X = lane_base + T * 128, ACC = lane_base + 256 + T * 128;
    uint8_t* a_tile = smem + SM_A + T * 32768;
    float* kvs = reinterpret_cast<float*>(smem + SM_KV) + T * 1024;
    const long long tile_tok0 = ((long long)blockIdx.x * 2 + T) * 128;
    const bool tile_ok = tile_tok0 < P.M;
    const bool tl_on = tid == 0;
    TL(0, 0);
    const long long tok = tile_ok ? tile_tok0 + r : (P.M - 128 + r);  // an out-of-range tile recomputes the last one
    const long long b = tok / P.N;
    const int p = (int)(tok - b * P.N);
    uint32_t ph_acc[2] = {0, 0}, ph_x = 0;
    float vm[MAX_TOKENS];
#pragma unroll
    for (int j = 0; j < MAX_TOKENS; ++j) vm[j] = P.valid != nullptr ? __ldg(P.valid + b * MAX_TOKENS + j) : 1.f;

    // ---- proj_in (13 -> 128) + pre_norm, result (the residual stream) into TMEM ----
    // proj_in weights are staged transposed ([feature][output], fp32) in the (still unused) U-tile region so that one
    // broadcast LDS.128 feeds two FFMA2 (4 outputs) -- the straightforward per-output uniform LDG version of this
    // prologue cost 11% of the kernel.
    {
      float* wt = reinterpret_cast<float*>(smem + SM_U);  // [13][128] weights, [128] bias, [128] pre_norm.w, [128] pre_norm.b
      const int et = tid;                                  // 256 epilogue threads
      for (int i = et; i < D_MODEL * 13; i += 256) {
        const int k = i / 13, c = i - k * 13;              // coalesced read of w_in[k][c]
        wt[c * D_MODEL + k] = __ldg(P.w_in + i);
      }
      if (et < D_MODEL) {
        wt[13 * D_MODEL + et] = __ldg(P.b_in + et);
        wt[14 * D_MODEL + et] = __ldg(P.pre_w + et);
        wt[15 * D_MODEL + et] = __ldg(P.pre_b + et);
      }
      float f[13];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        f[c] = __ldg(P.x + (b * 3 + c) * P.N + p);
        f[3 + c] = __ldg(P.anchors + (b * 3 + c) * P.N + p);
        const float v = __ldg(P.variances + (b * 3 + c) * P.N + p);
        f[6 + c] = (P.flags & DFB200_NET_INCLUDE_STD) ? sqrtf(v) : v;
      }
      const int part = __ldg(P.assign + tok);
#pragma unroll
      for (int c = 0; c < 4; ++c) f[9 + c] = part == c ? 1.f : 0.f;
      named_bar_sync(3, 256);
      float2 s2 = f2s(0.f), q2 = f2s(0.f);
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {
        float h[32];
#pragma unroll
        for (int kq = 0; kq < 8; ++kq) {
          const float4 bi = *reinterpret_cast<const float4*>(wt + 13 * D_MODEL + cb * 32 + kq * 4);
          float2 a0 = f2(bi.x, bi.y), a1 = f2(bi.z, bi.w);
#pragma unroll
          for (int c = 0; c < 13; ++c) {
            const float4 w = *reinterpret_cast<const float4*>(wt + c * D_MODEL + cb * 32 + kq * 4);
            a0 = __ffma2_rn(f2s(f[c]), f2(w.x, w.y), a0);
            a1 = __ffma2_rn(f2s(f[c]), f2(w.z, w.w), a1);
          }
          h[kq * 4] = a0.x; h[kq * 4 + 1] = a0.y; h[kq * 4 + 2] = a1.x; h[kq * 4 + 3] = a1.y;
          s2 = __fadd2_rn(s2, __fadd2_rn(a0, a1));
          q2 = __ffma2_rn(a0, a0, q2);
          q2 = __ffma2_rn(a1, a1, q2);
        }
        tmem_st32(X + cb * 32, h);
      }
      tmem_wait_st();
      const float mean = (s2.x + s2.y) * (1.f / D_MODEL);
      const float rstd = rsqrtf(fmaxf((q2.x + q2.y) * (1.f / D_MODEL) - mean * mean, 0.f) + LN_EPS);
      const float2 rs = f2s(rstd), nm = f2s(-mean * rstd);
#pragma unroll 1
      for (int cb = 0; cb < 4; ++cb) {
        float h[32];
        tmem_ld32(X + cb * 32, h);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float2 g = *reinterpret_cast<const float2*>(wt + 14 * D_MODEL + cb * 32 + 2 * k);
          const float2 be = *reinterpret_cast<const float2*>(wt + 15 * D_MODEL + cb * 32 + 2 * k);
          const float2 y = __ffma2_rn(__ffma2_rn(f2(h[2 * k], h[2 * k + 1]), rs, nm), g, be);
          h[2 * k] = y.x; h[2 * k + 1] = y.y;
        }
        tmem_st32(X + cb * 32, h);
      }
      tmem_wait_st();
      named_bar_sync(3, 256);  // everyone is done with the staged weights before the U region is reused
    }

    TL(0, 1);
    for (int l = 0; l < P.depth; ++l) {
      TL(0, 2 + l * 40);
      // K/V of this sample and block -> smem (every row of the tile belongs to the same sample)
      {
        const float4* src = reinterpret_cast<const float4*>(P.kv + ((size_t)b * P.depth + l) * 1024);
        float4* dst = reinterpret_cast<float4*>(kvs);
        dst[r * 2] = __ldg(src + r * 2);
        dst[r * 2 + 1] = __ldg(src + r * 2 + 1);
      }
      // ---- LN2 -> A ----
      float mean, rstd;
      row_stats(X, mean, rstd);
      row_normalize_to_tile(X, mean, rstd, a_tile, r);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars[BAR_A + T]);
      named_bar_sync(1 + T, 128);  // kvs visible to the tile's 128 threads
      TL(0, 3 + l * 40);

      // ---- attention over the 4 part tokens, head by head, out of the Q accumulator ----
      mbar_wait(&bars[BAR_ACC + T * 2 + 0], ph_acc[0]);
      ph_acc[0] ^= 1;
      tc_fence_after();
      TL(0, 4 + l * 40);
#pragma unroll 1
      for (int h = 0; h < 8; ++h) {
        float qv[16];
        tmem_ld16(ACC + h * 16, qv);
        tmem_wait_ld();
        float sim[MAX_TOKENS];
#pragma unroll
        for (int j = 0; j < MAX_TOKENS; ++j) {
          const float4* kp = reinterpret_cast<const float4*>(kvs + j * D_MODEL + h * 16);
          float2 s2 = f2s(0.f);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 k4 = kp[i];
            s2 = __ffma2_rn(f2(qv[4 * i], qv[4 * i + 1]), f2(k4.x, k4.y), s2);
            s2 = __ffma2_rn(f2(qv[4 * i + 2], qv[4 * i + 3]), f2(k4.z, k4.w), s2);
          }
          sim[j] = vm[j] == 0.f ? -FLT_MAX : (s2.x + s2.y) * 0.25f;  // masked_fill(~mask, -finfo.max)
        }
        const float mx = fmaxf(fmaxf(sim[0], sim[1]), fmaxf(sim[2], sim[3]));
        float pj[MAX_TOKENS], den = 0.f;
#pragma unroll
        for (int j = 0; j < MAX_TOKENS; ++j) { pj[j] = __expf(sim[j] - mx); den += pj[j]; }
        const float inv = __fdividef(1.f, den);
        float2 o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = f2s(0.f);
#pragma unroll
        for (int j = 0; j < MAX_TOKENS; ++j) {
          const float2 w = f2s(pj[j] * inv);
          const float4* vp = reinterpret_cast<const float4*>(kvs + 512 + j * D_MODEL + h * 16);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 v4 = vp[i];
            o[2 * i] = __ffma2_rn(w, f2(v4.x, v4.y), o[2 * i]);
            o[2 * i + 1] = __ffma2_rn(w, f2(v4.z, v4.w), o[2 * i + 1]);
          }
        }
        uint4 v0, v1;
        v0.x = pack_bf16(o[0].x, o[0].y); v0.y = pack_bf16(o[1].x, o[1].y); v0.z = pack_bf16(o[2].x, o[2].y); v0.w = pack_bf16(o[3].x, o[3].y);
        v1.x = pack_bf16(o[4].x, o[4].y); v1.y = pack_bf16(o[5].x, o[5].y); v1.z = pack_bf16(o[6].x, o[6].y); v1.w = pack_bf16(o[7].x, o[7].y);
        *reinterpret_cast<uint4*>(a_tile + (2 * h) * 2048 + r * 16) = v0;
        *reinterpret_cast<uint4*>(a_tile + (2 * h + 1) * 2048 + r * 16) = v1;
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars[BAR_A + T]);

      TL(0, 5 + l * 40);
      // ---- x += attn @ Wo^T + bo (accumulated in TMEM by the MMA warp);  LN3 -> A ----
      mbar_wait(&bars[BAR_X + T], ph_x);
      ph_x ^= 1;
      tc_fence_after();
      TL(0, 6 + l * 40);
      row_stats(X, mean, rstd);
      row_normalize_to_tile(X, mean, rstd, a_tile, r);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars[BAR_A + T]);

      TL(0, 7 + l * 40);
      // ---- GEGLU feed-forward, 16 chunks of 32 value + 32 gate columns ----
#pragma unroll 1
      for (int c = 0; c < FF_CHUNKS; ++c) {
        const int hb = c & 1;
        const int g = l * FF_CHUNKS + c;  // global chunk counter: U[T][hb] is reused every 2 chunks
        mbar_wait(&bars[BAR_ACC + T * 2 + hb], ph_acc[hb]);
        ph_acc[hb] ^= 1;
        tc_fence_after();
        TL(0, 8 + l * 40 + c * 2);
        float a[32], gt[32];
        tmem_ld32(ACC + hb * 64, a);
        tmem_ld32(ACC + hb * 64 + 32, gt);
        tmem_wait_ld();
        uint32_t u[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const float2 y = geglu2(f2(a[2 * k], a[2 * k + 1]), f2(gt[2 * k], gt[2 * k + 1]));
          u[k] = pack_bf16(y.x, y.y);
        }
        // the FF-out MMAs of chunk g-2 must have finished reading this U buffer
        mbar_wait(&bars[BAR_UFREE + T * 2 + hb], (((uint32_t)g >> 1) & 1u) ^ 1u);
        uint8_t* ut = smem + SM_U + (T * 2 + hb) * 8192;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(ut + j * 2048 + r * 16) = make_uint4(u[4 * j], u[4 * j + 1], u[4 * j + 2], u[4 * j + 3]);
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(&bars[BAR_UREADY + T * 2 + hb]);
        TL(0, 9 + l * 40 + c * 2);
      }
      mbar_wait(&bars[BAR_X + T], ph_x);
      ph_x ^= 1;
      tc_fence_after();
    }

    TL(0, 2 + P.depth * 40);
    // ---- post_norm (folded) + proj_out (128 -> 3) ----
    {
      float mean, rstd;
      row_stats(X, mean, rstd);
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
      for (int cb = 0; cb < 4; ++cb) {
        float h[32];
        tmem_ld32(X + cb * 32, h);
        tmem_wait_ld();
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(P.head + cb * 32) + k4);
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(P.head + D_MODEL + cb * 32) + k4);
          const float4 w2 = __ldg(reinterpret_cast<const float4*>(P.head + 2 * D_MODEL + cb * 32) + k4);
          const float y0 = (h[4 * k4] - mean) * rstd, y1 = (h[4 * k4 + 1] - mean) * rstd;
          const float y2 = (h[4 * k4 + 2] - mean) * rstd, y3 = (h[4 * k4 + 3] - mean) * rstd;
          o0 = fmaf(y3, w0.w, fmaf(y2, w0.z, fmaf(y1, w0.y, fmaf(y0, w0.x, o0))));
          o1 = fmaf(y3, w1.w, fmaf(y2, w1.z, fmaf(y1, w1.y, fmaf(y0, w1.x, o1))));
          o2 = fmaf(y3, w2.w, fmaf(y2, w2.z, fmaf(y1, w2.y, fmaf(y0, w2.x, o2))));
        }
      }
      if (tile_ok) {
        P.eps_out[(b * 3 + 0) * P.N + p] = o0 + __ldg(P.head + 3 * D_MODEL + 0);
        P.eps_out[(b * 3 + 1) * P.N + p] = o1 + __ldg(P.head + 3 * D_MODEL + 1);
        P.eps_out[(b * 3 + 2) * P.N + p] = o2 + __ldg(P.head + 3 * D_MODEL + 2);
      }
    }
    TL(0, 3 + P.depth * 40);
    tc_fence_before();
  } else if (warp == 8) {
    // =========================== MMA issuer ===========================
    // The whole warp runs this control flow (waits included); one elected lane issues tcgen05.mma / commit.
    // Everything that feeds a descriptor is warp-uniform by construction (constants, loop counters, the TMEM
    // base which is 0 for a 512-column allocation), so the issue sequence stays on the uniform datapath.
    constexpr uint32_t idesc128 = make_idesc_bf16(128, 128), idesc64 = make_idesc_bf16(128, 64);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t ring = sbase + SM_RING, a_base = sbase + SM_A, u_base = sbase + SM_U;
    const uint64_t ones_desc = make_smem_desc(sbase + SM_ONES, 2048, TILE_SBO);
    uint32_t ph_a0 = 0, ph_a1 = 0, ph_u = 0;  // ph_u: bit hb (both tiles advance together)
    const bool tl_on = lane == 0;
    auto pkt_addr = [&](int G) -> uint32_t {  // wait until packet G has landed; its smem address
      mbar_wait(&bars[BAR_WFULL + G % NSLOT], (uint32_t)(G / NSLOT) & 1u);
      return ring + (uint32_t)(G % NSLOT) * SLOT_BYTES;
    };
    auto wait_a = [&](int T) {
      if (T == 0) { mbar_wait(&bars[BAR_A + 0], ph_a0); ph_a0 ^= 1; }
      else { mbar_wait(&bars[BAR_A + 1], ph_a1); ph_a1 ^= 1; }
    };
    // Both tiles advance in LOCKSTEP through the MMA schedule and their K-steps are interleaved (T0,k),(T1,k): one set of
    // barrier/packet waits serves both tiles, and the epilogues of the two tiles (8 warps) run concurrently.  MMA/epilogue
    // overlap comes from the double-buffered FF hidden chunks (ACC halves), not from skewing the tiles.
    for (int l = 0; l < P.depth; ++l) {
      const int G0 = l * PKT_PER_LAYER;
      TL(1, 2 + l * 40);
      // ---- Q = LN2(x) Wq'^T + bq' ----
      {
        wait_a(0); wait_a(1);
        const uint32_t p0 = pkt_addr(G0 + 0), p1 = pkt_addr(G0 + 1);
        tc_fence_after();
        if (elect_one()) {
          umma_gemm2<128, 4>(256, 384, a_base, a_base + 32768, p0, idesc128, 0u);
          umma_gemm2<128, 4>(256, 384, a_base + 4 * 4096, a_base + 32768 + 4 * 4096, p1, idesc128, 1u);
          const uint64_t bdsc = make_smem_desc(p0 + BIAS_OFF_W128, 0, TILE_SBO);
          umma_bf16(256, ones_desc, bdsc, idesc128, 1u);
          umma_bf16(384, ones_desc, bdsc, idesc128, 1u);
          umma_commit(&bars[BAR_ACC + 0]);
          umma_commit(&bars[BAR_ACC + 2]);
          umma_commit(&bars[BAR_WEMPTY + (G0 + 0) % NSLOT]);
          umma_commit(&bars[BAR_WEMPTY + (G0 + 1) % NSLOT]);
        }
        __syncwarp();
      }
      TL(1, 3 + l * 40);
      // ---- x += O Wo^T + bo ----
      {
        wait_a(0);
        TL(1, 4 + l * 40);
        wait_a(1);
        const uint32_t p2 = pkt_addr(G0 + 2), p3 = pkt_addr(G0 + 3);
        tc_fence_after();
        if (elect_one()) {
          umma_gemm2<128, 4>(0, 128, a_base, a_base + 32768, p2, idesc128, 1u);
          umma_gemm2<128, 4>(0, 128, a_base + 4 * 4096, a_base + 32768 + 4 * 4096, p3, idesc128, 1u);
          const uint64_t bdsc = make_smem_desc(p2 + BIAS_OFF_W128, 0, TILE_SBO);
          umma_bf16(0, ones_desc, bdsc, idesc128, 1u);
          umma_bf16(128, ones_desc, bdsc, idesc128, 1u);
          umma_commit(&bars[BAR_X + 0]);
          umma_commit(&bars[BAR_X + 1]);
          umma_commit(&bars[BAR_WEMPTY + (G0 + 2) % NSLOT]);
          umma_commit(&bars[BAR_WEMPTY + (G0 + 3) % NSLOT]);
        }
        __syncwarp();
      }
      // ---- feed-forward ----
      // H_c = LN3(x) W1'_c^T + b1'_c for both tiles -> ACC_T half (c & 1); issued inside an elected region
      auto ff_in2 = [&](uint32_t pw, int c) {
        const uint32_t d0 = 256 + (c & 1) * 64, d1 = 384 + (c & 1) * 64;
        umma_gemm2<64, 8>(d0, d1, a_base, a_base + 32768, pw, idesc64, 0u);
        const uint64_t bdsc = make_smem_desc(pw + BIAS_OFF_W128, 0, TILE_SBO);
        umma_bf16(d0, ones_desc, bdsc, idesc64, 1u);
        umma_bf16(d1, ones_desc, bdsc, idesc64, 1u);
        umma_commit(&bars[BAR_ACC + (c & 1)]);
        umma_commit(&bars[BAR_ACC + 2 + (c & 1)]);
      };
      TL(1, 5 + l * 40);
      wait_a(0); wait_a(1);
      TL(1, 6 + l * 40);
      {
        const uint32_t pa = pkt_addr(G0 + 4), pb = pkt_addr(G0 + 5);
        tc_fence_after();
        if (elect_one()) {
          ff_in2(pa, 0);
          ff_in2(pb, 1);
          umma_commit(&bars[BAR_WEMPTY + (G0 + 4) % NSLOT]);
          umma_commit(&bars[BAR_WEMPTY + (G0 + 5) % NSLOT]);
        }
        __syncwarp();
      }
      TL(1, 7 + l * 40);
#pragma unroll 1
      for (int c = 0; c < FF_CHUNKS; ++c) {
        const int hb = c & 1;
        TL(1, 8 + l * 40 + c * 2);
        mbar_wait(&bars[BAR_UREADY + hb], (ph_u >> hb) & 1u);
        mbar_wait(&bars[BAR_UREADY + 2 + hb], (ph_u >> hb) & 1u);
        ph_u ^= 1u << hb;
        TL(1, 9 + l * 40 + c * 2);
        const uint32_t pw2 = pkt_addr(G0 + pkt_w2(c));
        uint32_t pw1 = 0;
        if (c + 2 < FF_CHUNKS) pw1 = pkt_addr(G0 + pkt_w1(c + 2));
        tc_fence_after();
        if (elect_one()) {
          umma_gemm2<128, 2>(0, 128, u_base + hb * 8192, u_base + (2 + hb) * 8192, pw2, idesc128, 1u);
          if (c == FF_CHUNKS - 1) {
            const uint64_t bdsc = make_smem_desc(pw2 + BIAS_OFF_W2, 0, TILE_SBO);
            umma_bf16(0, ones_desc, bdsc, idesc128, 1u);
            umma_bf16(128, ones_desc, bdsc, idesc128, 1u);
          }
          umma_commit(&bars[BAR_UFREE + hb]);
          umma_commit(&bars[BAR_UFREE + 2 + hb]);
          if (c == FF_CHUNKS - 1) { umma_commit(&bars[BAR_X + 0]); umma_commit(&bars[BAR_X + 1]); }
          umma_commit(&bars[BAR_WEMPTY + (G0 + pkt_w2(c)) % NSLOT]);
          if (c + 2 < FF_CHUNKS) {
            ff_in2(pw1, c + 2);
            umma_commit(&bars[BAR_WEMPTY + (G0 + pkt_w1(c + 2)) % NSLOT]);
          }
        }
        __syncwarp();
      }
    }
    tc_fence_before();
  } else {
    // =========================== weight producer ===========================
    const int total = P.depth * PKT_PER_LAYER;
    for (int G = 0; G < total; ++G) {
      const int slot = G % NSLOT;
      mbar_wait(&bars[BAR_WEMPTY + slot], ((uint32_t)(G / NSLOT) & 1u) ^ 1u);
      if (elect_one()) {
        const uint32_t bytes = (uint32_t)pkt_bytes(G % PKT_PER_LAYER);
        mbar_arrive_expect_tx(&bars[BAR_WFULL + slot], bytes);
        bulk_g2s(smem + SM_RING + slot * SLOT_BYTES, P.stream + (size_t)G * SLOT_BYTES, bytes, &bars[BAR_WFULL + slot]);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

static long long* g_tc_timeline = nullptr;  // set by dfb200_debug_tc_timeline

int denoiser_forward_tc(const PackLayout& L, const void* packed, int B, int N, const float* x, const float* anchors,
                        const float* variances, const int* assign, const float* valid_id, float* eps_out, Workspace& ws,
                        cudaStream_t st) {
  DFB_REQUIRE(N % 128 == 0, DFB200_ERR_UNSUPPORTED, "denoiser (bf16 mode): N must be a multiple of 128 (got %d); use fp32 mode", N);
  static bool attr_set = false;
  if (!attr_set) {
    DFB_CUDA(cudaFuncSetAttribute(denoiser_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    attr_set = true;
  }
  const float* Pf = reinterpret_cast<const float*>(packed);
  const uint8_t* S = reinterpret_cast<const uint8_t*>(packed) + L.tc_stream_off;
  TcParams p{};
  p.stream = S;
  p.head = reinterpret_cast<const float*>(S + (size_t)L.d.depth * PKT_PER_LAYER * SLOT_BYTES);
  p.w_in = Pf + L.g[P_IN_W]; p.b_in = Pf + L.g[P_IN_B]; p.pre_w = Pf + L.g[P_PRE_W]; p.pre_b = Pf + L.g[P_PRE_B];
  p.kv = ws.kv;
  p.x = x; p.anchors = anchors; p.variances = variances; p.assign = assign; p.valid = valid_id;
  p.eps_out = eps_out;
  p.N = N; p.depth = L.d.depth; p.flags = L.d.flags;
  p.M = (long long)B * N;
  p.dbg = g_tc_timeline;
  const int grid = cdiv(p.M, 256);
  denoiser_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(p);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

// ---------------------------------------------------------------------------------------------
// UMMA self-test: D[128 x N] = (Cin) + A[128 x K] . W[N x K]^T (+ bias), bf16 operands, one CTA.
// Exercises exactly the building blocks of the fused kernel: canonical no-swizzle K-major tiles written
// by threads, bulk-copied B tile, TMEM alloc / st / ld, accumulate onto pre-stored TMEM, bias-by-ones-MMA.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_selftest_kernel(int variant, int N, int K, const float* __restrict__ A, const float* __restrict__ W,
                     const float* __restrict__ bias, const float* __restrict__ Cin, float* __restrict__ D,
                     uint8_t* __restrict__ scratch) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* a_tile = smem;                 // 128 x 128 bf16 max = 32768
  uint8_t* b_tile = smem + 32768;         // 128 x 128 bf16 max = 32768
  uint8_t* ones = smem + 65536;           // 4096
  uint8_t* bslab = smem + 69632;          // 2 slabs x 2048
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 73728);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 73728 + 64);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool swap_lbo_sbo = variant & 1, bias_two_slabs = variant & 2, use_bulk = variant & 4;
  if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
  // A: row tid
  for (int k = 0; k < K; k += 2)
    *reinterpret_cast<uint32_t*>(a_tile + tile_off(128, tid, k)) = pack_bf16(A[tid * K + k], A[tid * K + k + 1]);
  uint8_t* bdst = use_bulk ? scratch : b_tile;
  for (int i = tid; i < N * K / 2; i += 128) {
    const int n = i / (K / 2), k = (i - n * (K / 2)) * 2;
    *reinterpret_cast<uint32_t*>(bdst + tile_off(N, n, k)) = pack_bf16(W[n * K + k], W[n * K + k + 1]);
  }
  {
    *reinterpret_cast<uint4*>(ones + tid * 16) = make_uint4(0x3F803F80u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(ones + 2048 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    if (tid < N) {
      const float v = bias != nullptr ? bias[tid] : 0.f;
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      uint32_t w0 = (uint32_t)(*reinterpret_cast<const uint16_t*>(&hi)) | ((uint32_t)(*reinterpret_cast<const uint16_t*>(&lo)) << 16);
      *reinterpret_cast<uint4*>(bslab + tid * 16) = make_uint4(w0, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(bslab + N * 16 + tid * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  fence_proxy_async();
  __threadfence();
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t row_addr = tmem + ((uint32_t)(warp * 32) << 16);
  if (use_bulk && tid == 0) {
    mbar_arrive_expect_tx(&bars[1], (uint32_t)(N * K * 2));
    bulk_g2s(b_tile, scratch, (uint32_t)(N * K * 2), &bars[1]);
  }
  if (Cin != nullptr) {
    for (int cb = 0; cb < N / 32; ++cb) {
      float h[32];
      for (int k = 0; k < 32; ++k) h[k] = Cin[tid * N + cb * 32 + k];
      tmem_st32(row_addr + cb * 32, h);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    if (use_bulk) mbar_wait(&bars[1], 0);
    tc_fence_after();
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(128, N);
      const uint32_t a_lbo = swap_lbo_sbo ? TILE_SBO : 2048u, a_sbo = swap_lbo_sbo ? 2048u : TILE_SBO;
      const uint32_t b_lbo = swap_lbo_sbo ? TILE_SBO : (uint32_t)(N * 16), b_sbo = swap_lbo_sbo ? (uint32_t)(N * 16) : TILE_SBO;
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t ad = make_smem_desc(smem_u32(a_tile) + ks * 4096, a_lbo, a_sbo);
        const uint64_t bd = make_smem_desc(smem_u32(b_tile) + ks * (N * 32), b_lbo, b_sbo);
        umma_bf16(tmem, ad, bd, idesc, (ks > 0 || Cin != nullptr) ? 1u : 0u);
      }
      if (bias != nullptr) {
        const uint64_t od = make_smem_desc(smem_u32(ones), a_lbo, a_sbo);
        const uint64_t bd = bias_two_slabs ? make_smem_desc(smem_u32(bslab), b_lbo, b_sbo)
                                           : (swap_lbo_sbo ? make_smem_desc(smem_u32(bslab), TILE_SBO, 0u)
                                                           : make_smem_desc(smem_u32(bslab), 0u, TILE_SBO));
        umma_bf16(tmem, od, bd, idesc, 1u);
      }
      umma_commit(&bars[0]);
    }
    __syncwarp();
  }
  mbar_wait(&bars[0], 0);
  tc_fence_after();
  for (int cb = 0; cb < N / 32; ++cb) {
    float h[32];
    tmem_ld32(row_addr + cb * 32, h);
    tmem_wait_ld();
    for (int k = 0; k < 32; ++k) D[tid * N + cb * 32 + k] = h[k];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// ---------------------------------------------------------------------------------------------
// UMMA issue-rate microbenchmark: `iters` back-to-back K=16 MMAs (M=128, N) from smem operands, cycles from the
// first issue to the commit's arrival.  layout 0 = no-swizzle canonical tiles (as used by the fused kernel),
// 1 = SWIZZLE_128B descriptors.  Operand contents are irrelevant (timing only).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int layout, int N, int iters, int ksteps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 131072);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 131072 + 64);
  const int tid = threadIdx.x;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  for (int i = tid; i < 131072 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid == 0) { mbar_init(&bars[0], 1); fence_barrier_init(); }
  fence_proxy_async();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    const uint32_t sbase = smem_u32(smem);
    const uint32_t idesc = make_idesc_bf16(128, N);
    const bool same_acc = (layout & 2) != 0;  // every MMA accumulates into the same TMEM tile (dependent chain)
    layout &= 1;
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      // descriptors precomputed; the issue loop is 8 unrolled MMAs per iteration (like the fused kernel's sequences)
      uint64_t ad[8], bd[8];
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int kk = ks % ksteps;
        if (layout == 0) {
          ad[ks] = make_smem_desc(sbase + kk * 4096, 2048, TILE_SBO);
          bd[ks] = make_smem_desc(sbase + 65536 + kk * (N * 32), N * 16, TILE_SBO);
        } else {  // SWIZZLE_128B K-major: rows of 128 B, 8-row atoms of 1024 B; a K=16 step advances the start by 32 B
          ad[ks] = make_smem_desc(sbase + (kk >> 2) * 16384 + (kk & 3) * 32, 16, 1024) | ((uint64_t)2 << 61);
          bd[ks] = make_smem_desc(sbase + 65536 + (kk >> 2) * (N * 128) + (kk & 3) * 32, 16, 1024) | ((uint64_t)2 << 61);
        }
      }
      t0 = clock64();
      for (int i = 0; i < iters; i += 8) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) umma_bf16(tmem + (same_acc ? 0 : (ks & 1) * 256), ad[ks], bd[ks], idesc, 1u);
      }
      umma_commit(&bars[0]);
    }
    __syncwarp();
    mbar_wait(&bars[0], 0);
    t1 = clock64();
    if (elect_one()) { out[0] = t1 - t0; }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

}  // namespace dfb200

using namespace dfb200;

extern "C" int dfb200_bench_umma(int layout, int N, int iters, int ksteps, long long* out_cycles, dfb200_stream_t stream) {
  DFB_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && ksteps >= 1 && ksteps <= 8, DFB200_ERR_INVALID_ARG, "bench_umma: bad shape");
  const int smem = 131072 + 128;
  DFB_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_rate_kernel<<<1, 128, smem, as_stream(stream)>>>(layout, N, iters, ksteps, out_cycles);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

// Debug hook: device buffer of 1024 int64 that CTA 0 of the fused kernel fills with clock64() stamps (NULL = off).
extern "C" int dfb200_debug_tc_timeline(long long* device_buffer) {
  g_tc_timeline = device_buffer;
  return DFB200_OK;
}

extern "C" int dfb200_selftest_umma(int variant, int N, int K, const float* A, const float* W, const float* bias,
                                    const float* Cin, float* D, void* scratch, dfb200_stream_t stream) {
  DFB_REQUIRE((N == 32 || N == 64 || N == 128) && K >= 16 && K <= 128 && K % 16 == 0, DFB200_ERR_INVALID_ARG,
              "selftest_umma: N in {32,64,128}, K multiple of 16 in [16,128]");
  DFB_REQUIRE(!(variant & 4) || scratch != nullptr, DFB200_ERR_INVALID_ARG, "selftest_umma: bulk variant needs scratch");
  const int smem = 73728 + 128;
  DFB_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(variant, N, K, A, W, bias, Cin, D, reinterpret_cast<uint8_t*>(scratch));
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}
