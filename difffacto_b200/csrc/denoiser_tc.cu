// DFB200_MODE_BF16: the whole cross-diffusion denoiser as ONE fused tcgen05 kernel.
//
// Reference computation: python/difffacto/models/diffusions/nets/attention.py:385-440 (+ blocks
// :161-306).  The reference materialises every activation in HBM (512 B/token per tensor, 4 KB/token
// for the GEGLU intermediate, ~180 kernels per step); here a CTA owns two 128-token tiles for the whole
// network and HBM sees only the 13 input features and the 3 output channels per token:
//   * residual stream x (fp32) lives in TENSOR MEMORY: 128 lanes x 128 columns per tile; the attention-out and
//     FF-out GEMMs accumulate straight into it (tcgen05.mma D += A.B), so the residual add is free;
//   * LayerNorm / softmax-over-4-part-tokens / GEGLU run on CUDA cores out of TMEM (one thread = one token
//     row, packed FFMA2 math) and write the next bf16 A operand into shared memory in the canonical UMMA
//     K-major layout; LN gains/biases and all Linear biases are folded into the packed weights (bias = one
//     extra K=16 MMA against a constant "ones" tile, bf16 hi+lo split so it is fp32-accurate);
//   * cross-attention is FOLDED: the keys/values of a sample are 4 tokens, so per (sample, block) the context
//     kernel builds W_sim = 0.25 K_h.Wq' (32 x 128) and W_pv = Wo_h.V_h (128 x 32); the attention is then
//     S = LN2(x).W_sim^T (N=32 MMA) -> 8 softmaxes over 4 logits per token -> x += P.W_pv^T (K=32 MMA);
//   * weights stream L2 -> smem as pre-packed bf16 UMMA tiles through a 6-slot cp.async.bulk ring fed by
//     a producer warp; both tiles consume each packet (the kernel is shared-memory-bandwidth bound: operand
//     reads of the SS-mode MMAs + ring fills + epilogue stores, so FF runs on N=128 MMAs);
//   * one warp issues all MMAs (uniform datapath); the two tiles PING-PONG through the feed-forward: while the
//     epilogue of one tile applies GEGLU to its 128-column hidden chunk, the tensor pipe runs the other tile's
//     FF-out + next FF-in.
// TMEM map (512 columns): X0 [0,128) X1 [128,256) ACC0 [256,384) ACC1 [384,512).
#include <float.h>

#include "ddpm.cuh"
#include "denoiser.cuh"
#include "philox.cuh"
#include "tc_common.cuh"

namespace dfb200 {
using namespace tc;

// ---------------------------------------------------------------------------------------------
// packets (18 KB slots).  Per block, in MMA consumption order:
//   0, 1 : "fold" packets of tile 0 / tile 1 (per sample+block, rebuilt every step by context_fold_kernel):
//          W_sim tile (32 x 128, 8 KB) | bias_sim slab (512 B) | W_pv tile (128 x 32, 8 KB)
//   static stream, 24 packets per block at an 18 KB stride:
//   s0: W1'_0 k[0,64)                  s1: W1'_0 k[64,128) + bo slab
//   then for c = 0..6:  W2_c | W1'_{c+1} k[0,64) | W1'_{c+1} k[64,128)                      and finally W2_7 + b2 slab
// W1'_c (128 x 128): rows [0,64) = value units 64c.. (scaled by 1/2, the GELU's 1/2), rows [64,128) = gate units
// 512+64c..; W1' = W1.diag(norm3.w), b1' = b1 + W1.norm3.b.  W2_c: 128 rows x k[64c, 64c+64).
// ---------------------------------------------------------------------------------------------
constexpr int SLOT_BYTES = 18432;
constexpr int NSLOT = 8;
constexpr int STATIC_PER_LAYER = 24;
constexpr int PKT_PER_LAYER = 26;
constexpr int FF_CHUNKS = 8;
constexpr int SLAB_OFF = 16384;  // bias slab position inside a static packet
constexpr int FOLD_WSIM = 0, FOLD_BSIM = 8192, FOLD_WPV = 8704, FOLD_BYTES = 16896;

// layer-local packet index p (0..25) -> bytes / static-stream index
__host__ __device__ inline int pkt_bytes(int p) {
  if (p < 2) return FOLD_BYTES;
  const int s = p - 2;
  return (s == 1 || s == 23) ? 16384 + 2048 : 16384;  // W1'_0 B-half carries the bo slab, W2_7 the b2 slab
}
__host__ __device__ inline int spkt_w1a(int c) { return c == 0 ? 0 : 3 + 3 * (c - 1); }  // static index of W1'_c k[0,64)
__host__ __device__ inline int spkt_w1b(int c) { return spkt_w1a(c) + 1; }
__host__ __device__ inline int spkt_w2(int c) { return 2 + 3 * c; }

// Extras appended after the static stream:
//   the "in/head" image (INHEAD_BYTES, copied into shared memory once per CTA):
//     IH_INTILE   proj_in + pre_norm as ONE K=32 UMMA B tile (128 x 32 bf16), see pack_inhead_kernel
//     IH_HEADTILE post_norm + proj_out as a 16 x 128 bf16 B tile (rows 3..15 zero), IH_HEADSLAB its bias slab
//     IH_CONST    fp32 constants of the analytic pre_norm variance: Gt[45] | gp[4][9] | cp[4]
//   then per block WqG (128x128, = 0.25*Wq.diag(norm2.w)), bqG (128, = 0.25*Wq.norm2.b) and WoT (128x128, = Wo^T) for the
//   fold kernel (fp32).
constexpr uint32_t IH_INTILE = 0, IH_HEADTILE = 8192, IH_HEADSLAB = 12288, IH_CONST = 12544, INHEAD_BYTES = 13056;
constexpr int IHC_GT = 0, IHC_GP = 45, IHC_CP = 81;  // float offsets inside IH_CONST (85 floats)
//   then per block b1p (1024 fp32): the folded GEGLU-in bias b1' = b1 + W1.norm3.b in MMA column order -- per chunk c 128 floats,
//   [0,64) = b1'[64c + k] / 2 (value units, the GELU's 1/2), [64,128) = b1'[512 + 64c + k] (gate units).  The feed-forward
//   epilogue adds it on CUDA cores (it used to be a 14th K=16 MMA per tile-chunk against the ones tile).
constexpr size_t HEAD_FLOATS = INHEAD_BYTES / 4;
constexpr size_t FOLDW_FLOATS = (size_t)D_MODEL * D_MODEL * 2 + D_MODEL;
constexpr size_t B1P_FLOATS = 2 * D_FF;

size_t tc_stream_bytes_for(const NetDims& d) {
  return (size_t)d.depth * STATIC_PER_LAYER * SLOT_BYTES + sizeof(float) * (HEAD_FLOATS + d.depth * (FOLDW_FLOATS + B1P_FLOATS));
}
size_t tc_fold_bytes_for(const NetDims& d, int B) { return (size_t)B * d.depth * FOLD_BYTES; }

// ---- pack kernels (run once per weight update) -----------------------------------------------------
// dst: UMMA tile of R rows x KC k-values; element (r,k) = src[rowmap(r)*ld + k0 + k] * (gamma ? gamma[k0+k] : 1)
// geglu_chunk >= 0: rows [0,64) -> value unit 64c+r (x 1/2), rows [64,128) -> gate unit 512+64c+(r-64)
__global__ void pack_tile_kernel(uint8_t* __restrict__ dst, int R, int KC, const float* __restrict__ src, int ld, int k0,
                                 const float* __restrict__ gamma, int geglu_chunk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * KC) return;
  const int r = i / KC, k = i - r * KC;
  int row = r;
  if (geglu_chunk >= 0) row = r < 64 ? 64 * geglu_chunk + r : D_FF + 64 * geglu_chunk + (r - 64);
  float v = __ldg(src + (size_t)row * ld + k0 + k);
  if (gamma != nullptr) v *= __ldg(gamma + k0 + k);
  if (geglu_chunk >= 0 && r < 64) v *= 0.5f;  // a*gelu(g) = (a/2)*g*(1+tanh(..))
  *reinterpret_cast<__nv_bfloat16*>(dst + tile_off(R, r, k)) = __float2bfloat16_rn(v);
}
// bias slab: R rows x 8 k-values (16 B per row); k=0: bf16 hi, k=1: bf16 lo of  bias[row] + W[row,:].beta
__global__ void pack_bias_kernel(uint8_t* __restrict__ dst, int R, const float* __restrict__ bias,
                                 const float* __restrict__ W, int ld, const float* __restrict__ beta, int geglu_chunk) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  int row = r;
  if (geglu_chunk >= 0) row = r < 64 ? 64 * geglu_chunk + r : D_FF + 64 * geglu_chunk + (r - 64);
  float v = bias != nullptr ? __ldg(bias + row) : 0.f;
  if (beta != nullptr)
    for (int k = 0; k < ld; ++k) v = fmaf(__ldg(W + (size_t)row * ld + k), __ldg(beta + k), v);
  if (geglu_chunk >= 0 && r < 64) v *= 0.5f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dst + r * 16);
  o[0] = hi;
  o[1] = lo;
#pragma unroll
  for (int k = 2; k < 8; ++k) o[k] = __float2bfloat16_rn(0.f);
}
// b1p (see B1P_FLOATS): dst[c*128 + j] for chunk c, MMA column j
__global__ void pack_b1p_kernel(float* __restrict__ dst, const float* __restrict__ b1, const float* __restrict__ W1,
                                const float* __restrict__ beta3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * D_FF) return;
  const int c = i >> 7, j = i & 127;
  const int row = j < 64 ? 64 * c + j : D_FF + 64 * c + (j - 64);
  float v = __ldg(b1 + row);
  for (int k = 0; k < D_MODEL; ++k) v = fmaf(__ldg(W1 + (size_t)row * D_MODEL + k), __ldg(beta3 + k), v);
  dst[i] = j < 64 ? 0.5f * v : v;
}
// proj_in (13 -> 128) + pre_norm as one MMA.  With h = W f + b (f = the 13 point features, of which f[9..12] is a one-hot
// class p), LN(h)_n = rstd * ((W_n - wbar).f + (b_n - bbar)) * g_n + beta_n where wbar/bbar are the means over the 128
// outputs, and var(h) is a quadratic form of f that every token evaluates on CUDA cores from 58 constants (IH_CONST):
//   var = f9' Gt f9 + gp[p].f9 + cp[p]          (f9 = f[0..8]; Gt upper-triangular with doubled off-diagonals)
// so the token scales its features by rstd itself and the MMA produces the normalised residual stream directly.
// A row (K = 32 bf16), written by the token's thread:   k = 2i, 2i+1 : hi, lo of rstd*f_i (i < 9)
//   k = 18+3c+{0,1,2} : (rhi, rlo, rhi) of rstd if c == p else 0          k = 30, 31 : 1, 1
// B row n (this kernel):   k = 2i, 2i+1 : W''_ni = (W_ni - wbar_i) g_n      k = 18+3c+{0,1,2} : (chi, chi, clo) of
//   W''_n,9+c + (b_n - bbar) g_n          k = 30, 31 : hi, lo of beta_n.
// Also the folded head: w_out' = w_out.diag(post_norm.w) (bf16 tile), b_out' = b_out + w_out.post_norm.b (hi+lo slab).
__global__ void __launch_bounds__(128)
pack_inhead_kernel(uint8_t* __restrict__ dst, const float* __restrict__ w_in, const float* __restrict__ b_in,
                   const float* __restrict__ pre_w, const float* __restrict__ pre_b, const float* __restrict__ w_out,
                   const float* __restrict__ b_out, const float* __restrict__ post_w, const float* __restrict__ post_b) {
  __shared__ float wc[D_MODEL][14];  // centred [W_n - wbar | b_n - bbar]
  __shared__ float mean14[14];
  __shared__ float G[14][14];
  const int n = threadIdx.x;
  for (int c = 0; c < 13; ++c) wc[n][c] = __ldg(w_in + n * 13 + c);
  wc[n][13] = __ldg(b_in + n);
  __syncthreads();
  if (n < 14) {
    float m = 0.f;
    for (int k = 0; k < D_MODEL; ++k) m += wc[k][n];
    mean14[n] = m * (1.f / D_MODEL);
  }
  __syncthreads();
  for (int c = 0; c < 14; ++c) wc[n][c] -= mean14[c];
  __syncthreads();
  for (int e = n; e < 196; e += 128) {
    const int i = e / 14, j = e - i * 14;
    float a = 0.f;
    for (int k = 0; k < D_MODEL; ++k) a = fmaf(wc[k][i], wc[k][j], a);
    G[i][j] = a * (1.f / D_MODEL);
  }
  __syncthreads();
  float* cst = reinterpret_cast<float*>(dst + IH_CONST);
  if (n < 45) {  // upper triangle of the 9 x 9 feature block, row-major (i, j >= i)
    int i = 0, e = n;
    while (e >= 9 - i) { e -= 9 - i; ++i; }
    const int j = i + e;
    cst[IHC_GT + n] = G[i][j] * (i == j ? 1.f : 2.f);
  } else if (n < 81) {
    const int c = (n - 45) / 9, i = (n - 45) - c * 9;
    cst[IHC_GP + c * 9 + i] = 2.f * (G[i][9 + c] + G[i][13]);
  } else if (n < 85) {
    const int c = n - 81;
    cst[IHC_CP + c] = G[9 + c][9 + c] + 2.f * G[9 + c][13] + G[13][13];
  }
  // B tile row n
  const float g = __ldg(pre_w + n), beta = __ldg(pre_b + n);
  auto put = [&](int k, float v) { *reinterpret_cast<__nv_bfloat16*>(dst + IH_INTILE + tile_off(128, n, k)) = __float2bfloat16_rn(v); };
  for (int i = 0; i < 9; ++i) { put(2 * i, wc[n][i] * g); put(2 * i + 1, wc[n][i] * g); }
  for (int c = 0; c < 4; ++c) {
    const float v = (wc[n][9 + c] + wc[n][13]) * g;
    const float hi = __bfloat162float(__float2bfloat16_rn(v));
    put(18 + 3 * c, hi); put(19 + 3 * c, hi); put(20 + 3 * c, v - hi);
  }
  {
    const float hi = __bfloat162float(__float2bfloat16_rn(beta));
    put(30, hi); put(31, beta - hi);
  }
  // head tile (16 rows) + slab
  for (int r = 0; r < 16; ++r)
    *reinterpret_cast<__nv_bfloat16*>(dst + IH_HEADTILE + tile_off(16, r, n)) =
        __float2bfloat16_rn(r < 3 ? __ldg(w_out + r * D_MODEL + n) * __ldg(post_w + n) : 0.f);
  if (n < 16) {
    float v = 0.f;
    if (n < 3) {
      v = __ldg(b_out + n);
      for (int k = 0; k < D_MODEL; ++k) v = fmaf(__ldg(w_out + n * D_MODEL + k), __ldg(post_b + k), v);
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(dst + IH_HEADSLAB + n * 16);
    o[0] = hi; o[1] = lo;
    for (int k = 2; k < 8; ++k) o[k] = __float2bfloat16_rn(0.f);
  }
}
// fold-kernel weights: WqG[n][k] = 0.25 Wq[n][k] g2[k];  bqG[n] = 0.25 sum_k Wq[n][k] b2[k];  WoT[k][c] = Wo[c][k]
__global__ void pack_foldw_kernel(float* __restrict__ dst, const float* __restrict__ wq, const float* __restrict__ g2,
                                  const float* __restrict__ b2, const float* __restrict__ wo) {
  const int n = blockIdx.x, k = threadIdx.x;  // 128 x 128
  const float w = __ldg(wq + n * D_MODEL + k);
  dst[n * D_MODEL + k] = 0.25f * w * __ldg(g2 + k);
  float acc = w * __ldg(b2 + k);
  for (int d = 16; d >= 1; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
  __shared__ float part[4];
  if ((k & 31) == 0) part[k >> 5] = acc;
  __syncthreads();
  if (k == 0) dst[D_MODEL * D_MODEL + n] = 0.25f * (part[0] + part[1] + part[2] + part[3]);
  dst[D_MODEL * D_MODEL + D_MODEL + k * D_MODEL + n] = __ldg(wo + n * D_MODEL + k);  // WoT[k][n] = Wo[n][k]
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
  float* extras = reinterpret_cast<float*>(S + (size_t)L.d.depth * STATIC_PER_LAYER * SLOT_BYTES);
  for (int l = 0; l < L.d.depth; ++l) {
    const size_t* o = L.blk[l];
    uint8_t* base = S + (size_t)l * STATIC_PER_LAYER * SLOT_BYTES;
    auto pk = [&](int p) { return base + (size_t)p * SLOT_BYTES; };
    for (int c = 0; c < FF_CHUNKS; ++c) {
      tile(pk(spkt_w1a(c)), 128, 64, P + o[B_W1], D_MODEL, 0, P + o[B_N3_W], c);
      tile(pk(spkt_w1b(c)), 128, 64, P + o[B_W1], D_MODEL, 64, P + o[B_N3_W], c);
      tile(pk(spkt_w2(c)), 128, 64, P + o[B_W2], D_FF, 64 * c, nullptr, -1);
    }
    bias(pk(spkt_w1b(0)) + SLAB_OFF, 128, P + o[B_BO], nullptr, 0, nullptr, -1);
    bias(pk(spkt_w2(FF_CHUNKS - 1)) + SLAB_OFF, 128, P + o[B_B2], nullptr, 0, nullptr, -1);
    pack_foldw_kernel<<<D_MODEL, D_MODEL, 0, st>>>(extras + HEAD_FLOATS + (size_t)l * FOLDW_FLOATS, P + o[B_WQ], P + o[B_N2_W],
                                                   P + o[B_N2_B], P + o[B_WO]);
    count_launch();
    pack_b1p_kernel<<<cdiv(2 * D_FF, 256), 256, 0, st>>>(extras + HEAD_FLOATS + (size_t)L.d.depth * FOLDW_FLOATS + (size_t)l * B1P_FLOATS,
                                                         P + o[B_B1], P + o[B_W1], P + o[B_N3_B]);
    count_launch();
  }
  pack_inhead_kernel<<<1, 128, 0, st>>>(reinterpret_cast<uint8_t*>(extras), P + L.g[P_IN_W], P + L.g[P_IN_B], P + L.g[P_PRE_W], P + L.g[P_PRE_B],
                                        P + L.g[P_OUT_W], P + L.g[P_OUT_B], P + L.g[P_POST_W], P + L.g[P_POST_B]);
  count_launch();
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

// ---------------------------------------------------------------------------------------------
// per-step fold kernel: K/V of a sample (4 tokens) folded into the attention weights of each block
//   W_sim[(h,j)][k] = sum_d K[j][16h+d] WqG[16h+d][k]        (0.25 and norm2 gain already in WqG)
//   b_sim[(h,j)]    = sum_d K[j][16h+d] bqG[16h+d]
//   W_pv[c][(h,j)]  = sum_d Wo[c][16h+d] V[j][16h+d]
// written as bf16 UMMA tiles (the "fold" packet of (b, l)).  grid (B, depth), 256 threads.
// ---------------------------------------------------------------------------------------------
// kv: [B][depth][2][4][128] (full K/V, or their static half when kv_time != NULL);  kv_time: [T][depth][2][128] time half of
// timestep step_t[blockIdx.z] (or t_first - blockIdx.z when step_t == NULL), broadcast over the 4 tokens.  grid (B, depth, steps).
// W_sim and b_sim carry an extra factor log2(e): the softmax of the fused kernel is exp2(s' - max s') / sum, one MUFU.EX2 per logit.
constexpr float LOG2E = 1.4426950408889634f;
constexpr int FOLD_SPB = 8;  // sampling steps folded per CTA: its 128 + 128 weight columns are read once (registers) for all of them
__global__ void __launch_bounds__(512)
context_fold_kernel(int depth, const float* __restrict__ kv, const float* __restrict__ kv_time, int t_first,
                    const int* __restrict__ step_t, int steps, const float* __restrict__ extras, uint8_t* __restrict__ fold_all) {
  __shared__ float K[MAX_TOKENS][D_MODEL], V[MAX_TOKENS][D_MODEL];
  // the packet is assembled in shared memory and leaves as coalesced 16-byte stores: the tile layout scatters a thread's
  // bf16 values 16..512 bytes apart, and 66 2-byte global stores per thread made this kernel store-issue bound
  __shared__ __align__(16) uint8_t img[FOLD_BYTES];
  const int b = blockIdx.x, l = blockIdx.y, t = threadIdx.x;
  const float* src = kv + ((size_t)b * depth + l) * 1024;
  const float* WqG = extras + HEAD_FLOATS + (size_t)l * FOLDW_FLOATS;
  const float* bqG = WqG + D_MODEL * D_MODEL;
  const float* WoT = bqG + D_MODEL;
  // 512 threads: threads [0,256) build W_sim, [256,512) W_pv; within a role thread = (column, half of the heads)
  const int col = t & 127, half = (t >> 7) & 1, role = t >> 8;
  // Every step of a (sample, block) multiplies the SAME weight columns: one L2 read per CTA instead of one per step (the kernel
  // was L2-bound on these 128 KB per packet: 2.7 ms per 1000-step loop at the BASELINE size).
  float w[4][16];
#pragma unroll
  for (int hh = 0; hh < 4; ++hh) {
#pragma unroll
    for (int d = 0; d < 16; ++d) w[hh][d] = __ldg((role == 0 ? WqG : WoT) + (16 * (half * 4 + hh) + d) * D_MODEL + col);
  }
  const int s_lo = blockIdx.z * FOLD_SPB, s_hi = min(steps, s_lo + FOLD_SPB);
  for (int sidx = s_lo; sidx < s_hi; ++sidx) {
    uint8_t* fold = fold_all + (size_t)sidx * gridDim.x * depth * FOLD_BYTES;
    __syncthreads();  // the previous step's packet has left `img`, K / V are free
    for (int i = t; i < 512; i += 512) {
      float kt = 0.f, vt = 0.f;
      if (kv_time != nullptr) {
        const int tcur = step_t != nullptr ? __ldg(step_t + sidx) : t_first - sidx;
        const float* tt = kv_time + ((size_t)tcur * depth + l) * 2 * D_MODEL;
        kt = __ldg(tt + (i & 127));
        vt = __ldg(tt + D_MODEL + (i & 127));
      }
      (&K[0][0])[i] = __ldg(src + i) + kt;
      (&V[0][0])[i] = __ldg(src + 512 + i) + vt;
    }
    __syncthreads();
    uint8_t* gout = fold + ((size_t)b * depth + l) * FOLD_BYTES;
    uint8_t* out = img;
    // role 0, W_sim: thread = column k, rows r = (h,j) with h in this half's 4 heads;  role 1, W_pv: thread = output channel c,
    // columns r = (h,j)
    const float (*KV)[D_MODEL] = role == 0 ? K : V;
#pragma unroll
    for (int hh = 0; hh < 4; ++hh) {
      const int h = half * 4 + hh;
      float acc[MAX_TOKENS] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int d = 0; d < 16; ++d) {
#pragma unroll
        for (int j = 0; j < MAX_TOKENS; ++j) acc[j] = fmaf(KV[j][16 * h + d], w[hh][d], acc[j]);
      }
#pragma unroll
      for (int j = 0; j < MAX_TOKENS; ++j) {
        if (role == 0) *reinterpret_cast<__nv_bfloat16*>(out + FOLD_WSIM + tile_off(32, h * 4 + j, col)) = __float2bfloat16_rn(LOG2E * acc[j]);
        else *reinterpret_cast<__nv_bfloat16*>(out + FOLD_WPV + tile_off(128, col, h * 4 + j)) = __float2bfloat16_rn(acc[j]);
      }
    }
    if (t < 32) {  // bias slab of the logits
      const int h = t >> 2, j = t & 3;
      float v = 0.f;
      for (int d = 0; d < 16; ++d) v = fmaf(K[j][16 * h + d], __ldg(bqG + 16 * h + d), v);
      v *= LOG2E;
      const __nv_bfloat16 hi = __float2bfloat16_rn(v);
      const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out + FOLD_BSIM + t * 16);
      o[0] = hi; o[1] = lo;
#pragma unroll
      for (int k = 2; k < 8; ++k) o[k] = __float2bfloat16_rn(0.f);
    }
    __syncthreads();
    static_assert(FOLD_BYTES % 16 == 0, "fold packet must be a whole number of 16-byte chunks");
    for (int i = t; i < FOLD_BYTES / 16; i += 512)
      __stcs(reinterpret_cast<uint4*>(gout) + i, reinterpret_cast<const uint4*>(img)[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// the fused kernel
// ---------------------------------------------------------------------------------------------
// FF-out A operand (the gated activations U) is handed to the tensor core through TENSOR MEMORY: the epilogue writes the
// bf16 pairs over the value columns of the hidden chunk it has just consumed and the FF-out MMA runs in TS form.  This
// removes 16 KB of st.shared + 16 KB of MMA operand reads per tile-chunk from the shared-memory pipe.
constexpr int TC_THREADS = 320;  // warps 0-3: tile 0 epilogue, 4-7: tile 1 epilogue, 8: MMA issuer, 9: weight producer
constexpr uint32_t SM_A = 0;               // 2 x 32768  A operand tiles (128 x 128 bf16)
constexpr uint32_t SM_ONES = 65536;        // 4096  ones tile (128 x 16 bf16: k=0,1 -> 1)
constexpr uint32_t SM_RING = 69632;        // NSLOT x 18432
constexpr uint32_t SM_BAR = SM_RING + NSLOT * SLOT_BYTES;  // mbarriers
constexpr uint32_t SM_TMEM = SM_BAR + 256;
constexpr uint32_t SM_INHEAD = SM_BAR + 512;  // the in/head image (INHEAD_BYTES), staged once per CTA
constexpr uint32_t TC_SMEM_BYTES = SM_INHEAD + INHEAD_BYTES;
static_assert(TC_SMEM_BYTES <= 232448, "shared memory budget");

enum Bar { BAR_A = 0 /*[2]*/, BAR_ACC = 2 /*[2]*/, BAR_UREADY = 4 /*[2]*/, BAR_X = 6 /*[2]*/, BAR_WFULL = 8 /*[NSLOT]*/,
           BAR_WEMPTY = 8 + NSLOT /*[NSLOT]*/, BAR_COUNT = 8 + 2 * NSLOT };
static_assert(BAR_COUNT * 8 <= 256, "barrier block");

struct TcParams {
  const uint8_t* stream;  // static packets
  const uint8_t* fold;    // [B][depth] fold packets
  const uint8_t* inhead;  // in/head image (see IH_*)
  const float* x; const float* anchors; const float* variances; const int* assign; const float* valid;
  float* eps_out;
  int N, depth, flags;
  long long M;
  long long* dbg;  // optional timeline buffer: CTA 0 records clock64() at the phase boundaries of its dbg_item-th work item
  int dbg_item;
  int abl;  // diagnostic build only: ablation bits (timing experiments, results invalid): 1 no GEGLU math, 2 no FF LDTM, 4 no ring copies, 8 no FF-in MMAs, 16 no FF-out MMAs
  // persistent work list: item idx = step_local * n_units + unit, CTA c takes idx = c, c + gridDim.x, ...
  int n_units, n_steps, t_first;  // units of 2 tiles; timesteps t_first, t_first-1, ... (n_steps of them) unless step_t is given
  const int* step_t;              // optional device list of this launch's timesteps in execution order (DDIM strides)
  int step_base;                  // sampling steps completed by earlier launches of the loop (base of the `done` counters)
  size_t fold_step_bytes;         // distance between the fold packets of consecutive steps
  int* done;                      // per unit: number of tile-steps completed since the loop began (cross-CTA dependency), or NULL
  const float* b1p;               // folded GEGLU-in biases, [depth][8][128] fp32 (see B1P_FLOATS)
  // fused eps -> x_{t-1} update (sampling loop): active when upd_sched != NULL
  const float* upd_sched; int upd_T; const float* upd_noise; size_t noise_step_elems; uint64_t upd_seed; float* x_out;
  float* traj; int traj_interval;
  float* step_sample; float* step_xstart;  // optional (n_steps,B,3,N): `sample` / `pred_xstart` of every step of this launch
  float* step_s[TC_MAX_LIST_STEPS]; float* step_x[TC_MAX_LIST_STEPS];  // or one (B,3,N) buffer per step (NULL entries = not wanted)
  const float* ddim_acp; const float* ddim_dir; float ddim_eta;  // DDIM update instead of the ancestral one when ddim_acp != NULL
  // classifier-free guidance (anchored_diffusion.py:263-266): a unit is ONE 128-token tile run twice -- tile slot 0 with the
  // conditional fold packets, slot 1 with the unconditional ones (fold entries [B, 2B) of a step) -- and the head mixes
  // eps = (1 - w) * eps_uncond + w * eps_cond before the update.
  int guidance; float guid_w; int B;
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
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// GEGLU for two columns: (a/2 arrives from the MMA) * g * (1 + tanh(g*(c0 + c1 g^2))).
// tanh-form GELU with (c0, c1) refit against the exact erf GELU (max abs error 2.7e-4 over R, below the bf16
// rounding of the result); 5 packed FMA-pipe instructions + 2 MUFU.TANH per column pair (erff costs ~45/column).
__device__ __forceinline__ float2 geglu2(float2 a_half, float2 g, float2 ba_half, float2 bg) {
  a_half = __fadd2_rn(a_half, ba_half);  // the GEGLU-in bias b1' (value half pre-scaled by 1/2), added here instead of by an MMA
  g = __fadd2_rn(g, bg);
  const float2 g2 = __fmul2_rn(g, g);
  const float2 in = __fmul2_rn(g, __ffma2_rn(g2, f2s(0.034700932528f), f2s(0.800156991001f)));
  const float2 t = f2(tanh_approx(in.x), tanh_approx(in.y));
  const float2 ag = __fmul2_rn(a_half, g);
  return __ffma2_rn(ag, t, ag);
}

// timeline instrumentation (diagnostic build only, and off unless a buffer is supplied): slot layout [who][event],
// who 0 = tile-0 row 0, 1 = MMA lane
#ifdef DFB200_DIAGNOSTICS
#define ABL(bit) ((P.abl & (bit)) != 0)
#define TL(who, ev)                                                                                  \
  do {                                                                                               \
    if (P.dbg != nullptr && blockIdx.x == 0 && item_n == P.dbg_item && tl_on) P.dbg[(who) * 512 + (ev)] = clock64(); \
  } while (0)
#else
#define ABL(bit) false
#define TL(who, ev) do { (void)tl_on; } while (0)
#endif

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
// LayerNorm of the row -> bf16 A-operand row.  (A one-pass form that keeps the row in 128 registers was measured slower: the phase
// is bound by the ~190 FMA-pipe instructions per row, not by the TMEM round trips, and the registers pushed state to local memory.)
__device__ __forceinline__ void row_layernorm_to_tile(uint32_t taddr, uint8_t* tile, int r) {
  float mean, rstd;
  row_stats(taddr, mean, rstd);
  row_normalize_to_tile(taddr, mean, rstd, tile, r);
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

__global__ void __launch_bounds__(TC_THREADS, 1) denoiser_tc_kernel(const TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_TMEM);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);  // warp-uniform by construction (helps uniform-datapath codegen)

  // ---- one-time setup ----
  if (warp == 8 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars[BAR_A + i], 128); mbar_init(&bars[BAR_X + i], 1);
      mbar_init(&bars[BAR_ACC + i], 1); mbar_init(&bars[BAR_UREADY + i], 256);
    }
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&bars[BAR_WFULL + i], 1); mbar_init(&bars[BAR_WEMPTY + i], 1); }
    fence_barrier_init();
  }
  if (tid < 256) {  // ones tile: slab 0 (k 0..7) = {1,1,0,...}, slab 1 (k 8..15) = 0
    uint4 v = make_uint4(tid < 128 ? 0x3F803F80u : 0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(smem + SM_ONES + tid * 16) = v;
  }
  for (int i = tid; i < (int)(INHEAD_BYTES / 16); i += TC_THREADS)
    reinterpret_cast<uint4*>(smem + SM_INHEAD)[i] = __ldg(reinterpret_cast<const uint4*>(P.inhead) + i);
  fence_proxy_async();
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tmem != 0u) __trap();  // a 512-column allocation owns the whole TMEM: base = lane 0, column 0 (the MMA path relies on it)

  // sample index of each tile (all 128 rows of a tile belong to one sample: N % 128 == 0); an out-of-range second tile
  // recomputes the last valid one
  const long long n_tiles = P.M / 128;
  const long long total_items = (long long)P.n_units * P.n_steps;

  if (warp < 8) {
    // =========================== epilogue warps: one thread per token row ===========================
    const int T = warp >> 2, r = tid & 127;
    const uint32_t lane_base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t X = lane_base + T * 128, ACC = lane_base + 256 + T * 128;
    uint8_t* a_tile = smem + SM_A + T * 32768;
    const bool tl_on = tid == 0;
    uint32_t ph_acc = 0, ph_acc_oth = 0, ph_x = 0;  // ph_acc_oth: BAR_ACC of the OTHER tile (FF phase works on both)
    const bool upd = P.upd_sched != nullptr;
    int item_n = 0;
#pragma unroll 1
    for (long long idx = blockIdx.x; idx < total_items; idx += gridDim.x, ++item_n) {
    const int unit = (int)(idx % P.n_units), step_local = (int)(idx / P.n_units);
    const int t_cur = P.step_t != nullptr ? __ldg(P.step_t + step_local) : P.t_first - step_local;
    const long long tile_id = P.guidance ? (long long)unit : (long long)unit * 2 + T;  // guidance: both slots run the same tile
    const bool tile_ok = tile_id < n_tiles;
    const long long tok = (tile_ok ? tile_id : n_tiles - 1) * 128 + r;
    const long long b = tok / P.N;
    const int p = (int)(tok - b * P.N);
    uint32_t vmask = 0;  // bit j set = part token j is valid
#pragma unroll
    for (int j = 0; j < MAX_TOKENS; ++j)
      if (P.valid == nullptr || __ldg(P.valid + b * MAX_TOKENS + j) != 0.f) vmask |= 1u << j;
    TL(0, 0);
#ifdef DFB200_DIAGNOSTICS
    const long long dbg_t0 = (P.dbg != nullptr && tl_on) ? clock64() : 0;
    if (P.dbg != nullptr && blockIdx.x == 0 && item_n == P.dbg_item && tl_on) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      P.dbg[220] = (long long)ns;
    }
#endif
    // Per-token operands that do not change over the sampling steps are requested BEFORE the cross-step dependency wait.
    float f[9];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      f[3 + c] = __ldg(P.anchors + (b * 3 + c) * P.N + p);
      const float v = __ldg(P.variances + (b * 3 + c) * P.N + p);
      f[6 + c] = (P.flags & DFB200_NET_INCLUDE_STD) ? sqrtf(v) : v;
    }
    const int part = __ldg(P.assign + tok);
    if (P.done != nullptr) {
      // x of this unit at this step is produced by the item (unit, previous step), possibly on another SM
      if (r == 0) {
        const int need = 2 * (P.step_base + step_local);
        int have;
        do {
          asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(have) : "l"(P.done + unit) : "memory");
        } while (have < need);
      }
      named_bar_sync(1 + T, 128);
    }
    TL(0, 210);
#ifdef DFB200_DIAGNOSTICS
    if (P.dbg != nullptr && tl_on && blockIdx.x < 148) P.dbg[742 + blockIdx.x] += clock64() - dbg_t0;  // per-CTA dependency-wait cycles
#endif

    // ---- proj_in (13 -> 128) + pre_norm on the tensor core: analytic variance -> scaled K=32 A row -> one MMA into X ----
    {
      const float* cst = reinterpret_cast<const float*>(smem + SM_INHEAD + IH_CONST);
#pragma unroll
      for (int c = 0; c < 3; ++c) f[c] = __ldcg(P.x + (b * 3 + c) * P.N + p);  // x is rewritten every step by other SMs: read through L2
      TL(0, 211);
      float var = cst[IHC_CP + part];
      {
        int e = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          float t = cst[IHC_GP + part * 9 + i];
#pragma unroll
          for (int j = i; j < 9; ++j, ++e) t = fmaf(cst[IHC_GT + e], f[j], t);
          var = fmaf(f[i], t, var);
        }
      }
      const float rstd = rsqrtf(fmaxf(var, 0.f) + LN_EPS);
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 9; ++i) {
        const float sv = rstd * f[i];
        const float hi = __bfloat162float(__float2bfloat16_rn(sv));
        w[i] = pack_bf16(hi, sv - hi);
      }
      {
        const float rhi = __bfloat162float(__float2bfloat16_rn(rstd));
        const float rlo = rstd - rhi;
        // entries k = 18 + 3c + m: (rhi, rlo, rhi) for c == part, else 0; two entries per word, word 9 starts at k = 18
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const int k0 = 2 * q, k1 = 2 * q + 1;  // entry indices inside the 12-entry class block
          const float e0 = (k0 / 3 == part) ? (k0 % 3 == 1 ? rlo : rhi) : 0.f;
          const float e1 = (k1 / 3 == part) ? (k1 % 3 == 1 ? rlo : rhi) : 0.f;
          w[9 + q] = pack_bf16(e0, e1);
        }
      }
      w[15] = 0x3F803F80u;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(a_tile + j * 2048 + r * 16) = make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars[BAR_A + T]);
      TL(0, 212);
      mbar_wait(&bars[BAR_X + T], ph_x);  // x = pre_norm(proj_in(features)) is in TMEM
      ph_x ^= 1;
      tc_fence_after();
    }

    TL(0, 1);
    for (int l = 0; l < P.depth; ++l) {
      TL(0, 2 + l * 40);
      // ---- LN2 -> A ----
      row_layernorm_to_tile(X, a_tile, r);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars[BAR_A + T]);
      TL(0, 3 + l * 40);

      // ---- folded cross-attention: logits S[(h,j)] (x log2 e) arrive from the MMA; softmax over the 4 part tokens per head ----
      mbar_wait(&bars[BAR_ACC + T], ph_acc);
      ph_acc ^= 1;
      tc_fence_after();
      TL(0, 4 + l * 40);
      {
        float sv[32];
        tmem_ld32(ACC, sv);
        tmem_wait_ld();
        uint32_t pw[16];
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          float s0 = (vmask & 1u) ? sv[4 * h] : -FLT_MAX, s1 = (vmask & 2u) ? sv[4 * h + 1] : -FLT_MAX;  // masked_fill(~mask, -finfo.max)
          float s2 = (vmask & 4u) ? sv[4 * h + 2] : -FLT_MAX, s3 = (vmask & 8u) ? sv[4 * h + 3] : -FLT_MAX;
          const float mx = fmaxf(fmaxf(s0, s1), fmaxf(s2, s3));
          s0 = ex2_approx(s0 - mx); s1 = ex2_approx(s1 - mx); s2 = ex2_approx(s2 - mx); s3 = ex2_approx(s3 - mx);
          const float inv = __fdividef(1.f, (s0 + s1) + (s2 + s3));
          pw[2 * h] = pack_bf16(s0 * inv, s1 * inv);
          pw[2 * h + 1] = pack_bf16(s2 * inv, s3 * inv);
        }
        // P (128 x 32 bf16) reuses the first 4 k-slabs of the A tile (the logits MMA has finished reading it)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(a_tile + j * 2048 + r * 16) = make_uint4(pw[4 * j], pw[4 * j + 1], pw[4 * j + 2], pw[4 * j + 3]);
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars[BAR_A + T]);
      TL(0, 5 + l * 40);

      // ---- x += P W_pv^T + bo (accumulated in TMEM by the MMA warp);  LN3 -> A ----
      mbar_wait(&bars[BAR_X + T], ph_x);
      ph_x ^= 1;
      tc_fence_after();
      TL(0, 6 + l * 40);
      row_layernorm_to_tile(X, a_tile, r);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars[BAR_A + T]);
      TL(0, 7 + l * 40);

      // ---- GEGLU feed-forward: 8 chunks of 64 value + 64 gate columns ----
      // All 8 epilogue warps work on EACH tile's chunk: warps 0-3 take value/gate columns [0,32), warps 4-7 columns
      // [32,64) of the same 128 rows (warp w and w+4 own the same TMEM lanes).  The GEGLU latency of a tile-chunk is
      // halved, which takes it below the tensor-pipe time of the other tile's FF-out + FF-in: the FF phase becomes
      // MMA bound.  The gated activations go back to TMEM over the value columns this thread has just consumed
      // (U columns [32h, 32h+16) for column half h) and FF-out reads them in TS form.
      // Skip the other tile's logits phase of BAR_ACC: it has completed (this thread is past its own tile's P.W_pv commit,
      // which the MMA warp issued after both logits MMAs), and it must NOT be waited for here: the barrier may already be
      // two phases on (the other tile's H_0 does not depend on this thread), where a parity wait would never return.
      ph_acc_oth ^= 1;
      const float4* b1p = reinterpret_cast<const float4*>(P.b1p + (size_t)l * B1P_FLOATS) + T * 8;  // this column half's biases
#pragma unroll 1
      for (int c = 0; c < FF_CHUNKS; ++c) {
        // the chunk's folded biases (uniform addresses, L2-resident), requested before the accumulator wait; both tiles use them
        float ba[32], bg[32];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 va = __ldg(b1p + c * 32 + k), vg = __ldg(b1p + c * 32 + 16 + k);
          ba[4 * k] = va.x; ba[4 * k + 1] = va.y; ba[4 * k + 2] = va.z; ba[4 * k + 3] = va.w;
          bg[4 * k] = vg.x; bg[4 * k + 1] = vg.y; bg[4 * k + 2] = vg.z; bg[4 * k + 3] = vg.w;
        }
#pragma unroll
        for (int TT = 0; TT < 2; ++TT) {
          const uint32_t ACCt = lane_base + 256 + TT * 128 + T * 32;
          mbar_wait(&bars[BAR_ACC + TT], TT == T ? ph_acc : ph_acc_oth);  // H_c of tile TT is ready
          if (TT == T) ph_acc ^= 1; else ph_acc_oth ^= 1;
          tc_fence_after();
          if (TT == 0) TL(0, 8 + l * 40 + c * 2);
          float a[32], gt[32];
#ifdef DFB200_DIAGNOSTICS
#pragma unroll
          for (int k = 0; k < 32; ++k) a[k] = gt[k] = 0.f;
          if (!ABL(2))
#endif
          {
            tmem_ld32(ACCt, a);
            tmem_ld32(ACCt + 64, gt);
          }
          tmem_wait_ld();
          uint32_t u[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            float2 y;
            if (ABL(1))  // diagnostic build: no GEGLU math (timing experiment, results invalid)
              y = f2(a[2 * k] + ba[2 * k] + bg[2 * k + 1], gt[2 * k + 1] + a[2 * k + 1] + gt[2 * k] + ba[2 * k + 1] + bg[2 * k]);
            else
              y = geglu2(f2(a[2 * k], a[2 * k + 1]), f2(gt[2 * k], gt[2 * k + 1]), f2(ba[2 * k], ba[2 * k + 1]), f2(bg[2 * k], bg[2 * k + 1]));
            u[k] = pack_bf16(y.x, y.y);
          }
          tmem_st16(ACCt, u);
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&bars[BAR_UREADY + TT]);
          if (TT == 0) TL(0, 9 + l * 40 + c * 2);
        }
      }
      mbar_wait(&bars[BAR_X + T], ph_x);
      ph_x ^= 1;
      tc_fence_after();
    }

    TL(0, 2 + P.depth * 40);
    // ---- post_norm (folded) + proj_out (128 -> 3) on the tensor core (N = 16 MMA, 3 live columns) ----
    {
      TL(0, 204);
      row_layernorm_to_tile(X, a_tile, r);
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(&bars[BAR_A + T]);
      ph_acc_oth ^= 1;  // the other tile's head phase of BAR_ACC (never waited for here; see the FF loop)
      // While the head MMA runs: everything of the update that does not depend on eps (operand loads, noise).
      StepCoef cf{};
      float xv[3] = {0.f, 0.f, 0.f}, av[3] = {0.f, 0.f, 0.f}, vv[3] = {0.f, 0.f, 0.f}, zq[3] = {0.f, 0.f, 0.f};
      if (upd) {
        cf = load_step_coef(P.upd_sched, P.upd_T, t_cur);
        const float* znoise = P.upd_noise != nullptr ? P.upd_noise + (size_t)step_local * P.noise_step_elems : nullptr;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const long long e = (b * 3 + c) * P.N + p;
          xv[c] = __ldcg(P.x + e); av[c] = __ldg(P.anchors + e); vv[c] = __ldg(P.variances + e);
          if (znoise != nullptr) zq[c] = __ldg(znoise + e);
        }
        if (znoise == nullptr) {
          // Philox: the 4 lanes of a quad own 4 consecutive points, i.e. the 4 outputs of ONE Philox call per channel.
          // Lane j of the quad (j < 3) draws channel j's float4; 4 shuffle rounds transpose it (round m: lane j reads lane
          // (j+m)%4, which exposes component (src - m) % 4 = j).  One call per thread instead of three.
          const int j = lane & 3;
          float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < 3) z4 = philox_normal4((uint64_t)(((b * 3 + j) * P.N + (p & ~3)) >> 2), (uint64_t)t_cur, P.upd_seed);
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            const int comp = (j - m) & 3;  // component this lane exposes in round m
            const float mine = comp == 0 ? z4.x : comp == 1 ? z4.y : comp == 2 ? z4.z : z4.w;
            const int src = (j + m) & 3;   // channel received in round m
            const float got = __shfl_sync(0xFFFFFFFFu, mine, (lane & ~3) | src);
            if (src == 0) zq[0] = got; else if (src == 1) zq[1] = got; else if (src == 2) zq[2] = got;
          }
        }
      }
      mbar_wait(&bars[BAR_ACC + T], ph_acc);
      ph_acc ^= 1;
      tc_fence_after();
      float ev[16];
      tmem_ld16(ACC, ev);
      tmem_wait_ld();
      float eps3[3] = {ev[0], ev[1], ev[2]};
      TL(0, 205);
      if (P.guidance) {
        // slot 1 (unconditional pass) hands its eps to slot 0 through its A tile (the head MMA has finished reading it)
        float* xch = reinterpret_cast<float*>(smem + SM_A + 32768);
        if (T == 1) {
#pragma unroll
          for (int c = 0; c < 3; ++c) xch[c * 128 + r] = eps3[c];
        }
        named_bar_sync(3, 256);
        if (T == 0) {
#pragma unroll
          for (int c = 0; c < 3; ++c) eps3[c] = __fadd_rn(__fmul_rn(1.f - P.guid_w, xch[c * 128 + r]), __fmul_rn(P.guid_w, eps3[c]));
        }
        named_bar_sync(3, 256);  // slot 1 must not start its next item's A row before slot 0 has read the exchange
      }
      const bool writer = tile_ok && (!P.guidance || T == 0);
      if (writer && P.eps_out != nullptr) {
#pragma unroll
        for (int c = 0; c < 3; ++c) P.eps_out[(b * 3 + c) * P.N + p] = eps3[c];
      }
      if (upd) {
        // anchored DDPM (or DDIM) update fused into the epilogue (same arithmetic as dfb200_ddpm_step / dfb200_ddim_step; every
        // sample shares t); x, anchors, variances and the noise are already in registers
        const bool keep = P.traj != nullptr && t_cur > 0 && t_cur % P.traj_interval == 0;
        float* tr = keep ? P.traj + (size_t)(t_cur / P.traj_interval - 1) * (size_t)P.M * 3 : nullptr;
        float* ss = P.step_sample != nullptr ? P.step_sample + (size_t)step_local * (size_t)P.M * 3 : P.step_s[step_local & (TC_MAX_LIST_STEPS - 1)];
        float* sx = P.step_xstart != nullptr ? P.step_xstart + (size_t)step_local * (size_t)P.M * 3 : P.step_x[step_local & (TC_MAX_LIST_STEPS - 1)];
        float acp_s = 0.f, dir_c = 0.f;
        if (P.ddim_acp != nullptr) {
          acp_s = __fsqrt_rn(__ldg(P.ddim_acp + t_cur));
          dir_c = __ldg(P.ddim_dir + t_cur);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const long long e = (b * 3 + c) * P.N + p;
          const float x0 = ddpm_xstart(cf, xv[c], av[c], vv[c], eps3[c]);
          const float xp = P.ddim_acp != nullptr ? ddim_prev(cf, av[c], vv[c], x0, eps3[c], zq[c], acp_s, dir_c, P.ddim_eta)
                                                 : ddpm_prev(cf, xv[c], av[c], vv[c], x0, zq[c]);
          if (writer) {
            P.x_out[e] = xp;
            if (keep) tr[e] = xp;
            if (ss != nullptr) ss[e] = xp;
            if (sx != nullptr) sx[e] = x0;
          }
        }
      }
    }
    TL(0, 206);
    if (P.done != nullptr) {
      named_bar_sync(1 + T, 128);      // all 128 rows of the tile have stored x_{t-1} (CTA-scope order) ...
      if (r == 0 && (!P.guidance || T == 0))  // ... and ONE gpu-scope release publishes the tile-step (cumulative over the barrier)
        asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(P.done + unit), "r"(P.guidance ? 2 : 1) : "memory");
    }
    TL(0, 3 + P.depth * 40);
#ifdef DFB200_DIAGNOSTICS
    if (P.dbg != nullptr && tl_on && blockIdx.x < 148) P.dbg[230 + blockIdx.x] += clock64() - dbg_t0;  // per-CTA item cycles
    if (P.dbg != nullptr && blockIdx.x == 0 && item_n == P.dbg_item && tl_on) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      P.dbg[221] = (long long)ns;
    }
#endif
    }  // items
    tc_fence_before();
  } else if (warp == 8) {
    // =========================== MMA issuer ===========================
    // The whole warp runs this control flow (waits included); one elected lane issues tcgen05.mma / commit.
    // Everything that feeds a descriptor is warp-uniform by construction (constants, loop counters, the TMEM
    // base which is 0 for a 512-column allocation), so the issue sequence stays on the uniform datapath.
    constexpr uint32_t idesc128 = make_idesc_bf16(128, 128), idesc32 = make_idesc_bf16(128, 32), idesc16 = make_idesc_bf16(128, 16);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t ring = sbase + SM_RING, a_base = sbase + SM_A;
    const uint64_t ones_desc = make_smem_desc(sbase + SM_ONES, 2048, TILE_SBO);
    uint32_t ph_a0 = 0, ph_a1 = 0, ph_u0 = 0, ph_u1 = 0;
    const bool tl_on = lane == 0;
    auto pkt_addr = [&](int G) -> uint32_t {  // wait until packet G has landed; its smem address
      mbar_wait(&bars[BAR_WFULL + G % NSLOT], (uint32_t)(G / NSLOT) & 1u);
      return ring + (uint32_t)(G % NSLOT) * SLOT_BYTES;
    };
    auto wait_a = [&](int T) {
      if (T == 0) { mbar_wait(&bars[BAR_A + 0], ph_a0); ph_a0 ^= 1; }
      else { mbar_wait(&bars[BAR_A + 1], ph_a1); ph_a1 ^= 1; }
    };
    auto wait_u = [&](int T) {
      if (T == 0) { mbar_wait(&bars[BAR_UREADY + 0], ph_u0); ph_u0 ^= 1; }
      else { mbar_wait(&bars[BAR_UREADY + 1], ph_u1); ph_u1 ^= 1; }
    };
    // H_c(T) = LN3(x_T) W1'_c^T -> ACC_T (128 columns: 64 value | 64 gate); the bias b1'_c is added by the GEGLU epilogue
    auto ff_in = [&](int T, uint32_t pa, uint32_t pb) {
      const uint32_t d = 256 + T * 128, at = a_base + T * 32768;
      if (!ABL(8)) {
      umma_gemm<128, 4>(d, at, pa, idesc128, 0u);
      umma_gemm<128, 4>(d, at + 4 * 4096, pb, idesc128, 1u);
      }
      umma_commit(&bars[BAR_ACC + T]);
    };
    int item_n = 0;
#pragma unroll 1
    for (long long idx = blockIdx.x; idx < total_items; idx += gridDim.x, ++item_n) {
    // ---- x_T = pre_norm(proj_in(features)): one K=32 GEMM against the resident in-tile ----
#pragma unroll
    for (int T = 0; T < 2; ++T) {
      wait_a(T);
      tc_fence_after();
      if (elect_one()) {
        umma_gemm<128, 2>(T * 128, a_base + T * 32768, sbase + SM_INHEAD + IH_INTILE, idesc128, 0u);
        umma_commit(&bars[BAR_X + T]);
      }
      __syncwarp();
    }
    for (int l = 0; l < P.depth; ++l) {
      const int G0 = (item_n * P.depth + l) * PKT_PER_LAYER;
      TL(1, 2 + l * 40);
      // ---- logits: S_T = LN2(x_T) W_sim^T + b_sim -> ACC_T columns [0,32) ----
      uint32_t pf0 = 0, pf1 = 0;
#pragma unroll
      for (int T = 0; T < 2; ++T) {
        wait_a(T);
        const uint32_t pf = pkt_addr(G0 + T);
        if (T == 0) pf0 = pf; else pf1 = pf;
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = 256 + T * 128;
          umma_gemm<32, 8>(d, a_base + T * 32768, pf + FOLD_WSIM, idesc32, 0u);
          umma_bf16(d, ones_desc, make_smem_desc(pf + FOLD_BSIM, 0, TILE_SBO), idesc32, 1u);
          umma_commit(&bars[BAR_ACC + T]);
        }
        __syncwarp();
      }
      TL(1, 3 + l * 40);
      // ---- x_T += P_T W_pv^T + bo ----
      const uint32_t ps0 = pkt_addr(G0 + 2), ps1 = pkt_addr(G0 + 3);  // W1'_0 halves; ps1 also carries the bo slab
#pragma unroll
      for (int T = 0; T < 2; ++T) {
        wait_a(T);
        if (T == 0) TL(1, 4 + l * 40);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = T * 128, pf = T == 0 ? pf0 : pf1;
          umma_gemm<128, 2>(d, a_base + T * 32768, pf + FOLD_WPV, idesc128, 1u);
          umma_bf16(d, ones_desc, make_smem_desc(ps1 + SLAB_OFF, 0, TILE_SBO), idesc128, 1u);
          umma_commit(&bars[BAR_X + T]);
        }
        __syncwarp();
      }
      if (elect_one()) { umma_commit(&bars[BAR_WEMPTY + (G0 + 0) % NSLOT]); umma_commit(&bars[BAR_WEMPTY + (G0 + 1) % NSLOT]); }
      __syncwarp();
      TL(1, 5 + l * 40);
      // ---- feed-forward, the two tiles ping-pong ----
#pragma unroll
      for (int T = 0; T < 2; ++T) {
        wait_a(T);
        tc_fence_after();
        if (elect_one()) ff_in(T, ps0, ps1);
        __syncwarp();
      }
      if (elect_one()) { umma_commit(&bars[BAR_WEMPTY + (G0 + 2) % NSLOT]); umma_commit(&bars[BAR_WEMPTY + (G0 + 3) % NSLOT]); }
      __syncwarp();
      TL(1, 7 + l * 40);
#pragma unroll 1
      for (int c = 0; c < FF_CHUNKS; ++c) {
        const bool last = c == FF_CHUNKS - 1;
        const int gw2 = G0 + 2 + spkt_w2(c);
        const uint32_t pw2 = pkt_addr(gw2);
        uint32_t pa = 0, pb = 0;
        if (!last) { pa = pkt_addr(G0 + 2 + spkt_w1a(c + 1)); pb = pkt_addr(G0 + 2 + spkt_w1b(c + 1)); }
#pragma unroll
        for (int T = 0; T < 2; ++T) {
          if (T == 0) TL(1, 8 + l * 40 + c * 2);
          wait_u(T);
          if (T == 0) TL(1, 9 + l * 40 + c * 2);
#ifdef DFB200_DIAGNOSTICS
          if (ABL(32) || ABL(64)) {  // idle cycles on the critical chain (no work, ~no energy): is the sustained rate cycle- or power-bound?
            const long long t_end = clock64() + (ABL(64) ? 300 : 150);
            while (clock64() < t_end) __nanosleep(20);
          }
#endif
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d = T * 128;
            // x_T += U_T W2_c^T with U read from TMEM (8 columns per K=16 step; k [0,32) at columns [0,16), k [32,64) at [32,48))
            if (!ABL(16))
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              umma_bf16_ts(d, 256 + T * 128 + (ks >> 1) * 32 + (ks & 1) * 8, make_smem_desc(pw2 + ks * 4096, 2048, TILE_SBO), idesc128, 1u);
            if (last) {
              umma_bf16(d, ones_desc, make_smem_desc(pw2 + SLAB_OFF, 0, TILE_SBO), idesc128, 1u);
              umma_commit(&bars[BAR_X + T]);
            } else {
              ff_in(T, pa, pb);
            }
          }
          __syncwarp();
        }
        if (elect_one()) {
          umma_commit(&bars[BAR_WEMPTY + gw2 % NSLOT]);
          if (!last) {
            umma_commit(&bars[BAR_WEMPTY + (G0 + 2 + spkt_w1a(c + 1)) % NSLOT]);
            umma_commit(&bars[BAR_WEMPTY + (G0 + 2 + spkt_w1b(c + 1)) % NSLOT]);
          }
        }
        __syncwarp();
      }
    }
    // ---- eps = proj_out(post_norm(x_T)): N=16 GEMM against the resident head tile -> ACC_T columns [0,16) ----
#pragma unroll
    for (int T = 0; T < 2; ++T) {
      wait_a(T);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d = 256 + T * 128;
        umma_gemm<16, 8>(d, a_base + T * 32768, sbase + SM_INHEAD + IH_HEADTILE, idesc16, 0u);
        umma_bf16(d, ones_desc, make_smem_desc(sbase + SM_INHEAD + IH_HEADSLAB, 0, TILE_SBO), idesc16, 1u);
        umma_commit(&bars[BAR_ACC + T]);
      }
      __syncwarp();
    }
    }  // items
    tc_fence_before();
  } else {
    // =========================== weight producer ===========================
    int G = 0;
#pragma unroll 1
    for (long long idx = blockIdx.x; idx < total_items; idx += gridDim.x) {
      const int unit = (int)(idx % P.n_units), step_local = (int)(idx / P.n_units);
      const long long tile_id0 = P.guidance ? (long long)unit : (long long)unit * 2, tile_id1 = P.guidance ? (long long)unit : tile_id0 + 1;
      const long long t0 = tile_id0 < n_tiles ? tile_id0 : n_tiles - 1, t1 = tile_id1 < n_tiles ? tile_id1 : n_tiles - 1;
      const long long b0 = t0 * 128 / P.N, b1 = t1 * 128 / P.N + (P.guidance ? P.B : 0);
      const uint8_t* fold = P.fold + (size_t)step_local * P.fold_step_bytes;
      for (int lp = 0; lp < P.depth * PKT_PER_LAYER; ++lp, ++G) {
        const int slot = G % NSLOT;
        mbar_wait(&bars[BAR_WEMPTY + slot], ((uint32_t)(G / NSLOT) & 1u) ^ 1u);
        if (elect_one()) {
          const int l = lp / PKT_PER_LAYER, p = lp - l * PKT_PER_LAYER;
          const uint32_t bytes = (uint32_t)pkt_bytes(p);
          const uint8_t* src = p < 2 ? fold + ((size_t)(p == 0 ? b0 : b1) * P.depth + l) * FOLD_BYTES
                                     : P.stream + ((size_t)l * STATIC_PER_LAYER + (p - 2)) * SLOT_BYTES;
          if (ABL(4) && p >= 2 && G >= NSLOT) {
            mbar_arrive(&bars[BAR_WFULL + slot]);
          } else {
          mbar_arrive_expect_tx(&bars[BAR_WFULL + slot], bytes);
          bulk_g2s(smem + SM_RING + slot * SLOT_BYTES, src, bytes, &bars[BAR_WFULL + slot]);
          }
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

#ifdef DFB200_DIAGNOSTICS
static long long* g_tc_timeline = nullptr;  // set by dfb200_debug_tc_timeline
static int g_tc_timeline_item = 0;
#endif

static const float* tc_extras(const PackLayout& L, const void* packed) {
  const uint8_t* S = reinterpret_cast<const uint8_t*>(packed) + L.tc_stream_off;
  return reinterpret_cast<const float*>(S + (size_t)L.d.depth * STATIC_PER_LAYER * SLOT_BYTES);
}

// shared with the tf32 kernel (denoiser_tf32.cu): the per-block fold weights (WqG | bqG | WoT) and the folded GEGLU-in biases
const float* tc_foldw_base(const PackLayout& L, const void* packed, size_t* stride_floats) {
  *stride_floats = FOLDW_FLOATS;
  return tc_extras(L, packed) + HEAD_FLOATS;
}
const float* tc_b1p_base(const PackLayout& L, const void* packed) {
  return tc_extras(L, packed) + HEAD_FLOATS + (size_t)L.d.depth * FOLDW_FLOATS;
}

int launch_context_fold(const PackLayout& L, const void* packed, int B, const float* kv_static, const float* kv_time, int t_first,
                        const int* step_t, int steps, void* fold, cudaStream_t st) {
  if (B == 0 || steps == 0) return DFB200_OK;
  context_fold_kernel<<<dim3(B, L.d.depth, cdiv(steps, FOLD_SPB)), 512, 0, st>>>(L.d.depth, kv_static, kv_time, t_first, step_t, steps,
                                                                                  tc_extras(L, packed), reinterpret_cast<uint8_t*>(fold));
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

int denoiser_step_tc(const PackLayout& L, const void* packed, int B, int N, const float* x, const float* anchors,
                     const float* variances, const int* assign, const float* valid_id, const void* fold, float* eps_out,
                     const TcUpdate* upd, cudaStream_t st) {
  DFB_REQUIRE(N % 128 == 0, DFB200_ERR_UNSUPPORTED, "denoiser (bf16 mode): N must be a multiple of 128 (got %d); use fp32 mode", N);
  static DeviceOnce attr_once;  // cudaFuncSetAttribute is per device
  if (attr_once.first_time())
    DFB_CUDA(cudaFuncSetAttribute(denoiser_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
  TcParams p{};
  p.stream = reinterpret_cast<const uint8_t*>(packed) + L.tc_stream_off;
  p.fold = reinterpret_cast<const uint8_t*>(fold);
  p.inhead = reinterpret_cast<const uint8_t*>(tc_extras(L, packed));
  p.b1p = tc_extras(L, packed) + HEAD_FLOATS + (size_t)L.d.depth * FOLDW_FLOATS;
  p.x = x; p.anchors = anchors; p.variances = variances; p.assign = assign; p.valid = valid_id;
  p.eps_out = eps_out;
  p.N = N; p.depth = L.d.depth; p.flags = L.d.flags;
  p.M = (long long)B * N;
#ifdef DFB200_DIAGNOSTICS
  p.dbg = g_tc_timeline;
  p.dbg_item = g_tc_timeline_item;
  { const char* e = getenv("DFB200_TC_ABLATE"); p.abl = e != nullptr ? atoi(e) : 0; }
#endif
  p.B = B;
  p.guidance = upd != nullptr && upd->guidance;
  p.guid_w = upd != nullptr ? upd->guid_w : 0.f;
  p.n_units = p.guidance ? (int)(p.M / 128) : cdiv(p.M, 256);
  p.n_steps = 1;
  p.t_first = 0;
  if (upd != nullptr) {
    p.upd_sched = upd->sched; p.upd_T = upd->T; p.t_first = upd->t; p.upd_noise = upd->noise; p.upd_seed = upd->seed; p.x_out = upd->x_out;
    p.n_steps = upd->n_steps; p.fold_step_bytes = upd->fold_step_bytes; p.noise_step_elems = (size_t)p.M * 3;
    p.done = upd->done; p.traj = upd->traj; p.traj_interval = upd->traj_interval;
    p.step_t = upd->step_t; p.step_base = upd->step_base;
    p.step_sample = upd->step_sample; p.step_xstart = upd->step_xstart;
    if (upd->step_sample_list != nullptr || upd->step_xstart_list != nullptr) {
      DFB_REQUIRE(p.n_steps <= TC_MAX_LIST_STEPS, DFB200_ERR_INVALID_ARG, "denoiser (bf16 mode): at most %d steps per launch with per-step output buffers",
                  TC_MAX_LIST_STEPS);
      for (int k = 0; k < p.n_steps; ++k) {
        if (upd->step_sample_list != nullptr) p.step_s[k] = upd->step_sample_list[k];
        if (upd->step_xstart_list != nullptr) p.step_x[k] = upd->step_xstart_list[k];
      }
    }
    p.ddim_acp = upd->ddim_acp; p.ddim_dir = upd->ddim_dir; p.ddim_eta = upd->ddim_eta;
    DFB_REQUIRE(p.n_steps == 1 || p.done != nullptr, DFB200_ERR_INVALID_ARG, "denoiser (bf16 mode): multi-step launch without dependency counters");
  }
  // the spin-wait on P.done assumes every CTA of the grid is resident: size it by the LAUNCHING device's SM count
  const int n_sm = current_device_sm_count();
  DFB_REQUIRE(n_sm > 0, DFB200_ERR_CUDA, "denoiser (bf16 mode): cannot query the SM count of the current device");
  // persistent: one CTA per SM walks the (step, unit) work list; dependencies between steps of a unit go through P.done
  const long long items = (long long)p.n_units * p.n_steps;
  const int grid = (int)(items < n_sm ? items : n_sm);
  denoiser_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(p);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

int denoiser_forward_tc(const PackLayout& L, const void* packed, int B, int N, const float* x, const float* anchors,
                        const float* variances, const int* assign, const float* valid_id, float* eps_out, Workspace& ws,
                        cudaStream_t st) {
  // K/V (ws.kv, from launch_context_kv) -> per-(sample, block) folded attention tiles -> fused kernel
  int rc = launch_context_fold(L, packed, B, ws.kv, nullptr, 0, nullptr, 1, ws.fold, st);
  if (rc != DFB200_OK) return rc;
  return denoiser_step_tc(L, packed, B, N, x, anchors, variances, assign, valid_id, ws.fold, eps_out, nullptr, st);
}

}  // namespace dfb200

using namespace dfb200;

#ifdef DFB200_DIAGNOSTICS
// Debug hook (diagnostic build only): device buffer of 1024 int64 that CTA 0 of the fused kernel fills with clock64()
// stamps (NULL = off).  Declared in include/difffacto_b200_diag.h.
extern "C" int dfb200_debug_tc_timeline(long long* device_buffer, int item) {
  g_tc_timeline = device_buffer;
  g_tc_timeline_item = item;
  return DFB200_OK;
}
#endif
