// PointNet++ set-abstraction ops for sm_100a behind the C ABI of include/difffacto_b200.h.
//
// What each kernel must reproduce (reference: pointnet2_ops_lib/pointnet2_ops/_ext-src/src/):
//   sampling_gpu.cu:8-20,34-47   gather_points (+grad)
//   sampling_gpu.cu:59-173       furthest_point_sampling  (tie rule of the smem tree, |p|^2<=1e-3 skip)
//   ball_query_gpu.cu:9-44       query_ball_point (first nsample hits in index order, first-hit padding)
//   group_points_gpu.cu:8-28,43-64   group_points (+grad)
//   interpolate_gpu.cu:9-59,72-101,116-143   three_nn, three_interpolate (+grad)
// The reference runs ONE block per cloud for everything except gather; here every op is tiled
// over (cloud, centre/point tile, channel tile) so that a batch fills the 148 SMs, xyz tiles are
// staged in shared memory, and idx/out traffic is coalesced and 128-bit vectorised.
#include <math.h>

#include "common.cuh"

namespace dfb200 {

// ============================================================================================
// gather_points / group_points : out[b,c,s] = points[b,c,idx[b,s]],  s over npoints*nsample
// ============================================================================================
// HBM-bound, write-dominated.  Each thread owns 4 consecutive s (one int4 idx load, reused from
// registers for every channel of its channel tile; one streaming float4 store per channel); the
// gathered `points` row (n floats) stays in L1/L2.
template <bool VEC4>
__global__ void __launch_bounds__(256)
group_points_kernel(int c, int n, long long S, int c_per_block, const float* __restrict__ points,
                    const int* __restrict__ idx, float* __restrict__ out) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * c_per_block;
  const int c1 = min(c, c0 + c_per_block);
  const float* pts = points + (size_t)b * c * n;
  const int* id = idx + (size_t)b * S;
  float* o = out + (size_t)b * c * S;
  if (VEC4) {
    const long long s4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (s4 >= S) return;
    const int4 ii = __ldg(reinterpret_cast<const int4*>(id + s4));
#pragma unroll 4
    for (int l = c0; l < c1; ++l) {
      const float* row = pts + (size_t)l * n;
      float4 v;
      v.x = __ldg(row + ii.x);
      v.y = __ldg(row + ii.y);
      v.z = __ldg(row + ii.z);
      v.w = __ldg(row + ii.w);
      __stcs(reinterpret_cast<float4*>(o + (size_t)l * S + s4), v);
    }
  } else {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int ii = __ldg(id + s);
#pragma unroll 4
    for (int l = c0; l < c1; ++l) o[(size_t)l * S + s] = __ldg(pts + (size_t)l * n + ii);
  }
}

// Shared-memory staged variant for large S (grouping: npoints*nsample outputs per source row).
// A CTA owns a tile of G4_S consecutive outputs and a tile of channel QUADS.  The idx tile is read ONCE into registers
// (4 x int4 per thread).  For every quad of channels the source rows are staged in shared memory INTERLEAVED, one
// float4 = (ch0, ch1, ch2, ch3) per point, with coalesced row reads and conflict-free STS.128; the gather is then one
// random LDS.128 per output position that serves four channels (a quarter-warp phase moves 8 x 16 B, so the expected
// bank-conflict cost per gathered byte is a third of a 4-byte gather's), a 4x4 register transpose, and one streaming
// STG.128 per channel.  The random reads never leave the SM; HBM sees only the streaming idx read and output write, L2
// one extra read of the (L2-resident) source per G4_S outputs.
constexpr int G4_THREADS = 512;
constexpr int G4_VEC = 4;                               // int4 idx vectors (16 outputs) per thread
constexpr int G4_S = G4_THREADS * G4_VEC * 4;           // 8192 outputs per CTA
__global__ void __launch_bounds__(G4_THREADS)
group_points_c4_kernel(int c, int n, long long S, int quads_per_block, const float* __restrict__ points,
                       const int* __restrict__ idx, float* __restrict__ out) {
  extern __shared__ float4 g4_tile[];  // [n]
  const int b = blockIdx.z, tid = threadIdx.x;
  const int nquad = (c + 3) >> 2;
  const int q0 = blockIdx.y * quads_per_block, q1 = min(nquad, q0 + quads_per_block);
  const float* pts = points + (size_t)b * c * n;
  const int* id = idx + (size_t)b * S;
  float* o = out + (size_t)b * c * S;
  const long long sbase = (long long)blockIdx.x * G4_S + tid * 4;
  int4 ii[G4_VEC];
  bool ok[G4_VEC];
#pragma unroll
  for (int j = 0; j < G4_VEC; ++j) {
    const long long s = sbase + (long long)j * G4_THREADS * 4;
    ok[j] = s < S;
    ii[j] = ok[j] ? __ldg(reinterpret_cast<const int4*>(id + s)) : make_int4(0, 0, 0, 0);
  }
  for (int q = q0; q < q1; ++q) {
    const int l0 = q * 4, nl = min(4, c - l0);
    const float* r0 = pts + (size_t)l0 * n;
    const float* r1 = r0 + (nl > 1 ? n : 0);
    const float* r2 = r0 + (nl > 2 ? 2 * (size_t)n : 0);
    const float* r3 = r0 + (nl > 3 ? 3 * (size_t)n : 0);
    __syncthreads();  // the previous quad's gathers are done
    for (int i = tid; i < n; i += G4_THREADS) g4_tile[i] = make_float4(__ldg(r0 + i), __ldg(r1 + i), __ldg(r2 + i), __ldg(r3 + i));
    __syncthreads();
    float* o0 = o + (size_t)l0 * S;
#pragma unroll
    for (int j = 0; j < G4_VEC; ++j) {
      if (!ok[j]) continue;
      const long long s = sbase + (long long)j * G4_THREADS * 4;
      const float4 a = g4_tile[ii[j].x], bq = g4_tile[ii[j].y], cq = g4_tile[ii[j].z], d = g4_tile[ii[j].w];
      __stcs(reinterpret_cast<float4*>(o0 + s), make_float4(a.x, bq.x, cq.x, d.x));
      if (nl > 1) __stcs(reinterpret_cast<float4*>(o0 + (size_t)S + s), make_float4(a.y, bq.y, cq.y, d.y));
      if (nl > 2) __stcs(reinterpret_cast<float4*>(o0 + 2 * (size_t)S + s), make_float4(a.z, bq.z, cq.z, d.z));
      if (nl > 3) __stcs(reinterpret_cast<float4*>(o0 + 3 * (size_t)S + s), make_float4(a.w, bq.w, cq.w, d.w));
    }
  }
}

static int launch_group(int b, int c, int n, long long S, const float* points, const int* idx,
                        float* out, cudaStream_t st) {
  if (b == 0 || c == 0 || S == 0) return DFB200_OK;
  {
    // staged path: 16 B-aligned vectorisable outputs, an interleaved source tile that fits in shared memory, enough reuse of it
    const bool vec_ok = (S % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const size_t smem = sizeof(float4) * (size_t)n;
    if (vec_ok && S >= 4 * (long long)n && S >= 4096 && smem <= 96 * 1024 && b <= 65535) {
      const int gx = cdiv(S, G4_S);
      const int nquad = (c + 3) / 4;
      int qpb = nquad;
      while (qpb > 1 && (long long)gx * b * cdiv(nquad, qpb) < 148 * 4) qpb = (qpb + 1) / 2;
      dim3 grid(gx, cdiv(nquad, qpb), b);
      if (grid.y <= 65535) {
        static size_t smem_set = 48 * 1024;
        if (smem > smem_set) {
          DFB_CUDA(cudaFuncSetAttribute(group_points_c4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
          smem_set = 96 * 1024;
        }
        group_points_c4_kernel<<<grid, G4_THREADS, smem, st>>>(c, n, S, qpb, points, idx, out);
        DFB_LAUNCH_CHECK();
        return DFB200_OK;
      }
    }
  }
  const bool vec = (S % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const long long work = vec ? S / 4 : S;
  const int gx = cdiv(work, 256);
  // channel tile: enough blocks to cover the machine several times, while amortising the idx load
  int c_per_block = c;
  while (c_per_block > 4 && (long long)gx * b * cdiv(c, c_per_block) < 148 * 8) c_per_block = (c_per_block + 1) / 2;
  dim3 grid(gx, cdiv(c, c_per_block), b);
  DFB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, DFB200_ERR_INVALID_ARG, "group_points: grid too large");
  if (vec)
    group_points_kernel<true><<<grid, 256, 0, st>>>(c, n, S, c_per_block, points, idx, out);
  else
    group_points_kernel<false><<<grid, 256, 0, st>>>(c, n, S, c_per_block, points, idx, out);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

// grad_points[b,c,idx[b,s]] += grad_out[b,c,s]   (red.global.add.f32; order nondeterministic as
// in the reference's atomicAdd scatter)
__global__ void __launch_bounds__(256)
group_points_grad_kernel(int c, int n, long long S, int c_per_block,
                         const float* __restrict__ grad_out, const int* __restrict__ idx,
                         float* __restrict__ grad_points) {
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * c_per_block;
  const int c1 = min(c, c0 + c_per_block);
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  const int ii = __ldg(idx + (size_t)b * S + s);
  const float* go = grad_out + (size_t)b * c * S;
  float* gp = grad_points + (size_t)b * c * n;
  for (int l = c0; l < c1; ++l) atomicAdd(gp + (size_t)l * n + ii, __ldg(go + (size_t)l * S + s));
}

static int launch_group_grad(int b, int c, int n, long long S, const float* grad_out,
                             const int* idx, float* grad_points, cudaStream_t st) {
  DFB_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * c * n, st));
  if (b == 0 || c == 0 || S == 0) return DFB200_OK;
  const int gx = cdiv(S, 256);
  int c_per_block = c;
  while (c_per_block > 4 && (long long)gx * b * cdiv(c, c_per_block) < 148 * 8) c_per_block = (c_per_block + 1) / 2;
  dim3 grid(gx, cdiv(c, c_per_block), b);
  DFB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, DFB200_ERR_INVALID_ARG, "group_points_grad: grid too large");
  group_points_grad_kernel<<<grid, 256, 0, st>>>(c, n, S, c_per_block, grad_out, idx, grad_points);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

// ============================================================================================
// furthest point sampling
// ============================================================================================
// One CTA per cloud (the m-1 rounds are inherently sequential).  Unlike the reference, the
// running min-distance array lives in REGISTERS (PPT points per thread), the cloud lives in
// shared memory (the "last selected point" fetch is an LDS, not a global load), and the
// per-round arg-max is two REDUX instructions per warp + ONE __syncthreads (double-buffered
// cross-warp slots) instead of a 9-level shared-memory tree with 10 barriers.
//
// Tie rule.  The reference picks, among equal maxima, the winner of its smem tree: thread-local
// strict `>` keeps the lowest k of a thread (k = tid, tid+bs, ...), and __update() keeps the
// LOWER slot on ties at strides bs/2, bs/4, ..., 1.  The slot surviving at stride s holds tids
// congruent mod 2s, so ties are resolved on bit0 of tid first, then bit1, ...: the winner has
// the smallest BIT-REVERSED tid (log2(bs) bits), then the smallest k.  We encode that as
//   rank(k) = (bitrev(k mod bs) << qbits) | (k / bs)            (unique per point)
// and reduce (value desc, rank asc) -- a total order, so any reduction tree gives the reference's
// winner and the point->thread mapping is free.
__device__ __forceinline__ unsigned fps_rank(int k, int log2bs, int qbits) {
  const unsigned v = (unsigned)k & ((1u << log2bs) - 1u);
  const unsigned q = (unsigned)k >> log2bs;
  const unsigned rv = log2bs ? (__brev(v) >> (32 - log2bs)) : 0u;
  return (rv << qbits) | q;
}
__device__ __forceinline__ int fps_unrank(unsigned rank, int log2bs, int qbits) {
  const unsigned q = rank & ((1u << qbits) - 1u);
  const unsigned rv = rank >> qbits;
  const unsigned v = log2bs ? (__brev(rv) >> (32 - log2bs)) : 0u;
  return (int)(v + (q << log2bs));
}

template <int PPT, int THREADS, bool REGPTS>
__global__ void __launch_bounds__(THREADS)
fps_kernel(int n, int m, int log2bs, int qbits, const float* __restrict__ dataset,
           float* __restrict__ temp, int* __restrict__ idxs) {
  extern __shared__ float fps_smem[];  // xs[n], ys[n], zs[n]
  constexpr int NW = THREADS / 32;
  constexpr int NR = REGPTS ? PPT : 1;
  __shared__ uint2 slot[2][NW];
  float* xs = fps_smem;
  float* ys = xs + n;
  float* zs = ys + n;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* pts = dataset + (size_t)b * n * 3;
  int* out = idxs + (size_t)b * m;

  for (int i = tid; i < n * 3; i += THREADS) {  // coalesced AoS read -> SoA smem
    const float v = __ldg(pts + i);
    const int k = i / 3, ch = i - k * 3;
    (ch == 0 ? xs : ch == 1 ? ys : zs)[k] = v;
  }
  __syncthreads();

  // REGPTS: the thread's points are cached in registers; otherwise (large n, 64-register budget
  // at 1024 threads) they are re-read from shared memory every round.
  float px[NR], py[NR], pz[NR], td[PPT];
  unsigned okmask = 0u;
#pragma unroll
  for (int i = 0; i < PPT; ++i) {
    const int k = tid + i * THREADS;
    const bool in = k < n;
    const float x = in ? xs[k] : 0.f, y = in ? ys[k] : 0.f, z = in ? zs[k] : 0.f;
    if (REGPTS) { px[i] = x; py[i] = y; pz[i] = z; }
    td[i] = 1e10f;
    // reference: `float mag = x*x+y*y+z*z; if (mag <= 1e-3) continue;` -- fp32 mag, DOUBLE compare
    const float mag = sq3(x, y, z);
    if (in && !((double)mag <= 1e-3)) okmask |= 1u << i;
  }

  int old = 0;
  if (tid == 0 && m > 0) out[0] = 0;
  for (int j = 1; j < m; ++j) {
    const float x1 = xs[old], y1 = ys[old], z1 = zs[old];
    float best = -1.f;
    int bestk = 0;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      if ((okmask >> i) & 1u) {
        const int k = tid + i * THREADS;
        const float x = REGPTS ? px[i] : xs[k], y = REGPTS ? py[i] : ys[k], z = REGPTS ? pz[i] : zs[k];
        const float d = sq3(x - x1, y - y1, z - z1);
        const float d2 = fminf(d, td[i]);
        td[i] = d2;
        bool take = d2 > best;
        if (d2 == best) take = fps_rank(k, log2bs, qbits) < fps_rank(bestk, log2bs, qbits);
        best = take ? d2 : best;
        bestk = take ? k : bestk;
      }
    }
    // d2 >= 0 (or best == -1: nothing valid) -> order-preserving unsigned key, 0 = "nothing"
    const unsigned key = best < 0.f ? 0u : (__float_as_uint(best) + 1u);
    const unsigned brank = best < 0.f ? 0xFFFFFFFFu : fps_rank(bestk, log2bs, qbits);
    unsigned wkey = __reduce_max_sync(0xFFFFFFFFu, key);
    unsigned wrank = __reduce_min_sync(0xFFFFFFFFu, key == wkey ? brank : 0xFFFFFFFFu);
    if (NW > 1) {
      if (lane == 0) slot[j & 1][warp] = make_uint2(wkey, wrank);
      __syncthreads();
      const uint2 s = lane < NW ? slot[j & 1][lane] : make_uint2(0u, 0xFFFFFFFFu);
      wkey = __reduce_max_sync(0xFFFFFFFFu, s.x);
      wrank = __reduce_min_sync(0xFFFFFFFFu, s.x == wkey ? s.y : 0xFFFFFFFFu);
    }
    old = (wkey == 0u) ? 0 : fps_unrank(wrank, log2bs, qbits);
    if (tid == 0) out[j] = old;
  }
  if (temp != nullptr) {
    float* t = temp + (size_t)b * n;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const int k = tid + i * THREADS;
      if (k < n) t[k] = td[i];
    }
  }
}

// reference cuda_utils.h:15-19 -- evaluated in double exactly as there (the quotient of logs can
// land just below an integer, which changes the block size and therefore the tie rule).
static int ref_opt_n_threads(int work_size) {
  const int pow_2 = (int)(log((double)work_size) / log(2.0));
  int v = 1 << pow_2;
  if (v > 512) v = 512;
  if (v < 1) v = 1;
  return v;
}

template <int PPT, int THREADS, bool REGPTS>
static int launch_fps(int b, int n, int m, int log2bs, int qbits, const float* dataset, float* temp,
                      int* idxs, cudaStream_t st) {
  const size_t smem = sizeof(float) * 3 * (size_t)n;
  if (smem > 48 * 1024) {
    DFB_CUDA(cudaFuncSetAttribute(fps_kernel<PPT, THREADS, REGPTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  fps_kernel<PPT, THREADS, REGPTS><<<b, THREADS, smem, st>>>(n, m, log2bs, qbits, dataset, temp, idxs);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

// ============================================================================================
// ball query
// ============================================================================================
// One WARP per query centre: 32 candidate points per step tested in parallel, hits compacted in
// index order with ballot + prefix-popcount (keeps "first nsample in ascending k"), exact early
// exit once nsample hits are found, coalesced idx writes.  The cloud tile is staged once per CTA
// in shared memory (SoA) and reused by CENTRES_PER_WARP * 8 centres.
template <bool SMEM>
__global__ void __launch_bounds__(256)
ball_query_kernel(int n, int m, float radius2, int nsample, int centres_per_warp,
                  const float* __restrict__ new_xyz, const float* __restrict__ xyz,
                  int* __restrict__ idx) {
  extern __shared__ float bq_smem[];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* pts = xyz + (size_t)b * n * 3;
  float* xs = bq_smem;
  float* ys = xs + n;
  float* zs = ys + n;
  if (SMEM) {
    for (int i = threadIdx.x; i < n * 3; i += blockDim.x) {
      const float v = __ldg(pts + i);
      const int k = i / 3, ch = i - k * 3;
      (ch == 0 ? xs : ch == 1 ? ys : zs)[k] = v;
    }
    __syncthreads();
  }
  const int j0 = (blockIdx.x * 8 + warp) * centres_per_warp;
  for (int jj = 0; jj < centres_per_warp; ++jj) {
    const int j = j0 + jj;
    if (j >= m) break;
    const float* cq = new_xyz + ((size_t)b * m + j) * 3;
    const float cx = __ldg(cq), cy = __ldg(cq + 1), cz = __ldg(cq + 2);
    int* o = idx + ((size_t)b * m + j) * nsample;
    int cnt = 0, first = 0;
    for (int k0 = 0; k0 < n && cnt < nsample; k0 += 128) {
      unsigned mask[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u * 32 + lane;
        bool hit = false;
        if (k < n) {
          float x, y, z;
          if (SMEM) {
            x = xs[k]; y = ys[k]; z = zs[k];
          } else {
            x = __ldg(pts + 3 * k); y = __ldg(pts + 3 * k + 1); z = __ldg(pts + 3 * k + 2);
          }
          const float d2 = sq3(cx - x, cy - y, cz - z);
          hit = d2 < radius2;  // reference: `if (d2 < radius2)` (PTX setp.geu + branch)
        }
        mask[u] = __ballot_sync(0xFFFFFFFFu, hit);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (mask[u] != 0u && cnt < nsample) {
          if (cnt == 0) first = k0 + u * 32 + __ffs(mask[u]) - 1;
          const int pos = cnt + __popc(mask[u] & lanemask_lt());
          if (((mask[u] >> lane) & 1u) && pos < nsample) o[pos] = k0 + u * 32 + lane;
          cnt += __popc(mask[u]);
        }
      }
    }
    // slots never reached keep the first hit (or 0 for an empty ball)
    for (int l = min(cnt, nsample) + lane; l < nsample; l += 32) o[l] = first;
  }
}

// ============================================================================================
// three_nn / three_interpolate
// ============================================================================================
// Thread per unknown point, `known` staged in smem tiles (broadcast LDS).  The reference keeps its
// three bests in double initialised to 1e40 but compares against an fp32 distance; fp32 bests
// initialised to +inf select the same indices and produce the same fp32 outputs (float(1e40) is
// +inf), without touching the FP64 pipe.
constexpr int NN_TILE = 2048;
__global__ void __launch_bounds__(256)
three_nn_kernel(int n, int m, const float* __restrict__ unknown, const float* __restrict__ known,
                float* __restrict__ dist2, int* __restrict__ idx) {
  __shared__ float kx[NN_TILE], ky[NN_TILE], kz[NN_TILE];
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const float* kn = known + (size_t)b * m * 3;
  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (j < n) {
    const float* u = unknown + ((size_t)b * n + j) * 3;
    ux = __ldg(u); uy = __ldg(u + 1); uz = __ldg(u + 2);
  }
  float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
  int i1 = 0, i2 = 0, i3 = 0;
  for (int k0 = 0; k0 < m; k0 += NN_TILE) {
    const int tile = min(NN_TILE, m - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < tile * 3; i += blockDim.x) {
      const float v = __ldg(kn + (size_t)k0 * 3 + i);
      const int k = i / 3, ch = i - k * 3;
      (ch == 0 ? kx : ch == 1 ? ky : kz)[k] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < tile; ++k) {
      const float d = sq3(ux - kx[k], uy - ky[k], uz - kz[k]);
      if (d < b1) {
        b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k0 + k;
      } else if (d < b2) {
        b3 = b2; i3 = i2; b2 = d; i2 = k0 + k;
      } else if (d < b3) {
        b3 = d; i3 = k0 + k;
      }
    }
  }
  if (j < n) {
    float* d = dist2 + ((size_t)b * n + j) * 3;
    int* o = idx + ((size_t)b * n + j) * 3;
    d[0] = b1; d[1] = b2; d[2] = b3;
    o[0] = i1; o[1] = i2; o[2] = i3;
  }
}

// out[b,l,j] = p1*w1 + p2*w2 + p3*w3 rounded as the reference's SASS does:
// fma(p3,w3, fma(p1,w1, mul(p2,w2))).  idx/weight of a point are loaded once and reused for the
// whole channel tile; writes are coalesced along j.
__global__ void __launch_bounds__(256)
three_interpolate_kernel(int c, int m, int n, int c_per_block, const float* __restrict__ points,
                         const int* __restrict__ idx, const float* __restrict__ weight,
                         float* __restrict__ out) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int c0 = blockIdx.y * c_per_block, c1 = min(c, c0 + c_per_block);
  const int* ii = idx + ((size_t)b * n + j) * 3;
  const float* ww = weight + ((size_t)b * n + j) * 3;
  const int i1 = __ldg(ii), i2 = __ldg(ii + 1), i3 = __ldg(ii + 2);
  const float w1 = __ldg(ww), w2 = __ldg(ww + 1), w3 = __ldg(ww + 2);
  const float* pts = points + (size_t)b * c * m;
  float* o = out + (size_t)b * c * n;
#pragma unroll 4
  for (int l = c0; l < c1; ++l) {
    const float* row = pts + (size_t)l * m;
    const float v = __fmaf_rn(__ldg(row + i3), w3, __fmaf_rn(__ldg(row + i1), w1, __fmul_rn(__ldg(row + i2), w2)));
    __stcs(o + (size_t)l * n + j, v);
  }
}

// Shared-memory staged variant (same idea as group_points_c4_kernel): per quad of channels the m source columns are staged
// interleaved in shared memory (one float4 = 4 channels per source point), so each of the 3 neighbours costs ONE random
// LDS.128 for 4 channels and the gathers never leave the SM; a thread owns TI_PPT query points (idx / weights in registers
// for the whole channel tile) and the 4 output rows are written with coalesced streaming stores.  (A variant with 4
// consecutive points per thread and STG.128 stores measured slower: its idx/weight loads are 48-byte strided.)
constexpr int TI_THREADS = 256, TI_PPT = 4;
__global__ void __launch_bounds__(TI_THREADS)
three_interpolate_c4_kernel(int c, int m, int n, int quads_per_block, const float* __restrict__ points,
                            const int* __restrict__ idx, const float* __restrict__ weight, float* __restrict__ out) {
  extern __shared__ float4 ti_tile[];  // [m]
  const int b = blockIdx.z, tid = threadIdx.x;
  const int nquad = (c + 3) >> 2;
  const int q0 = blockIdx.y * quads_per_block, q1 = min(nquad, q0 + quads_per_block);
  const float* pts = points + (size_t)b * c * m;
  float* o = out + (size_t)b * c * n;
  int i1[TI_PPT], i2[TI_PPT], i3[TI_PPT];
  float w1[TI_PPT], w2[TI_PPT], w3[TI_PPT];
  bool ok[TI_PPT];
#pragma unroll
  for (int u = 0; u < TI_PPT; ++u) {
    const int j = (blockIdx.x * TI_PPT + u) * TI_THREADS + tid;  // consecutive threads -> consecutive points (coalesced stores)
    ok[u] = j < n;
    const int* ii = idx + ((size_t)b * n + (ok[u] ? j : 0)) * 3;
    const float* ww = weight + ((size_t)b * n + (ok[u] ? j : 0)) * 3;
    i1[u] = __ldg(ii); i2[u] = __ldg(ii + 1); i3[u] = __ldg(ii + 2);
    w1[u] = __ldg(ww); w2[u] = __ldg(ww + 1); w3[u] = __ldg(ww + 2);
  }
  for (int q = q0; q < q1; ++q) {
    const int l0 = q * 4, nl = min(4, c - l0);
    const float* r0 = pts + (size_t)l0 * m;
    const float* r1 = r0 + (nl > 1 ? m : 0);
    const float* r2 = r0 + (nl > 2 ? 2 * (size_t)m : 0);
    const float* r3 = r0 + (nl > 3 ? 3 * (size_t)m : 0);
    __syncthreads();
    for (int i = tid; i < m; i += TI_THREADS) ti_tile[i] = make_float4(__ldg(r0 + i), __ldg(r1 + i), __ldg(r2 + i), __ldg(r3 + i));
    __syncthreads();
    float* o0 = o + (size_t)l0 * n;
#pragma unroll
    for (int u = 0; u < TI_PPT; ++u) {
      if (!ok[u]) continue;
      const int j = (blockIdx.x * TI_PPT + u) * TI_THREADS + tid;
      const float4 a = ti_tile[i1[u]], bq = ti_tile[i2[u]], cq = ti_tile[i3[u]];
      // fma(p3,w3, fma(p1,w1, mul(p2,w2))) per channel, as in the plain kernel
      __stcs(o0 + j, __fmaf_rn(cq.x, w3[u], __fmaf_rn(a.x, w1[u], __fmul_rn(bq.x, w2[u]))));
      if (nl > 1) __stcs(o0 + (size_t)n + j, __fmaf_rn(cq.y, w3[u], __fmaf_rn(a.y, w1[u], __fmul_rn(bq.y, w2[u]))));
      if (nl > 2) __stcs(o0 + 2 * (size_t)n + j, __fmaf_rn(cq.z, w3[u], __fmaf_rn(a.z, w1[u], __fmul_rn(bq.z, w2[u]))));
      if (nl > 3) __stcs(o0 + 3 * (size_t)n + j, __fmaf_rn(cq.w, w3[u], __fmaf_rn(a.w, w1[u], __fmul_rn(bq.w, w2[u]))));
    }
  }
}

__global__ void __launch_bounds__(256)
three_interpolate_grad_kernel(int c, int n, int m, int c_per_block,
                              const float* __restrict__ grad_out, const int* __restrict__ idx,
                              const float* __restrict__ weight, float* __restrict__ grad_points) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int c0 = blockIdx.y * c_per_block, c1 = min(c, c0 + c_per_block);
  const int* ii = idx + ((size_t)b * n + j) * 3;
  const float* ww = weight + ((size_t)b * n + j) * 3;
  const int i1 = __ldg(ii), i2 = __ldg(ii + 1), i3 = __ldg(ii + 2);
  const float w1 = __ldg(ww), w2 = __ldg(ww + 1), w3 = __ldg(ww + 2);
  const float* go = grad_out + (size_t)b * c * n;
  float* gp = grad_points + (size_t)b * c * m;
  for (int l = c0; l < c1; ++l) {
    const float g = __ldg(go + (size_t)l * n + j);
    float* row = gp + (size_t)l * m;
    atomicAdd(row + i1, g * w1);
    atomicAdd(row + i2, g * w2);
    atomicAdd(row + i3, g * w3);
  }
}

static int channel_tile(int c, long long blocks_without_c) {
  int cpb = c;
  while (cpb > 4 && blocks_without_c * cdiv(c, cpb) < 148 * 8) cpb = (cpb + 1) / 2;
  return cpb;
}

}  // namespace dfb200

using namespace dfb200;

// ============================================================================================
// C ABI
// ============================================================================================
extern "C" int dfb200_gather_points(int b, int c, int n, int npoints, const float* points,
                                    const int* idx, float* out, dfb200_stream_t stream) {
  DFB_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0, DFB200_ERR_INVALID_ARG, "gather_points: negative size");
  DFB_REQUIRE(npoints == 0 || n > 0, DFB200_ERR_INVALID_ARG, "gather_points: gathering from an empty cloud");
  return launch_group(b, c, n, npoints, points, idx, out, as_stream(stream));
}

extern "C" int dfb200_gather_points_grad(int b, int c, int n, int npoints, const float* grad_out,
                                         const int* idx, float* grad_points, dfb200_stream_t stream) {
  DFB_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0, DFB200_ERR_INVALID_ARG, "gather_points_grad: negative size");
  return launch_group_grad(b, c, n, npoints, grad_out, idx, grad_points, as_stream(stream));
}

extern "C" int dfb200_group_points(int b, int c, int n, int npoints, int nsample, const float* points,
                                   const int* idx, float* out, dfb200_stream_t stream) {
  DFB_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, DFB200_ERR_INVALID_ARG, "group_points: negative size");
  DFB_REQUIRE((long long)npoints * nsample == 0 || n > 0, DFB200_ERR_INVALID_ARG, "group_points: grouping from an empty cloud");
  return launch_group(b, c, n, (long long)npoints * nsample, points, idx, out, as_stream(stream));
}

extern "C" int dfb200_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                        const float* grad_out, const int* idx, float* grad_points,
                                        dfb200_stream_t stream) {
  DFB_REQUIRE(b >= 0 && c >= 0 && n >= 0 && npoints >= 0 && nsample >= 0, DFB200_ERR_INVALID_ARG, "group_points_grad: negative size");
  return launch_group_grad(b, c, n, (long long)npoints * nsample, grad_out, idx, grad_points, as_stream(stream));
}

extern "C" int dfb200_furthest_point_sampling(int b, int n, int m, const float* dataset, float* temp,
                                              int* idxs, dfb200_stream_t stream) {
  DFB_REQUIRE(b >= 0 && n >= 0 && m >= 0, DFB200_ERR_INVALID_ARG, "furthest_point_sampling: negative size");
  if (b == 0 || m == 0) return DFB200_OK;
  DFB_REQUIRE(n > 0, DFB200_ERR_INVALID_ARG, "furthest_point_sampling: empty cloud");
  DFB_REQUIRE(n <= 16384, DFB200_ERR_UNSUPPORTED, "furthest_point_sampling: n=%d > 16384 (shared-memory resident limit) not supported", n);
  cudaStream_t st = as_stream(stream);
  const int bs = ref_opt_n_threads(n);
  int log2bs = 0;
  while ((1 << log2bs) < bs) ++log2bs;
  const int q = cdiv(n, bs);
  int qbits = 0;
  while ((1 << qbits) < q) ++qbits;
#define FPS_CASE(PPT, THREADS, REG) return launch_fps<PPT, THREADS, REG>(b, n, m, log2bs, qbits, dataset, temp, idxs, st)
  if (n <= 32) FPS_CASE(1, 32, true);
  if (n <= 128) FPS_CASE(4, 32, true);
  if (n <= 256) FPS_CASE(4, 64, true);
  if (n <= 512) FPS_CASE(4, 128, true);
  if (n <= 1024) FPS_CASE(4, 256, true);
  if (n <= 2048) FPS_CASE(4, 512, true);
  if (n <= 4096) FPS_CASE(4, 1024, true);
  if (n <= 8192) FPS_CASE(8, 1024, true);
  FPS_CASE(16, 1024, false);
#undef FPS_CASE
}

extern "C" int dfb200_query_ball_point(int b, int n, int m, float radius, int nsample,
                                       const float* new_xyz, const float* xyz, int* idx,
                                       dfb200_stream_t stream) {
  DFB_REQUIRE(b >= 0 && n >= 0 && m >= 0 && nsample >= 0, DFB200_ERR_INVALID_ARG, "query_ball_point: negative size");
  if (b == 0 || m == 0 || nsample == 0) return DFB200_OK;
  DFB_REQUIRE(b <= 65535, DFB200_ERR_INVALID_ARG, "query_ball_point: b > 65535");
  cudaStream_t st = as_stream(stream);
  const float radius2 = radius * radius;  // one fp32 multiply, as in the reference
  // centres per warp: keep >= ~4 waves of CTAs on 148 SMs but amortise the smem staging
  int cpw = 8;
  while (cpw > 1 && (long long)b * cdiv(m, 8 * cpw) < 148 * 4) cpw /= 2;
  dim3 grid(cdiv(m, 8 * cpw), b);
  const size_t smem = sizeof(float) * 3 * (size_t)n;
  if (smem <= 160 * 1024) {
    if (smem > 48 * 1024)
      DFB_CUDA(cudaFuncSetAttribute(ball_query_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ball_query_kernel<true><<<grid, 256, smem, st>>>(n, m, radius2, nsample, cpw, new_xyz, xyz, idx);
  } else {
    ball_query_kernel<false><<<grid, 256, 0, st>>>(n, m, radius2, nsample, cpw, new_xyz, xyz, idx);
  }
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_three_nn(int b, int n, int m, const float* unknown, const float* known,
                               float* dist2, int* idx, dfb200_stream_t stream) {
  DFB_REQUIRE(b >= 0 && n >= 0 && m >= 0, DFB200_ERR_INVALID_ARG, "three_nn: negative size");
  if (b == 0 || n == 0) return DFB200_OK;
  DFB_REQUIRE(b <= 65535, DFB200_ERR_INVALID_ARG, "three_nn: b > 65535");
  const int threads = (long long)b * cdiv(n, 256) >= 148 * 2 ? 256 : 64;
  dim3 grid(cdiv(n, threads), b);
  three_nn_kernel<<<grid, threads, 0, as_stream(stream)>>>(n, m, unknown, known, dist2, idx);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_three_interpolate(int b, int c, int m, int n, const float* points,
                                        const int* idx, const float* weight, float* out,
                                        dfb200_stream_t stream) {
  DFB_REQUIRE(b >= 0 && c >= 0 && n >= 0 && m >= 0, DFB200_ERR_INVALID_ARG, "three_interpolate: negative size");
  if (b == 0 || c == 0 || n == 0) return DFB200_OK;
  DFB_REQUIRE(m > 0, DFB200_ERR_INVALID_ARG, "three_interpolate: empty source cloud");
  {
    // staged path: the interleaved source tile fits in shared memory and is reused by enough outputs
    const size_t smem = sizeof(float4) * (size_t)m;
    if (smem <= 48 * 1024 && n >= 4 * m && c >= 4 && b <= 65535) {
      const int gx4 = cdiv(n, TI_THREADS * TI_PPT);
      const int nquad = (c + 3) / 4;
      int qpb = nquad;
      while (qpb > 1 && (long long)gx4 * b * cdiv(nquad, qpb) < 148 * 8) qpb = (qpb + 1) / 2;
      dim3 grid4(gx4, cdiv(nquad, qpb), b);
      if (grid4.y <= 65535) {
        three_interpolate_c4_kernel<<<grid4, TI_THREADS, smem, as_stream(stream)>>>(c, m, n, qpb, points, idx, weight, out);
        DFB_LAUNCH_CHECK();
        return DFB200_OK;
      }
    }
  }
  const int gx = cdiv(n, 256);
  const int cpb = channel_tile(c, (long long)gx * b);
  dim3 grid(gx, cdiv(c, cpb), b);
  DFB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, DFB200_ERR_INVALID_ARG, "three_interpolate: grid too large");
  three_interpolate_kernel<<<grid, 256, 0, as_stream(stream)>>>(c, m, n, cpb, points, idx, weight, out);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_three_interpolate_grad(int b, int c, int n, int m, const float* grad_out,
                                             const int* idx, const float* weight, float* grad_points,
                                             dfb200_stream_t stream) {
  DFB_REQUIRE(b >= 0 && c >= 0 && n >= 0 && m >= 0, DFB200_ERR_INVALID_ARG, "three_interpolate_grad: negative size");
  cudaStream_t st = as_stream(stream);
  DFB_CUDA(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * c * m, st));
  if (b == 0 || c == 0 || n == 0) return DFB200_OK;
  const int gx = cdiv(n, 256);
  const int cpb = channel_tile(c, (long long)gx * b);
  dim3 grid(gx, cdiv(c, cpb), b);
  DFB_REQUIRE(grid.y <= 65535 && grid.z <= 65535, DFB200_ERR_INVALID_ARG, "three_interpolate_grad: grid too large");
  three_interpolate_grad_kernel<<<grid, 256, 0, st>>>(c, n, m, cpb, grad_out, idx, weight, grad_points);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}
