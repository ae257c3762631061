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
#include <stdlib.h>

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

// Fast path: THREADS == the reference's block size bs (>= 32).  Then point k = tid + i*bs has
// rank (bitrev(tid) << qbits) | i, ascending in i inside a thread, so the in-thread arg-max needs
// only the strict `>` of the reference (no rank compare per point) and the thread's rank is
// `rbase | best_i`.  Skipped points (|p|^2 <= 1e-3) carry td = -1: fminf keeps them at -1 and
// `-1 > best` never fires, so there is no per-point branch; pairs of points go through the packed
// fp32x2 pipe (FADD2/FMUL2/FFMA2 round each half exactly like the scalar ops).  A round costs
// ~7 issue slots per point pair instead of ~44.
__device__ __forceinline__ float2 sq3x2(float2 dx, float2 dy, float2 dz) {
  return __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
}

template <int PPT, int THREADS>
__global__ void __launch_bounds__(THREADS)
fps_fast_kernel(int n, int m, int log2bs, int qbits, const float* __restrict__ dataset,
                float* __restrict__ temp, int* __restrict__ idxs) {
  extern __shared__ float fps_smem[];  // xs[n], ys[n], zs[n]
  constexpr int NW = THREADS / 32;
  constexpr int NP = (PPT + 1) / 2;
  __shared__ uint2 slot[2][NW];
  float* xs = fps_smem;
  float* ys = xs + n;
  float* zs = ys + n;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* pts = dataset + (size_t)b * n * 3;
  int* out = idxs + (size_t)b * m;

  for (int i = tid; i < n * 3; i += THREADS) {
    const float v = __ldg(pts + i);
    const int k = i / 3, ch = i - k * 3;
    (ch == 0 ? xs : ch == 1 ? ys : zs)[k] = v;
  }
  __syncthreads();

  float2 px[NP], py[NP], pz[NP], td[NP];
#pragma unroll
  for (int i = 0; i < 2 * NP; ++i) {
    const int k = tid + i * THREADS;
    const bool in = (i < PPT) && (k < n);
    const float x = in ? xs[k] : 0.f, y = in ? ys[k] : 0.f, z = in ? zs[k] : 0.f;
    const float mag = sq3(x, y, z);
    const bool ok = in && !((double)mag <= 1e-3);  // reference: fp32 mag, double compare
    const float t0 = ok ? 1e10f : -1.f;
    if (i & 1) { px[i >> 1].y = x; py[i >> 1].y = y; pz[i >> 1].y = z; td[i >> 1].y = t0; }
    else       { px[i >> 1].x = x; py[i >> 1].x = y; pz[i >> 1].x = z; td[i >> 1].x = t0; }
  }
  const unsigned rbase = (__brev((unsigned)tid) >> (32 - log2bs)) << qbits;

  int old = 0;
  if (tid == 0 && m > 0) out[0] = 0;
  for (int j = 1; j < m; ++j) {
    const float nx = -xs[old], ny = -ys[old], nz = -zs[old];  // x - x1 == x + (-x1) exactly
    const float2 nx2 = make_float2(nx, nx), ny2 = make_float2(ny, ny), nz2 = make_float2(nz, nz);
    float best = -1.f;
    int besti = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const float2 d = sq3x2(__fadd2_rn(px[i], nx2), __fadd2_rn(py[i], ny2), __fadd2_rn(pz[i], nz2));
      const float a = fminf(d.x, td[i].x), c = fminf(d.y, td[i].y);
      td[i] = make_float2(a, c);
      bool take = a > best;
      best = take ? a : best;
      besti = take ? 2 * i : besti;
      take = c > best;
      best = take ? c : best;
      besti = take ? 2 * i + 1 : besti;
    }
    const unsigned key = best < 0.f ? 0u : (__float_as_uint(best) + 1u);
    const unsigned brank = best < 0.f ? 0xFFFFFFFFu : (rbase | (unsigned)besti);
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
  if (temp != nullptr) {  // skipped points keep the caller's 1e10 initialisation
    float* t = temp + (size_t)b * n;
#pragma unroll
    for (int i = 0; i < PPT; ++i) {
      const int k = tid + i * THREADS;
      const float v = (i & 1) ? td[i >> 1].y : td[i >> 1].x;
      if (k < n) t[k] = v < 0.f ? 1e10f : v;
    }
  }
}

template <int PPT, int THREADS>
static int launch_fps_fast(int b, int n, int m, int log2bs, int qbits, const float* dataset, float* temp,
                           int* idxs, cudaStream_t st) {
  const size_t smem = sizeof(float) * 3 * (size_t)n;
  if (smem + 1024 > 48 * 1024) {  // static slots count against the 48 KB default too
    DFB_CUDA(cudaFuncSetAttribute(fps_fast_kernel<PPT, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  fps_fast_kernel<PPT, THREADS><<<b, THREADS, smem, st>>>(n, m, log2bs, qbits, dataset, temp, idxs);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

// Wide variant for n >= 512 (reference block size bs = 512): THREADS = 512 >> S threads own PPT = 2^S * q points each, so
// a round issues the same FP work from fewer warps (half / a quarter of the REDUX, barrier and slot traffic) and the
// arg-max inside a thread is a TREE over the thread's points instead of a chain.
// Tie order: point k = tid + i*THREADS has rank (bitrev9(k mod 512) << qbits) | (k / 512) with
// bitrev9(k mod 512) = bitrev9(tid) | bitrev_S(i mod 2^S).  Register slot o = c * 2^qbits + w holds the point
// i = w * 2^S + bitrev_S(c); then rank = (bitrev9(tid) << qbits) | o: ascending in the slot number, so "left operand wins
// ties" at every tree level reproduces the reference's winner.
__host__ __device__ constexpr int fps_brev_small(int c, int bits) {
  int r = 0;
  for (int i = 0; i < bits; ++i) r |= ((c >> i) & 1) << (bits - 1 - i);
  return r;
}

template <int PPT, int THREADS, int S>
__global__ void __launch_bounds__(THREADS)
fps_wide_kernel(int n, int m, int qbits, const float* __restrict__ dataset, float* __restrict__ temp,
                int* __restrict__ idxs) {
  extern __shared__ float fps_smem[];  // xs[n], ys[n], zs[n]
  constexpr int NW = THREADS / 32;
  constexpr int NP = PPT / 2;
  constexpr int QP = PPT >> S;  // slots per residue class = 2^qbits (launch_fps_wide checks it)
  static_assert(PPT >= 2 && (PPT & (PPT - 1)) == 0, "PPT must be a power of two >= 2");
  __shared__ uint2 slot[2][NW];
  float* xs = fps_smem;
  float* ys = xs + n;
  float* zs = ys + n;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* pts = dataset + (size_t)b * n * 3;
  int* out = idxs + (size_t)b * m;

  for (int i = tid; i < n * 3; i += THREADS) {
    const float v = __ldg(pts + i);
    const int k = i / 3, ch = i - k * 3;
    (ch == 0 ? xs : ch == 1 ? ys : zs)[k] = v;
  }
  __syncthreads();

  float2 px[NP], py[NP], pz[NP], td[NP];
#pragma unroll
  for (int o = 0; o < PPT; ++o) {
    const int i = (o % QP) * (1 << S) + fps_brev_small(o / QP, S);
    const int k = tid + i * THREADS;
    const bool in = k < n;
    const float x = in ? xs[k] : 0.f, y = in ? ys[k] : 0.f, z = in ? zs[k] : 0.f;
    const float mag = sq3(x, y, z);
    const bool ok = in && !((double)mag <= 1e-3);  // reference: fp32 mag, double compare
    const float t0 = ok ? 1e10f : -1.f;
    if (o & 1) { px[o >> 1].y = x; py[o >> 1].y = y; pz[o >> 1].y = z; td[o >> 1].y = t0; }
    else       { px[o >> 1].x = x; py[o >> 1].x = y; pz[o >> 1].x = z; td[o >> 1].x = t0; }
  }
  const unsigned rbase = (__brev((unsigned)tid) >> 23) << qbits;  // bitrev9(tid); its low S bits are zero

  int old = 0;
  if (tid == 0 && m > 0) out[0] = 0;
  for (int j = 1; j < m; ++j) {
    const float nx = -xs[old], ny = -ys[old], nz = -zs[old];
    const float2 nx2 = make_float2(nx, nx), ny2 = make_float2(ny, ny), nz2 = make_float2(nz, nz);
    float bv[NP];
    int bo[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const float2 d = sq3x2(__fadd2_rn(px[i], nx2), __fadd2_rn(py[i], ny2), __fadd2_rn(pz[i], nz2));
      const float a = fminf(d.x, td[i].x), c = fminf(d.y, td[i].y);
      td[i] = make_float2(a, c);
      const bool take = c > a;  // strict: the lower slot keeps ties
      bv[i] = take ? c : a;
      bo[i] = 2 * i + (take ? 1 : 0);
    }
#pragma unroll
    for (int w = 1; w < NP; w <<= 1) {
#pragma unroll
      for (int i = 0; i + w < NP; i += 2 * w) {
        const bool take = bv[i + w] > bv[i];
        bv[i] = take ? bv[i + w] : bv[i];
        bo[i] = take ? bo[i + w] : bo[i];
      }
    }
    const float best = bv[0];
    const unsigned key = best < 0.f ? 0u : (__float_as_uint(best) + 1u);
    const unsigned brank = best < 0.f ? 0xFFFFFFFFu : (rbase | (unsigned)bo[0]);
    unsigned wkey = __reduce_max_sync(0xFFFFFFFFu, key);
    unsigned wrank = __reduce_min_sync(0xFFFFFFFFu, key == wkey ? brank : 0xFFFFFFFFu);
    if (NW > 1) {
      if (lane == 0) slot[j & 1][warp] = make_uint2(wkey, wrank);
      __syncthreads();
      const uint2 sv = lane < NW ? slot[j & 1][lane] : make_uint2(0u, 0xFFFFFFFFu);
      wkey = __reduce_max_sync(0xFFFFFFFFu, sv.x);
      wrank = __reduce_min_sync(0xFFFFFFFFu, sv.x == wkey ? sv.y : 0xFFFFFFFFu);
    }
    old = (wkey == 0u) ? 0 : fps_unrank(wrank, 9, qbits);
    if (tid == 0) out[j] = old;
  }
  if (temp != nullptr) {  // skipped points keep the caller's 1e10 initialisation
    float* t = temp + (size_t)b * n;
#pragma unroll
    for (int o = 0; o < PPT; ++o) {
      const int i = (o % QP) * (1 << S) + fps_brev_small(o / QP, S);
      const int k = tid + i * THREADS;
      const float v = (o & 1) ? td[o >> 1].y : td[o >> 1].x;
      if (k < n) t[k] = v < 0.f ? 1e10f : v;
    }
  }
}

template <int PPT, int THREADS, int S>
static int launch_fps_wide(int b, int n, int m, int qbits, const float* dataset, float* temp, int* idxs, cudaStream_t st) {
  DFB_REQUIRE((PPT >> S) == (1 << qbits), DFB200_ERR_INVALID_ARG, "fps: internal slot/rank mismatch");
  const size_t smem = sizeof(float) * 3 * (size_t)n;
  if (smem + 1024 > 48 * 1024) {
    DFB_CUDA(cudaFuncSetAttribute(fps_wide_kernel<PPT, THREADS, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  fps_wide_kernel<PPT, THREADS, S><<<b, THREADS, smem, st>>>(n, m, qbits, dataset, temp, idxs);
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
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
  if (smem + 1024 > 48 * 1024) {  // static slots count against the 48 KB default too
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
// Ordered scan of one centre by one warp (brute force, exact early exit).
template <bool SMEM>
__device__ __forceinline__ void bq_warp_scan(int n, float radius2, int nsample, float cx, float cy, float cz,
                                             const float* xs, const float* ys, const float* zs,
                                             const float* __restrict__ pts, int* __restrict__ o, int lane) {
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

template <bool SMEM>
__global__ void __launch_bounds__(256)
ball_query_kernel(int n, int m, float radius2, int nsample, int centres_per_warp, int only_marked,
                  const float* __restrict__ new_xyz, const float* __restrict__ xyz,
                  int* __restrict__ idx) {
  extern __shared__ float bq_smem[];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // second launch of the grid path: only the tiles the grid kernel handed over (sentinel -1 in the
  // first output slot of the tile, which this CTA alone overwrites)
  if (only_marked && idx[((size_t)b * m + (size_t)blockIdx.x * 8 * centres_per_warp) * nsample] != -1) return;
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
    bq_warp_scan<SMEM>(n, radius2, nsample, cx, cy, cz, xs, ys, zs, pts, idx + ((size_t)b * m + j) * nsample, lane);
  }
}

// Shared-memory variant of the ordered scan on the packed fp32x2 pipe: the cloud is staged NEGATED (c - p == c + (-p)
// exactly) as three arrays padded to a multiple of 4 points, a lane owns 4 CONSECUTIVE candidates per 128-candidate step
// (three LDS.128 instead of twelve LDS, 12 packed FP instructions instead of 24) and the compaction counts the hits of
// the lower lanes over the four ballots -- index order is lane-major, so the output is unchanged.
__global__ void __launch_bounds__(256)
ball_query_scan4_kernel(int n, int m, float radius2, int nsample, int centres_per_warp, int only_marked,
                        const float* __restrict__ new_xyz, const float* __restrict__ xyz, int* __restrict__ idx) {
  extern __shared__ __align__(16) float bq4_smem[];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (only_marked && idx[((size_t)b * m + (size_t)blockIdx.x * 8 * centres_per_warp) * nsample] != -1) return;
  const int npad = (n + 3) & ~3;
  const float* pts = xyz + (size_t)b * n * 3;
  float* xs = bq4_smem;
  float* ys = xs + npad;
  float* zs = ys + npad;
  for (int i = threadIdx.x; i < npad * 3; i += blockDim.x) {
    const int k = i / 3, ch = i - k * 3;
    (ch == 0 ? xs : ch == 1 ? ys : zs)[k] = k < n ? -__ldg(pts + i) : 0.f;
  }
  __syncthreads();
  const float4* xs4 = reinterpret_cast<const float4*>(xs);
  const float4* ys4 = reinterpret_cast<const float4*>(ys);
  const float4* zs4 = reinterpret_cast<const float4*>(zs);
  const unsigned lt = lanemask_lt();
  const int j0 = (blockIdx.x * 8 + warp) * centres_per_warp;
  for (int jj = 0; jj < centres_per_warp; ++jj) {
    const int j = j0 + jj;
    if (j >= m) break;
    const float* cq = new_xyz + ((size_t)b * m + j) * 3;
    const float cx = __ldg(cq), cy = __ldg(cq + 1), cz = __ldg(cq + 2);
    const float2 cx2 = make_float2(cx, cx), cy2 = make_float2(cy, cy), cz2 = make_float2(cz, cz);
    int* o = idx + ((size_t)b * m + j) * nsample;
    int cnt = 0, first = 0;
    for (int k0 = 0; k0 < n && cnt < nsample; k0 += 128) {
      const int k = k0 + 4 * lane;
      bool h0 = false, h1 = false, h2 = false, h3 = false;
      if (k < n) {
        const float4 X = xs4[k >> 2], Y = ys4[k >> 2], Z = zs4[k >> 2];
        const float2 da = sq3x2(__fadd2_rn(cx2, make_float2(X.x, X.y)), __fadd2_rn(cy2, make_float2(Y.x, Y.y)),
                                __fadd2_rn(cz2, make_float2(Z.x, Z.y)));
        const float2 db = sq3x2(__fadd2_rn(cx2, make_float2(X.z, X.w)), __fadd2_rn(cy2, make_float2(Y.z, Y.w)),
                                __fadd2_rn(cz2, make_float2(Z.z, Z.w)));
        h0 = da.x < radius2;  // reference: `if (d2 < radius2)`
        h1 = (k + 1 < n) && da.y < radius2;
        h2 = (k + 2 < n) && db.x < radius2;
        h3 = (k + 3 < n) && db.y < radius2;
      }
      const unsigned m0 = __ballot_sync(0xFFFFFFFFu, h0), m1 = __ballot_sync(0xFFFFFFFFu, h1),
                     m2 = __ballot_sync(0xFFFFFFFFu, h2), m3 = __ballot_sync(0xFFFFFFFFu, h3);
      const unsigned any = m0 | m1 | m2 | m3;
      if (any != 0u) {
        if (cnt == 0) {
          const int fl = __ffs(any) - 1;  // lowest lane with a hit; its lowest hit is the first hit
          const unsigned hb = ((m0 >> fl) & 1u) | (((m1 >> fl) & 1u) << 1) | (((m2 >> fl) & 1u) << 2) | (((m3 >> fl) & 1u) << 3);
          first = k0 + 4 * fl + __ffs(hb) - 1;
        }
        int pos = cnt + __popc(m0 & lt) + __popc(m1 & lt) + __popc(m2 & lt) + __popc(m3 & lt);
        if (h0) { if (pos < nsample) o[pos] = k; ++pos; }
        if (h1) { if (pos < nsample) o[pos] = k + 1; ++pos; }
        if (h2) { if (pos < nsample) o[pos] = k + 2; ++pos; }
        if (h3) { if (pos < nsample) o[pos] = k + 3; ++pos; }
        cnt += __popc(m0) + __popc(m1) + __popc(m2) + __popc(m3);
      }
    }
    for (int l = min(cnt, nsample) + lane; l < nsample; l += 32) o[l] = first;
  }
}

// --------------------------------------------------------------------------------------------
// Thread-per-centre scan (round 2).  The warp-per-centre scans above spend most of their issue slots on the ORDERED
// compaction (4 ballots + prefix popcounts per 128 candidates) and on per-centre set-up; here a THREAD owns a centre and walks
// the whole cloud in index order, so "the first nsample hits in ascending k" needs no cross-lane work at all:
//   * the cloud is staged once per CTA, negated (c - p == c + (-p) exactly), as three arrays padded to a multiple of 32 points
//     with a never-hit value; every lane of a warp reads the SAME four points per step (one broadcast LDS.128 per coordinate);
//   * 4 candidates cost 6 packed fp32x2 instructions (same rounding as the reference's fma chain, sq3x2) + 4 compares that
//     set bits of a per-thread 32-candidate hit mask;
//   * after 32 candidates the (few) set bits are appended, in ascending order, to the thread's slot list in shared memory
//     ([slot][33] halfwords per warp: conflict-free for the thread-private writes and for the transposed read-back);
//   * a warp leaves the loop as soon as all its 32 centres hold nsample hits (dense balls stop early, sparse ones read the
//     cloud once: 6.4 issue slots per candidate and centre, no grid, no sort, no bitmap);
//   * the warp then writes the 32 rows of idx with coalesced stores, padding with the first hit (0 for an empty ball).
// Needs n <= 65535 (halfword slots) and the cloud + slot lists in shared memory.
// --------------------------------------------------------------------------------------------
// A thread owns BQT_C centres (register tile): the broadcast loads return 512 bytes per LDS.128 to the register file, and with
// one centre per thread that return path (128 B/clk per SM) bound the kernel at 3 loads per 4 candidates; the loaded points
// are reused for BQT_C centres.
constexpr int BQT_THREADS = 256;
__host__ __device__ inline size_t bqt_smem_bytes(int n, int nsample, int C) {
  const size_t npad = ((size_t)n + 31) & ~(size_t)31;
  return npad * 12 + (size_t)(BQT_THREADS / 32) * C * nsample * 33 * 2;
}
template <int BQT_C>
__global__ void __launch_bounds__(BQT_THREADS)
ball_query_tpc_kernel(int n, int m, float radius2, int nsample, const float* __restrict__ new_xyz, const float* __restrict__ xyz,
                      int* __restrict__ idx) {
  extern __shared__ __align__(16) float bqt_smem[];
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npad = (n + 31) & ~31;
  float* xs = bqt_smem;
  float* ys = xs + npad;
  float* zs = ys + npad;
  unsigned short* slots = reinterpret_cast<unsigned short*>(zs + npad) + (size_t)warp * BQT_C * nsample * 33;  // [c][slot][33]
  const float* pts = xyz + (size_t)b * n * 3;
  for (int i = threadIdx.x; i < npad * 3; i += BQT_THREADS) {
    const int k = i / 3, ch = i - k * 3;
    (ch == 0 ? xs : ch == 1 ? ys : zs)[k] = k < n ? -__ldg(pts + i) : -3.0e38f;  // padding: (c + 3e38)^2 overflows to +inf, never < r^2
  }
  __syncthreads();
  // centre c of this thread: j0 + c * BQT_THREADS + tid  (a warp's 32 lanes hold 32 consecutive centres for every c)
  const int j0 = blockIdx.x * (BQT_THREADS * BQT_C);
  float2 cx2[BQT_C], cy2[BQT_C], cz2[BQT_C];
  int cnt[BQT_C];
#pragma unroll
  for (int c = 0; c < BQT_C; ++c) {
    const int j = j0 + c * BQT_THREADS + threadIdx.x;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (j < m) {
      const float* cq = new_xyz + ((size_t)b * m + j) * 3;
      cx = __ldg(cq); cy = __ldg(cq + 1); cz = __ldg(cq + 2);
    }
    cx2[c] = make_float2(cx, cx); cy2[c] = make_float2(cy, cy); cz2[c] = make_float2(cz, cz);
    cnt[c] = j < m ? 0 : nsample;  // lanes without a centre count as full
  }
  const float4* xs4 = reinterpret_cast<const float4*>(xs);
  const float4* ys4 = reinterpret_cast<const float4*>(ys);
  const float4* zs4 = reinterpret_cast<const float4*>(zs);
  const int r2bits = __float_as_int(radius2);
#pragma unroll 1
  for (int k0 = 0; k0 < npad; k0 += 32) {
    // candidate i of the block ends up at bit 31 - i of mask[c]: d2 >= 0 (or NaN / +inf, which never hit), so its bit pattern
    // compares like an integer and (d2bits - r2bits) has its sign bit set exactly when d2 < r2 (reference: `if (d2 < radius2)`);
    // a funnel shift moves that bit into the mask
    unsigned mask[BQT_C];
#pragma unroll
    for (int c = 0; c < BQT_C; ++c) mask[c] = 0u;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 X = xs4[(k0 >> 2) + q], Y = ys4[(k0 >> 2) + q], Z = zs4[(k0 >> 2) + q];  // warp-uniform addresses: broadcast
#pragma unroll
      for (int c = 0; c < BQT_C; ++c) {
        const float2 da = sq3x2(__fadd2_rn(cx2[c], make_float2(X.x, X.y)), __fadd2_rn(cy2[c], make_float2(Y.x, Y.y)),
                                __fadd2_rn(cz2[c], make_float2(Z.x, Z.y)));
        const float2 db = sq3x2(__fadd2_rn(cx2[c], make_float2(X.z, X.w)), __fadd2_rn(cy2[c], make_float2(Y.z, Y.w)),
                                __fadd2_rn(cz2[c], make_float2(Z.z, Z.w)));
        mask[c] = __funnelshift_l((unsigned)(__float_as_int(da.x) - r2bits), mask[c], 1);
        mask[c] = __funnelshift_l((unsigned)(__float_as_int(da.y) - r2bits), mask[c], 1);
        mask[c] = __funnelshift_l((unsigned)(__float_as_int(db.x) - r2bits), mask[c], 1);
        mask[c] = __funnelshift_l((unsigned)(__float_as_int(db.y) - r2bits), mask[c], 1);
      }
    }
    bool full = true;
#pragma unroll
    for (int c = 0; c < BQT_C; ++c) {
      unsigned mk = mask[c];
      unsigned short* sl = slots + (size_t)c * nsample * 33;
      while (mk != 0u && cnt[c] < nsample) {  // ascending k within the block; blocks are visited in ascending order
        const int bpos = __clz(mk);
        sl[cnt[c] * 33 + lane] = (unsigned short)(k0 + bpos);
        ++cnt[c];
        mk &= ~(0x80000000u >> bpos);
      }
      full = full && cnt[c] >= nsample;
    }
    if (__all_sync(0xFFFFFFFFu, full)) break;
  }
  // write-out: the warp emits its 32 rows per centre set one after the other (coalesced), padding with the first hit / 0
  __syncwarp();  // the slot lists are read across lanes below (the vote above does not order shared-memory accesses)
#pragma unroll
  for (int c = 0; c < BQT_C; ++c) {
    const unsigned short* sl = slots + (size_t)c * nsample * 33;
    const int jw = j0 + c * BQT_THREADS + warp * 32;
    const int first = (jw + lane < m && cnt[c] > 0) ? (int)sl[lane] : 0;
    for (int r = 0; r < 32 && jw + r < m; ++r) {
      const int c_r = __shfl_sync(0xFFFFFFFFu, cnt[c], r), f_r = __shfl_sync(0xFFFFFFFFu, first, r);
      int* o = idx + ((size_t)b * m + jw + r) * nsample;
      for (int l = lane; l < nsample; l += 32) o[l] = l < c_r ? (int)sl[l * 33 + r] : f_r;
    }
  }
}

// --------------------------------------------------------------------------------------------
// Grid path (1024 <= n <= 8192): sparse balls (r = 0.1 / 0.2 on a unit-scale cloud hold 6 / 43 of
// 2048 points on average, so 77-92 % of the centres scan ALL n points in the brute-force kernel).
// Each CTA bins the cloud into a uniform grid of cell size >= 1.001 r (at most 16 cells per axis)
// with a shared-memory counting sort, then a GROUP OF 8 LANES per centre walks the <= 9
// x-contiguous runs of the centre's 3x3x3 cell neighbourhood (8.7x / 57x fewer pair tests at
// r = 0.2 / 0.1; 8 lanes match the typical run length, a full warp would idle on short runs and a
// single thread is latency- and divergence-bound -- measured).  The cells are visited out of index
// order, so every hit sets a bit (shared-memory atomicOr) in a per-centre n-bit map; the group then
// reads the map back in ascending index order (popcount prefix scan) to emit "the first nsample
// hits in index order, padded with the first hit" -- bit-identical to the ordered scan.
// Degenerate inputs (non-finite bounding box, r <= 0) and dense balls (few occupied cells: no
// pruning, and the unordered walk cannot stop early) are handed over, per cloud, to the ordered-scan
// kernel launched right after on the same stream: the grid kernel writes the sentinel -1 into the
// first output slot of every scan tile it skips, and a scan CTA that does not find it exits at once.
constexpr int BQG_G = 16;
constexpr int BQG_CELLS = BQG_G * BQG_G * BQG_G;
constexpr int BQG_MIN_OCCUPIED = 135;  // ~27 n / occupied candidates per centre; above ~n/5 the ordered scan with early exit wins (measured)
constexpr int BQG_T = 256;            // threads per CTA
constexpr int BQG_LPC = 8;            // lanes per centre
constexpr int BQG_GPW = 32 / BQG_LPC; // centres per warp and pass
constexpr int BQG_MAXSORT = 2048;     // clouds with at most this many centres get them sorted by cell
__device__ __forceinline__ int bqg_phys(int c) { return c + (c >> 5); }  // de-conflicts the chunked scan
// words of one bitmap, padded so that the 4 maps of a warp start 8 banks apart
static __host__ __device__ inline int bqg_stride(int n) { return ((n + 1023) / 1024) * 32 + BQG_LPC; }

static size_t bqg_smem_bytes(int n) {
  const size_t bitmaps = 4 * (size_t)(BQG_T / BQG_LPC) * bqg_stride(n);
  const size_t counters = 4 * (size_t)(BQG_CELLS + BQG_CELLS / 32 + 1);
  const size_t u = bitmaps > counters ? bitmaps : counters;
  return 16 * (size_t)n + (size_t)((2 * (BQG_CELLS + 1) + 15) / 16 * 16) + (u + 15) / 16 * 16 + 4 * (size_t)BQG_MAXSORT;
}
static size_t bqg_keys_offset(int n) { return bqg_smem_bytes(n) - 4 * (size_t)BQG_MAXSORT; }

template <bool WPC>  // WPC: one warp per centre over a flattened candidate stream (A/B experiment); default: 8 lanes per centre
__global__ void __launch_bounds__(BQG_T, 3)
ball_query_grid_kernel(int nb, int n, int m, int mpad, int keys_off, int min_occupied, float radius, float radius2,
                       int nsample, int scan_tile,
                       const float* __restrict__ new_xyz, const float* __restrict__ xyz,
                       int* __restrict__ idx) {
  extern __shared__ __align__(16) unsigned char bqg_smem[];
  constexpr int T = BQG_T, NW = T / 32, LPC = BQG_LPC, GPW = BQG_GPW;
  constexpr int CHUNK = BQG_CELLS / T;
  __shared__ float red[NW][6];
  __shared__ unsigned wsum[NW];
  __shared__ int s_occ;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int words = (n + 31) >> 5;
  float4* S = reinterpret_cast<float4*>(bqg_smem);                                    // sorted (x, y, z, index)
  unsigned short* E = reinterpret_cast<unsigned short*>(bqg_smem + 16 * (size_t)n);    // E[c] = start of cell c
  unsigned* U = reinterpret_cast<unsigned*>(bqg_smem + 16 * (size_t)n + (2 * (BQG_CELLS + 1) + 15) / 16 * 16);
  unsigned* K = reinterpret_cast<unsigned*>(bqg_smem + keys_off);  // (cell << 16 | centre) keys, sorted
  // Balanced static partition: the work items are (cloud, pass of 32 centres), cloud-major; CTA i owns
  // a contiguous slice, i.e. a few whole or partial clouds, and builds the grid once per cloud it touches.
  const int ppc = (m + NW * GPW - 1) / (NW * GPW);
  const long long items = (long long)nb * ppc;
  const long long it_lo = items * blockIdx.x / gridDim.x, it_hi = items * (blockIdx.x + 1) / gridDim.x;
  for (long long it0 = it_lo; it0 < it_hi;) {
  const int b = (int)(it0 / ppc);
  const long long it1 = min(it_hi, (long long)(b + 1) * ppc);
  const int jlo = (int)(it0 - (long long)b * ppc) * (NW * GPW), jhi = min(m, (int)(it1 - (long long)b * ppc) * (NW * GPW));
  it0 = it1;
  const float* pts = xyz + (size_t)b * n * 3;

  // ---- bounding box of the cloud -------------------------------------------------------------
  float lx = INFINITY, ly = INFINITY, lz = INFINITY, hx = -INFINITY, hy = -INFINITY, hz = -INFINITY;
  bool fin = true;
#pragma unroll 4
  for (int k = tid; k < n; k += T) {
    const float x = __ldg(pts + 3 * k), y = __ldg(pts + 3 * k + 1), z = __ldg(pts + 3 * k + 2);
    fin = fin && isfinite(x) && isfinite(y) && isfinite(z);
    lx = fminf(lx, x); ly = fminf(ly, y); lz = fminf(lz, z);
    hx = fmaxf(hx, x); hy = fmaxf(hy, y); hz = fmaxf(hz, z);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    lx = fminf(lx, __shfl_xor_sync(0xFFFFFFFFu, lx, d)); ly = fminf(ly, __shfl_xor_sync(0xFFFFFFFFu, ly, d));
    lz = fminf(lz, __shfl_xor_sync(0xFFFFFFFFu, lz, d)); hx = fmaxf(hx, __shfl_xor_sync(0xFFFFFFFFu, hx, d));
    hy = fmaxf(hy, __shfl_xor_sync(0xFFFFFFFFu, hy, d)); hz = fmaxf(hz, __shfl_xor_sync(0xFFFFFFFFu, hz, d));
  }
  if (tid == 0) s_occ = 0;
  if (lane == 0) { red[warp][0] = lx; red[warp][1] = ly; red[warp][2] = lz; red[warp][3] = hx; red[warp][4] = hy; red[warp][5] = hz; }
  const int all_finite = __syncthreads_and(fin ? 1 : 0);
#pragma unroll
  for (int w = 0; w < NW; ++w) {
    lx = fminf(lx, red[w][0]); ly = fminf(ly, red[w][1]); lz = fminf(lz, red[w][2]);
    hx = fmaxf(hx, red[w][3]); hy = fmaxf(hy, red[w][4]); hz = fmaxf(hz, red[w][5]);
  }
  bool brute = !(all_finite && radius > 0.f && isfinite(radius));

  // ---- grid: cell size >= 1.001 r, <= 16 cells per axis ----------------------------------------
  // |p - c| < r implies the cell coordinates of p and c differ by < 1/1.001 + O(1e-5) on every axis,
  // so all hits of a centre lie in its 3x3x3 neighbourhood, with margin for the fp32 rounding of
  // the coordinates and of the d2 < r2 test.
  const float cell = radius * 1.001f;
  const float csx = fmaxf(cell, (hx - lx) * (1.0f / 15.99f)), csy = fmaxf(cell, (hy - ly) * (1.0f / 15.99f)),
              csz = fmaxf(cell, (hz - lz) * (1.0f / 15.99f));
  const float ivx = 1.0f / csx, ivy = 1.0f / csy, ivz = 1.0f / csz;
  int gx = 1, gy = 1, gz = 1;
  if (!brute) {
    gx = min((int)((hx - lx) * ivx) + 1, BQG_G);
    gy = min((int)((hy - ly) * ivy) + 1, BQG_G);
    gz = min((int)((hz - lz) * ivz) + 1, BQG_G);
  }
  const int ncell = gx * gy * gz;
  // coarse grids cannot have enough occupied cells: skip the build (r = 0.4 on a unit-scale cloud)
  brute = brute || ncell < 2 * min_occupied;
  auto cell_of = [&](float x, float y, float z) {
    const int ix = min(max((int)floorf((x - lx) * ivx), 0), gx - 1);
    const int iy = min(max((int)floorf((y - ly) * ivy), 0), gy - 1);
    const int iz = min(max((int)floorf((z - lz) * ivz), 0), gz - 1);
    return (iz * gy + iy) * gx + ix;
  };

  if (!brute) {
    // ---- counting sort of the points by cell --------------------------------------------------
    for (int i = tid; i < BQG_CELLS + BQG_CELLS / 32 + 1; i += T) U[i] = 0u;
    __syncthreads();
#pragma unroll 4
    for (int k = tid; k < n; k += T) {
      const float x = __ldg(pts + 3 * k), y = __ldg(pts + 3 * k + 1), z = __ldg(pts + 3 * k + 2);
      atomicAdd(&U[bqg_phys(cell_of(x, y, z))], 1u);
    }
    __syncthreads();
    unsigned local = 0u;
    int occ = 0;
#pragma unroll 4
    for (int i = 0; i < CHUNK; ++i) {
      const int c = tid * CHUNK + i;
      const unsigned v = c < ncell ? U[bqg_phys(c)] : 0u;
      local += v;
      occ += v != 0u;
    }
    unsigned incl = local;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const unsigned v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += v;
    }
    occ = __reduce_add_sync(0xFFFFFFFFu, occ);
    if (lane == 31) wsum[warp] = incl;
    if (lane == 0 && occ) atomicAdd(&s_occ, occ);
    __syncthreads();
    unsigned run = incl - local;
#pragma unroll
    for (int w = 0; w < NW; ++w) run += w < warp ? wsum[w] : 0u;
#pragma unroll 4
    for (int i = 0; i < CHUNK; ++i) {
      const int c = tid * CHUNK + i;
      if (c < ncell) {
        const unsigned v = U[bqg_phys(c)];
        U[bqg_phys(c)] = run;  // start of cell c; the scatter below advances it to the end of c
        run += v;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int k = tid; k < n; k += T) {
      const float x = __ldg(pts + 3 * k), y = __ldg(pts + 3 * k + 1), z = __ldg(pts + 3 * k + 2);
      const unsigned pos = atomicAdd(&U[bqg_phys(cell_of(x, y, z))], 1u);
      S[pos] = make_float4(x, y, z, __int_as_float(k));
    }
    __syncthreads();
    if (tid == 0) E[0] = 0;
    for (int c = tid; c < ncell; c += T) E[c + 1] = (unsigned short)U[bqg_phys(c)];
    brute = s_occ < min_occupied;
    __syncthreads();
  }

  if (brute) {  // hand this slice over to the ordered-scan kernel launched next on the same stream
    for (int j = jlo + tid; j < jhi; j += T) {
      if (j % scan_tile == 0) idx[((size_t)b * m + j) * nsample] = -1;
    }
    __syncthreads();  // the next cloud's build reuses shared memory
    continue;
  }

  if (WPC) {
    // ---- ONE WARP per centre over a FLATTENED candidate stream (round 2) ------------------------------------------------
    // The 9 x-contiguous runs of the 3x3x3 neighbourhood hold ~9 candidates each at r = 0.2: with 8 lanes per centre and one
    // run at a time the pair-test loop ran at 35 % lane utilisation and a centre cost 775 warp-instructions.  Here lanes 0..8
    // fetch the 9 run bounds, a 9-wide prefix sum lays the runs end to end, and lane L tests flat positions L, L+32, ...: every
    // lane keeps a cursor (current run, its flat end) that it advances past finished / empty runs, so all 32 lanes test
    // candidates until the stream is exhausted (~7 iterations for ~216 candidates).  Hits set bits in the warp's n-bit map
    // (shared-memory atomicOr: candidates arrive in cell order), which the warp reads back in ascending index order -- lane L
    // owns words [L*wpl, (L+1)*wpl) -- with ONE 32-lane popcount prefix scan: bit-identical to the ordered scan.  No centre
    // sorting (no inter-centre divergence inside a warp), no group shuffles.
    __shared__ int2 s_runs[NW][10];
    const int wstride = ((n + 1023) / 1024) * 32;  // words of one warp's map (multiple of 32)
    const int wpl = wstride / 32;                  // words per lane, <= 8 (n <= 8192)
    unsigned* BMw = U + (size_t)warp * wstride;
    for (int i = lane; i < wstride; i += 32) BMw[i] = 0u;
    __syncwarp();
    const int npass = (jhi - jlo + NW * GPW - 1) / (NW * GPW);
    for (int pass = 0; pass < npass; ++pass) {
#pragma unroll 1
      for (int cc = 0; cc < GPW; ++cc) {
        const int j = jlo + pass * NW * GPW + warp * GPW + cc;
        if (j >= jhi) break;  // warp-uniform
        const float* cq = new_xyz + ((size_t)b * m + j) * 3;
        const float cx = __ldg(cq), cy = __ldg(cq + 1), cz = __ldg(cq + 2);
        const int icx = (int)fminf(fmaxf(floorf((cx - lx) * ivx), -1.f), (float)gx);
        const int icy = (int)fminf(fmaxf(floorf((cy - ly) * ivy), -1.f), (float)gy);
        const int icz = (int)fminf(fmaxf(floorf((cz - lz) * ivz), -1.f), (float)gz);
        const int x0 = max(icx - 1, 0), x1 = min(icx + 1, gx - 1);
        int kb = 0, len = 0;
        if (lane < 9) {
          const int dz = lane / 3, dy = lane - 3 * dz;
          const int yy = icy - 1 + dy, zz = icz - 1 + dz;
          if (x0 <= x1 && yy >= 0 && yy < gy && zz >= 0 && zz < gz) {
            const int row = (zz * gy + yy) * gx;
            kb = E[row + x0];
            len = (int)E[row + x1 + 1] - kb;
          }
        }
        int incl = len;
#pragma unroll
        for (int d = 1; d < 16; d <<= 1) {
          const int v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
          if (lane >= d) incl += v;
        }
        const int total = __shfl_sync(0xFFFFFFFFu, incl, 8);
        if (lane < 9) s_runs[warp][lane] = make_int2(kb, incl);  // (start in S, flat END of the run)
        __syncwarp();
        {
          int rcur = 0, obeg = 0;
          int2 cur = s_runs[warp][0];
          for (int p = lane; p - lane < total; p += 32) {  // warp-uniform trip count
            if (p < total) {
              while (p >= cur.y) {  // past the current run (or an empty one): total = s_runs[8].y > p bounds rcur by 8
                obeg = cur.y;
                cur = s_runs[warp][++rcur];
              }
              const float4 pt = S[cur.x + (p - obeg)];
              if (sq3(cx - pt.x, cy - pt.y, cz - pt.z) < radius2) {
                const int q = __float_as_int(pt.w);
                atomicOr(&BMw[q >> 5], 1u << (q & 31));
              }
            }
          }
        }
        __syncwarp();
        // read-back in ascending index order
        int* o = idx + ((size_t)b * m + j) * nsample;
        unsigned wv[8];
        int c = 0, myfirst = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          wv[i] = i < wpl ? BMw[lane * wpl + i] : 0u;
          if (c == 0 && wv[i] != 0u) myfirst = (lane * wpl + i) * 32 + __ffs(wv[i]) - 1;
          c += __popc(wv[i]);
        }
        int inc2 = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int v = __shfl_up_sync(0xFFFFFFFFu, inc2, d);
          if (lane >= d) inc2 += v;
        }
        const int cnt = __shfl_sync(0xFFFFFFFFu, inc2, 31);
        const unsigned nz = __ballot_sync(0xFFFFFFFFu, c != 0);
        const int ffirst = __shfl_sync(0xFFFFFFFFu, myfirst, nz != 0u ? __ffs(nz) - 1 : 0);
        const int first = cnt > 0 ? ffirst : 0;
        int pos = inc2 - c;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          unsigned v = wv[i];
          if (v != 0u) {
            BMw[lane * wpl + i] = 0u;  // ready for the warp's next centre
            const int base = (lane * wpl + i) * 32;
            while (v != 0u && pos < nsample) {
              o[pos++] = base + __ffs(v) - 1;
              v &= v - 1u;
            }
          }
        }
        for (int l = min(cnt, nsample) + lane; l < nsample; l += 32) o[l] = first;
        __syncwarp();
      }
    }
    __syncthreads();  // the next cloud's build overwrites the maps / sorted array
    continue;
  }

  // ---- 8 lanes per centre: walk the neighbourhood, mark hits, read the map back in index order ----
  // All loops below are WARP-uniform (the four groups of a warp run in lockstep, idle lanes are
  // predicated): group-divergent control flow was measured 4x slower (the groups serialise).
  const int stride = bqg_stride(n);
  const int wpl = (stride - LPC) / LPC;  // map words per lane (a multiple of 4)
  const int g = lane / LPC, gl = lane % LPC;
  unsigned* BM = U + (size_t)(warp * GPW + g) * stride;
  for (int i = gl; i < stride; i += LPC) BM[i] = 0u;
  __syncwarp();
  // Centres in FPS order are scattered over the cloud, so the four groups of a warp see very different
  // run lengths (the warp pays the maximum).  Sort the cloud's centres by cell (bitonic sort of unique
  // (cell << 16 | centre) keys: deterministic, every CTA working on this cloud derives the same order) and
  // let the slices index the sorted order: 33 -> 21 pair-test iterations per warp and pass at r = 0.2.
  // The sort costs ~9 us per cloud and CTA: worth it only when a CTA runs several passes of the cloud and
  // the pair-test loop dominates (many candidates per centre).  Both tests use cloud-/launch-level numbers
  // only, so every CTA working on this cloud takes the same decision.
  const bool sorted = mpad > 0 && items / gridDim.x >= 6 && 27 * n >= 128 * s_occ;
  if (sorted) {
    for (int i = tid; i < mpad; i += T) {
      unsigned key = 0xFFFFFFFFu;
      if (i < m) {
        const float* cq = new_xyz + ((size_t)b * m + i) * 3;
        key = ((unsigned)cell_of(__ldg(cq), __ldg(cq + 1), __ldg(cq + 2)) << 16) | (unsigned)i;
      }
      K[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= mpad; k <<= 1) {
      for (int jj = k >> 1; jj > 0; jj >>= 1) {
        for (int i = tid; i < mpad; i += T) {
          const int ixj = i ^ jj;
          if (ixj > i) {
            const unsigned a = K[i], c = K[ixj];
            if ((a > c) == ((i & k) == 0)) { K[i] = c; K[ixj] = a; }
          }
        }
        __syncthreads();
      }
    }
  }
  const int npass = (jhi - jlo + NW * GPW - 1) / (NW * GPW);
  for (int pass = 0; pass < npass; ++pass) {
    const int jpos = jlo + pass * NW * GPW + warp * GPW + g;
    const bool valid = jpos < jhi;
    const int j = (valid && sorted) ? (int)(K[jpos] & 0xFFFFu) : jpos;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (valid) {
      const float* cq = new_xyz + ((size_t)b * m + j) * 3;
      cx = __ldg(cq); cy = __ldg(cq + 1); cz = __ldg(cq + 2);
    }
    // cell of the centre, clamped to [-1, g] (centres may lie outside the cloud's box; NaN -> -1)
    const int icx = (int)fminf(fmaxf(floorf((cx - lx) * ivx), -1.f), (float)gx);
    const int icy = (int)fminf(fmaxf(floorf((cy - ly) * ivy), -1.f), (float)gy);
    const int icz = (int)fminf(fmaxf(floorf((cz - lz) * ivz), -1.f), (float)gz);
    const int x0 = max(icx - 1, 0), x1 = min(icx + 1, gx - 1);
    // bounds of the 9 (dy, dz) runs (x-contiguous cells are one run): lane r of the group loads run r,
    // lane 0 also run 8 -- one shared-memory latency for all of them
    auto run_bounds = [&](int r, int& kb, int& e) {
      const int dz = r / 3, dy = r - 3 * dz;
      const int yy = icy - 1 + dy, zz = icz - 1 + dz;
      kb = 0; e = 0;
      if (valid && x0 <= x1 && yy >= 0 && yy < gy && zz >= 0 && zz < gz) {
        const int row = (zz * gy + yy) * gx;
        kb = E[row + x0];
        e = E[row + x1 + 1];
      }
    };
    int kbA, eA, kbB = 0, eB = 0;
    run_bounds(gl, kbA, eA);
    if (gl == 0) run_bounds(8, kbB, eB);
#pragma unroll 1
    for (int r = 0; r < 9; ++r) {
      const int src = g * LPC + (r & 7);
      const int e = __shfl_sync(0xFFFFFFFFu, r < 8 ? eA : eB, src);
      int k = __shfl_sync(0xFFFFFFFFu, r < 8 ? kbA : kbB, src) + gl;
      while (__any_sync(0xFFFFFFFFu, k < e)) {
        const int k1 = k + LPC;
        float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
        if (k < e) p0 = S[k];
        if (k1 < e) p1 = S[k1];
        const float d0 = sq3(cx - p0.x, cy - p0.y, cz - p0.z);
        const float d1 = sq3(cx - p1.x, cy - p1.y, cz - p1.z);
        if (k < e && d0 < radius2) {
          const int q = __float_as_int(p0.w);
          atomicOr(&BM[q >> 5], 1u << (q & 31));
        }
        if (k1 < e && d1 < radius2) {
          const int q = __float_as_int(p1.w);
          atomicOr(&BM[q >> 5], 1u << (q & 31));
        }
        k += 2 * LPC;
      }
    }
    __syncwarp();

    // read-back: lane gl owns the wpl consecutive words [gl*wpl, (gl+1)*wpl) of its group's map
    // (ascending index order across the group), so ONE prefix scan over the 8 lane totals places
    // every lane's hits; two 128-bit loads fetch a lane's words of a 2048-point cloud.
    int* o = idx + ((size_t)b * m + (valid ? j : 0)) * nsample;
    uint4* BM4 = reinterpret_cast<uint4*>(BM + gl * wpl);
    int c = 0, myfirst = 0;
    for (int i = 0; i < wpl / 4; ++i) {
      const uint4 v = BM4[i];
      const unsigned wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (c == 0 && wv[u] != 0u) myfirst = (gl * wpl + i * 4 + u) * 32 + __ffs(wv[u]) - 1;
        c += __popc(wv[u]);
      }
    }
    int incl = c;
#pragma unroll
    for (int d = 1; d < LPC; d <<= 1) {
      const int v = __shfl_up_sync(0xFFFFFFFFu, incl, d, LPC);
      if (gl >= d) incl += v;
    }
    const int cnt = __shfl_sync(0xFFFFFFFFu, incl, LPC - 1, LPC);
    const unsigned nz = (__ballot_sync(0xFFFFFFFFu, c != 0) >> (g * LPC)) & ((1u << LPC) - 1u);
    const int ffirst = __shfl_sync(0xFFFFFFFFu, myfirst, (__ffs(nz) - 1) & (LPC - 1), LPC);  // all lanes: full-mask shuffle
    const int first = cnt > 0 ? ffirst : 0;
    int pos = incl - c;
    for (int i = 0; i < wpl / 4; ++i) {
      const uint4 v = BM4[i];
      if (c != 0) BM4[i] = make_uint4(0u, 0u, 0u, 0u);  // ready for the group's next centre
      unsigned wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int base = (gl * wpl + i * 4 + u) * 32;
        while (wv[u] != 0u && pos < nsample) {
          o[pos++] = base + __ffs(wv[u]) - 1;
          wv[u] &= wv[u] - 1u;
        }
      }
    }
    if (valid) {
      for (int l = min(cnt, nsample) + gl; l < nsample; l += LPC) o[l] = first;
    }
    __syncwarp();
  }
  __syncthreads();  // the next cloud's build overwrites the maps / sorted array
  }  // clouds of this CTA
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

// Two unknown points per thread on the packed fp32x2 pipe (3 FP issue slots per pair test instead of 6), known points
// staged NEGATED as float4 (u - k == u + (-k) exactly; one broadcast LDS.128 per known point), and ONE compare against
// the current third-best as the common case: the three-way insertion (the reference's strict `<` chain, so equal
// distances keep the earlier index) runs only for the ~3 ln(m) candidates per point that enter the top three.
__device__ __forceinline__ void nn3_insert(float d, int k, float& b1, float& b2, float& b3, int& i1, int& i2, int& i3) {
  if (d < b1) {
    b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k;
  } else if (d < b2) {
    b3 = b2; i3 = i2; b2 = d; i2 = k;
  } else {
    b3 = d; i3 = k;
  }
}

constexpr int NN2_TILE = 1024;
__global__ void __launch_bounds__(256)
three_nn2_kernel(int n, int m, const float* __restrict__ unknown, const float* __restrict__ known,
                 float* __restrict__ dist2, int* __restrict__ idx) {
  __shared__ float4 kt[NN2_TILE];
  const int b = blockIdx.y;
  const int j0 = blockIdx.x * (2 * blockDim.x) + threadIdx.x, j1 = j0 + blockDim.x;
  const float* kn = known + (size_t)b * m * 3;
  float2 ux = make_float2(0.f, 0.f), uy = ux, uz = ux;
  if (j0 < n) {
    const float* u = unknown + ((size_t)b * n + j0) * 3;
    ux.x = __ldg(u); uy.x = __ldg(u + 1); uz.x = __ldg(u + 2);
  }
  if (j1 < n) {
    const float* u = unknown + ((size_t)b * n + j1) * 3;
    ux.y = __ldg(u); uy.y = __ldg(u + 1); uz.y = __ldg(u + 2);
  }
  float a1 = INFINITY, a2 = INFINITY, a3 = INFINITY, c1 = INFINITY, c2 = INFINITY, c3 = INFINITY;
  int ia1 = 0, ia2 = 0, ia3 = 0, ic1 = 0, ic2 = 0, ic3 = 0;
  for (int k0 = 0; k0 < m; k0 += NN2_TILE) {
    const int tile = min(NN2_TILE, m - k0);
    __syncthreads();
    for (int k = threadIdx.x; k < tile; k += blockDim.x) {
      const float* p = kn + (size_t)(k0 + k) * 3;
      kt[k] = make_float4(-__ldg(p), -__ldg(p + 1), -__ldg(p + 2), 0.f);
    }
    __syncthreads();
#pragma unroll 4
    for (int k = 0; k < tile; ++k) {
      const float4 q = kt[k];
      const float2 dx = __fadd2_rn(ux, make_float2(q.x, q.x)), dy = __fadd2_rn(uy, make_float2(q.y, q.y)),
                   dz = __fadd2_rn(uz, make_float2(q.z, q.z));
      const float2 d = __ffma2_rn(dz, dz, __ffma2_rn(dx, dx, __fmul2_rn(dy, dy)));
      if (d.x < a3) nn3_insert(d.x, k0 + k, a1, a2, a3, ia1, ia2, ia3);
      if (d.y < c3) nn3_insert(d.y, k0 + k, c1, c2, c3, ic1, ic2, ic3);
    }
  }
  if (j0 < n) {
    float* d = dist2 + ((size_t)b * n + j0) * 3;
    int* o = idx + ((size_t)b * n + j0) * 3;
    d[0] = a1; d[1] = a2; d[2] = a3;
    o[0] = ia1; o[1] = ia2; o[2] = ia3;
  }
  if (j1 < n) {
    float* d = dist2 + ((size_t)b * n + j1) * 3;
    int* o = idx + ((size_t)b * n + j1) * 3;
    d[0] = c1; d[1] = c2; d[2] = c3;
    o[0] = ic1; o[1] = ic2; o[2] = ic3;
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
constexpr int TI_THREADS = 256;
template <int TI_PPT, int TI_STAGE>  // TI_STAGE float4 slots staged per thread and quad: m <= TI_THREADS * TI_STAGE
__global__ void __launch_bounds__(TI_THREADS)
three_interpolate_c4_kernel(int c, int m, int n, int quads_per_block, const float* __restrict__ points,
                            const int* __restrict__ idx, const float* __restrict__ weight, float* __restrict__ out) {
  extern __shared__ float4 ti_tile[];  // [2][m]
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
  // The source tile is DOUBLE-BUFFERED: the 4 rows of quad q+1 are fetched into registers before quad q is computed and stored to
  // the other buffer after it, so the global-load latency of the staging overlaps the gathers / stores and a quad costs one
  // barrier instead of two (the kernel is HBM-bound on its output stream; the barriers were its bubbles).
  float4 st[TI_STAGE];
  auto fetch = [&](int q) {
    const int l0 = q * 4, nl = min(4, c - l0);
    const float* r0 = pts + (size_t)l0 * m;
    const float* r1 = r0 + (nl > 1 ? m : 0);
    const float* r2 = r0 + (nl > 2 ? 2 * (size_t)m : 0);
    const float* r3 = r0 + (nl > 3 ? 3 * (size_t)m : 0);
#pragma unroll
    for (int s = 0; s < TI_STAGE; ++s) {
      const int i = tid + s * TI_THREADS;
      if (i < m) st[s] = make_float4(__ldg(r0 + i), __ldg(r1 + i), __ldg(r2 + i), __ldg(r3 + i));
    }
  };
  auto stash = [&](float4* buf) {
#pragma unroll
    for (int s = 0; s < TI_STAGE; ++s) {
      const int i = tid + s * TI_THREADS;
      if (i < m) buf[i] = st[s];
    }
  };
  fetch(q0);
  stash(ti_tile);
  __syncthreads();
  for (int q = q0; q < q1; ++q) {
    const int l0 = q * 4, nl = min(4, c - l0);
    const float4* cur = ti_tile + ((q - q0) & 1) * m;
    float4* nxt = ti_tile + (((q - q0) & 1) ^ 1) * m;
    if (q + 1 < q1) fetch(q + 1);
    float* o0 = o + (size_t)l0 * n;
#pragma unroll
    for (int u = 0; u < TI_PPT; ++u) {
      if (!ok[u]) continue;
      const int j = (blockIdx.x * TI_PPT + u) * TI_THREADS + tid;
      const float4 a = cur[i1[u]], bq = cur[i2[u]], cq = cur[i3[u]];
      // fma(p3,w3, fma(p1,w1, mul(p2,w2))) per channel, as in the plain kernel
      __stcs(o0 + j, __fmaf_rn(cq.x, w3[u], __fmaf_rn(a.x, w1[u], __fmul_rn(bq.x, w2[u]))));
      if (nl > 1) __stcs(o0 + (size_t)n + j, __fmaf_rn(cq.y, w3[u], __fmaf_rn(a.y, w1[u], __fmul_rn(bq.y, w2[u]))));
      if (nl > 2) __stcs(o0 + 2 * (size_t)n + j, __fmaf_rn(cq.z, w3[u], __fmaf_rn(a.z, w1[u], __fmul_rn(bq.z, w2[u]))));
      if (nl > 3) __stcs(o0 + 3 * (size_t)n + j, __fmaf_rn(cq.w, w3[u], __fmaf_rn(a.w, w1[u], __fmul_rn(bq.w, w2[u]))));
    }
    if (q + 1 < q1) stash(nxt);  // the other buffer was last read in iteration q-1, before that iteration's barrier
    __syncthreads();
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
  // fast path: one thread block of exactly the reference's block size (see fps_fast_kernel)
#define FPS_FAST(PPT, THREADS) return launch_fps_fast<PPT, THREADS>(b, n, m, log2bs, qbits, dataset, temp, idxs, st)
  // threads per cloud = 512 >> S.  Measured 2048 -> 512 at batch 32 / 256: S=0 186 / 304 us, S=1 146 / 204 us, S=2 175 / 204 us
  // (DFB200_FPS_S overrides for A/B measurements)
  static const int fps_s = [] { const char* e = getenv("DFB200_FPS_S"); return e != nullptr ? atoi(e) : 1; }();
#define FPS_WIDE(PPT, THREADS, S) return launch_fps_wide<PPT, THREADS, S>(b, n, m, qbits, dataset, temp, idxs, st)
  if (bs == 512 && fps_s == 1 && q >= 1 && q <= 8) {
    if (qbits == 0) FPS_WIDE(2, 256, 1);
    if (qbits == 1) FPS_WIDE(4, 256, 1);
    if (qbits == 2) FPS_WIDE(8, 256, 1);
    if (qbits == 3) FPS_WIDE(16, 256, 1);
  }
  if (bs == 512 && fps_s == 2 && q >= 1 && q <= 4) {
    if (qbits == 0) FPS_WIDE(4, 128, 2);
    if (qbits == 1) FPS_WIDE(8, 128, 2);
    if (qbits == 2) FPS_WIDE(16, 128, 2);
  }
#undef FPS_WIDE
  if (bs == 512) {
    if (q <= 1) FPS_FAST(1, 512);
    if (q <= 2) FPS_FAST(2, 512);
    if (q <= 4) FPS_FAST(4, 512);
    if (q <= 8) FPS_FAST(8, 512);
    if (q <= 16) FPS_FAST(16, 512);
  } else if (q <= 2) {
    if (bs == 256) FPS_FAST(2, 256);
    if (bs == 128) FPS_FAST(2, 128);
    if (bs == 64) FPS_FAST(2, 64);
    if (bs == 32) FPS_FAST(2, 32);
  }
#undef FPS_FAST
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
  // sparse-ball path: uniform grid + 8 lanes per centre (see ball_query_grid_kernel), then the ordered scan for handed-over clouds;
  // DFB200_BALL_QUERY=scan forces the ordered brute-force scan (A/B measurements)
  static const bool force_scan = [] { const char* e = getenv("DFB200_BALL_QUERY"); return e != nullptr && e[0] == 's'; }();
  // thread-per-centre scan (ball_query_tpc_kernel): DFB200_BALL_QUERY=tpc forces it, =grid forces the grid path (A/B measurements)
  static const int tpc_mode = [] { const char* e = getenv("DFB200_BALL_QUERY"); return e == nullptr ? 0 : e[0] == 't' ? 1 : e[0] == 'g' ? -1 : 0; }();
  {
    // centres per thread (DFB200_BQT_C, A/B): 1 is the measured optimum at batch 256 (104 / 118 / 138 us for r = 0.1 / 0.2 / 0.4; 2:
    // 117 / 127 / -, 4 with 128 threads: 145 / 174 / -: the register tile halves the broadcast-load traffic that bounds the
    // kernel, but one cloud per CTA leaves 1.7 CTAs per SM)
    static const int tpc_c = [] { const char* e = getenv("DFB200_BQT_C"); return e != nullptr ? atoi(e) : 1; }();
    const int C = tpc_c == 4 ? 4 : tpc_c == 2 ? 2 : 1;
    const size_t tsm = bqt_smem_bytes(n, nsample, C);
    const bool fits = n >= 1 && n <= 65535 && tsm <= 110 * 1024;
    // default policy: nsample is the caller's estimate of the ball population.  Balls expected to hold >= 48 points go to the
    // thread-per-centre scan (radius-independent: 118 vs 159 us at r = 0.2, 138 vs 190 us at r = 0.4 against the grid path, same
    // inputs); sparser queries keep the grid path, whose pruning wins there (97 vs 104 us at r = 0.1, nsample 16).
    const bool want = tpc_mode == 1 || (tpc_mode == 0 && !force_scan && nsample >= 48 && n >= 256 && m >= 64 && (long long)b * m >= 148LL * BQT_THREADS);  // at least one CTA per SM
    if (fits && want) {
      static DeviceOnce once;
      if (once.first_time()) {
        DFB_CUDA(cudaFuncSetAttribute(ball_query_tpc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        DFB_CUDA(cudaFuncSetAttribute(ball_query_tpc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        DFB_CUDA(cudaFuncSetAttribute(ball_query_tpc_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
      }
      const dim3 grid(cdiv(m, BQT_THREADS * C), b);
      if (C == 1) ball_query_tpc_kernel<1><<<grid, BQT_THREADS, tsm, st>>>(n, m, radius2, nsample, new_xyz, xyz, idx);
      else if (C == 2) ball_query_tpc_kernel<2><<<grid, BQT_THREADS, tsm, st>>>(n, m, radius2, nsample, new_xyz, xyz, idx);
      else ball_query_tpc_kernel<4><<<grid, BQT_THREADS, tsm, st>>>(n, m, radius2, nsample, new_xyz, xyz, idx);
      DFB_LAUNCH_CHECK();
      return DFB200_OK;
    }
  }
  if (!force_scan && n >= 1024 && n <= 8192 && m >= 32) {
    const size_t gsm = bqg_smem_bytes(n);
    // persistent CTAs (as many as fit: shared memory or 8 x 256 threads per SM), balanced static partition
    // DFB200_BQ_QUERY=wpc: the warp-per-centre query over a flattened candidate stream (round-2 experiment, measured SLOWER than
    // the default 8-lanes-per-centre query: 180 vs 132 us at batch 256, r = 0.2 -- 804 vs 633 warp-instructions per centre, the
    // per-lane run cursor and the 32-lane read-back cost more than the idle lanes they remove; profiles/ncu_ball_query_r2.csv)
    static const bool wpc = [] { const char* e = getenv("DFB200_BQ_QUERY"); return e != nullptr && e[0] == 'w'; }();
    auto kernel = wpc ? ball_query_grid_kernel<true> : ball_query_grid_kernel<false>;
    DFB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsm));
    int per_sm = 0;  // resident CTAs per SM (registers AND shared memory): more CTAs than that would run as a second, unbalanced wave
    DFB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, BQG_T, gsm));
    per_sm = per_sm > 8 ? 8 : (per_sm < 1 ? 1 : per_sm);
    const int ppc = cdiv(m, BQG_T / BQG_LPC);  // passes of 32 centres per cloud
    const int n_sm = current_device_sm_count();
    const long long items = (long long)b * ppc, slots = (long long)(n_sm > 0 ? n_sm : 148) * per_sm;
    int ctas;
    if (2LL * b <= slots) {  // few clouds: k CTAs per cloud (slices never straddle two clouds -> one build each)
      const int k = (int)(slots / b < ppc ? slots / b : ppc);
      ctas = b * k;
    } else {
      ctas = (int)(items < slots ? items : slots);
    }
    int scpw = 8;
    while (scpw > 1 && (long long)b * cdiv(m, 8 * scpw) < 148 * 4) scpw /= 2;
    int mpad = 0;  // centres are sorted by cell when they fit the key array
    if (m <= BQG_MAXSORT) {
      mpad = 32;
      while (mpad < m) mpad <<= 1;
    }
    static const int min_occ = [] { const char* e = getenv("DFB200_BQ_MIN_OCCUPIED"); return e != nullptr ? atoi(e) : BQG_MIN_OCCUPIED; }();
    if (wpc) mpad = 0;  // no centre sorting: a warp works on one centre at a time
    kernel<<<ctas, BQG_T, gsm, st>>>(b, n, m, mpad, (int)bqg_keys_offset(n), min_occ, radius, radius2, nsample, 8 * scpw, new_xyz, xyz, idx);
    DFB_LAUNCH_CHECK();
    const size_t ssm = sizeof(float) * 3 * (size_t)((n + 3) & ~3);
    if (ssm > 48 * 1024)
      DFB_CUDA(cudaFuncSetAttribute(ball_query_scan4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));
    ball_query_scan4_kernel<<<dim3(cdiv(m, 8 * scpw), b), 256, ssm, st>>>(n, m, radius2, nsample, scpw, 1, new_xyz, xyz, idx);
    DFB_LAUNCH_CHECK();
    return DFB200_OK;
  }
  // centres per warp: keep >= ~4 waves of CTAs on 148 SMs but amortise the smem staging
  int cpw = 8;
  while (cpw > 1 && (long long)b * cdiv(m, 8 * cpw) < 148 * 4) cpw /= 2;
  dim3 grid(cdiv(m, 8 * cpw), b);
  const size_t smem = sizeof(float) * 3 * (size_t)((n + 3) & ~3);
  static const bool scan1 = [] { const char* e = getenv("DFB200_BALL_QUERY"); return e != nullptr && e[0] == 's' && e[1] == '1'; }();  // A/B
  if (smem <= 160 * 1024 && !scan1) {
    if (smem > 48 * 1024)
      DFB_CUDA(cudaFuncSetAttribute(ball_query_scan4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ball_query_scan4_kernel<<<grid, 256, smem, st>>>(n, m, radius2, nsample, cpw, 0, new_xyz, xyz, idx);
  } else if (smem <= 160 * 1024) {
    if (smem > 48 * 1024)
      DFB_CUDA(cudaFuncSetAttribute(ball_query_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ball_query_kernel<true><<<grid, 256, smem, st>>>(n, m, radius2, nsample, cpw, 0, new_xyz, xyz, idx);
  } else {
    ball_query_kernel<false><<<grid, 256, 0, st>>>(n, m, radius2, nsample, cpw, 0, new_xyz, xyz, idx);
  }
  DFB_LAUNCH_CHECK();
  return DFB200_OK;
}

extern "C" int dfb200_three_nn(int b, int n, int m, const float* unknown, const float* known,
                               float* dist2, int* idx, dfb200_stream_t stream) {
  DFB_REQUIRE(b >= 0 && n >= 0 && m >= 0, DFB200_ERR_INVALID_ARG, "three_nn: negative size");
  if (b == 0 || n == 0) return DFB200_OK;
  DFB_REQUIRE(b <= 65535, DFB200_ERR_INVALID_ARG, "three_nn: b > 65535");
  static const bool legacy = [] { const char* e = getenv("DFB200_THREE_NN"); return e != nullptr && e[0] == '1'; }();  // A/B
  if (!legacy && (long long)b * cdiv(n, 128) >= 148 * 2) {  // enough work for two unknown points per thread
    const int threads = (long long)b * cdiv(n, 512) >= 148 * 2 ? 256 : 64;
    three_nn2_kernel<<<dim3(cdiv(n, 2 * threads), b), threads, 0, as_stream(stream)>>>(n, m, unknown, known, dist2, idx);
    DFB_LAUNCH_CHECK();
    return DFB200_OK;
  }
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
    const size_t smem = 2 * sizeof(float4) * (size_t)m;  // double-buffered source tile
    if (m <= TI_THREADS * 4 && n >= 4 * m && c >= 4 && b <= 65535) {
      const int ppt = (n >= TI_THREADS * 8 && getenv("DFB200_TI_PPT4") == nullptr) ? 8 : 4;  // points per thread (idx / weights in registers)
      const int gx4 = cdiv(n, TI_THREADS * ppt);
      const int nquad = (c + 3) / 4;
      int qpb = nquad;
      while (qpb > 1 && (long long)gx4 * b * cdiv(nquad, qpb) < 148 * 8) qpb = (qpb + 1) / 2;
      dim3 grid4(gx4, cdiv(nquad, qpb), b);
      if (grid4.y <= 65535) {
        // (a variant with 4 consecutive points per thread -- 128-bit idx/weight loads, one STG.128 per channel -- was
        //  measured slower: 224 vs 191 us at batch 256, 69 registers and less gather parallelism per warp)
#define TI_LAUNCH(P, S) three_interpolate_c4_kernel<P, S><<<grid4, TI_THREADS, smem, as_stream(stream)>>>(c, m, n, qpb, points, idx, weight, out)
        if (ppt == 8 && m <= TI_THREADS * 2) TI_LAUNCH(8, 2);
        else if (ppt == 8) TI_LAUNCH(8, 4);
        else if (m <= TI_THREADS * 2) TI_LAUNCH(4, 2);
        else TI_LAUNCH(4, 4);
#undef TI_LAUNCH
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
