// Reverse-DDPM sampling loop (x_T -> x_0) behind one C-ABI call.
// Reference: python/difffacto/models/diffusions/anchored_diffusion.py:528-588
// (p_sample_loop_progressive) driven by AnchorDiffAE.decode, models/networks/anchor_gen.py:145-169.
// The reference crosses host->device ~11 times per step (torch.tensor([i]*B) + 10 schedule-table
// uploads) and launches ~180 kernels per step; here the schedule lives on the device, the step
// index is a kernel argument, and a step is the denoiser launch sequence + one fused update (fp32
// mode) or a single fused kernel (bf16 mode).  The loop is stream-ordered: no host sync inside.
#include <stdlib.h>

#include "ddpm.cuh"
#include "denoiser.cuh"

namespace dfb200 {

int launch_ddpm_step(int B, int N, int T, const float* sched, const int* t, const float* x_t,
                     const float* eps, const float* anchors, const float* variance, const float* noise,
                     bool philox, uint64_t seed, uint64_t offset, float* x_prev, float* pred_xstart,
                     cudaStream_t st);
int launch_xT_init(long long total, float* x, const float* anchors, const float* variance, bool philox,
                   uint64_t seed, uint64_t offset, cudaStream_t st);

__global__ void fill_step_kernel(int B, int i, float* __restrict__ t_f, int* __restrict__ t_i) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) {
    t_f[b] = (float)i;
    t_i[b] = i;
  }
}

struct LoopWorkspace {
  float* eps;   // [B,3,N]              (fp32 mode)
  float* eps_u; // [B,3,N]              (fp32 mode, guidance: unconditional pass)
  float* t_f;   // [B]                  (fp32 mode)
  int* t_i;     // [B]                  (fp32 mode)
  void* net;    // denoiser workspace   (fp32 mode)
  float* ctx0;  // [B,context_dim,n_tok] zeros: context of the unconditional pass (guidance)
  // bf16 mode: time / sample tables built once per loop, fold tiles built per chunk of steps
  float* t_all;      // [T] timestep values 0..T-1
  float* temb_h;     // [T,1024]
  float* temb;       // [T,256]
  float* kv_time;    // [T,depth,2,128]
  float* kv_static;  // [B (x2 with guidance),depth,2,4,128]
  void* fold;        // [chunk][B (x2)][depth] fold packets
  int* done;         // [units] cross-step dependency counters of the persistent kernel
  int chunk;         // sampling steps per fold launch / persistent kernel launch
  size_t net_bytes;
  size_t bytes;
};

// steps after which a round-robin walk of the (step, unit) list has given every SM the same number of items
static long long balance_period(long long units) {
  int n_sm = current_device_sm_count();
  if (n_sm <= 0) n_sm = 148;  // sizing query on a box without a GPU
  long long a = units, b = n_sm;
  while (b) { const long long t = a % b; a = b; b = t; }
  return a > 0 ? n_sm / a : 1;
}

static LoopWorkspace carve_loop(const NetDims& d, int mode, int B, int N, int T, void* base) {
  LoopWorkspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? reinterpret_cast<char*>(base) + off : nullptr;
    off += (bytes + 255) & ~(size_t)255;
    return p;
  };
  // sized for the largest variant (guidance doubles the per-sample tables), so one workspace serves every option set
  w.ctx0 = reinterpret_cast<float*>(take(sizeof(float) * (size_t)B * (d.c_ctx_static - d.n_tok) * d.n_tok));
  if (mode != DFB200_MODE_BF16) {
    w.eps = reinterpret_cast<float*>(take(sizeof(float) * (size_t)B * 3 * N));
    w.eps_u = reinterpret_cast<float*>(take(sizeof(float) * (size_t)B * 3 * N));
    w.t_f = reinterpret_cast<float*>(take(sizeof(float) * (size_t)B));
    w.t_i = reinterpret_cast<int*>(take(sizeof(int) * (size_t)B));
    w.net_bytes = carve_workspace(d, mode, B, N, nullptr).bytes;
    w.net = take(w.net_bytes);
  } else {
    w.t_all = reinterpret_cast<float*>(take(sizeof(float) * (size_t)T));
    w.temb_h = reinterpret_cast<float*>(take(sizeof(float) * (size_t)T * D_TEMB_H));
    w.temb = reinterpret_cast<float*>(take(sizeof(float) * (size_t)T * D_TEMB));
    w.kv_time = reinterpret_cast<float*>(take(sizeof(float) * (size_t)T * d.depth * 2 * D_MODEL));
    w.kv_static = reinterpret_cast<float*>(take(sizeof(float) * (size_t)2 * B * d.depth * 2 * d.n_tok * D_MODEL));
    // Fold tiles (per step: B samples x depth blocks x 16.5 KB) are built for a whole chunk of steps by ONE launch, and ONE persistent
    // launch of the fused kernel runs the chunk.  Every chunk boundary costs the fold launch plus the drain / ramp of the work list
    // (measured at the BASELINE size with 24-step chunks: 0.33 ms per chunk, 10 % of the loop), so the chunk is as long as the
    // fold workspace allows -- 4 GiB by default (DFB200_FOLD_WORKSPACE_MB), i.e. all 1000 steps of a 32-shape batch in one
    // launch -- and, when the loop still needs several chunks, a multiple of the step count after which every SM has run the same
    // number of (step, unit) items (n_sm / gcd(units, n_sm) steps: 37 for 256 units on 148 SMs).
    const size_t per_step = tc_fold_bytes_for(d, B);
    size_t cap_mb = 4096;
    if (const char* e = getenv("DFB200_FOLD_WORKSPACE_MB")) {
      const long v = atol(e);
      if (v > 0) cap_mb = (size_t)v;
    }
    long long chunk = per_step ? (long long)((cap_mb << 20) / per_step) : 1;
    if (chunk < 1) chunk = 1;
    if (chunk > T) chunk = T;
    if (chunk < T) {
      const long long period = balance_period(cdiv((long long)B * N, 256));
      if (chunk >= period) chunk -= chunk % period;
    }
    w.chunk = (int)chunk;
    w.fold = take(per_step * (size_t)(chunk > 1 ? chunk : 2));  // >= 2 steps' worth: one guidance step needs 2B entries
    w.done = reinterpret_cast<int*>(take(sizeof(int) * (size_t)cdiv((long long)B * N, 128)));
  }
  w.bytes = off;
  return w;
}

__global__ void arange_kernel(int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)i;
}

}  // namespace dfb200

using namespace dfb200;

extern "C" size_t dfb200_ddpm_sample_loop_workspace_bytes(const dfb200_denoiser_cfg* cfg, int mode, int B, int N, int T) {
  NetDims d;
  if (make_net_dims(cfg, &d) != DFB200_OK || B < 0 || N < 0 || T < 1) return 0;
  return carve_loop(d, mode, B, N, T, nullptr).bytes;
}

extern "C" int dfb200_sample_loop_chunk(const dfb200_denoiser_cfg* cfg, int mode, int B, int N, int T) {
  NetDims d;
  if (make_net_dims(cfg, &d) != DFB200_OK || B < 0 || N < 0 || T < 1) return 0;
  if (mode != DFB200_MODE_BF16) return T < 32 ? T : 32;  // stepwise launches: any chunk is as good as another
  // a generator-style caller keeps `sample` + `pred_xstart` of every step of a chunk alive: the shortest balanced chunk of >= 24 steps
  const long long full = carve_loop(d, mode, B, N, T, nullptr).chunk;
  const long long period = balance_period(cdiv((long long)B * N, 256));
  long long c = period * ((24 + period - 1) / period);
  if (c > TC_MAX_LIST_STEPS) c = TC_MAX_LIST_STEPS;  // per-step output buffers: at most this many steps per launch
  if (c > full) c = full;
  return (int)c;
}

extern "C" int dfb200_sample_loop(const dfb200_denoiser_cfg* cfg, const void* packed, int mode, int B, int N, int T,
                                  const float* sched, float* x, int x_T_from_noise, const float* ctx, const float* anchors,
                                  const float* variance, const int* anchor_assignment, const float* valid_id,
                                  const float* noise, uint64_t seed, float* traj, int traj_interval,
                                  const dfb200_sample_opts* opts, void* workspace, size_t workspace_bytes,
                                  dfb200_stream_t stream) {
  PackLayout L;
  int rc = make_pack_layout(cfg, &L);
  if (rc != DFB200_OK) return rc;
  dfb200_sample_opts o{};
  if (opts != nullptr) o = *opts;
  DFB_REQUIRE(B >= 0 && N >= 0 && T >= 1, DFB200_ERR_INVALID_ARG, "sample_loop: bad sizes B=%d N=%d T=%d", B, N, T);
  DFB_REQUIRE(mode == DFB200_MODE_FP32 || mode == DFB200_MODE_BF16 || mode == DFB200_MODE_TF32, DFB200_ERR_INVALID_ARG,
              "sample_loop: unknown mode %d", mode);
  DFB_REQUIRE(traj == nullptr || traj_interval >= 1, DFB200_ERR_INVALID_ARG, "sample_loop: traj_interval must be >= 1");
  // the step list: every timestep T-1 .. 0, or the caller's strictly decreasing device list (DDIM strides)
  const int n_list = o.timesteps != nullptr ? o.n_timesteps : T;
  DFB_REQUIRE(o.timesteps == nullptr || (o.timesteps_host != nullptr && o.n_timesteps >= 1), DFB200_ERR_INVALID_ARG,
              "sample_loop: a timestep list needs its host copy and n_timesteps >= 1");
  for (int k = 0; o.timesteps != nullptr && k < n_list; ++k)
    DFB_REQUIRE(o.timesteps_host[k] >= 0 && o.timesteps_host[k] < T && (k == 0 || o.timesteps_host[k] < o.timesteps_host[k - 1]),
                DFB200_ERR_INVALID_ARG, "sample_loop: timesteps must be strictly decreasing values in [0, T)");
  const int first = o.first_step, count = o.num_steps > 0 ? o.num_steps : n_list - first;
  DFB_REQUIRE(first >= 0 && count >= 0 && first + count <= n_list, DFB200_ERR_INVALID_ARG,
              "sample_loop: steps [%d, %d) outside the list of %d", first, first + count, n_list);
  DFB_REQUIRE(!o.ddim || (o.alphas_cumprod_prev != nullptr && o.xt_dir_coeff != nullptr), DFB200_ERR_INVALID_ARG,
              "sample_loop: the DDIM update needs the alphas_cumprod_prev and xt_dir_coeff tables");
  if (B == 0 || N == 0) return DFB200_OK;
  LoopWorkspace lw = carve_loop(L.d, mode, B, N, T, workspace);
  DFB_REQUIRE(workspace != nullptr && workspace_bytes >= lw.bytes, DFB200_ERR_WORKSPACE,
              "sample_loop: workspace too small (%zu < %zu)", workspace_bytes, lw.bytes);
  cudaStream_t st = as_stream(stream);
  const float* P = reinterpret_cast<const float*>(packed);
  const float* valid = (L.d.flags & DFB200_NET_MASK_UNREFERENCED) ? valid_id : nullptr;
  const long long total = (long long)B * 3 * N;
  const bool philox = noise == nullptr;
  const bool prepare = first == 0 || !o.tables_ready;
  auto step_time = [&](int k) { return o.timesteps != nullptr ? o.timesteps_host[k] : T - 1 - k; };

  // x_T = sqrt(var) * z + anchors  (anchored_diffusion.py:564); Philox draw `T` is the x_T noise
  // x_T_from_noise: 0 = x already holds x_T, 1 = x holds N(0,1) noise, 2 = draw it from Philox
  if (x_T_from_noise && first == 0) {
    rc = launch_xT_init(total, x, anchors, variance, x_T_from_noise == 2, seed, (uint64_t)T, st);
    if (rc != DFB200_OK) return rc;
  }
  // AnchorDiffAE.decode keeps x_T itself under key T when T % ret_interval == 0 (anchor_gen.py:164-165): last slot
  if (traj != nullptr && first == 0 && T % traj_interval == 0)
    DFB_CUDA(cudaMemcpyAsync(traj + (size_t)(T / traj_interval - 1) * total, x, sizeof(float) * total, cudaMemcpyDeviceToDevice, st));
  const size_t ctx_floats = (size_t)B * (L.d.c_ctx_static - L.d.n_tok) * L.d.n_tok;
  if (o.guidance && prepare) DFB_CUDA(cudaMemsetAsync(lw.ctx0, 0, sizeof(float) * ctx_floats, st));

  if (mode != DFB200_MODE_BF16) {  // fp32 (CUDA cores) and tf32 (tensor cores): step-wise launches, same update kernels
    Workspace ws = carve_workspace(L.d, mode, B, N, lw.net);
    auto forward = [&](float* eps_dst) {
      return mode == DFB200_MODE_TF32
                 ? denoiser_forward_tf32(L, packed, B, N, x, anchors, variance, anchor_assignment, valid, eps_dst, ws, st)
                 : denoiser_forward_fp32(L, P, B, N, x, anchors, variance, anchor_assignment, valid, eps_dst, ws, st);
    };
    for (int k = first; k < first + count; ++k) {
      const int i = step_time(k);
      fill_step_kernel<<<cdiv(B, 256), 256, 0, st>>>(B, i, lw.t_f, lw.t_i);
      DFB_LAUNCH_CHECK();
      const float* eps = lw.eps;
      if (o.guidance) {  // unconditional pass first (its K/V workspace is overwritten by the conditional one)
        rc = launch_context_kv(L, P, B, lw.t_f, lw.ctx0, ws, st);
        if (rc != DFB200_OK) return rc;
        rc = forward(lw.eps_u);
        if (rc != DFB200_OK) return rc;
      }
      rc = launch_context_kv(L, P, B, lw.t_f, ctx, ws, st);
      if (rc != DFB200_OK) return rc;
      rc = forward(lw.eps);
      if (rc != DFB200_OK) return rc;
      if (o.guidance) {
        rc = dfb200_guidance_mix((size_t)total, o.classifier_weight, lw.eps_u, lw.eps, lw.eps, stream);
        if (rc != DFB200_OK) return rc;
      }
      const float* z = philox ? nullptr : noise + (size_t)(k - first) * total;
      float* xs = o.step_xstart != nullptr ? o.step_xstart + (size_t)(k - first) * total
                                           : o.step_xstart_list != nullptr ? o.step_xstart_list[k - first] : nullptr;
      if (o.ddim) {
        float* zbuf = nullptr;
        if (philox) {  // DDIM with in-kernel noise: draw this step's Philox normals into the (now free) unconditional buffer
          zbuf = lw.eps_u;
          rc = dfb200_philox_normal(zbuf, (size_t)total, seed, (uint64_t)i, stream);
          if (rc != DFB200_OK) return rc;
        }
        rc = dfb200_ddim_step(B, N, T, sched, lw.t_i, x, eps, anchors, variance, philox ? zbuf : z, o.alphas_cumprod_prev,
                              o.xt_dir_coeff, o.ddim_eta, x, xs, stream);
      } else {
        rc = launch_ddpm_step(B, N, T, sched, lw.t_i, x, eps, anchors, variance, z, philox, seed, (uint64_t)i, x, xs, st);
      }
      if (rc != DFB200_OK) return rc;
      float* ssd = o.step_sample != nullptr ? o.step_sample + (size_t)(k - first) * total
                                            : o.step_sample_list != nullptr ? o.step_sample_list[k - first] : nullptr;
      if (ssd != nullptr) DFB_CUDA(cudaMemcpyAsync(ssd, x, sizeof(float) * total, cudaMemcpyDeviceToDevice, st));
      if (traj != nullptr && i > 0 && i % traj_interval == 0)
        DFB_CUDA(cudaMemcpyAsync(traj + (size_t)(i / traj_interval - 1) * total, x, sizeof(float) * total, cudaMemcpyDeviceToDevice, st));
    }
    return DFB200_OK;
  }

  // ---- bf16 mode: everything that does not depend on x is hoisted out of the step loop ----
  //   time tables for all T steps (timestep MLP + time half of K/V), static half of K/V per sample, then per chunk of steps
  //   ONE fold launch (K/V -> attention weight tiles) and ONE persistent launch of the fused kernel (denoiser + update).
  const int nb = o.guidance ? 2 * B : B;  // guidance: entries [B, 2B) are the unconditional (zero-context) samples
  if (prepare) {
    arange_kernel<<<cdiv(T, 256), 256, 0, st>>>(T, lw.t_all);
    DFB_LAUNCH_CHECK();
    rc = launch_context_kv_time(L, P, T, lw.t_all, lw.temb_h, lw.temb, lw.kv_time, st);
    if (rc != DFB200_OK) return rc;
    rc = launch_context_kv_static(L, P, B, ctx, lw.kv_static, st);
    if (rc != DFB200_OK) return rc;
    if (o.guidance) {
      rc = launch_context_kv_static(L, P, B, lw.ctx0, lw.kv_static + (size_t)B * L.d.depth * 2 * L.d.n_tok * D_MODEL, st);
      if (rc != DFB200_OK) return rc;
    }
    DFB_CUDA(cudaMemsetAsync(lw.done, 0, sizeof(int) * (size_t)cdiv((long long)B * N, 128), st));
  }
  const size_t per_step = tc_fold_bytes_for(L.d, nb);
  int chunk = o.guidance ? (lw.chunk > 1 ? lw.chunk / 2 : 1) : lw.chunk;  // guidance: 2B fold entries per step
  const bool lists = o.step_sample_list != nullptr || o.step_xstart_list != nullptr;
  if (lists && chunk > TC_MAX_LIST_STEPS) chunk = TC_MAX_LIST_STEPS;  // per-step output pointers travel in the kernel parameters
  for (int k0 = first; k0 < first + count; k0 += chunk) {
    const int steps = first + count - k0 < chunk ? first + count - k0 : chunk;
    const int* list = o.timesteps != nullptr ? o.timesteps + k0 : nullptr;
    rc = launch_context_fold(L, packed, nb, lw.kv_static, lw.kv_time, step_time(k0), list, steps, lw.fold, st);
    if (rc != DFB200_OK) return rc;
    // ONE persistent launch runs `steps` timesteps for every unit (148 CTAs walk the (step, unit) list; a unit's next step
    // waits on its previous one through lw.done), so no SM idles at step boundaries.
    TcUpdate u{};
    u.sched = sched; u.T = T; u.t = step_time(k0); u.n_steps = steps; u.fold_step_bytes = per_step;
    u.step_t = list; u.step_base = k0;
    u.noise = philox ? nullptr : noise + (size_t)(k0 - first) * total;
    u.seed = seed;
    u.x_out = x;
    u.done = lw.done;
    u.traj = traj; u.traj_interval = traj_interval;
    u.step_sample = o.step_sample != nullptr ? o.step_sample + (size_t)(k0 - first) * total : nullptr;
    u.step_xstart = o.step_xstart != nullptr ? o.step_xstart + (size_t)(k0 - first) * total : nullptr;
    u.step_sample_list = o.step_sample_list != nullptr ? o.step_sample_list + (k0 - first) : nullptr;
    u.step_xstart_list = o.step_xstart_list != nullptr ? o.step_xstart_list + (k0 - first) : nullptr;
    if (o.ddim) { u.ddim_acp = o.alphas_cumprod_prev; u.ddim_dir = o.xt_dir_coeff; u.ddim_eta = o.ddim_eta; }
    u.guidance = o.guidance; u.guid_w = o.classifier_weight;
    rc = denoiser_step_tc(L, packed, B, N, x, anchors, variance, anchor_assignment, valid, lw.fold, nullptr, &u, st);
    if (rc != DFB200_OK) return rc;
  }
  return DFB200_OK;
}

extern "C" int dfb200_ddpm_sample_loop(const dfb200_denoiser_cfg* cfg, const void* packed, int mode, int B, int N,
                                       int T, const float* sched, float* x, int x_T_from_noise, const float* ctx,
                                       const float* anchors, const float* variance, const int* anchor_assignment,
                                       const float* valid_id, const float* noise, uint64_t seed, float* traj,
                                       int traj_interval, void* workspace, size_t workspace_bytes,
                                       dfb200_stream_t stream) {
  return dfb200_sample_loop(cfg, packed, mode, B, N, T, sched, x, x_T_from_noise, ctx, anchors, variance, anchor_assignment,
                            valid_id, noise, seed, traj, traj_interval, nullptr, workspace, workspace_bytes, stream);
}
