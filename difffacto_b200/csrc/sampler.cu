// Reverse-DDPM sampling loop (x_T -> x_0) behind one C-ABI call.
// Reference: python/difffacto/models/diffusions/anchored_diffusion.py:528-588
// (p_sample_loop_progressive) driven by AnchorDiffAE.decode, models/networks/anchor_gen.py:145-169.
// The reference crosses host->device ~11 times per step (torch.tensor([i]*B) + 10 schedule-table
// uploads) and launches ~180 kernels per step; here the schedule lives on the device, the step
// index is a kernel argument, and a step is the denoiser launch sequence + one fused update (fp32
// mode) or a single fused kernel (bf16 mode).  The loop is stream-ordered: no host sync inside.
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
  float* t_f;   // [B]                  (fp32 mode)
  int* t_i;     // [B]                  (fp32 mode)
  void* net;    // denoiser workspace   (fp32 mode)
  // bf16 mode: time / sample tables built once per loop, fold tiles built per chunk of steps
  float* t_all;      // [T] timestep values 0..T-1
  float* temb_h;     // [T,1024]
  float* temb;       // [T,256]
  float* kv_time;    // [T,depth,2,128]
  float* kv_static;  // [B,depth,2,4,128]
  void* fold;        // [chunk][B][depth] fold packets
  int* done;         // [units] cross-step dependency counters of the persistent kernel
  int chunk;         // sampling steps per fold launch / persistent kernel launch
  size_t net_bytes;
  size_t bytes;
};

static LoopWorkspace carve_loop(const NetDims& d, int mode, int B, int N, int T, void* base) {
  LoopWorkspace w{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? reinterpret_cast<char*>(base) + off : nullptr;
    off += (bytes + 255) & ~(size_t)255;
    return p;
  };
  if (mode == DFB200_MODE_FP32) {
    w.eps = reinterpret_cast<float*>(take(sizeof(float) * (size_t)B * 3 * N));
    w.t_f = reinterpret_cast<float*>(take(sizeof(float) * (size_t)B));
    w.t_i = reinterpret_cast<int*>(take(sizeof(int) * (size_t)B));
    w.net_bytes = carve_workspace(d, mode, B, N, nullptr).bytes;
    w.net = take(w.net_bytes);
  } else {
    w.t_all = reinterpret_cast<float*>(take(sizeof(float) * (size_t)T));
    w.temb_h = reinterpret_cast<float*>(take(sizeof(float) * (size_t)T * D_TEMB_H));
    w.temb = reinterpret_cast<float*>(take(sizeof(float) * (size_t)T * D_TEMB));
    w.kv_time = reinterpret_cast<float*>(take(sizeof(float) * (size_t)T * d.depth * 2 * D_MODEL));
    w.kv_static = reinterpret_cast<float*>(take(sizeof(float) * (size_t)B * d.depth * 2 * d.n_tok * D_MODEL));
    const size_t per_step = tc_fold_bytes_for(d, B);
    long long chunk = per_step ? (long long)((64u << 20) / per_step) : 1;  // ~64 MB of fold tiles in flight
    if (chunk < 1) chunk = 1;
    if (chunk > 64) chunk = 64;
    if (chunk > T) chunk = T;
    w.chunk = (int)chunk;
    w.fold = take(per_step * (size_t)chunk);
    w.done = reinterpret_cast<int*>(take(sizeof(int) * (size_t)cdiv((long long)B * N, 256)));
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

extern "C" int dfb200_ddpm_sample_loop(const dfb200_denoiser_cfg* cfg, const void* packed, int mode, int B, int N,
                                       int T, const float* sched, float* x, int x_T_from_noise, const float* ctx,
                                       const float* anchors, const float* variance, const int* anchor_assignment,
                                       const float* valid_id, const float* noise, uint64_t seed, float* traj,
                                       int traj_interval, void* workspace, size_t workspace_bytes,
                                       dfb200_stream_t stream) {
  PackLayout L;
  int rc = make_pack_layout(cfg, &L);
  if (rc != DFB200_OK) return rc;
  DFB_REQUIRE(B >= 0 && N >= 0 && T >= 1, DFB200_ERR_INVALID_ARG, "ddpm_sample_loop: bad sizes B=%d N=%d T=%d", B, N, T);
  DFB_REQUIRE(mode == DFB200_MODE_FP32 || mode == DFB200_MODE_BF16, DFB200_ERR_INVALID_ARG, "ddpm_sample_loop: unknown mode %d", mode);
  DFB_REQUIRE(traj == nullptr || traj_interval >= 1, DFB200_ERR_INVALID_ARG, "ddpm_sample_loop: traj_interval must be >= 1");
  if (B == 0 || N == 0) return DFB200_OK;
  LoopWorkspace lw = carve_loop(L.d, mode, B, N, T, workspace);
  DFB_REQUIRE(workspace != nullptr && workspace_bytes >= lw.bytes, DFB200_ERR_WORKSPACE,
              "ddpm_sample_loop: workspace too small (%zu < %zu)", workspace_bytes, lw.bytes);
  cudaStream_t st = as_stream(stream);
  const float* P = reinterpret_cast<const float*>(packed);
  const float* valid = (L.d.flags & DFB200_NET_MASK_UNREFERENCED) ? valid_id : nullptr;
  const long long total = (long long)B * 3 * N;
  const bool philox = noise == nullptr;

  // x_T = sqrt(var) * z + anchors  (anchored_diffusion.py:564); Philox draw `T` is the x_T noise
  // x_T_from_noise: 0 = x already holds x_T, 1 = x holds N(0,1) noise, 2 = draw it from Philox
  if (x_T_from_noise) {
    rc = launch_xT_init(total, x, anchors, variance, x_T_from_noise == 2, seed, (uint64_t)T, st);
    if (rc != DFB200_OK) return rc;
  }
  auto keep_traj = [&](int i) -> int {
    if (traj != nullptr && i > 0 && i % traj_interval == 0)
      DFB_CUDA(cudaMemcpyAsync(traj + (size_t)(i / traj_interval - 1) * total, x, sizeof(float) * total, cudaMemcpyDeviceToDevice, st));
    return DFB200_OK;
  };

  if (mode == DFB200_MODE_FP32) {
    Workspace ws = carve_workspace(L.d, mode, B, N, lw.net);
    for (int i = T - 1; i >= 0; --i) {
      fill_step_kernel<<<cdiv(B, 256), 256, 0, st>>>(B, i, lw.t_f, lw.t_i);
      DFB_LAUNCH_CHECK();
      rc = launch_context_kv(L, P, B, lw.t_f, ctx, ws, st);
      if (rc != DFB200_OK) return rc;
      rc = denoiser_forward_fp32(L, P, B, N, x, anchors, variance, anchor_assignment, valid, lw.eps, ws, st);
      if (rc != DFB200_OK) return rc;
      const float* z = philox ? nullptr : noise + (size_t)(T - 1 - i) * total;
      rc = launch_ddpm_step(B, N, T, sched, lw.t_i, x, lw.eps, anchors, variance, z, philox, seed, (uint64_t)i, x, nullptr, st);
      if (rc != DFB200_OK) return rc;
      rc = keep_traj(i);
      if (rc != DFB200_OK) return rc;
    }
    return DFB200_OK;
  }

  // ---- bf16 mode: everything that does not depend on x is hoisted out of the step loop ----
  //   time tables for all T steps (timestep MLP + time half of K/V), static half of K/V per sample, then per chunk of steps
  //   ONE fold launch (K/V -> attention weight tiles) and per step ONE fused kernel (denoiser + eps -> x_{t-1} update).
  arange_kernel<<<cdiv(T, 256), 256, 0, st>>>(T, lw.t_all);
  DFB_LAUNCH_CHECK();
  rc = launch_context_kv_time(L, P, T, lw.t_all, lw.temb_h, lw.temb, lw.kv_time, st);
  if (rc != DFB200_OK) return rc;
  rc = launch_context_kv_static(L, P, B, ctx, lw.kv_static, st);
  if (rc != DFB200_OK) return rc;
  const size_t per_step = tc_fold_bytes_for(L.d, B);
  DFB_CUDA(cudaMemsetAsync(lw.done, 0, sizeof(int) * (size_t)cdiv((long long)B * N, 256), st));
  for (int i0 = T - 1; i0 >= 0; i0 -= lw.chunk) {
    const int steps = i0 + 1 < lw.chunk ? i0 + 1 : lw.chunk;
    rc = launch_context_fold(L, packed, B, lw.kv_static, lw.kv_time, i0, steps, lw.fold, st);
    if (rc != DFB200_OK) return rc;
    // ONE persistent launch runs `steps` timesteps for every 256-token unit (148 CTAs walk the (step, unit) list; a unit's
    // next step waits on its previous one through lw.done), so no SM idles at step boundaries.
    TcUpdate u{};
    u.sched = sched; u.T = T; u.t = i0; u.n_steps = steps; u.fold_step_bytes = per_step;
    u.noise = philox ? nullptr : noise + (size_t)(T - 1 - i0) * total;
    u.seed = seed;
    u.x_out = x;
    u.done = lw.done;
    u.traj = traj; u.traj_interval = traj_interval;
    rc = denoiser_step_tc(L, packed, B, N, x, anchors, variance, anchor_assignment, valid, lw.fold, nullptr, &u, st);
    if (rc != DFB200_OK) return rc;
  }
  return DFB200_OK;
}
