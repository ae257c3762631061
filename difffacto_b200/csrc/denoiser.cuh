// Packed weight image + workspace layout of the cross-diffusion denoiser, shared by the fp32
// (CUDA-core) path, the bf16 (tcgen05) path and the sampling loop.
#pragma once
#include "common.cuh"

namespace dfb200 {

constexpr int D_MODEL = 128;   // inner_dim = n_heads * d_head
constexpr int D_FF = 512;      // GEGLU inner dim (mult=4): proj 128 -> 2*512, out 512 -> 128
constexpr int D_TEMB = 256;    // timestep embedding width (attention.py:357, 393)
constexpr int D_TEMB_H = 1024; // time_embed GEGLU inner dim
constexpr int MAX_TOKENS = 4;  // n_class part tokens
constexpr float LN_EPS = 1e-5f;

// Parameter indices of dfb200_denoiser_pack (see include/difffacto_b200.h)
enum GlobalParam { P_PRE_W, P_PRE_B, P_POST_W, P_POST_B, P_IN_W, P_IN_B, P_TE0_W, P_TE0_B, P_TE2_W, P_TE2_B,
                   P_OUT_W, P_OUT_B, N_GLOBAL_PARAMS };
enum BlockParam { B_N2_W, B_N2_B, B_N3_W, B_N3_B, B_WQ, B_WK, B_WV, B_WO, B_BO, B_W1, B_B1, B_W2, B_B2,
                  N_BLOCK_PARAMS };
constexpr int MAX_DEPTH = 16;

struct NetDims {
  int c_in;    // channels of the concatenated point features (13)
  int c_out;   // 3
  int c_ctx;   // key/value input width (522 = context_dim + n_class + 256)
  int c_ctx_static;  // context_dim + n_class one-hot (266): the part that does not depend on t
  int depth, n_tok, n_heads, d_head, flags;
};

// Offsets (in floats) of every fp32 parameter inside the packed image, and (in bytes from the image
// start) of the bf16 UMMA operand stream consumed by the tcgen05 kernel.
struct PackLayout {
  NetDims d;
  size_t g[N_GLOBAL_PARAMS];
  size_t blk[MAX_DEPTH][N_BLOCK_PARAMS];
  size_t freqs;          // 128 sinusoid frequencies (appended by pack, see denoiser_pack)
  size_t fp32_floats;    // end of the fp32 region
  size_t tc_stream_off;  // byte offset of the bf16 packet stream (16 B aligned)
  size_t tc_stream_bytes;
  size_t tf32_stream_off;  // byte offset of the tf32 packet stream (denoiser_tf32.cu)
  size_t tf32_stream_bytes;
  size_t total_bytes;
};

int make_net_dims(const dfb200_denoiser_cfg* cfg, NetDims* d);          // validates cfg
int make_pack_layout(const dfb200_denoiser_cfg* cfg, PackLayout* L);    // host-side, deterministic
size_t param_numel(const NetDims& d, bool global, int which);

// ---- workspace ------------------------------------------------------------------------------
struct Workspace {
  float* temb_h;  // [B, 1024]
  float* temb;    // [B, 256]
  float* kv;      // [B, depth, 2, n_tok, 128]
  float* x;       // [M, 128]   residual stream (fp32 path)
  float* q;       // [M, 128]   q, then attention output in place (fp32 path)
  float* u;       // [M, 512]   GEGLU output (fp32 path)
  void* fold;     // [B, depth] folded attention packets (bf16 path)
  size_t bytes;
};
Workspace carve_workspace(const NetDims& d, int mode, int B, int N, void* base);

// ---- launches shared between the paths (denoiser_ctx.cu) --------------------------------------
// K/V of every block for every sample: kv[b,l,{k,v},j,:] = W{k,v}_l . [ctx[b,:,j] | onehot(j) | temb(t[b])]
int launch_context_kv(const PackLayout& L, const float* packed, int B, const float* t, const float* ctx,
                      Workspace& ws, cudaStream_t st);

// Time-independent / sample-independent halves of the K/V projection (used by the fused sampling loop):
//   kv_static[b,l,{k,v},j,:] = W[:, :c_static] . [ctx[b,:,j] | onehot(j)]      kv_time[t,l,{k,v},:] = W[:, c_static:] . temb(t)
int launch_context_kv_static(const PackLayout& L, const float* packed, int B, const float* ctx, float* kv_static, cudaStream_t st);
int launch_context_kv_time(const PackLayout& L, const float* packed, int T, const float* t_values, float* temb_h, float* temb,
                           float* kv_time, cudaStream_t st);

constexpr int TC_MAX_LIST_STEPS = 64;  // per-step output pointers travel in the kernel parameters (power of two)
// fused-update arguments of the bf16 kernel (all-zero = plain forward writing eps)
struct TcUpdate {
  const float* sched;  // device schedule table [DFB200_SCHED_ROWS][T]
  int T, t;            // first timestep of this launch (same for every sample of the batch)
  int n_steps;         // timesteps t, t-1, ..., t-n_steps+1 (or step_t[0..n_steps)) run inside ONE persistent launch
  const int* step_t;   // optional device list of this launch's timesteps in execution order (strided DDIM lists)
  int step_base;       // sampling steps completed by earlier launches of the same loop (the `done` counters keep counting)
  size_t fold_step_bytes;  // distance between consecutive steps' fold packets
  const float* noise;  // (n_steps,B,3,N) N(0,1), first slice = step t, or NULL -> Philox(seed, offset = timestep)
  uint64_t seed;
  float* x_out;        // x_{t-1}; must alias x when n_steps > 1
  int* done;           // per 256-token unit: tile-steps completed since the loop began (zeroed by the caller); required if n_steps > 1
  float* traj; int traj_interval;  // optional trajectory slots (see dfb200_ddpm_sample_loop)
  float* step_sample; float* step_xstart;  // optional (n_steps,B,3,N): sample / pred_xstart after each step of this launch
  float* const* step_sample_list; float* const* step_xstart_list;  // or HOST arrays of n_steps (B,3,N) device buffers (n_steps <= TC_MAX_LIST_STEPS)
  const float* ddim_acp; const float* ddim_dir; float ddim_eta;  // DDIM update when ddim_acp != NULL (device float32[T] tables)
  int guidance; float guid_w;  // classifier-free guidance: `fold` holds 2B entries per step (conditional, then unconditional)
};
int denoiser_step_tc(const PackLayout& L, const void* packed, int B, int N, const float* x, const float* anchors,
                     const float* variances, const int* assign, const float* valid_id, const void* fold, float* eps_out,
                     const TcUpdate* upd, cudaStream_t st);
// fold tiles for `steps` timesteps t_first, t_first-1, ... (or step_t[0..steps) when given): fold[(s*B + b)*depth + l]
int launch_context_fold(const PackLayout& L, const void* packed, int B, const float* kv_static, const float* kv_time,
                        int t_first, const int* step_t, int steps, void* fold, cudaStream_t st);
size_t tc_fold_bytes_for(const NetDims& d, int B);

// DFB200_MODE_TF32 (denoiser_tf32.cu): kind::tf32 tensor-core path at reference tolerance; ws.kv must hold the K/V of this call
int denoiser_forward_tf32(const PackLayout& L, const void* packed, int B, int N, const float* x, const float* anchors,
                          const float* variances, const int* assign, const float* valid_id, float* eps_out, Workspace& ws,
                          cudaStream_t st);
size_t tf32_fold_bytes_for(const NetDims& d, int B);

int denoiser_forward_fp32(const PackLayout& L, const float* packed, int B, int N, const float* x,
                          const float* anchors, const float* variances, const int* assign,
                          const float* valid_id, float* eps_out, Workspace& ws, cudaStream_t st);

}  // namespace dfb200
