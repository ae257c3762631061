"""ORACLE -- TEST INFRASTRUCTURE ONLY: fp32 PyTorch restatement of the reference's sampling step.

Never imported by the product path (difffacto_b200/).  Used by tests/ as the checker of the CUDA
denoiser / DDPM kernels, and by bench.py's cpu_baseline and `--impl reference` legs as the CPU
port of the reference's own PyTorch path (the reference's Python cannot travel to the GPU box).

Each function restates, op for op and in the same order, the cited reference lines (all under
/root/reference/python/difffacto/models/diffusions/):
  nets/utils.py:7-24                 timestep_embedding
  nets/attention.py:50-57, 77-94     GEGLU / FeedForward
  nets/attention.py:179-204          CrossAttention.forward (masked softmax over the 4 part tokens)
  nets/attention.py:296-306          BasicTransformerBlock._forward (single_attn)
  nets/attention.py:385-440          TransformerNet.forward / _forward_attn
  anchored_diffusion.py:62-112       schedule tables (float64 numpy)
  anchored_diffusion.py:148-173      q_sample
  anchored_diffusion.py:227-395, 401-409, 175-193, 450-484   p_mean_variance / p_sample (config path)
  anchored_diffusion.py:528-588      p_sample_loop_progressive

Parity pinning: this restatement is checked against the reference implementation itself --
imported from /root/reference in the build container -- both live (tests/test_oracle_vs_reference.py,
skipped when /root/reference is absent) and through the committed golden vectors
tests/golden/denoiser_golden.npz produced by tests/golden/make_golden.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ---------------------------------------------------------------------------------------------
# synthetic weights: deterministic, reference-free, same names/shapes as the reference state_dict
# ---------------------------------------------------------------------------------------------
GEN_CHAIR_NET = dict(in_channels=3, out_channels=3, n_heads=8, d_head=16, depth=5, context_dim=262, n_class=4)


def param_shapes(depth=5, c_in=13, c_ctx=522, c_out=3):
    s = {
        "pre_norm.weight": (128,), "pre_norm.bias": (128,), "post_norm.weight": (128,), "post_norm.bias": (128,),
        "proj_in.weight": (128, c_in), "proj_in.bias": (128,),
        "time_embed.net.0.proj.weight": (2048, 256), "time_embed.net.0.proj.bias": (2048,),
        "time_embed.net.2.weight": (256, 1024), "time_embed.net.2.bias": (256,),
        "proj_out.weight": (c_out, 128), "proj_out.bias": (c_out,),
    }
    for i in range(depth):
        p = f"transformer_blocks.{i}."
        s.update({
            p + "norm2.weight": (128,), p + "norm2.bias": (128,), p + "norm3.weight": (128,), p + "norm3.bias": (128,),
            p + "attn2.to_q.weight": (128, 128), p + "attn2.to_k.weight": (128, c_ctx), p + "attn2.to_v.weight": (128, c_ctx),
            p + "attn2.to_out.0.weight": (128, 128), p + "attn2.to_out.0.bias": (128,),
            p + "ff.net.0.proj.weight": (1024, 128), p + "ff.net.0.proj.bias": (1024,),
            p + "ff.net.2.weight": (128, 512), p + "ff.net.2.bias": (128,),
        })
    return s


def synthetic_state_dict(seed=0, depth=5, c_ctx=522):
    """Weights ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)) (nn.Linear's default scale), LayerNorm gains
    around 1, all biases non-zero, from numpy's PCG64 so that no torch RNG detail is involved."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in param_shapes(depth=depth, c_ctx=c_ctx).items():
        if "norm" in name and name.endswith("weight"):
            v = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name.endswith("bias"):
            v = 0.1 * rng.standard_normal(shape)
        else:
            bound = 1.0 / math.sqrt(shape[1])
            v = rng.uniform(-bound, bound, size=shape)
        sd[name] = torch.from_numpy(v.astype(np.float32))
    return sd


def synthetic_inputs(seed, B, N, all_valid=False):
    """Part-segmented synthetic batch in the shapes AnchorDiffAE.decode feeds the sampler
    (models/networks/anchor_gen.py:145-169; encoders/part_encoders.py:1052-1110)."""
    rng = np.random.default_rng(seed)
    code = rng.standard_normal((B, 256, 4)).astype(np.float32)
    mean = (0.3 * rng.standard_normal((B, 3, 4))).astype(np.float32)
    logvar = rng.uniform(math.log(0.01), math.log(0.1), size=(B, 3, 4)).astype(np.float32)
    valid = np.ones((B, 4), np.float32)
    if not all_valid:
        valid = (rng.random((B, 4)) < 0.8).astype(np.float32)
        valid[np.arange(B), rng.integers(0, 4, B)] = 1.0
    # seg_mask = arange(4)*valid + argmax(valid)*(1-valid), each part N/4 points (part_encoders.py:1105-1106)
    first = valid.argmax(1)
    part = (np.arange(4)[None] * valid + first[:, None] * (1 - valid)).astype(np.int32)
    assign = np.repeat(part, N // 4, axis=1).astype(np.int32)
    var = np.exp(logvar)
    anchors = np.take_along_axis(mean, np.broadcast_to(assign[:, None, :], (B, 3, N)).astype(np.int64), axis=2)
    variance = np.take_along_axis(var, np.broadcast_to(assign[:, None, :], (B, 3, N)).astype(np.int64), axis=2)
    params = np.concatenate([mean, var], axis=1).astype(np.float32)  # (B, 6, 4)
    x = (np.sqrt(variance) * rng.standard_normal((B, 3, N)) + anchors).astype(np.float32)
    t = rng.integers(0, 100, size=(B,)).astype(np.int64)
    noise = rng.standard_normal((B, 3, N)).astype(np.float32)
    T = torch.from_numpy
    C = lambda a: T(np.ascontiguousarray(a, dtype=np.float32))  # noqa: E731  (take_along_axis output is not C-contiguous)
    return dict(x=T(x), t=T(t), code=T(code), params=T(params), anchors=C(anchors), variance=C(variance),
                assign=T(assign), valid=T(valid), noise=T(noise))


# ---------------------------------------------------------------------------------------------
# denoiser
# ---------------------------------------------------------------------------------------------
def timestep_embedding(timesteps, dim, max_period=10000):  # nets/utils.py:7-24
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].to(timesteps.dtype) * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _geglu_ff(x, w0, b0, w2, b2):  # attention.py:50-57, 77-94 (dropout is identity in eval)
    h = F.linear(x, w0, b0)
    a, gate = h.chunk(2, dim=-1)
    return F.linear(a * F.gelu(gate), w2, b2)


def _cross_attention(x, context, mask, wq, wk, wv, wo, bo, heads=8):  # attention.py:179-204
    B, n, _ = x.shape
    q = F.linear(x, wq)
    k = F.linear(context, wk)
    v = F.linear(context, wv)

    def split(t):  # 'b n (h d) -> (b h) n d'
        return t.reshape(B, t.shape[1], heads, -1).permute(0, 2, 1, 3).reshape(B * heads, t.shape[1], -1)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum("bid,bjd->bij", q, k) * (q.shape[-1] ** -0.5)
    if mask is not None:
        m = mask.to(bool)[:, None, None, :].expand(B, heads, 1, mask.shape[1]).reshape(B * heads, 1, -1)
        sim = sim.masked_fill(~m, -torch.finfo(sim.dtype).max)
    sim = sim.softmax(dim=-1)
    out = torch.einsum("bij,bjd->bid", sim, v)
    out = out.reshape(B, heads, n, -1).permute(0, 2, 1, 3).reshape(B, n, -1)
    return F.linear(out, wo, bo)


def denoiser_forward(sd, x, t, ctx_list, anchors, variances, valid_id, anchor_assignment, depth=5, n_class=4,
                     mask_out_unreferenced_code=True, include_std=False):
    """TransformerNet.forward on the gen_chair configuration (attention.py:385-440).
    x, anchors, variances: (B,3,N) channel-major; ctx_list: [(B,256,4), (B,6,4)]; t: (B,) integer."""
    ctx = torch.cat(ctx_list, dim=1) if isinstance(ctx_list, (list, tuple)) else ctx_list
    ctx = ctx.transpose(1, 2).contiguous()  # b c n -> b n c
    B = x.shape[0]
    class_embed = torch.eye(n_class).to(x).unsqueeze(0).repeat_interleave(B, dim=0)
    ctx = torch.cat([ctx, class_embed], dim=-1)
    t_embed = _geglu_ff(timestep_embedding(t, 256), sd["time_embed.net.0.proj.weight"], sd["time_embed.net.0.proj.bias"],
                        sd["time_embed.net.2.weight"], sd["time_embed.net.2.bias"])
    ctx = torch.cat([ctx, t_embed.unsqueeze(1).expand(-1, ctx.shape[1], -1)], dim=-1)
    v = torch.sqrt(variances) if include_std else variances
    feat = torch.cat([x, anchors, v, F.one_hot(anchor_assignment.to(torch.long), num_classes=n_class).transpose(1, 2).to(x)], dim=1)
    h = F.linear(feat.transpose(1, 2).contiguous(), sd["proj_in.weight"], sd["proj_in.bias"])
    h = F.layer_norm(h, (128,), sd["pre_norm.weight"], sd["pre_norm.bias"])
    mask = valid_id if mask_out_unreferenced_code else None
    for i in range(depth):
        p = f"transformer_blocks.{i}."
        a = F.layer_norm(h, (128,), sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        h = _cross_attention(a, ctx, mask, sd[p + "attn2.to_q.weight"], sd[p + "attn2.to_k.weight"], sd[p + "attn2.to_v.weight"],
                             sd[p + "attn2.to_out.0.weight"], sd[p + "attn2.to_out.0.bias"]) + h
        f = F.layer_norm(h, (128,), sd[p + "norm3.weight"], sd[p + "norm3.bias"])
        h = _geglu_ff(f, sd[p + "ff.net.0.proj.weight"], sd[p + "ff.net.0.proj.bias"], sd[p + "ff.net.2.weight"],
                      sd[p + "ff.net.2.bias"]) + h
    h = F.layer_norm(h, (128,), sd["post_norm.weight"], sd["post_norm.bias"])
    out = F.linear(h, sd["proj_out.weight"], sd["proj_out.bias"])
    return out.transpose(1, 2).contiguous()  # (B,3,N)


# ---------------------------------------------------------------------------------------------
# anchored DDPM
# ---------------------------------------------------------------------------------------------
SCHED_ROWS = ["sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
              "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_mean_coef1", "posterior_mean_coef2",
              "posterior_mean_coef3"]


def schedule(T, beta_1=1e-4, beta_T=0.02):
    """float64 tables of anchored_diffusion.py:62-112 (linear mode)."""
    betas = np.linspace(beta_1, beta_T, num=T, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    s = {
        "betas": betas,
        "sqrt_alphas_cumprod": np.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
        "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1),
        "posterior_variance": betas * (1.0 - ac_prev) / (1.0 - ac),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
        "posterior_mean_coef3": 1.0 + ((np.sqrt(ac) - 1.0) * (np.sqrt(ac_prev) + np.sqrt(alphas))) / (1.0 - ac),
    }
    return s


def schedule_table(T, **kw):
    """(8, T) float32 table in the row order of include/difffacto_b200.h (DFB200_SCHED_*)."""
    s = schedule(T, **kw)
    return np.stack([s[k].astype(np.float32) for k in SCHED_ROWS])


def _extract(arr, t, shape):  # diffusion_utils.py:42-66
    res = torch.from_numpy(arr).to(t.device).float()[t]
    while res.dim() < len(shape):
        res = res[..., None]
    return res.expand(shape)


def q_sample(s, x_start, t, anchors, variance, noise):  # anchored_diffusion.py:148-173
    L = torch.sqrt(variance)
    return (_extract(s["sqrt_alphas_cumprod"], t, x_start.shape) * (x_start - anchors) + anchors
            + _extract(s["sqrt_one_minus_alphas_cumprod"], t, x_start.shape) * L * noise)


def ddpm_step(s, x, t, eps, anchors, variance, noise):
    """eps -> (x_{t-1}, pred_xstart): anchored_diffusion.py:306-314, 401-409, 184-188, 476-483."""
    L = torch.sqrt(variance)
    model_variance = _extract(s["posterior_variance"], t, x.shape) * variance
    pred_xstart = (_extract(s["sqrt_recip_alphas_cumprod"], t, x.shape) * (x - anchors) + anchors
                   - _extract(s["sqrt_recipm1_alphas_cumprod"], t, x.shape) * L * eps)
    mean = (_extract(s["posterior_mean_coef1"], t, x.shape) * pred_xstart
            + _extract(s["posterior_mean_coef2"], t, x.shape) * x
            + _extract(s["posterior_mean_coef3"], t, x.shape) * anchors)
    nonzero_mask = (t != 0).float().view(-1, 1, 1)
    sample = mean + nonzero_mask * torch.sqrt(model_variance) * noise
    return sample, pred_xstart


def ddim_steps(T, nsteps=10, discretize="uniform"):  # anchored_diffusion.py:117-126
    if discretize == "uniform":
        return list(range(0, T, T // nsteps))
    return ((np.linspace(0., math.sqrt(T * 0.8), nsteps) ** 2).astype(np.int32)).tolist()


def ddim_step(s, x, t, eps, anchors, variance, noise, eta):
    """DDIM variant: anchored_diffusion.py:114-116 (xt_dir_coeff), :368-374 (xt_dir), :480-481 (sample)."""
    betas = s["betas"]
    ac = np.cumprod(1.0 - betas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    xt_dir_coeff = np.sqrt(1. - ac - eta * eta * s["posterior_variance"])
    L = torch.sqrt(variance)
    model_variance = _extract(s["posterior_variance"], t, x.shape) * variance
    pred_xstart = (_extract(s["sqrt_recip_alphas_cumprod"], t, x.shape) * (x - anchors) + anchors
                   - _extract(s["sqrt_recipm1_alphas_cumprod"], t, x.shape) * L * eps)
    xt_dir = L * _extract(xt_dir_coeff, t, x.shape) * eps
    nonzero_mask = (t != 0).float().view(-1, 1, 1)
    sample = ((pred_xstart - anchors) * torch.sqrt(_extract(ac_prev, t, x.shape)) + anchors + xt_dir
              + eta * nonzero_mask * torch.sqrt(model_variance) * noise)
    return sample, pred_xstart


@torch.no_grad()
def p_sample(sd, s, x, t, ctx_list, anchors, variance, assign, valid, noise, depth=5, guidance_weight=None, ddim_eta=None):
    """anchored_diffusion.py:450-484; guidance_weight: classifier-free guidance (:263-266); ddim_eta: DDIM step."""
    eps = denoiser_forward(sd, x, t, ctx_list, anchors, variance, valid, assign, depth=depth)
    if guidance_weight is not None:
        unc = denoiser_forward(sd, x, t, [torch.zeros_like(r) for r in ctx_list], anchors, variance, valid, assign, depth=depth)
        eps = (1. - guidance_weight) * unc + guidance_weight * eps
    if ddim_eta is not None:
        return ddim_step(s, x, t, eps, anchors, variance, noise, ddim_eta) + (eps,)
    return ddpm_step(s, x, t, eps, anchors, variance, noise) + (eps,)


@torch.no_grad()
def p_sample_loop(sd, T, ctx_list, anchors, variance, assign, valid, noise_T, noises, steps=None, depth=5):
    """anchored_diffusion.py:528-588 with the per-step noise supplied (noises[k] is the k-th draw of
    the loop, i.e. the one used at t = T-1-k).  `steps` truncates the loop (for timing samples)."""
    s = schedule(T)
    x = torch.sqrt(variance) * noise_T + anchors
    B = x.shape[0]
    for k, i in enumerate(range(T - 1, -1, -1)):
        if steps is not None and k >= steps:
            break
        t = torch.tensor([i] * B)
        x, _, _ = p_sample(sd, s, x, t, ctx_list, anchors, variance, assign, valid, noises[k], depth=depth)
    return x
