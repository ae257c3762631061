"""TEST INFRASTRUCTURE ONLY (never imported by the product): CPU restatement in plain PyTorch of the encoder side of
sampling - PartEncoder.sample_latents and what it calls - pinned to the real reference by tests/golden/latents_golden.npz.

Follows /root/reference/python/difffacto/models/encoders/part_encoders.py:1052-1110 (sample_latents), :20-143
(PartAlignerTransformer), :417-428 (gather_all), :1317-1327 (prepare_ctx), encoders/flow.py:7-78 (coupling flows) and
diffusions/nets/attention.py:179-204, 296-306 (self-attention block), for the configuration of configs/gen_chair.py."""
import math

import torch
import torch.nn.functional as F


def encoder_param_shapes(depth=5, flow_depth=14, zdim=256, hidden=256, inner=256, noise_dim=32, n_class=4):
    s = {"part_aligner.class_emb.weight": (n_class, inner), "part_aligner.pre_norm.weight": (inner,), "part_aligner.pre_norm.bias": (inner,),
         "part_aligner.post_norm.weight": (inner,), "part_aligner.post_norm.bias": (inner,),
         "part_aligner.proj_in.weight": (inner, zdim + noise_dim), "part_aligner.proj_in.bias": (inner,),
         "part_aligner.proj_out.weight": (6, inner), "part_aligner.proj_out.bias": (6,)}
    for i in range(depth):
        p = f"part_aligner.transformer_blocks.{i}."
        s.update({p + "ff.net.0.proj.weight": (8 * inner, inner), p + "ff.net.0.proj.bias": (8 * inner,),
                  p + "ff.net.2.weight": (inner, 4 * inner), p + "ff.net.2.bias": (inner,),
                  p + "attn2.to_q.weight": (inner, inner), p + "attn2.to_k.weight": (inner, inner), p + "attn2.to_v.weight": (inner, inner),
                  p + "attn2.to_out.0.weight": (inner, inner), p + "attn2.to_out.0.bias": (inner,),
                  p + "norm2.weight": (inner,), p + "norm2.bias": (inner,), p + "norm3.weight": (inner,), p + "norm3.bias": (inner,)})
    half = zdim - zdim // 2
    for c in range(n_class):
        for j in range(flow_depth):
            p = f"flow.{c}.chain.{j}.net_s_t."
            s.update({p + "0.weight": (hidden, half), p + "0.bias": (hidden,), p + "2.weight": (hidden, hidden), p + "2.bias": (hidden,),
                      p + "4.weight": ((zdim - half) * 2, hidden), p + "4.bias": ((zdim - half) * 2,)})
    return s


def pointnet_param_shapes(zdim=256, n_class=4, prefix="encoder."):
    """PointNetV2(per_part_mlp=True) parameters and buffers (reference models/encoders/pointnet.py:122-181)."""
    s = {}
    for i, (ci, co) in enumerate([(3, 128), (128, 128), (128, 256), (256, 512)], 1):
        s[f"{prefix}conv{i}.weight"], s[f"{prefix}conv{i}.bias"] = (co, ci, 1), (co,)
        for n in ("weight", "bias", "running_mean", "running_var"):
            s[f"{prefix}bn{i}.{n}"] = (co,)
        s[f"{prefix}bn{i}.num_batches_tracked"] = ()
    g = n_class
    for head in ("mlp_m", "mlp_v"):
        for idx, (ci, co) in zip((0, 3, 6), [(512, 256), (256, 128), (128, zdim)]):
            s[f"{prefix}{head}.{idx}.weight"], s[f"{prefix}{head}.{idx}.bias"] = (co * g, ci, 1), (co * g,)
        for idx, c in zip((1, 4), (256, 128)):
            for n in ("weight", "bias", "running_mean", "running_var"):
                s[f"{prefix}{head}.{idx}.{n}"] = (c * g,)
            s[f"{prefix}{head}.{idx}.num_batches_tracked"] = ()
    return s


def synthetic_encoder_state_dict(seed=0, with_pointnet=False, with_aligner=True, **kw):
    """Deterministic weights (name order, seeded generator): Linear ~ N(0, 1/fan_in), biases ~ 0.05 N, LayerNorm gain 1 + 0.1 N."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    shapes = {k: v for k, v in encoder_param_shapes(**kw).items() if with_aligner or not k.startswith("part_aligner.")}
    if with_pointnet:
        shapes.update(pointnet_param_shapes())
    for name, shape in shapes.items():
        if name.endswith("num_batches_tracked"):
            sd[name] = torch.zeros((), dtype=torch.long)
            continue
        if name.endswith("running_mean"):
            sd[name] = 0.1 * torch.randn(shape, generator=g)
            continue
        if name.endswith("running_var"):
            sd[name] = 0.5 + torch.rand(shape, generator=g)
            continue
        if ".bn" in name or (name.split(".")[-2] in ("1", "4") and ("mlp_m" in name or "mlp_v" in name)):
            sd[name] = (1.0 + 0.1 * torch.randn(shape, generator=g)) if name.endswith("weight") else 0.05 * torch.randn(shape, generator=g)
            continue
        if name.endswith("bias"):
            sd[name] = 0.05 * torch.randn(shape, generator=g)
        elif "norm" in name:
            sd[name] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif "class_emb" in name:
            sd[name] = 0.5 * torch.randn(shape, generator=g)
        elif "net_s_t.4" in name:
            sd[name] = 0.3 * torch.randn(shape, generator=g) / math.sqrt(shape[1])  # keep 14 stacked coupling layers well conditioned
        else:
            sd[name] = torch.randn(shape, generator=g) / math.sqrt(shape[1])
    return sd


def flow_reverse(sd, c, x, depth=14):
    """SequentialFlow(reverse=True) of part c: flow.py:24-71."""
    d = x.shape[1] - x.shape[1] // 2
    for j in range(depth - 1, -1, -1):
        p = f"flow.{c}.chain.{j}.net_s_t."
        swap = j % 2 == 0
        if swap:
            x = torch.cat([x[:, d:], x[:, :d]], 1)
        h = F.relu(F.linear(x[:, :d], sd[p + "0.weight"], sd[p + "0.bias"]))
        h = F.relu(F.linear(h, sd[p + "2.weight"], sd[p + "2.bias"]))
        s_t = F.linear(h, sd[p + "4.weight"], sd[p + "4.bias"])
        out_dim = x.shape[1] - d
        scale = torch.sigmoid(s_t[:, :out_dim] + 2.)
        shift = s_t[:, out_dim:]
        y1 = (x[:, d:] - shift) / scale
        x = torch.cat([x[:, :d], y1], 1) if not swap else torch.cat([y1, x[:, :d]], 1)
    return x


def part_aligner(sd, x, mask, noise, depth=5, heads=8, noise_scale=100.0, n_class=4):
    """PartAlignerTransformer.forward (cimle, cond_noise_type 0, add_class_cond): part_encoders.py:86-143."""
    P = "part_aligner."
    B = x.shape[0]
    x = torch.cat([x, (noise * noise_scale).unsqueeze(-1).expand(-1, -1, n_class)], dim=1)
    h = F.linear(x.transpose(1, 2), sd[P + "proj_in.weight"], sd[P + "proj_in.bias"]) + sd[P + "class_emb.weight"].unsqueeze(0)
    inner = h.shape[-1]
    # (sic) with cimle=True and cond_noise_type=0 the reference never applies pre_norm: its `else: x = self.pre_norm(x)`
    # belongs to `if self.cimle:` (part_encoders.py:115-128)
    for i in range(depth):
        p = P + f"transformer_blocks.{i}."
        a = F.layer_norm(h, (inner,), sd[p + "norm2.weight"], sd[p + "norm2.bias"])
        q, k, v = (F.linear(a, sd[p + f"attn2.to_{n}.weight"]) for n in "qkv")
        split = lambda t: t.reshape(B, n_class, heads, -1).permute(0, 2, 1, 3)  # noqa: E731  b n (h d) -> b h n d
        q, k, v = split(q), split(k), split(v)
        sim = torch.einsum("bhid,bhjd->bhij", q, k) * (q.shape[-1] ** -0.5)
        sim = sim.masked_fill(~mask.to(bool)[:, None, None, :], -torch.finfo(sim.dtype).max)
        o = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v).permute(0, 2, 1, 3).reshape(B, n_class, inner)
        h = F.linear(o, sd[p + "attn2.to_out.0.weight"], sd[p + "attn2.to_out.0.bias"]) + h
        f = F.layer_norm(h, (inner,), sd[p + "norm3.weight"], sd[p + "norm3.bias"])
        u = F.linear(f, sd[p + "ff.net.0.proj.weight"], sd[p + "ff.net.0.proj.bias"])
        a_, g_ = u.chunk(2, dim=-1)
        h = F.linear(a_ * F.gelu(g_), sd[p + "ff.net.2.weight"], sd[p + "ff.net.2.bias"]) + h
    h = F.layer_norm(h, (inner,), sd[P + "post_norm.weight"], sd[P + "post_norm.bias"])
    out = F.linear(h, sd[P + "proj_out.weight"], sd[P + "proj_out.bias"]).transpose(1, 2)
    return torch.split(out, 3, dim=1)


@torch.no_grad()
def sample_latents(sd, prior, noise, valid_id, fixed_id, sample_points, K, n_class=4):
    """part_encoders.py:1052-1110 with the prior draw (B,256,4) and the cIMLE noise (B*K,32) supplied."""
    B = prior.shape[0]
    part_code = torch.stack([flow_reverse(sd, c, prior[..., c]) for c in range(n_class)], dim=-1)
    fixed_codes = part_code[0].unsqueeze(0)
    fixed_valid = (valid_id[0].unsqueeze(0) + fixed_id[None]).clamp(min=0, max=1)
    part_code = part_code * (1 - fixed_id).unsqueeze(0).unsqueeze(0) + fixed_id.unsqueeze(0).unsqueeze(0) * fixed_codes
    valid_id = valid_id * (1 - fixed_id).unsqueeze(0) + fixed_id.unsqueeze(0) * fixed_valid
    if torch.any(fixed_id == 1):
        noise = noise.reshape(B, K, -1)[0].unsqueeze(0).expand(B, -1, -1).reshape(B * K, -1)
    part_code = part_code.repeat_interleave(K, dim=0)
    valid_id = valid_id.repeat_interleave(K, dim=0)
    mean, logvar = part_aligner(sd, part_code, valid_id, noise)
    ids = torch.arange(n_class).unsqueeze(0) * valid_id + torch.argmax(valid_id, dim=1).unsqueeze(1) * (1 - valid_id)
    seg = ids.to(torch.int32).unsqueeze(-1).expand(-1, -1, sample_points // n_class).reshape(B * K, sample_points)
    idx = seg.long().unsqueeze(1).expand(-1, 3, -1)
    mean_pp, logvar_pp = torch.gather(mean, 2, idx), torch.gather(logvar, 2, idx)
    ctx = [part_code, torch.cat([mean, torch.exp(logvar)], dim=1)]
    return ctx, mean_pp, logvar_pp, seg, valid_id, [part_code, mean, logvar, noise]
