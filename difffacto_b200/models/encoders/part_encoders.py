"""ENCODERS['PartEncoderForTransformerDecoder'] / ENCODERS['PartAlignerTransformer']: the encoder side of SAMPLING.

Mirror of the generation-time slice of the reference classes (python/difffacto/models/encoders/part_encoders.py:
PartAlignerTransformer :20-143, PartEncoder.__init__ :313-387, gather_all :417-428, get_params_from_part_code :447-459,
sample_latents :1052-1110, PartEncoderForTransformerDecoder.prepare_ctx :1317-1327; encoders/flow.py:7-78) for the
configuration of configs/gen_*.py: part codes are drawn from the prior, pushed through the 4 latent coupling flows in
reverse, the part-aligner transformer (5 self-attention blocks over the 4 part tokens, cIMLE noise concatenated to the
input) predicts per-part mean / log-variance, and the per-point anchors / variances / assignments and the context list
of the denoiser are assembled.  Same constructor keywords, same parameter names (`part_aligner.*`, `flow.{i}.chain.{j}.
net_s_t.{0,2,4}.*`: reference checkpoints restore with strict=False - the point-cloud encoder `encoder.*`, used only for
reconstruction/training, is not built), same return tuple as `sample_latents`.

The modules only HOLD parameters; the arithmetic runs in this repo's CUDA kernels through the C ABI (dfb200_sgemm,
dfb200_layernorm_forward, dfb200_token_attention, dfb200_geglu_forward, dfb200_coupling_reverse, dfb200_gather_points, ...).
This is per-batch work before the sampling loop (launch-latency bound, < 1 % of a 1000-step sampling run)."""
import math

import torch
import torch.nn as nn

from ... import _lib
from ... import train_ops as T
from ..._lib import check, ptr, stream
from ...pointnet2_ops.pointnet2_utils import gather_operation
from ...utils.registry import ENCODERS, build_from_cfg
from ..diffusions.nets.attention import _CrossAttention, _FeedForward


def _call(name, *args, device):
    with _lib.on(device):
        check(getattr(_lib.load(), name)(*args, stream()))


def _layernorm(x2d, ln):
    y = torch.empty_like(x2d)
    _call("dfb200_layernorm_forward", x2d.shape[0], x2d.shape[1], ptr(x2d), ptr(ln.weight), ptr(ln.bias), ptr(y), device=x2d.device)
    return y


def _relu_(x):
    _call("dfb200_relu", x.numel(), ptr(x), device=x.device)
    return x


def _scale(x, alpha):
    x = x.contiguous()
    y = torch.empty_like(x)
    _call("dfb200_scale", x.numel(), float(alpha), ptr(x), ptr(y), device=x.device)
    return y


def _exp_shift(x, shift=0.0):
    x = x.contiguous()
    y = torch.empty_like(x)
    _call("dfb200_exp_shift", x.numel(), float(shift), ptr(x), ptr(y), device=x.device)
    return y


class CouplingLayer(nn.Module):  # reference encoders/flow.py:7-45
    def __init__(self, d, intermediate_dim, swap=False):
        super().__init__()
        self.d = d - (d // 2)
        self.swap = swap
        self.net_s_t = nn.Sequential(nn.Linear(self.d, intermediate_dim), nn.ReLU(inplace=True),
                                     nn.Linear(intermediate_dim, intermediate_dim), nn.ReLU(inplace=True),
                                     nn.Linear(intermediate_dim, (d - self.d) * 2))

    def reverse_(self, x):
        """In-place reverse pass on x (B, D) contiguous: the conditioning half is x[:, :d] (x[:, d:] when swap), the other
        half becomes (half - shift) / sigmoid(s + 2).  (The reference swaps the halves, transforms, and swaps back.)"""
        B, D = x.shape
        d = self.d
        cond = x[:, d:] if self.swap else x[:, :d]
        target = x[:, :D - d] if self.swap else x[:, d:]
        h = torch.empty(B, self.net_s_t[0].out_features, device=x.device)
        T._sgemm(True, True, B, h.shape[1], cond.shape[1], cond, D, self.net_s_t[0].weight, cond.shape[1], h, h.shape[1], bias=self.net_s_t[0].bias)
        h = T.linear(_relu_(h), self.net_s_t[2].weight, self.net_s_t[2].bias)
        s_t = T.linear(_relu_(h), self.net_s_t[4].weight, self.net_s_t[4].bias)
        _call("dfb200_coupling_reverse", B, D - d, ptr(s_t), ptr(target), D, device=x.device)
        return x


    def forward_with_logdet(self, x, logpx):
        """Forward direction (training prior loss), differentiable: reference flow.py:24-45 with reverse=False."""
        d = self.d
        if self.swap:
            x = torch.cat([x[:, d:], x[:, :d]], 1)
        cond, x2 = x[:, :d].contiguous(), x[:, d:].contiguous()
        h = T.relu(T.linear(cond, self.net_s_t[0].weight, self.net_s_t[0].bias))
        h = T.relu(T.linear(h, self.net_s_t[2].weight, self.net_s_t[2].bias))
        s_t = T.linear(h, self.net_s_t[4].weight, self.net_s_t[4].bias)
        y1, logdet = T.coupling_forward(s_t, x2)
        y = torch.cat([cond, y1], 1) if not self.swap else torch.cat([y1, cond], 1)
        return y, logpx - logdet.view(-1, 1)


class SequentialFlow(nn.Module):  # reference encoders/flow.py:48-71
    def __init__(self, layers):
        super().__init__()
        self.chain = nn.ModuleList(layers)

    def forward(self, x, logpx=None, reverse=False, inds=None):
        if not reverse:
            assert logpx is not None
            for i in (range(len(self.chain)) if inds is None else inds):
                x, logpx = self.chain[i].forward_with_logdet(x, logpx)
            return x, logpx
        if logpx is not None:
            raise NotImplementedError("difffacto_b200: the reverse direction of the latent flow is built without log-density tracking")
        x = x.to(torch.float32).contiguous().clone()
        for i in range(len(self.chain) - 1, -1, -1):
            self.chain[i].reverse_(x)
        return x


def build_latent_flow(latent_flow_depth, latent_flow_hidden_dim, latent_dim):
    return SequentialFlow([CouplingLayer(latent_dim, latent_flow_hidden_dim, swap=(i % 2 == 0)) for i in range(latent_flow_depth)])


class _AlignerBlock(nn.Module):  # BasicTransformerBlock(single_attn=True, context_dim=None): attention.py:259-306
    def __init__(self, dim, n_heads, d_head, dropout):
        super().__init__()
        self.ff = _FeedForward(dim, dropout=dropout)
        self.attn2 = _CrossAttention(dim, dim, n_heads, d_head, dropout)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)


@ENCODERS.register_module()
class PartAlignerTransformer(nn.Module):
    def __init__(self, in_channels, n_heads, d_head, out_channels, depth=1, dropout=0., use_linear=False, n_class=4,
                 use_checkpoint=False, single_attn=False, class_cond=True, mask_out_unreferenced_code=True, cimle=False,
                 noise_dim=32, noise_scale=10, cimle_start_epoch=0, add_class_cond=False, cond_noise_type=0,
                 cond_noise_as_token=False):
        super().__init__()
        unsupported = dict(use_linear=not use_linear, single_attn=not single_attn, class_cond=not class_cond,
                           add_class_cond=not add_class_cond, cond_noise_type=cond_noise_type != 0, cond_noise_as_token=cond_noise_as_token)
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError("difffacto_b200.PartAlignerTransformer implements the configuration of configs/gen_*.py; "
                                      f"unsupported setting(s): {bad}")
        self.n_class, self.cimle, self.noise_scale, self.noise_dim = n_class, cimle, noise_scale, noise_dim
        self.cimle_start_epoch = cimle_start_epoch
        self.mask_out_unreferenced_code = mask_out_unreferenced_code
        self.n_heads, self.d_head = n_heads, d_head
        self.in_channels = in_channels + int(cimle) * noise_dim
        self.inner_dim = inner = n_heads * d_head
        self.class_emb = nn.Embedding(n_class, inner)
        self.pre_norm = nn.LayerNorm(inner)
        self.post_norm = nn.LayerNorm(inner)
        self.proj_in = nn.Linear(self.in_channels, inner)
        self.transformer_blocks = nn.ModuleList([_AlignerBlock(inner, n_heads, d_head, dropout) for _ in range(depth)])
        self.proj_out = nn.Linear(inner, out_channels)

    @torch.no_grad()
    def forward(self, x, mask=None, noise=None):
        """x (B, C, n_class) part codes, mask (B, n_class) valid parts, noise (B, noise_dim) -> mean, logvar (B, 3, n_class)."""
        assert x.shape[-1] == self.n_class
        _lib.require_cuda(x, mask, noise)
        B, K = x.shape[0], self.n_class
        tok = x.to(torch.float32).transpose(1, 2)                                   # (B, K, C)
        if self.cimle:
            assert noise is not None
            if noise.shape[1] != self.noise_dim:
                noise = torch.zeros(B, self.noise_dim, device=x.device)
            noise = _scale(noise.to(torch.float32), self.noise_scale)
            tok = torch.cat([tok, noise.unsqueeze(1).expand(-1, K, -1)], dim=-1)
        tok = tok.reshape(B * K, self.in_channels).contiguous()
        cls = self.class_emb.weight.unsqueeze(0).expand(B, -1, -1).reshape(B * K, self.inner_dim)
        h = T.linear(tok, self.proj_in.weight, self.proj_in.bias, cls)               # proj_in(x) + class_emb (:110-113)
        if not self.cimle:  # (sic) reference :115-128: `else: x = self.pre_norm(x)` pairs with `if self.cimle:`, so the cIMLE
            h = _layernorm(h, self.pre_norm)  # configuration (cond_noise_type 0) never applies pre_norm
        valid = mask.to(torch.float32).contiguous() if (mask is not None and self.mask_out_unreferenced_code) else None
        for blk in self.transformer_blocks:
            a = _layernorm(h, blk.norm2)
            q = T.linear(a, blk.attn2.to_q.weight)
            k = T.linear(a, blk.attn2.to_k.weight)
            v = T.linear(a, blk.attn2.to_v.weight)
            o = torch.empty_like(q)
            _call("dfb200_token_attention", B, K, self.n_heads, self.d_head, ptr(q), ptr(k), ptr(v), ptr(valid), ptr(o), device=x.device)
            h = T.linear(o, blk.attn2.to_out[0].weight, blk.attn2.to_out[0].bias, h)
            f = _layernorm(h, blk.norm3)
            u = T.geglu(T.linear(f, blk.ff.net[0].proj.weight, blk.ff.net[0].proj.bias))
            h = T.linear(u, blk.ff.net[2].weight, blk.ff.net[2].bias, h)
        h = _layernorm(h, self.post_norm)
        out = T.linear(h, self.proj_out.weight, self.proj_out.bias).view(B, K, -1).transpose(1, 2).contiguous()  # (B, 6, K)
        mean, logvar = torch.split(out, 3, dim=1)
        return mean, logvar


@ENCODERS.register_module()
class PartEncoderForTransformerDecoder(nn.Module):
    def __init__(self, encoder=None, n_class=4, part_aligner=None, fit_loss_weight=1.0, include_z=True, include_part_code=False,
                 include_params=False, use_gt_params=False, encode_ref=False, scale_var=1.0, fit_loss_type=0, origin_scale=False,
                 kl_weight=0.001, use_flow=False, latent_flow_depth=14, latent_flow_hidden_dim=256, gen=False, prior_var=1.0,
                 detach_params_in_ctx=False, selective_noise_sampling=False, selective_noise_sampling_global=False, **kwargs):
        super().__init__()
        unsupported = dict(encode_ref=encode_ref, selective_noise_sampling=selective_noise_sampling,
                           selective_noise_sampling_global=selective_noise_sampling_global, gen=not gen)
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError("difffacto_b200.PartEncoderForTransformerDecoder implements sample_latents of configs/gen_*.py; "
                                      f"unsupported setting(s): {bad}")
        self.zdim = int((encoder or {}).get("zdim", 256)) if isinstance(encoder, dict) else 256
        self.encoder = build_from_cfg(encoder, ENCODERS, num_anchors=n_class) if encoder is not None and encoder.get("type") in ENCODERS else None
        self.use_gt_params_cfg, self.origin_scale, self.kl_weight = use_gt_params, origin_scale, kl_weight
        self.kl_weight_annealing = kwargs.get("kl_weight_annealing", False)
        self.min_kl_weight = kwargs.get("min_kl_weight", 1e-7)
        self.kl_weight_annealing_end_epoch = kwargs.get("kl_weight_annealing_end_epoch", 3000)
        self.part_aligner = build_from_cfg(part_aligner, ENCODERS)
        self.n_class, self.prior_var = n_class, prior_var
        self.include_part_code, self.include_params = include_part_code, include_params
        self.log_scale_var = math.log(scale_var)
        self.use_flow = use_flow
        if use_flow:
            self.flow = nn.ModuleList([build_latent_flow(latent_flow_depth, latent_flow_hidden_dim, self.zdim) for _ in range(n_class)])

    # ---- training forward (stage 1: use_gt_params=True, no part aligner; reference :1185-1257) ----------------------------
    def get_prior_loss(self, part_code, mean, logvar, valid_id, epoch=-1):
        """reference :1143-1183.  The flows (Linear / ReLU / coupling with log-determinant) run on this repo's kernels; the
        O(B * 256) Gaussian log-likelihood / entropy bookkeeping around them is host-side torch, as in the reference."""
        B, C, M = part_code.shape
        entropy = 0.5 * logvar.reshape(B * self.n_class, -1).sum(dim=1) + 0.5 * float(C) * (1. + math.log(math.pi * 2))  # gaussian_entropy
        log_p_part = torch.zeros(B, self.n_class, device=part_code.device)
        for i in range(self.n_class):
            _id = valid_id[:, i] == 1
            if _id.any():
                pc = part_code[_id, :, i].contiguous()
                b = pc.shape[0]
                w, delta_log_pw = self.flow[i](pc, torch.zeros(b, 1, device=pc.device), reverse=False)
                log_z = -0.5 * math.log(2 * math.pi) * w.shape[1]                      # (sic) gaussian_log_likelihood(dim=w.shape[1]) per element
                log_pw = (-math.log(self.prior_var) + log_z - w.pow(2) / (2. * self.prior_var)).view(b, -1).sum(dim=1)
                log_p_part = log_p_part.index_put((_id.nonzero(as_tuple=True)[0], torch.full((b,), i, device=pc.device)), log_pw - delta_log_pw.view(b))
        entropy = entropy.view(B, self.n_class)
        loss_prior = ((-log_p_part - entropy) * valid_id).sum(1) / valid_id.sum(1)
        if self.kl_weight_annealing and self.kl_weight_annealing_end_epoch > epoch:
            kl_weight = self.min_kl_weight + (self.kl_weight - self.min_kl_weight) * epoch / self.kl_weight_annealing_end_epoch
        else:
            kl_weight = self.kl_weight
        out = {"prior_loss": kl_weight * loss_prior.mean(), "kl_weight": torch.ones(1, device=part_code.device) * kl_weight}
        mlog_p, ment = (log_p_part * valid_id).sum(0) / valid_id.sum(0), (entropy * valid_id).sum(0) / valid_id.sum(0)
        for i in range(self.n_class):
            out[f"log_p_part_{i}"], out[f"entropy_{i}"] = mlog_p[i], ment[i]
        return out

    def forward(self, pcds, device, noise=None, epoch=-1, eps=None):
        """Training forward of the stage-1 configuration (configs/train_chair_stage1.py: PointNetV2 encoder, latent flows as
        the prior, ground-truth part parameters): returns (ctx, mean_per_point, logvar_per_point, flag_per_point, loss_dict,
        [part_code, mean, logvar, noise]) like the reference (:1185-1257).  `eps` overrides the reparameterisation draw."""
        if self.encoder is None or self.part_aligner is not None or not self.use_gt_params_cfg:
            raise NotImplementedError("difffacto_b200: the training forward is built for the stage-1 configuration "
                                      "(encoder=PointNetV2, part_aligner=None, use_gt_params=True)")
        inp = pcds['input'].to(device)
        valid_id = pcds['present'].to(device).float()
        ref = pcds['ref'].to(device).transpose(1, 2)
        seg_mask = pcds['ref_seg_mask'].to(device).to(torch.int32)
        seg_flag = pcds['ref_attn_map'].to(device)
        B = ref.shape[0]
        gt_shift = pcds.get('part_shift', torch.zeros(B, 3, self.n_class)).to(device)
        gt_var = pcds.get('part_scale', torch.ones(B, 3, self.n_class)).to(device)
        if not self.origin_scale:
            gt_var = gt_var ** 2
        m, v = self.encoder(inp, seg_flag)                                            # (B, 4, 256) each
        if eps is None:
            eps = torch.randn(v.size()).to(m)
        part_code = (m + torch.exp(0.5 * v) * eps).transpose(1, 2)                    # reparameterize_gaussian (:282-285), (B, 256, 4)
        loss_dict = dict(self.get_prior_loss(part_code, m, v, valid_id, epoch=epoch))
        mean, logvar = gt_shift, torch.log(gt_var)                                    # use_gt_params (:456-458)
        mean_pp, logvar_pp, flag_pp = self.gather_all(seg_mask, anchors=mean, variances=logvar, valid_id=valid_id)
        loss_dict['fit_loss'] = torch.zeros(1, device=ref.device)
        ctx = self.prepare_ctx(part_code, mean, logvar, anchor_assignments=seg_mask)
        return ctx, mean_pp, logvar_pp + self.log_scale_var, flag_pp, loss_dict, [part_code, mean, logvar, noise]

    def gather_all(self, anchor_assignments, anchors=None, variances=None, valid_id=None):  # reference :417-428
        B, N = anchor_assignments.shape
        dev = anchor_assignments.device
        a = gather_operation(anchors.contiguous(), anchor_assignments).reshape(B, 3, N) if anchors is not None else torch.zeros(B, 3, N, device=dev)
        v = gather_operation(variances.contiguous(), anchor_assignments).reshape(B, 3, N) if variances is not None else torch.zeros(B, 3, N, device=dev)
        f = gather_operation(valid_id.unsqueeze(1).contiguous(), anchor_assignments).reshape(B, 1, N) if valid_id is not None \
            else torch.ones(B, 1, N, device=dev)
        return a, v, f

    def prepare_ctx(self, part_code, mean, logvar, **kwargs):  # reference :1317-1327
        ctx = []
        if self.include_part_code:
            ctx.append(part_code)
        if self.include_params:
            ctx.append(torch.cat([mean, _exp_shift(logvar, self.log_scale_var)], dim=1))
        return ctx

    @torch.no_grad()
    def sample_latents(self, sample_num, sample_points, device, fixed_id=None, valid_id=None, epoch=0, K=None, part_code=None, **kwargs):
        """reference :1052-1110 -> (ctx, mean_per_point, logvar_per_point, seg_mask, valid_id, [part_code, mean, logvar, noise])."""
        if part_code is None:
            part_code = torch.randn(sample_num, self.zdim, self.n_class).to(device)
            if self.prior_var != 1.0:
                part_code = _scale(part_code, math.sqrt(self.prior_var))
            if self.use_flow:
                part_code = torch.stack([self.flow[i](part_code[..., i], reverse=True).view(sample_num, self.zdim)
                                         for i in range(self.n_class)], dim=-1)
        if self.part_aligner is not None and self.part_aligner.cimle:
            K = 10 if K is None else K
            noise = torch.randn(sample_num * K, self.part_aligner.noise_dim).to(device)
            if self.part_aligner.cimle_start_epoch > epoch:
                noise = torch.zeros_like(noise)
        else:
            K, noise = 1, None
        if fixed_id is None:
            fixed_id = torch.zeros(self.n_class, device=device)
        if valid_id is None:
            valid_id = torch.ones(sample_num, self.n_class, device=device)
        fixed_id, valid_id = fixed_id.to(device).float(), valid_id.to(device).float()
        # part mixing with the first sample (pure selection arithmetic on 0/1 masks; reference :1072-1084)
        fixed_codes = part_code[0].unsqueeze(0)
        fixed_valid_id = (valid_id[0].unsqueeze(0) + fixed_id[None]).clamp(min=0, max=1)
        keep = (fixed_id == 0)
        part_code = torch.where(keep.view(1, 1, -1), part_code, fixed_codes.expand_as(part_code))
        valid_id = torch.where(keep.view(1, -1), valid_id, fixed_valid_id.expand_as(valid_id))
        if noise is not None and torch.any(fixed_id == 1):
            noise = noise.reshape(sample_num, K, -1)[0].unsqueeze(0).expand(sample_num, -1, -1).reshape(sample_num * K, -1)
        part_code = part_code.repeat_interleave(K, dim=0)
        valid_id = valid_id.repeat_interleave(K, dim=0)
        mean, logvar = self.part_aligner(part_code, valid_id, noise=noise)
        assert part_code.shape[-1] == self.n_class
        ar = torch.arange(self.n_class, device=device).unsqueeze(0)
        first = torch.argmax(valid_id, dim=1).unsqueeze(1)
        ids = torch.where(valid_id != 0, ar.expand_as(valid_id), first.expand_as(valid_id))  # = arange*valid + argmax*(1-valid) for 0/1 masks
        seg_mask = ids.to(torch.int32).unsqueeze(-1).expand(-1, -1, sample_points // self.n_class).reshape(sample_num * K, sample_points).contiguous()
        logvar_s = logvar if self.log_scale_var == 0.0 else logvar + self.log_scale_var
        mean_per_point, logvar_per_point, _ = self.gather_all(seg_mask, anchors=mean, variances=logvar_s)
        ctx = self.prepare_ctx(part_code, mean, logvar_s, anchor_assignments=seg_mask)  # (sic) the reference adds log_scale_var twice
        return ctx, mean_per_point, logvar_per_point, seg_mask, valid_id, [part_code, mean, logvar, noise]
