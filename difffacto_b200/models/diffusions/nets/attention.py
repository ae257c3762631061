"""NETS['TransformerNet']: the cross-diffusion denoiser, host side.

Mirror of the reference class (python/difffacto/models/diffusions/nets/attention.py:308-440): same
constructor keywords, same parameter names/shapes (state_dicts interchange with the reference and
its pretrained checkpoints), same `forward(x, t, ctx, anchors=, variances=, valid_id=,
anchor_assignment=)` contract.  The modules below only HOLD parameters; the forward pass is the
hand-written CUDA of difffacto_b200/csrc/denoiser_*.cu reached through the C ABI
(dfb200_denoiser_forward).  There is no PyTorch fallback: CPU tensors raise.

When a gradient is required (training: `training_losses(...).backward()`), the forward runs as a
composition of the differentiable fp32 primitives of difffacto_b200/train_ops.py (each a
torch.autograd.Function over this repo's CUDA kernels), op for op like the reference modules.
"""
import math
import os

import torch
import torch.nn as nn

from .... import _lib
from .... import train_ops as T
from ...._lib import DenoiserCfg, check, ptr, stream
from ....utils.registry import NETS


class _GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class _FeedForward(nn.Module):  # reference attention.py:77-94 (glu=True): net = [GEGLU, Dropout, Linear]
    def __init__(self, dim, mult=4, dropout=0.0):
        super().__init__()
        inner = int(dim * mult)
        self.net = nn.Sequential(_GEGLU(dim, inner), nn.Dropout(dropout), nn.Linear(inner, dim))


class _CrossAttention(nn.Module):  # reference attention.py:161-177
    def __init__(self, query_dim, context_dim, heads, dim_head, dropout=0.0):
        super().__init__()
        inner = heads * dim_head
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(dropout))


class _Block(nn.Module):  # reference attention.py:259-294 with single_attn=True
    def __init__(self, dim, n_heads, d_head, dropout, context_dim):
        super().__init__()
        self.ff = _FeedForward(dim, dropout=dropout)
        self.attn2 = _CrossAttention(dim, context_dim, n_heads, d_head, dropout)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)


_GLOBAL_ORDER = ["pre_norm.weight", "pre_norm.bias", "post_norm.weight", "post_norm.bias", "proj_in.weight", "proj_in.bias",
                 "time_embed.net.0.proj.weight", "time_embed.net.0.proj.bias", "time_embed.net.2.weight",
                 "time_embed.net.2.bias", "proj_out.weight", "proj_out.bias"]
_BLOCK_ORDER = ["norm2.weight", "norm2.bias", "norm3.weight", "norm3.bias", "attn2.to_q.weight", "attn2.to_k.weight",
                "attn2.to_v.weight", "attn2.to_out.0.weight", "attn2.to_out.0.bias", "ff.net.0.proj.weight",
                "ff.net.0.proj.bias", "ff.net.2.weight", "ff.net.2.bias"]


def pack_order(depth):
    """Parameter names in the order dfb200_denoiser_pack expects (include/difffacto_b200.h)."""
    return _GLOBAL_ORDER + [f"transformer_blocks.{i}.{n}" for i in range(depth) for n in _BLOCK_ORDER]


@NETS.register_module()
class TransformerNet(nn.Module):
    def __init__(self, in_channels, n_heads, d_head, out_channels, depth=1, dropout=0., context_dim=None,
                 use_linear=False, use_checkpoint=False, single_attn=False, class_cond=False, n_class=4,
                 cat_params_to_x=False, mask_out_unreferenced_code=True, cat_class_to_x=False, use_sine_proj_in=False,
                 add_t_to_x=False, res=False, add_class_cond=False, context_proj=False, include_std=False,
                 precision=None):
        super().__init__()
        unsupported = dict(use_linear=not use_linear, single_attn=not single_attn, class_cond=not class_cond,
                           cat_params_to_x=not cat_params_to_x, cat_class_to_x=not cat_class_to_x,
                           use_sine_proj_in=use_sine_proj_in, add_t_to_x=add_t_to_x, res=res,
                           add_class_cond=add_class_cond, context_proj=context_proj)
        bad = [k for k, v in unsupported.items() if v]
        if bad or n_heads * d_head != 128:
            raise NotImplementedError(
                "difffacto_b200.TransformerNet implements the configuration of the shipped configs "
                f"(configs/gen_*.py, train_*.py); unsupported setting(s): {bad or 'inner_dim != 128'}")
        self.raw_in_channels = in_channels
        self.in_channels = in_channels + 6 + n_class
        self.out_channels = out_channels
        self.n_heads, self.d_head, self.depth = n_heads, d_head, depth
        self.n_class = n_class
        self.raw_context_dim = context_dim
        self.context_dim = context_dim + 256 + n_class
        self.inner_dim = n_heads * d_head
        self.mask_out_unreferenced_code = mask_out_unreferenced_code
        self.include_std = include_std
        self.class_cond, self.add_class_cond = class_cond, add_class_cond
        self.cat_params_to_x, self.cat_class_to_x = cat_params_to_x, cat_class_to_x
        self.add_t_to_x, self.res, self.context_proj, self.use_linear = add_t_to_x, res, context_proj, use_linear
        self.dropout = dropout

        self.pre_norm = nn.LayerNorm(self.inner_dim)
        self.post_norm = nn.LayerNorm(self.inner_dim)
        self.proj_in = nn.Linear(self.in_channels, self.inner_dim)
        self.time_embed = _FeedForward(256, dropout=dropout)
        self.transformer_blocks = nn.ModuleList(
            [_Block(self.inner_dim, n_heads, d_head, dropout, self.context_dim) for _ in range(depth)])
        self.proj_out = nn.Linear(self.inner_dim, out_channels)

        # "bf16": tcgen05 tensor cores (bf16 operands, fp32 accumulation), the throughput mode; "tf32": tcgen05 kind::tf32 with
        # every non-GEMM op in fp32, the tensor-core mode at reference tolerance; "fp32": CUDA-core path (reference numerics)
        self.precision = precision or os.environ.get("DFB200_PRECISION", "bf16")
        # training path: "fp32" (CUDA-core GEMMs, bit-faithful gradients) or "bf16" (tcgen05 GEMMs, mixed precision)
        self.train_precision = os.environ.get("DFB200_TRAIN_PRECISION", "fp32")
        self._packed = None
        self._packed_key = None
        self._workspace = None

    # ---- C-ABI plumbing -------------------------------------------------------------------
    def c_cfg(self):
        flags = _lib.NET_CLASS_COND | _lib.NET_CAT_PARAMS_TO_X | _lib.NET_CAT_CLASS_TO_X
        if self.mask_out_unreferenced_code:
            flags |= _lib.NET_MASK_UNREFERENCED
        if self.include_std:
            flags |= _lib.NET_INCLUDE_STD
        return DenoiserCfg(self.raw_in_channels, self.out_channels, self.n_heads, self.d_head, self.depth,
                           self.raw_context_dim, self.n_class, flags)

    def mode(self):
        if self.precision not in ("fp32", "bf16", "tf32"):
            raise ValueError(f"precision must be 'fp32', 'tf32' or 'bf16', got {self.precision!r}")
        return {"fp32": _lib.MODE_FP32, "bf16": _lib.MODE_BF16, "tf32": _lib.MODE_TF32}[self.precision]

    def packed_weights(self):
        """Device weight image for the kernels; rebuilt whenever a parameter was modified."""
        sd = dict(self.named_parameters())
        params = [sd[n] for n in pack_order(self.depth)]
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("CPU not supported: move the TransformerNet to a CUDA device")
        key = (dev,) + tuple((p.data_ptr(), p._version) for p in params)
        if self._packed is None or self._packed_key != key:
            lib = _lib.load()
            cfg = self.c_cfg()
            nbytes = lib.dfb200_denoiser_packed_bytes(cfg)
            if nbytes == 0:
                check(2)
            buf = torch.empty(nbytes + 1024, dtype=torch.uint8, device=dev)
            off = (-buf.data_ptr()) % 1024
            packed = buf[off:off + nbytes]
            tensors = [p.detach().to(torch.float32).contiguous() for p in params]
            # sinusoid frequencies exactly as the reference builds them (on the CPU, nets/utils.py:17-19)
            freqs = torch.exp(-math.log(10000) * torch.arange(start=0, end=128, dtype=torch.float32) / 128).to(dev)
            tensors.append(freqs)
            arr = (_lib.P * len(tensors))(*[t.data_ptr() for t in tensors])
            with _lib.on(dev):
                check(lib.dfb200_denoiser_pack(cfg, arr, len(tensors), ptr(packed), stream()))
                torch.cuda.current_stream().synchronize()
            self._packed, self._packed_key, self._packed_buf = packed, key, buf
        return self._packed

    def workspace(self, nbytes, dev):
        if self._workspace is None or self._workspace.numel() < nbytes or self._workspace.device != dev:
            self._workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        return self._workspace

    # ---- forward ----------------------------------------------------------------------------
    def forward(self, x, t, ctx, anchors=None, variances=None, valid_id=None, anchor_assignment=None, **kwargs):
        """x (B,3,N); t (B,); ctx list of (B,C,4) or a (B,262,4) tensor; anchors/variances (B,N,3)
        (the reference passes `.transpose(1,2)` views of (B,3,N) tensors); valid_id (B,4);
        anchor_assignment (B,N) int32.  Returns eps (B,3,N).  Reference attention.py:385-440."""
        if isinstance(ctx, (list, tuple)):
            ctx = torch.cat(list(ctx), dim=1)
        if torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or
                                        any(torch.is_tensor(a) and a.requires_grad for a in (x, ctx, anchors, variances))):
            # precision "bf16": the large Linear layers run on the tensor cores (bf16 operands, fp32 accumulation); everything
            # else (LayerNorm, attention core, GEGLU, small GEMMs) stays fp32
            with T.gemm_precision(getattr(self, "train_precision", None) or "fp32"):
                return self._forward_train(x, t, ctx, anchors, variances, valid_id, anchor_assignment)
        _lib.require_cuda(x, ctx, anchors, variances, anchor_assignment, valid_id)
        B, C, N = x.shape
        assert C == self.raw_in_channels
        assert ctx.shape[1] + 256 + self.n_class == self.context_dim and ctx.shape[2] == self.n_class
        dev = x.device
        f32 = torch.float32
        x = x.to(f32).contiguous()
        ctx = ctx.to(f32).contiguous()
        anchors_cm = anchors.transpose(1, 2).to(f32).contiguous()      # back to (B,3,N) channel-major
        variances_cm = variances.transpose(1, 2).to(f32).contiguous()
        assign = anchor_assignment.to(torch.int32).contiguous()
        valid = None
        if self.mask_out_unreferenced_code and valid_id is not None:
            assert valid_id.shape == (B, self.n_class)
            valid = valid_id.to(f32).contiguous()
        tf = t.to(f32).contiguous()
        assert tf.shape == (B,)
        eps = torch.empty(B, self.out_channels, N, dtype=f32, device=dev)
        lib = _lib.load()
        cfg, mode = self.c_cfg(), self.mode()
        packed = self.packed_weights()
        nws = lib.dfb200_denoiser_workspace_bytes(cfg, mode, B, N)
        ws = self.workspace(nws, dev)
        with _lib.on(dev):
            check(lib.dfb200_denoiser_forward(cfg, ptr(packed), mode, B, N, ptr(x), ptr(tf), ptr(ctx), ptr(anchors_cm),
                                              ptr(variances_cm), ptr(assign), ptr(valid), ptr(eps), ptr(ws), nws, stream()))
        return eps

    # ---- training forward (differentiable) ----------------------------------------------------
    def _ff(self, ff, x, residual=None):
        """FeedForward (glu=True): Linear -> GEGLU -> Dropout -> Linear (reference attention.py:77-94)."""
        u = T.ff_in(x, ff.net[0].proj.weight, ff.net[0].proj.bias, self.dropout, self.training)  # Linear + GEGLU + Dropout, one node
        return T.linear(u, ff.net[2].weight, ff.net[2].bias, residual)

    def _freqs_on(self, dev):
        """The 128 sinusoid frequencies of timestep_embedding (nets/utils.py:7-24), resident per device (a host-side build +
        upload inside the forward would not be capturable in a CUDA graph)."""
        cache = self.__dict__.setdefault("_freqs_cache", {})
        if dev not in cache:
            cache[dev] = torch.exp(-math.log(10000) * torch.arange(start=0, end=128, dtype=torch.float32) / 128).to(dev)
        return cache[dev]

    def _forward_train(self, x, t, ctx, anchors, variances, valid_id, anchor_assignment):
        """Reference attention.py:385-440 (+ :296-306, :179-204) on the differentiable primitives; fp32."""
        _lib.require_cuda(x, ctx, anchors, variances, anchor_assignment, valid_id)
        if self.include_std:
            raise NotImplementedError("difffacto_b200.TransformerNet: include_std is not supported on the training path")
        B, C, N = x.shape
        f32 = torch.float32
        dev = x.device
        T.begin_step()
        if torch.is_grad_enabled():  # one zero fill for all atomically accumulated gradients of this step's backward pass
            T.begin_zero_arena(dev, sum(p.numel() for p in self.parameters()) + 2 * self.depth * B * self.n_class * 128 + 64 * 160)
        # context (B,4,522) = [part code | mean, var | one-hot class | t_embed]  (:389-398)
        ctx = ctx.to(f32).transpose(1, 2)
        class_embed = torch.eye(self.n_class, device=dev, dtype=f32).unsqueeze(0).expand(B, -1, -1)
        freqs = self._freqs_on(dev)
        t_embed = self._ff(self.time_embed, T.timestep_embedding(t, freqs))
        ctx = torch.cat([ctx, class_embed, t_embed.unsqueeze(1).expand(-1, self.n_class, -1)], dim=-1).contiguous()
        ctx2d = ctx.reshape(B * self.n_class, self.context_dim)
        # point features (B*N,13) = [x | anchors | variances | one-hot part]  (:400-407)
        onehot = torch.nn.functional.one_hot(anchor_assignment.long(), num_classes=self.n_class).to(f32)
        feat = torch.cat([x.to(f32).transpose(1, 2), anchors.to(f32), variances.to(f32), onehot], dim=-1).reshape(B * N, self.in_channels)
        valid = None
        if self.mask_out_unreferenced_code and valid_id is not None:
            valid = valid_id.to(f32).contiguous()
        h = T.linear(feat.contiguous(), self.proj_in.weight, self.proj_in.bias)
        h = T.layernorm128(h, self.pre_norm.weight, self.pre_norm.bias)
        # K / V of the 4 part tokens for ALL blocks in one GEMM (the context does not change over the blocks, :179-182): the
        # 2 x depth projections of a (4B, 522) matrix were 10 launches of a single-wave kernel, forward, dgrad and wgrad alike;
        # autograd's cat / slicing hand the gradient of the stacked weight back to the per-block parameters.
        wkv = torch.cat([w for blk in self.transformer_blocks for w in (blk.attn2.to_k.weight, blk.attn2.to_v.weight)], dim=0)
        kv_all = T.linear(ctx2d, wkv).view(B * self.n_class, 2 * len(self.transformer_blocks), self.inner_dim)
        kv_parts = kv_all.transpose(0, 1).contiguous().unbind(0)  # 2 * depth tensors (4B, 128); backward = one stack
        for li, blk in enumerate(self.transformer_blocks):
            a, h = T.layernorm128_res(h, blk.norm2.weight, blk.norm2.bias)  # h: the same values, as the residual of this sub-block
            q = T.linear(a, blk.attn2.to_q.weight)
            k = kv_parts[2 * li].view(B, self.n_class, self.inner_dim)
            v = kv_parts[2 * li + 1].view(B, self.n_class, self.inner_dim)
            o = T.part_attention(q, k, v, valid, B, N)
            if self.training and self.dropout > 0:  # to_out = [Linear, Dropout], then the residual add (:203, :303)
                h = T.dropout(T.linear(o, blk.attn2.to_out[0].weight, blk.attn2.to_out[0].bias), self.dropout, True, residual=h)
            else:
                h = T.linear(o, blk.attn2.to_out[0].weight, blk.attn2.to_out[0].bias, h)
            f, h = T.layernorm128_res(h, blk.norm3.weight, blk.norm3.bias)
            h = self._ff(blk.ff, f, residual=h)
        h = T.layernorm128(h, self.post_norm.weight, self.post_norm.bias)
        out = T.linear(h, self.proj_out.weight, self.proj_out.bias)
        return out.view(B, N, self.out_channels).transpose(1, 2).contiguous()
