"""Differentiable building blocks of the denoiser's TRAINING path: torch.autograd.Function wrappers around the fp32
CUDA primitives of difffacto_b200/csrc/train_ops.cu (C ABI: dfb200_sgemm, dfb200_layernorm128_*, dfb200_geglu_*,
dfb200_part_attention_*, dfb200_dropout, dfb200_timestep_embedding, dfb200_q_sample_backward).

torch supplies the autograd graph, memory and pure data movement (cat / transpose / one_hot); every multiply-add of the
forward AND backward pass runs in this repo's kernels.  Composed by TransformerNet._forward_train
(models/diffusions/nets/attention.py), which follows the reference's module structure op for op
(python/difffacto/models/diffusions/nets/attention.py:50-57, 77-94, 161-204, 259-306, 385-440).
"""
import os

import torch
from torch.autograd import Function

from . import _lib
from ._lib import check, ptr, stream


def _c(t):
    return t.contiguous() if not t.is_contiguous() else t


class _ZeroArena:
    """Zero-initialised scratch for the backward pass.  Every weight / bias / LayerNorm / K-V gradient is accumulated with
    atomics (split-K wgrad, column sums) and used to cost one torch.zeros (allocation + fill launch) each: ~80 launches of a
    launch-bound step.  begin() zero-fills ONE flat buffer per training forward; zeros() hands out disjoint, never reused
    slices of it (bump pointer), and falls back to torch.zeros when the buffer is absent, on another device or exhausted."""
    buf = None
    off = 0


def begin_zero_arena(device, numel):
    _ZeroArena.buf = torch.zeros(int(numel), device=device, dtype=torch.float32)
    _ZeroArena.off = 0


def zeros(shape, device):
    n = 1
    for d in shape:
        n *= int(d)
    a = _ZeroArena
    if a.buf is not None and a.buf.device == device and a.off + n <= a.buf.numel():
        out = a.buf[a.off:a.off + n].view(shape)
        a.off += (n + 63) // 64 * 64  # keep every slice 256-byte aligned
        return out
    return torch.zeros(shape, device=device, dtype=torch.float32)


def _sgemm(a_kc, b_kc, M, N, K, A, lda, B, ldb, C, ldc, bias=None, beta=0, split_k=1, bf16=False):
    """fp32 CUDA-core GEMM, or - bf16=True and the problem is big enough for 128x128x64 tensor-core tiles - the tcgen05 GEMM
    with bf16-rounded operands and fp32 accumulation (same layout contract)."""
    tc = bf16 and M >= 128 and N >= 64 and K >= 64
    fn = "dfb200_gemm_bf16" if tc else "dfb200_sgemm"
    if split_k == 1 and K >= 256:
        # few output tiles and a long reduction (e.g. the K/V projections of the 4 part tokens: M = 4B, K = 522): split K over
        # more CTAs (atomic accumulation onto a zeroed / residual-initialised C)
        t = 128 if tc else 64
        tiles = ((M + t - 1) // t) * ((N + t - 1) // t)
        if tiles * 4 <= 148:
            split_k = max(1, min(K // 64, 148 // tiles))
            if split_k > 1 and not beta:
                C.zero_()
    with _lib.on(C.device):
        check(getattr(_lib.load(), fn)(int(a_kc), int(b_kc), M, N, K, ptr(A), lda, ptr(B), ldb, ptr(C), ldc, ptr(bias), int(beta),
                                       int(split_k), stream()))


def _dgrad(M, K, N, dy, weight, dx, bf16):
    """dx (M,K) = dy (M,N) . W (N,K).  On the tensor cores the weight is transposed first (a few hundred KB): with both operands
    k-contiguous the GEMM takes the asynchronous-copy kernel (csrc/gemm_tc.cu) instead of the register-staged one."""
    if bf16 and M >= 128 and K >= 64 and N >= 64 and N * K >= 65536:  # 128 x 128 weights: the transpose costs what the faster kernel saves
        wt = weight.t().contiguous()  # (K, N): B(n, k) = wt[k * N + n]
        _sgemm(True, True, M, K, N, dy, N, wt, N, dx, K, bf16=True)
    else:
        _sgemm(True, False, M, K, N, dy, N, weight, K, dx, K, bf16=bf16)


_GEMM_BF16 = False  # set by gemm_precision(): Linear layers built inside use the tensor cores (forward AND backward)


class gemm_precision:
    """with gemm_precision("bf16"): ... -> LinearFn created inside runs its GEMMs (fwd, dgrad, wgrad) on the tensor cores."""

    def __init__(self, precision):
        assert precision in ("fp32", "bf16")
        self.bf16 = precision == "bf16"

    def __enter__(self):
        global _GEMM_BF16
        self.prev, _GEMM_BF16 = _GEMM_BF16, self.bf16
        return self

    def __exit__(self, *a):
        global _GEMM_BF16
        _GEMM_BF16 = self.prev


class LinearFn(Function):
    """y = x W^T (+ b) (+ residual);  x (M,K), W (N,K) as in nn.Linear.  Backward: dx = dy W, dW = dy^T x (split-K over the
    M rows), db = column sums of dy; the residual receives dy unchanged."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual):
        x, weight = _c(x), _c(weight)
        M, K = x.shape
        N = weight.shape[0]
        if residual is not None:
            y = residual.contiguous().clone()
        else:
            y = torch.empty(M, N, device=x.device, dtype=torch.float32)
        ctx.bf16 = _GEMM_BF16
        _sgemm(True, True, M, N, K, x, K, weight, K, y, N, bias=bias, beta=residual is not None, bf16=ctx.bf16)
        ctx.save_for_backward(x, weight)
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = _c(dy)
        M, K = x.shape
        N = weight.shape[0]
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, device=x.device, dtype=torch.float32)
            _dgrad(M, K, N, dy, weight, dx, ctx.bf16)  # dx(i,k) = sum_n dy(i,n) W(n,k)
        if ctx.needs_input_grad[1]:
            dw = zeros((N, K), x.device)
            tc = ctx.bf16 and N >= 128 and K >= 64 and M >= 64
            t = 128 if tc else 64
            tiles = ((N + t - 1) // t) * ((K + t - 1) // t)
            split = max(1, min((M + 255) // 256, (148 * (2 if tc else 4) + tiles - 1) // tiles))
            _sgemm(False, False, N, K, M, dy, N, x, K, dw, K, split_k=split, bf16=ctx.bf16)  # dW(n,k) = sum_m dy(m,n) x(m,k)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = zeros((N,), x.device)
            with _lib.on(x.device):
                check(_lib.load().dfb200_colsum_accumulate(M, N, ptr(dy), N, ptr(db), stream()))
        return dx, dw, db, (dy if ctx.has_res and ctx.needs_input_grad[3] else None)


def linear(x, weight, bias=None, residual=None):
    return LinearFn.apply(x, weight, bias, residual)


class LayerNorm128Fn(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta):
        x = _c(x)
        M = x.shape[0]
        assert x.shape[1] == 128
        y = torch.empty_like(x)
        mean = torch.empty(M, device=x.device, dtype=torch.float32)
        rstd = torch.empty(M, device=x.device, dtype=torch.float32)
        with _lib.on(x.device):
            check(_lib.load().dfb200_layernorm128_forward(M, ptr(x), ptr(_c(gamma)), ptr(_c(beta)), ptr(y), ptr(mean), ptr(rstd), stream()))
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(x)
        dg = zeros((128,), x.device)
        db = zeros((128,), x.device)
        with _lib.on(x.device):
            check(_lib.load().dfb200_layernorm128_backward(x.shape[0], ptr(x), ptr(_c(gamma)), ptr(mean), ptr(rstd), ptr(dy), ptr(dx),
                                                           ptr(dg), ptr(db), stream()))
        return dx, dg, db


def layernorm128(x, gamma, beta):
    return LayerNorm128Fn.apply(x, gamma, beta)


_LN_RES = os.environ.get("DFB200_LN_RES", "1") != "0"


class LayerNorm128ResFn(Function):
    """(LayerNorm(x), x): the second output is x itself, to be used as the residual of the block the LayerNorm opens
    (x + branch(LayerNorm(x)), attention.py:296-306).  With x consumed by this ONE node, its two gradients (through the LayerNorm and
    through the residual) arrive together and are summed inside the LayerNorm backward kernel; through `layernorm128` autograd
    accumulates them with a separate add over the (M, 128) tensor (12 such passes per training step)."""

    @staticmethod
    def forward(ctx, x, gamma, beta):
        x = _c(x)
        M = x.shape[0]
        assert x.shape[1] == 128
        y = torch.empty_like(x)
        mean = torch.empty(M, device=x.device, dtype=torch.float32)
        rstd = torch.empty(M, device=x.device, dtype=torch.float32)
        with _lib.on(x.device):
            check(_lib.load().dfb200_layernorm128_forward(M, ptr(x), ptr(_c(gamma)), ptr(_c(beta)), ptr(y), ptr(mean), ptr(rstd), stream()))
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dres):
        x, gamma, mean, rstd = ctx.saved_tensors
        if dy is None:  # the normalised branch is unused: only the residual gradient flows
            return dres, None, None
        dx = torch.empty_like(x)
        dg = zeros((128,), x.device)
        db = zeros((128,), x.device)
        with _lib.on(x.device):
            check(_lib.load().dfb200_layernorm128_backward_residual(x.shape[0], ptr(x), ptr(_c(gamma)), ptr(mean), ptr(rstd), ptr(_c(dy)),
                                                                    ptr(_c(dres) if dres is not None else None), ptr(dx), ptr(dg), ptr(db),
                                                                    stream()))
        return dx, dg, db


def layernorm128_res(x, gamma, beta):
    """-> (LayerNorm(x), x) with the residual's gradient folded into the LayerNorm backward (LayerNorm128ResFn).
    DFB200_LN_RES=0 (A/B switch): plain LayerNorm128Fn, autograd accumulates the two gradients of x."""
    if not _LN_RES:
        return LayerNorm128Fn.apply(x, gamma, beta), x
    return LayerNorm128ResFn.apply(x, gamma, beta)


class GegluFn(Function):
    @staticmethod
    def forward(ctx, h):
        h = _c(h)
        M, H2 = h.shape
        u = torch.empty(M, H2 // 2, device=h.device, dtype=torch.float32)
        with _lib.on(h.device):
            check(_lib.load().dfb200_geglu_forward(M, H2 // 2, ptr(h), ptr(u), stream()))
        ctx.save_for_backward(h)
        return u

    @staticmethod
    def backward(ctx, du):
        (h,) = ctx.saved_tensors
        dh = torch.empty_like(h)
        with _lib.on(h.device):
            check(_lib.load().dfb200_geglu_backward(h.shape[0], h.shape[1] // 2, ptr(h), ptr(_c(du)), ptr(dh), stream()))
        return dh


def geglu(h):
    return GegluFn.apply(h)


_FUSED_FF_IN = os.environ.get("DFB200_FUSED_FF_IN", "1") != "0"  # A/B switch: 0 = GEMM, then the GEGLU + dropout kernel


class FFInFn(Function):
    """FeedForward's first half as ONE autograd node: u = dropout(geglu(x W1^T + b1)) (attention.py:77-94).  The GEGLU kernel
    applies the dropout mask (same Philox stream as DropoutFn) and, in the backward pass, also produces the bias gradient (column
    sums of dh), so neither the (M, H) activation nor the (M, 2H) gradient is read a second time."""

    @staticmethod
    def forward(ctx, x, weight, bias, p, seed, offset, counter):
        x, weight = _c(x), _c(weight)
        M, K = x.shape
        H2 = weight.shape[0]
        h = torch.empty(M, H2, device=x.device, dtype=torch.float32)
        u = torch.empty(M, H2 // 2, device=x.device, dtype=torch.float32)
        ctx.bf16 = _GEMM_BF16
        if ctx.bf16 and _FUSED_FF_IN and (H2 // 2) % 64 == 0 and K % 4 == 0 and M >= 128:
            # GEMM + GEGLU + dropout in one kernel (csrc/gemm_tc.cu, GEGLU epilogue): h is written once and not read back
            with _lib.on(x.device):
                check(_lib.load().dfb200_ff_in_forward(M, H2 // 2, K, ptr(x), K, ptr(weight), K, ptr(bias), float(p), int(seed), int(offset),
                                                       ptr(counter), ptr(h), ptr(u), stream()))
        else:
            _sgemm(True, True, M, H2, K, x, K, weight, K, h, H2, bias=bias, bf16=ctx.bf16)
            with _lib.on(x.device):
                check(_lib.load().dfb200_geglu_dropout_forward(M, H2 // 2, float(p), int(seed), int(offset), ptr(counter), ptr(h), ptr(u), stream()))
        ctx.save_for_backward(x, weight, h)
        ctx.args = (float(p), int(seed), int(offset))
        ctx.counter, ctx.has_bias = counter, bias is not None
        return u

    @staticmethod
    def backward(ctx, du):
        x, weight, h = ctx.saved_tensors
        p, seed, offset = ctx.args
        M, K = x.shape
        H2 = weight.shape[0]
        dh = torch.empty_like(h)
        db = zeros((H2,), x.device) if ctx.has_bias and ctx.needs_input_grad[2] else None
        with _lib.on(x.device):
            check(_lib.load().dfb200_geglu_dropout_backward(M, H2 // 2, p, seed, offset, ptr(ctx.counter), ptr(h), ptr(_c(du)), ptr(dh), ptr(db),
                                                            stream()))
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, device=x.device, dtype=torch.float32)
            _dgrad(M, K, H2, dh, weight, dx, ctx.bf16)
        if ctx.needs_input_grad[1]:
            dw = zeros((H2, K), x.device)
            tc = ctx.bf16 and H2 >= 128 and K >= 64 and M >= 64
            t = 128 if tc else 64
            tiles = ((H2 + t - 1) // t) * ((K + t - 1) // t)
            split = max(1, min((M + 255) // 256, (148 * (2 if tc else 4) + tiles - 1) // tiles))
            _sgemm(False, False, H2, K, M, dh, H2, x, K, dw, K, split_k=split, bf16=ctx.bf16)
        return dx, dw, db, None, None, None, None


def ff_in(x, weight, bias, p, training):
    """dropout(geglu(linear(x))) of FeedForward; the dropout key follows dropout()'s conventions (host generator, or the device
    step counter of a CUDA-graph step)."""
    global _dropout_calls, _step_calls
    if not training or p == 0.0:
        return FFInFn.apply(x, weight, bias, 0.0, 0, 0, None)
    if _step_counter is not None:
        _step_calls += 1
        return FFInFn.apply(x, weight, bias, p, _step_seed, _step_calls, _step_counter)
    _dropout_calls += 1
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return FFInFn.apply(x, weight, bias, p, seed, _dropout_calls, None)


class PartAttentionFn(Function):
    """softmax(q k^T / 4, masked) v over the 4 part tokens; q (B*N,128), k/v (B,4,128), valid (B,4) float or None."""

    @staticmethod
    def forward(ctx, q, k, v, valid, B, N):
        q, k, v = _c(q), _c(k), _c(v)
        o = torch.empty_like(q)
        probs = torch.empty(B * N, 32, device=q.device, dtype=torch.float32)
        with _lib.on(q.device):
            check(_lib.load().dfb200_part_attention_forward(B, N, ptr(q), ptr(k), ptr(v), ptr(valid), ptr(o), ptr(probs), stream()))
        ctx.save_for_backward(q, k, v, probs)
        ctx.valid, ctx.B, ctx.N = valid, B, N
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, k, v, probs = ctx.saved_tensors
        dq = torch.empty_like(q)
        dk = zeros(tuple(k.shape), k.device)
        dv = zeros(tuple(v.shape), v.device)
        with _lib.on(q.device):
            check(_lib.load().dfb200_part_attention_backward(ctx.B, ctx.N, ptr(q), ptr(k), ptr(v), ptr(ctx.valid), ptr(probs), ptr(_c(d_o)),
                                                             ptr(dq), ptr(dk), ptr(dv), stream()))
        return dq, dk, dv, None, None, None


def part_attention(q, k, v, valid, B, N):
    return PartAttentionFn.apply(q, k, v, valid, B, N)


_dropout_calls = 0
_step_counter = None   # device int64 tensor (1,) while a training step is captured / replayed as a CUDA graph (train_graph.py)
_step_seed = 0


def set_step_counter(counter, seed=0):
    """counter: a (1,) int64 CUDA tensor that the caller increments once per training step, or None (default).  While it is
    set, dropout() keys its masks by (seed, call index within the step, *counter) and touches neither the host generator nor
    the process-wide call count: the launch sequence is identical on every step, as a CUDA-graph replay needs."""
    global _step_counter, _step_seed, _step_calls
    _step_counter, _step_seed, _step_calls = counter, int(seed), 0


_step_calls = 0


def begin_step():
    """Resets the per-step dropout call index (called by TransformerNet._forward_train at the start of a forward pass)."""
    global _step_calls
    _step_calls = 0


class DropoutFn(Function):
    """y = dropout(x) (+ residual): inverted dropout with a Philox mask keyed by (seed, offset[, *step counter])."""

    @staticmethod
    def forward(ctx, x, p, seed, offset, residual, counter):
        x = _c(x)
        y = torch.empty_like(x)
        with _lib.on(x.device):
            if counter is None:
                check(_lib.load().dfb200_dropout(x.numel(), float(p), int(seed), int(offset), ptr(x),
                                                 ptr(None if residual is None else _c(residual)), ptr(y), stream()))
            else:
                check(_lib.load().dfb200_dropout_stepped(x.numel(), float(p), int(seed), int(offset), ptr(counter), ptr(x),
                                                         ptr(None if residual is None else _c(residual)), ptr(y), stream()))
        ctx.args = (float(p), int(seed), int(offset))
        ctx.counter = counter
        ctx.has_res = residual is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _c(dy)
        dx = torch.empty_like(dy)
        p, seed, offset = ctx.args
        with _lib.on(dy.device):
            if ctx.counter is None:
                check(_lib.load().dfb200_dropout(dy.numel(), p, seed, offset, ptr(dy), None, ptr(dx), stream()))
            else:  # the counter is incremented after the optimizer step: forward and backward of a step see the same value
                check(_lib.load().dfb200_dropout_stepped(dy.numel(), p, seed, offset, ptr(ctx.counter), ptr(dy), None, ptr(dx), stream()))
        return dx, None, None, None, (dy if ctx.has_res else None), None


def dropout(x, p, training, residual=None):
    """nn.Dropout (optionally fused with the residual add that follows it): identity in eval mode or for p == 0; the mask
    seed comes from torch's CPU generator (so torch.manual_seed makes a run reproducible), the offset counts the dropout
    calls of the process.  With a step counter set (CUDA-graph training, set_step_counter) the key is (seed, call index, *counter)."""
    global _dropout_calls, _step_calls
    if not training or p == 0.0:
        return x if residual is None else x + residual
    if _step_counter is not None:
        _step_calls += 1
        return DropoutFn.apply(x, p, _step_seed, _step_calls, residual, _step_counter)
    _dropout_calls += 1
    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
    return DropoutFn.apply(x, p, seed, _dropout_calls, residual, None)


def timestep_embedding(t, freqs):
    """(B,) float timesteps -> (B,256) sinusoid (no parameters, no gradient)."""
    t = _c(t.to(torch.float32))
    out = torch.empty(t.shape[0], 256, device=t.device, dtype=torch.float32)
    with _lib.on(t.device):
        check(_lib.load().dfb200_timestep_embedding(t.shape[0], ptr(t), ptr(freqs), ptr(out), stream()))
    return out


class QSampleFn(Function):
    """x_t = sqrt_ac (x0 - a) + a + sqrt_1mac sqrt(var) noise with gradients to x0, anchors and variance."""

    @staticmethod
    def forward(ctx, x_start, anchors, variance, noise, t_i32, sched, T):
        x_start, anchors, variance, noise = _c(x_start), _c(anchors), _c(variance), _c(noise)
        B, C, N = x_start.shape
        out = torch.empty_like(x_start)
        with _lib.on(x_start.device):
            check(_lib.load().dfb200_q_sample(B, N, T, ptr(sched), ptr(t_i32), ptr(x_start), ptr(anchors), ptr(variance), ptr(noise),
                                              ptr(out), stream()))
        ctx.save_for_backward(variance, noise, t_i32, sched)
        ctx.T = T
        return out

    @staticmethod
    def backward(ctx, g):
        variance, noise, t_i32, sched = ctx.saved_tensors
        g = _c(g)
        B, C, N = g.shape
        need = ctx.needs_input_grad
        dx0 = torch.empty_like(g) if need[0] else None
        da = torch.empty_like(g) if need[1] else None
        dv = torch.empty_like(g) if need[2] else None
        with _lib.on(g.device):
            check(_lib.load().dfb200_q_sample_backward(B, N, ctx.T, ptr(sched), ptr(t_i32), ptr(variance), ptr(noise), ptr(g), ptr(dx0),
                                                       ptr(da), ptr(dv), stream()))
        return dx0, da, dv, None, None, None, None


# ---- training-side encoder primitives (row f3) ---------------------------------------------------------------------------
class BatchNormFn(Function):
    """nn.BatchNorm1d over the M rows of a (M, C) matrix (+ fused ReLU).  training=True: batch statistics + running-stat
    update; else the running statistics.  Backward (training statistics only): dx, dgamma, dbeta."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, training, relu, momentum):
        x = _c(x)
        M, C = x.shape
        y = torch.empty_like(x)
        lib = _lib.load()
        if training:
            mean = torch.empty(C, device=x.device)
            rstd = torch.empty(C, device=x.device)
            scratch = torch.empty(2 * C, device=x.device)
            with _lib.on(x.device):
                check(lib.dfb200_batchnorm_forward(M, C, int(relu), ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd),
                                                   ptr(running_mean), ptr(running_var), float(momentum), ptr(scratch), stream()))
        else:
            mean = running_mean
            rstd = torch.rsqrt(running_var + 1e-5)  # C numbers of statistics bookkeeping
            with _lib.on(x.device):
                check(lib.dfb200_batchnorm_apply(M, C, int(relu), ptr(x), ptr(_c(mean)), ptr(rstd), ptr(gamma), ptr(beta), ptr(y), stream()))
        ctx.save_for_backward(x, y, gamma, mean, rstd)
        ctx.relu, ctx.training = relu, training
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, gamma, mean, rstd = ctx.saved_tensors
        if not ctx.training:
            raise NotImplementedError("difffacto_b200: BatchNorm backward is implemented for training-mode statistics")
        dy = _c(dy)
        M, C = x.shape
        dx = torch.empty_like(x)
        dg = torch.empty(C, device=x.device)
        db = torch.empty(C, device=x.device)
        with _lib.on(x.device):
            check(_lib.load().dfb200_batchnorm_backward(M, C, int(ctx.relu), ptr(x), ptr(y), ptr(dy), ptr(gamma), ptr(mean), ptr(rstd),
                                                        ptr(dx), ptr(dg), ptr(db), stream()))
        return dx, dg, db, None, None, None, None, None


def batchnorm(x, bn, relu=False):
    """x (M, C) through an nn.BatchNorm1d module `bn` (parameters + running statistics), optionally fused with ReLU."""
    training = bn.training or bn.running_mean is None
    if training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return BatchNormFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, training, relu, bn.momentum if bn.momentum is not None else 0.1)


class ReluFn(Function):
    @staticmethod
    def forward(ctx, x):
        y = _c(x).clone()
        with _lib.on(y.device):
            check(_lib.load().dfb200_relu(y.numel(), ptr(y), stream()))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dx = torch.empty_like(y)
        with _lib.on(y.device):
            check(_lib.load().dfb200_relu_backward(y.numel(), ptr(y), ptr(_c(dy)), ptr(dx), stream()))
        return dx


def relu(x):
    return ReluFn.apply(x)


class WeightedMaxPoolFn(Function):
    """out[b,c,a] = max_n x[b,n,c] * w[b,n,a] * scale (PointNetV2's anchor-weighted max-pool, no (B,C,N,A) intermediate)."""

    @staticmethod
    def forward(ctx, x, w, scale):
        x, w = _c(x), _c(w.to(torch.float32))
        B, N, C = x.shape
        A = w.shape[2]
        out = torch.empty(B, C, A, device=x.device)
        arg = torch.empty(B, C, A, dtype=torch.int32, device=x.device)
        with _lib.on(x.device):
            check(_lib.load().dfb200_weighted_maxpool_forward(B, N, C, A, float(scale), ptr(x), ptr(w), ptr(out), ptr(arg), stream()))
        ctx.save_for_backward(w, arg)
        ctx.dims, ctx.scale = (B, N, C, A), float(scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        w, arg = ctx.saved_tensors
        B, N, C, A = ctx.dims
        dx = torch.zeros(B, N, C, device=dout.device)
        with _lib.on(dout.device):
            check(_lib.load().dfb200_weighted_maxpool_backward(B, N, C, A, ctx.scale, ptr(w), ptr(arg), ptr(_c(dout)), ptr(dx), stream()))
        return dx, None, None


def weighted_maxpool(x, w, scale):
    return WeightedMaxPoolFn.apply(x, w, scale)


class CouplingForwardFn(Function):
    """y1 = x2 * sigmoid(s + 2) + shift, logdet = sum log sigmoid(s + 2);  s_t (B, 2d) = [s | shift], x2 (B, d)."""

    @staticmethod
    def forward(ctx, s_t, x2):
        s_t, x2 = _c(s_t), _c(x2)
        B, d = x2.shape
        y1 = torch.empty_like(x2)
        logdet = torch.empty(B, device=x2.device)
        with _lib.on(x2.device):
            check(_lib.load().dfb200_coupling_forward(B, d, ptr(s_t), ptr(x2), d, ptr(y1), d, ptr(logdet), stream()))
        ctx.save_for_backward(s_t, x2)
        return y1, logdet

    @staticmethod
    def backward(ctx, dy1, dlogdet):
        s_t, x2 = ctx.saved_tensors
        B, d = x2.shape
        ds_t = torch.empty_like(s_t)
        dx2 = torch.empty_like(x2)
        with _lib.on(x2.device):
            check(_lib.load().dfb200_coupling_backward(B, d, ptr(s_t), ptr(x2), d, ptr(_c(dy1)), d, ptr(_c(dlogdet)), ptr(ds_t), ptr(dx2), d,
                                                       stream()))
        return ds_t, dx2


def coupling_forward(s_t, x2):
    return CouplingForwardFn.apply(s_t, x2)
