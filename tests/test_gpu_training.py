"""-m gpu: the differentiable training path (difffacto_b200/train_ops.py over csrc/train_ops.cu) against torch autograd
run over the oracle's PyTorch port of the reference on the CPU.

Tolerances: forward eps 1e-4 abs; gradients 2e-3 relative to the largest entry of each tensor (fp32 accumulation order;
weight gradients are reductions over B*N tokens with split-K atomics)."""
import numpy as np
import pytest
import torch

from oracle import denoiser_ref as R

pytestmark = pytest.mark.gpu
from test_gpu_denoiser import DIFF_CFG  # noqa: E402


def _build(T=100, dropout=None):
    import difffacto_b200 as D
    cfg = dict(DIFF_CFG)
    if dropout is not None:
        cfg["net"] = dict(cfg["net"], dropout=dropout)
    d = D.build_from_cfg(cfg, D.DIFFUSIONS, num_timesteps=T)
    d.model.load_state_dict(R.synthetic_state_dict(1234), strict=True)
    return d.cuda()


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


@pytest.mark.parametrize("case", [(11, 3, 64, False), (12, 2, 256, True)])
def test_training_loss_and_gradients_match_autograd_of_the_port(case):
    seed, B, N, av = case
    d = _build().eval()  # eval(): dropout off (the reference's dropout consumes torch's RNG and cannot be reproduced)
    inp = R.synthetic_inputs(seed, B, N, av)
    s = R.schedule(100)
    # ---- reference side: torch autograd over the CPU port ----
    sd = {k: v.clone().requires_grad_(True) for k, v in R.synthetic_state_dict(1234).items()}
    leaves = {k: inp[k].clone().requires_grad_(True) for k in ("x", "anchors", "variance", "code", "params")}
    x_t = R.q_sample(s, leaves["x"], inp["t"], leaves["anchors"], leaves["variance"], inp["noise"])
    eps = R.denoiser_forward(sd, x_t, inp["t"], [leaves["code"], leaves["params"]], leaves["anchors"], leaves["variance"], inp["valid"],
                             inp["assign"])
    flags = torch.ones(B, 1, N)
    loss_ref = (((inp["noise"] - eps) ** 2) * flags).mean(1).sum() / flags.sum()
    loss_ref.backward()
    # ---- ours ----
    g = {k: inp[k].cuda().requires_grad_(True) for k in ("x", "anchors", "variance", "code", "params")}
    out = d.training_losses(g["x"], inp["t"].cuda(), anchors=g["anchors"], variance=g["variance"], ctx=[g["code"], g["params"]],
                            anchor_assignment=inp["assign"].cuda(), valid_id=inp["valid"].cuda(), flags=flags.cuda(),
                            noise=inp["noise"].cuda())
    loss = out["mse_loss"]
    assert abs(loss.item() - loss_ref.item()) < 1e-4 * max(1.0, abs(loss_ref.item()))
    loss.backward()
    for k in g:
        assert g[k].grad is not None, k
        assert _rel(g[k].grad.cpu(), leaves[k].grad) < 2e-3, (k, _rel(g[k].grad.cpu(), leaves[k].grad))
    ours = dict(d.model.named_parameters())
    assert set(ours) == set(sd)
    worst = max((_rel(ours[k].grad.cpu(), sd[k].grad), k) for k in sd)
    assert worst[0] < 2e-3, worst
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in d.model.parameters())


def test_training_forward_equals_inference_kernels():
    d = _build().eval()
    i = {k: v.cuda() for k, v in R.synthetic_inputs(12, 2, 256, True).items()}
    kw = dict(anchors=i["anchors"].transpose(1, 2), anchor_assignment=i["assign"], variances=i["variance"].transpose(1, 2), valid_id=i["valid"])
    d.model.precision = "fp32"
    with torch.no_grad():
        ref = d.model(i["x"], i["t"], [i["code"], i["params"]], **kw)
    tr = d.model(i["x"], i["t"], [i["code"], i["params"]], **kw)  # grad enabled -> composed differentiable path
    assert tr.requires_grad and (tr - ref).abs().max().item() < 1e-4


def test_sgemm_layouts_and_split_k():
    from difffacto_b200 import train_ops as T
    torch.manual_seed(0)
    for (M, N, K) in [(70, 13, 131), (300, 128, 522), (1000, 3, 128), (257, 1024, 128)]:
        x = torch.randn(M, K, device="cuda", requires_grad=True)
        w = torch.randn(N, K, device="cuda", requires_grad=True)
        b = torch.randn(N, device="cuda", requires_grad=True)
        r = torch.randn(M, N, device="cuda", requires_grad=True)
        y = T.linear(x, w, b, r)
        yr = torch.nn.functional.linear(x.double(), w.double(), b.double()) + r.double()
        assert (y.double() - yr).abs().max().item() < 1e-3
        go = torch.randn(M, N, device="cuda")
        y.backward(go)
        gx, gw, gb = torch.autograd.grad(yr, (x, w, b), go.double())
        assert _rel(x.grad.double(), gx) < 1e-5 and _rel(w.grad.double(), gw) < 1e-5 and _rel(b.grad.double(), gb) < 1e-5
        assert torch.equal(r.grad, go)


def test_dropout_mask_statistics_and_backward():
    from difffacto_b200 import train_ops as T
    torch.manual_seed(3)
    x = torch.ones(1 << 20, device="cuda", requires_grad=True)
    y = T.dropout(x, 0.2, True)
    kept = (y != 0).float().mean().item()
    assert abs(kept - 0.8) < 5e-3 and torch.allclose(y[y != 0], torch.tensor(1.25, device="cuda"))
    y.sum().backward()
    assert torch.equal(x.grad != 0, y.detach() != 0)         # the backward pass regenerates the same mask
    assert T.dropout(x, 0.2, False) is x and T.dropout(x, 0.0, True) is x
    d = _build(dropout=0.2).train()
    i = {k: v.cuda() for k, v in R.synthetic_inputs(12, 2, 256, True).items()}
    torch.manual_seed(7)
    a = d.training_losses(i["x"], i["t"], anchors=i["anchors"], variance=i["variance"], ctx=[i["code"], i["params"]],
                          anchor_assignment=i["assign"], valid_id=i["valid"], flags=torch.ones(2, 1, 256, device="cuda"), noise=i["noise"])
    a["mse_loss"].backward()
    assert torch.isfinite(a["mse_loss"]) and all(torch.isfinite(p.grad).all() for p in d.model.parameters())


def test_tensor_core_gemm_layouts_match_bf16_reference():
    """dfb200_gemm_bf16 (tcgen05) in the four operand layouts, edge tiles, bias, accumulate and split-K.  k-contiguous operands take
    the asynchronous-copy kernel (fp32 containers read as tf32: 10 mantissa bits, truncated), the other layouts the register-staged
    kernel (operands rounded to bf16: 8 bits): both are checked against the exact fp64 product with the bf16 operand-rounding bound
    (measured: 1.2e-2 sqrt(K) for bf16, 4e-3 sqrt(K) for tf32 on N(0,1) operands; a layout bug gives O(sqrt(K)))."""
    from difffacto_b200 import train_ops as T
    torch.manual_seed(1)
    d64 = lambda t: t.double()  # noqa: E731
    for (M, N, K) in [(128, 128, 64), (300, 200, 136), (1000, 128, 512), (257, 1024, 128), (4100, 260, 100)]:
        x, w, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda"), torch.randn(N, device="cuda")
        c0 = torch.randn(M, N, device="cuda")
        y = c0.clone()
        T._sgemm(True, True, M, N, K, x, K, w, K, y, N, bias=b, beta=1, bf16=True)
        ref = d64(c0) + d64(x) @ d64(w).t() + d64(b)
        assert (d64(y) - ref).abs().max().item() < 6e-3 * K ** 0.5, (M, N, K)   # tf32 path (or bf16 if K is not 16-byte aligned)
        y3 = torch.zeros(M, N, device="cuda")
        T._sgemm(True, True, M, N, K, x, K, w, K, y3, N, split_k=3, bf16=True)  # split-K on the same path
        assert (d64(y3) - d64(x) @ d64(w).t()).abs().max().item() < 6e-3 * K ** 0.5, (M, N, K)
        dy = torch.randn(M, N, device="cuda")
        dx = torch.empty(M, K, device="cuda")
        T._sgemm(True, False, M, K, N, dy, N, w, K, dx, K, bf16=True)                 # dgrad layout
        assert (d64(dx) - d64(dy) @ d64(w)).abs().max().item() < 1.8e-2 * N ** 0.5
        dx2 = torch.empty(M, K, device="cuda")
        T._dgrad(M, K, N, dy, w, dx2, True)                                            # dgrad through the transposed weight
        assert (d64(dx2) - d64(dy) @ d64(w)).abs().max().item() < 1.8e-2 * N ** 0.5
        dw = torch.zeros(N, K, device="cuda")
        T._sgemm(False, False, N, K, M, dy, N, x, K, dw, K, split_k=3, bf16=True)     # wgrad layout, split-K with atomics
        assert (d64(dw) - d64(dy).t() @ d64(x)).abs().max().item() < 1.8e-2 * M ** 0.5
        xt = x.t().contiguous()
        y2 = torch.empty(M, N, device="cuda")
        T._sgemm(False, True, M, N, K, xt, M, w, K, y2, N, bf16=True)                 # A row-contiguous, B k-contiguous
        assert (d64(y2) - d64(x) @ d64(w).t()).abs().max().item() < 1.8e-2 * K ** 0.5


def test_bf16_training_gradients_close_to_fp32():
    """precision bf16 on the training path: loss within 2e-3, gradients within 5 % (relative to each tensor's largest entry) of
    the fp32 path - the error of bf16 operand rounding, not of the kernels (layouts are checked exactly above)."""
    d = _build().eval()
    inp = {k: v.cuda() for k, v in R.synthetic_inputs(12, 2, 256, True).items()}
    flags = torch.ones(2, 1, 256, device="cuda")
    res = {}
    for prec in ("fp32", "bf16"):
        d.model.train_precision = prec
        d.zero_grad(set_to_none=True)
        loss = d.training_losses(inp["x"], inp["t"], anchors=inp["anchors"], variance=inp["variance"], ctx=[inp["code"], inp["params"]],
                                 anchor_assignment=inp["assign"], valid_id=inp["valid"], flags=flags, noise=inp["noise"])["mse_loss"]
        loss.backward()
        res[prec] = (loss.item(), {k: p.grad.clone() for k, p in d.model.named_parameters()})
    assert abs(res["bf16"][0] - res["fp32"][0]) < 2e-3 * max(1.0, abs(res["fp32"][0]))
    worst = max((_rel(res["bf16"][1][k], res["fp32"][1][k]), k) for k in res["fp32"][1])
    assert worst[0] < 5e-2, worst


def _graph_inputs(seed, B, N):
    inp = R.synthetic_inputs(seed, B, N, False)
    return dict(x0=inp["x"].cuda(), t=inp["t"].cuda(), anchors=inp["anchors"].cuda(), variance=inp["variance"].cuda(), code=inp["code"].cuda(),
                params=inp["params"].cuda(), assign=inp["assign"].cuda(), valid=inp["valid"].cuda(), flags=torch.ones(B, 1, N).cuda(),
                noise=inp["noise"].cuda())


def _loss_fn(d):
    def f(x0, t, anchors, variance, code, params, assign, valid, flags, noise):
        return d.training_losses(x0, t, anchors=anchors, variance=variance, ctx=[code, params], anchor_assignment=assign, valid_id=valid,
                                 flags=flags, noise=noise)["mse_loss"]
    return f


def test_cuda_graph_training_step_equals_the_eager_step():
    """difffacto_b200/train_graph.py: forward + backward + Adam replayed as ONE CUDA graph gives the losses and the weights of
    the eager loop (dropout off so both draw no masks; the warm-up steps of the capture are part of the sequence on both sides)."""
    from difffacto_b200.train_graph import GraphedTrainStep
    B, N, steps, warm = 2, 256, 4, 3
    batches = [_graph_inputs(40 + i, B, N) for i in range(1 + steps)]
    # eager reference
    d0 = _build(dropout=0.0).train()
    # SGD with momentum: Adam normalises noise-level gradients to +-lr, so the summation order of the split-K atomics would show up
    # as O(lr) weight differences that have nothing to do with the capture
    opt0 = torch.optim.SGD(d0.parameters(), lr=0.05, momentum=0.9)
    # the capture runs `warm` eager steps on the example batch (the capture itself only records, it does not execute)
    losses0 = []
    for i in range(warm):
        opt0.zero_grad(set_to_none=True)
        l = _loss_fn(d0)(**batches[0]); l.backward(); opt0.step()
    for i in range(steps):
        opt0.zero_grad(set_to_none=True)
        l = _loss_fn(d0)(**batches[1 + i]); l.backward(); opt0.step()
        losses0.append(l.item())
    # graphed
    d1 = _build(dropout=0.0).train()
    opt1 = torch.optim.SGD(d1.parameters(), lr=0.05, momentum=0.9)
    step = GraphedTrainStep(_loss_fn(d1), list(d1.parameters()), opt1, batches[0], warmup=warm)
    losses1 = [step(**batches[1 + i]).item() for i in range(steps)]
    assert np.allclose(losses0, losses1, rtol=2e-4, atol=1e-6), (losses0, losses1)
    for (k, p0), (_, p1) in zip(d0.model.named_parameters(), d1.model.named_parameters()):
        assert _rel(p1.detach(), p0.detach()) < 2e-3, k  # split-K atomics: summation order differs run to run


def test_cuda_graph_training_step_draws_a_fresh_dropout_mask_per_replay():
    """With dropout on, the replayed graph reads the device step counter: the same batch gives a different loss on every replay
    (same weights up to a tiny Adam step), and the counter advances by one per step."""
    from difffacto_b200.train_graph import GraphedTrainStep
    d = _build(dropout=0.2).train()
    opt = torch.optim.Adam(d.parameters(), lr=1e-8, capturable=True)
    b = _graph_inputs(77, 2, 256)
    step = GraphedTrainStep(_loss_fn(d), list(d.parameters()), opt, b, warmup=3)
    c0 = int(step.counter.item())
    losses = [step(**b).item() for _ in range(4)]
    assert int(step.counter.item()) == c0 + 4
    assert all(np.isfinite(losses)) and len({round(l, 7) for l in losses}) == 4, losses


def test_fused_ff_in_equals_linear_geglu_dropout_chain():
    """FFInFn (Linear + GEGLU + Dropout as one node, bias gradient from the GEGLU backward) against the unfused chain
    linear -> geglu -> DropoutFn with the same Philox key: outputs and all gradients."""
    from difffacto_b200 import train_ops as T
    torch.manual_seed(5)
    M, K, H = 1000, 128, 512
    x = torch.randn(M, K, device="cuda")
    w = (torch.randn(2 * H, K, device="cuda") * 0.1)
    b = torch.randn(2 * H, device="cuda") * 0.1
    dy = torch.randn(M, H, device="cuda")
    for p, seed, off in ((0.0, 0, 0), (0.3, 1234, 7)):
        a = [t.clone().requires_grad_(True) for t in (x, w, b)]
        u0 = T.geglu(T.linear(a[0], a[1], a[2]))
        if p > 0:
            u0 = T.DropoutFn.apply(u0, p, seed, off, None, None)
        u0.backward(dy)
        c = [t.clone().requires_grad_(True) for t in (x, w, b)]
        u1 = T.FFInFn.apply(c[0], c[1], c[2], p, seed, off, None)
        u1.backward(dy)
        assert torch.allclose(u1, u0, rtol=1e-5, atol=1e-6)
        if p > 0:
            kept = (u1 != 0).float().mean().item()
            assert abs(kept - (1 - p)) < 0.01, kept
        for g1, g0, name in zip(c, a, "xwb"):
            assert _rel(g1.grad, g0.grad) < 2e-5, (p, name, _rel(g1.grad, g0.grad))


def test_fused_adam_matches_torch_adam_and_is_graph_capturable():
    """difffacto_b200/optim.py FusedAdam (one launch for the whole group, dfb200_adam_step) against torch.optim.Adam on odd-sized
    tensors (vector and scalar tails, a tensor spanning several 2 048-element chunks, an unaligned view), with weight decay; then
    the same under CUDA-graph replay, and a state_dict round trip into torch.optim.Adam."""
    from difffacto_b200.optim import FusedAdam
    torch.manual_seed(11)
    shapes = [(7,), (128, 33), (5000,), (3, 3), (1,), (2049,)]
    base = torch.randn(10001, device="cuda")
    def make():
        ps = [torch.nn.Parameter(torch.randn(*s, device="cuda", generator=torch.Generator("cuda").manual_seed(i))) for i, s in enumerate(shapes)]
        return ps
    p0, p1 = make(), make()
    kw = dict(lr=3e-3, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.01)
    o0, o1 = torch.optim.Adam(p0, **kw), FusedAdam(p1, **kw)
    for it in range(5):
        for a, b in zip(p0, p1):
            g = torch.randn_like(a) * (0.1 + it)
            a.grad, b.grad = g.clone(), g.clone()
        o0.step(); o1.step()
    for a, b in zip(p0, p1):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (a - b).abs().max()
    assert int(o1.state[p1[0]]["step"].item()) == 5 and o1.state[p1[0]]["step"] is o1.state[p1[3]]["step"]
    # grad_scale = folded clipping coefficient
    sc = torch.tensor(0.25, device="cuda")
    for a, b in zip(p0, p1):
        g = torch.randn_like(a)
        a.grad, b.grad = g * 0.25, g.clone()
    o0.step(); o1.step(grad_scale=sc)
    for a, b in zip(p0, p1):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7)
    # checkpoint goes both ways
    o2 = torch.optim.Adam(p1, **kw)
    o2.load_state_dict(o1.state_dict())
    o3 = FusedAdam(p0, **kw)
    o3.load_state_dict(o0.state_dict())
    for a, b in zip(p0, p1):
        g = torch.randn_like(a)
        a.grad, b.grad = g.clone(), g.clone()
    o3.step(); o2.step()
    for a, b in zip(p0, p1):
        assert torch.allclose(a, b, rtol=4e-6, atol=4e-7)
    # CUDA graph: the step count is read from the device at replay time
    q0, q1 = make(), make()
    e0, e1 = torch.optim.Adam(q0, **kw), FusedAdam(q1, **kw)
    gs = [torch.randn_like(a) for a in q1]
    for b, g in zip(q1, gs):
        b.grad = g.clone()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        e1.step()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        e1.step()
    for _ in range(3):
        graph.replay()
    for _ in range(4):  # 1 eager + 3 replays (the capture itself does not execute)
        for a, g in zip(q0, gs):
            a.grad = g.clone()
        e0.step()
    for a, b in zip(q0, q1):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7)
    with pytest.raises(NotImplementedError):
        FusedAdam(q1, amsgrad=True)


@pytest.mark.parametrize("M,N,ld", [(5000, 128, 128), (4097, 132, 136), (1024, 4, 4), (100, 7, 7), (3000, 130, 131)])
def test_colsum_accumulate_vector_and_scalar_forms(M, N, ld):
    """dfb200_colsum_accumulate (bias gradients): out[j] += sum_i X[i, j] - the 128-bit form (N, ld multiples of 4, M >= 1024) and the
    scalar form, on a strided matrix, accumulating into a non-zero output."""
    from difffacto_b200 import _lib
    torch.manual_seed(M + N)
    X = torch.randn(M, ld, device="cuda")
    out = torch.randn(N, device="cuda")
    want = out.double() + X[:, :N].double().sum(0)
    _lib.check(_lib.load().dfb200_colsum_accumulate(M, N, _lib.ptr(X), ld, _lib.ptr(out), _lib.stream()))
    assert torch.allclose(out.double(), want, rtol=1e-5, atol=2e-4 * M ** 0.5)


def test_layernorm_residual_node_equals_layernorm_plus_autograd_accumulation():
    """T.layernorm128_res returns (LayerNorm(x), x): used as `x + f(LN(x))` it must give the values and gradients of the two-consumer
    form (LayerNorm128Fn + autograd's accumulation), with either output unused as well."""
    from difffacto_b200 import train_ops as T
    torch.manual_seed(21)
    M = 777
    x0 = torch.randn(M, 128, device="cuda")
    g0, b0 = torch.randn(128, device="cuda"), torch.randn(128, device="cuda")
    w = torch.randn(128, 128, device="cuda") * 0.1
    go = torch.randn(M, 128, device="cuda")
    def run(fused):
        x, g, b = (t.clone().requires_grad_(True) for t in (x0, g0, b0))
        if fused:
            a, r = T.layernorm128_res(x, g, b)
        else:
            a, r = T.layernorm128(x, g, b), x
        y = T.linear(a, w, None, r)
        y.backward(go)
        return y.detach(), x.grad, g.grad, b.grad
    for u, v in zip(run(True), run(False)):
        assert torch.allclose(u, v, rtol=1e-5, atol=1e-5)
    x = x0.clone().requires_grad_(True)
    a, r = T.layernorm128_res(x, g0, b0)
    (r * go).sum().backward()                      # only the pass-through output is used
    assert torch.equal(x.grad, go)
    x = x0.clone().requires_grad_(True)
    a, r = T.layernorm128_res(x, g0, b0)
    (a * go).sum().backward()                      # only the normalised output is used
    x2 = x0.clone().requires_grad_(True)
    (T.layernorm128(x2, g0, b0) * go).sum().backward()
    assert torch.allclose(x.grad, x2.grad, rtol=1e-6, atol=1e-6)
