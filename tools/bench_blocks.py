"""Secondary blocks of the bench line (bench.py): everything the driver should see measured on its own fresh box besides the
headline -- precision modes, the generator path, op / evaluation kernels against their rooflines and the reference's own
kernels, the training step, and the eager-PyTorch GPU baseline.  Every block is bounded (a few seconds) and self-describing.

`oracle/` is used here only as the CHECKER / BASELINE (eps error vs the oracle port, the eager baseline = the oracle port run on
the GPU, oracle/_ref = the reference's own kernels built for sm_100a, timed beside ours); the product path never sees it.
"""
import importlib.util
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 74.4: 148 SMs x 128 FFMA lanes x 2 flop x max SM clock (no measured figure)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    d = json.load(open(p)) if os.path.exists(p) else {}
    return float(d.get("hbm_gbs", 6550.0)), float(d.get("bf16_tflops_sustained", 1400.0))


def _timer(torch, flush):
    def timeit(fn, iters=10, warm=3, l2_flush=True):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(iters):
            if l2_flush:
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]
    return timeit


def load_ref_kernels():
    """The reference's own CUDA extensions, unmodified, built for sm_100a by oracle/build_ref.py (absent -> {})."""
    out = {}
    for name in ("ref_pointnet2_ext", "ref_chamfer", "ref_emd"):
        p = os.path.join(ROOT, "oracle", "_ref", name + ".so")
        if os.path.exists(p):
            try:
                spec = importlib.util.spec_from_file_location(name, p)
                m = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(m)
                out[name] = m
            except Exception:
                pass
    return out


def part_cloud(torch, B, N, seed=0):
    """Part-Gaussian clouds in unit scale: 4 parts, means ~ N(0, 0.3^2), variances ~ logU(0.01, 0.1)."""
    g = torch.Generator().manual_seed(seed)
    mean = 0.3 * torch.randn(B, 4, 3, generator=g)
    std = torch.empty(B, 4, 3).uniform_(-4.6052, -2.3026, generator=g).exp().sqrt()
    part = torch.randint(0, 4, (B, N), generator=g)
    idx = part[..., None].expand(B, N, 3)
    return (torch.gather(mean, 1, idx) + torch.gather(std, 1, idx) * torch.randn(B, N, 3, generator=g)).contiguous()


def ops_block(torch, flush, B=256):
    """PointNet++ ops at batch 256 on the SA/FP shapes of the encoders: median us (CUDA events, L2 flushed), algorithmic
    bytes and fraction of the measured HBM peak; for the O(n*m) pair-test ops also the fraction of the FP32-ALU peak; the reference's
    own kernel (oracle/_ref, sm_100a build of the unmodified source) timed beside each."""
    from difffacto_b200.pointnet2_ops import pointnet2_utils as pu
    hbm, _ = _peaks()
    E = load_ref_kernels().get("ref_pointnet2_ext")
    timeit = _timer(torch, flush)
    rows = []

    def row(op, shape, ours, ref, nbytes, pair_tests=None):
        r = {"op": op, "shape": shape, "us": round(ours, 1), "ref_kernel_us": None if ref is None else round(ref, 1),
             "alg_bytes": nbytes, "frac_hbm": round(nbytes / ours / 1e3 / hbm, 4)}
        if pair_tests is not None:
            r["frac_fp32_alu"] = round(pair_tests * 8 / ours / 1e6 / FP32_PEAK_TFLOPS, 4)
        rows.append(r)

    xyz = part_cloud(torch, B, 2048).cuda()
    n, m = 2048, 512
    o = timeit(lambda: pu.furthest_point_sample(xyz, m))
    r = timeit(lambda: E.furthest_point_sampling(xyz, m)) if E else None
    row("furthest_point_sampling", f"B={B} n={n} m={m}", o, r, B * (12 * n + 4 * m))
    sel = pu.furthest_point_sample(xyz, m)
    xyz_t = xyz.transpose(1, 2).contiguous()
    new_xyz = pu.gather_operation(xyz_t, sel).transpose(1, 2).contiguous()
    o = timeit(lambda: pu.gather_operation(xyz_t, sel))
    r = timeit(lambda: E.gather_points(xyz_t, sel)) if E else None
    row("gather_points", f"B={B} C=3 n={n} m={m}", o, r, B * 4 * (3 * n + m + 3 * m))
    for rad, ns in ((0.1, 16), (0.2, 64), (0.4, 128)):
        o = timeit(lambda: pu.ball_query(rad, ns, xyz, new_xyz))
        r = timeit(lambda: E.ball_query(new_xyz, xyz, rad, ns)) if E else None
        row("ball_query", f"B={B} n={n} m={m} r={rad} ns={ns}", o, r, B * (12 * n + 12 * m + 4 * m * ns), B * m * n)
    idx = pu.ball_query(0.2, 64, xyz, new_xyz)
    for C in (7, 131):
        feats = torch.randn(B, C, n, device="cuda")
        o = timeit(lambda: pu.grouping_operation(feats, idx))
        r = timeit(lambda: E.group_points(feats, idx)) if E else None
        row("group_points", f"B={B} C={C} n={n} np={m} ns=64", o, r, B * 4 * (C * n + m * 64 + C * m * 64))
        del feats
    o = timeit(lambda: pu.three_nn(xyz, new_xyz))
    r = timeit(lambda: E.three_nn(xyz, new_xyz)) if E else None
    row("three_nn", f"B={B} n={n} m={m}", o, r, B * (12 * n + 12 * m + 24 * n), B * m * n)
    dist, i3 = pu.three_nn(xyz, new_xyz)
    w = torch.softmax(-dist, -1).contiguous()
    f3 = torch.randn(B, 256, m, device="cuda")
    o = timeit(lambda: pu.three_interpolate(f3, i3, w))
    r = timeit(lambda: E.three_interpolate(f3, i3, w)) if E else None
    row("three_interpolate", f"B={B} c=256 m={m} n={n}", o, r, B * 4 * (256 * m + 6 * n + 256 * n))
    return rows


def eval_block(torch, flush):
    """BASELINE configs[4], bounded: Chamfer / auction EMD over 512-8192 points and batch 1-1024 (a diagonal of the sweep that fits in
    a few seconds; tools/bench_eval_sweep.py runs the full grid), each beside the reference's own kernel (oracle/_ref)."""
    from difffacto_b200.metrics import emdFunction
    from difffacto_b200.metrics.chamfer import chamfer_forward
    ref = load_ref_kernels()
    C, Em = ref.get("ref_chamfer"), ref.get("ref_emd")
    timeit = _timer(torch, flush)
    rows = []
    for B, n in ((1024, 512), (256, 1024), (256, 2048), (32, 2048), (16, 4096), (4, 8192), (1, 8192)):
        a, b = torch.rand(B, n, 3, device="cuda"), torch.rand(B, n, 3, device="cuda")
        o = timeit(lambda: chamfer_forward(a, b), iters=5, warm=2)
        r = timeit(lambda: C.forward(a, b), iters=5, warm=2) if C else None
        rows.append({"op": "chamfer_forward", "B": B, "n": n, "us": round(o, 1), "ref_kernel_us": None if r is None else round(r, 1),
                     "Tpair_per_s": round(2 * B * n * n / o / 1e6, 2), "frac_fp32_alu": round(2 * B * n * n * 8 / o / 1e6 / FP32_PEAK_TFLOPS, 4)})
    for B, n, eps, iters in ((32, 2048, 0.002, 10000), (32, 2048, 0.005, 50), (512, 1024, 0.005, 50), (64, 4096, 0.005, 50), (4, 8192, 0.005, 50),
                             (1, 8192, 0.005, 50)):
        a, b = torch.rand(B, n, 3, device="cuda"), torch.rand(B, n, 3, device="cuda")
        o = timeit(lambda: emdFunction.apply(a, b, eps, iters), iters=3, warm=1, l2_flush=False)
        r = None
        if Em:
            z = lambda *s, dt=torch.float32: torch.zeros(*s, device="cuda", dtype=dt)  # noqa: E731

            def run_ref():
                Em.forward(a, b, z(B, n), z(B, n, dt=torch.int32) - 1, z(B, n), z(B, n, dt=torch.int32) - 1, z(B, n, dt=torch.int32),
                           z(B, n), z(B, n), z(B * n, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32),
                           z(512, dt=torch.int32), z(B * n, dt=torch.int32), eps, iters)
            r = timeit(run_ref, iters=3, warm=1, l2_flush=False)
        rows.append({"op": "emd_forward", "B": B, "n": n, "eps": eps, "iters": iters, "us": round(o, 1),
                     "ref_kernel_us": None if r is None else round(r, 1)})
    return rows


def precision_block(torch, build_model, synthetic_batch, B, N, T, flop_per_point_step, modes=("bf16", "tf32", "fp32"), steps_bounded=40):
    """Every precision mode of the denoiser on the BASELINE batch: shapes/s of the fused reverse process (bf16: measured by the headline;
    the others on a bounded sample of `steps_bounded` sampling steps, scaled to T -- per-step cost does not depend on t) and
    max |eps - oracle| of one denoiser forward on that batch against the oracle's fp32 CPU port of the reference."""
    from oracle import denoiser_ref as R
    out = {}
    b = synthetic_batch(7, B, N)
    dev = {k: v.cuda() for k, v in b.items()}
    g = torch.Generator().manual_seed(5)
    x = torch.sqrt(b["variance"]) * torch.randn(B, 3, N, generator=g) + b["anchors"]
    t = torch.randint(0, T, (B,), generator=g)
    ref_model = None
    for mode in modes:
        try:
            torch.manual_seed(0)  # the same random-init weights in every mode
            diff = build_model(T, mode).cuda().eval()
            if ref_model is None:
                torch.set_num_threads(os.cpu_count() or 1)
                sd = {k: v.detach().cpu() for k, v in diff.model.state_dict().items()}
                with torch.no_grad():
                    ref_eps = R.denoiser_forward(sd, x, t, [b["code"], b["params"]], b["anchors"], b["variance"], b["valid"], b["assign"])
                ref_model = True
            with torch.no_grad():
                eps = diff.model(x.cuda(), t.cuda(), [dev["code"], dev["params"]], anchors=dev["anchors"].transpose(1, 2),
                                 anchor_assignment=dev["assign"], variances=dev["variance"].transpose(1, 2), valid_id=dev["valid"])
            err = (eps.cpu() - ref_eps).abs().max().item()
            short = build_model(steps_bounded, mode).cuda().eval()
            short.model.load_state_dict(diff.model.state_dict())

            def run():
                return short.p_sample_loop([B, 3, N], dev["anchors"], ctx=[dev["code"], dev["params"]], variance=dev["variance"],
                                           anchor_assignment=dev["assign"], valid_id=dev["valid"], rng="philox", seed=3)
            run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms_per_net_step = e0.elapsed_time(e1) / reps / steps_bounded
            out[mode] = {"shapes_per_s": round(B / (ms_per_net_step * 1e-3 * T), 2), "us_per_sampling_step": round(ms_per_net_step * 1e3, 1),
                         "algorithmic_TFLOPs": round(B * N * flop_per_point_step / (ms_per_net_step * 1e-3) / 1e12, 1),
                         "max_abs_eps_err_vs_oracle": float(f"{err:.3e}"),
                         "sample": f"{reps} x {steps_bounded} sampling steps of the batch-{B} workload, scaled to T={T}"}
            del diff, short
        except Exception as e:  # a mode this build does not have is reported, not hidden
            out[mode] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    out["eps_rms"] = float(f"{ref_eps.pow(2).mean().sqrt().item():.3e}") if ref_model else None
    return out


def gpu_eager_block(torch, synthetic_batch, B, N, T, steps=20):
    """The reference's path as unfused eager PyTorch fp32 on this GPU (BASELINE.md section 5's planned bar): the oracle's port of
    TransformerNet.forward + p_sample (same op sequence as the reference module), cuBLAS/ATen kernels, TF32 off."""
    from oracle import denoiser_ref as R
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = {k: v.cuda() for k, v in R.synthetic_state_dict(0).items()}
    b = {k: v.cuda() for k, v in synthetic_batch(0, B, N).items()}
    s = R.schedule(T)
    x = torch.sqrt(b["variance"]) * torch.randn(B, 3, N, device="cuda") + b["anchors"]

    def step(k, xx):
        t = torch.full((B,), T - 1 - k, dtype=torch.long, device="cuda")
        return R.p_sample(sd, s, xx, t, [b["code"], b["params"]], b["anchors"], b["variance"], b["assign"], b["valid"],
                          torch.randn(B, 3, N, device="cuda"))[0]
    with torch.no_grad():
        for k in range(3):
            x = step(k, x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(steps):
            x = step(3 + k, x)
        torch.cuda.synchronize()
    per_step = (time.perf_counter() - t0) / steps
    return {"value": round(B / (per_step * T), 3), "unit": "shapes/s", "ms_per_sampling_step": round(per_step * 1e3, 3),
            "what": "oracle PyTorch port of the reference path (unfused eager ops, fp32, TF32 off) on this GPU",
            "sample": f"{steps} denoiser+update steps at batch {B}, scaled to T={T} (wall clock incl. host launch overhead)"}


def generator_block(torch, diff, dev, B, N, iters=3):
    """The reference's own caller on top of the plugin: AnchorDiffAE.decode's loop (models/networks/anchor_gen.py:145-169) over
    `p_sample_loop_progressive` with ret_traj / ret_interval = 10 at the BASELINE size, torch noise as in the reference."""
    def decode():
        final = dict()
        for t, sample in diff.p_sample_loop_progressive([B, 3, N], anchors=dev["anchors"], variance=dev["variance"],
                                                        ctx=[dev["code"], dev["params"]], noise=None, anchor_assignment=dev["assign"],
                                                        valid_id=dev["valid"], device="cuda", progress=False):
            if t == 0:
                final["pred"] = sample["sample"].transpose(2, 1)
            elif t % 10 == 0:
                final[t] = sample["sample"].transpose(2, 1)
        return final
    # two warm-up passes: a pass keeps 101 of its 2 000 per-step buffers alive while the next one runs, so the caching allocator
    # reaches its steady state (no cudaMalloc, which waits for the running 5 ms chunk kernel every time) only on the third pass
    # (tools/diag_generator_order.py: 107 / 51 / 0 device allocations, 358 / 368 / 133 ms)
    final = decode()
    final = decode()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        final = decode()
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / iters
    return {"value": round(B / (ms * 1e-3), 2), "unit": "shapes/s", "ms_per_step": round(ms, 2), "wall_ms_per_step": round(wall / iters * 1e3, 2),
            "kept_keys": len(final), "what": "decode() loop over the generator (chunks of steps served by the persistent fused kernel), "
            "T=1000, torch.randn noise per step as the reference draws it"}


def train_block(torch, dist, build_model, synthetic_batch, world, rank, local, flop_per_point_step, B=16, N=2048, T=200, iters=12, warm=4):
    """BASELINE configs[3], denoiser part: one training step = forward + backward + Adam of AnchoredDiffusion.training_losses at
    16 shapes x 2048 points per GPU (batch 128 on 8 GPUs), bf16 tcgen05 GEMMs, under DistributedDataParallel (NCCL gradient
    all-reduce overlapped with backward by DDP's buckets) when N > 1.  Time = CUDA events, max over ranks."""
    dev = torch.device("cuda", local)
    torch.manual_seed(1234)  # identical initial weights on every rank (DDP broadcasts rank 0's anyway)
    d = build_model(T, "fp32").to(dev).train()
    d.model.train_precision = "bf16"
    model = d
    # DDP's bucketed NCCL all-reduces are captured inside the step's CUDA graph (NCCL >= 2.9.6; needs TORCH_NCCL_ASYNC_ERROR_HANDLING=0
    # before init_process_group -- bench.py sets it -- DDP built on a side stream and 11 eager warm-up steps, per the PyTorch
    # CUDA-graphs notes); DFB200_TRAIN_GRAPH_DDP=0 issues the DDP step eagerly instead
    import os
    graph_ddp = os.environ.get("DFB200_TRAIN_GRAPH_DDP", "1") != "0" and os.environ.get("TORCH_NCCL_ASYNC_ERROR_HANDLING", "") == "0"
    if world > 1:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            model = torch.nn.parallel.DistributedDataParallel(d, device_ids=[local], gradient_as_bucket_view=True)
        torch.cuda.current_stream(dev).wait_stream(side)
    b = {k: v.to(dev) for k, v in synthetic_batch(500 + rank, B, N).items()}
    g = torch.Generator(device=dev).manual_seed(77 + rank)
    x0 = torch.sqrt(b["variance"]) * torch.randn(B, 3, N, device=dev, generator=g) + b["anchors"]
    flags = torch.ones(B, 1, N, device=dev)
    graphed = world == 1 or graph_ddp  # one CUDA graph per step (difffacto_b200/train_graph.py)
    if os.environ.get("DFB200_TORCH_ADAM") == "1":  # A/B: torch's fused multi-tensor Adam (8 + 3 launches)
        opt = torch.optim.Adam(d.parameters(), lr=1e-4, fused=True, capturable=graphed)
    else:
        from difffacto_b200.optim import FusedAdam
        opt = FusedAdam(d.parameters(), lr=1e-4)  # torch.optim.Adam's update, one launch for all tensors
    ts, loss = [], None

    def loss_fn(x0, t, anchors, variance, code, params, assign, valid, flags):
        kw = dict(anchors=anchors, variance=variance, ctx=[code, params], anchor_assignment=assign, valid_id=valid, flags=flags)
        return (model(x0, t, **kw) if world > 1 else d.training_losses(x0, t, **kw))["mse_loss"]

    inputs = dict(x0=x0, t=torch.randint(0, T, (B,), device=dev, generator=g), anchors=b["anchors"], variance=b["variance"], code=b["code"],
                  params=b["params"], assign=b["assign"], valid=b["valid"], flags=flags)
    step = None
    if graphed:
        from difffacto_b200.train_graph import GraphedTrainStep
        step = GraphedTrainStep(loss_fn, list(d.parameters()), opt, inputs, warmup=11 if world > 1 else 3)
    for it in range(warm + iters):
        inputs["t"] = torch.randint(0, T, (B,), device=dev, generator=g)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if step is not None:
            loss = step(**inputs)  # copies the batch into the graph's static buffers, replays fwd + bwd + Adam
        else:
            opt.zero_grad(set_to_none=True)
            loss = loss_fn(**inputs)
            loss.backward()
            opt.step()
        e1.record()
        torch.cuda.synchronize()
        if it >= warm:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = torch.tensor([ts[len(ts) // 2]], device=dev, dtype=torch.float64)
    in_sync = None
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # the gradient all-reduce really ran in every (replayed) step: after warm + iters optimizer steps on DIFFERENT per-rank
        # batches the weights of all ranks are still bit-identical
        flat = torch.cat([p.detach().reshape(-1) for p in d.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        diff = (flat - ref).abs().max().reshape(1)
        dist.all_reduce(diff, op=dist.ReduceOp.MAX)
        in_sync = bool(diff.item() == 0.0)
    ms = ms.item()
    flop = 3 * B * N * flop_per_point_step
    return {"ms_per_step": round(ms, 3), "shapes_per_s": round(B * world / ms * 1e3, 1), "batch_per_gpu": B, "global_batch": B * world,
            "algorithmic_TFLOPs_per_gpu": round(flop / ms / 1e9, 1), "n_gpus": world, "loss": float(loss.detach()),
            "cuda_graph": bool(graphed), "ranks_in_sync": in_sync,
            "what": "denoiser training step (fwd+bwd+fused Adam), bf16 tcgen05 GEMMs, DistributedDataParallel over NCCL when n_gpus > 1; "
                    "the whole step (incl. DDP's bucketed all-reduces) is one CUDA graph when cuda_graph is true; algorithmic FLOP = 3 x forward"}
