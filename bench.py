#!/usr/bin/env python
"""Headline benchmark: shapes/sec of full reverse-DDPM sampling (2048 points x 4 parts, T=1000 steps,
gen_chair denoiser) -- BASELINE.json `configs[1]` (batch 32 per B200; N GPUs = N x 32 shapes, weak scaling,
one NCCL all-gather of the finished points).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|tf32|fp32] [--suite all|headline]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch: x_T -> x_0 for 32 shapes per GPU.
Prints ONE JSON line (rank 0).  `--impl reference` times the reference's CPU implementation of the same
path (oracle/denoiser_ref.py, the PyTorch port pinned to the reference by tests/golden) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "shapes/sec reverse DDPM (2048 pts, 4 parts, 1000 steps)"
UNIT = "shapes/s"
B_PER_GPU, NPTS, T_STEPS = int(os.environ.get("DFB200_BENCH_BATCH", "32")), 2048, 1000  # BASELINE configs[1]: 32 per GPU
FLOP_PER_POINT_STEP = 2308096  # proj_in + 5 x (Q, QK^T, PV, out, GEGLU-in, FF-out) + proj_out
# dram__bytes_read.sum + dram__bytes_write.sum of one persistent denoiser launch (24 sampling steps of this workload) from the
# committed `ncu --set full` capture profiles/ncu_denoiser_tc_r1_v7.csv: 69.73 MB + 1.89 MB, i.e. per sampling step:
NCU_DRAM_BYTES_PER_STEP = (69.734912e6 + 1.889280e6) / 24
WORKLOAD = f"gen_chair full 1000-step reverse sampling, batch={B_PER_GPU} per GPU, 2048 pts x 4 parts"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("bf16_tflops_sustained", 1400.9)), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    return 1400.0, "fallback (B200_PROFILING.md sustained figure)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def build_model(T, precision):
    import torch  # noqa: F401
    import difffacto_b200 as D
    from difffacto_b200.config import Config
    cfg = Config(os.path.join(ROOT, "configs", "gen_chair.py"))
    diff = D.build_from_cfg(cfg.model.diffusion, D.DIFFUSIONS, num_timesteps=T)
    diff.model.precision = precision
    return diff


def synthetic_batch(seed, B, N):
    """Part-segmented synthetic clouds: part codes ~ N(0,1), part means ~ N(0,0.3^2),
    log-variances ~ U(ln 0.01, ln 0.1), equal split of the points over the 4 parts, all parts valid."""
    import torch
    g = torch.Generator().manual_seed(seed)
    code = torch.randn(B, 256, 4, generator=g)
    mean = 0.3 * torch.randn(B, 3, 4, generator=g)
    logvar = torch.empty(B, 3, 4).uniform_(-4.6052, -2.3026, generator=g)
    valid = torch.ones(B, 4)
    assign = torch.arange(4, dtype=torch.int32).repeat_interleave(N // 4)[None].repeat(B, 1).contiguous()
    idx = assign.long()[:, None, :].expand(B, 3, N)
    anchors = torch.gather(mean, 2, idx).contiguous()
    variance = torch.gather(logvar.exp(), 2, idx).contiguous()
    params = torch.cat([mean, logvar.exp()], dim=1).contiguous()
    return dict(code=code, params=params, anchors=anchors, variance=variance, assign=assign, valid=valid)


def config_dict(world):
    """The workload both arms run -- `--impl reference` prints the identical dict (the driver compares them)."""
    return {"workload": WORKLOAD, "timesteps": T_STEPS, "points": NPTS, "parts": 4, "batch_per_gpu": B_PER_GPU,
            "global_batch": B_PER_GPU * world,
            "parallelism": f"GPU arm: batch-sharded x{world} (weak scaling), one all-gather of final points; reference arm: the per-GPU "
                           "batch on the host cores of rank 0",
            "l2": "GPU arm: flushed (256 MB write) between timed iterations; inputs of the e2e loop arrive from pinned host memory"}


def run_reference(args):
    """The reference's CPU path (PyTorch fp32 port in oracle/) on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from oracle import denoiser_ref as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = R.synthetic_state_dict(0)
    B, N, S = B_PER_GPU, NPTS, 10  # bounded sample: S denoiser+update steps of the full batch per timed step, scaled to T
    b = synthetic_batch(0, B, N)
    s = R.schedule(T_STEPS)
    x = torch.sqrt(b["variance"]) * torch.randn(B, 3, N) + b["anchors"]

    def sample_steps():
        xx = x
        for k in range(S):
            t = torch.full((B,), T_STEPS - 1 - k, dtype=torch.long)
            xx, _, _ = R.p_sample(sd, s, xx, t, [b["code"], b["params"]], b["anchors"], b["variance"], b["assign"], b["valid"],
                                  torch.randn(B, 3, N))
        return xx

    with torch.no_grad():
        for _ in range(args.warmup):
            sample_steps()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sample_steps()
    dt = (time.perf_counter() - t0) / args.steps  # seconds per timed step = S sampling steps of the batch
    value = B / (dt / S * T_STEPS)
    sample = (f"each timed step = {S} denoiser+update steps of the batch-{B} workload ({S}/{T_STEPS} of one reverse process); value scaled "
              f"to T={T_STEPS} (per-step cost does not depend on t); ms_per_step is the measured time of the bounded step")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(max(1, args.gpus)),
        "impl_detail": {"precision": "fp32", "rng": "torch.randn per step", "threads": cores},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline(budget_s=15.0):
    import torch
    from oracle import denoiser_ref as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = R.synthetic_state_dict(0)
    B, N = 8, NPTS
    b = synthetic_batch(1, B, N)
    s = R.schedule(T_STEPS)
    x = torch.sqrt(b["variance"]) * torch.randn(B, 3, N) + b["anchors"]
    n, t0 = 0, time.perf_counter()
    with torch.no_grad():
        while True:
            t = torch.full((B,), T_STEPS - 1 - n, dtype=torch.long)
            if n == 1:
                t0 = time.perf_counter()  # first step = warm-up
            x, _, _ = R.p_sample(sd, s, x, t, [b["code"], b["params"]], b["anchors"], b["variance"], b["assign"], b["valid"],
                                 torch.randn(B, 3, N))
            n += 1
            if n >= 3 and (time.perf_counter() - t0 > budget_s or n >= 40):
                break
    per_step = (time.perf_counter() - t0) / (n - 1)
    return {"value": B / (per_step * T_STEPS), "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n - 1} denoiser+update steps at batch {B} (oracle PyTorch port of the reference path), scaled to T={T_STEPS}"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from difffacto_b200 import _lib
    from difffacto_b200.parallel import gather_shapes, rank_seed
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the training block captures DDP's NCCL all-reduces in a CUDA graph: the NCCL watchdog's event polling is not allowed
        # while a stream is capturing (PyTorch CUDA-graphs notes, "DDP + whole-network capture")
        os.environ["TORCH_NCCL_ASYNC_ERROR_HANDLING"] = "0"  # torchrun exports 1
        dist.init_process_group("nccl", device_id=dev)
    B, N, T = B_PER_GPU, NPTS, T_STEPS
    diff = build_model(T, args.precision).to(dev).eval()
    host = [{k: v.pin_memory() for k, v in synthetic_batch(100 + rank, B, N).items()} for _ in range(2)]  # double-buffered staging
    res = {k: v.to(dev) for k, v in host[0].items()}
    out_host = [torch.empty(B, N, 3).pin_memory() for _ in range(2)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def hot(step, d):
        """one pass of the hot path over one batch; returns the (B*world,N,3) points"""
        x0 = diff.p_sample_loop([B, 3, N], d["anchors"], ctx=[d["code"], d["params"]], variance=d["variance"],
                                anchor_assignment=d["assign"], valid_id=d["valid"], rng="philox",
                                seed=rank_seed(1000 + step, rank, world))
        pts = x0.transpose(1, 2).contiguous()
        return gather_shapes(pts, B * world) if world > 1 else pts

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return tt.item()
        return ms

    def timed_resident(n):
        """`value`: inputs resident in HBM; each iteration timed by CUDA events between barriers, L2 flushed in between."""
        evs, launches0 = [], _lib.launch_count()
        for s in range(n):
            flush.fill_(s & 0xFF)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            hot(s, res)
            e1.record()
            barrier()
            evs.append(e0.elapsed_time(e1))
        return max_over_ranks(sum(evs)), _lib.launch_count() - launches0

    def timed_e2e(n):
        """`e2e`: the same call fed from PINNED HOST buffers, result read back to the host, every iteration, all inside ONE timed
        region of n iterations.  Staging is double-buffered on side streams (inputs of iteration s+1 upload while s computes, the
        points of s download while s+1 computes), as a serving loop would run it; the L2 flush stays between iterations."""
        cur = torch.cuda.current_stream(dev)
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        dbuf = [{k: torch.empty_like(v, device=dev) for k, v in host[0].items()} for _ in range(2)]
        in_ready = [None, None]
        comp_done = [None, None]
        out_done = [None, None]
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cur)
        s_in.wait_stream(cur)
        for s in range(n + 1):
            k = s & 1
            if s < n:  # upload inputs of iteration s (its device buffers were last read by iteration s-2)
                with torch.cuda.stream(s_in):
                    if comp_done[k] is not None:
                        s_in.wait_event(comp_done[k])
                    for name, v in host[k].items():
                        dbuf[k][name].copy_(v, non_blocking=True)
                    in_ready[k] = torch.cuda.Event()
                    in_ready[k].record(s_in)
            if s >= 1:  # compute iteration s-1 (its inputs were uploaded during iteration s-2's compute), then download
                j = (s - 1) & 1
                cur.wait_event(in_ready[j])
                flush.fill_(s & 0xFF)
                pts = hot(s - 1, dbuf[j])
                comp_done[j] = torch.cuda.Event()
                comp_done[j].record(cur)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(comp_done[j])
                    if out_done[j] is not None:
                        out_done[j].synchronize()  # host buffer j is free again (a consumer would have taken it by now)
                    out_host[j].copy_(pts[rank * B:(rank + 1) * B] if world > 1 else pts, non_blocking=True)
                    pts.record_stream(s_out)
                    out_done[j] = torch.cuda.Event()
                    out_done[j].record(s_out)
        cur.wait_stream(s_out)
        e1.record(cur)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    for s in range(args.warmup):
        hot(s, res)
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms, launches = timed_resident(args.steps)
    clocks = sampler.stop() if rank == 0 else None
    e2e_steps = max(10, min(args.steps, 20))
    e2e_ms = timed_e2e(e2e_steps)

    extra = {}
    if args.suite == "all":
        from tools import bench_blocks as BB
        t_extra = time.perf_counter()
        try:  # BASELINE configs[3], denoiser part: every rank takes part (DDP) -- at N=1 a single-GPU step
            extra["train"] = BB.train_block(torch, dist, build_model, synthetic_batch, world, rank, local, FLOP_PER_POINT_STEP)
        except Exception as e:
            extra["train"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        if world == 1:
            for name, fn in (
                    ("e2e_generator", lambda: BB.generator_block(torch, diff, res, B, N)),
                    ("precision_modes", lambda: BB.precision_block(torch, build_model, synthetic_batch, B, N, T, FLOP_PER_POINT_STEP)),
                    ("ops", lambda: BB.ops_block(torch, flush)),
                    ("eval", lambda: BB.eval_block(torch, flush)),
                    ("gpu_eager_baseline", lambda: BB.gpu_eager_block(torch, synthetic_batch, B, N, T))):
                try:
                    extra[name] = fn()
                except Exception as e:
                    extra[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
        extra["extra_blocks_wall_s"] = round(time.perf_counter() - t_extra, 1)
    if rank == 0:
        value = B * world * args.steps / (total_ms * 1e-3)
        e2e = B * world * e2e_steps / (e2e_ms * 1e-3)
        peak, peak_src = peaks()
        ms_per_net_step = total_ms / args.steps / T
        flop = B * N * FLOP_PER_POINT_STEP
        achieved = flop / (ms_per_net_step * 1e-3) / 1e12
        h2d = sum(v.numel() * v.element_size() for v in host[0].values())
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "tf32": "tf32"}.get(args.precision, "f32"), "data": "synthetic",
            "config": config_dict(world),
            "impl_detail": {"precision": args.precision, "rng": "in-kernel philox", "loop": "one persistent fused launch per reverse process"},
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES_PER_STEP * 1e-9,
                         "note": f"algorithmic {flop / 1e9:.1f} GFLOP per denoiser step (B*N*{FLOP_PER_POINT_STEP}) / mean time of one "
                                 f"sampling step ({ms_per_net_step * 1e3:.1f} us: context kernels + fused denoiser + update), CUDA events "
                                 f"over the timed region; peak {peak_src}; traffic = measured DRAM GB per sampling step (ncu), algorithmic "
                                 f"{B * N * 64 / 1e9:.4f} GB"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out_host[0].numel() * 4,
                    "iterations": e2e_steps, "staging": "pinned host buffers, double-buffered H2D / D2H on side streams, one timed region"},
            "gpu_launches": launches,
            "clocks": clocks,
            "cpu_baseline": cpu_baseline() if world == 1 else None,
        }
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "fp32"])
    ap.add_argument("--suite", default="all", choices=["all", "headline"],
                    help="all: headline + the secondary blocks (train, e2e_generator, precision_modes, ops, eval, gpu_eager_baseline); "
                         "headline: the timed loop only (use under ncu)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr",
               "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
