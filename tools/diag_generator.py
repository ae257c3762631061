"""Diagnostic (GPU box): decode()-style generator loop at the BASELINE size for several chunk lengths."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
B, N, T = 32, 2048, 1000
d = bench.build_model(T, "bf16").cuda().eval()
dev = {k: v.cuda() for k, v in bench.synthetic_batch(0, B, N).items()}
kw = dict(anchors=dev["anchors"], variance=dev["variance"], ctx=[dev["code"], dev["params"]], anchor_assignment=dev["assign"], valid_id=dev["valid"])
def run(chunk):
    n = 0
    for t, s in d.p_sample_loop_progressive([B, 3, N], device="cuda", chunk=chunk, **kw):
        n += 1
    return n
def fused():
    return d.p_sample_loop([B, 3, N], dev["anchors"], ctx=[dev["code"], dev["params"]], variance=dev["variance"], anchor_assignment=dev["assign"],
                           valid_id=dev["valid"], rng="philox", seed=1)
fused(); torch.cuda.synchronize()
t0 = time.perf_counter(); fused(); torch.cuda.synchronize(); print(f"fused one-call loop: {(time.perf_counter() - t0) * 1e3:.1f} ms")
for chunk in [None] + [int(a) for a in sys.argv[1:]]:
    run(chunk); torch.cuda.synchronize()
    t0 = time.perf_counter(); run(chunk); torch.cuda.synchronize()
    print(f"generator chunk={chunk}: {(time.perf_counter() - t0) * 1e3:.1f} ms wall")
