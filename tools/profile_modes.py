"""Profiling target (GPU box, run under ncu): a few denoiser forwards at the BASELINE batch in the given precision mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B, N = 32, 2048
torch.manual_seed(0)
d = bench.build_model(100, mode).cuda().eval()
dev = {k: v.cuda() for k, v in bench.synthetic_batch(0, B, N).items()}
x = torch.sqrt(dev["variance"]) * torch.randn(B, 3, N, device="cuda") + dev["anchors"]
t = torch.full((B,), 50, device="cuda")
with torch.no_grad():
    for _ in range(3):
        eps = d.model(x, t, [dev["code"], dev["params"]], anchors=dev["anchors"].transpose(1, 2), anchor_assignment=dev["assign"],
                      variances=dev["variance"].transpose(1, 2), valid_id=dev["valid"])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eps = d.model(x, t, [dev["code"], dev["params"]], anchors=dev["anchors"].transpose(1, 2), anchor_assignment=dev["assign"],
                      variances=dev["variance"].transpose(1, 2), valid_id=dev["valid"])
    e1.record(); torch.cuda.synchronize()
print(f"{mode}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us per forward (B={B}, N={N}), eps finite: {bool(torch.isfinite(eps).all())}")
