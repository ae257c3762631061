"""Profiling driver (GPU box): one denoiser training step (fwd + bwd + Adam) between cudaProfilerStart/Stop, after
warm-up; also prints host-side launch time vs device time of a step.  usage: profile_train.py [batch] [fp32|bf16]
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file X python tools/profile_train.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

B, N, T = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 2048, 200
PREC = sys.argv[2] if len(sys.argv) > 2 else "bf16"
d = bench.build_model(T, "fp32").cuda().train()
d.model.train_precision = PREC
b = {k: v.cuda() for k, v in bench.synthetic_batch(0, B, N).items()}
x0 = (torch.sqrt(b["variance"]) * torch.randn(B, 3, N, device="cuda") + b["anchors"])
from difffacto_b200.optim import FusedAdam
opt = FusedAdam(d.parameters(), lr=1e-4)  # torch.optim.Adam semantics, one launch (as in the bench's train block)
flags = torch.ones(B, 1, N, device="cuda")


def step():
    t = torch.randint(0, T, (B,), device="cuda")
    opt.zero_grad(set_to_none=True)
    loss = d.training_losses(x0, t, anchors=b["anchors"], variance=b["variance"], ctx=[b["code"], b["params"]],
                             anchor_assignment=b["assign"], valid_id=b["valid"], flags=flags)["mse_loss"]
    loss.backward()
    opt.step()
    return loss


for _ in range(4):
    step()
torch.cuda.synchronize()
host, dev = [], []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(); step(); e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    host.append((t1 - t0) * 1e3); dev.append(e0.elapsed_time(e1))
print(f"host launch time per step {min(host):.2f} ms, device time per step {min(dev):.2f} ms")
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
