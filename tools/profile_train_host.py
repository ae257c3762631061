"""Host-side profile (GPU box) of the denoiser training step: cProfile over 10 steps, top functions by own time.
usage: profile_train_host.py [batch] [fp32|bf16]"""
import cProfile, io, os, pstats, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

B, N, T = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 2048, 200
PREC = sys.argv[2] if len(sys.argv) > 2 else "bf16"
d = bench.build_model(T, "fp32").cuda().train()
d.model.train_precision = PREC
b = {k: v.cuda() for k, v in bench.synthetic_batch(0, B, N).items()}
x0 = (torch.sqrt(b["variance"]) * torch.randn(B, 3, N, device="cuda") + b["anchors"])
opt = torch.optim.Adam(d.parameters(), lr=1e-4)
flags = torch.ones(B, 1, N, device="cuda")


def step():
    t = torch.randint(0, T, (B,), device="cuda")
    opt.zero_grad(set_to_none=True)
    loss = d.training_losses(x0, t, anchors=b["anchors"], variance=b["variance"], ctx=[b["code"], b["params"]],
                             anchor_assignment=b["assign"], valid_id=b["valid"], flags=flags)["mse_loss"]
    loss.backward()
    opt.step()


for _ in range(4):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
print(s.getvalue())
