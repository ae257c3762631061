"""Training-step benchmark (GPU box): forward + backward of AnchoredDiffusion.training_losses on the train_chair_stage1
denoiser shape (batch 16 per GPU x 2048 points, fp32 primitives of csrc/train_ops.cu), CUDA events, median of 10."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench

B, N, T = int(sys.argv[1]) if len(sys.argv) > 1 else 16, 2048, 200
PREC = sys.argv[2] if len(sys.argv) > 2 else "fp32"   # GEMMs of the training path: fp32 CUDA cores or bf16 tcgen05
d = bench.build_model(T, "fp32").cuda().train()
d.model.train_precision = PREC
b = {k: v.cuda() for k, v in bench.synthetic_batch(0, B, N).items()}
x0 = (torch.sqrt(b["variance"]) * torch.randn(B, 3, N, device="cuda") + b["anchors"])
opt = torch.optim.Adam(d.parameters(), lr=1e-4)
flags = torch.ones(B, 1, N, device="cuda")
ts = []
for it in range(14):
    t = torch.randint(0, T, (B,), device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    opt.zero_grad(set_to_none=True)
    loss = d.training_losses(x0, t, anchors=b["anchors"], variance=b["variance"], ctx=[b["code"], b["params"]],
                             anchor_assignment=b["assign"], valid_id=b["valid"], flags=flags)["mse_loss"]
    loss.backward()
    opt.step()
    e1.record()
    torch.cuda.synchronize()
    if it >= 4:
        ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
flop = 3 * B * N * bench.FLOP_PER_POINT_STEP
print(json.dumps({"op": f"training step (denoiser fwd+bwd+Adam, {PREC} GEMMs)", "batch": B, "points": N, "ms": round(ms, 2),
                  "shapes_per_s": round(B / ms * 1e3, 1), "algorithmic_TFLOPs": round(flop / ms / 1e9, 2), "loss": float(loss.detach())}))
