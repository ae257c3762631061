"""Diagnostic (GPU box): bf16-mode error vs the fp32 CUDA path and per-call timing of both modes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from oracle import denoiser_ref as R
d = bench.build_model(1000, "bf16").cuda().eval()
d.model.load_state_dict(R.synthetic_state_dict(1234)); d.cuda()
for B in (4, 32):
    inp = R.synthetic_inputs(5, B, 2048, False)
    i = {k: v.cuda() for k, v in inp.items()}
    outs = {}
    for prec in ("fp32", "bf16"):
        d.model.precision = prec
        f = lambda: d.model(i["x"], i["t"], [i["code"], i["params"]], anchors=i["anchors"].transpose(1, 2),
                            anchor_assignment=i["assign"], variances=i["variance"].transpose(1, 2), valid_id=i["valid"])
        with torch.no_grad():
            outs[prec] = f(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): f()
            e1.record(); torch.cuda.synchronize()
        print(f"B={B} {prec}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per forward", flush=True)
    e = (outs["bf16"] - outs["fp32"]).abs()
    print(f"B={B}: |bf16 - fp32| max {e.max().item():.3e} mean {e.mean().item():.3e}; eps rms {outs['fp32'].pow(2).mean().sqrt().item():.3f}", flush=True)
