"""Profiling driver (GPU box, under ncu): a short reverse-sampling run of the BASELINE workload
(batch 32 x 2048 points) -- T sampling steps, repeated `reps` times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

T = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
precision = sys.argv[3] if len(sys.argv) > 3 else "bf16"
d = bench.build_model(T, precision).cuda().eval()
b = {k: v.cuda() for k, v in bench.synthetic_batch(0, bench.B_PER_GPU, bench.NPTS).items()}
for r in range(reps):
    x = d.p_sample_loop([bench.B_PER_GPU, 3, bench.NPTS], b["anchors"], ctx=[b["code"], b["params"]], variance=b["variance"],
                        anchor_assignment=b["assign"], valid_id=b["valid"], rng="philox", seed=r)
torch.cuda.synchronize()
print("done", float(x.abs().mean()))
