"""Diagnostic (GPU box): auction rounds executed / solo switch round / time of dfb200_emd_forward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from difffacto_b200.metrics import emdFunction
for (B, n, eps, iters) in ((32, 2048, 0.005, 50), (32, 2048, 0.002, 10000), (4, 8192, 0.002, 10000)):
    a, b = torch.rand(B, n, 3, device="cuda"), torch.rand(B, n, 3, device="cuda")
    emdFunction.apply(a, b, eps, iters)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); emdFunction.apply(a, b, eps, iters); e1.record(); torch.cuda.synchronize()
    r, s, u = (t.cpu().tolist() for t in emdFunction.last_stats)
    print(f"B={B} n={n} eps={eps} iters={iters}: {e0.elapsed_time(e1):.2f} ms; rounds min/max {min(r)}/{max(r)}; solo from round "
          f"{min(s)}..{max(s)}; unassigned at exit max {max(u)}", flush=True)
