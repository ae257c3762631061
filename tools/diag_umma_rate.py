"""Diagnostic (GPU box): tcgen05.mma issue rate by operand layout and N."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from difffacto_b200 import _lib
lib = _lib.load()
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for layout in (0, 2):
    for N in (64, 128, 256):
        for iters in (64, 512):
            res = []
            for rep in range(3):
                _lib.check(lib.dfb200_bench_umma(layout, N, iters, 8, _lib.ptr(out), _lib.stream()))
                torch.cuda.synchronize()
                res.append(out.item())
            print(f"{'alternating accumulators' if layout == 0 else 'same accumulator (dependent chain)'} N={N} iters={iters}: cycles {res} -> {min(res) / iters:.1f} cyc/MMA "
                  f"(math floor {N // 2})", flush=True)
