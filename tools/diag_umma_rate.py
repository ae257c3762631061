"""Diagnostic (GPU box): tcgen05.mma issue rate by operand layout and N."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from difffacto_b200 import _lib
lib = _lib.load_diag()
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for layout in (0, 2, 1):
    for N in (64, 128, 256):
        for iters in (64, 512):
            res = []
            for rep in range(3):
                _lib.check(lib.dfb200_bench_umma(layout, N, iters, 8, _lib.ptr(out), _lib.stream()))
                torch.cuda.synchronize()
                res.append(out.item())
            print(f"{'SWIZZLE_128B ' if layout & 1 else ''}{'alternating accumulators' if not layout & 2 else 'same accumulator (dependent chain)'} N={N} iters={iters}: cycles {res} -> {min(res) / iters:.1f} cyc/MMA "
                  f"(math floor {N // 2})", flush=True)
for mode in (0, 1, 2, 3, 4, 5):
    for N in (64, 128, 256):
        if (mode & 1) and N > 128 and not (mode & 2):
            continue
        res = []
        for rep in range(3):
            _lib.check(lib.dfb200_bench_umma2(mode, N, 512, 8, _lib.ptr(out), _lib.stream()))
            torch.cuda.synchronize()
            res.append(out.item())
        print(f"CTA pair (cta_group::2, M=256) {'SWIZZLE_128B ' if mode & 4 else ''}{'TS' if mode & 1 else 'SS'} {'same acc' if mode & 2 else 'alternating acc'} N={N} iters=512: "
              f"{min(res) / 512:.1f} cyc/MMA (per-SM math floor {N // 2})", flush=True)
