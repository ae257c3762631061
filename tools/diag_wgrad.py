"""wgrad-shaped GEMM (both operands row-contiguous over the reduced token dimension): time and check the register-staged kernel.
usage: python tools/diag_wgrad.py [MxNxK ...]   e.g. 1024x128x32768 (dW1), 128x512x32768 (dW2)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from difffacto_b200 import train_ops as T
torch.manual_seed(0)
def tm(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
SHAPES = [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]] or [(1024, 128, 32768), (128, 512, 32768), (128, 128, 32768)]
for (M, N, K) in SHAPES:
    dy, x = torch.randn(K, M, device="cuda"), torch.randn(K, N, device="cuda")
    tiles = ((M + 127) // 128) * ((N + 127) // 128)
    split = max(1, min((K + 255) // 256, (148 * 2 + tiles - 1) // tiles))
    dw = torch.zeros(M, N, device="cuda")
    T._sgemm(False, False, M, N, K, dy, M, x, N, dw, N, split_k=split, bf16=True)
    ref = dy.double().t() @ x.double()
    err = (dw.double() - ref).abs().max().item()
    dw2 = torch.zeros(M, N, device="cuda")
    us = tm(lambda: T._sgemm(False, False, M, N, K, dy, M, x, N, dw2, N, split_k=split, bf16=True))
    print(f"M={M} N={N} K={K} split={split}: max |err| vs fp64 {err:.3e} (sqrt(K) = {K**0.5:.0f}), {us:.1f} us, {(M*K+N*K)*4/us/1e3:.0f} GB/s")
