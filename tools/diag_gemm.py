import os, sys, torch
sys.path.insert(0, "/root/repo")
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
SHAPES = [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]]  # e.g. 32768x1024x128
for (M, N, K) in SHAPES or [(128, 128, 64), (300, 200, 136), (1000, 128, 512), (257, 1024, 128), (32768, 1024, 128), (32768, 128, 512), (32768, 128, 128), (32768, 128, 1024), (32768, 512, 128)]:
    x, w, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda"), torch.randn(N, device="cuda")
    c0 = torch.randn(M, N, device="cuda")
    y = c0.clone()
    T._sgemm(True, True, M, N, K, x, K, w, K, y, N, bias=b, beta=1, bf16=True)
    ref = c0.double() + x.double() @ w.double().t() + b.double()
    err = (y.double() - ref).abs().max().item()
    y2 = torch.empty(M, N, device="cuda")
    us = tm(lambda: T._sgemm(True, True, M, N, K, x, K, w, K, y2, N, bias=b, bf16=True))
    print(f"M={M} N={N} K={K}: max |err| vs fp64 {err:.3e} (sqrt(K) = {K**0.5:.1f}), {us:.1f} us, {2*M*N*K/us/1e6:.1f} TFLOP/s, {(M*K+N*K+M*N)*4/us/1e3:.0f} GB/s")
