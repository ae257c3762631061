"""Diagnostic (GPU box): run the UMMA self-test for every descriptor variant and print max errors."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from difffacto_b200 import _lib


def ref(A, W, bias, Cin):
    Ab, Wb = A.bfloat16().float(), W.bfloat16().float()
    D = Ab @ Wb.t()
    if bias is not None:
        hi = bias.bfloat16().float()
        lo = (bias - hi).bfloat16().float()
        D = D + (hi + lo)[None]
    if Cin is not None:
        D = D + Cin
    return D


def run(variant, N, K, use_bias, use_cin):
    torch.manual_seed(variant * 100 + N + K)
    A = torch.randn(128, K, device="cuda")
    W = torch.randn(N, K, device="cuda")
    bias = torch.randn(N, device="cuda") if use_bias else None
    Cin = torch.randn(128, N, device="cuda") if use_cin else None
    D = torch.full((128, N), float("nan"), device="cuda")
    scratch = torch.zeros(N * K * 2 + 256, dtype=torch.uint8, device="cuda")
    lib = _lib.load_diag()
    rc = lib.dfb200_selftest_umma(variant, N, K, _lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(Cin), _lib.ptr(D),
                                  _lib.ptr(scratch), _lib.stream())
    torch.cuda.synchronize()
    if rc != 0:
        return f"rc={rc} {lib.dfb200_last_error().decode()}"
    err = (D - ref(A, W, bias, Cin)).abs().max().item()
    return f"{err:.3e}"


if __name__ == "__main__":
    for variant in (0, 2, 4, 6, 8, 10, 12):
        for (N, K) in [(128, 16), (128, 128), (64, 128), (32, 32)]:
            try:
                print(f"variant={variant} N={N} K={K}: plain {run(variant, N, K, False, False)}  "
                      f"bias {run(variant, N, K, True, False)}  cin+bias {run(variant, N, K, True, True)}", flush=True)
            except Exception as e:  # a trap poisons the context: stop
                print(f"variant={variant} N={N} K={K}: EXCEPTION {e}", flush=True)
                sys.exit(0)
