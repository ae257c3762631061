"""Evaluation-kernel sweep (GPU box) for BASELINE.json configs[4]: Chamfer and auction EMD over 512-8192 points and batch
1-1024, against the reference's own kernels compiled for sm_100a (oracle/_ref) when present.  One JSON line per case
(CUDA events, median of 5 after 2 warm-ups; inputs U[0,1]^3 as in SURVEY.md section 8d)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from tools.bench_ops import REF, timeit
from difffacto_b200.metrics.chamfer import chamfer_forward
from difffacto_b200.metrics import emdFunction


def main():
    C, Em = REF.get("ref_chamfer"), REF.get("ref_emd")
    for n in (512, 1024, 2048, 4096, 8192):
        for B in (1, 16, 256, 1024):
            if B * n * n > 1024 * 4096 * 4096:
                continue
            a, b = torch.rand(B, n, 3, device="cuda"), torch.rand(B, n, 3, device="cuda")
            o = timeit(lambda: chamfer_forward(a, b), iters=5, warm=2)
            r = timeit(lambda: C.forward(a, b), iters=5, warm=2) if C else None
            print(json.dumps({"op": "chamfer_forward", "B": B, "n": n, "ours_us": round(o, 1), "ref_us": None if r is None else round(r, 1),
                              "speedup": None if r is None else round(r / o, 2), "Tpair_per_s": round(2 * B * n * n / o / 1e6, 2)}), flush=True)
    for (eps, iters) in ((0.005, 50), (0.002, 10000)):
        for n in (1024, 2048, 4096, 8192):
            for B in (1, 8, 64, 512):
                if B * n > 64 * 8192 or (iters == 10000 and (B > 64 or n > 2048)):
                    continue
                a, b = torch.rand(B, n, 3, device="cuda"), torch.rand(B, n, 3, device="cuda")
                o = timeit(lambda: emdFunction.apply(a, b, eps, iters), iters=3, warm=1, flush=False)
                r = None
                if Em:
                    z = lambda *s, dt=torch.float32: torch.zeros(*s, device="cuda", dtype=dt)

                    def ref():
                        Em.forward(a, b, z(B, n), z(B, n, dt=torch.int32) - 1, z(B, n), z(B, n, dt=torch.int32) - 1, z(B, n, dt=torch.int32),
                                   z(B, n), z(B, n), z(B * n, dt=torch.int32), z(512, dt=torch.int32), z(512, dt=torch.int32),
                                   z(512, dt=torch.int32), z(B * n, dt=torch.int32), eps, iters)
                    r = timeit(ref, iters=3, warm=1, flush=False)
                print(json.dumps({"op": "emd_forward", "B": B, "n": n, "eps": eps, "iters": iters, "ours_us": round(o, 1),
                                  "ref_us": None if r is None else round(r, 1), "speedup": None if r is None else round(r / o, 2)}), flush=True)


if __name__ == "__main__":
    main()
