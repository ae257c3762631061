"""Diagnostic (GPU box): phase timeline (SM cycles) of CTA 0 of the fused bf16 denoiser kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("DFB200_DIAGNOSTICS", "1")  # the whole run goes through the diagnostic build of the library
import torch
import bench
from difffacto_b200 import _lib
from oracle import denoiser_ref as R
TLEN = int(sys.argv[2]) if len(sys.argv) > 2 else 30   # length of the sampling loop the item is taken from
d = bench.build_model(TLEN, "bf16").cuda().eval()
inp = R.synthetic_inputs(5, 32, 2048, False)
i = {k: v.cuda() for k, v in inp.items()}
f = lambda: d.model(i["x"], i["t"], [i["code"], i["params"]], anchors=i["anchors"].transpose(1, 2),
                    anchor_assignment=i["assign"], variances=i["variance"].transpose(1, 2), valid_id=i["valid"])
ITEM = int(sys.argv[1]) if len(sys.argv) > 1 else 0
with torch.no_grad():
    f(); f()
    buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
    _lib.load().dfb200_debug_tc_timeline(_lib.ptr(buf), ITEM)
    if ITEM == 0:
        f()
    else:  # steady state of the persistent sampling loop: CTA 0's ITEM-th work item
        d.p_sample_loop([32, 3, 2048], i["anchors"], ctx=[i["code"], i["params"]], variance=i["variance"],
                        anchor_assignment=i["assign"], valid_id=i["valid"], rng="philox", seed=1)
    torch.cuda.synchronize()
    _lib.load().dfb200_debug_tc_timeline(None, 0)
b = buf.cpu().tolist()
E, M = b[:512], b[512:]
t0 = E[0]
rel = lambda v: v - t0 if v else None
print("epilogue(tile0,row0): start 0, init done", rel(E[1]), " end", rel(E[3 + 5 * 40]), " (final head start", rel(E[2 + 5 * 40]), ")")
if E[221] > E[220] > 0:
    print(f"item wall time {E[221] - E[220]} ns -> SM clock {(E[3 + 5 * 40] - E[0]) / (E[221] - E[220]) * 1e3:.0f} MHz")
print("init: dependency wait done", rel(E[210]), " inputs loaded", rel(E[211]), " proj_in done", rel(E[212]), " pre_norm done", rel(E[1]))
hs = E[2 + 5 * 40]
print("head (rel. to head start): post_norm stats", E[204] - hs, " proj_out", E[205] - hs, " update stored", E[206] - hs, " published", E[3 + 5 * 40] - hs)
for l in range(5):
    o = l * 40
    print(f"layer {l}: start {rel(E[2+o])}  LN2+kv done {rel(E[3+o])}  Q ready {rel(E[4+o])}  attn done {rel(E[5+o])}  "
          f"x ready {rel(E[6+o])}  LN3 done {rel(E[7+o])}")
    ff = [(rel(E[8 + o + 2 * c]), rel(E[9 + o + 2 * c])) for c in range(8)]
    print("   FF chunks (acc ready -> u ready):", " ".join(f"{a}->{b_}" for a, b_ in ff))
    print(f"   MMA: layer start {rel(M[2+o])} Q issued {rel(M[3+o])} O tile ready {rel(M[4+o])} out issued {rel(M[5+o])} "
          f"LN3 tiles ready {rel(M[6+o])} first FF-in issued {rel(M[7+o])}")
    mm = [(rel(M[8 + o + 2 * c]), rel(M[9 + o + 2 * c])) for c in range(8)]
    print("   MMA FF units T0 (begin wait u_ready -> got it):", " ".join(f"{a}->{b_}" for a, b_ in mm))

import statistics
it = [v for v in b[230:378] if v]
wt = b[742:890][:len(it)]
if it:
    print(f"per-CTA busy cycles over the launch (tile-0 thread 0): min {min(it)} median {statistics.median(it):.0f} max {max(it)}; "
          f"dependency-wait cycles: min {min(wt)} median {statistics.median(wt):.0f} max {max(wt)}")
    print("  slowest CTAs:", sorted(range(len(it)), key=lambda i: -it[i])[:8], " busy/wait of CTA 0:", it[0], wt[0])
