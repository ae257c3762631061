"""Diagnostic (GPU box): where does dfb200_ddpm_step differ from the golden reference output?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import denoiser_ref as R
from difffacto_b200 import _lib
g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/denoiser_golden.npz"))
for tag, (seed, B, N, av) in {"a": (11, 3, 64, False), "b": (12, 2, 128, True)}.items():
    inp = R.synthetic_inputs(seed, B, N, av)
    i = {k: v.cuda() for k, v in inp.items()}
    eps = torch.from_numpy(g[tag + "_eps"]).cuda()
    sched = torch.from_numpy(R.schedule_table(100)).cuda()
    out, x0 = torch.empty_like(eps), torch.empty_like(eps)
    ti = i["t"].to(torch.int32)
    _lib.check(_lib.load().dfb200_ddpm_step(B, N, 100, _lib.ptr(sched), _lib.ptr(ti), _lib.ptr(i["x"]), _lib.ptr(eps),
                                            _lib.ptr(i["anchors"]), _lib.ptr(i["variance"]), _lib.ptr(i["noise"]), _lib.ptr(out),
                                            _lib.ptr(x0), _lib.stream()))
    s = R.schedule(100)
    tg, x0g = R.ddpm_step(s, i["x"], i["t"], eps, i["anchors"], i["variance"], i["noise"])  # torch ops ON THE GPU
    o, gs, gx = out.cpu().numpy(), g[tag + "_sample"], g[tag + "_pred_xstart"]
    bad = np.argwhere(o != gs)
    print(tag, "t =", inp["t"].tolist(), "mismatch sample:", len(bad), "x0:", int((x0.cpu().numpy() != gx).sum()),
          "torch-gpu vs golden sample:", int((tg.cpu().numpy() != gs).sum()), "x0:", int((x0g.cpu().numpy() != gx).sum()))
    for idx in bad[:6]:
        b, c, p = idx
        print("   ", idx, "ours %.9e golden %.9e  x=%.9e a=%.9e v=%.9e eps=%.9e z=%.9e" % (
            o[b, c, p], gs[b, c, p], inp["x"][b, c, p], inp["anchors"][b, c, p], inp["variance"][b, c, p], g[tag + "_eps"][b, c, p],
            inp["noise"][b, c, p]), "x0 ours %.9e golden %.9e" % (x0.cpu().numpy()[b, c, p], gx[b, c, p]))
