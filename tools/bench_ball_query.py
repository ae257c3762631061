"""GPU box: ball_query timings (batch 256, n = 2048, m = 512) for r = 0.1 / 0.2 / 0.4, and at batch 32; CUDA events, L2 flushed.
    DFB200_BALL_QUERY=tpc|grid|scan python tools/bench_ball_query.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import bench_blocks as BB
from difffacto_b200.pointnet2_ops import pointnet2_utils as PU

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
timeit = BB._timer(torch, flush)
out = []
for B in (256, 32):
    xyz = BB.part_cloud(torch, B, 2048, seed=1).cuda()
    new_xyz = xyz[:, :512].contiguous()
    for r, ns in ((0.1, 16), (0.2, 64), (0.4, 128)):
        us = timeit(lambda: PU.ball_query(r, ns, xyz, new_xyz))
        out.append({"B": B, "r": r, "ns": ns, "us": round(us, 1)})
print(json.dumps({"mode": os.environ.get("DFB200_BALL_QUERY", "default"), "ball_query": out}))
