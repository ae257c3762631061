"""Profiling driver (GPU box, under ncu): ball_query at the batch-256 SA1 shape.  usage: profile_ball_query.py [radius] [nsample]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from gpu_util import cu, part_cloud
from difffacto_b200.pointnet2_ops import pointnet2_utils as pu

r = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 64
B = 256
rng = np.random.default_rng(0)
xyz = cu(part_cloud(rng, B, 2048))
sel = pu.furthest_point_sample(xyz, 512)
new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), sel).transpose(1, 2).contiguous()
for _ in range(3):
    idx = pu.ball_query(r, ns, xyz, new_xyz)
torch.cuda.synchronize()
print("done")
