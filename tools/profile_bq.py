"""Profiling target (GPU box, under ncu): ball_query at the batch-256 SA1 shape, r = 0.2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from gpu_util import cu, part_cloud
from difffacto_b200.pointnet2_ops import pointnet2_utils as pu
B = 256
rng = np.random.default_rng(0)
xyz = cu(part_cloud(rng, B, 2048))
sel = pu.furthest_point_sample(xyz, 512)
new_xyz = pu.gather_operation(xyz.transpose(1, 2).contiguous(), sel).transpose(1, 2).contiguous()
for _ in range(3):
    idx = pu.ball_query(0.2, 64, xyz, new_xyz)
torch.cuda.synchronize()
print("done")
