"""Profiling driver (GPU box, under ncu): one launch each of the PointNet++ / metric kernels at the batch-256 SA shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
from gpu_util import cu, part_cloud
from difffacto_b200.pointnet2_ops import pointnet2_utils as pu
from difffacto_b200.metrics.chamfer import chamfer_forward

B = 256
rng = np.random.default_rng(0)
xyz = cu(part_cloud(rng, B, 2048))
sel = pu.furthest_point_sample(xyz, 512)
xyz_t = xyz.transpose(1, 2).contiguous()
new_xyz = pu.gather_operation(xyz_t, sel).transpose(1, 2).contiguous()
idx = pu.ball_query(0.2, 64, xyz, new_xyz)
feats = torch.randn(B, 131, 2048, device="cuda")
g = pu.grouping_operation(feats, idx)
feats2 = torch.randn(B, 320, 512, device="cuda")
idx2 = pu.ball_query(0.4, 64, new_xyz, new_xyz[:, :128].contiguous())
g2 = pu.grouping_operation(feats2, idx2)
d, i3 = pu.three_nn(xyz, new_xyz)
w = torch.softmax(-d, -1).contiguous()
f3 = torch.randn(B, 256, 512, device="cuda")
o = pu.three_interpolate(f3, i3, w)
a, b = torch.rand(B, 2048, 3, device="cuda"), torch.rand(B, 2048, 3, device="cuda")
chamfer_forward(a, b)
torch.cuda.synchronize()
print("done")
