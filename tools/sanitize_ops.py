"""compute-sanitizer driver (GPU box): small launches that reach every branch of the round-1 v3 kernels -- FPS fast path
(several block sizes), ball_query grid path (plain, sorted centres, dense hand-over to the ordered scan, degenerate cloud),
packed Chamfer, EMD with the 16-CTA cluster.  Results are checked against the C oracle so that a silent wrong answer fails too.
  compute-sanitizer --tool memcheck  python tools/sanitize_ops.py
  compute-sanitizer --tool racecheck python tools/sanitize_ops.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from gpu_util import cu, part_cloud
from oracle import pointnet2_oracle as O
from difffacto_b200.pointnet2_ops import pointnet2_utils as pu
from difffacto_b200.metrics.chamfer import chamfer_forward
from difffacto_b200.metrics import emdFunction

rng = np.random.default_rng(0)
for (B, n, m) in ((2, 2048, 64), (1, 4096, 32), (1, 600, 40), (1, 300, 30), (1, 70, 20)):
    xyz = part_cloud(rng, B, n)
    assert np.array_equal(pu.furthest_point_sample(cu(xyz), m).cpu().numpy(), O.furthest_point_sampling(xyz, m))
print("fps ok")
cases = [(3, 2048, 200, 0.1, 16), (3, 2048, 200, 0.2, 32), (2, 2048, 100, 0.8, 64), (2, 1500, 77, 0.15, 8),
         (64, 1024, 2048, 0.2, 8)]  # last: >= 6 passes per CTA -> centres sorted by cell
for (B, n, m, r, ns) in cases:
    xyz = part_cloud(rng, B, n)
    c = np.stack([xyz[b, rng.integers(0, n, m)] for b in range(B)]).copy()
    c[:, 0] = 40.0
    assert np.array_equal(pu.ball_query(r, ns, cu(xyz), cu(c)).cpu().numpy(), O.ball_query(c, xyz, r, ns)), (B, n, m, r, ns)
bad = part_cloud(rng, 1, 2048); bad[0, 3] = np.nan
assert np.array_equal(pu.ball_query(0.2, 8, cu(bad), cu(bad[:, :64].copy())).cpu().numpy(), O.ball_query(bad[:, :64].copy(), bad, 0.2, 8))
print("ball_query ok")
a, b = rng.random((2, 1024, 3)).astype(np.float32), rng.random((2, 777, 3)).astype(np.float32)
d1, d2, i1, i2 = chamfer_forward(cu(a), cu(b))
od = O.chamfer_forward(a, b)
assert np.array_equal(i1.cpu().numpy(), od[2]) and np.array_equal(i2.cpu().numpy(), od[3])
print("chamfer ok")
a, b = rng.random((1, 4096, 3)).astype(np.float32), rng.random((1, 4096, 3)).astype(np.float32)
dist, ass = emdFunction.apply(cu(a), cu(b), 0.005, 8)
assert torch.isfinite(dist).all() and int(ass.min()) >= 0 and int(ass.max()) < 4096
torch.cuda.synchronize()
print("emd (16-CTA cluster) ok")
