"""Import-name shim: the reference does `import chamfer` (python/difffacto/metrics/chamfer_dist/__init__.py:10) and calls
`chamfer.forward(xyz1, xyz2)` / `chamfer.backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2)` -- the two functions its
compiled extension exports (chamfer_cuda.cpp:36-39).  With this repo on PYTHONPATH the name resolves to the B200 kernels
(difffacto_b200/csrc/metrics.cu) with the same signatures and return lists."""
from difffacto_b200.metrics.chamfer import chamfer_backward as _bwd
from difffacto_b200.metrics.chamfer import chamfer_forward as _fwd


def forward(xyz1, xyz2):
    """-> [dist1 (B,n), dist2 (B,m), idx1 (B,n) int32, idx2 (B,m) int32]   (chamfer_cuda.cpp:22-25)"""
    return list(_fwd(xyz1, xyz2))


def backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2):
    """-> [grad_xyz1, grad_xyz2]   (chamfer_cuda.cpp:27-34)"""
    return list(_bwd(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2))


__version__ = "2.0.0+b200"
