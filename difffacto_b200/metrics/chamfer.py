"""Chamfer distance on B200 behind the reference's API
(python/difffacto/metrics/chamfer_dist/__init__.py:14-97: ChamferFunction, ChamferDistanceL2,
ChamferDistanceL2_split, ChamferDistanceL1, registered in METRICS)."""
import torch

from .. import _lib
from .._lib import check, ptr, require_cuda, stream
from ..utils.registry import METRICS


def chamfer_forward(xyz1, xyz2):
    """-> dist1 (B,n), dist2 (B,m) squared NN distances, idx1, idx2 int32 (reference chamfer.forward)."""
    require_cuda(xyz1, xyz2)
    xyz1 = xyz1.contiguous().float()
    xyz2 = xyz2.contiguous().float()
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    dist1 = torch.empty(B, n, device=dev)
    dist2 = torch.empty(B, m, device=dev)
    idx1 = torch.empty(B, n, dtype=torch.int32, device=dev)
    idx2 = torch.empty(B, m, dtype=torch.int32, device=dev)
    with _lib.on(dev):
        check(_lib.load().dfb200_chamfer_forward(B, n, ptr(xyz1), m, ptr(xyz2), ptr(dist1), ptr(dist2), ptr(idx1), ptr(idx2), stream()))
    return dist1, dist2, idx1, idx2


def chamfer_backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2):
    """-> grad_xyz1 (B,n,3), grad_xyz2 (B,m,3)  (reference chamfer.backward, chamfer_cuda.cpp:27-34 / chamfer.cu:173-229)."""
    require_cuda(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2)
    xyz1 = xyz1.contiguous().float()
    xyz2 = xyz2.contiguous().float()
    g1 = torch.empty_like(xyz1)
    g2 = torch.empty_like(xyz2)
    B, n, _ = xyz1.shape
    with _lib.on(xyz1.device):
        check(_lib.load().dfb200_chamfer_backward(B, n, ptr(xyz1), xyz2.shape[1], ptr(xyz2), ptr(idx1.contiguous()),
                                                  ptr(idx2.contiguous()), ptr(grad_dist1.contiguous().float()),
                                                  ptr(grad_dist2.contiguous().float()), ptr(g1), ptr(g2), stream()))
    return g1, g2


class ChamferFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2):
        xyz1 = xyz1.contiguous().float()
        xyz2 = xyz2.contiguous().float()
        dist1, dist2, idx1, idx2 = chamfer_forward(xyz1, xyz2)
        ctx.save_for_backward(xyz1, xyz2, idx1, idx2)
        return dist1, dist2

    @staticmethod
    def backward(ctx, grad_dist1, grad_dist2):
        xyz1, xyz2, idx1, idx2 = ctx.saved_tensors
        return chamfer_backward(xyz1, xyz2, idx1, idx2, grad_dist1, grad_dist2)


class _ChamferBase(torch.nn.Module):
    def __init__(self, ignore_zeros=False, reduce=True):
        super().__init__()
        self.ignore_zeros = ignore_zeros
        self.reduce = reduce

    def _dists(self, xyz1, xyz2):
        if xyz1.size(0) == 1 and self.ignore_zeros:
            xyz1 = xyz1[torch.sum(xyz1, dim=2).ne(0)].unsqueeze(dim=0)
            xyz2 = xyz2[torch.sum(xyz2, dim=2).ne(0)].unsqueeze(dim=0)
        return ChamferFunction.apply(xyz1, xyz2)


@METRICS.register_module()
class ChamferDistanceL2(_ChamferBase):
    def forward(self, xyz1, xyz2):
        d1, d2 = self._dists(xyz1, xyz2)
        if self.reduce:
            d1, d2 = torch.mean(d1), torch.mean(d2)
        return d1 + d2


@METRICS.register_module()
class ChamferDistanceL2_split(_ChamferBase):
    def forward(self, xyz1, xyz2):
        d1, d2 = self._dists(xyz1, xyz2)
        if self.reduce:
            d1, d2 = torch.mean(d1), torch.mean(d2)
        return d1, d2


@METRICS.register_module()
class ChamferDistanceL1(_ChamferBase):
    def forward(self, xyz1, xyz2):
        d1, d2 = self._dists(xyz1, xyz2)
        d1, d2 = torch.sqrt(d1), torch.sqrt(d2)
        if self.reduce:
            d1, d2 = torch.mean(d1), torch.mean(d2)
        return (d1 + d2) / 2
