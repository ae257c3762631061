"""Approximate EMD (auction algorithm) on B200 behind the reference's API
(python/difffacto/metrics/emd/emd_module.py:32-87: emdFunction, EMD registered in METRICS).
One persistent kernel per call instead of 7 launches per auction round."""
import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _lib
from .._lib import check, ptr, require_cuda, stream
from ..utils.registry import METRICS


def emd_forward(xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_increments, max_increments, unass_idx, unass_cnt,
                unass_cnt_sum, cnt_tmp, max_idx, eps, iters):
    """The reference extension's `emd.forward` (emd.cpp:14-17): every buffer is caller-allocated, results land in `dist`
    (squared matched distances) and `assignment`.  The kernel initialises its own scratch, so the caller's fill values
    do not matter; `unass_cnt_sum` / `cnt_tmp` (>= batch int32 each) receive per-pair diagnostics (rounds run, round from
    which one CTA finished alone)."""
    require_cuda(xyz1, xyz2, dist, assignment)
    b, n, _ = xyz1.shape
    for t, dt, name in ((xyz1, torch.float32, "xyz1"), (xyz2, torch.float32, "xyz2"), (dist, torch.float32, "dist"),
                        (assignment, torch.int32, "assignment")):
        _lib.require(t, dt, name)
    with _lib.on(xyz1.device):
        check(_lib.load().dfb200_emd_forward(b, n, ptr(xyz1), ptr(xyz2), ptr(dist), ptr(assignment), ptr(price), ptr(assignment_inv),
                                             ptr(bid), ptr(bid_increments), ptr(max_increments), ptr(unass_idx), ptr(unass_cnt),
                                             ptr(unass_cnt_sum), ptr(cnt_tmp), ptr(max_idx), float(eps), int(iters), stream()))
    return 1


def emd_backward(xyz1, xyz2, gradxyz, graddist, idx):
    """The reference extension's `emd.backward` (emd.cpp:19-22, emd_cuda.cu:284-317): gradxyz (B,n,3) <- d(sum graddist*dist)/d xyz1."""
    require_cuda(xyz1, xyz2, gradxyz, graddist, idx)
    B, n, _ = xyz1.shape
    with _lib.on(xyz1.device):
        check(_lib.load().dfb200_emd_backward(B, n, ptr(xyz1), ptr(xyz2), ptr(gradxyz), ptr(graddist), ptr(idx), stream()))
    return 1


class emdFunction(Function):
    @staticmethod
    def forward(ctx, xyz1, xyz2, eps, iters):
        batchsize, n, _ = xyz1.size()
        _, m, _ = xyz2.size()
        assert n == m
        assert xyz1.size()[0] == xyz2.size()[0]
        assert n % 1024 == 0
        assert batchsize <= 512
        require_cuda(xyz1, xyz2)
        xyz1 = xyz1.contiguous().float()
        xyz2 = xyz2.contiguous().float()
        dev = xyz1.device
        i32 = dict(dtype=torch.int32, device=dev)
        dist = torch.empty(batchsize, n, device=dev)
        assignment = torch.empty(batchsize, n, **i32)
        assignment_inv = torch.empty(batchsize, m, **i32)
        price = torch.empty(batchsize, m, device=dev)
        bid = torch.empty(batchsize, n, **i32)
        bid_increments = torch.empty(batchsize, n, device=dev)
        max_increments = torch.empty(batchsize, m, device=dev)
        unass_idx = torch.empty(batchsize * n, **i32)
        max_idx = torch.empty(batchsize * m, **i32)
        unass_cnt = torch.zeros(512, **i32)
        rounds = torch.zeros(512, **i32)      # diagnostics: auction rounds executed per pair ...
        solo_from = torch.zeros(512, **i32)   # ... and the round from which one CTA finished alone
        emd_forward(xyz1, xyz2, dist, assignment, price, assignment_inv, bid, bid_increments, max_increments, unass_idx, unass_cnt,
                    rounds, solo_from, max_idx, eps, iters)
        emdFunction.last_stats = (rounds[:batchsize], solo_from[:batchsize], unass_cnt[:batchsize])
        ctx.save_for_backward(xyz1, xyz2, assignment)
        return dist, assignment

    @staticmethod
    def backward(ctx, graddist, gradidx):
        xyz1, xyz2, assignment = ctx.saved_tensors
        graddist = graddist.contiguous()
        gradxyz1 = torch.empty_like(xyz1)
        gradxyz2 = torch.zeros_like(xyz2)
        emd_backward(xyz1, xyz2, gradxyz1, graddist, assignment)
        return gradxyz1, gradxyz2, None, None


@METRICS.register_module()
class EMD(nn.Module):
    def __init__(self, eps, iters, dist_only=False):
        super().__init__()
        self.eps = eps
        self.iters = iters
        self.dist_only = dist_only

    def forward(self, input1, input2):
        if self.dist_only:
            return torch.sqrt(emdFunction.apply(input1, input2, self.eps, self.iters)[0]).mean(1)
        return emdFunction.apply(input1, input2, self.eps, self.iters)
