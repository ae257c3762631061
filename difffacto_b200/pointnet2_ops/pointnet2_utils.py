"""PointNet++ ops on B200: same Python API as the reference's pointnet2_ops.pointnet2_utils
(pointnet2_ops_lib/pointnet2_ops/pointnet2_utils.py:34-379), backed by the hand-written sm_100a
kernels of difffacto_b200/csrc/pointnet2.cu through the C ABI (no torch extension, no JIT build).

Argument checks and error behaviour follow the reference's C++ shims (`_ext-src/src/*.cpp`):
tensors must be contiguous, float32 / int32 and on a CUDA device ("CPU not supported").
Kernels are enqueued on the current torch CUDA stream.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from .. import _lib
from .._lib import check, ptr, require, require_cuda, stream


def _f32(t, name):
    require(t, torch.float32, name)


def _i32(t, name):
    require(t, torch.int32, name)


class FurthestPointSampling(Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        """xyz (B,N,3) float -> (B,npoint) int32 indices; reference pointnet2_utils.py:34-62."""
        _f32(xyz, "points")
        require_cuda(xyz)
        B, N, _ = xyz.shape
        out = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
        with _lib.on(xyz.device):
            check(_lib.load().dfb200_furthest_point_sampling(B, N, npoint, ptr(xyz), None, ptr(out), stream()))
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return ()


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        """features (B,C,N), idx (B,npoint) int32 -> (B,C,npoint); reference :68-101."""
        _f32(features, "points"); _i32(idx, "idx")
        require_cuda(features, idx)
        ctx.save_for_backward(idx, features)
        B, C, N = features.shape
        M = idx.shape[1]
        out = torch.empty(B, C, M, dtype=torch.float32, device=features.device)
        with _lib.on(features.device):
            check(_lib.load().dfb200_gather_points(B, C, N, M, ptr(features), ptr(idx), ptr(out), stream()))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, features = ctx.saved_tensors
        B, C, N = features.shape
        grad_out = grad_out.contiguous()
        grad = torch.empty(B, C, N, dtype=torch.float32, device=grad_out.device)
        with _lib.on(grad_out.device):
            check(_lib.load().dfb200_gather_points_grad(B, C, N, idx.shape[1], ptr(grad_out), ptr(idx), ptr(grad), stream()))
        return grad, None


gather_operation = GatherOperation.apply


class ThreeNN(Function):
    @staticmethod
    def forward(ctx, unknown, known):
        """unknown (B,n,3), known (B,m,3) -> (dist (B,n,3) L2 *not squared*, idx (B,n,3)); reference :104-136."""
        _f32(unknown, "unknowns"); _f32(known, "knows")
        require_cuda(unknown, known)
        B, n, _ = unknown.shape
        m = known.shape[1]
        dist2 = torch.empty(B, n, 3, dtype=torch.float32, device=unknown.device)
        idx = torch.empty(B, n, 3, dtype=torch.int32, device=unknown.device)
        with _lib.on(unknown.device):
            check(_lib.load().dfb200_three_nn(B, n, m, ptr(unknown), ptr(known), ptr(dist2), ptr(idx), stream()))
        dist = torch.sqrt(dist2)
        ctx.mark_non_differentiable(dist, idx)
        return dist, idx

    @staticmethod
    def backward(ctx, grad_dist, grad_idx):
        return ()


three_nn = ThreeNN.apply


class ThreeInterpolate(Function):
    @staticmethod
    def forward(ctx, features, idx, weight):
        """features (B,c,m), idx (B,n,3), weight (B,n,3) -> (B,c,n); reference :139-191."""
        _f32(features, "points"); _i32(idx, "idx"); _f32(weight, "weight")
        require_cuda(features, idx, weight)
        ctx.save_for_backward(idx, weight, features)
        B, c, m = features.shape
        n = idx.shape[1]
        out = torch.empty(B, c, n, dtype=torch.float32, device=features.device)
        with _lib.on(features.device):
            check(_lib.load().dfb200_three_interpolate(B, c, m, n, ptr(features), ptr(idx), ptr(weight), ptr(out), stream()))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, weight, features = ctx.saved_tensors
        B, c, m = features.shape
        n = idx.shape[1]
        grad_out = grad_out.contiguous()
        grad = torch.empty(B, c, m, dtype=torch.float32, device=grad_out.device)
        with _lib.on(grad_out.device):
            check(_lib.load().dfb200_three_interpolate_grad(B, c, n, m, ptr(grad_out), ptr(idx), ptr(weight), ptr(grad), stream()))
        return grad, torch.zeros_like(idx), torch.zeros_like(weight)


three_interpolate = ThreeInterpolate.apply


class GroupingOperation(Function):
    @staticmethod
    def forward(ctx, features, idx):
        """features (B,C,N), idx (B,npoint,nsample) -> (B,C,npoint,nsample); reference :194-240."""
        _f32(features, "points"); _i32(idx, "idx")
        require_cuda(features, idx)
        ctx.save_for_backward(idx, features)
        B, C, N = features.shape
        _, npoint, nsample = idx.shape
        out = torch.empty(B, C, npoint, nsample, dtype=torch.float32, device=features.device)
        with _lib.on(features.device):
            check(_lib.load().dfb200_group_points(B, C, N, npoint, nsample, ptr(features), ptr(idx), ptr(out), stream()))
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, features = ctx.saved_tensors
        B, C, N = features.shape
        _, npoint, nsample = idx.shape
        grad_out = grad_out.contiguous()
        grad = torch.empty(B, C, N, dtype=torch.float32, device=grad_out.device)
        with _lib.on(grad_out.device):
            check(_lib.load().dfb200_group_points_grad(B, C, N, npoint, nsample, ptr(grad_out), ptr(idx), ptr(grad), stream()))
        return grad, torch.zeros_like(idx)


grouping_operation = GroupingOperation.apply


class BallQuery(Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        """xyz (B,N,3), new_xyz (B,npoint,3) -> (B,npoint,nsample) int32; reference :243-276."""
        _f32(new_xyz, "new_xyz"); _f32(xyz, "xyz")
        require_cuda(new_xyz, xyz)
        B, N, _ = xyz.shape
        npoint = new_xyz.shape[1]
        out = torch.empty(B, npoint, nsample, dtype=torch.int32, device=xyz.device)
        with _lib.on(xyz.device):
            check(_lib.load().dfb200_query_ball_point(B, N, npoint, float(radius), int(nsample), ptr(new_xyz), ptr(xyz), ptr(out), stream()))
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return ()


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """ball_query -> group xyz (centred) [-> group features -> concat]; reference :279-333."""

    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        grouped_xyz = grouping_operation(xyz.transpose(1, 2).contiguous(), idx)  # (B,3,npoint,nsample)
        grouped_xyz -= new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is None:
            assert self.use_xyz, "Cannot have not features and not use xyz as a feature!"
            return grouped_xyz
        grouped_features = grouping_operation(features, idx)
        return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features


class GroupAll(nn.Module):
    """Group every point into one neighbourhood; reference :336-379."""

    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is None:
            return grouped_xyz
        grouped_features = features.unsqueeze(2)
        return torch.cat([grouped_xyz, grouped_features], dim=1) if self.use_xyz else grouped_features
