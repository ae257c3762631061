"""ENCODERS['PointNetV2']: the per-part point-cloud encoder of the training path.

Mirror of the reference class (python/difffacto/models/encoders/pointnet.py:122-214), same constructor keywords and
parameter names (Conv1d / BatchNorm1d modules hold the parameters, so reference checkpoints load unchanged).  The forward
pass runs on this repo's differentiable primitives (difffacto_b200/train_ops.py): the four 1x1 convolutions are GEMMs over
the B*N point rows, BatchNorm1d + ReLU is one fused kernel pair, and the anchor-weighted max-pool is computed WITHOUT the
(B, 512, N, 4) intermediate the reference materialises (2.1 GB at batch 128): one kernel keeps the running maximum and its
arg-max per (shape, channel, anchor); the backward pass scatters the gradient to the winners."""
import torch
import torch.nn as nn

from ... import _lib
from ... import train_ops as T
from ...utils.registry import ENCODERS


def _mlp(num_anchors, zdim):
    g = num_anchors
    return nn.Sequential(nn.Conv1d(512 * g, 256 * g, 1, groups=g), nn.BatchNorm1d(256 * g), nn.ReLU(),
                         nn.Conv1d(256 * g, 128 * g, 1, groups=g), nn.BatchNorm1d(128 * g), nn.ReLU(),
                         nn.Conv1d(128 * g, zdim * g, 1, groups=g))


@ENCODERS.register_module()
class PointNetV2(nn.Module):
    def __init__(self, point_dim=3, zdim=1024, num_anchors=4, reweight_by_anchor=True, use_ln=False, per_part_mlp=False):
        super().__init__()
        if use_ln or not per_part_mlp:
            raise NotImplementedError("difffacto_b200.PointNetV2 implements the configuration of configs/train_*.py "
                                      "(per_part_mlp=True, use_ln=False)")
        self.reweight_by_anchor, self.per_part_mlp, self.zdim, self.num_anchors, self.use_ln = reweight_by_anchor, per_part_mlp, zdim, num_anchors, use_ln
        self.conv1, self.conv2 = nn.Conv1d(point_dim, 128, 1), nn.Conv1d(128, 128, 1)
        self.conv3, self.conv4 = nn.Conv1d(128, 256, 1), nn.Conv1d(256, 512, 1)
        self.bn1, self.bn2, self.bn3, self.bn4 = nn.BatchNorm1d(128), nn.BatchNorm1d(128), nn.BatchNorm1d(256), nn.BatchNorm1d(512)
        self.mlp_m = _mlp(num_anchors, zdim)
        self.mlp_v = _mlp(num_anchors, zdim)

    @staticmethod
    def _conv(x2d, conv):
        return T.linear(x2d, conv.weight.view(conv.out_channels, -1), conv.bias)

    def _grouped(self, x_bgc, conv):
        """grouped 1x1 conv on a length-1 sequence = one Linear per anchor: x (B, G, Cin) -> (B, G*Cout)"""
        G = self.num_anchors
        co = conv.out_channels // G
        w = conv.weight.view(G, co, -1)
        b = conv.bias.view(G, co)
        return torch.cat([T.linear(x_bgc[:, g].contiguous(), w[g], b[g]) for g in range(G)], dim=1)

    def _head(self, x_bgc, mlp):
        B, G = x_bgc.shape[0], self.num_anchors
        h = T.batchnorm(self._grouped(x_bgc, mlp[0]), mlp[1], relu=True)
        h = T.batchnorm(self._grouped(h.view(B, G, -1), mlp[3]), mlp[4], relu=True)
        return self._grouped(h.view(B, G, -1), mlp[6]).view(B, G, -1)

    def forward(self, x, attn_weight):
        """x (B, N, point_dim); attn_weight (B, N, num_anchors) -> part-code mean, log-variance (B, num_anchors, zdim)."""
        _lib.require_cuda(x, attn_weight)
        B, N, _ = x.shape
        with T.gemm_precision(getattr(self, "train_precision", "fp32")):
            h = x.to(torch.float32).reshape(B * N, -1).contiguous()
            h = T.batchnorm(self._conv(h, self.conv1), self.bn1, relu=True)
            h = T.batchnorm(self._conv(h, self.conv2), self.bn2, relu=True)
            h = T.batchnorm(self._conv(h, self.conv3), self.bn3, relu=True)
            h = T.batchnorm(self._conv(h, self.conv4), self.bn4, relu=False)
            scale = float(self.num_anchors) if self.reweight_by_anchor else 1.0
            pooled = T.weighted_maxpool(h.view(B, N, 512), attn_weight, scale)       # (B, 512, A)
            xg = pooled.transpose(1, 2).contiguous()                                 # (B, A, 512): anchor-major, as the reference reshapes
            return self._head(xg, self.mlp_m), self._head(xg, self.mlp_v)
