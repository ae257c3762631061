"""DATASETS['SyntheticPartSeg']: synthetic part-segmented conditioning batches standing in for the
reference's ShapeNetSegPart + part stylizer on the sampling path (no dataset ships with either repo).

A batch is what `AnchorDiffAE.forward` hands to `decode` in generation mode after `encoder.sample_latents`
(reference models/networks/anchor_gen.py:1042-1045, encoders/part_encoders.py:1052-1110): per-part style codes
(B,256,4), per-part (mean, variance) (B,6,4), per-point anchors / variances (B,3,N), anchor assignments (B,N)
int32 and the valid-part mask (B,4).  Valid masks are drawn from the reference's chair part distribution
shape (datasets/dataset_utils.py:170-179) approximated by independent part-presence probabilities."""
import math

import torch

from .utils.registry import DATASETS


@DATASETS.register_module()
class SyntheticPartSeg:
    def __init__(self, batch_size=32, npoints=2048, n_parts=4, num_batches=1, seed=0, part_presence=(0.95, 0.98, 0.9, 0.6), **_):
        assert n_parts == 4 and npoints % n_parts == 0
        self.batch_size, self.npoints, self.n_parts, self.num_batches, self.seed = batch_size, npoints, n_parts, num_batches, seed
        self.part_presence = part_presence

    def __len__(self):
        return self.num_batches

    def batch(self, index, lo=0, hi=None):
        """Rows [lo, hi) of batch `index` (every rank regenerates the same global batch and takes its slice)."""
        B, N = self.batch_size, self.npoints
        g = torch.Generator().manual_seed(self.seed * 100003 + index)
        code = torch.randn(B, 256, 4, generator=g)
        mean = 0.3 * torch.randn(B, 3, 4, generator=g)
        logvar = torch.empty(B, 3, 4).uniform_(math.log(0.01), math.log(0.1), generator=g)
        valid = (torch.rand(B, 4, generator=g) < torch.tensor(self.part_presence)).float()
        valid[valid.sum(1) == 0, 0] = 1.0
        first = valid.argmax(1)
        part = (torch.arange(4)[None] * valid + first[:, None] * (1 - valid)).to(torch.int32)  # part_encoders.py:1105-1106
        assign = part.repeat_interleave(N // 4, dim=1).contiguous()
        idx = assign.long()[:, None, :].expand(B, 3, N)
        out = dict(code=code, params=torch.cat([mean, logvar.exp()], 1), anchors=torch.gather(mean, 2, idx).contiguous(),
                   variance=torch.gather(logvar.exp(), 2, idx).contiguous(), assign=assign, valid=valid)
        hi = B if hi is None else hi
        return {k: v[lo:hi].contiguous() for k, v in out.items()}

    def train_batch(self, index, lo=0, hi=None):
        """A training batch in the layout of ShapeNetSegPart items (reference datasets/shapenet_seg.py; keys read by
        PartEncoder.forward, part_encoders.py:1196-1204): clouds drawn from the per-part Gaussians of `batch(index)`."""
        b = self.batch(index, lo, hi)
        B, N = b["assign"].shape
        g = torch.Generator().manual_seed(self.seed * 7907 + index + 17)
        pts = (b["variance"].sqrt() * torch.randn(B, 3, N, generator=g) + b["anchors"]).transpose(1, 2).contiguous()
        seg = b["assign"].long()
        mean, var = b["params"][:, :3], b["params"][:, 3:]
        return {"input": pts, "ref": pts.clone(), "present": b["valid"], "ref_seg_mask": seg,
                "ref_attn_map": torch.nn.functional.one_hot(seg, self.n_parts).float(), "part_shift": mean.contiguous(),
                "part_scale": var.sqrt().contiguous(), "noise": torch.zeros(B, 32)}

    def __iter__(self):
        for i in range(self.num_batches):
            yield self.batch(i)
