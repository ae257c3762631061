"""Helpers shared by the -m gpu parity tests (synthetic part-segmented clouds etc.)."""
import numpy as np
import torch


def part_cloud(rng, B, N, dup_frac=0.1, origin_frac=0.02):
    """Part-Gaussian clouds in roughly unit scale, with duplicated points (datasets sample with
    replacement, reference shapenet_seg.py:465) and a few points at |p|^2 <= 1e-3 (FPS skip rule)."""
    mean = 0.3 * rng.standard_normal((B, 4, 3))
    std = np.sqrt(np.exp(rng.uniform(np.log(0.01), np.log(0.1), (B, 4, 3))))
    part = rng.integers(0, 4, (B, N))
    xyz = mean[np.arange(B)[:, None], part] + std[np.arange(B)[:, None], part] * rng.standard_normal((B, N, 3))
    ndup = int(N * dup_frac)
    if ndup:
        for b in range(B):
            src = rng.integers(0, N, ndup)
            dst = rng.integers(0, N, ndup)
            xyz[b, dst] = xyz[b, src]
    nor = int(N * origin_frac)
    if nor:
        for b in range(B):
            xyz[b, rng.integers(0, N, nor)] = 0.01 * rng.standard_normal((nor, 3))
    return xyz.astype(np.float32)


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()
