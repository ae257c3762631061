"""Runner for the sampling entry point (`tools/run_net.py --task val`), B200 path.

Counterpart of the slice of the reference Runner that the generation configs exercise
(python/difffacto/runner/runner.py:18-133 build/resume, :351-379 val): build the model pieces from the config
through the registries, restore `diffusion.model.*` weights from a reference checkpoint if one is given, sample
every conditioning batch with the fused reverse process, gather across ranks and save `results.npz` under
work_dir.  Only the sampling path is built: the encoder/stylizer that produces the conditioning is out of scope
(DESIGN.md section 6), so conditioning batches come from the registered dataset."""
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from .config import get_cfg
from .parallel import gather_shapes, rank_seed, shard_range
from .utils.registry import DATASETS, DIFFUSIONS, build_from_cfg


class Runner:
    def __init__(self, device=None, args=None):
        cfg = self.cfg = get_cfg()
        self.args = args
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        self.device = torch.device(device if device is not None else f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}")
        self.seed = int(getattr(args, "seed", 0) or 0)
        m = cfg.model
        self.num_timesteps = m.num_timesteps
        self.ret_traj, self.ret_interval = bool(m.ret_traj), int(m.ret_interval or 10)
        self.diffusion = build_from_cfg(m.diffusion, DIFFUSIONS, num_timesteps=m.num_timesteps)
        if cfg.precision:
            self.diffusion.model.precision = cfg.precision
        self.diffusion = self.diffusion.to(self.device).eval()
        self.val_dataset = build_from_cfg(cfg.dataset.val, DATASETS) if cfg.dataset and cfg.dataset.val else None
        self.work_dir = cfg.work_dir
        if cfg.resume_path and os.path.exists(cfg.resume_path):
            self.load(cfg.resume_path)

    def load(self, path):
        """Key-tolerant restore like the reference (runner.py:492-522): accepts the reference's checkpoint layout
        ({'model': state_dict} with optional 'module.' prefix) or a bare state_dict; only `diffusion.*` keys matter."""
        ckpt = torch.load(path, map_location="cpu")
        sd = ckpt.get("model", ckpt) if isinstance(ckpt, dict) else ckpt
        own = self.diffusion.state_dict()
        picked = {}
        for k, v in sd.items():
            k = k[len("module."):] if k.startswith("module.") else k
            k = k[len("diffusion."):] if k.startswith("diffusion.") else k
            if k in own and tuple(own[k].shape) == tuple(v.shape):
                picked[k] = v
        missing = sorted(set(own) - set(picked))
        self.diffusion.load_state_dict(picked, strict=False)
        print(f"[Runner] restored {len(picked)}/{len(own)} tensors from {path}" + (f"; missing {missing[:4]}..." if missing else ""))

    @torch.no_grad()
    def val(self, rng="philox"):
        assert self.val_dataset is not None, "config has no dataset.val"
        results, t0 = [], time.time()
        for bi in range(len(self.val_dataset)):
            B = self.val_dataset.batch_size
            lo, hi = shard_range(B, self.rank, self.world)
            b = {k: v.to(self.device) for k, v in self.val_dataset.batch(bi, lo, hi).items()}
            N = b["assign"].shape[1]
            out = self.diffusion.p_sample_loop([hi - lo, 3, N], b["anchors"], ctx=[b["code"], b["params"]], variance=b["variance"],
                                               anchor_assignment=b["assign"], valid_id=b["valid"], rng=rng,
                                               seed=rank_seed(self.seed * 7919 + bi, self.rank),
                                               traj_interval=self.ret_interval if self.ret_traj else None)
            x0, traj = out if self.ret_traj else (out, None)
            res = {"pred": gather_shapes(x0.transpose(1, 2).contiguous(), B)}  # (B,N,3), as AnchorDiffAE.decode returns
            if traj is not None:
                for s in range(traj.shape[0]):
                    res[(s + 1) * self.ret_interval] = gather_shapes(traj[s].transpose(1, 2).contiguous(), B)
            results.append({k: v.cpu().numpy() for k, v in res.items()})
        torch.cuda.synchronize(self.device)
        if self.rank == 0:
            os.makedirs(self.work_dir, exist_ok=True)
            path = os.path.join(self.work_dir, "results.npz")
            np.savez_compressed(path, **{f"batch{i}_{k}": v for i, r in enumerate(results) for k, v in r.items()})
            n = sum(r["pred"].shape[0] for r in results)
            print(f"[Runner] sampled {n} shapes x {results[0]['pred'].shape[1]} points, T={self.num_timesteps} in {time.time() - t0:.2f}s -> {path}")
        return results

    def run(self):
        raise NotImplementedError("training is outside the B200 sampling build (DESIGN.md section 6)")
