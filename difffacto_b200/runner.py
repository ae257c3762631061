"""Runner for the entry point (`tools/run_net.py --task val|train`), B200 path.

Counterpart of the slice of the reference Runner that the generation configs exercise
(python/difffacto/runner/runner.py:18-133 build/resume, :351-379 val): build the model pieces from the config
through the registries, restore `diffusion.model.*` weights from a reference checkpoint if one is given, sample
every conditioning batch with the fused reverse process, gather across ranks and save `results.npz` under
work_dir.  `--task train` (reference :135-230 train loop, :470-489 save) trains the cross-diffusion denoiser on the
differentiable path (difffacto_b200/train_ops.py): optimizer / max_norm / intervals from the config, one process per GPU
with DistributedDataParallel gradient all-reduce over NCCL, checkpoints in the reference's layout.  The encoder /
stylizer that produces the conditioning is out of scope, so conditioning batches - and, for
training, the target clouds - come from the registered dataset."""
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from .config import get_cfg
from .parallel import gather_shapes, rank_seed, shard_range
from .utils.registry import DATASETS, DIFFUSIONS, ENCODERS, build_from_cfg

# part-presence distribution of ShapeNet chairs (reference datasets/dataset_utils.py:170-179)
shapenet_chair_part_distribution = {
    '1110': 0.7209302325581395, '1111': 0.2630199803471995, '1101': 0.009498853586636095, '1001': 0.00032754667540124465,
    '1100': 0.002947920078611202, '0111': 0.0013101867016049786, '0110': 0.0016377333770062235, '1011': 0.00032754667540124465}


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
        self.encoder = None
        if m.encoder and m.encoder.type in ENCODERS:  # generation from the prior (`--task val_gen`): latent flows + part aligner
            self.encoder = build_from_cfg(m.encoder, ENCODERS).to(self.device).eval()
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
        if self.encoder is not None:  # `encoder.part_aligner.*`, `encoder.flow.*` of the same checkpoint
            eown = self.encoder.state_dict()
            epick = {}
            for k, v in sd.items():
                k = k[len("module."):] if k.startswith("module.") else k
                k = k[len("encoder."):] if k.startswith("encoder.") else k
                if k in eown and tuple(eown[k].shape) == tuple(v.shape):
                    epick[k] = v
            self.encoder.load_state_dict(epick, strict=False)
            print(f"[Runner] restored {len(epick)}/{len(eown)} encoder tensors")
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
                                               seed=rank_seed(self.seed * 7919 + bi, self.rank, self.world),
                                               traj_interval=self.ret_interval if self.ret_traj else None)
            x0, traj = out if self.ret_traj else (out, None)
            res = {"pred": gather_shapes(x0.transpose(1, 2).contiguous(), B)}  # (B,N,3), as AnchorDiffAE.decode returns
            if traj is not None:  # the keys AnchorDiffAE.decode keeps (anchor_gen.py:160-167), x_T under T included
                for t_key, slot in self.diffusion.traj_keys(self.ret_interval):
                    res[t_key] = gather_shapes(traj[slot].transpose(1, 2).contiguous(), B)
            results.append({k: v.cpu().numpy() for k, v in res.items()})
        torch.cuda.synchronize(self.device)
        if self.rank == 0:
            os.makedirs(self.work_dir, exist_ok=True)
            path = os.path.join(self.work_dir, "results.npz")
            np.savez_compressed(path, **{f"batch{i}_{k}": v for i, r in enumerate(results) for k, v in r.items()})
            n = sum(r["pred"].shape[0] for r in results)
            print(f"[Runner] sampled {n} shapes x {results[0]['pred'].shape[1]} points, T={self.num_timesteps} in {time.time() - t0:.2f}s -> {path}")
        return results

    @torch.no_grad()
    def generate_samples(self, num_gen, param_sample_num=1, batch_size=None, rng="philox"):
        """`--task val_gen` (reference runner.py:399-438): part-presence masks drawn from the ShapeNet-chair distribution,
        latents from the prior through the encoder's flows and part aligner (`sample_latents`, K = param_sample_num
        parameter draws per latent), then the fused reverse process.  Returns / saves {'pred', 'seg_mask_ref'}."""
        assert self.encoder is not None, "config has no model.encoder this build can construct"
        B = batch_size or (self.val_dataset.batch_size if self.val_dataset is not None else 32)
        keys = list(shapenet_chair_part_distribution)
        valid_ids = torch.tensor([[float(c) for c in k] for k in keys], device=self.device)
        probs = torch.tensor([shapenet_chair_part_distribution[k] for k in keys])
        torch.manual_seed(self.seed + self.rank)
        results = dict(pred=[], seg_mask_ref=[])
        npoints = int(self.cfg.model.npoints or 2048)
        # the num_gen shapes are sharded over the ranks (contiguous, balanced), generated in batches of <= B, gathered once
        lo, hi = shard_range(num_gen, self.rank, self.world)
        from .models.encoders.part_encoders import _exp_shift
        for bi, b0 in enumerate(range(lo, hi, B)):
            nb = min(B, hi - b0)
            valid_id = valid_ids[torch.multinomial(probs, nb, replacement=True).to(self.device)]
            ctx, mean_pp, logvar_pp, seg, vid, _ = self.encoder.sample_latents(nb, npoints, self.device, fixed_id=torch.zeros(4, device=self.device),
                                                                              valid_id=valid_id, K=param_sample_num)
            variance = _exp_shift(logvar_pp)
            x0 = self.diffusion.p_sample_loop(list(mean_pp.shape), mean_pp, ctx=ctx, variance=variance, anchor_assignment=seg, valid_id=vid,
                                              rng=rng, seed=rank_seed(self.seed * 7919 + bi, self.rank, self.world))
            results["pred"].append(x0.transpose(1, 2).contiguous())
            results["seg_mask_ref"].append(seg)
        n_local = (hi - lo) * param_sample_num
        pred = torch.cat(results["pred"], dim=0) if results["pred"] else torch.empty(0, npoints, 3, device=self.device)
        seg = torch.cat(results["seg_mask_ref"], dim=0) if results["seg_mask_ref"] else torch.empty(0, npoints, dtype=torch.int32, device=self.device)
        assert pred.shape[0] == n_local
        if self.world > 1 and param_sample_num == 1:
            pred, seg = gather_shapes(pred, num_gen), gather_shapes(seg.int(), num_gen)
        results = {"pred": pred.cpu().numpy(), "seg_mask_ref": seg.cpu().numpy()}
        if self.rank == 0:
            os.makedirs(os.path.join(self.work_dir, "val"), exist_ok=True)
            path = os.path.join(self.work_dir, "val", "gen_fixed0000.npz")
            np.savez_compressed(path, **results)
            print(f"[Runner] generated {results['pred'].shape[0]} shapes -> {path}")
        return results

    # ---- training of the diffusion denoiser -----------------------------------------------------------
    def save(self, epoch, it, optimizer):
        """Checkpoint in the reference's layout (runner.py:470-489): 'model' holds `diffusion.*` keys, 'decoder' the
        diffusion state_dict, plus meta / optimizer."""
        sd = self.diffusion.state_dict()
        model_sd = {f"diffusion.{k}": v for k, v in sd.items()}
        data = {"meta": {"epoch": epoch, "iter": it, "config": self.cfg.dump()}, "decoder": sd, "optimizer": optimizer.state_dict()}
        if self.encoder is not None:
            esd = self.encoder.state_dict()
            model_sd.update({f"encoder.{k}": v for k, v in esd.items()})
            if getattr(self.encoder, "encoder", None) is not None:
                data["encoder"] = self.encoder.encoder.state_dict()
        data["model"] = model_sd
        path = os.path.join(self.work_dir, "checkpoints", f"ckpt_{epoch}.pth")
        os.makedirs(os.path.dirname(path), exist_ok=True)
        torch.save(data, path)
        return path

    def run(self, max_iters=None):
        """Epsilon-objective training of `diffusion.model` (anchored_diffusion.py:760-852) on the dataset's batches:
        x_0 ~ per-part Gaussians N(mean_p, diag(var_p)) around the conditioning, uniform t."""
        cfg = self.cfg
        ds = build_from_cfg(cfg.dataset.train, DATASETS) if cfg.dataset and cfg.dataset.train else self.val_dataset
        assert ds is not None, "config has no dataset.train"
        self.diffusion.train()
        model = self.diffusion
        if self.world > 1:  # DDP all-reduces the gradients the autograd Functions hand to the parameters
            side = torch.cuda.Stream(self.device)  # built on a side stream so that a later CUDA-graph capture may include it
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                model = torch.nn.parallel.DistributedDataParallel(self.diffusion, device_ids=[self.device.index])
            torch.cuda.current_stream(self.device).wait_stream(side)
        # stage-1 configuration (train_chair_stage1.py): the PointNetV2 encoder + latent-flow prior are trained jointly with the
        # denoiser (AnchorDiffAE.forward, anchor_gen.py:995-1037); otherwise the conditioning comes from the dataset
        joint = self.encoder is not None and getattr(self.encoder, "encoder", None) is not None and self.encoder.part_aligner is None \
            and getattr(self.encoder, "use_gt_params_cfg", False)
        params = list(self.diffusion.parameters()) + (list(self.encoder.parameters()) if joint else [])
        if joint:
            self.encoder.train()
        ocfg = dict(cfg.optimizer.dump()) if cfg.optimizer else dict(type="Adam", lr=2e-3, weight_decay=0.)
        # cfg.cuda_graph = True: the denoiser step (forward + backward + clip + optimizer, DDP all-reduces included) is captured once
        # and replayed (difffacto_b200/train_graph.py); the joint stage-1 step (encoder + flows, data-dependent host logic) stays eager
        use_graph = bool(cfg.cuda_graph) and not joint
        if use_graph and self.world > 1 and os.environ.get("TORCH_NCCL_ASYNC_ERROR_HANDLING") != "0":
            print("[Runner] cuda_graph with DDP needs TORCH_NCCL_ASYNC_ERROR_HANDLING=0 before init_process_group: running the step eagerly")
            use_graph = False
        if use_graph and ocfg.get("type") in ("Adam", "AdamW"):
            ocfg["capturable"] = True
        if (ocfg.get("type") == "Adam" and not ocfg.get("amsgrad") and cfg.fused_adam is not False
                and set(ocfg) <= {"type", "lr", "betas", "eps", "weight_decay", "amsgrad", "capturable"}):
            # same update rule and state_dict layout as torch.optim.Adam, one launch for all tensors (difffacto_b200/optim.py)
            from .optim import FusedAdam
            opt = FusedAdam(params, **{k: v for k, v in ocfg.items() if k not in ("type", "amsgrad", "capturable")})
        else:
            opt = getattr(torch.optim, ocfg.pop("type"))(params, **ocfg)
        graphed = None
        max_epoch = int(cfg.max_epoch or 1)
        max_norm, log_interval, ckpt_interval = cfg.max_norm, int(cfg.log_interval or 50), int(cfg.checkpoint_interval or 500)
        torch.manual_seed(self.seed + self.rank)  # reference: seed + local_rank (runner.py:39)
        it, losses = 0, []
        for epoch in range(max_epoch):
            for bi in range(len(ds)):
                B = ds.batch_size
                lo, hi = shard_range(B, self.rank, self.world)
                t = torch.randint(0, self.num_timesteps, (hi - lo,), device=self.device)
                opt.zero_grad(set_to_none=True)
                if joint:
                    pcds = ds.train_batch(bi + epoch * len(ds), lo, hi)
                    ctx, mean_pp, logvar_pp, flag_pp, loss_dict, _ = self.encoder(pcds, self.device, epoch=epoch)
                    from .models.encoders.part_encoders import _exp_shift
                    x0 = pcds["ref"].to(self.device).transpose(1, 2).contiguous()
                    b = dict(anchors=mean_pp, variance=_exp_shift(logvar_pp), code=ctx[0], params=ctx[1],
                             assign=pcds["ref_seg_mask"].to(self.device).int(), valid=pcds["present"].to(self.device))
                    loss = _LossModule.forward_through(model, self.diffusion, x0, t, b, flag_pp) + loss_dict["prior_loss"] + loss_dict["fit_loss"].sum()
                else:
                    b = {k: v.to(self.device) for k, v in ds.batch(bi + epoch * len(ds), lo, hi).items()}
                    x0 = torch.sqrt(b["variance"]) * torch.randn_like(b["anchors"]) + b["anchors"]
                    flags = torch.ones(hi - lo, 1, x0.shape[2], device=self.device)
                    if use_graph:
                        inputs = dict(x0=x0, t=t, flags=flags, **{k: b[k] for k in ("anchors", "variance", "code", "params", "assign", "valid")})
                        if graphed is None:
                            from .train_graph import GraphedTrainStep
                            graphed = GraphedTrainStep(
                                lambda x0, t, flags, **bb: _LossModule.forward_through(model, self.diffusion, x0, t, bb, flags), params, opt,
                                inputs, max_norm=max_norm, warmup=11 if self.world > 1 else 3, seed=self.seed)
                        loss = graphed(**inputs).clone()
                        it += 1
                        losses.append(loss)
                        if it % log_interval == 0 and self.rank == 0:
                            print(f"[Runner] epoch {epoch} iter {it} loss {torch.stack(losses[-log_interval:]).mean().item():.5f}")
                        if max_iters is not None and it >= max_iters:
                            break
                        continue
                    # DDP hooks fire on the wrapped module's forward: route the loss through it
                    loss = _LossModule.forward_through(model, self.diffusion, x0, t, b, flags)
                loss.backward()
                if joint and self.world > 1:  # the encoder is not DDP-wrapped: average its gradients explicitly
                    for p in self.encoder.parameters():
                        if p.grad is not None:
                            dist.all_reduce(p.grad)
                            p.grad /= self.world
                if max_norm:
                    torch.nn.utils.clip_grad_norm_(params, max_norm)
                opt.step()
                it += 1
                losses.append(loss.detach())
                if it % log_interval == 0 and self.rank == 0:
                    print(f"[Runner] epoch {epoch} iter {it} loss {torch.stack(losses[-log_interval:]).mean().item():.5f}")
                if max_iters is not None and it >= max_iters:
                    break
            if ((epoch + 1) % ckpt_interval == 0 or epoch + 1 == max_epoch or (max_iters is not None and it >= max_iters)) and self.rank == 0:
                self.save(epoch + 1, it, opt)
            if max_iters is not None and it >= max_iters:
                break
        if joint:
            self.encoder.eval()
        self.diffusion.eval()
        return torch.stack(losses).cpu()


class _LossModule:
    """DistributedDataParallel synchronises gradients for the autograd graph built by ITS forward(); AnchoredDiffusion's
    forward is the training loss (as AnchorDiffAE.forward calls diffusion.training_losses, anchor_gen.py:1020-1037)."""

    @staticmethod
    def forward_through(wrapped, diffusion, x0, t, b, flags):
        kw = dict(anchors=b["anchors"], variance=b["variance"], ctx=[b["code"], b["params"]], anchor_assignment=b["assign"],
                  valid_id=b["valid"], flags=flags)
        if wrapped is diffusion:
            return diffusion.training_losses(x0, t, **kw)["mse_loss"]
        return wrapped(x0, t, **kw)["mse_loss"]
