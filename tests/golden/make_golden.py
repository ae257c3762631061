"""Generate tests/golden/denoiser_golden.npz from the REAL reference implementation.

Run in the build container only (needs /root/reference; it does not exist on the GPU box):
    python tests/golden/make_golden.py
The reference's Python hot path is imported with namespace shims (no package __init__ runs, so the
compiled `emd`/`chamfer`/`pointnet2_ops` extensions and tensorboardX/plyfile are not needed), the
`AnchoredDiffusion` + `TransformerNet` of configs/gen_chair.py are built through the reference's
own registry, their weights are overwritten with oracle.denoiser_ref.synthetic_state_dict(seed) and
they are run on oracle.denoiser_ref.synthetic_inputs(seed, ...).  Only outputs are stored: inputs
and weights are regenerated from the seeds by the tests.
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/python/difffacto"


def import_reference():
    def ns(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m

    for name, sub in [("difffacto", ""), ("difffacto.utils", "/utils"), ("difffacto.models", "/models"),
                      ("difffacto.models.diffusions", "/models/diffusions"),
                      ("difffacto.models.diffusions.nets", "/models/diffusions/nets")]:
        ns(name, REF + sub)
    p2 = types.ModuleType("pointnet2_ops")
    p2u = types.ModuleType("pointnet2_ops.pointnet2_utils")
    p2u.gather_operation = lambda f, idx: torch.gather(f, 2, idx.long().unsqueeze(1).expand(-1, f.shape[1], -1))
    p2u.furthest_point_sample = None
    p2.pointnet2_utils = p2u
    sys.modules["pointnet2_ops"] = p2
    sys.modules["pointnet2_ops.pointnet2_utils"] = p2u
    importlib.import_module("difffacto.models.diffusions.nets.attention")
    importlib.import_module("difffacto.models.diffusions.anchored_diffusion")
    from difffacto.utils.registry import DIFFUSIONS, build_from_cfg
    return DIFFUSIONS, build_from_cfg


def gen_chair_diffusion_cfg():
    # the `diffusion=dict(...)` block of /root/reference/configs/gen_chair.py:48-85, read from the file
    sys.path.insert(0, "/root/reference/configs")
    mod = importlib.import_module("gen_chair")
    sys.path.pop(0)
    return dict(mod.model["diffusion"])


class FixedNoise:
    """Replace torch.randn_like / torch.randn inside the reference by a queue of given tensors."""
    def __init__(self, tensors):
        self.q = list(tensors)

    def __enter__(self):
        self._rl, self._r = torch.randn_like, torch.randn
        torch.randn_like = lambda x, *a, **k: self.q.pop(0).to(x)
        torch.randn = lambda *a, **k: self.q.pop(0)
        return self

    def __exit__(self, *a):
        torch.randn_like, torch.randn = self._rl, self._r


def main():
    from oracle import denoiser_ref as R
    DIFFUSIONS, build_from_cfg = import_reference()
    torch.set_num_threads(4)
    out = {}
    T = 100
    diff = build_from_cfg(gen_chair_diffusion_cfg(), DIFFUSIONS, num_timesteps=T).eval()
    sd = R.synthetic_state_dict(seed=1234)
    missing = diff.model.load_state_dict(sd, strict=True)
    out["state_dict_keys"] = np.array(sorted(diff.model.state_dict().keys()))
    out["n_params"] = np.array(sum(p.numel() for p in diff.model.parameters()))
    # schedule tables as float32(np.float64 table)
    out["sched"] = np.stack([getattr(diff, k).astype(np.float32) for k in R.SCHED_ROWS])

    for tag, (seed, B, N, all_valid) in {"a": (11, 3, 64, False), "b": (12, 2, 128, True)}.items():
        inp = R.synthetic_inputs(seed, B, N, all_valid)
        ctx = [inp["code"], inp["params"]]
        with torch.no_grad():
            eps = diff.model(inp["x"], inp["t"], ctx, anchors=inp["anchors"].transpose(1, 2), anchor_assignment=inp["assign"],
                             variances=inp["variance"].transpose(1, 2), valid_id=inp["valid"])
            with FixedNoise([inp["noise"]]):
                ps = diff.p_sample(inp["x"], inp["t"], inp["anchors"], ctx=ctx, variance=inp["variance"],
                                   anchor_assignment=inp["assign"], valid_id=inp["valid"])
            # t == 0 row: no noise
            t0 = torch.zeros_like(inp["t"])
            with FixedNoise([inp["noise"]]):
                ps0 = diff.p_sample(inp["x"], t0, inp["anchors"], ctx=ctx, variance=inp["variance"],
                                    anchor_assignment=inp["assign"], valid_id=inp["valid"])
            xq = diff.q_sample(inp["x"], inp["t"], inp["anchors"], noise=inp["noise"], variance=inp["variance"])
            flags = torch.ones(B, 1, N)
            loss = diff.training_losses(inp["x"], inp["t"], anchors=inp["anchors"], variance=inp["variance"], ctx=ctx,
                                        anchor_assignment=inp["assign"], valid_id=inp["valid"], flags=flags,
                                        noise=inp["noise"])["mse_loss"]
        out[f"{tag}_eps"] = eps.numpy()
        out[f"{tag}_sample"] = ps["sample"].numpy()
        out[f"{tag}_pred_xstart"] = ps["pred_xstart"].numpy()
        out[f"{tag}_sample_t0"] = ps0["sample"].numpy()
        out[f"{tag}_q_sample"] = xq.numpy()
        out[f"{tag}_mse_loss"] = loss.numpy()

    # short full loop: T=6, B=2, N=128, noise supplied (x_T draw first, then one per step)
    Ts = 6
    diff6 = build_from_cfg(gen_chair_diffusion_cfg(), DIFFUSIONS, num_timesteps=Ts).eval()
    diff6.model.load_state_dict(sd, strict=True)
    inp = R.synthetic_inputs(21, 2, 128, False)
    rng = np.random.default_rng(99)
    noises = [torch.from_numpy(rng.standard_normal((2, 3, 128)).astype(np.float32)) for _ in range(Ts + 1)]
    import contextlib, io
    with FixedNoise(list(noises)), contextlib.redirect_stdout(io.StringIO()):
        traj = [(t, {k: v.clone() for k, v in o.items()}) for t, o in diff6.p_sample_loop_progressive(
            [2, 3, 128], anchors=inp["anchors"], ctx=[inp["code"], inp["params"]], variance=inp["variance"],
            anchor_assignment=inp["assign"], valid_id=inp["valid"], device="cpu")]
    out["loop_ts"] = np.array([t for t, _ in traj])
    out["loop_samples"] = np.stack([o["sample"].numpy() for _, o in traj])
    out["loop_noises"] = np.stack([n.numpy() for n in noises])
    out["loop_sched"] = np.stack([getattr(diff6, k).astype(np.float32) for k in R.SCHED_ROWS])
    path = os.path.join(HERE, "denoiser_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items()})

    # ---- sampling variants of the same class: DDIM steps and classifier-free guidance (variants_golden.npz) ----
    var = {}
    inp = R.synthetic_inputs(12, 2, 128, True)
    ctx = [inp["code"], inp["params"]]
    for name, kw in {"ddim_eta1": dict(ddim_sampling=True, ddim_nsteps=10, ddim_eta=1.0, ddim_discretize="uniform"),
                     "ddim_eta05_quad": dict(ddim_sampling=True, ddim_nsteps=8, ddim_eta=0.5, ddim_discretize="quad"),
                     "guidance_w2": dict(guidance=True, classifier_weight=2.0)}.items():
        cfg = gen_chair_diffusion_cfg()
        cfg.update(kw)
        d = build_from_cfg(cfg, DIFFUSIONS, num_timesteps=T).eval()
        d.model.load_state_dict(sd, strict=True)
        var[f"{name}_steps"] = np.array(d.steps)
        with torch.no_grad(), FixedNoise([inp["noise"]]):
            ps = d.p_sample(inp["x"], inp["t"], inp["anchors"], ctx=ctx, variance=inp["variance"],
                            anchor_assignment=inp["assign"], valid_id=inp["valid"])
        var[f"{name}_sample"] = ps["sample"].numpy()
        var[f"{name}_pred_xstart"] = ps["pred_xstart"].numpy()
    # a full DDIM loop (10 of 100 steps), noise supplied
    cfg = gen_chair_diffusion_cfg()
    cfg.update(ddim_sampling=True, ddim_nsteps=10, ddim_eta=1.0, ddim_discretize="uniform")
    d = build_from_cfg(cfg, DIFFUSIONS, num_timesteps=T).eval()
    d.model.load_state_dict(sd, strict=True)
    rng = np.random.default_rng(77)
    noises = [torch.from_numpy(rng.standard_normal((2, 3, 128)).astype(np.float32)) for _ in range(len(d.steps) + 1)]
    with FixedNoise(list(noises)), contextlib.redirect_stdout(io.StringIO()):
        x0 = d.p_sample_loop([2, 3, 128], inp["anchors"], ctx=ctx, variance=inp["variance"], anchor_assignment=inp["assign"],
                             valid_id=inp["valid"], device="cpu")
    var["ddim_loop_noises"] = np.stack([n.numpy() for n in noises])
    var["ddim_loop_x0"] = x0.numpy()
    path = os.path.join(HERE, "variants_golden.npz")
    np.savez_compressed(path, **var)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in var.items()})


def import_reference_eval():
    """python/difffacto/datasets/evaluation_utils.py with its compiled-extension imports stubbed (only the pure-torch
    functions knn / lgan_mmd_cov / lgan_mmd_cov_match are called)."""
    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m
    for name, sub in [("difffacto", ""), ("difffacto.datasets", "/datasets"), ("difffacto.metrics", "/metrics"),
                      ("difffacto.metrics.emd", "/metrics/emd"), ("difffacto.utils", "/utils")]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [REF + sub]
            sys.modules[name] = m
    stub("difffacto.metrics.chamfer_dist", ChamferDistanceL2_split=object)
    stub("difffacto.metrics.emd.emd_module", EMD=object)
    stub("difffacto.utils.misc", fps=None)
    return importlib.import_module("difffacto.datasets.evaluation_utils")


def eval_golden():
    """knn / lgan_mmd_cov / lgan_mmd_cov_match of the reference on seeded distance matrices -> eval_golden.npz"""
    import contextlib, io
    E = import_reference_eval()
    g = torch.Generator().manual_seed(4242)
    out = {}
    for tag, (S, Rn) in {"sq": (12, 12), "rect": (9, 14)}.items():
        Mrs = torch.rand(Rn, S, generator=g) + 0.05
        Mrr = torch.rand(Rn, Rn, generator=g); Mrr = (Mrr + Mrr.t()) / 2
        Mss = torch.rand(S, S, generator=g); Mss = (Mss + Mss.t()) / 2
        if tag == "rect":
            Mrs[3] += 2000.0  # a reference shape farther than the outlier threshold from every sample
        with contextlib.redirect_stdout(io.StringIO()):
            nn = E.knn(Mrr, Mrs, Mss, 1, sqrt=False)
            nn1 = E.knn(Mrr, Mrs, Mss, 1, sqrt=True, one_way=True)
            mc = E.lgan_mmd_cov(Mrs.t())
            mm, midx = E.lgan_mmd_cov_match(Mrs.t())
        out[f"{tag}_Mrs"], out[f"{tag}_Mrr"], out[f"{tag}_Mss"] = Mrs.numpy(), Mrr.numpy(), Mss.numpy()
        for k, v in nn.items():
            out[f"{tag}_knn_{k}"] = np.asarray(float(v))
        for k, v in nn1.items():
            out[f"{tag}_knn1way_{k}"] = np.asarray(float(v))
        for k, v in mc.items():
            out[f"{tag}_mmdcov_{k}"] = np.asarray(float(v))
        for k, v in mm.items():
            out[f"{tag}_match_{k}"] = np.asarray(float(v))
        out[f"{tag}_match_idx"] = midx.numpy()
    path = os.path.join(HERE, "eval_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


def latents_golden():
    """PartEncoderForTransformerDecoder.sample_latents of the REAL reference (configs/gen_chair.py encoder block, weights
    from oracle.latents_ref.synthetic_encoder_state_dict) -> latents_golden.npz (outputs only)."""
    import contextlib, io
    from oracle import latents_ref as L
    import_reference()
    m = types.ModuleType("difffacto.models.encoders")
    m.__path__ = [REF + "/models/encoders"]
    sys.modules["difffacto.models.encoders"] = m
    pe = importlib.import_module("difffacto.models.encoders.part_encoders")
    from difffacto.utils.registry import ENCODERS, build_from_cfg
    sys.path.insert(0, "/root/reference/configs")
    cfg = dict(importlib.import_module("gen_chair").model["encoder"])
    sys.path.pop(0)
    with contextlib.redirect_stdout(io.StringIO()):
        enc = build_from_cfg(cfg, ENCODERS).eval()
    sd = L.synthetic_encoder_state_dict(77)
    res = enc.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all(k.startswith("encoder.") for k in res.missing_keys), res
    out = {"param_names": np.array(sorted(sd))}
    g = torch.Generator().manual_seed(5)
    for tag, (B, K, npts, fixed) in {"a": (3, 2, 256, [0, 0, 0, 0]), "b": (2, 3, 128, [0, 1, 0, 0])}.items():
        prior = torch.randn(B, 256, 4, generator=g)
        noise = torch.randn(B * K, 32, generator=g)
        valid = torch.tensor([[1, 1, 1, 1], [1, 0, 1, 1], [0, 1, 1, 0]][:B], dtype=torch.float32)
        with FixedNoise([prior, noise]), contextlib.redirect_stdout(io.StringIO()), torch.no_grad():
            ctx, mpp, lpp, seg, vid, extra = enc.sample_latents(B, npts, "cpu", fixed_id=torch.tensor(fixed, dtype=torch.float32),
                                                                valid_id=valid.clone(), K=K)
        out.update({f"{tag}_prior": prior.numpy(), f"{tag}_noise": noise.numpy(), f"{tag}_valid_in": valid.numpy(),
                    f"{tag}_ctx0": ctx[0].numpy(), f"{tag}_ctx1": ctx[1].numpy(), f"{tag}_mean_pp": mpp.numpy(),
                    f"{tag}_logvar_pp": lpp.numpy(), f"{tag}_seg": seg.numpy(), f"{tag}_valid": vid.numpy(),
                    f"{tag}_mean": extra[1].numpy(), f"{tag}_logvar": extra[2].numpy()})
    path = os.path.join(HERE, "latents_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


def synthetic_pcds(seed, B, N):
    """A training batch shaped like ShapeNetSegPart's (datasets/shapenet_seg.py): points, part labels, presence, part boxes."""
    g = torch.Generator().manual_seed(seed)
    seg = torch.randint(0, 4, (B, N), generator=g)
    seg[1][seg[1] == 3] = 0                                     # shape 1 has no part 3
    present = torch.stack([(seg == k).any(1) for k in range(4)], 1).float()
    pts = 0.5 * torch.randn(B, N, 3, generator=g)
    attn = torch.nn.functional.one_hot(seg, 4).float()
    return {"input": pts, "ref": pts.clone(), "present": present, "ref_seg_mask": seg, "ref_attn_map": attn,
            "part_shift": 0.3 * torch.randn(B, 3, 4, generator=g), "part_scale": 0.2 + 0.3 * torch.rand(B, 3, 4, generator=g),
            "noise": torch.zeros(B, 32)}


def encoder_train_golden():
    """Training forward + backward of the REAL reference stage-1 encoder (configs/train_chair_stage1.py: PointNetV2 +
    latent flows, ground-truth part parameters) -> encoder_train_golden.npz (outputs and a few gradients)."""
    import contextlib, io
    from oracle import latents_ref as L
    import_reference()
    m = types.ModuleType("difffacto.models.encoders")
    m.__path__ = [REF + "/models/encoders"]
    sys.modules["difffacto.models.encoders"] = m
    importlib.import_module("difffacto.models.encoders.pointnet")
    importlib.import_module("difffacto.models.encoders.part_encoders")
    from difffacto.utils.registry import ENCODERS, build_from_cfg
    sys.path.insert(0, "/root/reference/configs")
    cfg = dict(importlib.import_module("train_chair_stage1").model["encoder"])
    sys.path.pop(0)
    with contextlib.redirect_stdout(io.StringIO()):
        enc = build_from_cfg(cfg, ENCODERS)
    sd = L.synthetic_encoder_state_dict(31, with_pointnet=True, with_aligner=False)
    enc.load_state_dict(sd, strict=True)
    enc.train()
    B, N = 8, 256
    pcds = synthetic_pcds(8, B, N)
    g = torch.Generator().manual_seed(123)
    eps = torch.randn(B, 4, 256, generator=g)
    R_ = torch.randn(B, 256, 4, generator=g)
    _cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self          # the reference writes `torch.ones(1).cuda()` in its loss dict
    try:
        with FixedNoise([eps]):
            ctx, mpp, lpp, flag, loss_dict, extra = enc(pcds, "cpu", epoch=10)
    finally:
        torch.Tensor.cuda = _cuda
    total = loss_dict["prior_loss"] * 1000.0 + (ctx[0] * R_).sum()
    total.backward()
    out = {"eps": eps.numpy(), "R": R_.numpy(), "ctx0": ctx[0].detach().numpy(), "ctx1": ctx[1].detach().numpy(), "mean_pp": mpp.numpy(),
           "logvar_pp": lpp.numpy(), "flag_pp": flag.numpy(), "prior_loss": loss_dict["prior_loss"].detach().numpy(),
           "total": total.detach().numpy(), "bn4_running_mean": enc.encoder.bn4.running_mean.numpy(),
           "bn4_running_var": enc.encoder.bn4.running_var.numpy(), "mlp_m1_running_var": enc.encoder.mlp_m[1].running_var.numpy()}
    params = dict(enc.named_parameters())
    for k in ("encoder.conv1.weight", "encoder.conv4.bias", "encoder.bn2.weight", "encoder.bn4.bias", "encoder.mlp_m.0.weight",
              "encoder.mlp_v.6.weight", "encoder.mlp_m.4.weight", "flow.0.chain.0.net_s_t.0.weight", "flow.2.chain.13.net_s_t.4.bias",
              "flow.3.chain.7.net_s_t.2.weight"):
        out["grad:" + k] = params[k].grad.numpy()[:8]  # first 8 rows keep the fixture small; grad_norms covers every tensor
    out["grad_norms"] = np.array([float(p.grad.norm()) if p.grad is not None else 0.0 for _, p in sorted(params.items())])
    path = os.path.join(HERE, "encoder_train_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays; prior_loss", float(loss_dict["prior_loss"]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "encoder_train":
        encoder_train_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == "latents":
        latents_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == "eval":
        eval_golden()
    else:
        main()
        eval_golden()
