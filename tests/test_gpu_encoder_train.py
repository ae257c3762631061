"""-m gpu: training-side encoder: PointNetV2 with BatchNorm / fused anchor-weighted max-pool and
the latent flows' forward direction, against (a) torch autograd of the same ops on the GPU for each primitive and (b) golden
outputs + gradients of the REAL reference stage-1 encoder (tests/golden/make_golden.py encoder_train)."""
import os

import numpy as np
import pytest
import torch

from oracle import latents_ref as L

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
ENC_CFG = dict(type='PartEncoderForTransformerDecoder', encoder=dict(type='PointNetV2', zdim=256, point_dim=3, per_part_mlp=True), n_class=4,
               kl_weight=5e-4, fit_loss_type=4, fit_loss_weight=1.0, use_flow=True, latent_flow_depth=14, latent_flow_hidden_dim=256,
               include_z=False, include_part_code=True, include_params=True, use_gt_params=True, kl_weight_annealing=False,
               min_kl_weight=1e-7, kl_weight_annealing_end_epoch=4000, gen=True, prior_var=1.0)


def _rel(a, b):
    return (a - b).abs().max().item() / max(b.abs().max().item(), 1e-12)


def test_batchnorm_relu_maxpool_coupling_match_torch_autograd():
    from difffacto_b200 import train_ops as T
    torch.manual_seed(0)
    # BatchNorm1d (+ReLU), training statistics and running-stat update
    x = torch.randn(1000, 96, device="cuda", requires_grad=True)
    bn, ref = torch.nn.BatchNorm1d(96).cuda().train(), torch.nn.BatchNorm1d(96).cuda().train()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(); ref.weight.copy_(bn.weight); ref.bias.copy_(bn.bias)
    y = T.batchnorm(x, bn, relu=True)
    xr = x.detach().clone().requires_grad_(True)
    yr = torch.relu(ref(xr))
    assert (y - yr).abs().max().item() < 1e-4
    go = torch.randn_like(y)
    y.backward(go); yr.backward(go)
    assert _rel(x.grad, xr.grad) < 1e-4 and _rel(bn.weight.grad, ref.weight.grad) < 1e-4 and _rel(bn.bias.grad, ref.bias.grad) < 1e-4
    assert torch.allclose(bn.running_mean, ref.running_mean, atol=1e-5) and torch.allclose(bn.running_var, ref.running_var, atol=1e-5)
    bn.eval(); ref.eval()
    assert (T.batchnorm(x.detach(), bn, relu=False) - ref(x.detach())).abs().max().item() < 1e-4
    # anchor-weighted max-pool without the (B,C,N,A) intermediate
    h = torch.randn(3, 200, 64, device="cuda", requires_grad=True)
    w = torch.nn.functional.one_hot(torch.randint(0, 4, (3, 200), device="cuda"), 4).float()
    out = T.weighted_maxpool(h, w, 4.0)
    hr = h.detach().clone().requires_grad_(True)
    outr = (hr.transpose(1, 2).unsqueeze(-1) * w.unsqueeze(1) * 4).max(2)[0]
    assert torch.equal(out, outr)
    g = torch.randn_like(out)
    out.backward(g); outr.backward(g)
    assert _rel(h.grad, hr.grad) < 1e-6
    # coupling layer forward with log-determinant
    s_t = torch.randn(37, 256, device="cuda", requires_grad=True)
    x2 = torch.randn(37, 128, device="cuda", requires_grad=True)
    y1, ld = T.coupling_forward(s_t, x2)
    sr, xr2 = s_t.detach().clone().requires_grad_(True), x2.detach().clone().requires_grad_(True)
    sc = torch.sigmoid(sr[:, :128] + 2.)
    y1r, ldr = xr2 * sc + sr[:, 128:], torch.log(sc).sum(1)
    assert (y1 - y1r).abs().max().item() < 1e-5 and (ld - ldr).abs().max().item() < 1e-4
    g1, g2 = torch.randn_like(y1), torch.randn_like(ld)
    (y1 * g1).sum().add((ld * g2).sum()).backward(); (y1r * g1).sum().add((ldr * g2).sum()).backward()
    assert _rel(s_t.grad, sr.grad) < 1e-4 and _rel(x2.grad, xr2.grad) < 1e-5


def _pcds(seed, B, N):  # same construction as tests/golden/make_golden.py::synthetic_pcds
    g = torch.Generator().manual_seed(seed)
    seg = torch.randint(0, 4, (B, N), generator=g)
    seg[1][seg[1] == 3] = 0
    present = torch.stack([(seg == k).any(1) for k in range(4)], 1).float()
    pts = 0.5 * torch.randn(B, N, 3, generator=g)
    attn = torch.nn.functional.one_hot(seg, 4).float()
    return {"input": pts, "ref": pts.clone(), "present": present, "ref_seg_mask": seg, "ref_attn_map": attn,
            "part_shift": 0.3 * torch.randn(B, 3, 4, generator=g), "part_scale": 0.2 + 0.3 * torch.rand(B, 3, 4, generator=g),
            "noise": torch.zeros(B, 32)}


def test_stage1_encoder_forward_backward_match_reference():
    import difffacto_b200 as D
    g = np.load(os.path.join(HERE, "golden", "encoder_train_golden.npz"))
    enc = D.build_from_cfg(ENC_CFG, D.ENCODERS)
    enc.load_state_dict(L.synthetic_encoder_state_dict(31, with_pointnet=True, with_aligner=False), strict=True)
    enc = enc.cuda().train()
    pcds = _pcds(8, 8, 256)
    ctx, mpp, lpp, flag, loss_dict, extra = enc(pcds, "cuda", epoch=10, eps=torch.from_numpy(g["eps"]).cuda())
    for got, name, tol in ((ctx[0], "ctx0", 5e-4), (ctx[1], "ctx1", 1e-5), (mpp, "mean_pp", 1e-6), (lpp, "logvar_pp", 1e-5), (flag, "flag_pp", 0)):
        ref = torch.from_numpy(g[name])
        assert got.shape == ref.shape, name
        assert (got.detach().cpu() - ref).abs().max().item() <= tol * max(1.0, ref.abs().max().item()), name
    assert abs(loss_dict["prior_loss"].item() - float(g["prior_loss"])) < 2e-4 * abs(float(g["prior_loss"]))
    total = loss_dict["prior_loss"] * 1000.0 + (ctx[0] * torch.from_numpy(g["R"]).cuda()).sum()
    assert abs(total.item() - float(g["total"])) < 5e-4 * abs(float(g["total"]))
    total.backward()
    params = dict(enc.named_parameters())
    for key in [k for k in g.files if k.startswith("grad:")]:
        ref = torch.from_numpy(g[key])
        got = params[key[5:]].grad.cpu()[:8]
        # (biases feeding a BatchNorm have a mathematically zero gradient: both sides hold rounding noise there, hence the atol)
        assert (got - ref).abs().max().item() <= 5e-3 * ref.abs().max().item() + 1e-4, (key, _rel(got, ref))
    norms = torch.tensor([float(p.grad.norm()) if p.grad is not None else 0.0 for _, p in sorted(params.items())])
    refn = torch.from_numpy(g["grad_norms"]).float()
    assert norms.shape == refn.shape and ((norms - refn).abs() <= 5e-3 * refn.abs() + 1e-3).all()
    # BatchNorm running statistics were updated as nn.BatchNorm1d does
    assert np.allclose(enc.encoder.bn4.running_mean.cpu().numpy(), g["bn4_running_mean"], atol=1e-5)
    assert np.allclose(enc.encoder.bn4.running_var.cpu().numpy(), g["bn4_running_var"], rtol=1e-4, atol=1e-6)
    assert np.allclose(enc.encoder.mlp_m[1].running_var.cpu().numpy(), g["mlp_m1_running_var"], rtol=1e-4, atol=1e-6)


def test_stage1_joint_training_step_encoder_plus_denoiser():
    """train_chair_stage1's loop body (anchor_gen.py:995-1037): encoder forward -> denoiser epsilon loss + prior loss ->
    backward reaches encoder, flows and denoiser; one Adam step changes all of them."""
    import difffacto_b200 as D
    from test_gpu_denoiser import DIFF_CFG
    torch.manual_seed(0)
    enc = D.build_from_cfg(ENC_CFG, D.ENCODERS).cuda().train()
    diff = D.build_from_cfg(DIFF_CFG, D.DIFFUSIONS, num_timesteps=50).cuda().train()
    pcds = _pcds(4, 4, 256)
    opt = torch.optim.Adam(list(enc.parameters()) + list(diff.parameters()), lr=1e-3)
    before = [p.detach().clone() for p in list(enc.parameters()) + list(diff.parameters())]
    ctx, mpp, lpp, flag, loss_dict, _ = enc(pcds, "cuda", epoch=0)
    var_pp = torch.exp(lpp)
    t = torch.randint(0, 50, (4,), device="cuda")
    x0 = pcds["ref"].cuda().transpose(1, 2).contiguous()
    mse = diff.training_losses(x0, t, anchors=mpp, variance=var_pp, ctx=ctx, anchor_assignment=pcds["ref_seg_mask"].cuda().int(),
                               valid_id=pcds["present"].cuda(), flags=flag)["mse_loss"]
    loss = mse + loss_dict["prior_loss"]
    assert torch.isfinite(loss)
    loss.backward()
    opt.step()
    after = list(enc.parameters()) + list(diff.parameters())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in after)
    assert sum(not torch.equal(a, b) for a, b in zip(after, before)) == len(before)
