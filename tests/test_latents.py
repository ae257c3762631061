"""Encoder side of sampling: PartEncoderForTransformerDecoder.sample_latents.
CPU: the oracle port (oracle/latents_ref.py) against golden outputs of the REAL reference (tests/golden/make_golden.py
latents).  GPU: difffacto_b200's CUDA path against the golden outputs and the port, then end to end into the sampler."""
import os

import numpy as np
import pytest
import torch

from oracle import latents_ref as L

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {"a": (3, 2, 256, [0, 0, 0, 0]), "b": (2, 3, 128, [0, 1, 0, 0])}
ENC_CFG = dict(type='PartEncoderForTransformerDecoder', encoder=dict(type='PointNetV2', zdim=256, point_dim=3, per_part_mlp=True),
               part_aligner=dict(type="PartAlignerTransformer", in_channels=256, out_channels=6, n_class=4, d_head=32, depth=5, n_heads=8,
                                 dropout=0., use_checkpoint=False, use_linear=True, class_cond=True, single_attn=True, add_class_cond=True,
                                 cimle=True, noise_scale=100, cond_noise_type=0),
               n_class=4, kl_weight=0, fit_loss_type=4, fit_loss_weight=1.0, use_flow=True, latent_flow_depth=14, latent_flow_hidden_dim=256,
               include_z=False, include_part_code=True, include_params=True, use_gt_params=False, kl_weight_annealing=False, gen=True,
               prior_var=1.0)


@pytest.fixture(scope="module")
def lg():
    return np.load(os.path.join(HERE, "golden", "latents_golden.npz"))


@pytest.mark.parametrize("tag", sorted(CASES))
def test_port_matches_reference_sample_latents(lg, tag):
    B, K, npts, fixed = CASES[tag]
    sd = L.synthetic_encoder_state_dict(77)
    assert sorted(sd) == list(lg["param_names"])
    ctx, mpp, lpp, seg, vid, extra = L.sample_latents(sd, torch.from_numpy(lg[tag + "_prior"]), torch.from_numpy(lg[tag + "_noise"]),
                                                      torch.from_numpy(lg[tag + "_valid_in"]), torch.tensor(fixed, dtype=torch.float32), npts, K)
    assert np.array_equal(seg.numpy(), lg[tag + "_seg"]) and np.array_equal(vid.numpy(), lg[tag + "_valid"])
    for got, name in ((ctx[0], "ctx0"), (ctx[1], "ctx1"), (mpp, "mean_pp"), (lpp, "logvar_pp"), (extra[1], "mean"), (extra[2], "logvar")):
        ref = lg[f"{tag}_{name}"]
        assert np.abs(got.numpy() - ref).max() < 2e-4 * max(1.0, np.abs(ref).max()), name


def _build():
    import difffacto_b200 as D
    enc = D.build_from_cfg(ENC_CFG, D.ENCODERS)
    res = enc.load_state_dict(L.synthetic_encoder_state_dict(77), strict=False)  # sampling path: part_aligner.* and flow.*
    assert not res.unexpected_keys and all(k.startswith("encoder.") for k in res.missing_keys)
    return enc.cuda().eval()


class _Fixed:
    """torch.randn -> the given tensors, in call order (prior draw, then cIMLE noise)"""
    def __init__(self, tensors):
        self.q = list(tensors)

    def __enter__(self):
        self._r = torch.randn
        torch.randn = lambda *a, **k: self.q.pop(0)
        return self

    def __exit__(self, *a):
        torch.randn = self._r


@pytest.mark.gpu
@pytest.mark.parametrize("tag", sorted(CASES))
def test_cuda_sample_latents_matches_reference(lg, tag):
    B, K, npts, fixed = CASES[tag]
    enc = _build()
    with _Fixed([torch.from_numpy(lg[tag + "_prior"]), torch.from_numpy(lg[tag + "_noise"])]):
        ctx, mpp, lpp, seg, vid, extra = enc.sample_latents(B, npts, "cuda", fixed_id=torch.tensor(fixed, dtype=torch.float32),
                                                            valid_id=torch.from_numpy(lg[tag + "_valid_in"]), K=K)
    assert seg.dtype == torch.int32 and np.array_equal(seg.cpu().numpy(), lg[tag + "_seg"])
    assert np.array_equal(vid.cpu().numpy(), lg[tag + "_valid"])
    for got, name in ((ctx[0], "ctx0"), (ctx[1], "ctx1"), (mpp, "mean_pp"), (lpp, "logvar_pp"), (extra[1], "mean"), (extra[2], "logvar")):
        ref = lg[f"{tag}_{name}"]
        assert got.shape == ref.shape, name
        assert np.abs(got.cpu().numpy() - ref).max() < 5e-4 * max(1.0, np.abs(ref).max()), name


@pytest.mark.gpu
def test_latents_feed_the_sampler_end_to_end():
    """noise -> flows -> part aligner -> ctx / anchors / variances -> fused reverse DDPM: finite clouds of the right shape,
    points of absent parts re-assigned to the first valid part (reference :1105-1106)."""
    import difffacto_b200 as D
    from test_gpu_denoiser import DIFF_CFG
    from oracle import denoiser_ref as R
    enc = _build()
    diff = D.build_from_cfg(DIFF_CFG, D.DIFFUSIONS, num_timesteps=8)
    diff.model.load_state_dict(R.synthetic_state_dict(1234), strict=True)
    diff = diff.cuda().eval()
    torch.manual_seed(0)
    valid = torch.tensor([[1, 1, 1, 1], [1, 0, 1, 1]], dtype=torch.float32)
    ctx, mpp, lpp, seg, vid, _ = enc.sample_latents(2, 256, "cuda", fixed_id=torch.zeros(4), valid_id=valid, K=2)
    assert mpp.shape == (4, 3, 256) and ctx[0].shape == (4, 256, 4) and ctx[1].shape == (4, 6, 4)
    assert set(seg[2].unique().tolist()) == {0, 2, 3}
    var = torch.exp(lpp.clamp(-12, 3))  # synthetic weights: keep the variance in a sane range
    x0 = diff.p_sample_loop([4, 3, 256], mpp, ctx=ctx, variance=var, anchor_assignment=seg, valid_id=vid, rng="philox", seed=3)
    assert x0.shape == (4, 3, 256) and torch.isfinite(x0).all()
