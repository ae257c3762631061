"""Build-container only: the oracle's PyTorch port against the LIVE reference implementation imported
from /root/reference (skipped where it does not exist, e.g. on the GPU box).  Complements the committed
golden vectors with fresh seeds / shapes / masks, including the degenerate all-parts-absent row."""
import os
import sys

import numpy as np
import pytest
import torch

if not os.path.isdir("/root/reference/python/difffacto"):
    pytest.skip("/root/reference not present", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import make_golden as MG  # noqa: E402
from oracle import denoiser_ref as R  # noqa: E402


@pytest.fixture(scope="module")
def ref_diffusion():
    DIFFUSIONS, build_from_cfg = MG.import_reference()
    torch.set_num_threads(4)
    diff = build_from_cfg(MG.gen_chair_diffusion_cfg(), DIFFUSIONS, num_timesteps=50).eval()
    sd = R.synthetic_state_dict(seed=77)
    diff.model.load_state_dict(sd, strict=True)
    return diff, sd


@pytest.mark.parametrize("seed,B,N,all_valid", [(1, 2, 96, False), (2, 1, 2048, True), (3, 4, 32, False)])
def test_denoiser_and_step_live(ref_diffusion, seed, B, N, all_valid):
    diff, sd = ref_diffusion
    inp = R.synthetic_inputs(seed, B, N, all_valid)
    inp["t"] = inp["t"] % 50
    if seed == 3:
        inp["valid"][0] = 0.0  # every part absent: attention degrades to uniform 0.25 (attention.py:195-197)
    ctx = [inp["code"], inp["params"]]
    with torch.no_grad():
        ref_eps = diff.model(inp["x"], inp["t"], ctx, anchors=inp["anchors"].transpose(1, 2), anchor_assignment=inp["assign"],
                             variances=inp["variance"].transpose(1, 2), valid_id=inp["valid"])
        eps = R.denoiser_forward(sd, inp["x"], inp["t"], ctx, inp["anchors"], inp["variance"], inp["valid"], inp["assign"])
        assert (eps - ref_eps).abs().max().item() < 5e-6
        with MG.FixedNoise([inp["noise"]]):
            ps = diff.p_sample(inp["x"], inp["t"], inp["anchors"], ctx=ctx, variance=inp["variance"],
                               anchor_assignment=inp["assign"], valid_id=inp["valid"])
    smp, x0 = R.ddpm_step(R.schedule(50), inp["x"], inp["t"], ref_eps, inp["anchors"], inp["variance"], inp["noise"])
    assert torch.equal(smp, ps["sample"]) and torch.equal(x0, ps["pred_xstart"])


def test_schedule_live(ref_diffusion):
    diff, _ = ref_diffusion
    s = R.schedule(50)
    for k in R.SCHED_ROWS + ["betas"]:
        assert np.array_equal(s[k], getattr(diff, k)), k  # float64 tables identical


def test_host_mirror_schedule_live(ref_diffusion):
    import difffacto_b200 as D
    diff, _ = ref_diffusion
    mine = D.build_from_cfg(MG.gen_chair_diffusion_cfg(), D.DIFFUSIONS, num_timesteps=50)
    for k in R.SCHED_ROWS + ["betas", "alphas_cumprod_prev", "posterior_log_variance_clipped"]:
        assert np.array_equal(getattr(mine, k), getattr(diff, k)), k
    assert sorted(mine.model.state_dict()) == sorted(diff.model.state_dict())
