"""-m gpu: the tools/run_net.py tasks through the Runner on a small synthetic config: `val` (sampling, results file) and
`train` (denoiser training on the differentiable path, reference checkpoint layout)."""
import os
import textwrap

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def small_cfg(tmp_path):
    from difffacto_b200.config import init_cfg
    p = tmp_path / "small.py"
    p.write_text(textwrap.dedent(f"""
        _base_ = '{ROOT}/configs/train_chair_stage1.py'
        model = dict(num_timesteps=8, npoints=256)
        dataset = dict(train=dict(type="SyntheticPartSeg", batch_size=4, npoints=256, n_parts=4, num_batches=3, seed=1),
                       val=dict(type="SyntheticPartSeg", batch_size=4, npoints=256, n_parts=4, num_batches=1, seed=0))
        max_epoch = 2
        checkpoint_interval = 1
        log_interval = 2
        work_dir = '{tmp_path}/work'
    """))
    init_cfg(str(p))
    return tmp_path


@pytest.fixture()
def small_gen_cfg(tmp_path):
    from difffacto_b200.config import init_cfg
    p = tmp_path / "small_gen.py"
    p.write_text(textwrap.dedent(f"""
        _base_ = '{ROOT}/configs/gen_chair.py'
        model = dict(num_timesteps=8, npoints=256, ret_traj=False)
        dataset = dict(val=dict(type="SyntheticPartSeg", batch_size=4, npoints=256, n_parts=4, num_batches=1, seed=0))
        work_dir = '{tmp_path}/work'
    """))
    init_cfg(str(p))
    return tmp_path


def test_train_task_updates_weights_and_writes_reference_layout_checkpoint(small_cfg):
    import difffacto_b200  # noqa: F401
    import difffacto_b200.datasets  # noqa: F401
    from difffacto_b200.runner import Runner
    r = Runner("cuda:0", None)
    before = {k: v.clone() for k, v in r.diffusion.state_dict().items()}
    assert r.encoder is not None and type(r.encoder.encoder).__name__ == "PointNetV2"   # stage 1: encoder + denoiser jointly
    enc_before = {k: v.clone() for k, v in r.encoder.named_parameters()}
    losses = r.run()
    assert losses.numel() == 6 and torch.isfinite(losses).all()
    after = r.diffusion.state_dict()
    changed = sum(not torch.equal(before[k], after[k]) for k in before)
    assert changed == len(before) == 77
    assert all(not torch.equal(enc_before[k], v) for k, v in r.encoder.named_parameters())
    ck = torch.load(os.path.join(str(small_cfg), "work", "checkpoints", "ckpt_2.pth"), map_location="cpu")
    assert set(ck) >= {"meta", "model", "decoder", "optimizer"} and ck["meta"]["iter"] == 6
    assert all(k.startswith(("diffusion.model.", "encoder.")) for k in ck["model"]) and "encoder" in ck
    # the checkpoint restores through the key-tolerant loader (as a reference checkpoint would)
    r2 = Runner("cuda:0", None)
    r2.load(os.path.join(str(small_cfg), "work", "checkpoints", "ckpt_2.pth"))
    assert all(torch.equal(a.cpu(), b.cpu()) for a, b in zip(r2.diffusion.state_dict().values(), after.values()))


def test_training_reduces_the_loss_on_a_fixed_batch():
    """40 Adam steps on one batch: the epsilon loss keeps falling (end-to-end check of the gradient path)."""
    import difffacto_b200 as D
    from difffacto_b200.config import Config
    from difffacto_b200.datasets import SyntheticPartSeg
    cfg = Config(os.path.join(ROOT, "configs", "gen_chair.py"))
    torch.manual_seed(1)
    d = D.build_from_cfg(cfg.model.diffusion, D.DIFFUSIONS, num_timesteps=50).cuda().eval()
    b = {k: v.cuda() for k, v in SyntheticPartSeg(batch_size=4, npoints=256).batch(0).items()}
    torch.manual_seed(0)
    x0 = torch.sqrt(b["variance"]) * torch.randn_like(b["anchors"]) + b["anchors"]
    noise, t = torch.randn_like(x0), torch.randint(0, 50, (4,), device="cuda")
    opt = torch.optim.Adam(d.parameters(), lr=1e-3)
    hist = []
    for _ in range(40):
        opt.zero_grad()
        loss = d.training_losses(x0, t, anchors=b["anchors"], variance=b["variance"], ctx=[b["code"], b["params"]],
                                 anchor_assignment=b["assign"], valid_id=b["valid"], flags=torch.ones(4, 1, 256, device="cuda"),
                                 noise=noise)["mse_loss"]
        loss.backward()
        opt.step()
        hist.append(loss.item())
    assert hist[-1] < 0.8 * hist[0] and hist[-1] < hist[20] < hist[5], hist[::8]


def test_val_task_writes_results(small_cfg):
    import difffacto_b200  # noqa: F401
    import difffacto_b200.datasets  # noqa: F401
    from difffacto_b200.runner import Runner
    r = Runner("cuda:0", None)
    res = r.val()
    assert res[0]["pred"].shape == (4, 256, 3) and np.isfinite(res[0]["pred"]).all()
    assert os.path.exists(os.path.join(str(small_cfg), "work", "results.npz"))


def test_val_gen_task_generates_from_the_prior(small_gen_cfg):
    small_cfg = small_gen_cfg
    """--task val_gen: prior -> flows -> part aligner -> fused sampler (random-init weights: finite output, right shapes)."""
    import difffacto_b200  # noqa: F401
    import difffacto_b200.datasets  # noqa: F401
    from difffacto_b200.runner import Runner
    r = Runner("cuda:0", None)
    assert r.encoder is not None
    lin = r.encoder.part_aligner.proj_out  # random-init weights: make the predicted Gaussians sane (mean 0, log-variance -3)
    torch.nn.init.zeros_(lin.weight)
    torch.nn.init.constant_(lin.bias, -3.0)
    with torch.no_grad():
        lin.bias[:3] = 0.0
    res = r.generate_samples(8, param_sample_num=2, batch_size=4)
    assert res["pred"].shape == (16, 256, 3) and res["seg_mask_ref"].shape == (16, 256)
    assert np.isfinite(res["pred"]).all()
    assert os.path.exists(os.path.join(str(small_cfg), "work", "val", "gen_fixed0000.npz"))


def test_train_task_with_cuda_graph(tmp_path):
    """cfg.cuda_graph = True: the denoiser-only training loop (conditioning from the dataset) replays one captured step per
    iteration (difffacto_b200/train_graph.py) - losses finite, every weight updated, gradient clipping inside the graph."""
    import difffacto_b200  # noqa: F401
    import difffacto_b200.datasets  # noqa: F401
    from difffacto_b200.config import init_cfg
    from difffacto_b200.runner import Runner
    p = tmp_path / "graph.py"
    p.write_text(textwrap.dedent(f"""
        _base_ = '{ROOT}/configs/gen_chair.py'
        model = dict(num_timesteps=8, npoints=256, ret_traj=False)
        dataset = dict(train=dict(type="SyntheticPartSeg", batch_size=4, npoints=256, n_parts=4, num_batches=3, seed=1),
                       val=dict(type="SyntheticPartSeg", batch_size=4, npoints=256, n_parts=4, num_batches=1, seed=0))
        optimizer = dict(type='Adam', lr=0.002, weight_decay=0.)
        max_epoch = 2
        max_norm = 10
        cuda_graph = True
        checkpoint_interval = 1
        log_interval = 2
        work_dir = '{tmp_path}/work'
    """))
    init_cfg(str(p))
    r = Runner("cuda:0", None)
    before = {k: v.clone() for k, v in r.diffusion.state_dict().items()}
    losses = r.run()
    assert losses.numel() == 6 and torch.isfinite(losses).all()
    after = r.diffusion.state_dict()
    assert sum(not torch.equal(before[k], after[k]) for k in before) == len(before) == 77
    assert os.path.exists(os.path.join(str(tmp_path), "work", "checkpoints", "ckpt_2.pth"))
