"""-m gpu: world_size 2 on ONE GPU (both ranks on cuda:0, gloo backend -- NCCL refuses two ranks per device): the denoiser's
training loss under DistributedDataParallel.  DDP all-reduces the gradients that difffacto_b200's autograd Functions hand to the
parameters (reference: runner/runner.py:67 wraps the model the same way); the averaged gradient of two half-batches must equal
the single-process gradient of the full batch (the loss is a mean over points, shards are equal)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import denoiser_ref as R
    from test_gpu_denoiser import DIFF_CFG
    import difffacto_b200 as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    T, B, N = 50, 4, 256
    d = D.build_from_cfg(DIFF_CFG, D.DIFFUSIONS, num_timesteps=T)
    d.model.load_state_dict(R.synthetic_state_dict(1234), strict=True)
    d = d.cuda().eval()  # eval: dropout off (its Philox mask depends on the batch layout), gradients still flow
    inp = {k: v.cuda() for k, v in R.synthetic_inputs(41, B, N, False).items()}
    g = torch.Generator(device="cuda").manual_seed(5)
    x0 = torch.sqrt(inp["variance"]) * torch.randn(B, 3, N, device="cuda", generator=g) + inp["anchors"]
    noise = torch.randn(B, 3, N, device="cuda", generator=g)
    t = torch.randint(0, T, (B,), device="cuda", generator=g)
    flags = torch.ones(B, 1, N, device="cuda")

    def loss_of(module, lo, hi):
        kw = dict(anchors=inp["anchors"][lo:hi], variance=inp["variance"][lo:hi], ctx=[inp["code"][lo:hi], inp["params"][lo:hi]],
                  anchor_assignment=inp["assign"][lo:hi], valid_id=inp["valid"][lo:hi], flags=flags[lo:hi], noise=noise[lo:hi])
        return module(x0[lo:hi], t[lo:hi], **kw)["mse_loss"]

    full = None
    if rank == 0:  # reference gradient: the whole batch in one process, no DDP
        loss_of(d, 0, B).backward()
        full = {n: p.grad.clone() for n, p in d.named_parameters()}
        d.zero_grad(set_to_none=True)
    ddp = torch.nn.parallel.DistributedDataParallel(d, device_ids=[0])
    lo, hi = rank * B // world, (rank + 1) * B // world
    loss_of(ddp, lo, hi).backward()
    torch.cuda.synchronize()
    if rank == 0:
        worst, n_checked = 0.0, 0
        for n, p in d.named_parameters():
            ref = full[n]
            worst = max(worst, ((p.grad - ref).abs().max() / (ref.abs().max() + 1e-12)).item())
            n_checked += 1
        q.put((n_checked, worst))
    dist.barrier()
    dist.destroy_process_group()


def test_ddp_world2_gradients_equal_full_batch():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29655, q)) for r in range(2)]
    for p in procs:
        p.start()
    n_checked, worst = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert n_checked == 77
    assert worst < 1e-4, worst  # atomics order in the split-K weight gradients: ~1e-6 relative
