"""CPU tests of the host-side mirror of the reference's plugin interface (registry, config, module
contracts) and of the multi-rank sharding logic (gloo, world_size 2)."""
import os
import textwrap

import numpy as np
import pytest
import torch

import difffacto_b200 as D
from difffacto_b200.config import Config
from difffacto_b200.parallel import gather_shapes, rank_seed, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_CONFIGS = "/root/reference/configs"


def test_registry_semantics():
    reg = D.Registry()

    @reg.register_module()
    class Foo:
        def __init__(self, a, b=2):
            self.a, self.b = a, b

    reg.register_module(name="Alias", module=Foo)
    assert D.build_from_cfg(dict(type="Foo", a=1), reg).b == 2
    assert D.build_from_cfg(dict(type="Alias", a=1), reg, b=5).b == 5   # kwargs override / extend the dict
    assert D.build_from_cfg("Foo", reg, a=3).a == 3
    assert D.build_from_cfg(None, reg) is None
    cfg = dict(type="Foo", a=1)
    D.build_from_cfg(cfg, reg)
    assert cfg == dict(type="Foo", a=1)                                  # the cfg dict is copied, not consumed
    with pytest.raises(AssertionError, match="is not registered"):
        D.build_from_cfg(dict(type="Nope"), reg)
    with pytest.raises(AssertionError, match="already registered"):
        reg.register_module(module=Foo)
    with pytest.raises(TypeError):
        D.build_from_cfg(dict(type="Foo"), reg)                           # missing required arg
    with pytest.raises(TypeError, match="not support"):
        D.build_from_cfg(3, reg)


def test_config_base_cover_and_missing_keys(tmp_path):
    (tmp_path / "base.py").write_text(textwrap.dedent("""
        model = dict(type='A', net=dict(depth=5, heads=8), keep=1)
        lr = 0.1
    """))
    (tmp_path / "child.py").write_text(textwrap.dedent("""
        _base_ = 'base.py'
        model = dict(net=dict(depth=3), extra=dict(_cover_=True, x=1))
        opt = dict(_cover_=True, type='Adam')
    """))
    c = Config(str(tmp_path / "child.py"))
    assert c.model.type == "A" and c.model.net.depth == 3 and c.model.net.heads == 8 and c.model.keep == 1
    assert c.model.extra.dump() == {"x": 1} and c.opt.dump() == {"type": "Adam"} and c.lr == 0.1
    assert c.name == "child" and c.work_dir == "work_dirs/child"
    assert c.this_key_does_not_exist is None and c.model.nope is None     # Runner relies on None for missing keys
    (tmp_path / "y.yaml").write_text("a: 1\nb: {c: 2}\n")
    y = Config(str(tmp_path / "y.yaml"))
    assert y.b.c == 2 and y.name == "y"


@pytest.mark.parametrize("name", ["gen_chair.py", "gen_airplane.py", "gen_car.py", "gen_lamp.py", "train_chair_stage1.py",
                                  "train_chair_stage2.py"])
def test_reference_configs_load_unmodified_and_resolve(name):
    path = os.path.join(REF_CONFIGS, name)
    if not os.path.exists(path):
        pytest.skip("reference configs are only present in the build container")
    c = Config(path)
    assert c.model.diffusion.type in D.DIFFUSIONS and c.model.diffusion.net.type in D.NETS
    diff = D.build_from_cfg(c.model.diffusion, D.DIFFUSIONS, num_timesteps=c.model.num_timesteps)
    assert diff.num_timesteps == c.model.num_timesteps
    assert sum(p.numel() for p in diff.model.parameters()) == 2615427


def test_repo_config_matches_reference_sampling_block():
    c = Config(os.path.join(ROOT, "configs", "gen_chair.py"))
    diff = D.build_from_cfg(c.model.diffusion, D.DIFFUSIONS, num_timesteps=c.model.num_timesteps)
    assert diff.num_timesteps == 100
    ref = os.path.join(REF_CONFIGS, "gen_chair.py")
    if os.path.exists(ref):
        r = Config(ref)
        assert c.model.diffusion.dump() == r.model.diffusion.dump()
        assert c.model.num_timesteps == r.model.num_timesteps and c.model.npoints == r.model.npoints


def test_module_contract_against_golden(golden):
    c = Config(os.path.join(ROOT, "configs", "gen_chair.py"))
    diff = D.build_from_cfg(c.model.diffusion, D.DIFFUSIONS, num_timesteps=100)
    assert sorted(diff.model.state_dict().keys()) == list(golden["state_dict_keys"])  # checkpoint key contract
    assert np.array_equal(diff.schedule_table(), golden["sched"])                       # float32(float64 tables), bit exact
    from difffacto_b200 import _lib
    from difffacto_b200.models.diffusions.nets.attention import pack_order
    assert len(pack_order(5)) == _lib.load().dfb200_denoiser_num_params(diff.model.c_cfg())
    from oracle.denoiser_ref import param_shapes
    shapes = {k: tuple(v.shape) for k, v in diff.model.state_dict().items()}
    assert shapes == param_shapes()


def test_unsupported_settings_fail_loudly():
    c = Config(os.path.join(ROOT, "configs", "gen_chair.py"))
    bad = c.model.diffusion.dump()
    bad["include_anchors"] = True
    with pytest.raises(NotImplementedError, match="include_anchors"):
        D.build_from_cfg(bad, D.DIFFUSIONS, num_timesteps=10)
    ddim = c.model.diffusion.dump()
    ddim.update(ddim_sampling=True, ddim_nsteps=5, ddim_discretize="uniform")
    assert D.build_from_cfg(ddim, D.DIFFUSIONS, num_timesteps=100).steps == [0, 20, 40, 60, 80]  # reference :117-119
    net = c.model.diffusion.net.dump()
    net["context_proj"] = True
    with pytest.raises(NotImplementedError, match="context_proj"):
        D.build_from_cfg(net, D.NETS)
    # no CPU fallback anywhere on the product path
    diff = D.build_from_cfg(c.model.diffusion, D.DIFFUSIONS, num_timesteps=10).eval()
    x = torch.zeros(1, 3, 128)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        diff.model(x, torch.zeros(1), [torch.zeros(1, 256, 4), torch.zeros(1, 6, 4)], anchors=x.transpose(1, 2),
                   variances=x.transpose(1, 2), valid_id=torch.ones(1, 4), anchor_assignment=torch.zeros(1, 128, dtype=torch.int32))


def test_product_never_imports_the_oracle():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import difffacto_b200, difffacto_b200.pointnet2_ops, difffacto_b200.parallel;"
            "assert not [m for m in sys.modules if m == 'oracle' or m.startswith('oracle.')]") % ROOT
    subprocess.check_call([sys.executable, "-c", code])
    for dirpath, _, files in os.walk(os.path.join(ROOT, "difffacto_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 32, 256, 257):
        for W in (1, 2, 3, 8):
            spans = [shard_range(total, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_rank_seed_is_unique_per_batch_and_rank():
    """Philox keys of (batch index, rank) must not collide: the kernels' counter is the LOCAL element index, so equal
    keys on two shards would draw identical x_T and step noise (the round-1 `seed + rank` did for consecutive seeds)."""
    assert rank_seed(5, 3, 4) == 23 and rank_seed(5, 0, 1) == 5
    for W in (1, 2, 8):
        keys = {rank_seed(7919 * s + bi, r, W) for s in range(3) for bi in range(50) for r in range(W)}
        assert len(keys) == 3 * 50 * W
    with pytest.raises(AssertionError):
        rank_seed(1, 4, 4)


def _gloo_worker(rank, world, total, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(total, rank, world)
    g = torch.Generator().manual_seed(1234)
    full = torch.randn(total, 16, 3, generator=g)           # every rank can rebuild the expected global batch
    out = gather_shapes(full[lo:hi].clone(), total)
    q.put((rank, bool(torch.equal(out, full))))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_gather_shapes_world_size_2_gloo(total):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + total
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, total, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_fused_adam_host_contract():
    """difffacto_b200/optim.py FusedAdam on the host side: torch.optim.Adam's constructor contract (defaults, param groups, state_dict
    keys), loud failures for what it does not implement, and no CPU fallback (a CPU parameter is an error, not a slow path)."""
    from difffacto_b200.optim import FusedAdam
    p = torch.nn.Parameter(torch.zeros(5))
    opt = FusedAdam([p], lr=2e-3, betas=(0.8, 0.9), eps=1e-6, weight_decay=0.1)
    g = opt.param_groups[0]
    assert (g["lr"], g["betas"], g["eps"], g["weight_decay"], g["capturable"]) == (2e-3, (0.8, 0.9), 1e-6, 0.1, True)
    assert set(opt.state_dict()) == {"state", "param_groups"}
    with pytest.raises(NotImplementedError):
        FusedAdam([p], amsgrad=True)
    for bad in (dict(lr=-1.0), dict(betas=(1.0, 0.9)), dict(betas=(0.9, -0.1)), dict(eps=-1e-8), dict(weight_decay=-0.1)):
        with pytest.raises(ValueError):
            FusedAdam([p], **bad)
    opt.step()                       # no gradient yet: nothing to do
    p.grad = torch.ones(5)
    with pytest.raises(RuntimeError, match="CUDA"):
        opt.step()
