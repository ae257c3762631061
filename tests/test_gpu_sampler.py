"""-m gpu: the generalised fused reverse process (dfb200_sample_loop): the generator served chunk by chunk from the
persistent kernel, the reference's `decode` loop on top of it, strided DDIM step lists, classifier-free guidance and the
x_T trajectory slot -- each against the step-wise path (denoiser forward + update kernels), which the golden vectors of
tests/test_gpu_denoiser.py pin to the reference."""
import pytest
import torch

from oracle import denoiser_ref as R
from test_gpu_denoiser import DIFF_CFG, build, dev

pytestmark = pytest.mark.gpu


def build_variant(T, precision, **kw):
    import difffacto_b200 as D
    cfg = dict(DIFF_CFG)
    cfg.update(kw)
    d = D.build_from_cfg(cfg, D.DIFFUSIONS, num_timesteps=T)
    d.model.load_state_dict(R.synthetic_state_dict(1234), strict=True)
    d.model.precision = precision
    return d.cuda().eval()


def reference_decode(diffusion, anchors, ctx, variance, anchor_assignments, valid_id, ret_traj, ret_interval, noise=None, **gen_kw):
    """AnchorDiffAE.decode (python/difffacto/models/networks/anchor_gen.py:145-169), restated statement by statement
    (save_pred_xstart=False; /root/reference does not exist on the GPU box)."""
    final = dict()
    npts = anchor_assignments.shape[1]
    bs = anchors.shape[0]
    for t, sample in diffusion.p_sample_loop_progressive([bs, 3, npts], anchors=anchors, variance=variance, ctx=ctx, noise=noise,
                                                         anchor_assignment=anchor_assignments, valid_id=valid_id, device="cuda",
                                                         progress=False, **gen_kw):
        if t == 0:
            final["pred"] = sample["sample"].transpose(2, 1)
        elif ret_traj and t % ret_interval == 0:
            final[t] = sample["sample"].transpose(2, 1)
    return final


@pytest.mark.parametrize("precision,T,B,N", [("bf16", 70, 2, 256), ("fp32", 9, 2, 128), ("bf16", 7, 3, 384)])
def test_generator_from_the_fused_loop_equals_the_stepwise_generator(precision, T, B, N):
    """Same torch seed -> the chunked fused generator and the reference-style step loop yield the same (t, sample,
    pred_xstart) sequence (T = 70 in chunks of 16 steps = 5 persistent launches); yielded tensors are never clobbered."""
    d = build(T, precision)
    i = dev(R.synthetic_inputs(31, B, N, False))
    kw = dict(anchors=i["anchors"], ctx=[i["code"], i["params"]], variance=i["variance"], anchor_assignment=i["assign"], valid_id=i["valid"])
    torch.manual_seed(3)
    slow = list(d.p_sample_loop_progressive([B, 3, N], fused=False, **kw))
    torch.manual_seed(3)
    fast = list(d.p_sample_loop_progressive([B, 3, N], chunk=16, **kw))
    assert [t for t, _ in fast] == [t for t, _ in slow] == list(range(T, -1, -1))
    assert set(fast[0][1]) == {"sample"} and all(set(o) == {"sample", "pred_xstart"} for _, o in fast[1:])
    snap = [o["sample"].clone() for _, o in fast]
    worst = 0.0
    for (_, a), (_, b) in zip(fast, slow):
        for k in a:
            worst = max(worst, (a[k] - b[k]).abs().max().item())
    # fp32 mode: identical kernels and arithmetic on both sides.  bf16 mode: the step-wise forward builds K/V in one pass, the loop from
    # its hoisted static + time halves (another fp32 summation order): an occasional 1-ulp flip of a bf16 attention-fold entry
    assert worst <= (1e-6 if precision == "fp32" else 2e-2), worst
    torch.cuda.synchronize()
    assert all(torch.equal(o["sample"], s) for (_, o), s in zip(fast, snap))


@pytest.mark.parametrize("T,interval", [(20, 5), (12, 5)])
def test_decode_loop_over_the_generator_equals_the_one_call_loop(T, interval):
    """The reference's decode() loop over the generator == p_sample_loop(traj_interval=) from the same torch seed: 'pred' and
    every kept trajectory key, the x_T slot under key T included when T % ret_interval == 0."""
    B, N = 2, 256
    d = build(T, "bf16")
    i = dev(R.synthetic_inputs(32, B, N, False))
    ctx = [i["code"], i["params"]]
    torch.manual_seed(11)
    final = reference_decode(d, i["anchors"], ctx, i["variance"], i["assign"], i["valid"], True, interval)
    torch.manual_seed(11)
    x0, traj = d.p_sample_loop([B, 3, N], i["anchors"], ctx=ctx, variance=i["variance"], anchor_assignment=i["assign"], valid_id=i["valid"],
                               rng="torch", traj_interval=interval, device="cuda")
    keys = d.traj_keys(interval)
    assert sorted(k for k in final if k != "pred") == [t for t, _ in keys]
    assert (T in final) == (T % interval == 0)
    assert (final["pred"] - x0.transpose(2, 1)).abs().max().item() <= 1e-6
    for t, slot in keys:
        assert (final[t] - traj[slot].transpose(2, 1)).abs().max().item() <= 1e-6, t


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("kw", [dict(ddim_sampling=True, ddim_nsteps=10, ddim_discretize="uniform", ddim_eta=1.0),
                                dict(ddim_sampling=True, ddim_nsteps=8, ddim_eta=0.5, ddim_discretize="quad"),
                                dict(guidance=True, classifier_weight=2.0),
                                dict(guidance=True, classifier_weight=0.5, ddim_sampling=True, ddim_nsteps=6, ddim_discretize="uniform", ddim_eta=0.0)])
def test_fused_ddim_and_guidance_loops_equal_the_stepwise_path(kw, precision):
    """DDIM step lists / classifier-free guidance inside the fused loop vs p_sample step by step (which
    test_ddim_and_guidance_match_reference pins to the real reference) from the same noise; trajectory slots included."""
    T, B, N = 100 if kw.get("ddim_sampling") else 12, 2, 128
    d = build_variant(T, precision, **kw)
    i = dev(R.synthetic_inputs(12, B, N, True))
    ctx = [i["code"], i["params"]]
    steps = list(d.steps[::-1])
    g = torch.Generator(device="cuda").manual_seed(9)
    xT = torch.sqrt(i["variance"]) * torch.randn(B, 3, N, device="cuda", generator=g) + i["anchors"]
    zs = [torch.randn(B, 3, N, device="cuda", generator=g) for _ in steps]
    x, kept = xT, {}
    for z, step in zip(zs, steps):
        t = torch.full((B,), step, dtype=torch.long, device="cuda")
        x = d.p_sample(x, t, i["anchors"], ctx=ctx, variance=i["variance"], anchor_assignment=i["assign"], valid_id=i["valid"], noise=z)["sample"]
        kept[step] = x
    # same noise through the one-call loop: torch.randn is patched to replay the draws in the loop's call order
    draws = [zs[k] for k in range(len(steps))]
    real_randn = torch.randn
    try:
        torch.randn = lambda *a, out=None, **k: out.copy_(draws.pop(0)) if out is not None else real_randn(*a, **k)
        y, traj = d.p_sample_loop([B, 3, N], i["anchors"], ctx=ctx, noise=xT, variance=i["variance"], anchor_assignment=i["assign"],
                                  valid_id=i["valid"], rng="torch", traj_interval=4)
    finally:
        torch.randn = real_randn
    assert not draws
    tol = 1e-6 if precision == "fp32" else 2e-2  # bf16: K/V summation order of the hoisted tables, see the generator test
    assert (y - x).abs().max().item() <= tol, (y - x).abs().max().item()
    for t, slot in d.traj_keys(4):
        ref = xT if t == T else kept[t]
        assert (traj[slot] - ref).abs().max().item() <= tol, t


def test_fused_ddim_philox_loop_is_deterministic_and_tracks_fp32():
    """25-step DDIM (the reference's default ddim_nsteps) with in-kernel Philox noise: bit-reproducible, and the tcgen05 path stays
    within the per-step bf16 tolerance accumulated over 25 steps of the CUDA-core fp32 path on the same draws."""
    T, B, N = 1000, 4, 2048
    kw = dict(ddim_sampling=True, ddim_nsteps=25, ddim_discretize="quad", ddim_eta=1.0)
    d16, d32 = build_variant(T, "bf16", **kw), build_variant(T, "fp32", **kw)
    assert len(d16.steps) <= 25
    i = dev(R.synthetic_inputs(13, B, N, False))
    args = dict(ctx=[i["code"], i["params"]], variance=i["variance"], anchor_assignment=i["assign"], valid_id=i["valid"], rng="philox", seed=4)
    a = d16.p_sample_loop([B, 3, N], i["anchors"], **args)
    b = d16.p_sample_loop([B, 3, N], i["anchors"], **args)
    f = d32.p_sample_loop([B, 3, N], i["anchors"], **args)
    assert torch.isfinite(a).all() and torch.equal(a, b)
    scale = (f - i["anchors"]).abs().mean().item()
    assert (a - f).abs().mean().item() < 2e-2 * scale


def test_philox_normals_are_standard_normal():
    """dfb200_philox_normal (Philox4x32-10 + Box-Muller on the SFU intrinsics; the draw every sampling kernel inlines): moments and
    tails of 8M samples, independence of consecutive draws, and a different stream per (seed, offset)."""
    from difffacto_b200 import _lib
    lib = _lib.load()
    n = 1 << 23
    z = torch.empty(3, n, device="cuda")
    for k, (seed, off) in enumerate([(1, 0), (1, 1), (2, 0)]):
        _lib.check(lib.dfb200_philox_normal(_lib.ptr(z[k]), n, seed, off, _lib.stream()))
    torch.cuda.synchronize()
    a = z[0].double()
    assert abs(a.mean().item()) < 2e-3 and abs(a.var().item() - 1) < 3e-3
    assert abs((a ** 3).mean().item()) < 6e-3 and abs((a ** 4).mean().item() - 3) < 2e-2
    assert 4.5 < a.abs().max().item() < 6.5                      # 8M draws: max |z| ~ 5.3; u is 24-bit -> |z| <= 5.9
    for q, p in ((1.0, 0.841345), (2.0, 0.977250), (3.0, 0.998650)):
        assert abs((a < q).double().mean().item() - p) < 6e-4
    assert abs((a[:-1] * a[1:]).mean().item()) < 2e-3 and abs((a[0::2] * a[1::2]).mean().item()) < 2e-3  # lag-1 / in-pair correlation
    assert abs((z[0] * z[1]).double().mean().item()) < 2e-3 and abs((z[0] * z[2]).double().mean().item()) < 2e-3
    assert not torch.equal(z[0], z[1]) and not torch.equal(z[0], z[2])


@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16"])
def test_guidance_and_ddim_generator_at_odd_shapes(precision):
    """Units that straddle samples (N = 384: a 128-token tile per unit in guidance mode, 256-token units otherwise), an odd batch,
    a strided DDIM list served through the chunked generator (chunk = 3 steps -> several launches with tables_ready), all against
    the step-wise generator from the same torch seed."""
    T, B, N = 40, 3, 384
    d = build_variant(T, precision, guidance=True, classifier_weight=1.5, ddim_sampling=True, ddim_nsteps=7, ddim_discretize="uniform",
                      ddim_eta=0.7)
    i = dev(R.synthetic_inputs(33, B, N, False))
    kw = dict(anchors=i["anchors"], ctx=[i["code"], i["params"]], variance=i["variance"], anchor_assignment=i["assign"], valid_id=i["valid"])
    torch.manual_seed(8)
    slow = list(d.p_sample_loop_progressive([B, 3, N], fused=False, **kw))
    torch.manual_seed(8)
    fast = list(d.p_sample_loop_progressive([B, 3, N], chunk=3, **kw))
    assert [t for t, _ in fast] == [t for t, _ in slow] == [T] + list(d.steps[::-1])
    tol = 1e-6 if precision == "fp32" else 5e-3 if precision == "tf32" else 5e-2  # tensor-core modes: hoisted K/V tables (see above), x |w|+|1-w| = 2
    for (_, a), (_, b) in zip(fast, slow):
        for k in a:
            assert (a[k] - b[k]).abs().max().item() <= tol, (k, (a[k] - b[k]).abs().max().item())


def test_sample_loop_rejects_bad_arguments():
    """Status codes instead of undefined behaviour: unsorted step lists, ranges outside the list, DDIM without its tables,
    a workspace that is too small."""
    from difffacto_b200 import _lib
    lib = _lib.load()
    T, B, N = 10, 1, 128
    d = build(T, "bf16")
    i = dev(R.synthetic_inputs(1, B, N, True))
    st = d._loop_state(B, N, i["anchors"], [i["code"], i["params"]], i["variance"], i["assign"], i["valid"], torch.device("cuda"))
    x = torch.zeros(B, 3, N, device="cuda")

    def call(o, ws_bytes=None):
        return lib.dfb200_sample_loop(st["cfg"], _lib.ptr(st["packed"]), st["mode"], B, N, T, _lib.ptr(st["sched"]), _lib.ptr(x), 0,
                                      _lib.ptr(st["ctx"]), _lib.ptr(st["anchors"]), _lib.ptr(st["variance"]), _lib.ptr(st["assign"]),
                                      _lib.ptr(st["valid"]), None, 1, None, 1, o, _lib.ptr(st["ws"]), st["nws"] if ws_bytes is None else ws_bytes,
                                      _lib.stream())
    bad = torch.tensor([5, 7, 2], dtype=torch.int32, device="cuda")
    o = _lib.SampleOpts()
    o.timesteps, o.timesteps_host, o.n_timesteps = bad.data_ptr(), (_lib.c_int * 3)(5, 7, 2), 3
    assert call(o) == 1 and b"strictly decreasing" in lib.dfb200_last_error()
    o = _lib.SampleOpts()
    o.first_step, o.num_steps = 8, 5
    assert call(o) == 1 and b"outside the list" in lib.dfb200_last_error()
    o = _lib.SampleOpts()
    o.ddim = 1
    assert call(o) == 1 and b"DDIM" in lib.dfb200_last_error()
    assert call(_lib.SampleOpts(), ws_bytes=16) == 4 and b"workspace too small" in lib.dfb200_last_error()
    assert call(_lib.SampleOpts()) == 0  # and the all-defaults call runs
    torch.cuda.synchronize()
    assert torch.isfinite(x).all()
