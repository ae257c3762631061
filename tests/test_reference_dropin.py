"""Build-container only (needs /root/reference): the drop-in claim of INTEGRATION.md section 2, executed.

  * `difffacto_b200.plugin.install()` into the REAL reference registry (python/difffacto/utils/registry.py:49-63);
    the reference's own `build_from_cfg` (:24-46) then returns the B200 classes for the type strings of
    configs/gen_chair.py, with the reference's state_dict keys;
  * the reference's `difffacto.metrics` package imports UNMODIFIED on top of this repo's `chamfer` / `emd` import shims
    (metrics/chamfer_dist/__init__.py:10 `import chamfer`, metrics/emd/emd_module.py:26 `import emd`) and `utils/misc.py:7`
    on top of the `pointnet2_ops` shim;
  * the reference's unmodified `AnchorDiffAE.decode` (models/networks/anchor_gen.py:145-169) drives the B200 generator
    with its own keyword set (the arithmetic needs a GPU: here the call must arrive at the "CPU not supported" guard).
Runs in a subprocess: the reference modules register themselves in process-wide registries.
"""
import os
import subprocess
import sys
import textwrap

import pytest

if not os.path.isdir("/root/reference/python/difffacto"):
    pytest.skip("/root/reference not present", allow_module_level=True)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent(r"""
    import importlib, sys, types
    sys.path.insert(0, %(root)r)                       # pointnet2_ops/, chamfer/, emd/ shims + difffacto_b200
    sys.path.insert(0, %(root)r + "/tests/golden")
    import torch
    REF = "/root/reference/python/difffacto"
    def ns(name, path):                                # namespace packages: no reference __init__ runs (tensorboardX, plyfile ...)
        m = types.ModuleType(name); m.__path__ = [path]; sys.modules[name] = m
    for name, sub in [("difffacto", ""), ("difffacto.utils", "/utils"), ("difffacto.models", "/models"),
                      ("difffacto.models.diffusions", "/models/diffusions"), ("difffacto.models.diffusions.nets", "/models/diffusions/nets"),
                      ("difffacto.models.networks", "/models/networks"),
                      ("difffacto.models.networks.language_utils", "/models/networks/language_utils")]:
        ns(name, REF + sub)
    reg = importlib.import_module("difffacto.utils.registry")
    importlib.import_module("difffacto.models.diffusions.nets.attention")
    importlib.import_module("difffacto.models.diffusions.anchored_diffusion")
    ref_metrics = importlib.import_module("difffacto.metrics")          # the REAL package __init__, on our chamfer / emd shims
    import chamfer, emd, pointnet2_ops
    assert "difffacto_b200" in chamfer.forward.__module__ or chamfer.__file__.startswith(%(root)r)
    assert emd.__file__.startswith(%(root)r) and pointnet2_ops.__file__.startswith(%(root)r)
    ref_diff_cls, ref_net_cls = reg.DIFFUSIONS.get("AnchoredDiffusion"), reg.NETS.get("TransformerNet")
    assert ref_diff_cls.__module__.startswith("difffacto.") and ref_net_cls.__module__.startswith("difffacto.")
    import make_golden as MG
    cfg = MG.gen_chair_diffusion_cfg()
    ref_keys = sorted(reg.build_from_cfg(cfg, reg.DIFFUSIONS, num_timesteps=10).state_dict())

    from difffacto_b200.plugin import install
    import difffacto_b200 as D
    replaced = install(reg)
    assert "DIFFUSIONS['AnchoredDiffusion']" in replaced and "NETS['TransformerNet']" in replaced and "METRICS['EMD']" in replaced, replaced
    ours = reg.build_from_cfg(cfg, reg.DIFFUSIONS, num_timesteps=10)    # the reference's own build_from_cfg
    assert type(ours) is D.DIFFUSIONS.get("AnchoredDiffusion") and type(ours.model) is D.NETS.get("TransformerNet")
    assert sorted(ours.state_dict()) == ref_keys and len(ref_keys) == 77
    assert type(reg.build_from_cfg(dict(type="EMD", eps=0.002, iters=50, dist_only=True), reg.METRICS)).__module__.startswith("difffacto_b200")
    try:
        reg.build_from_cfg(dict(type="NoSuchNet"), reg.NETS)
        raise SystemExit("unknown type must assert like the reference")
    except AssertionError as e:
        assert "not registered" in str(e)

    # the reference's decode(), unmodified, over the B200 generator
    ag = importlib.import_module("difffacto.models.networks.anchor_gen")
    # (save_pred_xstart=True makes the reference itself raise KeyError at t == T when T %% ret_interval == 0: its first yield has no pred_xstart)
    me = types.SimpleNamespace(diffusion=ours.eval(), ret_traj=True, ret_interval=5, save_pred_xstart=False)
    B, N = 2, 128
    anchors, variance = torch.zeros(B, 3, N), torch.ones(B, 3, N)
    ctx = [torch.zeros(B, 256, 4), torch.zeros(B, 6, 4)]
    try:
        ag.AnchorDiffAE.decode(me, anchors, ctx=ctx, variance=variance, anchor_assignments=torch.zeros(B, N, dtype=torch.int32),
                               valid_id=torch.ones(B, 4), device="cpu")
        raise SystemExit("expected the CPU guard")
    except RuntimeError as e:
        assert "CPU not supported" in str(e), e
    print("DROPIN_OK", len(replaced))
""")


def test_install_into_the_real_reference_registry_and_unmodified_callers():
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DROPIN_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
