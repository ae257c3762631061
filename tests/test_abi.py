"""The C-ABI library loads on a CPU-only box and exports every symbol include/difffacto_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols(name="difffacto_b200.h"):
    hdr = open(os.path.join(ROOT, "include", name)).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(dfb200_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from difffacto_b200 import _lib
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 26
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES out of sync with the header"
    assert lib.dfb200_abi_version() == 1
    assert isinstance(lib.dfb200_launch_count(), int)


def test_diagnostics_are_not_in_the_product_library():
    """Self-tests / microbenchmarks / the timeline hook live only in the -DDFB200_DIAGNOSTICS build."""
    from difffacto_b200 import _lib
    lib = _lib.load()
    diag_names = header_symbols("difffacto_b200_diag.h")
    assert sorted(_lib.DIAG_SIGNATURES) == diag_names and len(diag_names) == 5
    for n in diag_names:
        assert not hasattr(lib, n), f"{n} is a diagnostic and must not be exported by the product library"
    dlib = _lib.load_diag()
    for n in diag_names + header_symbols():
        assert hasattr(dlib, n), f"{n} missing from the diagnostic build"


def test_no_torch_in_abi():
    # plain C boundary: the .so must not link torch / ATen / c10
    from difffacto_b200 import _lib
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "c10" not in out, out


def test_cpu_tensors_are_rejected_like_the_reference():
    # reference: AT_ASSERT(false, "CPU not supported") (ball_query.cpp:28 etc.)
    import torch
    from difffacto_b200.pointnet2_ops import pointnet2_utils as pu
    xyz = torch.rand(1, 16, 3)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        pu.furthest_point_sample(xyz, 4)
    with pytest.raises(RuntimeError, match="CPU not supported"):
        pu.ball_query(0.2, 4, xyz, xyz[:, :2].contiguous())
    with pytest.raises(RuntimeError, match="must be a contiguous tensor"):
        pu.furthest_point_sample(torch.rand(1, 3, 16).transpose(1, 2), 4)
    with pytest.raises(RuntimeError, match="int tensor"):
        pu.gather_operation(torch.rand(1, 3, 16), torch.zeros(1, 4, dtype=torch.int64))


def test_denoiser_cfg_validation_and_sizes():
    from difffacto_b200 import _lib
    lib = _lib.load()
    flags = _lib.NET_CLASS_COND | _lib.NET_CAT_PARAMS_TO_X | _lib.NET_CAT_CLASS_TO_X | _lib.NET_MASK_UNREFERENCED
    cfg = _lib.DenoiserCfg(3, 3, 8, 16, 5, 262, 4, flags)
    assert lib.dfb200_denoiser_num_params(cfg) == 12 + 13 * 5
    assert lib.dfb200_denoiser_packed_bytes(cfg) >= 2615427 * 4
    assert lib.dfb200_denoiser_workspace_bytes(cfg, _lib.MODE_FP32, 2, 256) > 2 * 256 * (128 + 128 + 512) * 4
    bad = _lib.DenoiserCfg(3, 3, 4, 16, 5, 262, 4, flags)
    assert lib.dfb200_denoiser_num_params(bad) == -1
    assert b"inner_dim" in lib.dfb200_last_error()
