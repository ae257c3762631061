import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "denoiser_golden.npz"))


@pytest.fixture(scope="session")
def ref_ext():
    """The reference's own CUDA extensions built for sm_100a by oracle/build_ref.py (None if absent)."""
    import importlib.util
    out = {}
    d = os.path.join(ROOT, "oracle", "_ref")
    import torch  # noqa: F401  (the extensions link against torch)
    for name in ("ref_pointnet2_ext", "ref_chamfer", "ref_emd"):
        p = os.path.join(d, name + ".so")
        if os.path.exists(p):
            spec = importlib.util.spec_from_file_location(name, p)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            out[name] = mod
    return out


@pytest.fixture(scope="session")
def variants_golden():
    """DDIM / guidance outputs of the real reference (tests/golden/make_golden.py)."""
    import numpy as np
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "variants_golden.npz"))
