"""Import-name shim: `import pointnet2_ops` / `from pointnet2_ops import pointnet2_utils` resolve to
the B200 implementation (difffacto_b200.pointnet2_ops), as reference callers expect
(e.g. python/difffacto/utils/misc.py:7, models/encoders/pointnet2.py:3)."""
import sys

from difffacto_b200.pointnet2_ops import pointnet2_modules, pointnet2_utils  # noqa: F401

sys.modules[__name__ + ".pointnet2_utils"] = pointnet2_utils
sys.modules[__name__ + ".pointnet2_modules"] = pointnet2_modules
__version__ = "3.0.0+b200"
