"""Drop-in for the reference's `pointnet2_ops` package (pointnet2_ops_lib/pointnet2_ops/)."""
from . import pointnet2_modules, pointnet2_utils  # noqa: F401

__version__ = "3.0.0+b200"
