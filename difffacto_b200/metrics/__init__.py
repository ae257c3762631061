from .chamfer import ChamferDistanceL1, ChamferDistanceL2, ChamferDistanceL2_split, ChamferFunction  # noqa: F401
from .emd import EMD, emdFunction  # noqa: F401
