from .chamfer import ChamferDistanceL1, ChamferDistanceL2, ChamferDistanceL2_split, ChamferFunction  # noqa: F401
from .emd import EMD, emdFunction  # noqa: F401
from . import evaluation  # noqa: F401  (compute_all_metrics, _pairwise_EMD_CD_, knn, lgan_mmd_cov, ...)
