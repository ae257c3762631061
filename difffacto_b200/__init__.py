"""difffacto_b200 -- B200-native (sm_100a) reverse-diffusion sampling hot path of DiffFacto.

Importing the package registers the drop-in classes under the reference's registry type strings
(NETS['TransformerNet'], DIFFUSIONS['AnchoredDiffusion'], METRICS[...]); all device work goes
through the C ABI of include/difffacto_b200.h (difffacto_b200/lib/libdifffacto_b200.so).
"""
from .utils.registry import (DATASETS, DIFFUSIONS, ENCODERS, HOOKS, METRICS, MODELS, NETS, OPTIMS, SAMPLERS,  # noqa: F401
                             SCHEDULERS, Registry, build_from_cfg)
from .models.diffusions import AnchoredDiffusion, TransformerNet  # noqa: F401
from .models.encoders import PartAlignerTransformer, PartEncoderForTransformerDecoder  # noqa: F401
from . import metrics  # noqa: F401

__version__ = "0.1.0"
