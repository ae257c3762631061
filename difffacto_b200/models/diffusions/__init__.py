from .anchored_diffusion import AnchoredDiffusion  # noqa: F401
from .nets import TransformerNet  # noqa: F401
