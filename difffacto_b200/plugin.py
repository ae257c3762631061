"""Install the B200 classes into ANOTHER registry module (the reference's `difffacto.utils.registry`),
overriding the entries the shipped configs resolve: see INTEGRATION.md section 2a."""
from . import metrics as _metrics
from .models.diffusions import AnchoredDiffusion, TransformerNet

_OVERRIDES = {
    "DIFFUSIONS": {"AnchoredDiffusion": AnchoredDiffusion},
    "NETS": {"TransformerNet": TransformerNet},
    "METRICS": {"ChamferDistanceL2": _metrics.ChamferDistanceL2, "ChamferDistanceL2_split": _metrics.ChamferDistanceL2_split,
                "ChamferDistanceL1": _metrics.ChamferDistanceL1, "EMD": _metrics.EMD},
}


def install(registry_module):
    """registry_module: any module exposing Registry objects named DIFFUSIONS / NETS / METRICS whose entries live in
    a `_modules` dict (reference python/difffacto/utils/registry.py:1-22).  Existing entries are replaced."""
    replaced = []
    for reg_name, classes in _OVERRIDES.items():
        reg = getattr(registry_module, reg_name, None)
        if reg is None:
            continue
        for key, cls in classes.items():
            reg._modules[key] = cls
            replaced.append(f"{reg_name}[{key!r}]")
    # the diffusion builds its net through the registry it was given: make ours resolve there as well
    from .models.diffusions import anchored_diffusion as ad
    if getattr(registry_module, "NETS", None) is not None:
        ad.NETS = registry_module.NETS
    return replaced
