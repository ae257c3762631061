"""Class registries + build_from_cfg: the plugin boundary of the reference.

Same contract as reference python/difffacto/utils/registry.py:1-63: `@REG.register_module()`
registers a class under its __name__ (or an explicit name), `REG.get(name)` asserts the name is
known, `build_from_cfg(cfg, REG, **kw)` accepts a type string, a dict with a `type` key (copied,
kwargs merged, `type` popped, class called) or None.  Configs written for the reference resolve
to the classes registered here under the same type strings.
"""


class Registry:
    def __init__(self, name=None):
        self.name = name
        self._modules = {}

    def register_module(self, name=None, module=None):
        def deco(cls):
            key = cls.__name__ if name is None else name
            assert key not in self._modules, f"{key} is already registered."
            self._modules[key] = cls
            return cls

        return deco(module) if module is not None else deco

    def get(self, name):
        assert name in self._modules, f"{name} is not registered."
        return self._modules[name]

    def __contains__(self, name):
        return name in self._modules

    def keys(self):
        return self._modules.keys()


def build_from_cfg(cfg, registry, **kwargs):
    if cfg is None:
        return None
    if isinstance(cfg, str):
        return registry.get(cfg)(**kwargs)
    if isinstance(cfg, dict):
        args = dict(cfg)
        args.update(kwargs)
        cls = registry.get(args.pop("type"))
        try:
            return cls(**args)
        except TypeError as e:
            msg = str(e)
            raise TypeError(msg if "<class" in msg else f"{cls}.{msg}")
    if isinstance(cfg, (list, tuple)):
        import torch.nn as nn
        return nn.Sequential(*[build_from_cfg(c, registry, **kwargs) for c in cfg])
    raise TypeError(f"type {type(cfg)} not support")


DATASETS = Registry("datasets")
MODELS = Registry("models")
ENCODERS = Registry("encoders")
DECOMPOSERS = Registry("decomposers")
DIFFUSIONS = Registry("diffusions")
NETS = Registry("nets")
SCHEDULERS = Registry("schedulers")
HOOKS = Registry("hooks")
LOSSES = Registry("losses")
OPTIMS = Registry("optims")
SAMPLERS = Registry("samplers")
METRICS = Registry("metrics")
SEGMENTORS = Registry("segmentors")
GENERATORS = Registry("generators")
DISCRIMINATORS = Registry("discriminators")
