"""Dict-style run configuration with `_base_` inheritance and `_cover_` override.

Behavioural mirror of reference python/difffacto/config/config.py:16-153: a config is a .py file
(its module globals) or a .yaml file; `_base_` names one or more parent files merged first;
a dict carrying `_cover_: True` replaces the inherited dict instead of merging into it; attribute
access on a missing key returns None (the Runner relies on that); `name` / `work_dir` default to
the file stem / work_dirs/<name>.  Reference configs load unmodified.
"""
import copy
import importlib.util
import inspect
import os
from collections import OrderedDict

import yaml

BASE_KEY, COVER_KEY = "_base_", "_cover_"


def _strip_cover(node):
    if not isinstance(node, dict):
        return node
    return {k: _strip_cover(v) for k, v in copy.deepcopy(node).items() if k != COVER_KEY}


def _merge(dst, src):
    """merge src into dst in place (reference merge_dict_b2a semantics)."""
    if COVER_KEY in src:
        dst.clear()
        dst.update(_strip_cover(src))
        return
    for k, v in src.items():
        both_dicts = isinstance(v, dict) and isinstance(dst.get(k), dict)
        if k in dst and both_dicts and not v.get(COVER_KEY, False):
            _merge(dst[k], v)
        else:
            dst[k] = _strip_cover(v)


def _read_one(path):
    ext = os.path.splitext(path)[1]
    if ext == ".yaml":
        with open(path) as f:
            return yaml.safe_load(f.read())
    if ext == ".py":
        spec = importlib.util.spec_from_file_location("_dfb200_cfg_" + os.path.basename(path)[:-3], path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return {k: v for k, v in vars(mod).items() if not k.startswith("__")}
    raise AssertionError("unsupported config type.")


def _read(path):
    cfg = _read_one(path)
    if BASE_KEY in cfg:
        bases = cfg.pop(BASE_KEY)
        bases = [bases] if isinstance(bases, str) else list(bases)
        merged = {}
        for b in bases:
            _merge(merged, _read(os.path.join(os.path.dirname(path), b)))
        _merge(merged, cfg)
        cfg = merged
    return cfg


class Config(OrderedDict):
    def __init__(self, *args):
        super().__init__()
        assert len(args) <= 1
        if args:
            self.load_from_file(args[0])

    def __getattr__(self, name):
        return self[name] if name in self else None

    def __setattr__(self, name, value):
        self[name] = value

    @classmethod
    def _wrap(cls, node):
        if isinstance(node, dict):
            out = cls()
            for k, v in node.items():
                if not inspect.ismodule(v):
                    out[k] = cls._wrap(v)
            return out
        if isinstance(node, list):
            return [cls._wrap(v) for v in node if not inspect.ismodule(v)]
        return copy.deepcopy(node)

    def load_from_file(self, filename):
        tree = _read(filename)
        self.clear()
        self.update(self._wrap(tree))
        if self.name is None:
            self.name = os.path.splitext(os.path.basename(filename))[0]
        if self.work_dir is None:
            self.work_dir = f"work_dirs/{self.name}"

    def dump(self):
        def plain(v):
            if isinstance(v, Config):
                return {k: plain(x) for k, x in v.items()}
            if isinstance(v, list):
                return [plain(x) for x in v]
            return v
        return plain(self)


_cfg = Config()


def init_cfg(filename):
    print("Loading config from: ", filename)
    _cfg.load_from_file(filename)


def get_cfg():
    return _cfg


def update_cfg(**kwargs):
    _cfg.update(kwargs)


def save_cfg(save_file):
    with open(save_file, "w") as f:
        f.write(yaml.dump(_cfg.dump()))


def print_cfg():
    print(yaml.dump(_cfg.dump()))
