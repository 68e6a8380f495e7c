"""TEST INFRASTRUCTURE ONLY - registry / builder subset of nncore.nn (see ../__init__.py)."""
import torch.nn as _nn

Parameter = _nn.Parameter


class Registry:
    def __init__(self, name):
        self.name = name
        self._items = {}

    def register(self, name=None):
        def deco(cls):
            self._items[name or cls.__name__] = cls
            return cls
        return deco

    def get(self, key):
        return self._items[key]


MODELS = Registry("model")
LOSSES = Registry("loss")


def _build(registry, cfg, *args, **kwargs):
    if cfg is None:
        return None
    cfg = dict(cfg)
    typ = cfg.pop("type")
    cfg.update(kwargs)
    return registry.get(typ)(*args, **cfg)


def build_model(cfg, *args, **kwargs):
    return _build(MODELS, cfg, *args, **kwargs)


def build_loss(cfg, *args, **kwargs):
    try:
        return _build(LOSSES, cfg, *args, **kwargs)
    except KeyError:
        return None  # training-only losses (FocalLoss / L1Loss) are out of scope
