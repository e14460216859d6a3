"""The drop-in boundary: mmengine-style registries.

The reference resolves `type=` strings of its configs through
`mmseg.registry.MODELS` (Segmentation/mmseg/registry/registry.py:35-56) and
`mmdet.registry.MODELS` (Segmentation/mmdet/registry.py:62-121).  When mmseg /
mmdet / mmengine are importable the classes of this package are registered
into those registries (force=True) so the reference configs resolve to the
B200 implementations.  When they are not importable (this image), the small
registry below offers the same `register_module()` / `build()` contract
including the `mmdet.` scope prefix and ConfigDict attribute access.
"""
from __future__ import annotations


class ConfigDict(dict):
    """dict with attribute access -- detr_layers.py:307 reads `self_attn_cfg.embed_dims`."""

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    __setattr__ = dict.__setitem__


def to_config(obj):
    if isinstance(obj, dict):
        return ConfigDict({k: to_config(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return type(obj)(to_config(v) for v in obj)
    return obj


class Registry:
    def __init__(self, name, scope):
        self.name, self.scope = name, scope
        self._table = {}

    def register_module(self, name=None, force=False, module=None):
        def wrap(cls):
            key = name or cls.__name__
            if key in self._table and not force and self._table[key] is not cls:
                raise KeyError(f"{key} is already registered in {self.name}")
            self._table[key] = cls
            return cls

        return wrap(module) if module is not None else wrap

    def get(self, key):
        key = key.split(".")[-1]          # 'mmdet.X' -> 'X': scopes share one table here
        if key not in self._table:
            raise KeyError(f"{key} is not in the {self.name} registry")
        return self._table[key]

    def build(self, cfg, default_args=None):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise TypeError("cfg must be a dict containing the key 'type'")
        cfg = to_config(dict(cfg))
        if default_args:
            for k, v in default_args.items():
                cfg.setdefault(k, v)
        cls = self.get(cfg.pop("type"))
        return cls(**cfg)

    def __contains__(self, key):
        return key.split(".")[-1] in self._table


MODELS = Registry("model", scope="mmseg")


def register_everywhere(cls, scopes=("mmseg", "mmdet")):
    """Register in the built-in table and, when available, in the real mmseg / mmdet registries."""
    MODELS.register_module(force=True)(cls)
    for scope in scopes:
        try:
            reg = __import__(f"{scope}.registry", fromlist=["MODELS"]).MODELS
            reg.register_module(force=True)(cls)
        except Exception:
            pass
    return cls
