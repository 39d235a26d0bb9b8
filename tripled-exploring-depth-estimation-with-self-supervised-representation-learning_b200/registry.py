"""Name -> net-class table behind the reference's model-selection surface.

The reference picks its net with ``MONO.module_dict[cfg.model['name']](cfg.model)``
(train.py:98-99) after classes announce themselves with ``@MONO.register_module``
(mono/model/registry.py).  Only that surface is kept; the table itself is a plain dict
subclass, so ``MONO["mono_fm"]``, ``"mono_fm" in MONO`` and ``MONO.build(cfg)`` work too.
"""
from __future__ import annotations

import torch.nn as nn


class NetTable(dict):
    def __init__(self, label: str):
        super().__init__()
        self.label = label

    # -- reference-compatible surface ------------------------------------------------
    @property
    def name(self) -> str:
        return self.label

    @property
    def module_dict(self) -> "NetTable":
        return self

    def register_module(self, cls=None, *, name: str | None = None):
        """Usable bare (``@MONO.register_module``) or with an alias
        (``@MONO.register_module(name="alias")``)."""
        def add(klass):
            if not (isinstance(klass, type) and issubclass(klass, nn.Module)):
                raise TypeError(f"{self.label}: only nn.Module subclasses can be registered, got {klass!r}")
            key = name or klass.__name__
            if key in self:
                raise KeyError(f"{self.label}: '{key}' is already taken by {self[key].__module__}")
            self[key] = klass
            return klass
        return add if cls is None else add(cls)

    # -- convenience -------------------------------------------------------------------
    def build(self, model_cfg):
        """``model_cfg`` is the ``model = dict(name=..., ...)`` block of a config file."""
        key = model_cfg["name"]
        if key not in self:
            raise KeyError(f"{self.label}: unknown net '{key}' (known: {sorted(self)})")
        return self[key](model_cfg)


MONO = NetTable("mono")
