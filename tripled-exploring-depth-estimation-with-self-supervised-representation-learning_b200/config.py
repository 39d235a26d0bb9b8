"""mmcv-free stand-in for ``mmcv.Config.fromfile`` (train.py:70): the reference's configs are
plain Python modules (config/cfg_kitti_fm.py ...); nets read them as ``self.opt.<name>`` and
``self.opt.get(name, default)``.  mmcv 0.4.4 is not installable offline, so this shim executes
the file and wraps every dict in an attribute dict -- all option names are preserved."""
from __future__ import annotations

import os
import runpy


class ConfigDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value


def _wrap(v):
    if isinstance(v, dict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)):
        return type(v)(_wrap(x) for x in v)
    return v


class Config(ConfigDict):
    @staticmethod
    def fromfile(path):
        if not os.path.isfile(path):
            raise FileNotFoundError(path)
        ns = runpy.run_path(path)
        cfg = Config({k: _wrap(v) for k, v in ns.items()
                      if not k.startswith("_") and not callable(v) and not isinstance(v, type(os))})
        cfg["filename"] = path
        return cfg

    @staticmethod
    def fromdict(d):
        return Config({k: _wrap(v) for k, v in d.items()})
