"""Import the *real* reference loss classes from ``/root/reference``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  ``/root/reference``
exists only in the build container; ``oracle/make_ref.py`` places the unmodified loss
modules under the git-ignored ``oracle/_ref/`` so that they also reach the GPU box.
Everything that needs them is gated on ``reference_available()``: ``tests/test_oracle_vs_reference.py``
(pins ``oracle.restatement`` bit-for-bit against the reference), ``tests/golden/make_golden.py``
(writes the committed golden vectors) and the ``--impl reference`` / ``cpu_baseline`` legs of ``bench.py``.

The reference does not import as shipped (SURVEY.md "five facts" 3 and 4), so
the loader
  1. pre-seeds ``sys.modules['mono']`` / ``['mono.model']`` with empty
     namespace modules to skip the broken package ``__init__``s
     (mono/model/__init__.py:9-10 imports a missing package),
  2. stubs ``matplotlib`` and ``torchvision.models.utils`` (only imported, never
     used, by mono/model/mono_fm_joint/diffnet_encoder.py:6,8),
  3. makes ``Tensor.cuda`` the identity while a reference loss runs on the CPU
     (the loss hard-codes ``.cuda()``: mono/model/mono_fm/layers.py:58,60,
     mono/model/mono_fm/net.py:94).
None of the reference's sources are copied; they are executed where they lie.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    """The build container has the reference checkout; the GPU box only has ``oracle/_ref`` (the unmodified loss
    modules placed there by ``oracle/make_ref.py``, git-ignored, checked against their SHA-256 manifest)."""
    env = os.environ.get("TDL_REFERENCE_ROOT", "/root/reference")
    if os.path.isfile(os.path.join(env, "mono", "model", "mono_fm", "net.py")):
        return env, "checkout"
    ref = os.path.join(_HERE, "_ref")
    if os.path.isfile(os.path.join(ref, "mono", "model", "mono_fm", "net.py")):
        from . import make_ref
        if make_ref.verify(ref):
            return ref, "oracle/_ref"
    return env, None


REF_ROOT, REF_KIND = _find_root()


def reference_available() -> bool:
    return REF_KIND is not None


def _stub(name, path=None, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    if path:
        m.__path__ = [path]
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Returns a dict of the reference modules on the loss path."""
    if _loaded:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not found under {REF_ROOT}")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if "mono" not in sys.modules:
        _stub("mono", f"{REF_ROOT}/mono")
        _stub("mono.model", f"{REF_ROOT}/mono/model")
    if "torchvision.models.utils" not in sys.modules:
        _stub("torchvision.models.utils",
              load_state_dict_from_url=torch.hub.load_state_dict_from_url)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        _stub("matplotlib", "/nonexistent")
        _stub("matplotlib.pyplot")
    _loaded["fm"] = importlib.import_module("mono.model.mono_fm.net")
    _loaded["baseline"] = importlib.import_module("mono.model.mono_baseline.net")
    _loaded["layers"] = importlib.import_module("mono.model.mono_fm.layers")
    try:
        _loaded["inpaint"] = importlib.import_module("mono.model.mono_fm_joint_inpaint.net")
        _loaded["joint"] = importlib.import_module("mono.model.mono_fm_joint.net")
    except Exception as e:  # pragma: no cover - optional families
        _loaded["inpaint_error"] = repr(e)
    return _loaded


@contextlib.contextmanager
def cpu_cuda_shim(force=False):
    """``Tensor.cuda`` -> identity while the reference loss runs on CPU tensors (always when there is no GPU;
    ``force`` for the CPU arm of bench.py on the GPU box, where the reference's hard-coded ``.cuda()`` calls would
    otherwise move its constants to the device)."""
    if torch.cuda.is_available() and not force:
        yield
        return
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


class Opt(dict):
    """Stands in for the mmcv Config node the nets read as ``self.opt``."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def default_opt(batch, height, width, **kw):
    o = Opt(frame_ids=[0, -1, 1], imgs_per_gpu=batch, height=height, width=width,
            scales=[0, 1, 2, 3], min_depth=0.1, max_depth=100.0, automask=True,
            disp_norm=True, perception_weight=1e-3, smoothness_weight=1e-3,
            disparity_smoothness=1e-3)
    o.update(kw)
    return o


class _Feat(nn.Module):
    """Stand-in for ``extractor(img)[0]``: returns the pre-computed feature map
    registered for the image tensor it is called with (keyed by data_ptr)."""

    def __init__(self):
        super().__init__()
        self.table = {}

    def forward(self, img):
        return [self.table[img.data_ptr()]]


def make_loss_only_net(kind, opt):
    """Builds a reference net object WITHOUT its networks (``__new__`` +
    ``nn.Module.__init__``), carrying only what ``compute_losses`` reads.

    kind: 'baseline' (mono/model/mono_baseline/net.py:51-100),
          'fm'       (mono/model/mono_fm/net.py:69-133),
          'joint'    (mono/model/mono_fm_joint/net.py:73-155),
          'inpaint'  (mono/model/mono_fm_joint_inpaint/net.py:47-133),
          'tripled'  (mono/model/mono_fm_joint_inpaint/net.py:398-532, adds auto_res_loss).
    """
    mods = load()
    L = mods["layers"]
    if kind == "baseline":
        cls = mods["baseline"].Baseline
    elif kind == "fm":
        cls = mods["fm"].mono_fm
    elif kind == "inpaint":
        cls = mods["inpaint"].mono_fm_joint_inpaint
    elif kind == "joint":                        # mono/model/mono_fm_joint/net.py:73-155
        cls = mods["joint"].mono_fm_joint
    elif kind == "tripled":                      # the TripleD net of config/cfg_kitti_tripleD.py
        cls = mods["inpaint"].mono_fm_joint_inpaint_disentangle
    else:
        raise ValueError(kind)
    net = cls.__new__(cls)
    nn.Module.__init__(net)
    net.opt = opt
    net.ssim = L.SSIM()
    net.backproject = L.Backproject(opt.imgs_per_gpu, opt.height, opt.width)
    proj = L.Project(opt.imgs_per_gpu, opt.height, opt.width)
    net.project = proj
    net.project_3d = proj
    if kind == "fm":
        net.extractor = _Feat()
    if kind in ("inpaint", "tripled", "joint"):
        net.Encoder = _Feat()
    return net


def run_reference_loss(kind, opt, inputs, outputs, tgt_feat=None, src_feats=None, features=None):
    """Executes the reference's own ``compute_losses`` (the class named by ``kind``, see make_loss_only_net) on the
    given tensors -> loss_dict.  ``outputs`` is updated in place like the reference does (warped images, min_index).
    Feature maps stand in for the extractor's output through the data_ptr table of ``_Feat``."""
    net = make_loss_only_net(kind, opt)
    if tgt_feat is not None:
        table = {inputs[("color", 0, 0)].data_ptr(): tgt_feat}
        for f, t in src_feats.items():
            table[inputs[("color", f, 0)].data_ptr()] = t
        (net.extractor if kind == "fm" else net.Encoder).table = table
    with cpu_cuda_shim(force=not inputs[("color", 0, 0)].is_cuda):
        if kind in ("inpaint", "tripled", "joint"):
            return net.compute_losses(inputs, outputs, features)
        return net.compute_losses(inputs, outputs)
