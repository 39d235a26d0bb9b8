"""Golden vectors of the pose prologue, produced by the REAL reference (run in the build container only):
mono_fm.transformation_from_parameters / rot_from_axisangle / get_translation_matrix
(mono/model/mono_fm/net.py:201-253) on seeded axis-angle / translation draws, both `invert` settings, with the
gradients of <W, T> for a fixed random W.

    python tests/golden/make_pose_golden.py        ->  tests/golden/pose/pose_b16.pt
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def main():
    mods = ref_loader.load()
    cls = mods["fm"].mono_fm
    obj = types.SimpleNamespace()
    for name in ("transformation_from_parameters", "rot_from_axisangle", "get_translation_matrix"):
        setattr(obj, name, types.MethodType(getattr(cls, name), obj))
    g = torch.Generator().manual_seed(777)
    B = 16
    # the pose decoder scales its outputs by 0.01 (pose_decoder.py:23); a few larger and one exactly-zero rotation too
    aa = 0.01 * torch.randn(B, 1, 3, generator=g)
    aa[1] *= 50.0
    aa[2] *= 200.0
    aa[3] = 0.0
    tr = 0.01 * torch.randn(B, 1, 3, generator=g)
    tr[4] *= 100.0
    W = torch.randn(B, 4, 4, generator=g)
    rec = {"axisangle": aa, "translation": tr, "W": W}
    with ref_loader.cpu_cuda_shim():
        for inv in (False, True):
            a = aa.clone().requires_grad_(True)
            t = tr.clone().requires_grad_(True)
            T = obj.transformation_from_parameters(a, t, inv)
            (T * W).sum().backward()
            rec[("T", inv)] = T.detach().clone()
            rec[("d_axisangle", inv)] = a.grad.clone()
            rec[("d_translation", inv)] = t.grad.clone()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pose", "pose_b16.pt")
    torch.save(rec, out)
    print("wrote", out, {k: tuple(v.shape) for k, v in rec.items()})


if __name__ == "__main__":
    main()
