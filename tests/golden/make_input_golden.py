"""Golden vectors of the input pipeline, produced by the REAL reference code and the real torchvision / Pillow
(run in the build container only):

  * MonoDataset.preprocess (mono/datasets/mono_dataset.py:84-103) executed on PIL frames with a `color_aug` that is
    torchvision's own ColorJitter: the parameters come from ColorJitter.get_params under a seeded torch RNG, are recorded,
    and the recorded-parameter replay is checked against calling the ColorJitter object itself from the same RNG state;
  * KITTIInpaintDataset.preprocess_masks (mono/datasets/kitti_dataset.py:167-182) under a seeded torch RNG, the box
    corners recovered from the mask it returns.

    python tests/golden/make_input_golden.py        ->  tests/golden/input/input_b3_24x40.pt
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
from PIL import Image
from torchvision import transforms
import torchvision.transforms.functional as TF

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("TDL_REFERENCE_ROOT", "/root/reference")


def load_datasets():
    """mono.datasets.{mono_dataset,kitti_dataset} without the package __init__ (which pulls the whole training stack);
    scipy.misc (removed from scipy) and kitti_utils are only needed by classes this script does not touch."""
    pkg = types.ModuleType("refds")
    pkg.__path__ = [os.path.join(REF, "mono", "datasets")]
    sys.modules["refds"] = pkg
    if "scipy.misc" not in sys.modules:
        import scipy
        sys.modules["scipy.misc"] = types.ModuleType("scipy.misc")
        scipy.misc = sys.modules["scipy.misc"]
    mods = {}
    for name in ("kitti_utils", "mono_dataset", "kitti_dataset"):
        spec = importlib.util.spec_from_file_location("refds." + name, os.path.join(REF, "mono", "datasets", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["refds." + name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods


class Cfg(dict):
    __getattr__ = dict.__getitem__


def main():
    mods = load_datasets()
    MonoDataset = mods["mono_dataset"].MonoDataset
    Inpaint = mods["kitti_dataset"].KITTIInpaintDataset
    B, H, W = 3, 24, 40
    frame_ids = [0, -1, 1]
    g = torch.Generator().manual_seed(4242)
    # band-limited + white content, so that hue sectors, grey pixels and saturated values all occur
    base = torch.rand(B, len(frame_ids), H, W, 3, generator=g)
    smooth = torch.nn.functional.avg_pool2d(base.permute(0, 1, 4, 2, 3).reshape(-1, 3, H, W), 5, 1, 2).reshape(B, len(frame_ids), 3, H, W).permute(0, 1, 3, 4, 2)
    frames = torch.where(torch.rand(B, 1, 1, 1, 1, generator=g) > 0.5, base, smooth * 1.6 - 0.3).clamp(0, 1)
    frames_u8 = (frames * 255).round().to(torch.uint8)
    frames_u8[0, 0, :4, :4] = 128                                  # a grey block (min == max: hue / saturation 0)
    frames_u8[0, 0, 4:8, :4] = torch.tensor([255, 0, 0], dtype=torch.uint8)
    frames_u8[0, 0, 8:12, :4] = 0
    frames_u8[0, 0, 12:16, :4] = 255

    ds = object.__new__(Inpaint)                                   # no files: only the methods under test
    ds.height, ds.width, ds.interp = H, W, Image.LANCZOS
    ds.to_tensor = transforms.ToTensor()
    ds.resize = transforms.Resize((H, W), interpolation=transforms.InterpolationMode.LANCZOS)
    ds.cfg = Cfg(erase_count=5, erase_shape=[6, 9], get=lambda *a, **k: False)
    ds.cfg.get = lambda k, d=None: d

    torch.manual_seed(99)
    jitter = torch.zeros(B, len(frame_ids), 4)
    hue_factor = torch.zeros(B, len(frame_ids), dtype=torch.float64)
    order = torch.zeros(B, len(frame_ids), 4, dtype=torch.int32)
    do_aug = torch.tensor([1, 0, 1], dtype=torch.uint8)
    color = torch.zeros(len(frame_ids), B, 3, H, W)
    color_aug = torch.zeros(len(frame_ids), B, 3, H, W)
    mask = torch.zeros(B, 3, H, W)
    holes = torch.zeros(B, 5, 2, dtype=torch.int32)
    for b in range(B):
        inputs = {("color", f, -1): Image.fromarray(frames_u8[b, i].numpy()) for i, f in enumerate(frame_ids)}
        cj = transforms.ColorJitter((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1))      # mono_dataset.py:65-68,183
        calls = {"n": 0}

        def color_aug_fn(img, b=b, cj=cj, calls=calls):
            # what ColorJitter.forward does, with its parameters recorded; verified against the object itself
            i = calls["n"]
            calls["n"] += 1
            state = torch.get_rng_state()
            fn_idx, bf, cf, sf, hf = cj.get_params(cj.brightness, cj.contrast, cj.saturation, cj.hue)
            out = img
            for fn in fn_idx:
                if fn == 0:
                    out = TF.adjust_brightness(out, bf)
                elif fn == 1:
                    out = TF.adjust_contrast(out, cf)
                elif fn == 2:
                    out = TF.adjust_saturation(out, sf)
                elif fn == 3:
                    out = TF.adjust_hue(out, hf)
            after = torch.get_rng_state()
            torch.set_rng_state(state)
            assert np.array_equal(np.array(cj(img)), np.array(out)), "replay != ColorJitter.forward"
            torch.set_rng_state(after)
            # preprocess visits ("color", f, 0) in dict order == frame order
            order[b, i] = fn_idx.to(torch.int32)
            jitter[b, i, 0], jitter[b, i, 1], jitter[b, i, 2] = bf, cf, sf
            jitter[b, i, 3] = float(np.int32(hf * 255).astype(np.uint8))
            hue_factor[b, i] = hf
            return out

        MonoDataset.preprocess(ds, inputs, color_aug_fn if do_aug[b] else (lambda x: x))
        for i, f in enumerate(frame_ids):
            color[i, b] = inputs[("color", f, 0)]
            color_aug[i, b] = inputs[("color_aug", f, 0)]
        # preprocess_masks: rows / cols come from torch.LongTensor(1).random_(...); recover them by replaying the draws
        state = torch.get_rng_state()
        Inpaint.preprocess_masks(ds, inputs)
        mask[b] = inputs[("mask", 0, 0)].float()
        torch.set_rng_state(state)
        for c_ in range(5):
            holes[b, c_, 0] = int(torch.LongTensor(1).random_(0, H - 6 - 1)[0])
            holes[b, c_, 1] = int(torch.LongTensor(1).random_(0, W - 9 - 1)[0])
    rec = dict(frame_ids=frame_ids, frames_u8=frames_u8, jitter=jitter, hue_factor=hue_factor, order=order, do_aug=do_aug,
               holes=holes, erase_shape=(6, 9), color=color, color_aug=color_aug, mask=mask,
               versions=dict(torchvision=__import__("torchvision").__version__, pillow=__import__("PIL").__version__))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "input")
    os.makedirs(out, exist_ok=True)
    torch.save(rec, os.path.join(out, "input_b3_24x40.pt"))
    print("wrote", out, rec["versions"], "aug differs on", float((color != color_aug).float().mean()))


if __name__ == "__main__":
    main()
