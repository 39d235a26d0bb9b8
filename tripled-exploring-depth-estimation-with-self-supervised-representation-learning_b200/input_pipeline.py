"""On-GPU input pipeline: the per-item host work of the reference's data loader, for a whole batch on the device.

Reference (all executed per item inside DataLoader workers, on PIL images):

    MonoDataset.preprocess                 mono/datasets/mono_dataset.py:84-103   to_tensor, color_aug
    MonoDataset.__getitem__                mono/datasets/mono_dataset.py:140-141,182-187   do_color_aug / do_flip, ColorJitter
    KITTIInpaintDataset.preprocess_masks   mono/datasets/kitti_dataset.py:167-182  erase masks

Here the loader only has to hand over the resized uint8 frames (B,H,W,3); `GpuInputPipeline` samples the augmentation
parameters on the host the way torchvision's ColorJitter.get_params does (a few numbers per image) and one call of
libtdl.so's tdl_input_fwd (two launches) writes inputs[("color", f, 0)], inputs[("color_aug", f, 0)] and
inputs[("mask", 0, 0)] -- byte-exact against torchvision / Pillow on the same parameters (tests/test_input_pipeline.py).
No CPU fallback: the tensors must live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def hue_shift_byte(hue_factor: float) -> int:
    """The byte torchvision's PIL adjust_hue adds to the H channel (`np.int32(hue_factor * 255).astype(np.uint8)`)."""
    return int(np.int32(hue_factor * 255).astype(np.uint8))


class GpuInputPipeline:
    """Drop-in for MonoDataset.preprocess (+ KITTIInpaintDataset.preprocess_masks) at batch level.

    frame_ids, height, width: as in the dataset config; erase_count / erase_shape: cfg_kitti_tripleD.py:19-20 (0 = no
    mask); the jitter ranges default to the reference's (mono_dataset.py:65-68).
    """

    def __init__(self, frame_ids, height, width, erase_count=0, erase_shape=(16, 16), brightness=(0.8, 1.2),
                 contrast=(0.8, 1.2), saturation=(0.8, 1.2), hue=(-0.1, 0.1), is_train=True):
        self.frame_ids = list(frame_ids)
        if not 1 <= len(self.frame_ids) <= _lib.TDL_MAX_SRC + 1:
            raise ValueError("1..5 frames per item")
        self.H, self.W = int(height), int(width)
        self.erase_count = int(erase_count)
        self.erase_shape = (int(erase_shape[0]), int(erase_shape[1]))
        self.brightness, self.contrast, self.saturation, self.hue = brightness, contrast, saturation, hue
        self.is_train = is_train

    # ---- host side: the random draws of __getitem__ / ColorJitter.get_params / preprocess_masks
    def sample_params(self, B, generator=None):
        """Returns host tensors: jitter (B,F,4) float32, order (B,F,4) int32, do_aug (B) uint8, do_flip (B) uint8,
        holes (B,count,2) int32 or None."""
        F = len(self.frame_ids)
        g = generator
        # one vectorised draw per quantity (a few hundred numbers per batch): the same distributions as the per-item
        # calls of the reference -- random.random() > 0.5 twice per item (mono_dataset.py:140-141), then per frame
        # ColorJitter.get_params: a random permutation of the four operations and one uniform factor each
        coin = torch.rand(B, 2, generator=g)
        do_aug = ((coin[:, 0] > 0.5) & bool(self.is_train)).to(torch.uint8)
        do_flip = ((coin[:, 1] > 0.5) & bool(self.is_train)).to(torch.uint8)
        order = torch.rand(B, F, 4, generator=g).argsort(dim=-1).to(torch.int32)          # uniform over the 24 orders
        u = torch.rand(B, F, 4, generator=g, dtype=torch.float64)
        lo = torch.tensor([self.brightness[0], self.contrast[0], self.saturation[0], self.hue[0]], dtype=torch.float64)
        hi = torch.tensor([self.brightness[1], self.contrast[1], self.saturation[1], self.hue[1]], dtype=torch.float64)
        val = lo + (hi - lo) * u
        jitter = val.to(torch.float32)
        # the byte torchvision adds to the hue channel: uint8(int32(hue_factor * 255)), truncation toward zero then wrap
        jitter[..., 3] = torch.remainder(torch.trunc(val[..., 3] * 255), 256).to(torch.float32)
        holes = None
        if self.erase_count > 0:
            eh, ew = self.erase_shape
            holes = torch.zeros(B, self.erase_count, 2, dtype=torch.int32)
            if self.erase_count == 1:                                      # kitti_dataset.py:171-174: one centred box
                off = (self.H - eh) // 2
                holes[:, 0, 0] = off
                holes[:, 0, 1] = off
            else:
                holes[..., 0] = torch.randint(0, self.H - eh - 1, (B, self.erase_count), generator=g)   # :177
                holes[..., 1] = torch.randint(0, self.W - ew - 1, (B, self.erase_count), generator=g)   # :178
        return dict(jitter=jitter, order=order, do_aug=do_aug, do_flip=do_flip, holes=holes)

    # ---- device side
    def __call__(self, frames_u8, params=None, want_color_aug=True, out=None):
        """frames_u8: {frame_id: (B,H,W,3) uint8 CUDA tensor}; params: sample_params() output (host or device tensors).
        Returns the `inputs` entries the data loader would have produced: ("color", f, 0), ("color_aug", f, 0) and,
        with erase_count > 0, ("mask", 0, 0)."""
        first = frames_u8[self.frame_ids[0]]
        if not first.is_cuda:
            raise _lib.TdlError("GpuInputPipeline needs CUDA tensors (there is no CPU path)")
        dev = first.device
        B = first.shape[0]
        if params is None:
            params = self.sample_params(B)
        P = {k: (v.to(dev, non_blocking=True) if v is not None else None) for k, v in params.items()}
        a = _lib.InputArgs()
        a.B, a.H, a.W, a.nframes = B, self.H, self.W, len(self.frame_ids)
        a.erase_count, a.erase_h, a.erase_w = (self.erase_count if P.get("holes") is not None else 0), *self.erase_shape
        keep = []
        res = {} if out is None else out
        for i, f in enumerate(self.frame_ids):
            fr = frames_u8[f]
            if fr.dtype != torch.uint8 or tuple(fr.shape) != (B, self.H, self.W, 3) or not fr.is_contiguous():
                raise ValueError(f"frame {f}: expected a contiguous uint8 tensor of shape {(B, self.H, self.W, 3)}")
            keep.append(fr)
            a.frames[i] = fr.data_ptr()
            c = torch.empty(B, 3, self.H, self.W, device=dev, dtype=torch.float32)
            res[("color", f, 0)] = c
            a.color[i] = c.data_ptr()
            if want_color_aug:
                ca = torch.empty_like(c)
                res[("color_aug", f, 0)] = ca
                a.color_aug[i] = ca.data_ptr()
        if want_color_aug:
            a.jitter, a.order, a.do_aug = P["jitter"].data_ptr(), P["order"].data_ptr(), P["do_aug"].data_ptr()
            ws = torch.empty(int(_lib.lib().tdl_input_ws_bytes(B, a.nframes)), device=dev, dtype=torch.uint8)
            a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
            keep.append(ws)
        if P.get("do_flip") is not None:
            a.do_flip = P["do_flip"].data_ptr()
        if a.erase_count > 0:
            m = torch.empty(B, 3, self.H, self.W, device=dev, dtype=torch.float32)
            res[("mask", 0, 0)] = m
            a.mask, a.holes = m.data_ptr(), P["holes"].data_ptr()
        _lib.check(_lib.lib().tdl_input_fwd(C.byref(a), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "tdl_input_fwd")
        if not torch.cuda.is_current_stream_capturing():         # (inside a graph everything lives on the capture stream)
            for t in keep + [v for v in P.values() if v is not None]:
                t.record_stream(torch.cuda.current_stream(dev))
        return res
