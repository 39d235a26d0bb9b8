"""Seeded synthetic inputs for the view-synthesis loss.

No dataset or checkpoint is reachable (no network), so tests and bench.py feed
the loss with synthetic frames of the shapes BASELINE.json names.  Everything is
generated on the CPU with an explicit ``torch.Generator`` so that the golden
fixtures, the CPU oracle and the CUDA path all see identical bits.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .geometry import transformation_from_parameters


def _box(x, k):
    p = k // 2
    return F.avg_pool2d(F.pad(x, (p, p, p, p), mode="reflect"), k, 1)


def kitti_intrinsics(batch, height, width):
    """mono/datasets/kitti_dataset.py:126-129 scaled as mono_dataset.py:174-177."""
    K = torch.tensor([[0.58 * width, 0, 0.5 * width, 0],
                      [0, 1.92 * height, 0.5 * height, 0],
                      [0, 0, 1, 0],
                      [0, 0, 0, 1]], dtype=torch.float32)
    inv_K = torch.linalg.pinv(K)
    return K.repeat(batch, 1, 1).contiguous(), inv_K.repeat(batch, 1, 1).contiguous()


def make_inputs(batch, height, width, frame_ids=(0, -1, 1), scales=(0, 1, 2, 3), seed=1234,
                frames="smooth", feat_channels=0, with_noise=True):
    """Returns (inputs, outputs, extras) dicts keyed like the reference's.

    inputs : ("color", f, 0) (B,3,H,W) in [0,1], "K", "inv_K" (B,4,4)
    outputs: ("disp", 0, s) (B,1,H/2^(s+1),W/2^(s+1)), ("cam_T_cam", 0, f) (B,4,4)
    extras : "noise"[s][f] (B,1,H,W) automask tie-break noise (consumption order of
             mono/model/mono_fm/net.py:94), "tgt_feat", "src_feats"[f] (B,C,H/2,W/2)
    """
    g = torch.Generator().manual_seed(seed)
    B, H, W = batch, height, width
    base = torch.rand(B, 3, H + 16, W + 16, generator=g)
    if frames == "smooth":
        base = _box(_box(base, 9), 9)
        lo = base.amin((2, 3), True)
        hi = base.amax((2, 3), True)
        base = (base - lo) / (hi - lo)
    elif frames != "white":
        raise ValueError(frames)
    inputs, outputs, extras = {}, {}, {}
    inputs[("color", 0, 0)] = base[:, :, 8:8 + H, 8:8 + W].contiguous()
    for f in frame_ids[1:]:
        # source = target shifted by a sub-pixel translation + 2% noise
        shift = (torch.rand(2, generator=g) * 4 - 2)
        theta = torch.tensor([[1.0, 0.0, 2 * shift[0] / (W + 16)],
                              [0.0, 1.0, 2 * shift[1] / (H + 16)]]).repeat(B, 1, 1)
        grid = F.affine_grid(theta, list(base.shape), align_corners=False)
        moved = F.grid_sample(base, grid, padding_mode="border", align_corners=False)
        src = moved[:, :, 8:8 + H, 8:8 + W] + 0.02 * torch.randn(B, 3, H, W, generator=g)
        inputs[("color", f, 0)] = src.clamp(0, 1).contiguous()
    inputs["K"], inputs["inv_K"] = kitti_intrinsics(B, H, W)
    for s in scales:
        h, w = H >> (s + 1), W >> (s + 1)
        d = torch.randn(B, 1, h, w, generator=g)
        d = _box(_box(d, 3), 3) * 2.0 if min(h, w) >= 3 else d
        outputs[("disp", 0, s)] = torch.sigmoid(d).contiguous()
    for f in frame_ids[1:]:
        aa = 0.01 * torch.randn(B, 1, 3, generator=g)
        tr = 0.01 * torch.randn(B, 1, 3, generator=g)
        outputs[("cam_T_cam", 0, f)] = transformation_from_parameters(aa, tr, invert=(f < 0)).contiguous()
    if with_noise:
        extras["noise"] = {s: {f: torch.randn(B, 1, H, W, generator=g) for f in frame_ids[1:]}
                           for s in scales}
    if feat_channels:
        C, h, w = feat_channels, H // 2, W // 2

        def feat():
            return _box(torch.relu(torch.randn(B, C, h, w, generator=g)), 3).contiguous()
        tgt = feat()
        extras["tgt_feat"] = tgt
        extras["src_feats"] = {f: (tgt + 0.1 * feat()).contiguous() for f in frame_ids[1:]}
    return inputs, outputs, extras
