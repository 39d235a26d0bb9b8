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


def _waves(g, B, C, H, W, dx=0.0, dy=0.0, terms=3, max_cycles=2.0, coef=None):
    """Band-limited maps: sum of a few sinusoids with at most ``max_cycles`` periods across the
    image, evaluated at pixel centres shifted by (dx, dy).  Their bilinear interpolant has a nearly
    continuous derivative (second difference / first difference ~ 2*pi*cycles/W), which removes the
    one-pixel gradient discontinuities that fp32 rounding otherwise decides (tests/test_gpu_parity.py)."""
    if coef is None:
        coef = dict(k=(torch.rand(B, C, terms, 2, generator=g) * 2 - 1) * max_cycles,
                    ph=torch.rand(B, C, terms, generator=g) * 6.283185307179586,
                    a=torch.rand(B, C, terms, generator=g) + 0.5)
    ys = (torch.arange(H, dtype=torch.float64) + dy) / H
    xs = (torch.arange(W, dtype=torch.float64) + dx) / W
    out = torch.zeros(B, C, H, W, dtype=torch.float64)
    for t in range(terms):
        kx = coef["k"][:, :, t, 0].double()[..., None, None]
        ky = coef["k"][:, :, t, 1].double()[..., None, None]
        ph = coef["ph"][:, :, t].double()[..., None, None]
        a = coef["a"][:, :, t].double()[..., None, None]
        out += a * torch.sin(6.283185307179586 * (kx * xs[None, None, None, :] + ky * ys[None, None, :, None]) + ph)
    return out, coef


def _unit(x):
    lo, hi = x.amin((2, 3), True), x.amax((2, 3), True)
    return (x - lo) / (hi - lo)


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
    if frames == "waves":
        return _make_wave_inputs(g, B, H, W, frame_ids, scales, feat_channels, with_noise)
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
        pert = torch.randn(B, 3, H, W, generator=g)
        if frames == "smooth":
            # 2% low-passed perturbation: keeps warped / identity errors apart without making the bilinear
            # derivative jump at integer source coordinates (a white perturbation turns every such
            # crossing into a one-pixel discontinuity of the gradient, which fp32 rounding then decides)
            pert = _box(_box(pert, 9), 9)
            pert = pert / pert.std((1, 2, 3), keepdim=True)
        src = moved[:, :, 8:8 + H, 8:8 + W] + 0.02 * pert
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
            # stand-in for relu(bn1(conv1(img))) (mono/model/mono_autoencoder/encoder.py:37): non-negative,
            # low-passed so that its bilinear interpolant has a (nearly) continuous derivative
            f = _box(_box(torch.relu(torch.randn(B, C, h, w, generator=g)), 5), 5)
            return (f / f.std() * 0.5).contiguous()
        tgt = feat()
        extras["tgt_feat"] = tgt
        extras["src_feats"] = {f: (tgt + 0.1 * feat()).contiguous() for f in frame_ids[1:]}
    return inputs, outputs, extras


def _make_wave_inputs(g, B, H, W, frame_ids, scales, feat_channels, with_noise):
    """frames="waves": every map is band-limited (see _waves) -- the strict gradient-parity fixture."""
    inputs, outputs, extras = {}, {}, {}
    tgt, coef = _waves(g, B, 3, H, W)
    lo, hi = tgt.amin((2, 3), True), tgt.amax((2, 3), True)
    norm = lambda x: (0.1 + 0.8 * (x - lo) / (hi - lo))
    inputs[("color", 0, 0)] = norm(tgt).float().contiguous()
    for f in frame_ids[1:]:
        shift = (torch.rand(2, generator=g) * 4 - 2)
        moved, _ = _waves(g, B, 3, H, W, dx=float(shift[0]), dy=float(shift[1]), coef=coef)
        pert, _ = _waves(g, B, 3, H, W)
        # gain / offset change + 3% band-limited perturbation: a photometric error of a few 1e-2, so that
        # the loss is not dominated by the fp32 noise floor of SSIM ~ 0 (tests/test_gpu_parity.py)
        gain = 0.8 + 0.4 * float(torch.rand(1, generator=g))
        src = gain * norm(moved) + (1 - gain) * 0.4 + 0.03 * pert / pert.std()
        inputs[("color", f, 0)] = src.clamp(0, 1).float().contiguous()
    inputs["K"], inputs["inv_K"] = kitti_intrinsics(B, H, W)
    for s in scales:
        h, w = H >> (s + 1), W >> (s + 1)
        d, _ = _waves(g, B, 1, h, w)
        # + mild texture: a purely band-limited disparity has second differences ~1e-8 over wide
        # regions, where sign() in the smoothness gradient is decided by rounding
        tex = torch.randn(B, 1, h, w, generator=g)
        tex = _box(_box(tex, 3), 3) if min(h, w) >= 3 else tex
        outputs[("disp", 0, s)] = torch.sigmoid(d.float() + 0.3 * tex).contiguous()
    for f in frame_ids[1:]:
        aa = 0.01 * torch.randn(B, 1, 3, generator=g)
        tr = 0.01 * torch.randn(B, 1, 3, generator=g)
        outputs[("cam_T_cam", 0, f)] = transformation_from_parameters(aa, tr, invert=(f < 0)).contiguous()
    if with_noise:
        extras["noise"] = {s: {f: torch.randn(B, 1, H, W, generator=g) for f in frame_ids[1:]} for s in scales}
    if feat_channels:
        C, h, w = feat_channels, H // 2, W // 2
        tf, fcoef = _waves(g, B, C, h, w)
        extras["tgt_feat"] = (0.5 * _unit(tf)).float().contiguous()
        extras["src_feats"] = {}
        for f in frame_ids[1:]:
            shift = (torch.rand(2, generator=g) * 2 - 1)
            mf, _ = _waves(g, B, C, h, w, dx=float(shift[0]), dy=float(shift[1]), coef=fcoef)
            pf, _ = _waves(g, B, C, h, w)
            lo, hi = tf.amin((2, 3), True), tf.amax((2, 3), True)
            extras["src_feats"][f] = (0.5 * (mf - lo) / (hi - lo) + 0.02 * pf / pf.std()).float().contiguous()
    return inputs, outputs, extras
