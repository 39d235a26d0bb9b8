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


def stereo_transform(batch, sign=1.0):
    """inputs["stereo_T"] of the 's' frame: identity rotation, 0.1 baseline along x (mono/datasets/mono_dataset.py:194-199)."""
    T = torch.eye(4, dtype=torch.float32).repeat(batch, 1, 1)
    T[:, 0, 3] = sign * 0.1
    return T.contiguous()


def _add_poses(g, B, frame_ids, inputs, outputs):
    for f in frame_ids[1:]:
        if f == "s":
            inputs["stereo_T"] = stereo_transform(B)
            continue
        aa = 0.01 * torch.randn(B, 1, 3, generator=g)
        tr = 0.01 * torch.randn(B, 1, 3, generator=g)
        outputs[("cam_T_cam", 0, f)] = transformation_from_parameters(aa, tr, invert=(f < 0)).contiguous()


def _unit(x):
    lo, hi = x.amin((2, 3), True), x.amax((2, 3), True)
    return (x - lo) / (hi - lo)


def make_inputs(batch, height, width, frame_ids=(0, -1, 1), scales=(0, 1, 2, 3), seed=1234,
                frames="smooth", feat_channels=0, with_noise=True):
    """Returns (inputs, outputs, extras) dicts keyed like the reference's.

    frames : "smooth" (sources = target shifted by a global sub-pixel translation, pixel-level random
             disparity -- the un-warped sources win almost every arg-min, i.e. a fully auto-masked static scene),
             "white" (stress), "waves" (band-limited: the strict gradient-parity fixture), "scene" (a moving camera in a
             smooth 3-D scene: sources RENDERED from the target through ground-truth depth + pose, prediction = ground
             truth + a small error -- the warped sources win most pixels and the flow is coherent, like a training step
             on real video; see _make_scene_inputs)
    inputs : ("color", f, 0) (B,3,H,W) in [0,1], "K", "inv_K" (B,4,4)
    outputs: ("disp", 0, s) (B,1,H/2^(s+1),W/2^(s+1)), ("cam_T_cam", 0, f) (B,4,4)
    extras : "noise"[s][f] (B,1,H,W) automask tie-break noise (consumption order of
             mono/model/mono_fm/net.py:94), "tgt_feat", "src_feats"[f] (B,C,H/2,W/2)
    """
    g = torch.Generator().manual_seed(seed)
    B, H, W = batch, height, width
    if frames == "waves":
        return _make_wave_inputs(g, B, H, W, frame_ids, scales, feat_channels, with_noise)
    if frames == "scene":
        return _make_scene_inputs(g, B, H, W, frame_ids, scales, feat_channels, with_noise)
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
    _add_poses(g, B, frame_ids, inputs, outputs)
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
    _add_poses(g, B, frame_ids, inputs, outputs)
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


# ------------------------------------------------------------------------------------------------ "scene"
def _sample_px(field, px, py):
    """Bilinear sample of field (B,C,H,W) at fractional pixel-index coordinates px, py (B,H',W'); border clamp."""
    H, W = field.shape[-2:]
    grid = torch.stack([2 * px / (W - 1) - 1, 2 * py / (H - 1) - 1], -1)
    return F.grid_sample(field, grid.to(field.dtype), mode="bilinear", padding_mode="border", align_corners=True)


def _source_index_map(disp, K, inv_K, T, min_depth=0.1, max_depth=100.0):
    """Where the loss samples the source frame for every target pixel: the composition of disp_to_depth, Backproject,
    Project and F.grid_sample's align_corners=False un-normalisation (mono/model/mono_fm/layers.py:57-82,
    net.py:135-140,169) -> (ix, iy) in source pixel-index units, un-clipped, in float64."""
    B, _, H, W = disp.shape
    disp, K, inv_K, T = disp.double(), K.double(), inv_K.double(), T.double()
    depth = 1.0 / (1.0 / max_depth + (1.0 / min_depth - 1.0 / max_depth) * disp)
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(H * W, dtype=torch.float64)], 0)
    cam = depth.view(B, 1, -1) * torch.matmul(inv_K[:, :3, :3], pix)
    cam = torch.cat([cam, torch.ones(B, 1, H * W, dtype=torch.float64)], 1)
    p = torch.matmul(torch.matmul(K, T)[:, :3, :], cam)
    u = (p[:, 0] / (p[:, 2] + 1e-7)).view(B, H, W)
    v = (p[:, 1] / (p[:, 2] + 1e-7)).view(B, H, W)
    return u / (W - 1) * W - 0.5, v / (H - 1) * H - 0.5


def _render_source(target, ix, iy, iters=4):
    """The source frame that the map (ix, iy) warps back onto `target`: src[q] = target[p] with M(p) = q, solved by the
    fixed point p <- q - (M(p) - p) (the flow is smooth and a few pixels long, so this converges in a few rounds)."""
    B, _, H, W = target.shape
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    qx, qy = xs.expand(B, H, W), ys.expand(B, H, W)
    flow = torch.stack([ix - qx, iy - qy], 1)                       # u(p) on the integer target grid
    px, py = qx.clone(), qy.clone()
    for _ in range(iters):
        u = _sample_px(flow, px, py)
        px, py = qx - u[:, 0], qy - u[:, 1]
    return _sample_px(target.double(), px, py)


def _make_scene_inputs(g, B, H, W, frame_ids, scales, feat_channels, with_noise):
    """frames="scene": the representative training-step workload.

    * ground truth: a smooth disparity field in [0.02, 0.15] (depth 0.65 .. 4.8, nearer towards the bottom of the image
      like a road scene) and one camera motion per source frame -- mostly along the optical axis, a few 1e-2 units, plus a
      small rotation -- which gives a coherent flow of up to ~10 pixels;
    * frames: textured target (box-filtered noise, as "smooth"); every source frame is RENDERED from the target through
      that geometry with the loss's own index map (so warping it back with the true depth and pose reproduces the target
      up to one bilinear blur), then perturbed by a gain change and 2 % low-passed noise;
    * prediction handed to the loss: ground-truth disparity area-averaged to each decoder resolution + 2 % smooth error, the
      true pose + 2 % error.  The warped sources therefore win the per-pixel minimum almost everywhere (identity channels
      keep the far / low-parallax pixels and the one static image in eight: real auto-masking), and neighbouring pixels
      sample neighbouring source pixels;
    * features: 64-channel stand-ins rendered the same way at (H/2, W/2) through the half-resolution geometry.
    """
    inputs, outputs, extras = {}, {}, {}
    base = _box(_box(torch.rand(B, 3, H + 16, W + 16, generator=g), 9), 9)
    base = _unit(base)[:, :, 8:8 + H, 8:8 + W]
    fine = _box(_box(torch.rand(B, 3, H, W, generator=g), 3), 3) * 1.5   # finer texture: keeps SSIM informative at 1-pixel shifts
    target = (0.8 * base + 0.2 * fine).clamp(0, 1).contiguous()
    inputs[("color", 0, 0)] = target
    K, inv_K = kitti_intrinsics(B, H, W)
    inputs["K"], inputs["inv_K"] = K, inv_K
    # ground-truth disparity, full resolution
    lo = torch.randn(B, 1, max(H // 16, 3), max(W // 16, 3), generator=g)
    field = F.interpolate(lo, (H, W), mode="bicubic", align_corners=False)
    ramp = torch.linspace(-1.0, 1.0, H).view(1, 1, H, 1)
    disp_gt = (0.02 + 0.13 * torch.sigmoid(0.8 * field + 1.2 * ramp)).contiguous()
    extras["disp_gt"] = disp_gt
    for s in scales:
        h, w = H >> (s + 1), W >> (s + 1)
        d = F.interpolate(disp_gt, (h, w), mode="area")
        err = torch.randn(B, 1, h, w, generator=g)
        err = _box(_box(err, 3), 3) if min(h, w) >= 3 else err
        outputs[("disp", 0, s)] = (d * (1.0 + 0.02 * err)).clamp(1e-3, 1.0).contiguous()
    extras["T_gt"] = {}
    # every eighth image of a batch is a (nearly) static camera -- the car waiting at a light -- which is the case
    # auto-masking exists for: there the un-warped sources win and the backward sees almost no live pixels
    motion = torch.where(torch.arange(B) % 8 == 7, 0.02, 1.0).view(B, 1, 1)
    for f in frame_ids[1:]:
        if f == "s":
            raise ValueError('frames="scene" renders temporal neighbours only')
        sgn = -1.0 if f < 0 else 1.0
        aa = 0.004 * torch.randn(B, 1, 3, generator=g) * motion
        tr = (torch.randn(B, 1, 3, generator=g) * torch.tensor([0.01, 0.004, 0.01]) + torch.tensor([0.0, 0.0, 0.03 * sgn])) * motion
        T_gt = transformation_from_parameters(aa, tr, invert=False)
        extras["T_gt"][f] = T_gt
        ix, iy = _source_index_map(disp_gt, K, inv_K, T_gt)
        src = _render_source(target, ix, iy).float()
        pert = _box(_box(torch.randn(B, 3, H, W, generator=g), 9), 9)
        pert = pert / pert.std((1, 2, 3), keepdim=True)
        gain = 0.95 + 0.1 * torch.rand(B, 1, 1, 1, generator=g)
        inputs[("color", f, 0)] = (gain * src + 0.02 * pert).clamp(0, 1).contiguous()
        # predicted pose: the truth + 2 % error
        aa_p = aa * (1 + 0.02 * torch.randn(B, 1, 3, generator=g))
        tr_p = tr * (1 + 0.02 * torch.randn(B, 1, 3, generator=g))
        outputs[("cam_T_cam", 0, f)] = transformation_from_parameters(aa_p, tr_p, invert=False).contiguous()
    if with_noise:
        extras["noise"] = {s: {f: torch.randn(B, 1, H, W, generator=g) for f in frame_ids[1:]} for s in scales}
    if feat_channels:
        C, h, w = feat_channels, H // 2, W // 2
        t = _box(_box(torch.relu(torch.randn(B, C, h, w, generator=g)), 5), 5)
        t = (t / t.std() * 0.5).contiguous()
        extras["tgt_feat"] = t
        Kh = K.clone()
        Kh[:, 0:2, :] *= 0.5
        inv_Kh = inv_K.clone()
        inv_Kh[:, :, 0:2] *= 2.0
        disp_h = F.interpolate(disp_gt, (h, w), mode="area")
        extras["src_feats"] = {}
        for f in frame_ids[1:]:
            ix, iy = _source_index_map(disp_h, Kh, inv_Kh, extras["T_gt"][f])
            sf = _render_source(t, ix, iy).float()
            p = _box(_box(torch.relu(torch.randn(B, C, h, w, generator=g)), 5), 5)
            extras["src_feats"][f] = (sf + 0.05 * p / p.std() * 0.5).contiguous()
    return inputs, outputs, extras
