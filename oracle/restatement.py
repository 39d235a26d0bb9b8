"""Plain-PyTorch CPU restatement of the reference's view-synthesis loss path.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Everything here runs on
the CPU in fp32 (or fp64 for finite-difference style checks) with stock torch
ops, in the same operation order as the reference, so that on the same seeded
inputs it is bit-identical to the reference executed by the same torch build
(``tests/test_oracle_vs_reference.py`` checks that whenever the reference is reachable --
``/root/reference`` in the build container, ``oracle/_ref`` on the GPU box; ``tests/golden/*.pt``
pins it everywhere else).

File:line citations are relative to ``/root/reference``.

The restatement is functional: the learnable networks of the reference (depth
decoder, pose decoder, feature extractor) are *outside* this path, so their
outputs (``disp_s``, ``cam_T_cam``, feature maps) enter as plain tensors.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

SSIM_C1 = 0.01 ** 2     # mono/model/mono_fm/layers.py:94
SSIM_C2 = 0.03 ** 2     # mono/model/mono_fm/layers.py:95
L1_EPS = 1e-3           # mono/model/mono_fm/net.py:56


@dataclass
class LossSpec:
    """The hot-path options the reference reads from ``self.opt``
    (config/cfg_kitti_fm.py:20-38)."""
    height: int
    width: int
    frame_ids: Sequence[int] = (0, -1, 1)
    scales: Sequence[int] = (0, 1, 2, 3)
    min_depth: float = 0.1
    max_depth: float = 100.0
    automask: bool = True
    disp_norm: bool = True
    smoothness_weight: float = 1e-3
    perception_weight: float = 1e-3
    align_corners: bool = False       # torch>=1.3 default of F.grid_sample
    extra: dict = field(default_factory=dict)


# --------------------------------------------------------------------------- geometry
def disp_to_depth(disp, min_depth, max_depth):
    """mono/model/mono_fm/net.py:135-140 (same as layers.py:33-38)."""
    min_disp = 1 / max_depth
    max_disp = 1 / min_depth
    scaled = min_disp + (max_disp - min_disp) * disp
    return scaled, 1 / scaled


def pixel_grid(batch, height, width, dtype=torch.float32, device=None):
    """Homogeneous pixel coordinates (B,3,H*W): rows x, y, 1.
    mono/model/mono_fm/layers.py:49-55 (the reference builds them on the host and calls .cuda() per use;
    ``device`` lets bench.py time this same op sequence as eager PyTorch on the GPU)."""
    ys, xs = torch.meshgrid(torch.arange(height, dtype=dtype, device=device),
                            torch.arange(width, dtype=dtype, device=device), indexing="ij")
    pix = torch.stack([xs.reshape(-1), ys.reshape(-1), torch.ones(height * width, dtype=dtype, device=device)], 0)
    return pix.unsqueeze(0).repeat(batch, 1, 1)


def backproject(depth, inv_K):
    """mono/model/mono_fm/layers.py:57-61."""
    b, _, h, w = depth.shape
    rays = torch.matmul(inv_K[:, :3, :3], pixel_grid(b, h, w, depth.dtype, depth.device))
    cam = depth.view(b, 1, -1) * rays
    return torch.cat([cam, torch.ones(b, 1, h * w, dtype=depth.dtype, device=depth.device)], 1)


def project(cam_points, K, T, height, width, eps=1e-7):
    """mono/model/mono_fm/layers.py:73-82 -> sampling grid (B,H,W,2)."""
    b = cam_points.shape[0]
    P = torch.matmul(K, T)[:, :3, :]
    p = torch.matmul(P, cam_points)
    uv = p[:, :2, :] / (p[:, 2, :].unsqueeze(1) + eps)
    uv = uv.view(b, 2, height, width).permute(0, 2, 3, 1)
    uv[..., 0] /= width - 1
    uv[..., 1] /= height - 1
    return (uv - 0.5) * 2


def sampling_grid(disp_lr, K, inv_K, T, out_hw, spec: LossSpec):
    """Upsample -> depth -> backproject -> project.
    mono/model/mono_fm/net.py:158-168."""
    h, w = out_hw
    disp = F.interpolate(disp_lr, [h, w], mode="bilinear", align_corners=False)
    _, depth = disp_to_depth(disp, spec.min_depth, spec.max_depth)
    return project(backproject(depth, inv_K), K, T, h, w)


def warp(img, grid, spec: LossSpec):
    """mono/model/mono_fm/net.py:169 (grid_sample, bilinear, border)."""
    return F.grid_sample(img, grid, mode="bilinear", padding_mode="border",
                         align_corners=spec.align_corners)


def half_res_intrinsics(K):
    """K with rows 0,1 halved and its per-sample pseudo-inverse.
    mono/model/mono_fm/net.py:185-191."""
    Kh = K.clone()
    Kh[:, 0, :] /= 2
    Kh[:, 1, :] /= 2
    inv = torch.zeros_like(Kh)
    for i in range(inv.shape[0]):
        inv[i] = torch.pinverse(Kh[i])
    return Kh, inv


# --------------------------------------------------------------------------- photometric
def ssim(x, y):
    """mono/model/mono_fm/layers.py:97-107; x = prediction, y = target."""
    x = F.pad(x, (1, 1, 1, 1), mode="reflect")
    y = F.pad(y, (1, 1, 1, 1), mode="reflect")
    mu_x = F.avg_pool2d(x, 3, 1)
    mu_y = F.avg_pool2d(y, 3, 1)
    sigma_x = F.avg_pool2d(x ** 2, 3, 1) - mu_x ** 2
    sigma_y = F.avg_pool2d(y ** 2, 3, 1) - mu_y ** 2
    sigma_xy = F.avg_pool2d(x * y, 3, 1) - mu_x * mu_y
    n = (2 * mu_x * mu_y + SSIM_C1) * (2 * sigma_xy + SSIM_C2)
    d = (mu_x ** 2 + mu_y ** 2 + SSIM_C1) * (sigma_x + sigma_y + SSIM_C2)
    return torch.clamp((1 - n / d) / 2, 0, 1)


def robust_l1(pred, target):
    """mono/model/mono_fm/net.py:55-57."""
    return torch.sqrt(torch.pow(target - pred, 2) + L1_EPS ** 2)


def reprojection_loss(pred, target):
    """mono/model/mono_fm/net.py:63-67 -> (B,1,H,W)."""
    l1 = robust_l1(pred, target).mean(1, True)
    s = ssim(pred, target).mean(1, True)
    return 0.85 * s + 0.15 * l1


def perceptional_loss(tgt_f, src_f):
    """mono/model/mono_fm/net.py:59-61 -> (B,1,h,w)."""
    return robust_l1(tgt_f, src_f).mean(1, True)


# --------------------------------------------------------------------------- smoothness
def _grad(t):
    """mono/model/mono_fm/net.py:280-283 -> (d/dx, d/dy) first differences."""
    return t[:, :, :, 1:] - t[:, :, :, :-1], t[:, :, 1:] - t[:, :, :-1]


def smooth_loss(disp, img, a1=0.5, a2=0.5):
    """mono/model/mono_fm/net.py:255-278 (edge-aware 1st + 2nd order)."""
    _, _, h, w = disp.shape
    img = F.interpolate(img, (h, w), mode="area")
    d_dx, d_dy = _grad(disp)
    i_dx, i_dy = _grad(img)
    d_dxx, d_dxy = _grad(d_dx)
    d_dyx, d_dyy = _grad(d_dy)
    i_dxx, i_dxy = _grad(i_dx)
    i_dyx, i_dyy = _grad(i_dy)

    def term(dd, ii, a):
        return torch.mean(dd.abs() * torch.exp(-a * ii.abs().mean(1, True)))

    first = term(d_dx, i_dx, a1) + term(d_dy, i_dy, a1)
    second = (term(d_dxx, i_dxx, a2) + term(d_dxy, i_dxy, a2)
              + term(d_dyx, i_dyx, a2) + term(d_dyy, i_dyy, a2))
    return first + second


def normalise_disp(disp):
    """mono/model/mono_fm/net.py:123-125."""
    m = disp.mean(2, True).mean(3, True)
    return disp / (m + 1e-7)


# --------------------------------------------------------------------------- full losses
def draw_automask_noise(spec: LossSpec, batch, generator=None, dtype=torch.float32):
    """The reference draws ``torch.randn(B,1,H,W)`` from the global CPU
    generator per scale, per source frame (mono/model/mono_fm/net.py:94).
    Returns noise[scale][frame_id] in the same consumption order."""
    out = {}
    for s in spec.scales:
        out[s] = {}
        for f in spec.frame_ids[1:]:
            if generator is None:
                out[s][f] = torch.randn(batch, 1, spec.height, spec.width).to(dtype)
            else:
                out[s][f] = torch.randn(batch, 1, spec.height, spec.width, generator=generator).to(dtype)
    return out


def images_pred(spec: LossSpec, inputs, outputs, scale):
    """mono/model/mono_fm/net.py:157-170: adds ("color", f, scale)."""
    for f in spec.frame_ids[1:]:
        T = inputs["stereo_T"] if f == "s" else outputs[("cam_T_cam", 0, f)]
        grid = sampling_grid(outputs[("disp", 0, scale)], inputs["K"], inputs["inv_K"], T,
                             (spec.height, spec.width), spec)
        outputs[("color", f, scale)] = warp(inputs[("color", f, 0)], grid, spec)
    return outputs


def features_pred(spec: LossSpec, inputs, outputs, src_feats, inv_K_half=None):
    """mono/model/mono_fm/net.py:172-199: adds ("feature", f, 0).
    ``src_feats[f]`` stands in for ``extractor(color_f)[0]``."""
    h, w = int(spec.height / 2), int(spec.width / 2)
    Kh, inv = half_res_intrinsics(inputs["K"])
    if inv_K_half is not None:
        inv = inv_K_half
    for f in spec.frame_ids[1:]:
        T = inputs["stereo_T"] if f == "s" else outputs[("cam_T_cam", 0, f)]
        grid = sampling_grid(outputs[("disp", 0, 0)], Kh, inv, T, (h, w), spec)
        outputs[("feature", f, 0)] = warp(src_feats[f], grid, spec)
    return outputs


def _select_min(stacked, forced_index):
    """torch.min(dim=1) as the reference does, or -- for tie analysis in the tests -- the
    same reduction with the selected channel imposed (gather), which has the identical
    backward (gradient to the selected channel only)."""
    if forced_index is None:
        return torch.min(stacked, dim=1)
    idx = forced_index.to(torch.int64)
    return stacked.gather(1, idx.unsqueeze(1)).squeeze(1), idx


def photometric_scale(spec: LossSpec, inputs, outputs, scale, noise, forced_index=None):
    """Automask + minimum reprojection for one scale.
    mono/model/mono_fm/net.py:90-106 -> (loss scalar, min_index int64 (B,H,W))."""
    target = inputs[("color", 0, 0)]
    chans = []
    if spec.automask:
        for f in spec.frame_ids[1:]:
            ident = reprojection_loss(inputs[("color", f, 0)], target)
            ident = ident + noise[scale][f] * 1e-5
            chans.append(ident)
    for f in spec.frame_ids[1:]:
        chans.append(reprojection_loss(outputs[("color", f, scale)], target))
    stacked = torch.cat(chans, 1)
    outputs[("reproj_stack", scale)] = stacked.detach()       # test-only: lets the tests measure near-ties
    m, idx = _select_min(stacked, forced_index)
    return m.mean() / len(spec.scales), idx


def smooth_scale(spec: LossSpec, inputs, outputs, scale, weight):
    """mono/model/mono_fm/net.py:120-131."""
    disp = outputs[("disp", 0, scale)]
    if spec.disp_norm:
        disp = normalise_disp(disp)
    return weight * smooth_loss(disp, inputs[("color", 0, 0)]) / (2 ** scale) / len(spec.scales)


def compute_losses_baseline(spec: LossSpec, inputs, outputs, noise, forced=None):
    """mono/model/mono_baseline/net.py:51-100 (photometric + automask + smoothness).
    ``forced``: optional {("photo", scale): index map} imposing the arg-min (tests only)."""
    loss = {}
    forced = forced or {}
    for s in spec.scales:
        images_pred(spec, inputs, outputs, s)
        loss[("min_reconstruct_loss", s)], outputs[("min_index", s)] = \
            photometric_scale(spec, inputs, outputs, s, noise, forced.get(("photo", s)))
        outputs[("min_index_photo", s)] = outputs[("min_index", s)]
        loss[("smooth_loss", s)] = smooth_scale(spec, inputs, outputs, s, spec.smoothness_weight)
    return loss


def compute_losses_fm(spec: LossSpec, inputs, outputs, noise, tgt_feat, src_feats, inv_K_half=None,
                      forced=None):
    """mono/model/mono_fm/net.py:69-133 (adds the feature-metric term, evaluated
    once per scale exactly like the reference does)."""
    loss = {}
    forced = forced or {}
    for s in spec.scales:
        images_pred(spec, inputs, outputs, s)
        features_pred(spec, inputs, outputs, src_feats, inv_K_half)
        loss[("min_reconstruct_loss", s)], outputs[("min_index_photo", s)] = \
            photometric_scale(spec, inputs, outputs, s, noise, forced.get(("photo", s)))
        per = torch.cat([perceptional_loss(tgt_feat, outputs[("feature", f, 0)])
                         for f in spec.frame_ids[1:]], 1)
        outputs["feat_stack"] = per.detach()
        m, outputs[("min_index", s)] = _select_min(per, forced.get("feat"))
        loss[("min_perceptional_loss", s)] = spec.perception_weight * m.mean() / len(spec.scales)
        loss[("smooth_loss", s)] = smooth_scale(spec, inputs, outputs, s, spec.smoothness_weight)
    return loss


def compute_losses_inpaint_core(spec: LossSpec, inputs, outputs, noise, tgt_feat, src_feats,
                                inv_K_half=None, forced=None):
    """View-synthesis part of mono/model/mono_fm_joint_inpaint/net.py:47-133:
    one un-scaled feature-metric term (:58-70) + per-scale photometric/smooth
    (:96-131).  The autoencoder reconstruction / regularisation terms of that
    method are outside the hot path."""
    loss = {}
    forced = forced or {}
    features_pred(spec, inputs, outputs, src_feats, inv_K_half)
    per = torch.cat([perceptional_loss(tgt_feat, outputs[("feature", f, 0)])
                     for f in spec.frame_ids[1:]], 1)
    outputs["feat_stack"] = per.detach()
    m, outputs["min_index"] = _select_min(per, forced.get("feat"))
    loss["min_perceptional_loss"] = spec.perception_weight * m.mean()
    for s in spec.scales:
        images_pred(spec, inputs, outputs, s)
        loss[("min_reconstruct_loss", s)], outputs[("min_index", s)] = \
            photometric_scale(spec, inputs, outputs, s, noise, forced.get(("photo", s)))
        outputs[("min_index_photo", s)] = outputs[("min_index", s)]
        loss[("smooth_loss", s)] = smooth_scale(spec, inputs, outputs, s, spec.smoothness_weight)
    return loss


def compute_losses_joint(spec: LossSpec, inputs, outputs, noise, features, src_feats, inv_K_half=None, forced=None):
    """mono/model/mono_fm_joint/net.py:73-155: feature regularisation on the five encoder levels (:77-80), per scale the
    UN-masked autoencoder reconstruction term (:96-101), photometric + automask min (:103-131), the feature-metric
    term re-evaluated per scale like mono_fm (:133-143) and the smoothness term (:145-153)."""
    loss = {}
    forced = forced or {}
    target = inputs[("color", 0, 0)]
    for i in range(5):
        loss[("feature_regularization_loss", i)] = feature_regularization_loss(
            features[i], target, spec.extra["dis"], spec.extra["cvt"]) / (2 ** i) / 5
    for s in spec.scales:
        res_img = outputs[("res_img", 0, s)]
        _, _, h, w = res_img.size()
        target_resize = F.interpolate(target, [h, w], mode="bilinear", align_corners=False)
        loss[("img_reconstruct_loss", s)] = reprojection_loss(res_img, target_resize).mean() / len(spec.scales)
        images_pred(spec, inputs, outputs, s)
        features_pred(spec, inputs, outputs, src_feats, inv_K_half)
        loss[("min_reconstruct_loss", s)], outputs[("min_index_photo", s)] = \
            photometric_scale(spec, inputs, outputs, s, noise, forced.get(("photo", s)))
        per = torch.cat([perceptional_loss(features[0], outputs[("feature", f, 0)])
                         for f in spec.frame_ids[1:]], 1)
        outputs["feat_stack"] = per.detach()
        m, outputs[("min_index", s)] = _select_min(per, forced.get("feat"))
        loss[("min_perceptional_loss", s)] = spec.perception_weight * m.mean() / len(spec.scales)
        loss[("smooth_loss", s)] = smooth_scale(spec, inputs, outputs, s, spec.smoothness_weight)
    return loss


def img_reconstruct_loss(spec: LossSpec, inputs, outputs, scale, weight=1):
    """mono/model/mono_fm_joint_inpaint/net.py:80-91: SSIM + L1 of the autoencoder output against the
    bilinearly resized target, averaged over the erased (mask == 0) region."""
    target, mask = inputs[("color", 0, 0)], inputs[("mask", 0, 0)]
    res_img = outputs[("res_img", 0, scale)]
    _, _, h, w = res_img.size()
    target_resize = F.interpolate(target, [h, w], mode="bilinear", align_corners=False)
    mask_resize = F.interpolate(mask, [h, w], mode="bilinear", align_corners=False)
    loss = reprojection_loss(res_img, target_resize)
    loss = torch.sum(loss * (1 - mask_resize)) / torch.sum(1 - mask_resize)
    return loss / len(spec.scales) * weight


def auto_res_loss(inputs, outputs, weight):
    """mono/model/mono_fm_joint_inpaint/net.py:520-527 (un-reduced (B,1,H,W) map, as in the reference)."""
    return perceptional_loss(inputs[("color", 0, 0)], outputs[("auto_res_img", 0, 0)]) * weight


def feature_regularization_loss(feature, img, dis, cvt):
    """mono/model/mono_fm_joint/net.py:309-330 (exponent coefficient 1,
    ``-dis * first + cvt * second``)."""
    _, _, h, w = feature.shape
    img = F.interpolate(img, (h, w), mode="area")
    f_dx, f_dy = _grad(feature)
    i_dx, i_dy = _grad(img)
    f_dxx, f_dxy = _grad(f_dx)
    f_dyx, f_dyy = _grad(f_dy)
    i_dxx, i_dxy = _grad(i_dx)
    i_dyx, i_dyy = _grad(i_dy)

    def term(dd, ii):
        return torch.mean(dd.abs() * torch.exp(-ii.abs().mean(1, True)))

    first = term(f_dx, i_dx) + term(f_dy, i_dy)
    second = term(f_dxx, i_dxx) + term(f_dxy, i_dxy) + term(f_dyx, i_dyx) + term(f_dyy, i_dyy)
    return -dis * first + cvt * second


# --------------------------------------------------------------------------- pose helpers
def rot_from_axisangle(vec):
    """mono/model/mono_fm/net.py:225-253; vec (B,1,3) -> (B,4,4)."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = (axis[..., i].unsqueeze(1) for i in range(3))
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    rot = torch.zeros((vec.shape[0], 4, 4), dtype=vec.dtype)
    rot[:, 0, 0] = torch.squeeze(x * xC + ca)
    rot[:, 0, 1] = torch.squeeze(xyC - zs)
    rot[:, 0, 2] = torch.squeeze(zxC + ys)
    rot[:, 1, 0] = torch.squeeze(xyC + zs)
    rot[:, 1, 1] = torch.squeeze(y * yC + ca)
    rot[:, 1, 2] = torch.squeeze(yzC - xs)
    rot[:, 2, 0] = torch.squeeze(zxC - ys)
    rot[:, 2, 1] = torch.squeeze(yzC + xs)
    rot[:, 2, 2] = torch.squeeze(z * zC + ca)
    rot[:, 3, 3] = 1
    return rot


def transformation_from_parameters(axisangle, translation, invert=False):
    """mono/model/mono_fm/net.py:201-223."""
    R = rot_from_axisangle(axisangle)
    t = translation.clone()
    if invert:
        R = R.transpose(1, 2)
        t = t * -1
    Tm = torch.zeros(t.shape[0], 4, 4, dtype=t.dtype)
    Tm[:, 0, 0] = Tm[:, 1, 1] = Tm[:, 2, 2] = Tm[:, 3, 3] = 1
    Tm[:, :3, 3, None] = t.contiguous().view(-1, 3, 1)
    return torch.matmul(R, Tm) if invert else torch.matmul(Tm, R)
