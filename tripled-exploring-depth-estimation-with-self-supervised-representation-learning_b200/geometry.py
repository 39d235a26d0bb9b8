"""Pose / intrinsics helpers that sit just outside the fused kernels.

These are 12-floats-per-image computations; they stay in PyTorch so autograd
reaches the pose network untouched.  Device-agnostic.

Reference behaviour mirrored (paths relative to the reference repo):
  transformation_from_parameters  mono/model/mono_fm/net.py:201-212
  get_translation_matrix          mono/model/mono_fm/net.py:214-223
  rot_from_axisangle              mono/model/mono_fm/net.py:225-253
  K/2 + per-sample pinverse       mono/model/mono_fm/net.py:185-191
"""
from __future__ import annotations

import torch


def rot_from_axisangle(vec: torch.Tensor) -> torch.Tensor:
    """Axis-angle (B,1,3) -> homogeneous rotation (B,4,4) (Rodrigues)."""
    angle = torch.norm(vec, 2, 2, True)
    axis = vec / (angle + 1e-7)
    ca, sa = torch.cos(angle), torch.sin(angle)
    C = 1 - ca
    x, y, z = axis[..., 0:1], axis[..., 1:2], axis[..., 2:3]
    xs, ys, zs = x * sa, y * sa, z * sa
    xC, yC, zC = x * C, y * C, z * C
    xyC, yzC, zxC = x * yC, y * zC, z * xC
    zero = torch.zeros_like(ca)
    one = torch.ones_like(ca)
    rows = [x * xC + ca, xyC - zs, zxC + ys, zero,
            xyC + zs, y * yC + ca, yzC - xs, zero,
            zxC - ys, yzC + xs, z * zC + ca, zero,
            zero, zero, zero, one]
    return torch.cat(rows, 2).view(-1, 4, 4)


def get_translation_matrix(t: torch.Tensor) -> torch.Tensor:
    """(B,1,3) or (B,3) translation -> (B,4,4)."""
    t = t.contiguous().view(-1, 3, 1)
    T = torch.eye(4, dtype=t.dtype, device=t.device).repeat(t.shape[0], 1, 1)
    top = torch.cat([T[:, :3, :3], t], 2)
    return torch.cat([top, T[:, 3:, :]], 1)


def transformation_from_parameters(axisangle, translation, invert=False):
    R = rot_from_axisangle(axisangle)
    t = translation
    if invert:
        R = R.transpose(1, 2)
        t = t * -1
    T = get_translation_matrix(t)
    return torch.matmul(R, T) if invert else torch.matmul(T, R)


def half_res_intrinsics(K: torch.Tensor, inv_K: torch.Tensor | None = None):
    """K with rows 0,1 halved, and its inverse (mono/model/mono_fm/net.py:185-191).

    The reference runs a per-sample ``torch.pinverse`` (an SVD per image, per frame, per scale).
    K_half = D @ K with D = diag(1/2, 1/2, 1, 1), so inv(K_half) = inv(K) @ diag(2, 2, 1, 1): when the
    dataset's ``inv_K`` (= pinv(K), mono_dataset.py:176) is at hand the result is obtained by doubling its
    first two columns -- exact in fp32 -- instead of a decomposition; otherwise one batched pinv is used."""
    Kh = K.clone()
    Kh[:, 0:2, :] *= 0.5
    if inv_K is None:
        return Kh, torch.linalg.pinv(Kh)
    inv = inv_K.clone()
    inv[:, :, 0:2] *= 2.0
    return Kh, inv


def projection_matrix(K: torch.Tensor, T: torch.Tensor) -> torch.Tensor:
    """P = (K @ T)[:, :3, :] as in Project.forward (mono/model/mono_fm/layers.py:74).

    Evaluated as broadcast multiply + sum, NOT torch.matmul: with ``torch.backends.cuda.matmul.allow_tf32 = True`` (or
    ``torch.set_float32_matmul_precision("high")``, which training scripts commonly set for the networks) a matmul rounds K
    and T to 10 mantissa bits, which moves every projected pixel by ~1e-3 of its coordinate -- measured as a 1.3e-4
    relative error of the warped images.  These are 12 numbers per image; they stay full fp32 whatever the global flag."""
    return (K[:, :3, :, None] * T[:, None, :, :]).sum(2)
