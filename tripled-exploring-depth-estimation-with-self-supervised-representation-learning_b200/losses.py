"""Host-side mirror of the reference's loss methods, backed by the fused kernels.

``ViewSynthesisLossMixin`` gives a net class the reference's method surface

    compute_losses(inputs, outputs[, features]) -> loss_dict
    generate_images_pred(inputs, outputs, scale) -> outputs
    generate_features_pred(inputs, outputs)       -> outputs

with the same ``inputs`` / ``outputs`` / ``loss_dict`` keys, weights and divisions as
    mono/model/mono_baseline/net.py:51-100      (Baseline)
    mono/model/mono_fm/net.py:69-133            (mono_fm)
    mono/model/mono_fm_joint_inpaint/net.py:47-133 (TripleD family, view-synthesis part)
but evaluates them with one fused CUDA pass over all scales instead of ~150 eager
kernels per scale.  The mixin owns no parameters or buffers, so checkpoints stay
compatible with the reference's state-dict keys.

Automask tie-break noise (net.py:94).  The reference draws ``torch.randn`` on the CPU
per scale and source frame and copies it to the GPU.  ``noise_mode``:
    "philox"    (default) counter-based N(0,1) inside the kernel, no memory traffic;
    "reference" the reference's own draws: torch.randn from the global CPU generator in
                the same order (scale-major, then frame), then an H2D copy -- bit-parity
                with the reference's RNG stream;
    a dict noise[scale][frame_id] of (B,1,H,W) tensors may also be passed per call.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from .geometry import half_res_intrinsics, projection_matrix
import torch.nn.functional as F

from .ops import (projection_prologue, EdgeAwareSmoothness, EdgeAwareSmoothnessMulti, EdgeConfig, FeatConfig, FeatureMetricLoss, MaskedReconstructionLoss,
                  PhotoConfig, PhotometricSmoothLoss)


class LossDict(dict):
    """The reference's ``loss_dict`` (same keys, 0-dim tensors) plus ``total()``: the sum of the ``.mean()`` of all
    entries -- what ``batch_processor`` computes (mono/apis/trainer.py:39-48).  Entries that came out of one fused
    kernel call are registered as a *part* (the raw kernel output vector, the keys it covers); ``total()`` sums each
    part with one kernel and adds ``.mean()`` of every entry no part covers, so terms a subclass appends the
    reference's way (``loss_dict[k] = v`` / ``loss_dict.update(...)``) are never dropped from the objective."""

    def __init__(self):
        super().__init__()
        self._parts = []          # (tensor, multiplicity, keys covered)

    def add_part(self, tensor, mult=1, keys=()):
        self._parts.append((tensor, mult, tuple(keys)))

    def absorb(self, other: "LossDict"):
        """update() with another LossDict, keeping its parts."""
        super().update(other)
        self._parts += other._parts

    def __setitem__(self, key, value):
        # overwriting an entry a part stands in for drops that part: its other entries (views of the part's
        # tensor) are then summed one by one, which is the same value
        if key in self and self[key] is not value:
            self._parts = [p for p in self._parts if key not in p[2]]
        super().__setitem__(key, value)

    def __delitem__(self, key):
        self._parts = [p for p in self._parts if key not in p[2]]
        super().__delitem__(key)

    def update(self, *a, **kw):
        for k, v in dict(*a, **kw).items():
            self[k] = v

    _weights = {}                 # (device, ((numel, multiplicity), ...)) -> constant weight vector of the fused sum

    def _fused_parts_sum(self):
        """sum_i mult_i * sum(part_i) as ONE concatenation + ONE dot product (backward: one multiply, the parts' gradients
        are contiguous slices of it) instead of a reduce / scale / add chain per part and an expand + copy per gradient:
        the step is a serial chain of kernels, so each 2 us launch between the forward and the backward counts."""
        parts = [(t.reshape(-1), m) for t, m, _ in self._parts]
        dev = parts[0][0].device
        if any(t.device != dev or t.dtype != torch.float32 for t, _ in parts):
            return None
        key = (dev, tuple((t.numel(), float(m)) for t, m in parts))
        w = LossDict._weights.get(key)
        if w is None:
            if dev.type == "cuda" and torch.cuda.is_current_stream_capturing():
                return None       # (a host -> device copy cannot be captured; the eager warm-up steps fill the cache)
            w = torch.tensor([m for t, m in parts for _ in range(t.numel())], dtype=torch.float32, device=dev)
            LossDict._weights[key] = w
        flat = torch.cat([t for t, _ in parts])
        self._packed = flat.detach()
        return torch.dot(flat, w)

    def packed(self):
        """After total(): every kernel-produced loss value in ONE device vector (the parts in registration order), for
        logging with a single device->host copy instead of one ``.item()`` per entry (mono/apis/trainer.py:39-54);
        None when total() did not take the fused path."""
        return getattr(self, "_packed", None)

    def total(self):
        out = None
        covered = set()
        fused = self._fused_parts_sum() if len(self._parts) > 1 else None
        for t, m, keys in self._parts:
            covered.update(keys)
            if fused is not None:
                continue
            v = t.sum() if t.dim() else t
            v = v * m if m != 1 else v
            out = v if out is None else out + v
        if fused is not None:
            out = fused
        for k, v in self.items():
            if k not in covered:
                v = v.mean() if v.dim() else v
                out = v if out is None else out + v
        return out


def _scalar(t):
    """0-dim VIEW of a one-element kernel output.  (``t[0]`` would do, but its backward materialises a zero vector and
    copies the gradient into it: two launches per loss term on a step that is a serial chain of kernels.)"""
    return t.reshape(())


def _opt_get(opt, name, default=None):
    if hasattr(opt, "get"):
        return opt.get(name, default)
    return getattr(opt, name, default)


class ViewSynthesisLossMixin:
    """Expects ``self.opt`` with the reference's option names (config/cfg_kitti_fm.py:20-38)."""

    noise_mode = "philox"
    materialize_outputs = True          # write outputs[("color",f,s)] / ("feature",f,0) / ("min_index",s)
    grid_sample_align_corners = False   # torch >= 1.3 default, which is what the reference runs with today
    _smooth_weight_key = "smoothness_weight"
    overlap_streams = True              # feature-metric kernels on a second stream (see compute_losses_fm)
    _streams = {}

    def _side_stream(self, device):
        key = (device.type, device.index)
        st = ViewSynthesisLossMixin._streams.get(key)
        if st is None:
            st = ViewSynthesisLossMixin._streams[key] = torch.cuda.Stream(device)
        return st

    # ------------------------------------------------------------------ helpers
    def _src_frames(self):
        return list(self.opt.frame_ids[1:])

    def _pose(self, inputs, outputs, f):
        return inputs["stereo_T"] if f == "s" else outputs[("cam_T_cam", 0, f)]

    def _stack_P(self, inputs, outputs, K):
        return torch.stack([projection_matrix(K, self._pose(inputs, outputs, f)) for f in self._src_frames()], 1)

    def _camera(self, inputs, outputs):
        """(P_full, P_half, invK3, invKh3) for this call: ONE launch (tdl_proj_fwd) instead of ~15 small PyTorch kernels
        (matmul / slice / stack per frame, clone + scale of K and inv_K for the half-resolution feature path), and one
        instead of ~15 in the backward.  Cached on the outputs dict so that the photometric and the feature-metric term of
        one step share it."""
        Ts = [self._pose(inputs, outputs, f) for f in self._src_frames()]
        key = tuple((t.data_ptr(), t._version) for t in Ts + [inputs["K"], inputs["inv_K"]])
        cached = outputs.get("_tdl_camera")
        cam = cached[1] if cached is not None and cached[0] == key else None
        if cam is None:
            if inputs["K"].is_cuda:
                cam = projection_prologue(inputs["K"], inputs["inv_K"], Ts)
            else:                                   # (CPU tensors only reach this in host-logic tests; the kernels refuse them later)
                Kh, invKh = half_res_intrinsics(inputs["K"], inputs.get("inv_K"))
                cam = (self._stack_P(inputs, outputs, inputs["K"]), self._stack_P(inputs, outputs, Kh),
                       inputs["inv_K"][:, :3, :3], invKh[:, :3, :3].contiguous())
            outputs["_tdl_camera"] = (key, cam)
        return cam

    def _reference_noise(self, scales, batch, device):
        H, W = self.opt.height, self.opt.width
        return {s: {f: torch.randn(batch, 1, H, W).to(device, non_blocking=True) for f in self._src_frames()}
                for s in scales}

    # ------------------------------------------------------------------ fused photometric + smoothness
    def _photometric(self, inputs, outputs, scales: Sequence[int], noise=None, materialize=None):
        opt = self.opt
        frames = self._src_frames()
        target = inputs[("color", 0, 0)]
        n = len(opt.scales)
        if noise is None and opt.automask and self.noise_mode == "reference":
            noise = self._reference_noise(scales, target.shape[0], target.device)
        sw = _opt_get(opt, self._smooth_weight_key, 0.0)
        # Philox stream: (torch.initial_seed(), per-instance step counter) -- reproducible per model instance after
        # torch.manual_seed, independent of how many other loss objects live in the process
        step = self.__dict__.get("_tdl_noise_step", 0) + 1
        self.__dict__["_tdl_noise_step"] = step
        cfg = PhotoConfig(
            n_src=len(frames), n_scales=len(scales),
            min_depth=float(opt.min_depth), max_depth=float(opt.max_depth),
            automask=bool(opt.automask), disp_norm=bool(opt.disp_norm),
            align_corners=self.grid_sample_align_corners,
            photo_coef=tuple(1.0 / n for _ in scales),
            smooth_coef=tuple(sw / (2 ** s) / n for s in scales),
            materialize=self.materialize_outputs if materialize is None else materialize,
            noise_seed=(torch.initial_seed() * 1000003 + step) & (2 ** 63 - 1),
            has_noise=noise is not None and bool(opt.automask))
        P, _, invK, _ = self._camera(inputs, outputs)
        tensors = [inputs[("color", f, 0)] for f in frames] + [outputs[("disp", 0, s)] for s in scales]
        if cfg.has_noise:
            tensors += [noise[s][f] for s in scales for f in frames]
        res = PhotometricSmoothLoss.apply(cfg, target, P, invK, *tensors)
        losses = res[0]
        if cfg.materialize:
            k = 1
            for s in scales:
                for f in frames:
                    outputs[("color", f, s)] = res[k]
                    k += 1
            for s in scales:
                outputs[("min_index", s)] = res[k]
                outputs[("min_index_photo", s)] = res[k]      # survives mono_fm's overwrite (net.py:117)
                k += 1
        return losses

    def _refuse_silent_no_grad(self, what, tensors):
        """The stand-alone ``generate_*_pred`` calls return tensors WITHOUT autograd history (the fused kernels own the
        backward of the loss, not of the warped tensors).  The reference's versions are differentiable; a caller that
        builds its own loss term on them (as mono/model/mono_fm_joint_im_rot/net.py:53-133 does) would silently train
        nothing -- so that case is an error, not a silent no-op."""
        if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
            raise RuntimeError(
                f"{what}: the stand-alone call returns warped tensors without autograd history, but its inputs require "
                "grad.  Gradients of the view-synthesis losses flow through compute_losses(); call this method under "
                "torch.no_grad() if the warped tensors are only inspected (visualisation, metrics), or build the extra "
                "loss term from compute_losses' entries.")

    def generate_images_pred(self, inputs, outputs, scale):
        """mono/model/mono_fm/net.py:157-170.  Stand-alone use warps the sources for ONE scale; the images are produced by
        the fused kernel and carry no autograd history (gradients flow through compute_losses) -- see
        _refuse_silent_no_grad for what happens when a caller would need them to."""
        self._refuse_silent_no_grad("generate_images_pred", [outputs[("disp", 0, scale)]] +
                                    [self._pose(inputs, outputs, f) for f in self._src_frames()])
        with torch.no_grad():
            self._photometric(inputs, outputs, [scale], materialize=True)
        return outputs

    # ------------------------------------------------------------------ fused feature-metric
    def _feature_metric(self, inputs, outputs, tgt_f, src_fs: Dict, coef: float, materialize=None):
        opt = self.opt
        frames = self._src_frames()
        _, P, _, invKh3 = self._camera(inputs, outputs)
        cfg = FeatConfig(n_src=len(frames), min_depth=float(opt.min_depth), max_depth=float(opt.max_depth),
                         align_corners=self.grid_sample_align_corners, coef=float(coef),
                         materialize=self.materialize_outputs if materialize is None else materialize)
        res = FeatureMetricLoss.apply(cfg, tgt_f, outputs[("disp", 0, 0)], P, invKh3, *[src_fs[f] for f in frames])
        if cfg.materialize:
            for i, f in enumerate(frames):
                outputs[("feature", f, 0)] = res[1 + i]
        return _scalar(res[0]), (res[1 + len(frames)] if cfg.materialize else None)

    def _extract(self, img):
        """``extractor(img)[0]`` (mono/model/mono_fm/net.py:113,197).  Encoders that can stop after their first
        level (``first``) are not run through the four residual stages whose outputs the loss never reads."""
        ext = getattr(self, "extractor", None) or getattr(self, "Encoder")
        return ext.first(img) if hasattr(ext, "first") else ext(img)[0]

    def generate_features_pred(self, inputs, outputs):
        """mono/model/mono_fm/net.py:172-199 (stand-alone: warped features, no autograd history)."""
        self._refuse_silent_no_grad("generate_features_pred", [outputs[("disp", 0, 0)]] +
                                    [self._pose(inputs, outputs, f) for f in self._src_frames()] +
                                    list(getattr(self, "extractor", None).parameters() if hasattr(self, "extractor") else []))
        with torch.no_grad():
            src = {f: self._extract(inputs[("color", f, 0)]) for f in self._src_frames()}
            tgt = self._extract(inputs[("color", 0, 0)])
            self._feature_metric(inputs, outputs, tgt, src, 0.0, materialize=True)
        return outputs

    # ------------------------------------------------------------------ the three loss families
    def compute_losses_baseline(self, inputs, outputs, noise=None):
        """mono/model/mono_baseline/net.py:51-100."""
        scales = list(self.opt.scales)
        losses = self._photometric(inputs, outputs, scales, noise)
        loss_dict = LossDict()
        for i, s in enumerate(scales):
            loss_dict[("min_reconstruct_loss", s)] = losses[i]
            loss_dict[("smooth_loss", s)] = losses[len(scales) + i]
        loss_dict.add_part(losses, 1, [("min_reconstruct_loss", s) for s in scales] + [("smooth_loss", s) for s in scales])
        return loss_dict

    def compute_losses_fm(self, inputs, outputs, noise=None, tgt_f=None, src_fs=None):
        """mono/model/mono_fm/net.py:69-133.  The reference re-runs generate_features_pred and the
        extractor inside the scale loop (net.py:85,113); the values are scale-independent, so they are
        evaluated once and the same loss tensor is returned under every ('min_perceptional_loss', s)."""
        opt = self.opt
        scales = list(opt.scales)
        if src_fs is None:
            src_fs = {f: self._extract(inputs[("color", f, 0)]) for f in self._src_frames()}
        if tgt_f is None:
            tgt_f = self._extract(inputs[("color", 0, 0)])
        # The feature-metric kernels (HBM / L2-atomic bound) and the photometric kernels (instruction bound) are
        # independent until the losses are summed: issue them on two streams so that they share the SMs.  Autograd
        # runs each backward on its forward's stream, so the backward kernels overlap the same way.
        side = self._side_stream(tgt_f.device) if self.overlap_streams else None
        self._camera(inputs, outputs)          # on the CURRENT stream, before the fork: both streams read its outputs
        if side is not None:
            main = torch.cuda.current_stream(tgt_f.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                per, idx = self._feature_metric(inputs, outputs, tgt_f, src_fs, opt.perception_weight / len(scales))
            loss_dict = self.compute_losses_baseline(inputs, outputs, noise)
            main.wait_stream(side)
            # allocated while `side` was current but consumed on `main`: tell the caching allocator, so the blocks are
            # not handed to the next side-stream kernel while main-stream readers are still queued
            for t in [per, idx] + [outputs.get(("feature", f, 0)) for f in self._src_frames()]:
                if t is not None:
                    t.record_stream(main)
        else:
            loss_dict = self.compute_losses_baseline(inputs, outputs, noise)
            per, idx = self._feature_metric(inputs, outputs, tgt_f, src_fs, opt.perception_weight / len(scales))
        ordered = LossDict()
        for s in scales:
            ordered[("min_reconstruct_loss", s)] = loss_dict[("min_reconstruct_loss", s)]
            ordered[("min_perceptional_loss", s)] = per
            ordered[("smooth_loss", s)] = loss_dict[("smooth_loss", s)]
        ordered._parts = list(loss_dict._parts)
        ordered.add_part(per, len(scales), [("min_perceptional_loss", s) for s in scales])
        for s in scales:
            if idx is not None:
                outputs[("min_index", s)] = idx        # the reference overwrites it (net.py:117)
        return ordered

    def compute_losses_joint(self, inputs, outputs, features, noise=None, src_fs=None):
        """mono/model/mono_fm_joint/net.py:73-155 (the ``mono_fm_joint`` net): feature regularisation on the five encoder
        levels (:77-80); per scale the UN-masked autoencoder reconstruction term SSIM + L1 of ``("res_img", 0, s)``
        against the bilinearly resized target (:96-101), the photometric min-reprojection term (:103-131), the
        feature-metric term under per-scale keys (:133-143; scale independent, evaluated once like compute_losses_fm)
        and the smoothness term (:145-153)."""
        opt = self.opt
        scales = list(opt.scales)
        n = len(scales)
        target = inputs[("color", 0, 0)]
        loss_dict = LossDict()
        self._feature_regularization(loss_dict, features, target)
        if src_fs is None:
            src_fs = {f: self._extract(inputs[("color", f, 0)]) for f in self._src_frames()}
        per, idx = self._feature_metric(inputs, outputs, features[0], src_fs, opt.perception_weight / n)
        base = self.compute_losses_baseline(inputs, outputs, noise)
        for s in scales:
            res = outputs[("res_img", 0, s)]
            tgt_r = F.interpolate(target, list(res.shape[-2:]), mode="bilinear", align_corners=False)
            loss_dict[("img_reconstruct_loss", s)] = _scalar(MaskedReconstructionLoss.apply(1.0 / n, res, tgt_r, None))
            loss_dict[("min_reconstruct_loss", s)] = base[("min_reconstruct_loss", s)]
            loss_dict[("min_perceptional_loss", s)] = per
            loss_dict[("smooth_loss", s)] = base[("smooth_loss", s)]
            if idx is not None:
                outputs[("min_index", s)] = idx        # the reference overwrites the photometric map (net.py:142)
        loss_dict._parts += base._parts
        loss_dict.add_part(per, n, [("min_perceptional_loss", s) for s in scales])
        return loss_dict

    def compute_losses_joint_core(self, inputs, outputs, features, noise=None, src_fs=None):
        """View-synthesis part of mono/model/mono_fm_joint_inpaint/net.py:47-133: one un-divided
        feature-metric term on features[0] (:58-70) and the per-scale photometric / smoothness terms
        (:96-131), plus get_feature_regularization_loss on the five feature levels (:53-56)."""
        opt = self.opt
        loss_dict = LossDict()
        target = inputs[("color", 0, 0)]
        if features is not None:
            self._feature_regularization(loss_dict, features, target)
            if src_fs is None:
                src_fs = {f: self._extract(inputs[("color", f, 0)]) for f in self._src_frames()}
            per, idx = self._feature_metric(inputs, outputs, features[0], src_fs, opt.perception_weight)
            loss_dict["min_perceptional_loss"] = per
            if idx is not None:
                outputs["min_index"] = idx
        w_rec = _opt_get(opt, "img_reconstruct_weight", 1)
        if features is not None and w_rec != 0:
            mask = inputs[("mask", 0, 0)]
            for s in opt.scales:                 # autoencoder / in-painting term (net.py:80-91)
                res = outputs[("res_img", 0, s)]
                size = list(res.shape[-2:])
                tgt_r = F.interpolate(target, size, mode="bilinear", align_corners=False)
                msk_r = F.interpolate(mask, size, mode="bilinear", align_corners=False)
                rec = _scalar(MaskedReconstructionLoss.apply(float(w_rec) / len(opt.scales), res, tgt_r, msk_r))
                loss_dict[("img_reconstruct_loss", s)] = rec
        base = self.compute_losses_baseline(inputs, outputs, noise)
        loss_dict.absorb(base)
        return loss_dict

    def compute_auto_res_loss(self, inputs, outputs):
        """mono/model/mono_fm_joint_inpaint/net.py:520-527: one element-wise robust L1 (left to PyTorch);
        like the reference the entry is the un-reduced (B,1,H,W) map, which batch_processor averages."""
        loss_dict = {}
        if _opt_get(self.opt, "auto_res_weight", 0.0) > 0.0:
            target = inputs[("color", 0, 0)]
            res = outputs[("auto_res_img", 0, 0)]
            l1 = torch.sqrt(torch.pow(res - target, 2) + 1e-3 ** 2).mean(1, True)
            loss_dict["auto_res_loss"] = l1 * self.opt.auto_res_weight
        return loss_dict

    def _feature_regularization(self, loss_dict, features, target, n_levels=5):
        """('feature_regularization_loss', i) = get_feature_regularization_loss(features[i], target) / 2**i / 5 for the five
        encoder levels (mono/model/mono_fm_joint/net.py:77-80), all levels in one multi-level call: the per-level factor is
        folded into the coefficients, so every entry is a view of one kernel-written vector."""
        cfgs = tuple(EdgeConfig(alpha=1.0, first_coef=-float(self.opt.dis) / (2 ** i) / n_levels,
                                second_coef=float(self.opt.cvt) / (2 ** i) / n_levels) for i in range(n_levels))
        reg = EdgeAwareSmoothnessMulti.apply(cfgs, target, *features[:n_levels])
        for i in range(n_levels):
            loss_dict[("feature_regularization_loss", i)] = reg[i]
        loss_dict.add_part(reg, 1, [("feature_regularization_loss", i) for i in range(n_levels)])

    def get_feature_regularization_loss(self, feature, img):
        """mono/model/mono_fm_joint/net.py:309-330: -dis * first-order + cvt * second-order."""
        cfg = EdgeConfig(alpha=1.0, first_coef=-float(self.opt.dis), second_coef=float(self.opt.cvt))
        return _scalar(EdgeAwareSmoothness.apply(cfg, feature, img))
