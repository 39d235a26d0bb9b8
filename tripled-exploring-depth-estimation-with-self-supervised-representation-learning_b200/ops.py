"""torch.autograd.Functions over the C ABI of libtdl.so.

PyTorch is plumbing here: it owns the device memory and the CUDA stream; every
number is produced by the hand-written kernels behind include/tdl.h.  Inputs must be
CUDA tensors -- there is deliberately no CPU path.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import EdgeArgs, EdgeMultiArgs, FeatArgs, PhotoArgs, PoseArgs, ProjArgs, ReconArgs, TDL_MAX_LEVELS, TDL_MAX_SCALES, TDL_MAX_SRC


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.TdlError(f"{name}: expected a CUDA tensor (the fused loss has no CPU implementation)")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# --------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class PhotoConfig:
    """Scalar options of tdl_photo_args (config/cfg_kitti_fm.py:20-38 names in comments)."""
    n_src: int
    n_scales: int
    min_depth: float = 0.1              # min_depth
    max_depth: float = 100.0            # max_depth
    automask: bool = True               # automask
    disp_norm: bool = True              # disp_norm
    align_corners: bool = False         # torch >= 1.3 F.grid_sample default
    photo_coef: Tuple[float, ...] = ()  # 1/len(scales) per scale
    smooth_coef: Tuple[float, ...] = () # smoothness_weight / 2**s / len(scales)
    smooth_alpha: float = 0.5
    materialize: bool = True            # write outputs[("color",f,s)] and outputs[("min_index",s)]
    noise_seed: int = 0                 # Philox seed when no noise tensors are passed
    has_noise: bool = False


class PhotometricSmoothLoss(torch.autograd.Function):
    """losses[0:n] = min-reprojection loss per scale, losses[n:2n] = smoothness per scale.

    forward(cfg, target, P, invK, *src[S], *disp[n], *noise[n*S]?) ->
        (losses, *warped[n*S], *min_index[n])      (warped / min_index only if cfg.materialize)
    """

    @staticmethod
    def forward(ctx, cfg: PhotoConfig, target, P, invK, *tensors):
        L = _lib.lib()
        S, n = cfg.n_src, cfg.n_scales
        if not (1 <= S <= TDL_MAX_SRC and 1 <= n <= TDL_MAX_SCALES):
            raise _lib.TdlError(f"n_src={S} / n_scales={n} out of range")
        target = _f32c(target, "target")
        P = _f32c(P, "P")
        invK = _f32c(invK, "invK")
        srcs = [_f32c(t, "src") for t in tensors[:S]]
        disps = [_f32c(t, "disp") for t in tensors[S:S + n]]
        noise = [_f32c(t, "noise") for t in tensors[S + n:]] if cfg.has_noise else []
        if cfg.has_noise and len(noise) != n * S:
            raise _lib.TdlError("noise: expected n_scales * n_src tensors")
        B, _, H, W = target.shape
        if P.shape != (B, S, 3, 4) or invK.shape != (B, 3, 3):
            raise _lib.TdlError(f"P {tuple(P.shape)} / invK {tuple(invK.shape)}: expected (B,S,3,4) / (B,3,3)")
        dev = target.device
        a = PhotoArgs()
        a.B, a.H, a.W, a.S, a.nscales = B, H, W, S, n
        dh = (C.c_int32 * TDL_MAX_SCALES)()
        dw = (C.c_int32 * TDL_MAX_SCALES)()
        for s, d in enumerate(disps):
            if d.shape[0] != B or d.shape[1] != 1:
                raise _lib.TdlError(f"disp[{s}] {tuple(d.shape)}: expected (B,1,h,w)")
            dh[s], dw[s] = d.shape[2], d.shape[3]
            a.disp_h[s], a.disp_w[s] = d.shape[2], d.shape[3]
            a.disp[s] = d.data_ptr()
            a.photo_coef[s] = cfg.photo_coef[s]
            a.smooth_coef[s] = cfg.smooth_coef[s]
        a.automask, a.disp_norm, a.align_corners = int(cfg.automask), int(cfg.disp_norm), int(cfg.align_corners)
        a.min_depth, a.max_depth = cfg.min_depth, cfg.max_depth
        a.smooth_alpha = cfg.smooth_alpha
        a.noise_seed = cfg.noise_seed & 0xFFFFFFFFFFFFFFFF
        a.target = target.data_ptr()
        for f, t in enumerate(srcs):
            if t.shape != target.shape:
                raise _lib.TdlError("src/target shape mismatch")
            a.src[f] = t.data_ptr()
        a.P, a.invK = P.data_ptr(), invK.data_ptr()
        for i, t in enumerate(noise):
            a.noise[i // S][i % S] = t.data_ptr()
        warped, min_index = [], []
        if cfg.materialize:
            for s in range(n):
                for f in range(S):
                    w = torch.empty_like(target)
                    warped.append(w)
                    a.warped[s][f] = w.data_ptr()
                m = torch.empty((B, H, W), dtype=torch.int64, device=dev)
                min_index.append(m)
                a.min_index[s] = m.data_ptr()
        ws_bytes = L.tdl_photo_ws_bytes(B, H, W, S, n, dh, dw)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        losses = torch.empty(2 * n, dtype=torch.float32, device=dev)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws_bytes
        a.losses = losses.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(L.tdl_photo_fwd(C.byref(a), _stream()), "tdl_photo_fwd")
        ctx.cfg = cfg
        ctx.set_materialize_grads(False)      # no zero-filled "gradients" for the warped images / index maps
        ctx.n_warped = len(warped)
        # the backward re-reads the materialised warps instead of re-projecting the tile halo
        ctx.save_for_backward(target, P, invK, ws, *srcs, *disps, *warped)
        ctx.mark_non_differentiable(*warped, *min_index)
        return (losses, *warped, *min_index)

    @staticmethod
    def backward(ctx, g_losses, *_unused):
        L = _lib.lib()
        cfg = ctx.cfg
        S, n = cfg.n_src, cfg.n_scales
        target, P, invK, ws = ctx.saved_tensors[:4]
        if g_losses is None:
            g_losses = torch.zeros(2 * n, dtype=torch.float32, device=target.device)
        srcs = ctx.saved_tensors[4:4 + S]
        disps = ctx.saved_tensors[4 + S:4 + S + n]
        warped = ctx.saved_tensors[4 + S + n:4 + S + n + ctx.n_warped]
        B, _, H, W = target.shape
        a = PhotoArgs()
        a.B, a.H, a.W, a.S, a.nscales = B, H, W, S, n
        for i, t in enumerate(warped):
            a.warped[i // S][i % S] = t.data_ptr()
        d_disps = []
        for s, d in enumerate(disps):
            a.disp_h[s], a.disp_w[s] = d.shape[2], d.shape[3]
            a.disp[s] = d.data_ptr()
            a.photo_coef[s] = cfg.photo_coef[s]
            a.smooth_coef[s] = cfg.smooth_coef[s]
            g = torch.empty_like(d)
            d_disps.append(g)
            a.d_disp[s] = g.data_ptr()
        a.automask, a.disp_norm, a.align_corners = int(cfg.automask), int(cfg.disp_norm), int(cfg.align_corners)
        a.min_depth, a.max_depth = cfg.min_depth, cfg.max_depth
        a.smooth_alpha = cfg.smooth_alpha
        a.target = target.data_ptr()
        for f, t in enumerate(srcs):
            a.src[f] = t.data_ptr()
        a.P, a.invK = P.data_ptr(), invK.data_ptr()
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        scratch = torch.empty(2 * n, dtype=torch.float32, device=target.device)
        a.losses = scratch.data_ptr()
        g_losses = _f32c(g_losses, "grad")
        dP = torch.empty_like(P)
        a.dlosses, a.dP = g_losses.data_ptr(), dP.data_ptr()
        with torch.cuda.device(target.device):
            _lib.check(L.tdl_photo_bwd(C.byref(a), _stream()), "tdl_photo_bwd")
        n_noise = n * S if cfg.has_noise else 0
        return (None, None, dP, None, *([None] * S), *d_disps, *([None] * n_noise))


# --------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class FeatConfig:
    n_src: int
    min_depth: float = 0.1
    max_depth: float = 100.0
    align_corners: bool = False
    coef: float = 1e-3                  # perception_weight (/ len(scales) in mono_fm)
    materialize: bool = True            # write outputs[("feature",f,0)] and the argmin map
    layout: str = "auto"                # "auto": channel-last kernels when the target features ARE channels_last (what a
                                        # channels_last extractor hands over) or bf16; "nchw" / "nhwc" force one (with a copy)


def _feat_plan(cfg, tgt):
    """-> (nhwc: bool, torch dtype) of the feature tensors handed to tdl_feat_*."""
    C = tgt.shape[1]
    bf16 = tgt.dtype == torch.bfloat16
    is_cl = tgt.dim() == 4 and tgt.is_contiguous(memory_format=torch.channels_last) and not tgt.is_contiguous()
    if cfg.layout == "nhwc" or bf16:
        nhwc = True
    elif cfg.layout == "nchw":
        nhwc = False
    else:
        nhwc = is_cl
    if nhwc and C % 4 != 0:
        if bf16:
            raise _lib.TdlError("bf16 feature maps need C % 4 == 0")
        nhwc = False
    return nhwc, (torch.bfloat16 if bf16 else torch.float32)


def _feat_tensor(t, name, nhwc, dtype):
    if not t.is_cuda:
        raise _lib.TdlError(f"{name}: expected a CUDA tensor (the fused loss has no CPU implementation)")
    if t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous(memory_format=torch.channels_last) if nhwc else t.contiguous()


class FeatureMetricLoss(torch.autograd.Function):
    """forward(cfg, tgt, disp, P, invK, *src[S]) -> (loss[1], *warped[S], min_index)"""

    @staticmethod
    def forward(ctx, cfg: FeatConfig, tgt, disp, P, invK, *srcs):
        L = _lib.lib()
        S = cfg.n_src
        nhwc, fdtype = _feat_plan(cfg, tgt)
        tgt = _feat_tensor(tgt, "tgt_feat", nhwc, fdtype)
        disp = _f32c(disp, "disp")
        P = _f32c(P, "P")
        invK = _f32c(invK, "invK")
        srcs = [_feat_tensor(t, "src_feat", nhwc, fdtype) for t in srcs]
        if len(srcs) != S or not 1 <= S <= TDL_MAX_SRC:
            raise _lib.TdlError("src_feat: wrong number of source frames")
        B, Cc, h, w = tgt.shape
        if P.shape != (B, S, 3, 4) or invK.shape != (B, 3, 3):
            raise _lib.TdlError(f"P {tuple(P.shape)} / invK {tuple(invK.shape)}: expected (B,S,3,4) / (B,3,3)")
        a = FeatArgs()
        a.B, a.C, a.h, a.w, a.S = B, Cc, h, w, S
        a.disp_h, a.disp_w = disp.shape[2], disp.shape[3]
        a.align_corners = int(cfg.align_corners)
        a.layout = _lib.TDL_LAYOUT_NHWC if nhwc else _lib.TDL_LAYOUT_NCHW
        a.dtype = _lib.TDL_DTYPE_BF16 if fdtype == torch.bfloat16 else _lib.TDL_DTYPE_F32
        a.min_depth, a.max_depth, a.coef = cfg.min_depth, cfg.max_depth, cfg.coef
        a.tgt, a.disp, a.P, a.invK = tgt.data_ptr(), disp.data_ptr(), P.data_ptr(), invK.data_ptr()
        warped, extra = [], []
        for f, t in enumerate(srcs):
            if t.shape != tgt.shape:
                raise _lib.TdlError("src_feat/tgt_feat shape mismatch")
            a.src[f] = t.data_ptr()
            if cfg.materialize:
                wv = torch.empty_like(tgt)
                warped.append(wv)
                a.warped[f] = wv.data_ptr()
        if cfg.materialize:
            mi = torch.empty((B, h, w), dtype=torch.int64, device=tgt.device)
            extra.append(mi)
            a.min_index = mi.data_ptr()
        ws_bytes = L.tdl_feat_ws_bytes(B, Cc, h, w, S)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=tgt.device)
        loss = torch.empty(1, dtype=torch.float32, device=tgt.device)
        a.workspace, a.workspace_bytes, a.loss = ws.data_ptr(), ws_bytes, loss.data_ptr()
        with torch.cuda.device(tgt.device):
            _lib.check(L.tdl_feat_fwd(C.byref(a), _stream()), "tdl_feat_fwd")
        ctx.cfg = cfg
        ctx.nhwc = nhwc
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(tgt, disp, P, invK, ws, *srcs)
        ctx.mark_non_differentiable(*warped, *extra)
        return (loss, *warped, *extra)

    @staticmethod
    def backward(ctx, g_loss, *_unused):
        L = _lib.lib()
        cfg = ctx.cfg
        S = cfg.n_src
        tgt, disp, P, invK, ws = ctx.saved_tensors[:5]
        srcs = ctx.saved_tensors[5:5 + S]
        if g_loss is None:
            g_loss = torch.zeros(1, dtype=torch.float32, device=tgt.device)
        B, Cc, h, w = tgt.shape
        need = ctx.needs_input_grad            # (cfg, tgt, disp, P, invK, *srcs)
        a = FeatArgs()
        a.B, a.C, a.h, a.w, a.S = B, Cc, h, w, S
        a.disp_h, a.disp_w = disp.shape[2], disp.shape[3]
        a.align_corners = int(cfg.align_corners)
        a.layout = _lib.TDL_LAYOUT_NHWC if ctx.nhwc else _lib.TDL_LAYOUT_NCHW
        a.dtype = _lib.TDL_DTYPE_BF16 if tgt.dtype == torch.bfloat16 else _lib.TDL_DTYPE_F32
        a.min_depth, a.max_depth, a.coef = cfg.min_depth, cfg.max_depth, cfg.coef
        a.tgt, a.disp, a.P, a.invK = tgt.data_ptr(), disp.data_ptr(), P.data_ptr(), invK.data_ptr()
        for f, t in enumerate(srcs):
            a.src[f] = t.data_ptr()
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        scratch = torch.empty(1, dtype=torch.float32, device=tgt.device)
        a.loss = scratch.data_ptr()
        g_loss = _f32c(g_loss, "grad")
        d_disp, dP = torch.empty_like(disp), torch.empty_like(P)
        a.dloss, a.d_disp, a.dP = g_loss.data_ptr(), d_disp.data_ptr(), dP.data_ptr()
        d_tgt = None
        if need[1]:
            d_tgt = torch.empty_like(tgt)
            a.d_tgt = d_tgt.data_ptr()
        d_srcs = [None] * S
        if any(need[5:5 + S]):
            d_srcs = [torch.empty_like(tgt) for _ in range(S)]
            for f, t in enumerate(d_srcs):
                a.d_src[f] = t.data_ptr()
            # scratch of the bucketed d_src gather (lives for this call only; the library falls back to its
            # atomic scatter kernel when C % 4 != 0)
            sc_bytes = L.tdl_feat_bwd_scratch_bytes(B, Cc, h, w, S)
            scratch_buf = torch.empty(sc_bytes, dtype=torch.uint8, device=tgt.device)
            a.bwd_scratch, a.bwd_scratch_bytes = scratch_buf.data_ptr(), sc_bytes
        with torch.cuda.device(tgt.device):
            _lib.check(L.tdl_feat_bwd(C.byref(a), _stream()), "tdl_feat_bwd")
        return (None, d_tgt, d_disp, dP, None, *d_srcs)


# --------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class EdgeConfig:
    alpha: float = 1.0
    first_coef: float = 1.0
    second_coef: float = 1.0


class EdgeAwareSmoothness(torch.autograd.Function):
    """get_feature_regularization_loss (mono/model/mono_fm_joint/net.py:309-330) on one level:
    forward(cfg, feature (B,C,h,w), image (B,3,H,W)) -> loss[1]."""

    @staticmethod
    def forward(ctx, cfg: EdgeConfig, feature, image):
        L = _lib.lib()
        feature = _f32c(feature, "feature")
        image = _f32c(image, "image")
        a, ws = EdgeAwareSmoothness._args(L, cfg, feature, image)
        loss = torch.empty(1, dtype=torch.float32, device=feature.device)
        a.loss = loss.data_ptr()
        with torch.cuda.device(feature.device):
            _lib.check(L.tdl_edge_smooth_fwd(C.byref(a), _stream()), "tdl_edge_smooth_fwd")
        ctx.cfg = cfg
        ctx.save_for_backward(feature, image, ws)
        return loss

    @staticmethod
    def _args(L, cfg, feature, image, ws=None):
        B, Cc, h, w = feature.shape
        a = EdgeArgs()
        a.B, a.C, a.h, a.w, a.H, a.W = B, Cc, h, w, image.shape[2], image.shape[3]
        a.alpha, a.first_coef, a.second_coef = cfg.alpha, cfg.first_coef, cfg.second_coef
        a.feature, a.image = feature.data_ptr(), image.data_ptr()
        if ws is None:
            ws = torch.empty(L.tdl_edge_ws_bytes(B, Cc, h, w), dtype=torch.uint8, device=feature.device)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        return a, ws

    @staticmethod
    def backward(ctx, g_loss):
        L = _lib.lib()
        feature, image, ws = ctx.saved_tensors
        a, _ = EdgeAwareSmoothness._args(L, ctx.cfg, feature, image, ws)
        scratch = torch.empty(1, dtype=torch.float32, device=feature.device)
        g_loss = _f32c(g_loss, "grad")
        d_feat = torch.empty_like(feature)
        a.loss, a.dloss, a.d_feature = scratch.data_ptr(), g_loss.data_ptr(), d_feat.data_ptr()
        with torch.cuda.device(feature.device):
            _lib.check(L.tdl_edge_smooth_bwd(C.byref(a), _stream()), "tdl_edge_smooth_bwd")
        return None, d_feat, None


class EdgeAwareSmoothnessMulti(torch.autograd.Function):
    """get_feature_regularization_loss on ALL encoder levels in one call (the reference loops over five levels,
    mono/model/mono_fm_joint/net.py:77-80): forward(cfgs (EdgeConfig per level), image (B,3,H,W), *features) ->
    losses[n_levels].  Three launches forward, one backward, whatever the number of levels."""

    @staticmethod
    def forward(ctx, cfgs, image, *features):
        L = _lib.lib()
        n = len(features)
        if not 1 <= n <= TDL_MAX_LEVELS or len(cfgs) != n:
            raise _lib.TdlError(f"expected 1..{TDL_MAX_LEVELS} feature levels with one EdgeConfig each")
        image = _f32c(image, "image")
        feats = [_f32c(f, "feature") for f in features]
        losses = torch.empty(n, dtype=torch.float32, device=image.device)
        m, keep = EdgeAwareSmoothnessMulti._args(L, cfgs, feats, image, losses, None)
        with torch.cuda.device(image.device):
            _lib.check(L.tdl_edge_smooth_multi_fwd(C.byref(m), _stream()), "tdl_edge_smooth_multi_fwd")
        ctx.cfgs = cfgs
        ctx.save_for_backward(image, *feats, *keep)
        return losses

    @staticmethod
    def _args(L, cfgs, feats, image, losses, wss):
        m = EdgeMultiArgs()
        m.nlevels = len(feats)
        keep = []
        for l, (cfg, f) in enumerate(zip(cfgs, feats)):
            a, ws = EdgeAwareSmoothness._args(L, cfg, f, image, None if wss is None else wss[l])
            a.loss = losses[l:l + 1].data_ptr()
            m.level[l] = a
            keep.append(ws)
        return m, keep

    @staticmethod
    def backward(ctx, g_losses):
        L = _lib.lib()
        n = len(ctx.cfgs)
        image = ctx.saved_tensors[0]
        feats = ctx.saved_tensors[1:1 + n]
        wss = ctx.saved_tensors[1 + n:1 + 2 * n]
        scratch = torch.empty(n, dtype=torch.float32, device=image.device)
        m, _ = EdgeAwareSmoothnessMulti._args(L, ctx.cfgs, feats, image, scratch, wss)
        g_losses = _f32c(g_losses, "grad")
        d_feats = [torch.empty_like(f) for f in feats]
        for l in range(n):
            m.level[l].dloss = g_losses[l:l + 1].data_ptr()
            m.level[l].d_feature = d_feats[l].data_ptr()
        with torch.cuda.device(image.device):
            _lib.check(L.tdl_edge_smooth_multi_bwd(C.byref(m), _stream()), "tdl_edge_smooth_multi_bwd")
        return (None, None, *d_feats)


# --------------------------------------------------------------------------------------------------
class MaskedReconstructionLoss(torch.autograd.Function):
    """img_reconstruct_loss of the TripleD family (mono/model/mono_fm_joint_inpaint/net.py:80-91):
    forward(coef, pred (B,3,h,w), target (B,3,h,w), mask (B,3,h,w) | None) -> loss[1] =
    coef * sum(rho(pred, target) * (1 - mask)) / sum(1 - mask); gradient w.r.t. pred only."""

    @staticmethod
    def forward(ctx, coef: float, pred, target, mask):
        L = _lib.lib()
        pred = _f32c(pred, "pred")
        target = _f32c(target, "target")
        mask = None if mask is None else _f32c(mask, "mask")
        if target.shape != pred.shape or pred.shape[1] != 3 or (mask is not None and mask.shape != pred.shape):
            raise _lib.TdlError("pred / target / mask must all be (B,3,h,w)")
        a, ws = MaskedReconstructionLoss._args(L, coef, pred, target, mask)
        loss = torch.empty(1, dtype=torch.float32, device=pred.device)
        a.loss = loss.data_ptr()
        with torch.cuda.device(pred.device):
            _lib.check(L.tdl_recon_fwd(C.byref(a), _stream()), "tdl_recon_fwd")
        ctx.coef = coef
        ctx.has_mask = mask is not None
        ctx.save_for_backward(pred, target, ws, *([mask] if mask is not None else []))
        return loss

    @staticmethod
    def _args(L, coef, pred, target, mask, ws=None):
        a = ReconArgs()
        a.B, a.h, a.w, a.coef = pred.shape[0], pred.shape[2], pred.shape[3], coef
        a.pred, a.target = pred.data_ptr(), target.data_ptr()
        if mask is not None:
            a.mask = mask.data_ptr()
        if ws is None:
            ws = torch.empty(L.tdl_recon_ws_bytes(), dtype=torch.uint8, device=pred.device)
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        return a, ws

    @staticmethod
    def backward(ctx, g_loss):
        L = _lib.lib()
        pred, target, ws = ctx.saved_tensors[:3]
        mask = ctx.saved_tensors[3] if ctx.has_mask else None
        a, _ = MaskedReconstructionLoss._args(L, ctx.coef, pred, target, mask, ws)
        scratch = torch.empty(1, dtype=torch.float32, device=pred.device)
        g_loss = _f32c(g_loss, "grad")
        d_pred = torch.empty_like(pred)
        a.loss, a.dloss, a.d_pred = scratch.data_ptr(), g_loss.data_ptr(), d_pred.data_ptr()
        with torch.cuda.device(pred.device):
            _lib.check(L.tdl_recon_bwd(C.byref(a), _stream()), "tdl_recon_bwd")
        return None, d_pred, None, None


# --------------------------------------------------------------------------------------------------
class PoseTransform(torch.autograd.Function):
    """transformation_from_parameters (mono/model/mono_fm/net.py:201-253) in one launch:
    forward(axisangle (B,3) | (B,1,3), translation (B,3) | (B,1,3), invert: bool) -> cam_T_cam (B,4,4)."""

    @staticmethod
    def forward(ctx, axisangle, translation, invert):
        L = _lib.lib()
        aa = _f32c(axisangle, "axisangle").reshape(-1, 3)
        tr = _f32c(translation, "translation").reshape(-1, 3)
        if aa.shape != tr.shape:
            raise _lib.TdlError(f"axisangle {tuple(axisangle.shape)} / translation {tuple(translation.shape)} mismatch")
        B = aa.shape[0]
        T = torch.empty((B, 4, 4), dtype=torch.float32, device=aa.device)
        a = PoseArgs()
        a.B, a.invert = B, int(bool(invert))
        a.axisangle, a.translation, a.T = aa.data_ptr(), tr.data_ptr(), T.data_ptr()
        with torch.cuda.device(aa.device):
            _lib.check(L.tdl_pose_fwd(C.byref(a), _stream()), "tdl_pose_fwd")
        ctx.invert = bool(invert)
        ctx.shapes = (axisangle.shape, translation.shape)
        ctx.save_for_backward(aa, tr)
        return T

    @staticmethod
    def backward(ctx, gT):
        L = _lib.lib()
        aa, tr = ctx.saved_tensors
        gT = _f32c(gT, "grad")
        d_aa, d_tr = torch.empty_like(aa), torch.empty_like(tr)
        a = PoseArgs()
        a.B, a.invert = aa.shape[0], int(ctx.invert)
        a.axisangle, a.translation, a.dT = aa.data_ptr(), tr.data_ptr(), gT.data_ptr()
        a.d_axisangle, a.d_translation = d_aa.data_ptr(), d_tr.data_ptr()
        with torch.cuda.device(aa.device):
            _lib.check(L.tdl_pose_bwd(C.byref(a), _stream()), "tdl_pose_bwd")
        return d_aa.reshape(ctx.shapes[0]), d_tr.reshape(ctx.shapes[1]), None


def pose_transform(axisangle, translation, invert=False):
    """Drop-in for the reference's ``self.transformation_from_parameters(axisangle, translation, invert)`` on CUDA tensors."""
    return PoseTransform.apply(axisangle, translation, invert)


# --------------------------------------------------------------------------------------------------
class ProjectionPrologue(torch.autograd.Function):
    """Everything the loss kernels need from the camera matrices in one launch each way (include/tdl.h tdl_proj_args):
    forward(K (B,4,4), inv_K (B,4,4), *T (B,4,4) per source frame) -> (P_full (B,S,3,4), P_half (B,S,3,4),
    invK3 (B,3,3), invKh3 (B,3,3)); gradients flow to the T's only (K, inv_K are data)."""

    @staticmethod
    def forward(ctx, K, inv_K, *Ts):
        L = _lib.lib()
        K, inv_K = _f32c(K, "K"), _f32c(inv_K, "inv_K")
        Ts = [_f32c(t, "cam_T_cam") for t in Ts]
        S, B = len(Ts), K.shape[0]
        if not 1 <= S <= TDL_MAX_SRC or K.shape != (B, 4, 4) or inv_K.shape != (B, 4, 4) or any(t.shape != (B, 4, 4) for t in Ts):
            raise _lib.TdlError("ProjectionPrologue: expected K, inv_K and 1..4 transforms of shape (B,4,4)")
        dev = K.device
        P_full = torch.empty((B, S, 3, 4), dtype=torch.float32, device=dev)
        P_half = torch.empty_like(P_full)
        invK3 = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
        invKh3 = torch.empty_like(invK3)
        a = ProjArgs()
        a.B, a.S = B, S
        a.K, a.inv_K = K.data_ptr(), inv_K.data_ptr()
        for f, t in enumerate(Ts):
            a.T[f] = t.data_ptr()
        a.P_full, a.P_half, a.invK3, a.invKh3 = P_full.data_ptr(), P_half.data_ptr(), invK3.data_ptr(), invKh3.data_ptr()
        with torch.cuda.device(dev):
            _lib.check(L.tdl_proj_fwd(C.byref(a), _stream()), "tdl_proj_fwd")
        ctx.S = S
        ctx.save_for_backward(K)
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(invK3, invKh3)
        return P_full, P_half, invK3, invKh3

    @staticmethod
    def backward(ctx, g_full, g_half, *_unused):
        L = _lib.lib()
        (K,) = ctx.saved_tensors
        S, B = ctx.S, K.shape[0]
        need = ctx.needs_input_grad[2:2 + S]
        if (g_full is None and g_half is None) or not any(need):
            return (None, None, *([None] * S))
        a = ProjArgs()
        a.B, a.S = B, S
        a.K = K.data_ptr()
        if g_full is not None:
            g_full = _f32c(g_full, "grad")
            a.dP_full = g_full.data_ptr()
        if g_half is not None:
            g_half = _f32c(g_half, "grad")
            a.dP_half = g_half.data_ptr()
        dTs = [torch.empty((B, 4, 4), dtype=torch.float32, device=K.device) if need[f] else None for f in range(S)]
        for f, t in enumerate(dTs):
            if t is not None:
                a.dT[f] = t.data_ptr()
        with torch.cuda.device(K.device):
            _lib.check(L.tdl_proj_bwd(C.byref(a), _stream()), "tdl_proj_bwd")
        return (None, None, *dTs)


def projection_prologue(K, inv_K, Ts):
    return ProjectionPrologue.apply(K, inv_K, *Ts)
