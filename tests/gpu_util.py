"""Helpers for the GPU parity tests: run the CUDA path on a golden / synthetic record."""
import importlib

import torch
import torch.nn as nn

PKG = "tripled-exploring-depth-estimation-with-self-supervised-representation-learning_b200"


def pkg():
    return importlib.import_module(PKG)


def make_loss_net(opt_dict, kind):
    tdl = pkg()

    class LossNet(nn.Module, tdl.ViewSynthesisLossMixin):
        _smooth_weight_key = "disparity_smoothness" if kind == "baseline" else "smoothness_weight"

        def __init__(self, opt):
            super().__init__()
            self.opt = opt

    return LossNet(tdl.config.ConfigDict(opt_dict))


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


FEAT_LAYOUT = "nchw"      # "nhwc": hand the feature maps over in channels_last memory (tests/test_gpu_feat_nhwc.py)


def _leaf(k, v, dev):
    v = v.to(dev).clone()
    is_feat = k == "tgt_feat" or (isinstance(k, tuple) and k[0] == "src_feat")
    if is_feat and FEAT_LAYOUT == "nhwc":
        v = v.contiguous(memory_format=torch.channels_last)
    return v.requires_grad_(True)


def run_cuda(rec, noise, dev="cuda", kind=None):
    """rec: golden record (or same layout); noise[s][f] CPU tensors.  Returns loss_dict, outputs, grads."""
    meta = rec["meta"]
    kind = kind or meta["kind"]
    net = make_loss_net(meta["opt"], kind)
    inputs = {k: v.to(dev) for k, v in rec["inputs"].items()}
    leaves = {k: _leaf(k, v, dev) for k, v in rec["leaves"].items()}
    outputs = {k: v for k, v in leaves.items()
               if isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam", "res_img", "auto_res_img")}
    noise_d = {s: {f: n.to(dev) for f, n in d.items()} for s, d in noise.items()}
    frames = meta["opt"]["frame_ids"][1:]
    if kind == "baseline":
        loss = net.compute_losses_baseline(inputs, outputs, noise_d)
    else:
        src = {f: leaves[("src_feat", f)] for f in frames}
        if kind == "fm":
            loss = net.compute_losses_fm(inputs, outputs, noise_d, leaves["tgt_feat"], src)
        elif kind == "joint":
            feats = [leaves["tgt_feat"]] + [leaves[("feat_level", i)] for i in range(1, 5)]
            loss = net.compute_losses_joint(inputs, outputs, feats, noise_d, src)
        else:
            feats = [leaves["tgt_feat"]] + [leaves[("feat_level", i)] for i in range(1, 5)]
            loss = net.compute_losses_joint_core(inputs, outputs, feats, noise_d, src)
            if kind == "tripled":
                loss.update(net.compute_auto_res_loss(inputs, outputs))
    total = sum(v.mean() for v in loss.values())
    total.backward()
    torch.cuda.synchronize()
    grads = {k: (v.grad.detach().cpu() if v.grad is not None else None) for k, v in leaves.items()}
    loss = {k: v.detach().cpu() for k, v in loss.items()}
    outs = {k: v.detach().cpu() for k, v in outputs.items()
            if (isinstance(k, tuple) and k[0] in ("color", "feature", "min_index", "min_index_photo"))
            or k == "min_index"}
    return loss, outs, grads
