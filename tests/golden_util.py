"""Shared helpers for the golden-fixture tests (CPU and GPU)."""
import os

import torch

from oracle import restatement as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(f[:-3] for f in os.listdir(GOLDEN) if f.endswith(".pt"))


def load_case(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), weights_only=False)


def spec_from_meta(meta):
    o = meta["opt"]
    return R.LossSpec(height=meta["H"], width=meta["W"], frame_ids=tuple(o["frame_ids"]),
                      scales=tuple(o["scales"]), min_depth=o["min_depth"], max_depth=o["max_depth"],
                      automask=o["automask"], disp_norm=o["disp_norm"],
                      smoothness_weight=o["smoothness_weight"] if meta["kind"] != "baseline"
                      else o["disparity_smoothness"],
                      perception_weight=o["perception_weight"], extra=dict(o))


def reference_noise(spec, meta):
    """The reference draws the automask noise from the global CPU generator
    seeded with meta['seed'] (tests/golden/make_golden.py), scale-major."""
    torch.manual_seed(meta["seed"])
    return R.draw_automask_noise(spec, meta["B"])


def run_restatement(rec, dtype=torch.float32, forced=None):
    """Runs oracle.restatement on a golden record -> (loss_dict, outputs, leaves)."""
    meta = rec["meta"]
    spec = spec_from_meta(meta)
    inputs = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in rec["inputs"].items()}
    leaves = {k: v.to(dtype).clone().requires_grad_(True) for k, v in rec["leaves"].items()}
    outputs = {k: v for k, v in leaves.items()
               if isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam", "res_img", "auto_res_img")}
    noise = reference_noise(spec, meta)
    noise = {s: {f: n.to(dtype) for f, n in d.items()} for s, d in noise.items()}
    kind = meta["kind"]
    if kind == "baseline":
        loss = R.compute_losses_baseline(spec, inputs, outputs, noise, forced=forced)
    else:
        src = {f: leaves[("src_feat", f)] for f in spec.frame_ids[1:]}
        if kind == "fm":
            loss = R.compute_losses_fm(spec, inputs, outputs, noise, leaves["tgt_feat"], src, forced=forced)
        elif kind == "joint":
            feats = [leaves["tgt_feat"]] + [leaves[("feat_level", i)] for i in range(1, 5)]
            loss = R.compute_losses_joint(spec, inputs, outputs, noise, feats, src, forced=forced)
        else:
            loss = {}
            feats = [leaves["tgt_feat"]] + [leaves[("feat_level", i)] for i in range(1, 5)]
            for i, f in enumerate(feats):
                loss[("feature_regularization_loss", i)] = R.feature_regularization_loss(
                    f, inputs[("color", 0, 0)], spec.extra["dis"], spec.extra["cvt"]) / (2 ** i) / 5
            if kind == "tripled":
                for sc in spec.scales:
                    loss[("img_reconstruct_loss", sc)] = R.img_reconstruct_loss(spec, inputs, outputs, sc)
            loss.update(R.compute_losses_inpaint_core(spec, inputs, outputs, noise, leaves["tgt_feat"], src,
                                                       forced=forced))
            if kind == "tripled":
                loss["auto_res_loss"] = R.auto_res_loss(inputs, outputs, spec.extra["auto_res_weight"])
    return loss, outputs, leaves
