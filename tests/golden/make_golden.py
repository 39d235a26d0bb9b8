"""Generates tests/golden/*.pt by executing the REAL reference loss code
(/root/reference, torch 2.11 CPU, fp32) on seeded synthetic inputs.

Run in the build container only:  python tests/golden/make_golden.py
The reference ships no tests or golden vectors, so these
files are the parity pin: "the reference's own Python, executed".
"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

PKG = importlib.import_module(
    "tripled-exploring-depth-estimation-with-self-supervised-representation-learning_b200")
synth = importlib.import_module(PKG.__name__ + ".synth")

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # name: (kind, B, H, W, frames, feat_channels, seed)
    # "waves": band-limited frames/features (synth._waves) -- strict gradient parity (no derivative kinks)
    "baseline_waves_b2_64x96": ("baseline", 2, 64, 96, "waves", 0, 1234),
    "fm_waves_b2_64x96_c8": ("fm", 2, 64, 96, "waves", 8, 1236),
    "inpaint_waves_b1_96x128_c8": ("inpaint", 1, 96, 128, "waves", 8, 1237),
    # mono_fm_joint (mono/model/mono_fm_joint/net.py:73-155): per-scale perceptual keys, UN-masked reconstruction
    "joint_waves_b1_96x128_c8": ("joint", 1, 96, 128, "waves", 8, 1240),
    # TripleD net (cfg_kitti_tripleD): + masked autoencoder reconstruction (16 erased 8x8 holes) + auto_res_loss
    "tripled_waves_b1_96x128_c8": ("tripled", 1, 96, 128, "waves", 8, 1238),
    # "smooth": box-filtered noise frames -- realistic, judged with the kink-robust metric
    "baseline_smooth_b2_64x96": ("baseline", 2, 64, 96, "smooth", 0, 1234),
    # 32x64: disp_3 is 2x4, so d_dyy is empty and the reference's smooth_loss is NaN
    # (mean of an empty tensor) -- kept as the degenerate-shape edge case; white-noise frames.
    "baseline_white_b1_32x64": ("baseline", 1, 32, 64, "white", 0, 1235),
}


def run_reference(kind, B, H, W, frames, C, seed):
    opt = ref_loader.default_opt(B, H, W, dis=1e-3, cvt=1e-3, img_reconstruct_weight=1 if kind == "tripled" else 0,
                                 auto_res_weight=5e-3)
    inputs, outputs, extras = synth.make_inputs(B, H, W, seed=seed, frames=frames,
                                                feat_channels=C, with_noise=False)
    leaves = {}
    for k in list(outputs):
        outputs[k] = outputs[k].clone().requires_grad_(True)
        leaves[k] = outputs[k]
    net = ref_loader.make_loss_only_net(kind, opt)
    features = None
    if C:
        tgt = extras["tgt_feat"].clone().requires_grad_(True)
        leaves["tgt_feat"] = tgt
        table = {}
        for f, t in extras["src_feats"].items():
            t = t.clone().requires_grad_(True)
            leaves[("src_feat", f)] = t
            table[inputs[("color", f, 0)].data_ptr()] = t
        table[inputs[("color", 0, 0)].data_ptr()] = tgt
        (net.extractor if kind == "fm" else net.Encoder).table = table
        if kind in ("inpaint", "tripled", "joint"):
            # 5 feature levels; only level 0 enters the view-synthesis path, the
            # others feed get_feature_regularization_loss.
            g = torch.Generator().manual_seed(seed + 99)
            features = [tgt]
            for i in range(1, 5):
                f_i = torch.randn(B, 4, H >> (i + 1), W >> (i + 1), generator=g).requires_grad_(True)
                leaves[("feat_level", i)] = f_i
                features.append(f_i)
            mask = torch.ones(B, 3, H, W)
            if kind == "tripled":                  # erase mask (mono/datasets/kitti_dataset.py:167-182) + decoder outputs
                for _ in range(16):
                    y0 = int(torch.randint(0, H - 8, (1,), generator=g))
                    x0 = int(torch.randint(0, W - 8, (1,), generator=g))
                    mask[:, :, y0:y0 + 8, x0:x0 + 8] = 0
            if kind in ("tripled", "joint"):
                for s in range(4):
                    r = torch.sigmoid(torch.nn.functional.avg_pool2d(
                        torch.randn(B, 3, (H >> s) + 4, (W >> s) + 4, generator=g), 5, 1) * 2).requires_grad_(True)
                    leaves[("res_img", 0, s)] = r
                    outputs[("res_img", 0, s)] = r
            if kind == "tripled":
                a = (inputs[("color", 0, 0)] + 0.05 * torch.randn(B, 3, H, W, generator=g)).clamp(0, 1).requires_grad_(True)
                leaves[("auto_res_img", 0, 0)] = a
                outputs[("auto_res_img", 0, 0)] = a
            inputs[("mask", 0, 0)] = mask
    torch.manual_seed(seed)          # the reference draws automask noise from the global CPU RNG
    with ref_loader.cpu_cuda_shim(force=True):
        if kind in ("inpaint", "tripled", "joint"):
            loss_dict = net.compute_losses(inputs, outputs, features)
        else:
            loss_dict = net.compute_losses(inputs, outputs)
    total = sum(v.mean() for v in loss_dict.values())      # batch_processor: .mean() of every entry (trainer.py:39-48)
    total.backward()
    rec = {
        "inputs": {k: v.detach().clone() for k, v in inputs.items()},
        "leaves": {k: v.detach().clone() for k, v in leaves.items()},
        "meta": dict(kind=kind, B=B, H=H, W=W, frames=frames, C=C, seed=seed,
                     torch=torch.__version__, opt=dict(opt)),
        "loss": {k: v.detach().clone() for k, v in loss_dict.items()},
        "out": {k: (v.detach().to(torch.uint8) if v.dtype == torch.int64 else v.detach().clone())
                for k, v in outputs.items()
                if isinstance(k, tuple) and k[0] in ("color", "feature", "min_index") or k == "min_index"},
        "grad": {k: (v.grad.detach().clone() if v.grad is not None else None) for k, v in leaves.items()},
    }
    return rec


def main():
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        rec = run_reference(*case)
        path = os.path.join(HERE, name + ".pt")
        torch.save(rec, path)
        print(name, os.path.getsize(path) // 1024, "KiB",
              {str(k): float(v.mean()) for k, v in rec["loss"].items()})


if __name__ == "__main__":
    main()
