"""Diagnostic: identical frames -> rho must be 0.15e-3 exactly; prints the excess for several shapes / kernels."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
from gpu_util import pkg, run_cuda
tdl = pkg()
for (H, W, frames) in [(32, 32, "smooth"), (32, 64, "smooth"), (64, 32, "smooth"), (64, 96, "smooth"), (64, 96, "white"), (192, 640, "smooth")]:
    inputs, outputs, _ = tdl.synth.make_inputs(1, H, W, seed=7, with_noise=False, frames=frames)
    for f in (-1, 1):
        inputs[("color", f, 0)] = inputs[("color", 0, 0)].clone()
    rec = {"inputs": inputs, "leaves": dict(outputs),
           "meta": dict(kind="baseline", B=1, H=H, W=W, frames=frames, C=0, seed=7,
                        opt=dict(frame_ids=[0, -1, 1], imgs_per_gpu=1, height=H, width=W, scales=[0, 1, 2, 3], min_depth=0.1,
                                 max_depth=100.0, automask=True, disp_norm=True, disparity_smoothness=1e-3,
                                 smoothness_weight=1e-3, perception_weight=1e-3))}
    zero = {s: {f: torch.zeros(1, 1, H, W) for f in (-1, 1)} for s in range(4)}
    for v1 in (1, 0):
        with tdl._lib.options(photo_v1=v1):
            loss, outs, _ = run_cuda(rec, zero)
        ex = [float(loss[("min_reconstruct_loss", s)]) * 4 / 0.15e-3 - 1 for s in range(4)]
        nw = [int((outs[("min_index", s)] > 1).sum()) for s in range(4)]
        print(H, W, frames, "v1" if v1 else "v2", "relative excess", ["%.2e" % e for e in ex], "warped wins", nw)
