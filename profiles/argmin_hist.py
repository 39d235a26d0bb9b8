#!/usr/bin/env python
"""Arg-min channel histogram of the bench workload (which of [identity -1, identity +1, warp -1, warp +1] wins per pixel
and scale, and which source wins the feature-metric min): tells how sparse the backward's gradients are.
    python profiles/argmin_hist.py        (on a B200)"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

dev = torch.device("cuda", 0)
B, H, W = 8, 192, 640
host = bench.make_host_workload(B, H, W, 1234 + 1)
step = bench.DeviceStep(host, B, H, W, dev, True)
b = step.buf
inputs = {k[1]: v for k, v in b.items() if k[0] == "in"}
outputs = {k[1]: v for k, v in b.items() if k[0] == "leaf" and isinstance(k[1], tuple) and k[1][0] in ("disp", "cam_T_cam")}
src = {f: b[("leaf", ("src_feat", f))] for f in bench.FRAME_IDS[1:]}
step.net.compute_losses_fm(inputs, outputs, None, b[("leaf", "tgt_feat")], src)
torch.cuda.synchronize()
for k, v in sorted(outputs.items(), key=str):
    if isinstance(k, tuple) and k[0] in ("min_index", "min_index_photo") or k == "min_index":
        h = torch.bincount(v.flatten().cpu(), minlength=4).float()
        print(k, tuple(v.shape), [round(float(x), 4) for x in (h / h.sum())])
