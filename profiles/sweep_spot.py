#!/usr/bin/env python
"""A few large-batch points of profiles/sweep_loss.py (scalability check of the feature backward's overflow path)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import sweep_loss as s

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
torch.cuda.set_stream(torch.cuda.Stream(dev))
for (B, H, W) in [(8, 192, 640), (32, 192, 640), (64, 192, 640), (32, 384, 1280)]:
    ms, ips, gbs = s.run_case(B, H, W, (0, -1, 1), 10, dev)
    print(B, H, W, round(ms, 3), "ms", round(ips), "images/s", round(gbs), "GB/s", flush=True)
