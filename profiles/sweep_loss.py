#!/usr/bin/env python
"""BASELINE.json configs[4]: isolated fused-loss sweep (fwd+bwd, CUDA-graph replay, device-resident inputs) over
resolution x batch x number of source frames.  Prints a markdown table.

    python profiles/sweep_loss.py [--quick] [--frames smooth|scene] > profiles/r2_sweep.md      (on a B200)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def run_case(B, H, W, frame_ids, steps, device, frames="smooth"):
    bench.FRAME_IDS = tuple(frame_ids)
    host = bench.make_host_workload(B, H, W, 99, frames)
    step = bench.DeviceStep(host, B, H, W, device, True)
    for _ in range(2):
        step.run_eager()
    step.capture()
    for _ in range(3):
        step.replay()
    ms = bench.timed_region(step.replay, steps, device, False) / steps
    S = len(frame_ids) - 1
    alg = bench.algorithmic_bytes(B, H, W, S, bench.FEAT_C, 4, True)
    total = sum(alg[k] for k in ("photo_fwd", "photo_bwd", "feat_fwd", "feat_bwd", "memset_dsrc"))
    del step
    torch.cuda.empty_cache()
    return ms, B / (ms * 1e-3), total / (ms * 1e-3) / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--frames", default="smooth,scene")
    args = ap.parse_args()
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    torch.cuda.set_stream(torch.cuda.Stream(device))
    peak = 6539.5
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peak = float(json.load(open(pp))["hbm_gbs"])
    res = [(96, 320), (192, 640), (320, 1024), (384, 1280)]
    batches = [1, 8, 32] if args.quick else [1, 2, 4, 8, 16, 32, 64]
    print("# Fused mono_fm loss fwd+bwd sweep (BASELINE configs[4]) -- B200, fp32, C=64 channels_last features, 4 scales\n")
    print(f"Algorithmic GB/s = SURVEY 8(d) bytes / time; roofline = {peak:.0f} GB/s (measured HBM copy).  Workloads: 'smooth' = SURVEY 8(d) "
          "recipe (auto-masked, sparse backward), 'scene' = rendered moving camera (dense backward).\n")
    print("| workload | H x W | S | batch | ms/step | images/s | alg. GB/s | % of HBM roofline |")
    print("|---|---|---:|---:|---:|---:|---:|---:|")
    for frames in args.frames.split(","):
        for (H, W) in res:
            for fids in ((0, -1, 1), (0, -2, -1, 1, 2)):
                for B in batches:
                    if B * H * W > 64 * 192 * 640 * 2:          # bound the memory of the largest cases
                        continue
                    steps = max(5, min(50, int(2e6 / (B * H * W / 1000))))
                    try:
                        ms, ips, gbs = run_case(B, H, W, fids, steps, device, frames)
                        print(f"| {frames} | {H}x{W} | {len(fids) - 1} | {B} | {ms:.3f} | {ips:.0f} | {gbs:.0f} | {100 * gbs / peak:.1f} |", flush=True)
                    except Exception as exc:                      # keep the sweep going
                        print(f"| {frames} | {H}x{W} | {len(fids) - 1} | {B} | failed: {type(exc).__name__}: {str(exc)[:60]} | | | |", flush=True)
                        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
