#!/usr/bin/env python
"""Benchmark of the fused view-synthesis loss (forward + backward) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], cfg_kitti_fm): the mono_fm loss -- 4-scale photometric
(SSIM + L1, automask, min over 2 source frames) + feature-metric on 64-channel features at
(H/2, W/2) + edge-aware smoothness -- at 192x640, batch 8 per GPU, fp32, forward + backward,
on synthetic frames (package synth.py).  One "step" = one fwd+bwd pass over one batch.

Two synthetic workloads of that shape are measured (package synth.py):
  "smooth"  SURVEY.md 8(d)'s recipe (BASELINE's): globally shifted sources + pixel-level random disparity.  The un-warped
            sources win ~99 % of the arg-mins, i.e. a fully auto-masked scene: sparse photometric backward, scattered gathers.
  "scene"   a moving camera in a smooth 3-D scene, sources rendered through ground-truth depth + pose, prediction = truth +
            2 % error: the warped sources win ~92 % of the pixels (dense backward), the flow is coherent.  This is the
            regime of a real training step, so `roofline` names the dominant kernel of THIS workload.

Printed JSON line (rank 0):
  value      images/s on "smooth" with every input already resident in HBM (CUDA-graph replay of the step): the
             MEDIAN of `repeats.n` timed regions of exactly --steps replays each (min / max / all regions in `repeats`)
  scene      the same measurement on "scene": images/s, ms/step, per-kernel table, identity_frac
  e2e        images/s through the public host API with HOST (pinned) buffers: H2D of every input of the step, the step,
             D2H of the loss scalars -- inside the timed region; e2e_images_only: only what a data loader provides is
             uploaded (uint8 frames, K, inv_K), features come from a conv stem on the device
  roofline   dominant kernel of "scene": algorithmic bytes per launch / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the reference's own loss code (oracle/_ref, the unmodified reference modules; else the oracle port) timed
             on this box's host cores on a bounded sample (B=2) of the same workload
With --impl reference the reference arm is timed instead: the reference's mono_fm.compute_losses + backward on the host
cores at the SAME batch, steps and warm-up as asked for (kind "reference" when oracle/_ref travelled, else "port").
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "tripled-exploring-depth-estimation-with-self-supervised-representation-learning_b200"

FRAME_IDS = (0, -1, 1)
SCALES = (0, 1, 2, 3)
FEAT_C = 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=8, help="images per GPU")
    ap.add_argument("--height", type=int, default=192)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--frozen-extractor", action="store_true",
                    help="features do not require grad (pretrained, frozen extractor: mono_fm/net.py:24-25)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the full train-step measurement")
    ap.add_argument("--train-steps", type=int, default=10)
    ap.add_argument("--eager-train", action="store_true", help="do not CUDA-graph the train step")
    ap.add_argument("--train-model", default="both", choices=["both", "mono_fm", "tripled"],
                    help="mono_fm = cfg_kitti_fm (BASELINE configs[1]); tripled = cfg_kitti_tripleD (configs[2])")
    ap.add_argument("--bf16-comm", action="store_true", help="train step: all-reduce the gradients in bf16")
    ap.add_argument("--no-channels-last", action="store_true", help="train step: keep the networks in NCHW")
    ap.add_argument("--cpu-batch", type=int, default=2, help="batch of the bounded CPU sample of the cpu_baseline leg")
    ap.add_argument("--repeats", type=int, default=7, help="timed regions of --steps replays each (median is reported)")
    ap.add_argument("--no-scene", action="store_true", help="skip the representative-scene workload")
    ap.add_argument("--feat-layout", default="nhwc", choices=["nhwc", "nchw"],
                    help="memory layout of the 64-channel feature maps handed to the loss: nhwc = torch channels_last, what the "
                         "cuDNN extractor produces on B200 (default); nchw = the reference's contiguous layout")
    ap.add_argument("--no-bf16", action="store_true", help="skip the opt-in bf16-storage measurement of the scene workload")
    ap.add_argument("--ref-budget-s", type=float, default=240.0,
                    help="--impl reference: wall-clock budget; the per-step sample (batch) is halved until the run fits")
    ap.add_argument("--no-fused-adam", action="store_true",
                    help="train step: per-tensor Adam kernels instead of torch's fused multi-tensor Adam (33.0 vs 31.0 ms/step)")
    ap.add_argument("--seed-offset", type=int, default=0,
                    help="shifts the synthetic-data seed (rank r uses 1234 + r + offset): the sparse backward paths make the "
                         "step time data dependent, this shows by how much")
    return ap.parse_args()


def opt_dict(B, H, W):
    return dict(frame_ids=list(FRAME_IDS), imgs_per_gpu=B, height=H, width=W, scales=list(SCALES), min_depth=0.1,
                max_depth=100.0, automask=True, disp_norm=True, perception_weight=1e-3, smoothness_weight=1e-3)


# ------------------------------------------------------------------------------------------------ bytes
def algorithmic_bytes(B, H, W, S, C, n_scales, trainable_feat, materialize=True, noise_tensors=False):
    """compulsory fp32 traffic of one ideal pass, per kernel, for a batch of B."""
    N = H * W
    ns = [N // 4 ** (s + 1) for s in range(n_scales)]
    photo_fwd = sum(12 * N + 12 * S * N + (4 * S * N if noise_tensors else 0)
                    + ((12 * S * N + 8 * N) if materialize else 0) + 4 * n for n in ns)
    photo_bwd = sum(12 * N + 12 * S * N + 8 * N + 8 * n for n in ns)
    feat_fwd = C * N * (1 + S + (S if materialize else 0))            # C channels at N/4 pixels, 4 B each
    if trainable_feat:
        feat_bwd = C * N * (1 + S + 1 + S)        # reads tgt, src; writes d_tgt; atomically accumulates d_src
        memset_dsrc = C * N * S
    else:
        feat_bwd = C * N * (1 + S)
        memset_dsrc = 0
    # split forward: the warp kernel reads disp + sources and writes the warps; the scoring kernel reads target,
    # sources (identity terms) and the warps back, and writes min_index
    photo_warp = sum(12 * S * N + 12 * S * N + 4 * n for n in ns)
    photo_score = sum(12 * N + 12 * S * N + 12 * S * N + 8 * N for _ in ns)
    per_image = dict(photo_fwd=photo_fwd, photo_warp=photo_warp, photo_score=photo_score, photo_bwd=photo_bwd,
                     feat_fwd=feat_fwd, feat_bwd=feat_bwd, memset_dsrc=memset_dsrc)
    return {k: v * B for k, v in per_image.items()}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workload
def make_host_workload(B, H, W, seed, frames="smooth"):
    tdl = importlib.import_module(PKG)
    inputs, outputs, extras = tdl.synth.make_inputs(B, H, W, frame_ids=FRAME_IDS, scales=SCALES, seed=seed,
                                                    frames=frames, feat_channels=FEAT_C, with_noise=False)
    host = {}
    for k, v in inputs.items():
        host[("in", k)] = v
    for k, v in outputs.items():
        host[("leaf", k)] = v
    host[("leaf", "tgt_feat")] = extras["tgt_feat"]
    for f, t in extras["src_feats"].items():
        host[("leaf", ("src_feat", f))] = t
    return host


FEAT_LAYOUT = "nhwc"
FEAT_DTYPE = torch.float32


def _is_feat_key(k):
    return k[0] == "leaf" and (k[1] == "tgt_feat" or (isinstance(k[1], tuple) and k[1][0] == "src_feat"))


class DeviceStep:
    """The loss step on static device buffers: eager, or captured once into a CUDA graph and replayed."""

    def __init__(self, host, B, H, W, device, trainable_feat, feat_dtype=None):
        tdl = importlib.import_module(PKG)
        self.tdl, self.device = tdl, device
        feat_dtype = feat_dtype or FEAT_DTYPE
        if FEAT_LAYOUT == "nhwc" or feat_dtype != torch.float32:   # host buffers already in the layout / dtype the device wants
            host = {k: (v.to(feat_dtype).contiguous(memory_format=torch.channels_last) if _is_feat_key(k) else v)
                    for k, v in host.items()}
        self.buf = {k: (v.to(device) if _is_feat_key(k) else v.to(device).contiguous()) for k, v in host.items()}
        self.host = {k: v.pin_memory() for k, v in host.items()}
        self.trainable_feat = trainable_feat
        self.grad_keys = [k for k in self.buf if k[0] == "leaf" and
                          (trainable_feat or not (k[1] == "tgt_feat" or (isinstance(k[1], tuple) and k[1][0] == "src_feat")))]
        for k in self.grad_keys:
            self.buf[k].requires_grad_(True)

        import torch.nn as nn

        class LossNet(nn.Module, tdl.ViewSynthesisLossMixin):
            def __init__(self, opt):
                super().__init__()
                self.opt = opt

        self.net = LossNet(tdl.config.ConfigDict(opt_dict(B, H, W)))
        self.graph = None
        self.total = None
        self.losses_host = torch.empty(12, dtype=torch.float32).pin_memory()
        self.loss_vec = None

    def _step(self):
        b = self.buf
        for k in self.grad_keys:
            b[k].grad = None
        inputs = {k[1]: v for k, v in b.items() if k[0] == "in"}
        outputs = {k[1]: v for k, v in b.items() if k[0] == "leaf" and isinstance(k[1], tuple)
                   and k[1][0] in ("disp", "cam_T_cam")}
        src = {f: b[("leaf", ("src_feat", f))] for f in FRAME_IDS[1:]}
        loss_dict = self.net.compute_losses_fm(inputs, outputs, None, b[("leaf", "tgt_feat")], src)
        total = loss_dict.total()
        total.backward(self._seed(total))       # (a cached 1.0: no fill kernel per step)
        # the step's result for the host (the reference logs every entry, mono/apis/trainer.py:39-54): the packed vector
        # of the kernel outputs when the fused total produced one, else the stacked entries
        self.loss_vec = loss_dict.packed()
        if self.loss_vec is None:
            self.loss_vec = torch.stack([v.detach() for v in loss_dict.values()])
        self.last_outputs = outputs
        return total

    def _seed(self, total):
        if getattr(self, "_one", None) is None:
            self._one = torch.ones_like(total)
        return self._one

    def identity_frac(self):
        """Fraction of pixels whose photometric arg-min is an identity (auto-mask) channel, per scale."""
        S = len(FRAME_IDS) - 1
        return [round(float((self.last_outputs[("min_index_photo", s)] < S).float().mean()), 4) for s in SCALES]

    def run_eager(self):
        self.total = self._step()

    def capture(self):
        # the whole benchmark runs on one non-default stream (see main): autograd's AccumulateGrad nodes
        # remember the stream they were created on, and capture must not touch the legacy stream
        for _ in range(3):
            self._step()
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=torch.cuda.current_stream(self.device)):
            self.total = self._step()
        torch.cuda.synchronize(self.device)

    def replay(self):
        self.graph.replay()

    def h2d(self):
        for k, v in self.host.items():
            self.buf[k].detach().copy_(v, non_blocking=True)

    def d2h(self):
        if self.losses_host.numel() != self.loss_vec.numel():
            self.losses_host = torch.empty(self.loss_vec.numel(), dtype=torch.float32).pin_memory()
        self.losses_host.copy_(self.loss_vec, non_blocking=True)

    def h2d_bytes(self):
        return sum(v.numel() * v.element_size() for v in self.host.values())


SEED_OFFSET = 0


def rank_seed(rank):
    """Batch sharding: every rank draws its own synthetic batch (weak scaling, no data-path collective)."""
    return 1234 + rank + SEED_OFFSET


def max_over_ranks(ms, device, dist_on):
    """Multi-GPU numbers are the MAX over ranks of the device-timed region."""
    if not dist_on:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t)


def whole_job_images_per_s(world, batch, steps, ms):
    return world * batch * steps / (ms * 1e-3)


def timed_region(fn, steps, device, dist_on, warm=None):
    if warm is not None:
        warm()
    if dist_on:
        torch.distributed.barrier()
    torch.cuda.synchronize(device)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize(device)
    if dist_on:
        torch.distributed.barrier()
    return max_over_ranks(a.elapsed_time(b), device, dist_on)


def measure_workload(frames, args, device, rank, world, dist_on, trainable, feat_dtype=None):
    """Per-kernel device times (eager, CUDA events inside the library) and the graph-replay step time of one workload:
    `repeats` timed regions of exactly args.steps replays, each bracketed by barrier + synchronize, max over ranks;
    the median region is the reported figure."""
    tdl = importlib.import_module(PKG)
    B, H, W = args.batch, args.height, args.width
    host = make_host_workload(B, H, W, rank_seed(rank), frames)
    step = DeviceStep(host, B, H, W, device, trainable, feat_dtype)
    for _ in range(3):
        step.run_eager()
    torch.cuda.synchronize(device)
    prof_steps = min(args.steps, 20)
    # per-kernel times: eager steps on ONE stream (with the two-stream schedule a kernel's event pair also measures the
    # time it waits for the other stream's CTAs to drain, which is not its own duration)
    step.net.overlap_streams = False
    step.run_eager()
    torch.cuda.synchronize(device)
    tdl._lib.profile_begin()
    for _ in range(prof_steps):
        step.run_eager()
    kern = tdl._lib.profile_end()
    step.net.overlap_streams = True
    step.run_eager()
    torch.cuda.synchronize(device)
    ident = step.identity_frac()
    launches_per_step = sum(n for k, (n, _) in kern.items() if not k.startswith("memset")) // prof_steps
    step.capture()
    for _ in range(max(3, args.warmup)):
        step.replay()
    regions = [timed_region(step.replay, args.steps, device, dist_on) / args.steps for _ in range(max(1, args.repeats))]
    med = statistics.median(regions)
    return {"step": step, "host": host, "kern": kern, "prof_steps": prof_steps, "identity_frac": ident,
            "launches_per_step": launches_per_step, "ms_per_step": med,
            "repeats": {"n": len(regions), "steps_per_region": args.steps, "ms_per_step_median": round(med, 4),
                        "ms_per_step_min": round(min(regions), 4), "ms_per_step_max": round(max(regions), 4),
                        "ms_per_step_all": [round(r, 4) for r in regions]}}


def e2e_loop(sets, n, device):
    """n steps alternating between two device buffer sets: the copy stream uploads step i+1 while step i runs; the host
    waits for (and reads) the losses of every step."""
    main = torch.cuda.current_stream(device)
    copy_stream = e2e_loop.copy_stream.get(device)
    if copy_stream is None:
        copy_stream = e2e_loop.copy_stream[device] = torch.cuda.Stream(device)
    ev_copied = [torch.cuda.Event(), torch.cuda.Event()]
    ev_done = [torch.cuda.Event(), torch.cuda.Event()]
    for e in ev_done:
        e.record(main)
    with torch.cuda.stream(copy_stream):
        copy_stream.wait_stream(main)                 # the upload of step 0 starts inside the timed region
        sets[0].h2d()
        ev_copied[0].record(copy_stream)
    for i in range(n):
        cur, k = sets[i % 2], i % 2
        main.wait_event(ev_copied[k])
        if i + 1 < n:
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(ev_done[1 - k])    # the step that last read that buffer set is finished
                sets[1 - k].h2d()
                ev_copied[1 - k].record(copy_stream)
        cur.replay()
        cur.d2h()
        ev_done[k].record(main)
        ev_done[k].synchronize()                          # the host reads the loss of every step


e2e_loop.copy_stream = {}


class ImagesOnlyStep(DeviceStep):
    """e2e variant whose per-step upload is what a data loader provides -- uint8 frames, K, inv_K -- while the feature
    maps are produced on the device (a trainable 7x7 stride-2 conv + ReLU, the shape of the extractor's first level,
    mono/model/mono_autoencoder/encoder.py:37) and disparities / poses are resident stand-ins for the network outputs."""

    def __init__(self, host, B, H, W, device, trainable_feat=True):
        keep = {k: v for k, v in host.items() if not (k[0] == "leaf" and (k[1] == "tgt_feat" or (isinstance(k[1], tuple) and k[1][0] == "src_feat")))}
        super().__init__(keep, B, H, W, device, True)
        # what a decoder / PIL resize hands over: (B,H,W,3) uint8; the on-GPU input pipeline (tdl_input_fwd) turns them into
        # inputs[("color", f, 0)] and inputs[("color_aug", f, 0)] (torchvision ColorJitter, byte-exact) every step
        self.frames_u8 = {f: (host[("in", ("color", f, 0))] * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory()
                          for f in FRAME_IDS}
        self.frames_dev = {f: torch.empty_like(v, device=device) for f, v in self.frames_u8.items()}
        self.pipe = self.tdl.GpuInputPipeline(FRAME_IDS, H, W)
        self.params = {k: (v.to(device) if v is not None else None)
                       for k, v in self.pipe.sample_params(B, torch.Generator().manual_seed(5)).items()}
        self.host = {k: v for k, v in self.host.items() if k[0] == "in" and not isinstance(k[1], tuple)}   # K, inv_K
        g = torch.Generator().manual_seed(0)
        self.stem = (torch.randn(FEAT_C, 3, 7, 7, generator=g) * 0.1).to(device)
        if FEAT_LAYOUT == "nhwc":
            self.stem = self.stem.contiguous(memory_format=torch.channels_last)
        self.stem.requires_grad_(True)

    def _step(self):
        b = self.buf
        for k in self.grad_keys:
            b[k].grad = None
        self.stem.grad = None
        pre = self.pipe(self.frames_dev, self.params)
        imgs = {f: pre[("color", f, 0)] for f in FRAME_IDS}
        # the networks see the augmented frames, the loss the plain ones (mono/model/mono_fm/net.py:41-46)
        conv_in = {f: (pre[("color_aug", f, 0)].contiguous(memory_format=torch.channels_last) if FEAT_LAYOUT == "nhwc"
                       else pre[("color_aug", f, 0)]) for f in FRAME_IDS}
        inputs = {k[1]: v for k, v in b.items() if k[0] == "in" and not isinstance(k[1], tuple)}
        for f in FRAME_IDS:
            inputs[("color", f, 0)] = imgs[f]
        outputs = {k[1]: v for k, v in b.items() if k[0] == "leaf"}
        feats = {f: torch.relu(torch.nn.functional.conv2d(conv_in[f], self.stem, stride=2, padding=3)) for f in FRAME_IDS}
        loss_dict = self.net.compute_losses_fm(inputs, outputs, None, feats[0], {f: feats[f] for f in FRAME_IDS[1:]})
        total = loss_dict.total()
        total.backward(self._seed(total))       # (a cached 1.0: no fill kernel per step)
        # the step's result for the host (the reference logs every entry, mono/apis/trainer.py:39-54): the packed vector
        # of the kernel outputs when the fused total produced one, else the stacked entries
        self.loss_vec = loss_dict.packed()
        if self.loss_vec is None:
            self.loss_vec = torch.stack([v.detach() for v in loss_dict.values()])
        self.last_outputs = outputs
        return total

    def h2d(self):
        for f, v in self.frames_u8.items():
            self.frames_dev[f].copy_(v, non_blocking=True)
        for k, v in self.host.items():
            self.buf[k].detach().copy_(v, non_blocking=True)

    def h2d_bytes(self):
        return sum(v.numel() for v in self.frames_u8.values()) + sum(v.numel() * v.element_size() for v in self.host.values())


def input_pipeline_bench(B, H, W, device, steps, cpu_baseline=True):
    """The on-GPU input pipeline on the TripleD item shape: 3 frames, colour jitter on every image,
    16 erase boxes of 16x16 (cfg_kitti_tripleD.py:19-20).  Device time by CUDA events over `steps` calls; the CPU figure
    is the reference's way -- torchvision ColorJitter on PIL images + ToTensor + the mask, one item after the other."""
    tdl = importlib.import_module(PKG)
    g = torch.Generator().manual_seed(77)
    frames = {f: (torch.rand(B, H, W, 3, generator=g) ** 2 * 255).to(torch.uint8) for f in FRAME_IDS}
    pipe = tdl.GpuInputPipeline(FRAME_IDS, H, W, erase_count=16, erase_shape=(16, 16))
    params = pipe.sample_params(B, g)
    params["do_aug"][:] = 1
    dev_frames = {f: v.to(device) for f, v in frames.items()}
    dev_params = {k: (v.to(device) if v is not None else None) for k, v in params.items()}
    for _ in range(3):
        pipe(dev_frames, dev_params)
    torch.cuda.synchronize(device)
    graph = torch.cuda.CUDAGraph()                       # (device time of the launches, not of the Python call)
    with torch.cuda.graph(graph, stream=torch.cuda.current_stream(device)):
        keep = pipe(dev_frames, dev_params)
    for _ in range(3):
        graph.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    e0.record()
    for _ in range(steps):
        graph.replay()
    e1.record()
    torch.cuda.synchronize(device)
    del keep
    us = e0.elapsed_time(e1) / steps * 1e3
    nf, n = len(FRAME_IDS), B * H * W
    alg = n * nf * (3 + 3 + 24) + n * 12            # statistics pass + apply pass reads, color + color_aug writes, mask
    out = {"us_per_batch": round(us, 2), "images_per_s": round(B / us * 1e6, 1), "alg_bytes": alg,
           "gbs": round(alg / us * 1e-3, 1), "launches": 2,
           "what": f"tdl_input_fwd: {nf} uint8 frames x batch {B} at {H}x{W} -> color, color_aug (ColorJitter on every image), "
                   "16 erase boxes; bit-exact against torchvision / Pillow (tests/test_input_pipeline.py)"}
    if cpu_baseline:
        try:
            from PIL import Image
            from torchvision import transforms
            cj = transforms.ColorJitter((0.8, 1.2), (0.8, 1.2), (0.8, 1.2), (-0.1, 0.1))
            tt = transforms.ToTensor()
            imgs = [[Image.fromarray(frames[f][b].numpy()) for f in FRAME_IDS] for b in range(B)]
            t0 = time.perf_counter()
            reps = 2
            for _ in range(reps):
                for b in range(B):
                    for im in imgs[b]:
                        tt(im)
                        tt(cj(im))
                    m = torch.ones(3, H, W, dtype=torch.uint8)
                    for _c in range(16):
                        r_ = int(torch.LongTensor(1).random_(0, H - 17)[0])
                        c_ = int(torch.LongTensor(1).random_(0, W - 17)[0])
                        m[:, r_:r_ + 16, c_:c_ + 16] = 0
            dt = (time.perf_counter() - t0) / reps
            out["cpu_baseline"] = {"images_per_s": round(B / dt, 1), "ms_per_batch": round(dt * 1e3, 1), "cores": 1, "kind": "reference",
                                   "sample": f"{reps} batches of {B} items: torchvision ColorJitter + ToTensor on PIL images and the "
                                             "mask loop of kitti_dataset.py:167-182, one DataLoader worker's share (single thread)"}
        except Exception as exc:                       # pragma: no cover - baseline only
            out["cpu_baseline"] = {"unavailable": f"{type(exc).__name__}: {exc}"}
    return out


def bind_host_near_gpu(local):
    """Best effort: run this process on the CPUs that are NUMA-local to its GPU (sysfs local_cpulist of the PCI
    device), so that the pinned staging buffers it first-touches sit on that node -- at 8 ranks the e2e figure is
    otherwise bound by cross-socket host memory traffic rather than by the path."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return sorted(allowed)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_record(B, H, W, seed, frames="smooth"):
    tdl = importlib.import_module(PKG)
    inputs, outputs, extras = tdl.synth.make_inputs(B, H, W, frame_ids=FRAME_IDS, scales=SCALES, seed=seed,
                                                    frames=frames, feat_channels=FEAT_C, with_noise=False)
    leaves = dict(outputs)
    leaves["tgt_feat"] = extras["tgt_feat"]
    for f, t in extras["src_feats"].items():
        leaves[("src_feat", f)] = t
    return {"inputs": inputs, "leaves": leaves,
            "meta": dict(kind="fm", B=B, H=H, W=W, frames=frames, C=FEAT_C, seed=seed, opt=opt_dict(B, H, W))}


def cpu_arm_kind():
    """"reference": the reference's own modules are reachable (oracle/_ref on the GPU box, /root/reference in the
    build container) and are what gets timed; "port": oracle/restatement.py (bit-identical restatement)."""
    try:
        from oracle import ref_loader
        return "reference" if ref_loader.reference_available() else "port"
    except Exception:
        return "port"


def cpu_step(rec, kind):
    """One forward + backward of the reference's mono_fm loss on the host cores, on fresh leaf tensors."""
    if kind == "reference":
        from oracle import ref_loader
        meta = rec["meta"]
        opt = ref_loader.default_opt(meta["B"], meta["H"], meta["W"])
        opt.update(meta["opt"])
        leaves = {k: v.clone().requires_grad_(True) for k, v in rec["leaves"].items()}
        outputs = {k: v for k, v in leaves.items() if isinstance(k, tuple) and k[0] in ("disp", "cam_T_cam")}
        src = {f: leaves[("src_feat", f)] for f in FRAME_IDS[1:]}
        loss = ref_loader.run_reference_loss("fm", opt, dict(rec["inputs"]), outputs, leaves["tgt_feat"], src)
    else:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from golden_util import run_restatement
        loss, _, _ = run_restatement(rec)
    total = sum(v.mean() for v in loss.values())         # batch_processor (mono/apis/trainer.py:39-48)
    total.backward()
    return float(total.detach())


def time_cpu_arm(B, H, W, steps, warmup, kind, budget_s=None):
    """-> (images/s, s/step, batch actually used).  With a budget the per-step sample (batch) is halved until
    (steps + warmup) steps fit, judged from the first warm-up step."""
    torch.set_num_threads(os.cpu_count() or 1)
    import warnings
    warnings.filterwarnings("ignore", message="Default grid_sample")
    while True:
        rec = cpu_record(B, H, W, 4321)
        t0 = time.perf_counter()
        cpu_step(rec, kind)
        first = time.perf_counter() - t0
        if budget_s is None or B == 1 or first * (steps + warmup) <= budget_s:
            break
        B = max(1, B // 2)
    for _ in range(max(0, warmup - 1)):
        cpu_step(rec, kind)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(rec, kind)
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps, B


# ------------------------------------------------------------------------------------------------ train step
def train_step_bench(args, device, rank, world, dist_on, model_kind):
    """Full training step of cfg_kitti_fm (model_kind "mono_fm": ResNet-50 depth net, ResNet-18 pose net, ResNet-50
    extractor) or cfg_kitti_tripleD ("tripled": + in-loop encoder, image / colour decoders, erase mask) with the fused
    loss, Adam 1e-4, grad-clip 35, batch-sharded over the ranks.  The networks are plain PyTorch / cuDNN in
    channels_last (out of the hot path's scope); the step itself is package train_step.FlatGradStep: flat gradient
    buffer, ONE NCCL all-reduce, fused clip, fused Adam, everything captured in one CUDA graph at every rank count.
    Also timed in the same job: the identical step WITHOUT the collective (what one GPU alone does), which gives
    `efficiency_vs_n1` = local step time / distributed step time."""
    tdl = importlib.import_module(PKG)
    importlib.import_module(PKG + ".nets")
    ts = importlib.import_module(PKG + ".train_step")
    B, H, W = args.batch, args.height, args.width
    opt = opt_dict(B, H, W)
    tripled = model_kind == "tripled"
    name = "mono_fm_joint_inpaint_disentangle" if tripled else "mono_fm"
    opt.update(name=name, depth_num_layers=50, pose_num_layers=18, extractor_num_layers=50,
               extractor_pretrained_path=None, dis=1e-3, cvt=1e-3, auto_res_weight=5e-3, freeze_extractor=False)
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    host = make_host_workload(B, H, W, rank_seed(rank), "scene")
    inputs = {}
    for k, v in host.items():
        if k[0] == "in":
            inputs[k[1]] = v.to(device)
    for f in FRAME_IDS:
        inputs[("color_aug", f, 0)] = inputs[("color", f, 0)]
    if tripled:                                     # 16 erased 16x16 holes (cfg_kitti_tripleD.py:19-20)
        g = torch.Generator().manual_seed(rank_seed(rank))
        mask = torch.ones(B, 3, H, W)
        for _ in range(16):
            y0, x0 = int(torch.randint(0, H - 16, (1,), generator=g)), int(torch.randint(0, W - 16, (1,), generator=g))
            mask[:, :, y0:y0 + 16, x0:x0 + 16] = 0
        inputs[("mask", 0, 0)] = mask.to(device)

    def build(world_for_step):
        torch.manual_seed(0)
        model = tdl.MONO.module_dict[name](tdl.config.ConfigDict(opt)).to(device).train()
        if not args.no_channels_last:
            model = model.to(memory_format=torch.channels_last)
        stepper = ts.FlatGradStep(model, lr=1e-4, max_norm=35.0, world=world_for_step, bf16_comm=args.bf16_comm,
                                  fused_adam=not args.no_fused_adam)
        stepper.broadcast_parameters()

        def loss_fn():
            _, loss_dict = model(inputs)
            return loss_dict.total()
        return model, stepper, loss_fn

    def measure(world_for_step, sync_ranks):
        model, stepper, loss_fn = build(world_for_step)
        mode = "cuda-graph"
        try:
            if args.eager_train:
                raise RuntimeError("--eager-train")
            stepper.capture(loss_fn, warmup=3)
            run = stepper.replay
        except Exception as exc:
            if not args.eager_train:
                print(f"train-step graph capture failed ({type(exc).__name__}: {exc}); timing the eager step", file=sys.stderr)
            torch.cuda.synchronize(device)
            mode = "eager"

            def run():
                return stepper.step(loss_fn)
        for _ in range(3):
            loss = run()
        ms = timed_region(run, args.train_steps, device, sync_ranks)
        loss = run()
        out = (ms / args.train_steps, mode, float(loss.detach()), stepper.n_params())
        del model, stepper
        torch.cuda.empty_cache()
        return out

    ms_step, mode, final_loss, n_params = measure(world, dist_on)
    local_ms = None
    if dist_on:
        local_ms, _, _, _ = measure(1, dist_on)       # the same step without the collective, on every rank at the same time
    res = {"images_per_s": round(world * B / (ms_step * 1e-3), 1), "ms_per_step": round(ms_step, 3), "steps": args.train_steps,
           "execution": mode, "data": "scene",
           "model": ("TripleD mono_fm_joint_inpaint_disentangle (cfg_kitti_tripleD): ResNet-50 depth + ResNet-18 pose + "
                     "ResNet-50 in-loop encoder + image / colour decoders, " if tripled else
                     "mono_fm (cfg_kitti_fm): ResNet-50 depth + ResNet-18 pose + ResNet-50 extractor (level 0), ")
                    + f"{n_params / 1e6:.1f} M trainable params, PyTorch fp32 networks"
                    + ("" if args.no_channels_last else " (channels_last)") + " + fused loss, fused clip + Adam",
           "parallelism": f"batch-sharded x{world}: one flat NCCL all-reduce of {n_params * (2 if args.bf16_comm else 4) / 1e6:.0f} MB "
                          f"({'bf16' if args.bf16_comm else 'fp32'}) per step inside the graph, per-rank BatchNorm",
           "final_loss": final_loss}
    if local_ms is not None:
        res["local_ms_per_step"] = round(local_ms, 3)
        res["efficiency_vs_n1"] = round(local_ms / ms_step, 4)
    return res


def time_eager_gpu_port(B, H, W, device, steps=5, warmup=2):
    """The reference's loss as stock eager PyTorch ops on this same GPU (the oracle port run on CUDA tensors,
    ~150-200 kernel launches per scale): a reported baseline for the fused kernels, like the CPU port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import run_restatement
    rec = cpu_record(B, H, W, 4321)
    rec = {"inputs": {k: v.to(device) for k, v in rec["inputs"].items()},
           "leaves": {k: v.to(device) for k, v in rec["leaves"].items()}, "meta": rec["meta"]}
    import oracle.restatement as R
    draw = R.draw_automask_noise

    def draw_dev(spec, batch, generator=None, dtype=torch.float32):       # the reference: CPU randn + .cuda()
        return {s: {f: n.to(device) for f, n in d.items()} for s, d in draw(spec, batch, generator, dtype).items()}
    R.draw_automask_noise = draw_dev
    try:
        def one():
            loss, _, _ = run_restatement(rec)
            sum(loss.values()).backward()
        for _ in range(warmup):
            one()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        torch.cuda.synchronize(device)
        dt = (time.perf_counter() - t0) / steps
    finally:
        R.draw_automask_noise = draw
    return B / dt, dt


def config_dict(args, trainable):
    return {"workload": f"mono_fm loss fwd+bwd (cfg_kitti_fm): {args.height}x{args.width}, batch {args.batch}/GPU, "
                        f"frames [0,-1,1], 4 scales, {FEAT_C}-ch features at H/2xW/2, "
                        f"{'trainable' if trainable else 'frozen'} extractor features in {FEAT_LAYOUT.upper()} memory, reference-faithful outputs "
                        "(warped images/features + int64 min_index materialised), in-kernel Philox automask noise; "
                        "value = 'smooth' synthetic frames (SURVEY 8d), `scene` block = rendered moving-camera frames",
            "global_batch": args.batch * args.gpus, "per_gpu_batch": args.batch,
            "height": args.height, "width": args.width, "parallelism": f"dp{args.gpus} (batch-sharded, no collective)",
            "l2_policy": "per-step working set (~0.6 GB at batch 8) exceeds the 126 MB L2; inputs are re-streamed "
                         "from HBM every step",
            "execution": "CUDA-graph replay of the public compute_losses_fm + backward; e2e double-buffers the H2D upload of step i+1 behind step i"}


class _QuietStdout:
    """Routes the process's file descriptor 1 to stderr while libraries (NCCL prints its version banner on
    stdout) are active, so that the ONE JSON line is the only thing the benchmark writes to stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    with _QuietStdout():
        line = run()
    if line is not None:
        print(json.dumps(line), flush=True)


def run():
    global SEED_OFFSET, FEAT_LAYOUT
    args = parse()
    SEED_OFFSET = args.seed_offset
    FEAT_LAYOUT = args.feat_layout
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    trainable = not args.frozen_extractor

    if args.impl == "reference":
        if rank != 0:
            return
        kind = cpu_arm_kind()
        steps, warm = max(1, args.steps), max(0, args.warmup)
        ips, spstep, used_b = time_cpu_arm(args.batch, args.height, args.width, steps, warm, kind, args.ref_budget_s)
        cores = os.cpu_count() or 1
        what = ("the UNMODIFIED reference modules (oracle/_ref: mono/model/mono_fm/net.py compute_losses + autograd backward)"
                if kind == "reference" else "oracle/restatement.py (bit-identical port; the reference modules did not travel)")
        cfg = config_dict(args, trainable)
        cfg["reference_arm"] = (f"batch {used_b} per step (asked {args.batch}), {steps} timed steps after {warm} warm-ups, "
                                f"torch CPU ops on {cores} threads, features supplied as tensors like the GPU arm")
        line = {"impl": "reference", "metric": "fused_loss_fwd_bwd_images_per_s", "value": ips, "unit": "images/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": spstep * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg,
                "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": kind,
                                 "sample": f"batch {used_b} of the same {args.height}x{args.width} workload per step, "
                                           f"{steps} steps after {warm} warm-ups; {what}"},
                "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        return line

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the fused loss has no CPU path (use --impl reference for the CPU arm)")
    dist_on = world > 1
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    if dist_on:
        torch.distributed.init_process_group("nccl", device_id=device)
    tdl = importlib.import_module(PKG)
    tdl._lib.lib()                                   # fail loudly if libtdl.so is missing

    B, H, W = args.batch, args.height, args.width
    bind_host_near_gpu(local)                # pinned staging buffers are first touched by this process: keep it NUMA-local
    side = torch.cuda.Stream(device)
    torch.cuda.set_stream(side)              # every launch, copy and timing event below is on this stream
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                      # covers the warm-up replays and every timed region of both workloads
    smooth = measure_workload("smooth", args, device, rank, world, dist_on, trainable)
    scene = None if args.no_scene else measure_workload("scene", args, device, rank, world, dist_on, trainable)
    clocks = sampler.stop() if rank == 0 else None
    # opt-in storage mode (north_star "bf16/fp32 loads"): feature maps stored bf16, fp32 arithmetic; never the parity config
    scene_bf16 = None
    if not (args.no_scene or args.no_bf16):
        scene_bf16 = measure_workload("scene", args, device, rank, world, dist_on, trainable, torch.bfloat16)
        del scene_bf16["step"], scene_bf16["host"]
    step, host = smooth["step"], smooth["host"]
    ms_step = smooth["ms_per_step"]
    value = world * B / (ms_step * 1e-3)

    # ---- e2e: host (pinned) buffers -> H2D of every input of the step -> step -> D2H of the loss scalars, with a
    #      host synchronisation per step like the reference's per-iteration .item() (mono/apis/trainer.py:52-54).
    #      Two device buffer sets: the copy stream uploads step i+1 while the compute stream runs step i.
    step2 = DeviceStep(host, B, H, W, device, trainable)
    for _ in range(3):
        step2.run_eager()
    step2.capture()
    e2e_steps = min(args.steps, 40)
    ms_e2e = timed_region(lambda: e2e_loop([step, step2], e2e_steps, device), 1, device, dist_on, warm=lambda: e2e_loop([step, step2], 4, device))
    e2e_value = whole_job_images_per_s(world, B, e2e_steps, ms_e2e)
    del step2

    # ---- e2e_images_only: what a data loader hands over (uint8 frames, K, inv_K) is uploaded; disparities / poses are
    #      resident stand-ins for the network outputs, the features come from a conv stem ON THE DEVICE
    io_a, io_b = ImagesOnlyStep(host, B, H, W, device), ImagesOnlyStep(host, B, H, W, device)
    for st_ in (io_a, io_b):
        for _ in range(3):
            st_.run_eager()
        st_.capture()
    ms_io = timed_region(lambda: e2e_loop([io_a, io_b], e2e_steps, device), 1, device, dist_on, warm=lambda: e2e_loop([io_a, io_b], 4, device))
    io_value = whole_job_images_per_s(world, B, e2e_steps, ms_io)
    io_bytes, io_d2h = io_a.h2d_bytes(), io_a.losses_host.numel() * 4
    del io_a, io_b

    input_pipe = input_pipeline_bench(B, H, W, device, max(args.steps, 10), cpu_baseline=(rank == 0 and not args.no_cpu_baseline))

    train = train_tripled = None
    if not args.no_train:
        torch.cuda.empty_cache()
        if args.train_model in ("both", "mono_fm"):
            train = train_step_bench(args, device, rank, world, dist_on, "mono_fm")
        if args.train_model in ("both", "tripled"):
            train_tripled = train_step_bench(args, device, rank, world, dist_on, "tripled")

    if rank != 0:
        if dist_on:
            torch.distributed.destroy_process_group()
        return

    # ---- roofline: dominant kernel of the representative ("scene") workload
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    S = len(FRAME_IDS) - 1
    alg = algorithmic_bytes(B, H, W, S, FEAT_C, len(SCALES), trainable)
    total_alg = sum(alg[k] for k in ("photo_fwd", "photo_bwd", "feat_fwd", "feat_bwd", "memset_dsrc"))

    alg_bf16 = dict(alg)
    for k in ("feat_fwd", "feat_bwd", "memset_dsrc"):
        alg_bf16[k] = alg[k] // 2

    def kernel_table(m, alg=alg):
        kernels = {}
        for name, (n, total_ms) in m["kern"].items():
            avg_us = total_ms / n * 1e3
            row = {"launches_per_step": n / m["prof_steps"], "avg_us": round(avg_us, 2),
                   "us_per_step": round(total_ms / m["prof_steps"] * 1e3, 2)}
            if name in alg and alg[name]:
                row["alg_bytes"] = alg[name]
                row["gbs"] = round(alg[name] / (avg_us * 1e-6) / 1e9, 1)
            kernels[name] = row
        tot_us = sum(r["us_per_step"] for r in kernels.values())
        for r in kernels.values():
            r["share"] = round(r["us_per_step"] / tot_us, 3)
        return kernels

    def whole_step(m):
        gbs = total_alg / (m["ms_per_step"] * 1e-3) / 1e9
        return {"alg_bytes": total_alg, "gbs": round(gbs, 1), "frac": round(gbs / peak, 4)}

    k_smooth = kernel_table(smooth)
    dom_src, dom_tab = ("scene", kernel_table(scene)) if scene else ("smooth", k_smooth)
    dom = max((k for k in dom_tab if "alg_bytes" in dom_tab[k]), key=lambda k: dom_tab[k]["us_per_step"])
    achieved = dom_tab[dom]["gbs"]
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):       # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed
        tj = json.load(open(tpath))  # `ncu --set full` capture of this same default workload
        if tj.get("workload") == [B, H, W, S, FEAT_C] and dom in tj.get(dom_src, tj).get("kernels", {}):
            traffic = tj.get(dom_src, tj)["kernels"][dom]
    roofline = {"bound": "hbm", "kernel": dom, "workload": dom_src, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": alg[dom],
                "whole_step": whole_step(scene) if scene else whole_step(smooth),
                "whole_step_smooth": whole_step(smooth),
                "kernels": dom_tab}

    def block(m, kernels):
        return {"images_per_s": round(world * B / (m["ms_per_step"] * 1e-3), 1), "ms_per_step": round(m["ms_per_step"], 4),
                "repeats": m["repeats"], "identity_frac": m["identity_frac"], "launches_per_step": m["launches_per_step"],
                "kernels": kernels}

    cpu = eager = None
    if world == 1 and not args.no_cpu_baseline:
        kind = cpu_arm_kind()
        ips, spstep, used_b = time_cpu_arm(args.cpu_batch, H, W, 8, 1, kind)
        cpu = {"value": round(ips, 3), "unit": "images/s", "cores": os.cpu_count() or 1, "kind": kind,
               "sample": f"batch {used_b} of the same {H}x{W} workload, 8 timed steps after 1 warm-up "
                         f"({spstep:.2f} s/step), " + ("the unmodified reference modules (oracle/_ref)" if kind == "reference"
                                                         else "oracle/restatement.py") + " on all host threads"}
        try:
            ips, spstep = time_eager_gpu_port(B, H, W, device)
            eager = {"value": round(ips, 1), "unit": "images/s", "ms_per_step": round(spstep * 1e3, 2), "kind": "port",
                     "what": "the reference's loss op sequence (oracle/restatement.py) as eager PyTorch on this GPU, "
                             f"batch {B}, host-timed with synchronisation, 5 steps after 2 warm-ups"}
        except Exception as exc:                       # pragma: no cover - baseline only
            eager = {"unavailable": f"{type(exc).__name__}: {exc}"}

    line = {"metric": "fused_loss_fwd_bwd_images_per_s", "value": round(value, 1), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms_step, 4),
            "repeats": smooth["repeats"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, trainable), "clocks": clocks,
            "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "h2d_bytes_per_step": step.h2d_bytes(),
                    "d2h_bytes_per_step": step.losses_host.numel() * 4, "ms_per_step": round(ms_e2e / e2e_steps, 4),
                    "steps": e2e_steps},
            "e2e_images_only": {"value": round(io_value, 1), "unit": "images/s", "h2d_bytes_per_step": io_bytes,
                                "d2h_bytes_per_step": io_d2h, "ms_per_step": round(ms_io / e2e_steps, 4), "steps": e2e_steps,
                                "what": "uint8 frames + K + inv_K uploaded per step; on-GPU input pipeline (tdl_input_fwd: to_tensor + "
                                        "ColorJitter), a trainable 7x7/2 conv + ReLU feature stem on the augmented frames (3 frames), "
                                        "loss fwd + bwd through the stem, on the device"},
            "input_pipeline": input_pipe,
            "gpu_launches": smooth["launches_per_step"] * args.steps * smooth["repeats"]["n"],
            "gpu_launches_per_step": smooth["launches_per_step"],
            "smooth": block(smooth, k_smooth), "scene": block(scene, dom_tab) if scene else None,
            "scene_bf16_features": (dict(block(scene_bf16, kernel_table(scene_bf16, alg_bf16)),
                                         what="opt-in: 64-channel feature maps stored in bf16 (half the feature bytes), fp32 arithmetic; "
                                              "images, disparities and every loss / gradient of them stay fp32",
                                         alg_bytes=sum(alg_bf16[k] for k in ("photo_fwd", "photo_bwd", "feat_fwd", "feat_bwd", "memset_dsrc")))
                                    if scene_bf16 else None),
            "roofline": roofline, "cpu_baseline": cpu, "eager_torch_gpu_baseline": eager,
            "train_step": train, "train_step_tripled": train_tripled}
    if dist_on:
        torch.distributed.destroy_process_group()
    return line


if __name__ == "__main__":
    main()
