#!/usr/bin/env python
"""Benchmark of the fused view-synthesis loss (forward + backward) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], cfg_kitti_fm): the mono_fm loss -- 4-scale photometric
(SSIM + L1, automask, min over 2 source frames) + feature-metric on 64-channel features at
(H/2, W/2) + edge-aware smoothness -- at 192x640, batch 8 per GPU, fp32, forward + backward,
on synthetic frames (package synth.py).  One "step" = one fwd+bwd pass over one batch.

Printed JSON line (rank 0):
  value      images/s with every input already resident in HBM (CUDA-graph replay of the step)
  e2e        images/s through the public host API with HOST (pinned) buffers: H2D of every input
             of the step, the step, D2H of the loss scalars -- inside the timed region
  roofline   dominant kernel: algorithmic bytes per launch / CUDA-event duration vs MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (oracle/restatement.py, a port of the reference's loss code) timed on
             this box's host cores on a bounded sample (B=2) of the same workload
With --impl reference the reference arm is timed instead: the reference's loss code path on the
host cores (the oracle port; /root/reference does not exist on the GPU box).
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
    ap.add_argument("--train-model", default="mono_fm", choices=["mono_fm", "tripled"],
                    help="mono_fm = cfg_kitti_fm (BASELINE configs[1]); tripled = cfg_kitti_tripleD (configs[2])")
    ap.add_argument("--syncbn", action="store_true",
                    help="convert to SyncBatchNorm like cfg_kitti_fm (syncbn = True); off by default: its ~200 tiny "
                         "collectives per step halve 2-GPU throughput (measured: 186 vs 394 images/s)")
    ap.add_argument("--cpu-batch", type=int, default=2, help="batch of the bounded CPU sample")
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
def make_host_workload(B, H, W, seed):
    tdl = importlib.import_module(PKG)
    inputs, outputs, extras = tdl.synth.make_inputs(B, H, W, frame_ids=FRAME_IDS, scales=SCALES, seed=seed,
                                                    frames="smooth", feat_channels=FEAT_C, with_noise=False)
    host = {}
    for k, v in inputs.items():
        host[("in", k)] = v
    for k, v in outputs.items():
        host[("leaf", k)] = v
    host[("leaf", "tgt_feat")] = extras["tgt_feat"]
    for f, t in extras["src_feats"].items():
        host[("leaf", ("src_feat", f))] = t
    return host


class DeviceStep:
    """The loss step on static device buffers: eager, or captured once into a CUDA graph and replayed."""

    def __init__(self, host, B, H, W, device, trainable_feat):
        tdl = importlib.import_module(PKG)
        self.tdl, self.device = tdl, device
        self.buf = {k: v.to(device).contiguous() for k, v in host.items()}
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
        total.backward()
        self.loss_vec = torch.stack([v.detach() for v in loss_dict.values()])
        return total

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


def timed_region(fn, steps, device, dist_on):
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


# ------------------------------------------------------------------------------------------------ CPU arms
def cpu_port_step(rec_builder):
    """One fwd+bwd of the CPU oracle (the port of the reference's mono_fm loss) on a fresh copy of the inputs."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from golden_util import run_restatement
    rec = rec_builder()
    loss, _, _ = run_restatement(rec)
    sum(loss.values()).backward()
    return float(sum(v.detach() for v in loss.values()))


def cpu_record(B, H, W, seed):
    tdl = importlib.import_module(PKG)
    inputs, outputs, extras = tdl.synth.make_inputs(B, H, W, frame_ids=FRAME_IDS, scales=SCALES, seed=seed,
                                                    frames="smooth", feat_channels=FEAT_C, with_noise=False)
    leaves = dict(outputs)
    leaves["tgt_feat"] = extras["tgt_feat"]
    for f, t in extras["src_feats"].items():
        leaves[("src_feat", f)] = t
    return {"inputs": inputs, "leaves": leaves,
            "meta": dict(kind="fm", B=B, H=H, W=W, frames="smooth", C=FEAT_C, seed=seed, opt=opt_dict(B, H, W))}


def time_cpu_port(B, H, W, steps, warmup):
    torch.set_num_threads(os.cpu_count() or 1)
    rec = cpu_record(B, H, W, 4321)
    for _ in range(warmup):
        cpu_port_step(lambda: rec)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_port_step(lambda: rec)
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps


# ------------------------------------------------------------------------------------------------ train step
def train_step_bench(args, device, rank, world, dist_on):
    """Full cfg_kitti_fm training step (config/cfg_kitti_fm.py: ResNet-50 depth net, ResNet-18 pose net, ResNet-50
    extractor, Adam 1e-4, grad-clip 35, syncbn) with the fused loss, batch-sharded DDP over NCCL when world > 1.
    The networks are plain PyTorch (out of the hot path's scope); reported next to the loss-only figure."""
    tdl = importlib.import_module(PKG)
    importlib.import_module(PKG + ".nets")
    B, H, W = args.batch, args.height, args.width
    opt = opt_dict(B, H, W)
    tripled = args.train_model == "tripled"
    name = "mono_fm_joint_inpaint_disentangle" if tripled else "mono_fm"
    opt.update(name=name, depth_num_layers=50, pose_num_layers=18, extractor_num_layers=50,
               extractor_pretrained_path=None, dis=1e-3, cvt=1e-3, auto_res_weight=5e-3, freeze_extractor=False)
    torch.manual_seed(0)
    torch.backends.cudnn.benchmark = True
    model = tdl.MONO.module_dict[name](tdl.config.ConfigDict(opt)).to(device).train()
    n_params = sum(q.numel() for q in model.parameters() if q.requires_grad)
    syncbn = dist_on and args.syncbn
    if dist_on:
        if syncbn:
            model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
        model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[device.index], gradient_as_bucket_view=True)
    optim = torch.optim.Adam([q for q in model.parameters() if q.requires_grad], lr=1e-4, weight_decay=0,
                             capturable=not dist_on and not args.eager_train, fused=None if args.no_fused_adam else True)
    host = make_host_workload(B, H, W, rank_seed(rank))
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

    def step():
        optim.zero_grad(set_to_none=True)
        _, loss_dict = model(inputs)
        loss = loss_dict.total()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 35.0)
        optim.step()
        return loss

    for _ in range(3):
        loss = step()
    torch.cuda.synchronize(device)
    # The eager step is CPU-launch-bound (~1500 small kernels); capture forward + backward + clip + Adam into one
    # CUDA graph when possible (single process; DDP's bucketed NCCL all-reduce stays eager).
    mode = "eager"
    run = step
    if not dist_on and not args.eager_train:
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=torch.cuda.current_stream(device)):
                static_loss = step()
            torch.cuda.synchronize(device)

            def run():
                graph.replay()
                return static_loss
            mode = "cuda-graph"
            for _ in range(2):
                loss = run()
        except Exception as exc:                      # pragma: no cover - falls back to the eager step
            print(f"train-step graph capture failed ({type(exc).__name__}: {exc}); timing the eager step", file=sys.stderr)
            torch.cuda.synchronize(device)
            run, mode = step, "eager"
    ms = timed_region(run, args.train_steps, device, dist_on)
    loss = run()
    return {"images_per_s": round(whole_job_images_per_s(world, B, args.train_steps, ms), 1),
            "ms_per_step": round(ms / args.train_steps, 3), "steps": args.train_steps, "execution": mode,
            "model": ("TripleD mono_fm_joint_inpaint_disentangle (cfg_kitti_tripleD): ResNet-50 depth + ResNet-18 pose + "
                      "ResNet-50 in-loop encoder + image / colour decoders, " if tripled else
                      "mono_fm (cfg_kitti_fm): ResNet-50 depth + ResNet-18 pose + ResNet-50 extractor (level 0), ")
                     + f"{n_params / 1e6:.1f} M trainable params, PyTorch fp32 networks + fused loss, Adam",
            "parallelism": f"DDP x{world} (NCCL all-reduce of gradients" + (", SyncBatchNorm)" if syncbn else ")"),
            "final_loss": float(loss.detach())}


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
                        f"{'trainable' if trainable else 'frozen'} extractor features, reference-faithful outputs "
                        "(warped images/features + int64 min_index materialised), in-kernel Philox automask noise",
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
    global SEED_OFFSET
    args = parse()
    SEED_OFFSET = args.seed_offset
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    trainable = not args.frozen_extractor

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, min(args.steps, 5)), max(1, min(args.warmup, 2))
        ips, spstep = time_cpu_port(args.cpu_batch, args.height, args.width, steps, warm)
        cores = os.cpu_count() or 1
        line = {"impl": "reference", "metric": "fused_loss_fwd_bwd_images_per_s", "value": ips, "unit": "images/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": spstep * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_dict(args, trainable),
                "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                                 "sample": f"batch {args.cpu_batch} of the same {args.height}x{args.width} workload per "
                                           f"step, {steps} steps, torch CPU ops on {cores} threads "
                                           "(oracle/restatement.py; /root/reference is absent on the GPU box)"},
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
    host = make_host_workload(B, H, W, rank_seed(rank))
    side = torch.cuda.Stream(device)
    torch.cuda.set_stream(side)              # every launch, copy and timing event below is on this stream
    step = DeviceStep(host, B, H, W, device, trainable)

    # ---- per-kernel device time (CUDA events on the launch stream), eager, before the graph is built
    for _ in range(3):
        step.run_eager()
    torch.cuda.synchronize(device)
    prof_steps = min(args.steps, 20)
    tdl._lib.profile_begin()
    for _ in range(prof_steps):
        step.run_eager()
    kern = tdl._lib.profile_end()
    launches_per_step = sum(n for k, (n, _) in kern.items() if not k.startswith("memset")) // prof_steps

    # ---- value: device-resident inputs, CUDA-graph replay
    step.capture()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # samples cover the warm-up replays (same load) and the timed region
    for _ in range(max(3, args.warmup)):
        step.replay()
    ms = timed_region(step.replay, args.steps, device, dist_on)
    clocks = sampler.stop() if rank == 0 else None
    value = whole_job_images_per_s(world, B, args.steps, ms)

    # ---- e2e: host (pinned) buffers -> H2D of every input of the step -> step -> D2H of the loss scalars, with a
    #      host synchronisation per step like the reference's per-iteration .item() (mono/apis/trainer.py:52-54).
    #      Two device buffer sets: the copy stream uploads step i+1 while the compute stream runs step i.
    step2 = DeviceStep(host, B, H, W, device, trainable)
    for _ in range(3):
        step2.run_eager()
    step2.capture()
    sets = [step, step2]
    copy_stream = torch.cuda.Stream(device)
    ev_copied = [torch.cuda.Event(), torch.cuda.Event()]
    ev_done = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_loop(n):
        main = torch.cuda.current_stream(device)
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

    e2e_loop(4)
    e2e_steps = min(args.steps, 40)
    ms_e2e = timed_region(lambda: e2e_loop(e2e_steps), 1, device, dist_on)
    e2e_value = whole_job_images_per_s(world, B, e2e_steps, ms_e2e)

    train = None
    if not args.no_train:
        del step2, sets
        torch.cuda.empty_cache()
        train = train_step_bench(args, device, rank, world, dist_on)

    if rank != 0:
        if dist_on:
            torch.distributed.destroy_process_group()
        return

    # ---- roofline of the dominant kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    S = len(FRAME_IDS) - 1
    alg = algorithmic_bytes(B, H, W, S, FEAT_C, len(SCALES), trainable)
    kernels = {}
    for name, (n, total_ms) in kern.items():
        avg_us = total_ms / n * 1e3
        per_step_us = total_ms / prof_steps * 1e3
        row = {"launches_per_step": n / prof_steps, "avg_us": round(avg_us, 2), "us_per_step": round(per_step_us, 2)}
        if name in alg and alg[name]:
            row["alg_bytes"] = alg[name]
            row["gbs"] = round(alg[name] / (avg_us * 1e-6) / 1e9, 1)
        kernels[name] = row
    tot_us = sum(r["us_per_step"] for r in kernels.values())
    for r in kernels.values():
        r["share"] = round(r["us_per_step"] / tot_us, 3)
    dom = max((k for k in kernels if "alg_bytes" in kernels[k]), key=lambda k: kernels[k]["us_per_step"])
    achieved = kernels[dom]["gbs"]
    total_alg = sum(alg[k] for k in ("photo_fwd", "photo_bwd", "feat_fwd", "feat_bwd", "memset_dsrc"))
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):       # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed
        tj = json.load(open(tpath))  # `ncu --set full` capture of this same default workload
        if tj.get("workload") == [B, H, W, S, FEAT_C] and dom in tj.get("kernels", {}):
            traffic = tj["kernels"][dom]
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": alg[dom],
                "whole_step": {"alg_bytes": total_alg, "gbs": round(total_alg / (ms / args.steps * 1e-3) / 1e9, 1),
                               "frac": round(total_alg / (ms / args.steps * 1e-3) / 1e9 / peak, 4)},
                "kernels": kernels}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        ips, spstep = time_cpu_port(args.cpu_batch, H, W, 8, 1)
        cpu = {"value": round(ips, 3), "unit": "images/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"batch {args.cpu_batch} of the same {H}x{W} workload, 8 timed steps after 1 warm-up "
                         f"({spstep:.2f} s/step), oracle/restatement.py on all host threads"}

    eager = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            ips, spstep = time_eager_gpu_port(B, H, W, device)
            eager = {"value": round(ips, 1), "unit": "images/s", "ms_per_step": round(spstep * 1e3, 2), "kind": "port",
                     "what": "the reference's loss op sequence (oracle/restatement.py) as eager PyTorch on this GPU, "
                             f"batch {B}, host-timed with synchronisation, 5 steps after 2 warm-ups"}
        except Exception as exc:                       # pragma: no cover - baseline only
            eager = {"unavailable": f"{type(exc).__name__}: {exc}"}

    line = {"metric": "fused_loss_fwd_bwd_images_per_s", "value": round(value, 1), "unit": "images/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms / args.steps, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, trainable), "clocks": clocks,
            "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "h2d_bytes_per_step": step.h2d_bytes(),
                    "d2h_bytes_per_step": step.losses_host.numel() * 4, "ms_per_step": round(ms_e2e / e2e_steps, 4),
                    "steps": e2e_steps},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "eager_torch_gpu_baseline": eager,
            "train_step": train}
    if dist_on:
        torch.distributed.destroy_process_group()
    return line


if __name__ == "__main__":
    main()
