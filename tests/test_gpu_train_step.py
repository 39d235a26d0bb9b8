"""GPU: the batch-sharded data-parallel step (package train_step.FlatGradStep) --
  * world 1: flat-buffer gradients + fused clip + fused Adam == the textbook step (clip_grad_norm_ + Adam);
  * world 2: gradients averaged over two ranks (half the batch each) == single-GPU gradients on the concatenated batch,
    the invariant the reference's DDP training relies on (mono/apis/trainer.py:147-189,
    mono/core/utils/dist_utils.py:34-60).  NCCL over two GPUs when the box has them, else two gloo ranks on one GPU.
"""
import importlib
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from gpu_util import PKG, pkg

pytestmark = pytest.mark.gpu
H, W = 96, 128


def _cfg(tdl, B):
    return tdl.config.ConfigDict(dict(
        name="mono_fm", depth_num_layers=18, pose_num_layers=18, extractor_num_layers=18, frame_ids=[0, -1, 1],
        imgs_per_gpu=B, height=H, width=W, scales=[0, 1, 2, 3], min_depth=0.1, max_depth=100.0,
        automask=False,        # (the in-kernel tie-break noise is indexed by the LOCAL image number: keep it out of the comparison)
        disp_norm=True, perception_weight=1e-3, smoothness_weight=1e-3, extractor_pretrained_path=None))


def _model(tdl, B, dev, channels_last):
    importlib.import_module(PKG + ".nets")
    torch.manual_seed(0)
    m = tdl.MONO.module_dict["mono_fm"](_cfg(tdl, B)).to(dev).train()
    for mod in m.modules():          # batch statistics / dropout masks depend on how the batch is split: freeze them
        if isinstance(mod, (torch.nn.BatchNorm2d, torch.nn.Dropout)):
            mod.eval()
    return m.to(memory_format=torch.channels_last) if channels_last else m


def _inputs(tdl, B, dev, lo=0, hi=None):
    inputs, _, _ = tdl.synth.make_inputs(B, H, W, seed=21, with_noise=False, frames="scene")
    inputs = {k: v[lo:hi].to(dev).contiguous() for k, v in inputs.items()}
    for f in (0, -1, 1):
        inputs[("color_aug", f, 0)] = inputs[("color", f, 0)]
    return inputs


@pytest.mark.parametrize("channels_last", [False, True])
def test_flat_step_equals_textbook_step(channels_last):
    """(a) gradients land in the flat buffer and the fused clip equals clip_grad_norm_; (b) the fused Adam on the flat
    views updates every parameter like a plain Adam given the same gradients (element pairing under channels_last)."""
    tdl = pkg()
    ts = importlib.import_module(PKG + ".train_step")
    dev = "cuda"
    inputs = _inputs(tdl, 2, dev)
    a, b = _model(tdl, 2, dev, channels_last), _model(tdl, 2, dev, channels_last)
    step = ts.FlatGradStep(a, lr=1e-3, max_norm=0.05, world=1, capturable=False)       # max_norm small enough to clip
    pb = [q for q in b.parameters() if q.requires_grad]
    opt = torch.optim.Adam(pb, lr=1e-3)
    # (a)
    step.flat_grad.zero_()
    a(inputs)[1].total().backward()
    step.reduce_and_clip()
    b(inputs)[1].total().backward()
    norm = torch.nn.utils.clip_grad_norm_(pb, 0.05)
    assert float(norm) > 0.05                                   # the clip was active
    assert abs(float(step.grad_norm) - float(norm)) <= 1e-4 * float(norm)
    err2 = 0.0
    for qa, qb in zip(step.params, pb):
        assert qa.grad.data_ptr() >= step.flat_grad.data_ptr()  # still a view of the flat buffer (accumulated in place)
        err2 += float((qa.grad - qb.grad).double().pow(2).sum())
    # (the backward's atomics are not run-to-run deterministic: compare the whole gradient, not parameter by parameter)
    assert err2 ** 0.5 <= 1e-4 * float(step.flat_grad.norm()), err2 ** 0.5
    # (b) identical, well-conditioned gradients on both sides
    g = torch.Generator(device=dev).manual_seed(3)
    for qa, qb in zip(step.params, pb):
        r = torch.randn(qa.shape, generator=g, device=dev) + 3.0
        qa.grad.copy_(r)
        qb.grad.copy_(r)
    for _ in range(2):
        step.optim.step()
        opt.step()
    worst = max(float((qa - qb).abs().max()) for qa, qb in zip(step.params, pb))
    assert worst < 1e-6, worst                                  # a mis-paired element would be ~1e-3 off


def _rank_main(rank, world, port, backend, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dev = torch.device("cuda", rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.allow_tf32 = False          # cuDNN picks other algorithms for batch 2 than for batch 4: keep fp32 exact
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.distributed.init_process_group(backend, rank=rank, world_size=world)
    tdl = pkg()
    ts = importlib.import_module(PKG + ".train_step")
    B = 4
    model = _model(tdl, B // world, dev, True)
    step = ts.FlatGradStep(model, lr=1e-4, max_norm=None, world=world, capturable=False)
    step.broadcast_parameters()
    inputs = _inputs(tdl, B, dev, rank * (B // world), (rank + 1) * (B // world))       # this rank's shard of the batch
    step.flat_grad.zero_()
    model(inputs)[1].total().backward()
    step.reduce_and_clip()                                       # all-reduce + 1/world
    torch.cuda.synchronize(dev)
    if rank == 0:
        torch.save(step.flat_grad.cpu(), os.path.join(out_dir, "dist_grad.pt"))
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


def test_two_rank_gradients_equal_single_gpu_on_the_concatenated_batch(tmp_path):
    tdl = pkg()
    ts = importlib.import_module(PKG + ".train_step")
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = sck.getsockname()[1]
    mp.spawn(_rank_main, args=(2, port, backend, str(tmp_path)), nprocs=2, join=True)
    dist_grad = torch.load(os.path.join(str(tmp_path), "dist_grad.pt"))
    dev = "cuda"
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = _model(tdl, 4, dev, True)
    step = ts.FlatGradStep(model, lr=1e-4, max_norm=None, world=1, capturable=False)
    model(_inputs(tdl, 4, dev))[1].total().backward()
    single = step.flat_grad.cpu()
    err = float((dist_grad - single).norm() / single.norm())
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    assert err < 1e-4, (backend, err)
