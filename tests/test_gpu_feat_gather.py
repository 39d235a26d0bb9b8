"""GPU: the bucketed-gather backward of the feature-metric loss (feat_bwd_bucket + feat_gather + feat_overflow kernels)
against the atomic-scatter kernel of the same library (library option feat_atomic = 1) on inputs chosen to stress it:

  * pixel-level noisy disparity -> sampling points scattered by many pixels (buckets of very different sizes);
  * a large translation -> most samples clip to the image border, so border buckets overflow into the overflow list;
  * h*w not a multiple of the 32-pixel tile, C = 8 (less than one 64-channel chunk) and C = 72 (one chunk + a partial one).

Both paths evaluate grid_sample's backward (mono/model/mono_fm/net.py:172-199 via autograd); they differ only in the
fp32 summation order of d_src, hence the 1e-5 tolerance.  The oracle comparison of the same path is in
tests/test_gpu_parity.py (golden feature cases run through the bucketed path because their C % 4 == 0)."""
import os

import pytest
import torch

from gpu_util import pkg, rel_l2

pytestmark = pytest.mark.gpu


def _inputs(B, C, h, w, S, seed, shift):
    g = torch.Generator().manual_seed(seed)
    tgt = torch.randn(B, C, h, w, generator=g).relu_()
    srcs = [torch.randn(B, C, h, w, generator=g).relu_() for _ in range(S)]
    disp = torch.sigmoid(torch.randn(B, 1, h, w, generator=g))                  # pixel-level noise: scattered flow
    K = torch.tensor([[0.58 * w, 0, 0.5 * w, 0], [0, 1.92 * h, 0.5 * h, 0], [0, 0, 1, 0], [0, 0, 0, 1]],
                     dtype=torch.float32).repeat(B, 1, 1)
    invK = torch.linalg.inv(K)[:, :3, :3].contiguous()
    Ps = []
    for f in range(S):
        T = torch.eye(4).repeat(B, 1, 1)
        T[:, :3, 3] = 0.01 * torch.randn(B, 3, generator=g) + torch.tensor([shift * (1 if f == 0 else -1), 0.0, 0.0])
        Ps.append(torch.matmul(K, T)[:, :3, :])
    P = torch.stack(Ps, 1).contiguous()
    return tgt, disp, P, invK, srcs


def _run(args, atomic):
    tdl = pkg()
    tgt, disp, P, invK, srcs = args
    dev = "cuda"
    leaves = [t.to(dev).clone().requires_grad_(True) for t in (tgt, disp, P)] + \
             [t.to(dev).clone().requires_grad_(True) for t in srcs]
    cfg = tdl.ops.FeatConfig(n_src=len(srcs), coef=1.0)
    with tdl._lib.options(feat_atomic=int(atomic)):
        res = tdl.ops.FeatureMetricLoss.apply(cfg, leaves[0], leaves[1], leaves[2], invK.to(dev), *leaves[3:])
        res[0].sum().backward()
        torch.cuda.synchronize()
    return float(res[0].detach()), [t.grad.detach().cpu() for t in leaves]


@pytest.mark.parametrize("B,C,h,w,S,shift", [
    (2, 8, 50, 70, 2, 0.0),       # ragged plane (3500 pixels), partial channel chunk
    (2, 72, 48, 80, 2, 0.0),      # 64 + 8 channels: two chunks
    (1, 16, 64, 96, 2, 0.5),      # half of the samples clip to the left / right border: overflow list in use
    (1, 64, 96, 320, 1, 0.05),    # bench plane size, one source
    (1, 12, 40, 60, 4, 0.1),      # four source frames
])
def test_bucketed_gather_matches_atomic_scatter(B, C, h, w, S, shift):
    args = _inputs(B, C, h, w, S, 4000 + C + h, shift)
    loss_a, grads_a = _run(args, atomic=True)
    loss_b, grads_b = _run(args, atomic=False)
    assert loss_a == loss_b                                   # the forward is the same kernel
    names = ["d_tgt", "d_disp", "dP"] + [f"d_src{f}" for f in range(S)]
    for name, ga, gb in zip(names, grads_a, grads_b):
        assert torch.isfinite(gb).all(), name
        if name.startswith("d_src") or name == "d_tgt":
            assert rel_l2(gb, ga) < 1e-5, (name, rel_l2(gb, ga))
            # elements nobody samples must be exactly zero in both (no memset in the gather path)
            assert bool(((ga == 0) == (gb == 0)).all()), name
        else:
            assert rel_l2(gb, ga) < 1e-5, (name, rel_l2(gb, ga))


def test_batch_chunking_is_invisible():
    """The bucketed backward runs in batch chunks (G stays L2-resident); chunk sizes that do not divide the batch and a
    chunk of one image must give the same gradients as the un-chunked run."""
    args = _inputs(5, 8, 40, 64, 2, 4200, 0.05)
    with pkg()._lib.options(feat_chunk=0):
        loss0, grads0 = _run(args, atomic=False)
    for chunk in (2, 1, 4):
        with pkg()._lib.options(feat_chunk=chunk):
            loss1, grads1 = _run(args, atomic=False)
        assert loss1 == loss0
        for g0, g1 in zip(grads0, grads1):
            assert rel_l2(g1, g0) < 1e-6, chunk


def test_gather_path_without_scratch_falls_back():
    """C % 4 != 0 cannot use the 16-byte G rows: the library must silently take the atomic kernel, same results."""
    args = _inputs(1, 6, 32, 48, 2, 4100, 0.0)
    loss_a, grads_a = _run(args, atomic=True)
    loss_b, grads_b = _run(args, atomic=False)
    assert loss_a == loss_b
    for ga, gb in zip(grads_a, grads_b):
        assert rel_l2(gb, ga) < 1e-5
