"""Pose prologue: transformation_from_parameters / rot_from_axisangle /
get_translation_matrix, mono/model/mono_fm/net.py:201-253.

tests/golden/pose/pose_b16.pt was produced by the REAL reference methods (tests/golden/make_pose_golden.py): small pose-decoder
sized rotations, two large ones, an exactly-zero one, both `invert` settings, with autograd gradients of <W, T>.
  * CPU: the oracle restatement and the package's PyTorch helper reproduce it (the pin);
  * GPU: the fused tdl_pose_fwd / tdl_pose_bwd kernels match it through the C ABI."""
import os

import pytest
import torch

from gpu_util import pkg, rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))


def _golden():
    return torch.load(os.path.join(HERE, "golden", "pose", "pose_b16.pt"))


@pytest.mark.parametrize("invert", [False, True])
def test_oracle_reproduces_reference_pose(invert):
    from oracle import restatement as R
    tdl = pkg()
    rec = _golden()
    for fn in (R.transformation_from_parameters, tdl.geometry.transformation_from_parameters):
        a = rec["axisangle"].clone().requires_grad_(True)
        t = rec["translation"].clone().requires_grad_(True)
        T = fn(a, t, invert)
        assert torch.equal(T, rec[("T", invert)]), fn.__module__                   # bit-exact on the CPU
        (T * rec["W"]).sum().backward()
        torch.testing.assert_close(a.grad, rec[("d_axisangle", invert)], rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(t.grad, rec[("d_translation", invert)], rtol=1e-6, atol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("invert", [False, True])
def test_fused_pose_matches_reference(invert):
    tdl = pkg()
    rec = _golden()
    a = rec["axisangle"].cuda().requires_grad_(True)
    t = rec["translation"].cuda().requires_grad_(True)
    T = tdl.ops.pose_transform(a, t, invert)
    assert T.shape == (16, 4, 4)
    # same fp32 operation order; sin / cos / the 3-term norm may differ in the last bit between devices
    assert float((T.detach().cpu() - rec[("T", invert)]).abs().max()) <= 2e-7
    (T * rec["W"].cuda()).sum().backward()
    assert a.grad.shape == rec["axisangle"].shape and t.grad.shape == rec["translation"].shape
    assert rel_l2(a.grad.cpu(), rec[("d_axisangle", invert)]) < 1e-5
    assert rel_l2(t.grad.cpu(), rec[("d_translation", invert)]) < 1e-5
    assert float(a.grad[3].abs().max()) == 0.0                                     # zero rotation: the reference's sub-gradient


@pytest.mark.gpu
def test_fused_pose_matches_torch_composition_on_gpu():
    """Same device, plain PyTorch composition (the ~40-launch sequence the kernel replaces), pose-decoder sized inputs."""
    tdl = pkg()
    g = torch.Generator().manual_seed(5)
    aa = (0.01 * torch.randn(8, 1, 3, generator=g)).cuda()
    tr = (0.01 * torch.randn(8, 1, 3, generator=g)).cuda()
    W = torch.randn(8, 4, 4, generator=g).cuda()
    for invert in (False, True):
        a1, t1 = aa.clone().requires_grad_(True), tr.clone().requires_grad_(True)
        a2, t2 = aa.clone().requires_grad_(True), tr.clone().requires_grad_(True)
        T1 = tdl.ops.pose_transform(a1, t1, invert)
        T2 = tdl.geometry.transformation_from_parameters(a2, t2, invert)
        assert float((T1 - T2).detach().abs().max()) <= 2e-7
        (T1 * W).sum().backward()
        (T2 * W).sum().backward()
        assert rel_l2(a1.grad, a2.grad) < 1e-5 and rel_l2(t1.grad, t2.grad) < 1e-5


def test_pose_refuses_cpu_tensors():
    tdl = pkg()
    with pytest.raises(tdl._lib.TdlError, match="no CPU implementation"):
        tdl.ops.pose_transform(torch.zeros(2, 1, 3), torch.zeros(2, 1, 3), False)
