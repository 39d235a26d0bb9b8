"""GPU: the registered net classes run a training forward + backward through the fused loss, with the
reference's train-mode contract ``net(inputs) -> (outputs, loss_dict)`` and loss_dict / outputs keys."""
import importlib

import pytest
import torch

from gpu_util import PKG, pkg

pytestmark = pytest.mark.gpu


H, W = 96, 128      # level-4 feature maps are 3x4: the smallest size whose second differences are not empty (NaN mean)


def _cfg(tdl, name):
    return tdl.config.ConfigDict(dict(
        name=name, depth_num_layers=18, pose_num_layers=18, extractor_num_layers=18, frame_ids=[0, -1, 1],
        imgs_per_gpu=2, height=H, width=W, scales=[0, 1, 2, 3], min_depth=0.1, max_depth=100.0, automask=True,
        disp_norm=True, perception_weight=1e-3, smoothness_weight=1e-3, disparity_smoothness=1e-3, dis=1e-3, cvt=1e-3,
        auto_res_weight=5e-3, extractor_pretrained_path=None))


def _inputs(tdl, B=2):
    inputs, _, _ = tdl.synth.make_inputs(B, H, W, seed=3, with_noise=False)
    inputs = {k: v.cuda() for k, v in inputs.items()}
    for f in (0, -1, 1):
        inputs[("color_aug", f, 0)] = inputs[("color", f, 0)]
    mask = torch.ones(B, 3, H, W, device="cuda")
    mask[:, :, 10:26, 20:36] = 0
    inputs[("mask", 0, 0)] = mask
    return inputs


@pytest.mark.parametrize("name,keys", [
    ("Baseline", [("min_reconstruct_loss", 0), ("smooth_loss", 3)]),
    ("mono_fm", [("min_reconstruct_loss", 0), ("min_perceptional_loss", 2), ("smooth_loss", 3)]),
    ("mono_fm_joint", [("feature_regularization_loss", 4), ("min_perceptional_loss", 3), ("img_reconstruct_loss", 1),
                       ("min_reconstruct_loss", 0), ("smooth_loss", 2)]),
    ("mono_fm_joint_inpaint_disentangle", [("feature_regularization_loss", 4), "min_perceptional_loss",
                                           ("img_reconstruct_loss", 1), ("min_reconstruct_loss", 0), "auto_res_loss"]),
])
def test_registered_net_trains_one_step(name, keys):
    tdl = pkg()
    importlib.import_module(PKG + ".nets")
    torch.manual_seed(0)
    net = tdl.MONO.module_dict[name](_cfg(tdl, name)).cuda().train()
    outputs, loss_dict = net(_inputs(tdl))
    for k in keys:
        assert k in loss_dict, k
    assert ("color", -1, 0) in outputs and ("min_index", 3) in outputs and ("cam_T_cam", 0, 1) in outputs
    total = loss_dict.total()
    ref_total = sum(v.mean() for v in loss_dict.values())          # batch_processor's sum (trainer.py:39-48)
    assert torch.isfinite(total) and abs(float(total.detach()) - float(ref_total.detach())) < 1e-6 * abs(float(ref_total.detach())) + 1e-9
    total.backward()
    for mod in ("DepthDecoder", "PoseDecoder"):
        grads = [q.grad for q in getattr(net, mod).parameters() if q.grad is not None]
        assert grads and all(torch.isfinite(g).all() for g in grads) and any(float(g.abs().max()) > 0 for g in grads), mod
    net.eval()
    with torch.no_grad():
        out = net({("color_aug", 0, 0): torch.rand(1, 3, H, W, device="cuda")})
    assert set(out) == {("disp", 0, s) for s in range(4)}
