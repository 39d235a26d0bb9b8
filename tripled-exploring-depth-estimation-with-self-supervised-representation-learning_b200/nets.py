"""Net classes with the reference's names and train-mode contract -- ``net(inputs) -> (outputs, loss_dict)``
(mono/model/mono_fm/net.py:47-53) -- whose loss methods run on the fused CUDA kernels.  Registered in ``MONO``
under the reference's names so that ``MONO.module_dict[cfg.model['name']](cfg.model)`` (train.py:98-99) works."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .geometry import transformation_from_parameters
from .losses import ViewSynthesisLossMixin
from .networks import DepthDecoder, PoseDecoder, ResnetEncoder
from .registry import MONO


class _DepthPoseNet(ViewSynthesisLossMixin, nn.Module):
    def __init__(self, options):
        super().__init__()
        self.opt = options
        self.DepthEncoder = ResnetEncoder(self.opt.depth_num_layers)
        self.DepthDecoder = DepthDecoder(self.DepthEncoder.num_ch_enc)
        self.PoseEncoder = ResnetEncoder(self.opt.pose_num_layers, num_input_images=2)
        self.PoseDecoder = PoseDecoder(self.PoseEncoder.num_ch_enc)

    def predict_poses(self, inputs):
        """mono/model/mono_fm/net.py:142-155: temporal order of the image pair, inverted pose for f < 0."""
        out = {}
        size = [self.opt.height, self.opt.width]
        img = {f: F.interpolate(inputs["color_aug", f, 0], size, mode="bilinear", align_corners=False)
               if list(inputs["color_aug", f, 0].shape[-2:]) != size else inputs["color_aug", f, 0]
               for f in self.opt.frame_ids if f != "s"}
        for f in self.opt.frame_ids[1:]:
            if f == "s":
                continue
            pair = [img[f], img[0]] if f < 0 else [img[0], img[f]]
            axisangle, translation = self.PoseDecoder(self.PoseEncoder(torch.cat(pair, 1)))
            out[("cam_T_cam", 0, f)] = transformation_from_parameters(axisangle[:, 0], translation[:, 0], invert=f < 0)
        return out

    def forward(self, inputs):
        outputs = self.DepthDecoder(self.DepthEncoder(inputs["color_aug", 0, 0]))
        if self.training:
            outputs.update(self.predict_poses(inputs))
            return outputs, self.compute_losses(inputs, outputs)
        return outputs


@MONO.register_module
class Baseline(_DepthPoseNet):
    """monodepth2-style net: photometric + automask + smoothness (mono/model/mono_baseline/net.py)."""
    _smooth_weight_key = "disparity_smoothness"

    def compute_losses(self, inputs, outputs):
        return self.compute_losses_baseline(inputs, outputs)


@MONO.register_module
class mono_fm(_DepthPoseNet):
    """FeatDepth net: adds the feature-metric loss on extractor features (mono/model/mono_fm/net.py)."""

    def __init__(self, options):
        super().__init__(options)
        self.extractor = ResnetEncoder(self.opt.get("extractor_num_layers", 50))
        if self.opt.get("extractor_pretrained_path", None) is not None:      # frozen when pre-trained (net.py:24-25)
            state = torch.load(self.opt.extractor_pretrained_path, map_location="cpu")
            self.extractor.load_state_dict(state.get("state_dict", state), strict=False)
            for q in self.extractor.parameters():
                q.requires_grad = False
        # the loss reads extractor level 0 only: the deeper stages get no gradient (DDP would otherwise need
        # find_unused_parameters); their tensors stay in the state-dict
        for q in self.extractor.stages.parameters():
            q.requires_grad = False

    def compute_losses(self, inputs, outputs):
        return self.compute_losses_fm(inputs, outputs)
