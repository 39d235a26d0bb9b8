"""Net classes with the reference's names and train-mode contract -- ``net(inputs) -> (outputs, loss_dict)``
(mono/model/mono_fm/net.py:47-53) -- whose loss methods run on the fused CUDA kernels.  Registered in ``MONO``
under the reference's names so that ``MONO.module_dict[cfg.model['name']](cfg.model)`` (train.py:98-99) works."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .ops import pose_transform
from .losses import ViewSynthesisLossMixin
from .networks import DepthDecoder, ImageDecoder, PoseDecoder, ResnetEncoder
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
            # mono_fm/net.py:154 transformation_from_parameters: one fused launch (tdl_pose_fwd), no CPU path
            out[("cam_T_cam", 0, f)] = pose_transform(axisangle[:, 0], translation[:, 0], invert=f < 0)
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


@MONO.register_module
class mono_fm_joint(_DepthPoseNet):
    """Joint depth + autoencoder net (mono/model/mono_fm_joint/net.py:17-155): the feature encoder is trained in the
    loop on the un-erased target, its decoder reconstructs the image at four scales (no erase mask)."""

    def __init__(self, options):
        super().__init__(options)
        self.Encoder = ResnetEncoder(self.opt.get("extractor_num_layers", 50))
        self.Decoder = ImageDecoder(self.Encoder.num_ch_enc, "res_img")

    def forward(self, inputs):                     # net.py:47-58
        outputs = self.DepthDecoder(self.DepthEncoder(inputs["color_aug", 0, 0]))
        if not self.training:
            return outputs
        outputs.update(self.predict_poses(inputs))
        features = self.Encoder(inputs[("color", 0, 0)])
        outputs.update(self.Decoder(features, 0))
        return outputs, self.compute_losses(inputs, outputs, features)

    def compute_losses(self, inputs, outputs, features):
        return self.compute_losses_joint(inputs, outputs, features)


@MONO.register_module
class mono_fm_joint_inpaint(_DepthPoseNet):
    """Joint depth + in-painting autoencoder net (mono/model/mono_fm_joint_inpaint/net.py:19-133): the feature
    encoder is trained in the loop on the erased target, its decoder reconstructs the image at four scales."""

    def __init__(self, options):
        super().__init__(options)
        self.Encoder = ResnetEncoder(self.opt.get("extractor_num_layers", 50))
        self.Decoder = ImageDecoder(self.Encoder.num_ch_enc, "res_img")
        if self.opt.get("freeze_extractor", False):
            for q in self.Encoder.parameters():
                q.requires_grad = False

    def forward(self, inputs):
        outputs = self.DepthDecoder(self.DepthEncoder(inputs["color_aug", 0, 0]))
        if not self.training:
            return outputs
        outputs.update(self.predict_poses(inputs))
        features = self.Encoder(inputs[("color", 0, 0)] * inputs[("mask", 0, 0)])       # net.py:40
        if self.opt.get("img_reconstruct_weight", 1) != 0:
            outputs.update(self.Decoder(features, 0))
        self._decode_extra(inputs, outputs, features)
        return outputs, self.compute_losses(inputs, outputs, features)

    def _decode_extra(self, inputs, outputs, features):
        pass

    def compute_losses(self, inputs, outputs, features):
        return self.compute_losses_joint_core(inputs, outputs, features)


@MONO.register_module
class mono_fm_joint_inpaint_disentangle(mono_fm_joint_inpaint):
    """The TripleD net of config/cfg_kitti_tripleD.py (mono/model/mono_fm_joint_inpaint/net.py:398-532): adds a
    colour decoder whose full-resolution output feeds auto_res_loss."""

    def __init__(self, options):
        super().__init__(options)
        self.ColorDecoder = ImageDecoder(self.DepthEncoder.num_ch_enc, "auto_res_img")

    def _decode_extra(self, inputs, outputs, features):
        if self.opt.get("auto_res_weight", 0.0) > 0.0:
            outputs.update(self.ColorDecoder(self.DepthEncoder(inputs["color_aug", 0, 0]), 0))

    def compute_losses(self, inputs, outputs, features):
        loss_dict = super().compute_losses(inputs, outputs, features)
        loss_dict.update(self.compute_auto_res_loss(inputs, outputs))     # as the reference does (net.py:530)
        return loss_dict
