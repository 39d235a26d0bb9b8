"""The networks AROUND the fused loss.  They stay plain PyTorch/cuDNN (BASELINE.json north_star: "The ResNet
encoder/depth decoder/pose net stay in PyTorch") and exist here only so that a full training step can be run
and timed; they follow the reference's interfaces (what goes in, which keys / shapes come out), not its code:

  ResnetEncoder(num_layers, num_input_images)   5 feature maps at H/2 .. H/32
        (mono/model/mono_fm/depth_encoder.py, pose_encoder.py:11-49, mono_autoencoder/encoder.py:35-43)
  DepthDecoder(num_ch_enc)                      {("disp", 0, s): sigmoid map at H/2^(s+1)}, s = 0..3
        (mono/model/mono_fm/depth_decoder.py:45-98)
  PoseDecoder(num_ch_enc)                       axisangle, translation (B, 2, 1, 3), scaled by 0.01
        (mono/model/mono_fm/pose_decoder.py:16-26)
  ImageDecoder(num_ch_enc, key)                 {(key, 0, s): sigmoid RGB at H/2^s}, s = 0..3, from the deepest
        encoder level (mono/model/mono_fm_joint/decoder.py:7-57 "res_img", :60-112 "auto_res_img")
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision.models as tvm

_RESNETS = {18: tvm.resnet18, 34: tvm.resnet34, 50: tvm.resnet50, 101: tvm.resnet101}


class ResnetEncoder(nn.Module):
    def __init__(self, num_layers=18, num_input_images=1):
        super().__init__()
        if num_layers not in _RESNETS:
            raise ValueError(f"{num_layers} is not a valid number of resnet layers")
        net = _RESNETS[num_layers](weights=None)
        if num_input_images > 1:
            net.conv1 = nn.Conv2d(3 * num_input_images, 64, 7, 2, 3, bias=False)
        self.stem = nn.Sequential(net.conv1, net.bn1, net.relu)
        self.pool = net.maxpool
        self.stages = nn.ModuleList([net.layer1, net.layer2, net.layer3, net.layer4])
        wide = 4 if num_layers > 34 else 1
        self.num_ch_enc = [64, 64 * wide, 128 * wide, 256 * wide, 512 * wide]

    def first(self, x):
        """Only the first feature map (stem output, 64 ch at H/2): all the feature-metric loss reads."""
        return self.stem((x - 0.45) / 0.225)

    def forward(self, x):
        feats = [self.first(x)]
        y = self.pool(feats[0])
        for stage in self.stages:
            y = stage(y)
            feats.append(y)
        return feats


class _Refine(nn.Module):
    """conv3x3 + chained residual pooling (max-pool 5 -> conv1x1, summed) + conv3x3."""

    def __init__(self, cin, ch, hops=4):
        super().__init__()
        self.enter = nn.Conv2d(cin, ch, 3, padding=1)
        self.hops = nn.ModuleList([nn.Conv2d(ch, ch, 1, bias=False) for _ in range(hops)])
        self.leave = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        x = F.leaky_relu(self.enter(x))
        path = x
        for hop in self.hops:
            path = hop(F.max_pool2d(path, 5, 1, 2))
            x = x + path
        return F.leaky_relu(self.leave(x))


class DepthDecoder(nn.Module):
    def __init__(self, num_ch_enc, width=256):
        super().__init__()
        self.squeeze = nn.ModuleList([nn.Conv2d(c, 512 if i == 4 else width, 1, bias=False)
                                      for i, c in enumerate(num_ch_enc)][1:])          # levels 1..4
        self.refine = nn.ModuleList([_Refine(512 if lvl == 4 else 2 * width + 1, width) for lvl in (1, 2, 3, 4)])
        self.heads = nn.ModuleList([nn.Conv2d(width, 1, 3, padding=1) for _ in range(4)])
        self.drop = nn.Dropout(0.5)

    def forward(self, feats, frame_id=0):
        out, x, disp = {}, None, None
        for lvl in (4, 3, 2, 1):                       # coarse to fine; level l is at H / 2^(l+1)
            f = feats[lvl]
            if lvl >= 3:
                f = self.drop(f)
            y = self.squeeze[lvl - 1](f)
            if x is not None:
                y = torch.cat((y, x, disp), 1)
            x = self.refine[lvl - 1](y)
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            disp = torch.sigmoid(self.heads[lvl - 1](x))
            out[("disp", frame_id, lvl - 1)] = disp
        return out


class PoseDecoder(nn.Module):
    def __init__(self, num_ch_enc, num_frames=2):
        super().__init__()
        self.num_frames = num_frames
        self.body = nn.Sequential(nn.Conv2d(num_ch_enc[-1], 256, 1), nn.ReLU(True),
                                  nn.Conv2d(256, 256, 3, padding=1), nn.ReLU(True),
                                  nn.Conv2d(256, 256, 3, padding=1), nn.ReLU(True),
                                  nn.Conv2d(256, 6 * num_frames, 1))

    def forward(self, feats):
        pose = 0.01 * self.body(feats[-1]).mean((2, 3)).view(-1, self.num_frames, 1, 6)
        return pose[..., :3], pose[..., 3:]


class ImageDecoder(nn.Module):
    """Autoencoder / colour decoder of the joint nets: five (conv3x3 + ELU, x2 nearest up-sampling) stages from the
    deepest encoder map back to full resolution, RGB heads on the last four."""

    def __init__(self, num_ch_enc, key="res_img", widths=(256, 128, 64, 32, 16)):
        super().__init__()
        self.key = key
        chans = [num_ch_enc[-1], *widths]
        self.up = nn.ModuleList([nn.Sequential(nn.Conv2d(chans[i], chans[i + 1], 3, padding=1), nn.ELU(True),
                                               nn.Upsample(scale_factor=2, mode="nearest"),
                                               nn.Conv2d(chans[i + 1], chans[i + 1], 3, padding=1), nn.ELU(True))
                                 for i in range(5)])
        self.heads = nn.ModuleList([nn.Conv2d(w, 3, 3, padding=1) for w in widths[1:]])

    def forward(self, feats, frame_id=0):
        out, x = {}, feats[-1]
        for i, stage in enumerate(self.up):
            x = stage(x)
            if i >= 1:                                   # stages 1..4 are at H/8, H/4, H/2, H
                out[(self.key, frame_id, 4 - i)] = torch.sigmoid(self.heads[i - 1](x))
        return out
