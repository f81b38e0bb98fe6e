"""CPU oracle of the policy's observation encoder -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline leg may import this module; the product
(v2a_b200/) never does.  A functional restatement, driven by a reference-format ``state_dict``, of

  VisualCore.forward          diffuser/diffusion_policy/common/vision_nets.py:113-177
    ResNet18Conv               common/vision_nets.py:9-39  = torchvision.models.resnet18 children[:-2]
                               (third-party torchvision, installed here; BasicBlock / stem restated below from its
                               published architecture) with every BatchNorm2d(C) replaced by GroupNorm(C // 16, C)
                               (model/multi_image_obs_encoder.py:67-74)
    SpatialSoftmax.forward     common/base_nets.py:234-285 (1x1 conv to K maps, softmax over pixels / temperature,
                               expected (x, y) on a linspace(-1, 1) grid, [B, K, 2]; training noise * noise_std)
    Flatten -> Linear          common/vision_nets.py:133-143
  MultiImageObsEncoder.forward model/multi_image_obs_encoder.py:144-196 (independent per-key models, SORTED keys)

Pinned (tests/test_encoder_oracle.py) against the UNMODIFIED reference classes when /root/reference is mounted and
against tests/golden/encoder_golden.pt generated from them (tests/golden/make_encoder_golden.py).
Gradients come from torch autograd on these functional ops.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]
GN_EPS = 1e-5


def _gn(sd: SD, p: str, x: Tensor) -> Tensor:
    c = x.shape[1]
    return F.group_norm(x, c // 16, sd[p + "weight"], sd[p + "bias"], GN_EPS)


def _relu(x: Tensor, site: int) -> Tensor:
    return F.relu(x)


def basic_block(sd: SD, p: str, x: Tensor, stride: int, relu=_relu, site: int = 0) -> Tensor:
    """torchvision BasicBlock: conv3x3(stride) -> norm -> relu -> conv3x3 -> norm (+ downsample(x)) -> relu.
    `relu(x, site)` is the activation hook (tests substitute a fixed mask to compare gradients mask for mask)."""
    out = relu(_gn(sd, p + "bn1.", F.conv2d(x, sd[p + "conv1.weight"], stride=stride, padding=1)), site)
    out = _gn(sd, p + "bn2.", F.conv2d(out, sd[p + "conv2.weight"], padding=1))
    idn = x
    if p + "downsample.0.weight" in sd:
        idn = _gn(sd, p + "downsample.1.", F.conv2d(x, sd[p + "downsample.0.weight"], stride=stride))
    return relu(out + idn, site + 1)


def resnet18_gn_trunk(sd: SD, p: str, x: Tensor, relu=_relu) -> Tensor:
    """p = '...backbone.nets.': 0 conv7x7 s2 p3 (no bias), 1 norm, 2 relu, 3 maxpool(3, 2, 1), 4..7 = layer1..4.
    ReLU sites: 0 = stem, 2k+1 / 2k+2 = inner / output activation of BasicBlock k (k = 0..7)."""
    x = relu(_gn(sd, p + "1.", F.conv2d(x, sd[p + "0.weight"], stride=2, padding=3)), 0)
    x = F.max_pool2d(x, 3, 2, 1)
    k = 0
    for li in range(4, 8):
        for bi in range(2):
            x = basic_block(sd, f"{p}{li}.{bi}.", x, 2 if (li > 4 and bi == 0) else 1, relu, 2 * k + 1)
            k += 1
    return x


def spatial_softmax(sd: SD, p: str, feat: Tensor) -> Tensor:
    """p = '...pool.': nets (1x1 conv), temperature, pos_x, pos_y -> keypoints [B, K, 2] (base_nets.py:248-267)."""
    f = F.conv2d(feat, sd[p + "nets.weight"], sd[p + "nets.bias"])
    B, K, H, W = f.shape
    att = F.softmax(f.reshape(-1, H * W) / sd[p + "temperature"], dim=-1)
    ex = torch.sum(sd[p + "pos_x"] * att, dim=1, keepdim=True)
    ey = torch.sum(sd[p + "pos_y"] * att, dim=1, keepdim=True)
    return torch.cat([ex, ey], 1).view(-1, K, 2)


def visual_core_forward(sd: SD, p: str, x: Tensor, relu=_relu) -> Tensor:
    """p = prefix of one VisualCore ('obs_encoder.key_model_map.img_obs_1.'); x [B, 3, H, W] -> [B, 64]."""
    kp = spatial_softmax(sd, p + "pool.", resnet18_gn_trunk(sd, p + "backbone.nets.", x, relu))
    return F.linear(kp.flatten(1), sd[p + "nets.3.weight"], sd[p + "nets.3.bias"])


def obs_encoder_forward(sd: SD, p: str, obs: Dict[str, Tensor]) -> Tensor:
    """p = 'obs_encoder.'; features of the rgb keys in sorted order, concatenated (multi_image_obs_encoder.py:171-196)."""
    return torch.cat([visual_core_forward(sd, f"{p}key_model_map.{k}.", obs[k]) for k in sorted(obs)], dim=-1)
