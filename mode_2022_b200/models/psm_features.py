"""PSMNet feature extractor with SPP, used only for conv='Regular' (reference models/submodule.py:205-268).
Kept as plain cuDNN-backed PyTorch (not a kernel target); same state-dict keys."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .batchnorm import BatchNorm2d
from .submodule import convbn


class BasicBlock(nn.Module):
  expansion = 1

  def __init__(self, inplanes, planes, stride, downsample, pad, dilation):
    super().__init__()
    self.conv1 = nn.Sequential(convbn(inplanes, planes, 3, stride, pad, dilation), nn.ReLU(inplace=True))
    self.conv2 = convbn(planes, planes, 3, 1, pad, dilation)
    self.downsample = downsample
    self.stride = stride

  def forward(self, x):
    out = self.conv2(self.conv1(x))
    if self.downsample is not None:
      x = self.downsample(x)
    out += x
    return out


class feature_extraction(nn.Module):
  def __init__(self):
    super().__init__()
    self.inplanes = 32
    self.firstconv = nn.Sequential(convbn(3, 32, 3, 2, 1, 1), nn.ReLU(inplace=True), convbn(32, 32, 3, 1, 1, 1), nn.ReLU(inplace=True), convbn(32, 32, 3, 1, 1, 1),
                                   nn.ReLU(inplace=True))
    self.layer1 = self._make_layer(32, 3, 1, 1, 1)
    self.layer2 = self._make_layer(64, 16, 2, 1, 1)
    self.layer3 = self._make_layer(128, 3, 1, 1, 1)
    self.layer4 = self._make_layer(128, 3, 1, 1, 2)
    for i, k in enumerate((64, 32, 16, 8), 1):
      setattr(self, f'branch{i}', nn.Sequential(nn.AvgPool2d((k, k), stride=(k, k)), convbn(128, 32, 1, 1, 0, 1), nn.ReLU(inplace=True)))
    self.lastconv = nn.Sequential(convbn(320, 128, 3, 1, 1, 1), nn.ReLU(inplace=True), nn.Conv2d(128, 32, kernel_size=1, padding=0, stride=1, bias=False))

  def _make_layer(self, planes, blocks, stride, pad, dilation):
    downsample = None
    if stride != 1 or self.inplanes != planes:
      downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes, kernel_size=1, stride=stride, bias=False), BatchNorm2d(planes))
    layers = [BasicBlock(self.inplanes, planes, stride, downsample, pad, dilation)]
    self.inplanes = planes
    layers += [BasicBlock(planes, planes, 1, None, pad, dilation) for _ in range(1, blocks)]
    return nn.Sequential(*layers)

  def forward(self, x):
    out = self.layer1(self.firstconv(x))
    raw = self.layer2(out)
    skip = self.layer4(self.layer3(raw))
    size = skip.shape[2:]
    br = [F.interpolate(getattr(self, f'branch{i}')(skip), size, mode='bilinear', align_corners=True) for i in (1, 2, 3, 4)]
    return self.lastconv(torch.cat((raw, skip, br[3], br[2], br[1], br[0]), 1))
