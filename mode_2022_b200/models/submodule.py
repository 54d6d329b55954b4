"""Building blocks of the stereo network -- same module tree (hence the same state-dict keys) as the
reference models/submodule.py, so reference checkpoints load unchanged.  The regular 2-D layers are ordinary
cuDNN-backed nn modules (not a custom-kernel target); the spherical layers and everything
3-D are executed by libmode_b200 from ModeDisparity.forward, using these modules only as parameter holders; the
modules' own forward()s are the differentiable TRAINING path (SphereConv -> libmode_b200 forward/backward kernels).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .batchnorm import BatchNorm2d, BatchNorm3d
from .sphere_conv import SphereConv


def convbn(in_planes, out_planes, kernel_size, stride, pad, dilation):
  """Conv2d + BN (reference submodule.py:15-17)."""
  return nn.Sequential(nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, stride=stride, padding=dilation if dilation > 1 else pad, dilation=dilation, bias=False),
                       BatchNorm2d(out_planes))


def convbn_3d(in_planes, out_planes, kernel_size, stride, pad):
  """Conv3d + BN (reference submodule.py:20-22)."""
  return nn.Sequential(nn.Conv3d(in_planes, out_planes, kernel_size=kernel_size, padding=pad, stride=stride, bias=False), BatchNorm3d(out_planes))


def sphereConvbn(in_height, in_width, sphereType, in_planes, out_planes, kernel_size, stride, pad, dilation):
  """SphereConv + BN (reference submodule.py:61-75)."""
  return nn.Sequential(SphereConv(in_height, in_width, sphereType, in_planes, out_planes, kernel_size=kernel_size, stride=stride,
                                  padding=dilation if dilation > 1 else pad, dilation=dilation, bias=False), BatchNorm2d(out_planes))


class RegularBasicBlock(nn.Module):
  """Residual block of the regular layers (reference submodule.py:94-119)."""
  expansion = 1

  def __init__(self, inplanes, planes, stride, downsample, pad, dilation):
    super().__init__()
    self.conv1 = nn.Sequential(convbn(inplanes, planes, 3, stride, pad, dilation), nn.ReLU(inplace=True))
    self.conv2 = convbn(planes, planes, 3, 1, pad, dilation)
    self.relu = nn.ReLU(inplace=True)
    self.downsample = downsample
    self.stride = stride

  def forward(self, x):
    out = self.conv2(self.conv1(x))
    if self.downsample is not None:
      x = self.downsample(x)
    out += x
    return self.relu(out)


class SphereBasicBlock(nn.Module):
  """Residual block of layer4 (reference submodule.py:122-147): parameter holder; see ModeDisparity."""
  expansion = 1

  def __init__(self, in_height, in_width, sphereType, inplanes, planes, stride, downsample, pad, dilation):
    super().__init__()
    self.conv1 = nn.Sequential(sphereConvbn(in_height, in_width, sphereType, inplanes, planes, 3, stride, pad, dilation), nn.ReLU(inplace=True))
    self.conv2 = sphereConvbn(in_height // stride, in_width // stride, sphereType, planes, planes, 3, 1, pad, dilation)
    self.relu = nn.ReLU(inplace=True)
    self.downsample = downsample
    self.stride = stride

  def forward(self, x):
    """Training path (autograd through SphereConvFunction, libmode_b200 forward + backward kernels); inference goes
    through ModeDisparity's fused plan instead."""
    out = self.conv2(self.conv1(x))
    res = x if self.downsample is None else self.downsample(x)
    return self.relu(out + res)


def bn_affine_host(bn: nn.modules.batchnorm._BatchNorm):
  """Eval-mode BN as y = x*scale + shift, folded on the HOST in fp32: a plan folds ~75 BatchNorms once per set of weights;
  on the device that is ~900 tiny elementwise launches, on the host it is four small D2H copies per layer and no kernel."""
  w, b, m, v = (t.detach().float().cpu() for t in (bn.weight, bn.bias, bn.running_mean, bn.running_var))
  scale = w * torch.rsqrt(v + bn.eps)
  shift = b - m * scale
  return scale.contiguous(), shift.contiguous()


def bn_affine(bn: nn.modules.batchnorm._BatchNorm):
  """Eval-mode BN as y = x*scale + shift (fp32, on the module's device)."""
  scale, shift = bn_affine_host(bn)
  dev = bn.weight.device
  return scale.to(dev), shift.to(dev)


class sphere_feature_extraction(nn.Module):
  """Feature extractor (reference submodule.py:151-201): firstconv, layer1-3 regular, layer4 spherical, lastconv."""

  def __init__(self, in_height, in_width, sphereType):
    super().__init__()
    self.firstconv = nn.Sequential(convbn(3, 32, 7, 2, 3, 1), nn.ReLU(inplace=True), convbn(32, 32, 3, 1, 1, 1), nn.ReLU(inplace=True), convbn(32, 32, 3, 1, 1, 1),
                                   nn.ReLU(inplace=True))
    self.layer1 = self._make_layer(RegularBasicBlock, in_height // 2, in_width // 2, sphereType, 32, 64, 3, 1, 1, 1)
    self.layer2 = self._make_layer(RegularBasicBlock, in_height // 2, in_width // 2, sphereType, 64, 64, 8, 2, 1, 1)
    self.layer3 = self._make_layer(RegularBasicBlock, in_height // 4, in_width // 4, sphereType, 64, 64, 4, 1, 1, 2)
    self.layer4 = self._make_layer(SphereBasicBlock, in_height // 4, in_width // 4, sphereType, 64, 128, 8, 1, 1, 1)
    self.lastconv = nn.Sequential(convbn(256, 128, 1, 1, 0, 1), nn.ReLU(inplace=True), convbn(128, 128, 3, 1, 1, 1), nn.ReLU(inplace=True), convbn(128, 32, 1, 1, 0, 1),
                                  nn.ReLU(inplace=True))

  @staticmethod
  def _make_layer(block, height, width, sphereType, inplanes, planes, blocks, stride, pad, dilation):
    downsample = None
    if stride != 1 or inplanes != planes * block.expansion:
      downsample = nn.Sequential(nn.Conv2d(inplanes, planes * block.expansion, kernel_size=1, stride=stride, bias=False), BatchNorm2d(planes * block.expansion))
    if block is SphereBasicBlock:
      layers = [block(height, width, sphereType, inplanes, planes, stride, downsample, pad, dilation)]
      layers += [block(height // stride, width // stride, sphereType, planes, planes, 1, None, pad, dilation) for _ in range(1, blocks)]
    else:
      layers = [block(inplanes, planes, stride, downsample, pad, dilation)]
      layers += [block(planes, planes, 1, None, pad, dilation) for _ in range(1, blocks)]
    return nn.Sequential(*layers)

  def forward(self, x):
    """Training path (reference submodule.py:192-201): plain module execution, differentiable end to end."""
    raw = self.layer2(self.layer1(self.firstconv(x)))
    reg = self.layer3(raw)
    # the spherical layers read and write NCHW (the reference op's layout): when the regular layers run channels-last (training,
    # mode_disparity.py) convert ONCE here instead of once per spherical conv and residual add, and come back for lastconv
    cl = reg.dim() == 4 and not reg.is_contiguous() and reg.is_contiguous(memory_format=torch.channels_last)
    sph = self.layer4(reg.contiguous() if cl else reg)
    if cl:
      sph = sph.contiguous(memory_format=torch.channels_last)
    return self.lastconv(torch.cat((raw, reg, sph), 1))
