"""ModeFusion / Baseline -- interface- and checkpoint-compatible with the reference models/mode_fusion.py.

Stage 2 of MODE is a plain 2-D U-Net (395 GFLOP/frame, once per 6 stereo pairs).  SURVEY.md §8 keeps its
*interface* in scope (it must drop in under train_fusion.py / test_fusion.py and consume the stage-1 outputs)
but not its kernels: the layers are ordinary cuDNN convolutions with no custom op in the reference.  The module
tree below reproduces the reference's state-dict keys (251 entries for ModeFusion) and forward semantics
(mode_fusion.py:233-247, :297-307).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn


def _convbn(cin, cout, k, stride, pad, dilation):
  return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=dilation if dilation > 1 else pad, dilation=dilation, bias=False), nn.BatchNorm2d(cout))


class BasicBlock(nn.Module):
  """Two conv-BN-ReLU layers, no skip connection (reference mode_fusion.py:18-34)."""
  expansion = 1

  def __init__(self, inplanes, planes, stride=1, pad=1, dilation=1):
    super().__init__()
    self.conv1 = nn.Sequential(_convbn(inplanes, planes, 3, stride, pad, dilation), nn.ReLU(inplace=True))
    self.conv2 = nn.Sequential(_convbn(planes, planes, 3, 1, pad, dilation), nn.ReLU(inplace=True))

  def forward(self, x):
    return self.conv2(self.conv1(x))


def _stack(cin, planes, blocks, pool=False, up=False, head=False):
  layers = [nn.MaxPool2d(2, stride=2)] if pool else []
  layers += [BasicBlock(cin if i == 0 else planes, planes) for i in range(blocks)]
  if up:
    layers += [nn.ConvTranspose2d(planes, planes // 2, 2, 2), nn.BatchNorm2d(planes // 2), nn.ReLU(inplace=True)]
  if head:
    layers += [nn.Conv2d(planes, 1, kernel_size=1, padding=0, stride=1, bias=True), nn.Sigmoid()]
  return nn.Sequential(*layers)


class feature_extraction_Baseline(nn.Module):
  def __init__(self, maxdepth):
    super().__init__()
    chans = [6, 32, 64, 128, 256, 128, 64]
    for i in range(6):
      setattr(self, f'layer{i + 1}', _stack(chans[i], chans[i + 1], 2 if i == 0 else 1))
    self.layer7 = _stack(64, 32, 2, head=True)
    self.maxdepth = torch.tensor(maxdepth)

  def forward(self, x):
    for i in range(1, 8):
      x = getattr(self, f'layer{i}')(x)
    return x * self.maxdepth


class feature_extraction_MODE_Fusion(nn.Module):
  def __init__(self, maxdepth, channels, inplanes):
    super().__init__()
    c = channels
    self.depth_layer1 = _stack(inplanes['depth'], c[0], 2)
    self.depth_layer2 = _stack(c[0], c[1], 1, pool=True)
    self.depth_layer3 = _stack(c[1], c[2], 1, pool=True)
    self.rgb_layer1 = _stack(inplanes['rgb'], c[0], 2)
    self.rgb_layer2 = _stack(c[0], c[1], 1, pool=True)
    self.rgb_layer3 = _stack(c[1], c[2], 1, pool=True)
    self.fusion_layer1 = _stack(2 * c[0], c[0], 2)
    self.fusion_layer2 = _stack(2 * c[1], c[1], 2)
    self.fusion_layer3 = _stack(2 * c[2], c[2], 2)
    self.depth_layer4 = _stack(c[2], c[3], 1, pool=True, up=True)
    self.depth_layer5 = _stack(c[3], c[2], 1, up=True)
    self.depth_layer6 = _stack(c[2], c[1], 1, up=True)
    self.depth_layer7 = _stack(c[1], c[0], 2, head=True)
    self.maxdepth = torch.tensor(maxdepth)

  def forward(self, depth_input, rgb_input):
    d1 = self.depth_layer1(depth_input)
    d2 = self.depth_layer2(d1)
    d3 = self.depth_layer3(d2)
    d4 = self.depth_layer4(d3)
    r1 = self.rgb_layer1(rgb_input)
    r2 = self.rgb_layer2(r1)
    r3 = self.rgb_layer3(r2)
    f1 = self.fusion_layer1(torch.cat((d1, r1), 1))
    f2 = self.fusion_layer2(torch.cat((d2, r2), 1))
    f3 = self.fusion_layer3(torch.cat((d3, r3), 1))
    d5 = self.depth_layer5(torch.cat((f3, d4), 1))
    d6 = self.depth_layer6(torch.cat((f2, d5), 1))
    d7 = self.depth_layer7(torch.cat((f1, d6), 1))
    return d7 * self.maxdepth


def _reference_init(module):
  for m in module.modules():  # reference mode_fusion.py:266-275, 286-295
    if isinstance(m, nn.Conv2d):
      n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
      m.weight.data.normal_(0, math.sqrt(2. / n))
    elif isinstance(m, nn.BatchNorm2d):
      m.weight.data.fill_(1)
      m.bias.data.zero_()


def _run(module, net, precision, *inputs):
  """fp32: the reference's arithmetic (also the training path).  fp16 / bf16 (eval only): the folded 16-bit plan of
  models/fusion_plan.py -- channels_last cuDNN tensor-core kernels with fp32 accumulation, BN / bias / ReLU in the conv epilogues,
  weights folded once and re-folded when they change (the fusion stage is plain library code)."""
  if precision == 'fp32' or net.training:
    return net(*inputs)
  key = sum(t._version + t.data_ptr() for t in list(net.parameters()) + list(net.buffers())) & 0xFFFFFFFFFFFF
  if getattr(module, '_fplan', None) is None or module._fplan_key != key:
    from .fusion_plan import FusionPlan
    module._fplan, module._fplan_key = FusionPlan(net, torch.float16 if precision == 'fp16' else torch.bfloat16), key
  with torch.no_grad():
    return module._fplan(*inputs)


class Baseline(nn.Module):
  def __init__(self, maxdepth, precision='fp32'):
    super().__init__()
    self.feature_extraction = feature_extraction_Baseline(maxdepth)
    self.precision = precision
    _reference_init(self)

  def forward(self, depthes):
    return _run(self, self.feature_extraction, self.precision, torch.cat(depthes, 1))


class ModeFusion(nn.Module):
  def __init__(self, maxdepth, channels, inplanes, precision='fp32'):
    super().__init__()
    self.feature_extraction = feature_extraction_MODE_Fusion(maxdepth, channels, inplanes)
    self.precision = precision
    _reference_init(self)

  def forward(self, depthes, confs, rgbs):
    """depthes/confs: 6 x (B,1,H,W); rgbs: 4 x (B,3,H,W) -> (B,1,H,W) in [0, maxdepth]."""
    pairs = [t for dc in zip(depthes, confs) for t in dc]  # depth0, conf0, depth1, conf1, ...
    return _run(self, self.feature_extraction, self.precision, torch.cat(pairs, 1), torch.cat(rgbs, 1))
