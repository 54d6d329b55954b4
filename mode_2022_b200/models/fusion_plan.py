"""16-bit inference plan of the fusion stage: the same module tree as ModeFusion / Baseline, executed as
BN-folded channels_last cuDNN convolutions with fused bias + ReLU epilogues (no autocast: weights are folded and converted ONCE per
set of weights, not cast on every forward), the 2x2 stride-2 transposed convolutions with their BatchNorm folded in, and the
sigmoid head in fp32.  Static shapes, no host synchronisation: the whole stage can be captured in a CUDA graph
(mode_2022_b200.pipeline.FusionStage).  Reference dataflow: models/mode_fusion.py:91-247, 297-307."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .plan_bf16 import _FoldedConv2d, _probe_fused_cudnn
from . import plan_bf16
from .submodule import bn_affine_host


class _FoldedDeconv2x2:
  """ConvTranspose2d(k=2, s=2, bias) + eval-BN + ReLU folded into one transposed convolution."""

  def __init__(self, deconv: nn.ConvTranspose2d, bn: nn.BatchNorm2d, dtype):
    scale, shift = bn_affine_host(bn)
    dev = deconv.weight.device
    w = deconv.weight.detach().float().cpu() * scale.view(1, -1, 1, 1)  # (in, out, 2, 2): BN acts on the OUTPUT channels
    b = (deconv.bias.detach().float().cpu() if deconv.bias is not None else torch.zeros_like(shift)) * scale + shift
    self.w = w.to(dtype).contiguous(memory_format=torch.channels_last).to(dev)
    self.b = b.to(dtype).to(dev)
    self.stride = deconv.stride

  def __call__(self, x):
    return F.relu_(F.conv_transpose2d(x, self.w, self.b, self.stride))


class FusionPlan:
  def __init__(self, net: nn.Module, dtype=torch.float16):
    """`net`: feature_extraction_MODE_Fusion or feature_extraction_Baseline (eval mode)."""
    self.dtype = dtype
    plan_bf16._FUSED_CUDNN = plan_bf16._FUSED_CUDNN or _probe_fused_cudnn()
    self.maxdepth = float(net.maxdepth)
    self.stacks = {}
    for name, seq in net.named_children():
      if isinstance(seq, nn.Sequential):
        self.stacks[name] = self._compile(seq, dtype)
    self.is_fusion = hasattr(net, 'fusion_layer1')

  @staticmethod
  def _compile(seq: nn.Sequential, dtype):
    ops, mods, i = [], list(seq), 0
    while i < len(mods):
      m = mods[i]
      if isinstance(m, nn.MaxPool2d):
        ops.append(('pool', m.kernel_size, m.stride))
      elif hasattr(m, 'conv1') and hasattr(m, 'conv2'):  # BasicBlock: two conv-BN-ReLU layers
        ops.append(('conv', _FoldedConv2d(m.conv1[0][0], m.conv1[0][1], dtype)))
        ops.append(('conv', _FoldedConv2d(m.conv2[0][0], m.conv2[0][1], dtype)))
      elif isinstance(m, nn.ConvTranspose2d):  # followed by BatchNorm2d, ReLU
        ops.append(('deconv', _FoldedDeconv2x2(m, mods[i + 1], dtype)))
        i += 2
      elif isinstance(m, nn.Conv2d):  # 1x1 head, followed by Sigmoid
        ops.append(('head', m.weight.detach().to(dtype).contiguous(memory_format=torch.channels_last), m.bias.detach().to(dtype)))
        i += 1
      else:
        raise NotImplementedError(f'FusionPlan: unexpected module {type(m).__name__}')
      i += 1
    return ops

  def _run(self, name, x):
    for op in self.stacks[name]:
      if op[0] == 'pool':
        x = F.max_pool2d(x, op[1], op[2])
      elif op[0] == 'conv':
        x = op[1](x, True)
      elif op[0] == 'deconv':
        x = op[1](x)
      else:  # head: 1x1 conv in 16-bit, sigmoid and the depth scale in fp32
        x = torch.sigmoid(F.conv2d(x, op[1], op[2]).float()) * self.maxdepth
    return x

  def _in(self, t):
    return t.to(self.dtype).contiguous(memory_format=torch.channels_last)

  def __call__(self, depth_input, rgb_input=None):
    r = self._run
    if not self.is_fusion:  # Baseline: layer1..7 in sequence (reference mode_fusion.py:233-247)
      x = self._in(depth_input)
      for i in range(1, 8):
        x = r(f'layer{i}', x)
      return x
    d1 = r('depth_layer1', self._in(depth_input))
    d2 = r('depth_layer2', d1)
    d3 = r('depth_layer3', d2)
    d4 = r('depth_layer4', d3)
    r1 = r('rgb_layer1', self._in(rgb_input))
    r2 = r('rgb_layer2', r1)
    r3 = r('rgb_layer3', r2)
    f1 = r('fusion_layer1', torch.cat((d1, r1), 1))
    f2 = r('fusion_layer2', torch.cat((d2, r2), 1))
    f3 = r('fusion_layer3', torch.cat((d3, r3), 1))
    d5 = r('depth_layer5', torch.cat((f3, d4), 1))
    d6 = r('depth_layer6', torch.cat((f2, d5), 1))
    return r('depth_layer7', torch.cat((f1, d6), 1))
