"""16-bit throughput plan: NHWC / NDHWC bf16 (or fp16) activations, tensor-core kernels, fp32 accumulation, fp32 logits.

bf16 is the dtype BASELINE.json names; fp16 runs the identical kernels at the identical speed with 3 more mantissa
bits (8x smaller rounding error per stored activation), which is what brings the 3-D stack under the 0.01 px EPE budget
on un-trained networks (measured: bf16 0.03 px, fp16 0.004 px)."""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .plan import _PlanBase, _w
from .submodule import bn_affine, bn_affine_host


class _FoldedConv2d:
  """Conv2d + eval-BN folded into (bf16 channels_last weight, bias); cuDNN-backed."""

  def __init__(self, conv: nn.Conv2d, bn: nn.BatchNorm2d, dtype):
    scale, shift = bn_affine_host(bn)  # folded on the host: no elementwise launches while a plan is built
    dev = conv.weight.device
    w = conv.weight.detach().float().cpu() * scale.view(-1, 1, 1, 1)
    self.w = w.to(dtype).contiguous(memory_format=torch.channels_last).to(dev)
    self.shift = shift.to(dev)
    self.b = shift.to(dtype).to(dev)
    self.stride, self.padding, self.dilation = conv.stride, conv.padding, conv.dilation

  def move_bias_into(self, other: '_FoldedConv2d'):
    """This conv's output is only ever added to `other`'s pre-activation (a BasicBlock's downsample branch): give the
    BN shift to `other`, so that this conv runs bias-free instead of paying a separate elementwise bias pass."""
    other.shift = other.shift + self.shift
    other.b = other.shift.to(other.b.dtype)
    self.shift, self.b = None, None

  def __call__(self, x, relu, residual=None):
    """conv + folded-BN bias [+ residual] [+ ReLU]; uses cuDNN's fused conv-bias-(add)-ReLU when available."""
    if _FUSED_CUDNN and relu:
      if residual is None:
        return torch.cudnn_convolution_relu(x, self.w, self.b, self.stride, self.padding, self.dilation, 1)
      return torch.cudnn_convolution_add_relu(x, self.w, residual, 1.0, self.b, self.stride, self.padding, self.dilation, 1)
    y = F.conv2d(x, self.w, self.b, self.stride, self.padding, self.dilation)
    if residual is not None:
      y = y.add_(residual)
    return F.relu_(y) if relu else y


def _probe_fused_cudnn():
  """cuDNN runtime-fused conv+bias(+add)+ReLU for channels_last 16-bit tensors (plain cuDNN calls)."""
  try:
    for dt in (torch.bfloat16, torch.float16):
      x = torch.randn(1, 32, 8, 8, device='cuda', dtype=dt).contiguous(memory_format=torch.channels_last)
      w = torch.randn(32, 32, 3, 3, device='cuda', dtype=dt).contiguous(memory_format=torch.channels_last)
      b = torch.randn(32, device='cuda', dtype=dt)
      ref = F.relu(F.conv2d(x, w, b, 1, 1) + x)
      got = torch.cudnn_convolution_add_relu(x, w, x, 1.0, b, (1, 1), (1, 1), (1, 1), 1)
      got2 = torch.cudnn_convolution_relu(x, w, b, (1, 1), (1, 1), (1, 1), 1)
      if not (torch.allclose(got.float(), ref.float(), atol=0.25, rtol=0.05) and torch.allclose(got2.float(), F.relu(F.conv2d(x, w, b, 1, 1)).float(), atol=0.25, rtol=0.05)):
        return False
    return True
  except Exception:
    return False


_FUSED_CUDNN = False


class Bf16Plan(_PlanBase):
  def __init__(self, model, dtype=torch.bfloat16):
    global _FUSED_CUDNN
    self.dtype = dtype
    self.sphere_impl = 'bf16'
    self._cls_w = {}
    self._cv_w = None
    _FUSED_CUDNN = _probe_fused_cudnn()
    super().__init__(model)
    fe = model.feature_extraction
    if model.conv_type != 'Sphere':
      raise NotImplementedError("precision='bf16' supports conv='Sphere' (the MODE configuration); use precision='fp32' for conv='Regular'")
    fc = fe.firstconv
    c0 = fc[0][0]
    # firstconv[0] (3 -> 32, 7x7, stride 2): own tensor-core kernel reading the fp32 NCHW images directly (ops.stem_conv)
    self.stem = None
    if (c0.in_channels, c0.out_channels, c0.kernel_size, c0.stride, c0.padding, c0.dilation) == (3, 32, (7, 7), (2, 2), (3, 3), (1, 1)):
      self.stem = (_w(c0), *bn_affine(fc[0][1]))
    self.first = [_FoldedConv2d(fc[0][0], fc[0][1], dtype), _FoldedConv2d(fc[2][0], fc[2][1], dtype), _FoldedConv2d(fc[4][0], fc[4][1], dtype)]
    self.regular = []
    for layer in (fe.layer1, fe.layer2, fe.layer3):
      blocks = []
      for blk in layer:
        ds = _FoldedConv2d(blk.downsample[0], blk.downsample[1], dtype) if blk.downsample is not None else None
        c2 = _FoldedConv2d(blk.conv2[0], blk.conv2[1], dtype)
        if ds is not None:
          ds.move_bias_into(c2)
        blocks.append((_FoldedConv2d(blk.conv1[0][0], blk.conv1[0][1], dtype), c2, ds))
      self.regular.append(blocks)
    lc = fe.lastconv
    self.last = [_FoldedConv2d(lc[0][0], lc[0][1], dtype), _FoldedConv2d(lc[2][0], lc[2][1], dtype), _FoldedConv2d(lc[4][0], lc[4][1], dtype)]
    self.l4 = []
    for blk in fe.layer4:
      c1, b1 = blk.conv1[0][0], blk.conv1[0][1]
      c2, b2 = blk.conv2[0], blk.conv2[1]
      ds = _FoldedConv2d(blk.downsample[0], blk.downsample[1], dtype) if blk.downsample is not None else None
      if self.sphere_impl == 'bf16':
        s2, h2 = bn_affine(b2)
        if ds is not None:  # the downsample branch is only ever added to conv2's pre-activation: its BN shift rides in conv2's
          h2 = (h2 + ds.shift).contiguous()
          ds.shift, ds.b = None, None
        self.l4.append((c1, ops.sphere_conv_pack_weights(_w(c1), dtype), bn_affine(b1), ops.sphere_conv_pack_weights(_w(c2), dtype), (s2, h2), ds))
      else:
        self.l4.append((c1, _w(c1), bn_affine(b1), _w(c2), bn_affine(b2), ds))

  def pack_conv3d(self, w, scale, shift, mode):
    cout = w.shape[1] if mode == ops.DECONV_S2 else w.shape[0]
    return (ops.conv3d_pack_weights(w, mode, self.dtype), cout, scale, shift, mode)

  # ---- 2-D feature extractor: bf16 channels_last (physically NHWC) ------------------------------
  def _regular_layer(self, x, blocks):
    for c1, c2, ds in blocks:
      res = ds(x, False) if ds is not None else x
      x = c2(c1(x, True), True, residual=res)  # relu(conv2(relu(conv1(x))) + res), reference submodule.py:108-119
    return x

  def features(self, left, right=None):
    """left (B,3,H,W) [+ right (B,3,H,W)] -> (B or 2B, 32, H/4, W/4) features; the two batches are stacked along dim 0."""
    if self.stem is not None and left.shape[-1] <= 2048:
      w0, s0, h0 = self.stem
      x = ops.stem_conv(left.float(), None if right is None else right.float(), w0, s0, h0, True, self.dtype == torch.float16).permute(0, 3, 1, 2)
      rest = self.first[1:]
    else:
      x = left if right is None else torch.cat([left, right], 0)
      x = x.to(self.dtype).contiguous(memory_format=torch.channels_last)
      rest = self.first
    for c in rest:
      x = c(x, True)
    x = self._regular_layer(x, self.regular[0])
    raw = self._regular_layer(x, self.regular[1])
    reg = self._regular_layer(raw, self.regular[2])
    if self.sphere_impl == 'bf16':
      y = reg.permute(0, 2, 3, 1)  # NHWC view of the channels_last tensor (zero copy)
      if not y.is_contiguous():
        y = y.contiguous()
      for (c1, w1, (s1, h1), w2, (s2, h2), ds) in self.l4:
        pos = c1.position
        o = ops.sphere_conv_bf16(y, pos, w1, c1.out_channels, s1, h1, None, True)
        res = ds(y.permute(0, 3, 1, 2), False).permute(0, 2, 3, 1) if ds is not None else y
        y = ops.sphere_conv_bf16(o, pos, w2, c1.out_channels, s2, h2, res.contiguous(), True)
      if all(t.shape[1] % 8 == 0 for t in (raw, reg)) and y.shape[-1] % 8 == 0:
        # raw / reg are channels_last, i.e. NHWC in memory: stream the three maps (64 + 64 + 128) into the 256-channel operand of lastconv
        f = ops.concat3_nhwc(raw.permute(0, 2, 3, 1), reg.permute(0, 2, 3, 1), y).permute(0, 3, 1, 2)
        for c in self.last:
          f = c(f, True)
        return f
      sph = y.permute(0, 3, 1, 2)
    else:  # interim: fp32 CUDA-core sphere conv between bf16 neighbours
      y = reg.float().contiguous()
      for (c1, w1, (s1, h1), w2, (s2, h2), ds) in self.l4:
        pos = c1.position
        o = ops.sphere_conv_f32(y, pos, w1, s1, h1, None, True)
        res = ds(y.to(self.dtype).contiguous(memory_format=torch.channels_last), False).float().contiguous() if ds is not None else y
        y = ops.sphere_conv_f32(o, pos, w2, s2, h2, res, True)
      sph = y.to(self.dtype).contiguous(memory_format=torch.channels_last)
    f = torch.cat((raw, reg, sph), 1).contiguous(memory_format=torch.channels_last)
    f = self.last[0](f, True)
    f = self.last[1](f, True)
    f = self.last[2](f, True)
    return f  # (2B, 32, H/4, W/4) bf16 channels_last

  def cost_volume(self, fl, fr, d4):
    fl, fr = fl.permute(0, 2, 3, 1), fr.permute(0, 2, 3, 1)  # NHWC views
    return ops.cost_volume(fl.contiguous(), fr.contiguous(), d4)  # (B, D4, H4, W4, 64) bf16

  def first3d(self, fl, fr, d4):
    """Cost volume + dres0[0] fused (costvol_conv.cu): the 64-channel volume is never written.  Falls back to the two-kernel
    path for feature widths other than 32."""
    conv = self.model.dres0[0][0]
    if fl.shape[1] != 32 or tuple(conv.weight.shape) != (32, 64, 3, 3, 3) or os.environ.get('MODE_B200_NO_COSTVOL_FUSION'):
      return super().first3d(fl, fr, d4)
    if self._cv_w is None:
      self._cv_w = ops.costvol_conv_weights(_w(conv), self.dtype)
    _, _, scale, shift, _ = self.p3['dres0.0']
    fl, fr = fl.permute(0, 2, 3, 1).contiguous(), fr.permute(0, 2, 3, 1).contiguous()  # NHWC views of the channels_last features
    return None, ops.costvol_conv(fl, fr, self._cv_w[0], self._cv_w[1], scale, shift, d4, True)

  def conv3d(self, x, key, relu, residual=None):
    wp, cout, scale, shift, mode = self.p3[key]
    return ops.conv3d_bf16(x, wp, cout, scale, shift, residual, mode, relu, False)

  def logits(self, x, key, residual):
    w = getattr(self.model, key.split('.')[0])[2].weight
    if tuple(w.shape) == (1, 32, 3, 3, 3) and x.shape[-1] == 32:  # pointwise GEMM + shifted sum (conv3d_cls_tc.cu)
      if key not in self._cls_w:
        self._cls_w[key] = w.detach().float().contiguous()
      return ops.conv3d_classifier(x, self._cls_w[key], None if residual is None else residual[..., 0]).unsqueeze(-1)
    wp, cout, _, _, mode = self.p3[key]
    return ops.conv3d_bf16(x, wp, cout, None, None, residual, mode, False, True)  # (B, D4, H4, W4, 1) fp32

  def run(self, left, right, return_stages=False):
    B, _, H, W = left.shape
    feat = self.features(left, right)
    cost, c0 = self.first3d(feat[:B], feat[B:], self.maxdisp // 4)
    cost1, cost2, cost3 = self.regularise(c0)
    pred, conf = ops.disp_regress(cost3[..., 0], self.maxdisp, H, W)
    if return_stages:
      return pred, conf, dict(feat=feat, cost=cost, cost1=cost1[..., 0], cost2=cost2[..., 0], cost3=cost3[..., 0])
    return pred, conf
