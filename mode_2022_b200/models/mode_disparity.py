"""ModeDisparity -- drop-in for the reference models/mode_disparity.py (same constructor, forward signature,
output shapes and state-dict keys), executing the stereo hot path on libmode_b200 kernels:

  feature extraction   regular 2-D layers: cuDNN; layer4: spherical conv kernel (a2)
  cost volume          mode_cost_volume_*                                   (a4, reference :104-113)
  3-D regularisation   mode_conv3d_* with BN/residual/ReLU fused            (a5, reference :11-46,:115-129)
  regression + conf    mode_disp_regress                                    (a6/a7, reference :143-183)

`precision='fp32'` is the parity mode (NCHW/NCDHW fp32 everywhere, CUDA-core kernels); `precision='bf16'` is the
throughput mode BASELINE.json names (NHWC/NDHWC bf16 activations, tcgen05 tensor-core kernels with fp32 accumulation
and fp32 logits/regression); `precision='fp16'` runs the same kernels at the same speed with fp16 storage (3 more
mantissa bits: 8x less rounding noise per stored activation).  Left and right images share the feature
extractor, so they are run as one batch of 2B.

In training mode (`.train()`) forward returns the three supervised heads (pred1, pred2, pred3) like the reference and is
differentiable: the spherical layers use libmode_b200's forward and backward kernels, the remaining layers the library
modules the reference itself trains with.  The fused inference plans are eval-only.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .batchnorm import BatchNorm3d
from .submodule import bn_affine, convbn, convbn_3d, sphere_feature_extraction


class hourglass(nn.Module):
  """Parameter holder with the reference's layout (mode_disparity.py:11-25); executed by ModeDisparity."""

  def __init__(self, inplanes):
    super().__init__()
    self.conv1 = nn.Sequential(convbn_3d(inplanes, inplanes * 2, kernel_size=3, stride=2, pad=1), nn.ReLU(inplace=True))
    self.conv2 = convbn_3d(inplanes * 2, inplanes * 2, kernel_size=3, stride=1, pad=1)
    self.conv3 = nn.Sequential(convbn_3d(inplanes * 2, inplanes * 2, kernel_size=3, stride=2, pad=1), nn.ReLU(inplace=True))
    self.conv4 = nn.Sequential(convbn_3d(inplanes * 2, inplanes * 2, kernel_size=3, stride=1, pad=1), nn.ReLU(inplace=True))
    self.conv5 = nn.Sequential(nn.ConvTranspose3d(inplanes * 2, inplanes * 2, kernel_size=3, padding=1, output_padding=1, stride=2, bias=False),
                               BatchNorm3d(inplanes * 2))
    self.conv6 = nn.Sequential(nn.ConvTranspose3d(inplanes * 2, inplanes, kernel_size=3, padding=1, output_padding=1, stride=2, bias=False), BatchNorm3d(inplanes))

  def forward(self, x, presqu, postsqu):
    """Training path (reference mode_disparity.py:27-46)."""
    pre = self.conv2(self.conv1(x))
    pre = F.relu(pre if postsqu is None else pre + postsqu)
    out = self.conv4(self.conv3(pre))
    post = F.relu(self.conv5(out) + (pre if presqu is None else presqu))
    return self.conv6(post), pre, post


class ModeDisparity(nn.Module):
  def __init__(self, maxdisp, conv='Sphere', in_height=1024, in_width=512, sphereType='Cassini', out_conf=False, precision=None):
    super().__init__()
    self.maxdisp = maxdisp
    self.out_conf = out_conf
    self.train_channels_last = os.environ.get('MODE_B200_TRAIN_CHANNELS_LAST', '1') != '0'  # memory format of the 3-D stack in training (values unchanged)
    self.sphereType = sphereType
    self.precision = precision or os.environ.get('MODE_B200_PRECISION', 'fp16')
    if self.precision not in ('fp32', 'bf16', 'fp16'):
      raise ValueError("precision must be 'fp32', 'bf16' or 'fp16'")
    if conv == 'Regular':
      from .psm_features import feature_extraction
      self.feature_extraction = feature_extraction()
    elif conv == 'Sphere':
      self.feature_extraction = sphere_feature_extraction(in_height, in_width, sphereType)
    else:
      raise NotImplementedError('Convolution Type must be Regular or Sphere!')
    self.conv_type = conv

    self.dres0 = nn.Sequential(convbn_3d(64, 32, 3, 1, 1), nn.ReLU(inplace=True), convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True))
    self.dres1 = nn.Sequential(convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True), convbn_3d(32, 32, 3, 1, 1))
    self.dres2 = hourglass(32)
    self.dres3 = hourglass(32)
    self.dres4 = hourglass(32)
    self.classif1 = nn.Sequential(convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True), nn.Conv3d(32, 1, kernel_size=3, padding=1, stride=1, bias=False))
    self.classif2 = nn.Sequential(convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True), nn.Conv3d(32, 1, kernel_size=3, padding=1, stride=1, bias=False))
    self.classif3 = nn.Sequential(convbn_3d(32, 32, 3, 1, 1), nn.ReLU(inplace=True), nn.Conv3d(32, 1, kernel_size=3, padding=1, stride=1, bias=False))

    for m in self.modules():  # reference init, mode_disparity.py:82-96
      if isinstance(m, nn.Conv2d):
        n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
        m.weight.data.normal_(0, math.sqrt(2. / n))
      elif isinstance(m, nn.Conv3d):
        n = m.kernel_size[0] * m.kernel_size[1] * m.kernel_size[2] * m.out_channels
        m.weight.data.normal_(0, math.sqrt(2. / n))
      elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
        m.weight.data.fill_(1)
        m.bias.data.zero_()
    self._plan, self._plan_key = None, None
    self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_plan())

  # -------------------------------------------------------------------------------------------
  # folded-parameter cache ("plan"): BN affines, packed bf16 weights; rebuilt after weights change
  # -------------------------------------------------------------------------------------------
  def invalidate_plan(self):
    self._plan = None

  def _weights_key(self):
    """Cheap fingerprint of the weights a plan was folded from: storage addresses + in-place version counters."""
    k = 0
    for t in self.parameters():
      k = (k * 1000003 + t._version + t.data_ptr()) & 0xFFFFFFFFFFFF
    for t in self.buffers():
      k = (k * 1000003 + t._version + t.data_ptr()) & 0xFFFFFFFFFFFF
    return k

  def train(self, mode: bool = True):
    self._plan = None
    return super().train(mode)

  def _apply(self, fn, *a, **k):
    self._plan = None
    return super()._apply(fn, *a, **k)

  def load_state_dict(self, state_dict, *a, **k):
    # accept checkpoints saved from an nn.DataParallel wrapper ('module.' prefix; reference train_disparity.py:93)
    if state_dict and all(key.startswith('module.') for key in state_dict):
      state_dict = {key[7:]: v for key, v in state_dict.items()}
    return super().load_state_dict(state_dict, *a, **k)

  def _build_plan(self):
    from .plan import build_plan
    return build_plan(self)

  # -------------------------------------------------------------------------------------------
  def forward(self, left, right):
    if not left.is_cuda:
      raise NotImplementedError('ModeDisparity (B200) runs on CUDA tensors only')
    if left.shape != right.shape or left.dim() != 4 or left.shape[2] % 16 or left.shape[3] % 16:
      raise ValueError('left/right must be (B,3,H,W) with H, W multiples of 16')
    if self.maxdisp % 16:
      raise ValueError('maxdisp must be a multiple of 16')
    if self.training:
      return self._forward_train(left, right)
    if torch.is_grad_enabled() and (left.requires_grad or right.requires_grad):
      raise RuntimeError('ModeDisparity.eval() runs a fused, non-differentiable inference plan; call .train() for a differentiable forward')
    key = self._weights_key()
    if self._plan is None or self._plan_key != key:  # in-place weight edits (optimizer.step(), p.data.copy_()) bump tensor versions
      self._plan, self._plan_key = self._build_plan(), key
    with torch.no_grad():
      pred3, conf = self._plan.run(left, right)
    if self.out_conf:
      return pred3, conf
    return pred3

  # -------------------------------------------------------------------------------------------
  def _forward_train(self, left, right):
    """Training graph (reference mode_disparity.py:99-155): batch-stat BatchNorm, three supervised heads, differentiable end
    to end in fp32.  libmode_b200 forward AND backward kernels: the spherical layers (SphereConvFunction, SURVEY.md section 8
    a3), the cost volume and the three soft-argmin heads; the conv2d / conv3d / BatchNorm layers are the library code the
    reference trains with (cuDNN through autograd)."""
    left, right = left.float(), right.float()
    if self.train_channels_last:  # the regular 2-D layers run NHWC end to end (cuDNN's native layout); the spherical layers convert back
      left, right = left.contiguous(memory_format=torch.channels_last), right.contiguous(memory_format=torch.channels_last)
    fl = self.feature_extraction(left)
    fr = self.feature_extraction(right)
    d4 = self.maxdisp // 4
    # integer shifts: cost[:, :C, i, :, i:] = ref[..., i:], cost[:, C:, i, :, i:] = tgt[..., :W-i] -- one kernel; its backward is a
    # deterministic gather-sum over the shifts (mode_cost_volume_backward_f32), registered on the op with torch.library
    cost = ops.cost_volume(fl.contiguous(), fr.contiguous(), d4)
    if self.train_channels_last:
      # NDHWC for the whole 3-D stack: cuDNN's tensor-core conv3d kernels (fprop / dgrad / wgrad) are NDHWC kernels -- on NCDHW
      # tensors every call is wrapped in layout transforms (762 of them, 14 ms of a 117 ms step); BatchNorm3d (models/batchnorm.py)
      # normalises channels-last tensors in place, ReLU and the skip additions keep the format.  Values are unchanged.
      cost = cost.contiguous(memory_format=torch.channels_last_3d)
    cost0 = self.dres0(cost)
    cost0 = self.dres1(cost0) + cost0
    out1, pre1, post1 = self.dres2(cost0, None, None)
    out1 = out1 + cost0
    out2, _, post2 = self.dres3(out1, pre1, post1)
    out2 = out2 + cost0
    out3, _, _ = self.dres4(out2, pre1, post2)  # the third hourglass reuses pre1 (reference :124)
    out3 = out3 + cost0
    cost1 = self.classif1(out1)
    cost2 = self.classif2(out2) + cost1
    cost3 = self.classif3(out3) + cost2
    # the three supervised heads (reference :131-152): trilinear upsample + softmax + soft-argmin fused, forward AND backward
    # (mode_disp_regress / mode_disp_regress_backward): the reference materialises three (B,D,H,W) volumes per head for autograd
    preds = [ops.disp_regress(c, self.maxdisp, left.shape[2], left.shape[3])[0] for c in (cost1, cost2, cost3)]
    return tuple(preds)
