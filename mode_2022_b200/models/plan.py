"""Execution plans for ModeDisparity: the folded parameters + the op graph of the hot path.

A plan is built once per (weights, precision) and reused; it is the "engine" under the reference-shaped
nn.Module.  Dataflow follows models/mode_disparity.py:98-185 and hourglass.forward (:27-46) exactly; every
conv+BN(+residual)(+ReLU) group of the reference is ONE kernel launch here.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .submodule import bn_affine

S1, S2, DECONV = ops.CONV_S1, ops.CONV_S2, ops.DECONV_S2


def _w(conv):
  return conv.weight.detach().float().contiguous()


class _PlanBase:
  """Shared graph; subclasses provide layout-specific `features`, `cost_volume`, `conv3d`, `logits_f32`."""

  def __init__(self, model):
    self.model = model
    self.maxdisp = model.maxdisp
    self.p3 = {}
    m = model
    seqs = {
        'dres0.0': (m.dres0[0], S1), 'dres0.2': (m.dres0[2], S1), 'dres1.0': (m.dres1[0], S1), 'dres1.2': (m.dres1[2], S1),
        'classif1.0': (m.classif1[0], S1), 'classif2.0': (m.classif2[0], S1), 'classif3.0': (m.classif3[0], S1),
    }
    for name in ('dres2', 'dres3', 'dres4'):
      hg = getattr(m, name)
      seqs.update({f'{name}.conv1': (hg.conv1[0], S2), f'{name}.conv2': (hg.conv2, S1), f'{name}.conv3': (hg.conv3[0], S2), f'{name}.conv4': (hg.conv4[0], S1),
                   f'{name}.conv5': (hg.conv5, DECONV), f'{name}.conv6': (hg.conv6, DECONV)})
    for key, (seq, mode) in seqs.items():
      scale, shift = bn_affine(seq[1])
      self.p3[key] = self.pack_conv3d(_w(seq[0]), scale, shift, mode)
    for i in (1, 2, 3):
      self.p3[f'classif{i}.2'] = self.pack_conv3d(_w(getattr(m, f'classif{i}')[2]), None, None, S1)

  # ---- graph -----------------------------------------------------------------------------------
  def hourglass(self, name, x, presqu, postsqu, c0):
    c = self.conv3d
    out = c(x, f'{name}.conv1', relu=True)
    pre = c(out, f'{name}.conv2', relu=True, residual=postsqu)
    out = c(pre, f'{name}.conv3', relu=True)
    out = c(out, f'{name}.conv4', relu=True)
    post = c(out, f'{name}.conv5', relu=True, residual=presqu if presqu is not None else pre)
    out = c(post, f'{name}.conv6', relu=False, residual=c0)  # "+ cost0" of mode_disparity.py:119,122,125 fused
    return out, pre, post

  def first3d(self, fl, fr, d4):
    """Cost volume + dres0[0] (reference mode_disparity.py:104-115); returns (cost or None, activation)."""
    cost = self.cost_volume(fl, fr, d4)
    return cost, self.conv3d(cost, 'dres0.0', relu=True)

  def regularise(self, c0):
    """3-D stack from the output of dres0[0] on."""
    c = self.conv3d
    c0 = c(c0, 'dres0.2', relu=True)
    t = c(c0, 'dres1.0', relu=True)
    c0 = c(t, 'dres1.2', relu=False, residual=c0)
    out1, pre1, post1 = self.hourglass('dres2', c0, None, None, c0)
    out2, pre2, post2 = self.hourglass('dres3', out1, pre1, post1, c0)
    out3, pre3, post3 = self.hourglass('dres4', out2, pre1, post2, c0)
    cost1 = self.logits(c(out1, 'classif1.0', relu=True), 'classif1.2', None)
    cost2 = self.logits(c(out2, 'classif2.0', relu=True), 'classif2.2', cost1)
    cost3 = self.logits(c(out3, 'classif3.0', relu=True), 'classif3.2', cost2)
    return cost1, cost2, cost3

  def run(self, left, right, return_stages=False):
    B, _, H, W = left.shape
    feat = self.features(torch.cat([left, right], 0))
    cost, c0 = self.first3d(feat[:B], feat[B:], self.maxdisp // 4)
    cost1, cost2, cost3 = self.regularise(c0)
    pred, conf = ops.disp_regress(cost3, self.maxdisp, H, W)
    if return_stages:
      return pred, conf, dict(feat=feat, cost=cost, cost1=cost1, cost2=cost2, cost3=cost3)
    return pred, conf


class Fp32Plan(_PlanBase):
  """Parity mode: NCHW / NCDHW fp32 everywhere."""

  def __init__(self, model):
    super().__init__(model)
    fe = model.feature_extraction
    self.l4 = []
    if model.conv_type == 'Sphere':
      for blk in fe.layer4:
        c1, b1 = blk.conv1[0][0], blk.conv1[0][1]
        c2, b2 = blk.conv2[0], blk.conv2[1]
        self.l4.append((c1, _w(c1), bn_affine(b1), _w(c2), bn_affine(b2), blk.downsample))

  def pack_conv3d(self, w, scale, shift, mode):
    return (w, scale, shift, mode)

  def features(self, x):
    fe = self.model.feature_extraction
    x = x.float()
    if self.model.conv_type != 'Sphere':
      return fe(x)
    x = fe.layer1(fe.firstconv(x))
    raw = fe.layer2(x)
    reg = fe.layer3(raw)
    y = reg
    for (c1, w1, (s1, h1), w2, (s2, h2), ds) in self.l4:
      pos = c1.position
      o = ops.sphere_conv_f32(y, pos, w1, s1, h1, None, True)
      res = ds(y) if ds is not None else y
      y = ops.sphere_conv_f32(o, pos, w2, s2, h2, res, True)
    return fe.lastconv(torch.cat((raw, reg, y), 1))

  def cost_volume(self, fl, fr, d4):
    return ops.cost_volume(fl.contiguous(), fr.contiguous(), d4)

  def conv3d(self, x, key, relu, residual=None):
    w, scale, shift, mode = self.p3[key]
    return ops.conv3d_f32(x, w, scale, shift, residual, mode, relu)

  def logits(self, x, key, residual):
    w, _, _, mode = self.p3[key]
    return ops.conv3d_f32(x, w, None, None, residual, mode, False)


def build_plan(model):
  if model.precision == 'fp32':
    return Fp32Plan(model)
  import torch
  from .plan_bf16 import Bf16Plan
  return Bf16Plan(model, torch.float16 if model.precision == 'fp16' else torch.bfloat16)
