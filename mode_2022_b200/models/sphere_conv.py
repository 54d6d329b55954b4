"""SphereConv -- drop-in for models/basic/spherical_conv/sphere_conv.py of the reference.

Same constructor, attributes (`weight`, `bias`, `position`, `getPosition()`), state-dict keys (only
`weight`[/`bias`]; `position` is a plain attribute, not a buffer -- reference sphere_conv.py:150,156) and the
same functional entry point `sphere_conv(input, position, weight, bias, stride, padding, dilation, groups)`
(reference: SphereConvFunction.apply, sphere_conv.py:117).  The compute goes to libmode_b200
(`mode_sphere_conv_f32` / `mode_sphere_conv_bf16`) instead of im2col + addmm.
"""
from __future__ import annotations

import collections
import math
from functools import lru_cache

import numpy as np
import torch
import torch.nn as nn
from torch.nn.modules.utils import _pair, _single

from .. import ops


def _tap_pattern(Kh: int, Kw: int, height: int, width: int):
  """Tangent-plane (gnomonic) offsets of the Kh x Kw taps: kerX, kerY, rho, cos(nu), sin(nu)  (fp64)."""
  d_lat, d_lon = np.pi / height, 2 * np.pi / width
  ax = [np.arange(-(k // 2), k // 2 + 1) for k in (Kw, Kh)]
  ax = [a if k % 2 else np.delete(a, k // 2) for a, k in zip(ax, (Kw, Kh))]
  tx = np.tan(ax[0] * d_lon)
  ty = np.tan(ax[1] * d_lat) / np.cos(ax[1] * d_lon)
  kx, ky = np.meshgrid(tx, ty)
  rho = np.sqrt(kx**2 + ky**2)
  if Kh % 2 and Kw % 2:
    rho[Kh // 2][Kw // 2] = 1e-8  # centre tap: avoid 0/0 (reference sphere_conv.py:198-199)
  nu = np.arctan(rho)
  return kx, ky, rho, np.cos(nu), np.sin(nu)


@lru_cache(maxsize=None)
def sphere_position_numpy(height: int, width: int, sphere_type: str, Kh: int = 3, Kw: int = 3) -> np.ndarray:
  """Sampling grid (1, 2*Kh*Kw, H, W) fp32 for a sphere image whose short side is `height` (latitude) and long
  side `width` (longitude).  Must be BIT-EXACT with SphereConv.gen_sphere_position (reference
  sphere_conv.py:180-237): everything is fp64 numpy with the reference's per-element operation order, and the
  per-latitude trigonometry is evaluated on numpy *scalars* as the reference does (array and scalar sin/cos may
  differ in the last ulp).  Channel 2k is the row coordinate of tap k, channel 2k+1 the column coordinate."""
  kx, ky, rho, cos_nu, sin_nu = _tap_pattern(Kh, Kw, height, width)
  lats = ((np.arange(0, height, 1) / height) - 0.5) * np.pi
  lons = ((np.arange(0, width, 1) / width) - 0.5) * (2 * np.pi)
  lat_rows, dlon_rows = [], []
  for phi in lats:  # inverse gnomonic projection of the tap pattern centred at latitude phi
    s, c = np.sin(phi), np.cos(phi)
    lat_rows.append(np.arcsin(cos_nu * s + ky * sin_nu * c / rho))
    dlon_rows.append(np.arctan2(kx * sin_nu, (rho * c * cos_nu - ky * s * sin_nu)))
  lat = np.broadcast_to(np.array(lat_rows)[:, None], (height, width, Kh, Kw))
  lon = np.array(dlon_rows)[:, None] + lons[None, :, None, None]
  lat = (lat / np.pi + 0.5) * height
  lon = ((lon / (2 * np.pi) + 0.5) * width) % width
  if sphere_type == 'ERP':  # rows = latitude
    grid = np.stack((lat, lon)).astype(np.float32).transpose((3, 4, 0, 1, 2))
  else:  # Cassini: rows = longitude, image is the transpose
    grid = np.stack((lon, lat)).astype(np.float32).transpose((3, 4, 0, 2, 1))
  kh, kw, two, H, W = grid.shape
  return np.ascontiguousarray(grid.reshape(1, two * kh * kw, H, W))


_DEVICE_POS = collections.OrderedDict()  # LRU-bounded: one grid per (resolution, projection, kernel, device) in use
_DEVICE_POS_MAX = 16


def sphere_position(height, width, sphere_type, Kh, Kw, device) -> torch.Tensor:
  """Device-resident, shared copy of the grid (the reference keeps one copy per layer: 16 x 2.4 MB).  The upload is a
  synchronous host-to-device copy of pageable memory, so the tensor is complete for every stream when this returns."""
  key = (height, width, sphere_type, Kh, Kw, str(device))
  pos = _DEVICE_POS.get(key)
  if pos is None:
    pos = _DEVICE_POS[key] = torch.from_numpy(sphere_position_numpy(height, width, sphere_type, Kh, Kw)).to(device)
    if torch.device(device).type == 'cuda' and not torch.cuda.is_current_stream_capturing():
      torch.cuda.current_stream(device).synchronize()
    while len(_DEVICE_POS) > _DEVICE_POS_MAX:
      _DEVICE_POS.popitem(last=False)
  else:
    _DEVICE_POS.move_to_end(key)
  return pos


def sphere_conv(input, position, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
  """Functional form with the reference signature (sphere_conv.py:18,117).  As in the reference kernels,
  padding/dilation do not influence the sampling; only stride 1 / groups 1 (what MODE uses) are implemented."""
  if input is not None and input.dim() != 4:
    raise ValueError('Expected 4D tensor as input, got {}D tensor instead.'.format(input.dim()))
  if not input.is_cuda:
    raise NotImplementedError('Only support cuda tensor!')
  if _pair(stride) != (1, 1) or groups != 1:
    raise NotImplementedError('sphere_conv: only stride=1, groups=1 are supported')
  kh, kw = weight.shape[2:]
  ph, pw = _pair(padding)
  dh, dw = _pair(dilation)
  if (2 * ph - dh * (kh - 1), 2 * pw - dw * (kw - 1)) != (0, 0):
    raise NotImplementedError('sphere_conv: only "same" output size is supported')
  if torch.is_grad_enabled() and (input.requires_grad or weight.requires_grad or (bias is not None and bias.requires_grad)):
    return SphereConvFunction.apply(input.float(), position, weight, bias)
  return ops.sphere_conv_f32(input.float(), position, weight, None, bias, None, False)


class SphereConvFunction(torch.autograd.Function):
  """Autograd node of the reference (sphere_conv.py:16-90): forward and backward both go to libmode_b200; gradients flow to
  input, weight and bias, never to `position` (reference backward returns None for it, sphere_conv.py:89)."""

  @staticmethod
  def forward(ctx, input, position, weight, bias):
    ctx.save_for_backward(input, position, weight)
    ctx.has_bias = bias is not None
    return ops.sphere_conv_f32(input, position, weight, None, bias, None, False)

  @staticmethod
  @torch.autograd.function.once_differentiable
  def backward(ctx, grad_output):
    input, position, weight = ctx.saved_tensors
    if not grad_output.is_cuda:
      raise NotImplementedError  # sphere_conv.py:66-67
    gi, gw, gb = ops.sphere_conv_backward_f32(input, position, weight, grad_output.contiguous(), ctx.needs_input_grad[0], ctx.needs_input_grad[2],
                                               ctx.has_bias and ctx.needs_input_grad[3])
    return gi, None, gw, gb


class SphereConv(nn.Module):
  def __init__(self, in_height, in_width, sphereType, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=False):
    super().__init__()
    assert (sphereType is not None) and (sphereType in ['Cassini', 'ERP'])
    assert (in_height is not None) and (in_height > 0)
    assert (in_width is not None) and (in_width > 0)
    assert in_channels % groups == 0, 'in_channels {} cannot be divisible by groups {}'.format(in_channels, groups)
    assert out_channels % groups == 0, 'out_channels {} cannot be divisible by groups {}'.format(out_channels, groups)
    in_h, in_w = min(in_height, in_width), max(in_height, in_width)
    assert in_w == 2 * in_h
    self.in_height, self.in_width = in_h, in_w
    self.in_channels, self.out_channels = in_channels, out_channels
    self.kernel_size, self.stride = _pair(kernel_size), _pair(stride)
    self.padding, self.dilation = _pair(padding), _pair(dilation)
    self.groups, self.sphereType = groups, sphereType
    self.transposed, self.output_padding = False, _single(0)  # nn.Conv2d compatibility, as the reference
    self.input_size = (1, in_channels, in_h, in_w)
    self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, *self.kernel_size))
    if bias:
      self.bias = nn.Parameter(torch.zeros(out_channels))
    else:
      self.register_parameter('bias', None)
    self.reset_parameters()

  def reset_parameters(self):
    n = self.in_channels * self.kernel_size[0] * self.kernel_size[1]
    stdv = 1. / math.sqrt(n)
    self.weight.data.uniform_(-stdv, stdv)

  @property
  def position(self) -> torch.Tensor:
    return sphere_position(self.in_height, self.in_width, self.sphereType, *self.kernel_size, self.weight.device)

  def getPosition(self):
    return self.position

  def forward(self, x):
    return sphere_conv(x, self.position, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)

  def extra_repr(self):
    return f'{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, sphereType={self.sphereType}, grid={self.in_height}x{self.in_width}'
