"""CPU oracle for the MODE stereo hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain numpy / torch-CPU *restatement* of the reference algorithm
(nju-ee/MODE-2022).  It is the checker for the CUDA kernels: only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it.  The product package (`mode_2022_b200`) never does.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against the reference *itself*: `oracle/pin_against_reference.py`
imports the reference Python from /root/reference (with the two CPU shims of
SURVEY.md §8c), runs both on identical seeded inputs, asserts agreement and
writes the small fixtures under `tests/golden/`.  On the GPU box the sphere-conv
restatement is additionally checked against the *compiled, unmodified*
reference CUDA op (`oracle/_ref/`, built by `oracle/build_ref.py`).

Every function cites the reference file:line it restates (paths relative to
the reference repo root).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# a1. sampling grid  --  models/basic/spherical_conv/sphere_conv.py:180-237
# ----------------------------------------------------------------------------


def gen_sphere_position(in_height: int, in_width: int, sphere_type: str, kernel_size=(3, 3)) -> np.ndarray:
  """Tangent-plane sampling grid, fp64 numpy -> fp32, shape (1, 2*Kh*Kw, H, W).

  Restates SphereConv.__init__ (sphere_conv.py:131-139: short side becomes
  `height`, long side `width`) + gen_sphere_position (sphere_conv.py:180-237).
  The arithmetic order is kept identical so the result is bit-exact.
  """
  height = min(in_height, in_width)  # sphere_conv.py:131-136
  width = max(in_height, in_width)
  assert width == 2 * height
  Kh, Kw = kernel_size
  delta_lat = np.pi / height
  delta_lon = 2 * np.pi / width
  range_x = np.arange(-(Kw // 2), Kw // 2 + 1)
  if not Kw % 2:
    range_x = np.delete(range_x, Kw // 2)
  range_y = np.arange(-(Kh // 2), Kh // 2 + 1)
  if not Kh % 2:
    range_y = np.delete(range_y, Kh // 2)
  kerX = np.tan(range_x * delta_lon)
  kerY = np.tan(range_y * delta_lat) / np.cos(range_y * delta_lon)
  kerX, kerY = np.meshgrid(kerX, kerY)
  rho = np.sqrt(kerX**2 + kerY**2)
  if Kh % 2 and Kw % 2:
    rho[Kh // 2][Kw // 2] = 1e-8
  nu = np.arctan(rho)
  cos_nu = np.cos(nu)
  sin_nu = np.sin(nu)
  h_range = np.arange(0, height, 1)
  w_range = np.arange(0, width, 1)
  lat_range = ((h_range / height) - 0.5) * np.pi
  lon_range = ((w_range / width) - 0.5) * (2 * np.pi)
  lat = np.array([np.arcsin(cos_nu * np.sin(_lat) + kerY * sin_nu * np.cos(_lat) / rho) for _lat in lat_range])
  lat = np.array([lat for _ in lon_range]).transpose((1, 0, 2, 3))  # (H, W, Kh, Kw)
  lon = np.array([np.arctan2(kerX * sin_nu, (rho * np.cos(_lat) * cos_nu - kerY * np.sin(_lat) * sin_nu)) for _lat in lat_range])
  lon = np.array([lon + _lon for _lon in lon_range]).transpose((1, 0, 2, 3))
  lat = (lat / np.pi + 0.5) * height
  lon = ((lon / (2 * np.pi) + 0.5) * width) % width
  if sphere_type == 'ERP':
    ll = np.stack((lat, lon)).astype(np.float32).transpose((3, 4, 0, 1, 2))
  elif sphere_type == 'Cassini':
    ll = np.stack((lon, lat)).astype(np.float32).transpose((3, 4, 0, 2, 1))
  else:
    raise AssertionError(sphere_type)
  kh, kw, d, H, W = ll.shape
  return np.ascontiguousarray(ll.reshape((1, d * kh * kw, H, W)))


# ----------------------------------------------------------------------------
# a2. spherical convolution forward
#   sphere_conv_cuda_kernel.cu:83-113 (bilinear), :195-262 (im2col),
#   sphere_conv_cuda.cpp:129-210 (per-element GEMM)
# ----------------------------------------------------------------------------


def sphere_im2col(x: torch.Tensor, pos: torch.Tensor, kh: int = 3, kw: int = 3) -> torch.Tensor:
  """columns[b, c*KhKw + k, h, w]; stride 1 (the only configuration MODE uses)."""
  B, C, H, W = x.shape
  K = kh * kw
  p = pos.reshape(K, 2, H, W).to(x.dtype)
  h_im, w_im = p[:, 0], p[:, 1]  # kernel.cu:236-244
  valid = (h_im > -1) & (w_im > -1) & (h_im < H) & (w_im < W)  # kernel.cu:246
  h_low = torch.floor(h_im)
  w_low = torch.floor(w_im)
  lh = h_im - h_low
  lw = w_im - w_low
  hh, hw = 1 - lh, 1 - lw
  h_low = h_low.long()
  w_low = w_low.long()
  h_high, w_high = h_low + 1, w_low + 1
  flat = x.reshape(B, C, H * W)

  def corner(hi, wi, ok):
    ok = ok & valid
    idx = (hi.clamp(0, H - 1) * W + wi.clamp(0, W - 1)).reshape(-1)
    v = flat[:, :, idx].reshape(B, C, K, H, W)
    return v * ok.to(x.dtype)

  v1 = corner(h_low, w_low, (h_low >= 0) & (w_low >= 0))  # kernel.cu:97-107
  v2 = corner(h_low, w_high, (h_low >= 0) & (w_high <= W - 1))
  v3 = corner(h_high, w_low, (h_high <= H - 1) & (w_low >= 0))
  v4 = corner(h_high, w_high, (h_high <= H - 1) & (w_high <= W - 1))
  w1, w2, w3, w4 = hh * hw, hh * lw, lh * hw, lh * lw  # kernel.cu:109
  val = w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4  # kernel.cu:111
  return val.reshape(B, C * K, H, W)


def sphere_conv(x, pos, weight, bias=None):
  """out[b] = W.flatten(1) @ columns[b]   (sphere_conv_cuda.cpp:191-196)."""
  B, C, H, W = x.shape
  Co, _, kh, kw = weight.shape
  cols = sphere_im2col(x, pos, kh, kw).reshape(B, C * kh * kw, H * W)
  out = torch.matmul(weight.reshape(Co, -1), cols).reshape(B, Co, H, W)
  if bias is not None:
    out = out + bias.view(1, -1, 1, 1)  # cpp:207-209
  return out


# ----------------------------------------------------------------------------
# a4. cost volume  --  models/mode_disparity.py:104-113
# ----------------------------------------------------------------------------


def cost_volume(ref: torch.Tensor, tgt: torch.Tensor, d4: int) -> torch.Tensor:
  B, C, H, W = ref.shape
  cost = torch.zeros(B, 2 * C, d4, H, W, dtype=ref.dtype)
  for i in range(d4):
    if i > 0:
      cost[:, :C, i, :, i:] = ref[:, :, :, i:]
      cost[:, C:, i, :, i:] = tgt[:, :, :, :-i]
    else:
      cost[:, :C, i] = ref
      cost[:, C:, i] = tgt
  return cost


# ----------------------------------------------------------------------------
# a6/a7. upsample + softmax + soft-argmin + confidence
#   models/mode_disparity.py:131-152 (regression), :157-183 (confidence),
#   models/submodule.py:50-57 (disparityregression)
# ----------------------------------------------------------------------------


def disparity_regression(cost: torch.Tensor, maxdisp: int, H: int, W: int, want_conf: bool = False):
  """cost: (B,1,D/4,H/4,W/4) -> pred (B,1,H,W) [, conf (B,1,H,W)]."""
  up = F.interpolate(cost, [maxdisp, H, W], mode='trilinear', align_corners=True)
  p = F.softmax(up.squeeze(1), dim=1)
  disp = torch.arange(maxdisp, dtype=p.dtype).view(1, maxdisp, 1, 1)
  pred = torch.sum(p * disp, 1, keepdim=True)  # submodule.py:55-57
  if not want_conf:
    return pred
  # closed form of the three nearest-mode, border-padded grid_samples
  # (mode_disparity.py:159-181); verified bit-identical in pin_against_reference.py
  r = torch.round(pred).long()
  conf = (p.gather(1, r.clamp(0, maxdisp - 1)) + p.gather(1, (r - 1).clamp(0, maxdisp - 1)) + p.gather(1, (r + 1).clamp(0, maxdisp - 1)))
  return pred, conf


# ----------------------------------------------------------------------------
# full ModeDisparity forward from a state dict (functional restatement)
#   models/submodule.py:15-22, 94-201; models/mode_disparity.py:11-46, 98-185
# ----------------------------------------------------------------------------


class _SD:
  def __init__(self, sd, training_bn=False):
    self.sd = {(k[7:] if k.startswith('module.') else k): v for k, v in sd.items()}
    self.training_bn = training_bn

  def bn(self, x, key):
    s = self.sd
    if self.training_bn:  # calibration pass: running stats := this batch's stats (momentum 1)
      return F.batch_norm(x, s[key + '.running_mean'], s[key + '.running_var'], s[key + '.weight'], s[key + '.bias'], True, 1.0, 1e-5)
    return F.batch_norm(x, s[key + '.running_mean'], s[key + '.running_var'], s[key + '.weight'], s[key + '.bias'], False, 0.0, 1e-5)

  def convbn2d(self, x, key, stride, pad, dil):
    x = F.conv2d(x, self.sd[key + '.0.weight'], None, stride, dil if dil > 1 else pad, dil)
    return self.bn(x, key + '.1')

  def sconvbn(self, x, key, pos):
    return self.bn(sphere_conv(x, pos, self.sd[key + '.0.weight']), key + '.1')

  def convbn3d(self, x, key, stride):
    return self.bn(F.conv3d(x, self.sd[key + '.0.weight'], None, stride, 1), key + '.1')

  def deconvbn3d(self, x, key):
    return self.bn(F.conv_transpose3d(x, self.sd[key + '.0.weight'], None, 2, 1, 1), key + '.1')


def feature_extraction(m: _SD, x: torch.Tensor, pos: torch.Tensor, pre='feature_extraction.'):
  """sphere_feature_extraction.forward (submodule.py:192-201)."""
  x = F.relu(m.convbn2d(x, pre + 'firstconv.0', 2, 3, 1))
  x = F.relu(m.convbn2d(x, pre + 'firstconv.2', 1, 1, 1))
  x = F.relu(m.convbn2d(x, pre + 'firstconv.4', 1, 1, 1))

  def regular_layer(x, name, nblocks, stride, dil):
    for b in range(nblocks):
      k = f'{pre}{name}.{b}'
      s = stride if b == 0 else 1
      out = F.relu(m.convbn2d(x, k + '.conv1.0', s, 1, dil))
      out = m.convbn2d(out, k + '.conv2', 1, 1, dil)
      if (k + '.downsample.0.weight') in m.sd:
        x = m.bn(F.conv2d(x, m.sd[k + '.downsample.0.weight'], None, s), k + '.downsample.1')
      x = F.relu(out + x)  # submodule.py:116-117
    return x

  x = regular_layer(x, 'layer1', 3, 1, 1)
  raw = regular_layer(x, 'layer2', 8, 2, 1)
  reg = regular_layer(raw, 'layer3', 4, 1, 2)
  x = reg
  for b in range(8):  # SphereBasicBlock, submodule.py:122-147
    k = f'{pre}layer4.{b}'
    out = F.relu(m.sconvbn(x, k + '.conv1.0', pos))
    out = m.sconvbn(out, k + '.conv2', pos)
    if (k + '.downsample.0.weight') in m.sd:
      x = m.bn(F.conv2d(x, m.sd[k + '.downsample.0.weight']), k + '.downsample.1')
    x = F.relu(out + x)
  f = torch.cat((raw, reg, x), 1)
  f = F.relu(m.convbn2d(f, pre + 'lastconv.0', 1, 0, 1))
  f = F.relu(m.convbn2d(f, pre + 'lastconv.2', 1, 1, 1))
  f = F.relu(m.convbn2d(f, pre + 'lastconv.4', 1, 0, 1))
  return f


def hourglass(m: _SD, key: str, x, presqu, postsqu):
  """hourglass.forward (mode_disparity.py:27-46)."""
  out = F.relu(m.convbn3d(x, key + '.conv1.0', 2))
  pre = m.convbn3d(out, key + '.conv2', 1)
  pre = F.relu(pre + postsqu) if postsqu is not None else F.relu(pre)
  out = F.relu(m.convbn3d(pre, key + '.conv3.0', 2))
  out = F.relu(m.convbn3d(out, key + '.conv4.0', 1))
  if presqu is not None:
    post = F.relu(m.deconvbn3d(out, key + '.conv5') + presqu)
  else:
    post = F.relu(m.deconvbn3d(out, key + '.conv5') + pre)
  out = m.deconvbn3d(post, key + '.conv6')
  return out, pre, post


def regularise(m: _SD, cost: torch.Tensor):
  """dres0..classif3 (mode_disparity.py:115-129) -> cost1, cost2, cost3 at 1/4 res."""
  c0 = F.relu(m.convbn3d(cost, 'dres0.0', 1))
  c0 = F.relu(m.convbn3d(c0, 'dres0.2', 1))
  t = F.relu(m.convbn3d(c0, 'dres1.0', 1))
  c0 = m.convbn3d(t, 'dres1.2', 1) + c0
  out1, pre1, post1 = hourglass(m, 'dres2', c0, None, None)
  out1 = out1 + c0
  out2, pre2, post2 = hourglass(m, 'dres3', out1, pre1, post1)
  out2 = out2 + c0
  out3, pre3, post3 = hourglass(m, 'dres4', out2, pre1, post2)
  out3 = out3 + c0

  def classif(x, key):
    x = F.relu(m.convbn3d(x, key + '.0', 1))
    return F.conv3d(x, m.sd[key + '.2.weight'], None, 1, 1)

  cost1 = classif(out1, 'classif1')
  cost2 = classif(out2, 'classif2') + cost1
  cost3 = classif(out3, 'classif3') + cost2
  return cost1, cost2, cost3


def mode_disparity_forward(sd, left, right, maxdisp, sphere_type='Cassini', out_conf=True, stages=False):
  """Eval-mode ModeDisparity.forward (mode_disparity.py:98-185) on CPU."""
  m = _SD(sd)
  H, W = left.shape[2:]
  pos = torch.from_numpy(gen_sphere_position(H // 4, W // 4, sphere_type))
  with torch.no_grad():
    fl = feature_extraction(m, left, pos)
    fr = feature_extraction(m, right, pos)
    cost = cost_volume(fl, fr, maxdisp // 4)
    cost1, cost2, cost3 = regularise(m, cost)
    res = disparity_regression(cost3, maxdisp, H, W, want_conf=out_conf)
  if stages:
    return res, dict(feat_l=fl, feat_r=fr, cost=cost, cost1=cost1, cost2=cost2, cost3=cost3)
  return res


def calibrate_bn(sd, left, right, maxdisp, sphere_type='Cassini'):
  """BN calibration (SURVEY.md hard part 5): one train-mode-BN pass that overwrites every
  running_mean/var with the statistics of this input, so that a random-init network is
  un-saturated in eval mode.  Returns a new state dict; conv weights are untouched."""
  sd = {k: v.clone() for k, v in sd.items()}
  m = _SD(sd, training_bn=True)
  sd = m.sd
  H, W = left.shape[2:]
  pos = torch.from_numpy(gen_sphere_position(H // 4, W // 4, sphere_type))
  with torch.no_grad():
    fl = feature_extraction(m, left, pos)
    fr = feature_extraction(m, right, pos)
    regularise(m, cost_volume(fl, fr, maxdisp // 4))
  return sd


# ----------------------------------------------------------------------------
# deterministic synthetic weights (shared by the golden generator, the tests
# and bench.py so that nothing but a seed has to travel)
# ----------------------------------------------------------------------------


def synthetic_state_dict(shapes: dict, seed: int = 0, logit_gain: float = 1.0) -> dict:
  """Key-addressed deterministic init: every tensor depends only on (seed, key, shape).

  Convs: N(0, 2/fan_in).  BN: gamma ~ U(.6,1.0), beta ~ N(0,.05), mean ~ N(0,.05),
  var ~ U(.8,1.2) -- chosen so a random-init network stays un-saturated in eval
  mode (SURVEY.md hard part 5) without a calibration pass.
  """
  import zlib
  out = {}
  for k in sorted(shapes):
    shp = tuple(shapes[k])
    g = torch.Generator().manual_seed((zlib.crc32(k.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    if k.endswith('num_batches_tracked'):
      out[k] = torch.tensor(1, dtype=torch.long)
    elif k.endswith('running_mean'):
      out[k] = torch.randn(shp, generator=g) * 0.05
    elif k.endswith('running_var'):
      out[k] = torch.rand(shp, generator=g) * 0.4 + 0.8
    elif k.endswith('.bias'):
      out[k] = torch.randn(shp, generator=g) * 0.05
    elif len(shp) == 1:  # BN gamma
      out[k] = torch.rand(shp, generator=g) * 0.4 + 0.6
    else:
      fan_in = int(np.prod(shp[1:]))
      if 'conv5.0' in k or 'conv6.0' in k:  # ConvTranspose3d (in,out,k,k,k): 27/8 taps per output
        fan_in = shp[0] * 27 // 8
      std = math.sqrt(2.0 / fan_in)
      if k.startswith('classif') and k.endswith('.2.weight'):
        std = std * logit_gain
      out[k] = torch.randn(shp, generator=g) * std
  return out


# ----------------------------------------------------------------------------
# a8. disparity -> depth   save_output_disparity_stage.py:105-133
# ----------------------------------------------------------------------------

DEEP360_BASELINES = np.array([1, 1, math.sqrt(2), math.sqrt(2), 1, 1]).astype(np.float32)  # :109
CAM_PAIRS = ('12', '13', '14', '23', '24', '34')


def disp_to_depth(disp: np.ndarray, baseline: float) -> np.ndarray:
  """Sine-rule triangulation on a Cassini disparity map (H,W) fp32 -> depth fp32."""
  output_h, output_w = disp.shape
  phi_l_start = 0.5 * math.pi - (0.5 * math.pi / output_w)
  phi_l_end = -0.5 * math.pi
  phi_l_step = math.pi / output_w
  phi_l_range = np.arange(phi_l_start, phi_l_end, -phi_l_step)
  phi_l_map = np.array([phi_l_range for j in range(output_h)]).astype(np.float32)
  mask_disp_is_0 = disp == 0
  disp_not_0 = np.ma.array(disp, mask=mask_disp_is_0)
  phi_r_map = disp_not_0 * math.pi / output_w + phi_l_map
  depth_l = np.float32(baseline) * np.sin(math.pi / 2 - phi_r_map) / np.sin(phi_r_map - phi_l_map)
  depth_l = depth_l.filled(1000)
  depth_l[depth_l > 1000] = 1000
  depth_l[depth_l < 0] = 0
  return depth_l


# ----------------------------------------------------------------------------
# a9. rotateCassini   utils/geometry.py:48-91
# ----------------------------------------------------------------------------


def _rot(pitch, yaw, roll):
  Rx = np.array([[1, 0, 0], [0, np.cos(roll), -np.sin(roll)], [0, np.sin(roll), np.cos(roll)]])
  Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
  Ry = np.array([[np.cos(pitch), 0, -np.sin(pitch)], [0, 1, 0], [np.sin(pitch), 0, np.cos(pitch)]])
  return np.dot(np.dot(Rx, Rz), Ry)


def _cassini_angles(output_h, output_w):
  theta_start = np.pi - (np.pi / output_h)
  theta_step = 2 * np.pi / output_h
  theta_range = np.arange(theta_start, -np.pi, -theta_step)
  theta_map = np.array([theta_range for i in range(output_w)]).astype(np.float32).T
  phi_start = 0.5 * np.pi - (0.5 * np.pi / output_w)
  phi_step = np.pi / output_w
  phi_range = np.arange(phi_start, -0.5 * np.pi, -phi_step)
  phi_map = np.array([phi_range for j in range(output_h)]).astype(np.float32)
  return theta_map, phi_map


def rotate_cassini_grid(output_h, output_w, pitch, yaw, roll) -> np.ndarray:
  """Constant sampling grid (H,W,2) fp32 in grid_sample [-1,1] coords (geometry.py:49-86)."""
  R_I = np.linalg.inv(_rot(pitch, yaw, roll))
  theta_2_map, phi_2_map = _cassini_angles(output_h, output_w)
  x_2 = np.sin(phi_2_map)
  y_2 = np.cos(phi_2_map) * np.sin(theta_2_map)
  z_2 = np.cos(phi_2_map) * np.cos(theta_2_map)
  X_2 = np.expand_dims(np.dstack((x_2, y_2, z_2)), axis=-1)
  X_1 = np.matmul(R_I, X_2)
  theta_1_map = np.arctan2(X_1[:, :, 1, 0], X_1[:, :, 2, 0])
  phi_1_map = np.arcsin(np.clip(X_1[:, :, 0, 0], -1, 1))
  gx = np.clip(-phi_1_map / (0.5 * np.pi), -1, 1).astype(np.float32)
  gy = np.clip(-theta_1_map / np.pi, -1, 1).astype(np.float32)
  return np.stack([gx, gy], axis=-1)


def grid_sample_bilinear_border(img: np.ndarray, grid: np.ndarray) -> np.ndarray:
  """F.grid_sample(bilinear, align_corners=True, padding_mode='border') on (H,W[,C]) maps."""
  squeeze = img.ndim == 2
  if squeeze:
    img = img[:, :, None]
  src = torch.from_numpy(np.ascontiguousarray(img)).float().permute(2, 0, 1).unsqueeze(0)
  g = torch.from_numpy(np.ascontiguousarray(grid)).float().unsqueeze(0)
  out = F.grid_sample(src, g, mode='bilinear', align_corners=True, padding_mode='border')
  out = out[0].permute(1, 2, 0).numpy().astype(img.dtype)
  return out[:, :, 0] if squeeze else out


def rotate_cassini(cassini_1: np.ndarray, pitch, yaw, roll) -> np.ndarray:
  grid = rotate_cassini_grid(cassini_1.shape[0], cassini_1.shape[1], pitch, yaw, roll)
  return grid_sample_bilinear_border(cassini_1, grid)


# ----------------------------------------------------------------------------
# a10. depthViewTransWithConf   utils/geometry.py:94-156
# ----------------------------------------------------------------------------


def depth_view_trans_targets(view_1: np.ndarray, y0, z0, x0, pitch, yaw, roll):
  """fp64 geometry -> (r_2 fp64, I_2 int16, J_2 int16)  (geometry.py:94-137)."""
  R = _rot(pitch, yaw, roll)
  t = np.array([[x0], [y0], [z0]])
  output_h, output_w = view_1.shape
  theta_1_map, phi_1_map = _cassini_angles(output_h, output_w)
  r_1 = view_1
  x_1 = r_1 * np.sin(phi_1_map)
  y_1 = r_1 * np.cos(phi_1_map) * np.sin(theta_1_map)
  z_1 = r_1 * np.cos(phi_1_map) * np.cos(theta_1_map)
  X_1 = np.expand_dims(np.dstack((x_1, y_1, z_1)), axis=-1)
  X_2 = np.matmul(R, X_1 - t)
  r_2 = np.sqrt(np.square(X_2[:, :, 0, 0]) + np.square(X_2[:, :, 1, 0]) + np.square(X_2[:, :, 2, 0]))
  theta_2_map = np.arctan2(X_2[:, :, 1, 0], X_2[:, :, 2, 0])
  with np.errstate(invalid='ignore', divide='ignore'):
    phi_2_map = np.arcsin(np.clip(X_2[:, :, 0, 0] / r_2, -1, 1))
  I_2 = np.clip(np.rint(output_h / 2 - output_h * theta_2_map / (2 * np.pi)), 0, output_h - 1).astype(np.int16)
  J_2 = np.clip(np.rint(output_w / 2 - output_w * phi_2_map / np.pi), 0, output_w - 1).astype(np.int16)
  return r_2, I_2, J_2


def depth_view_trans_with_conf(view_1, conf_1, y0, z0, x0, pitch, yaw, roll):
  """Serial strict-`<` z-buffer forward warp (geometry.py:139-156)."""
  output_h, output_w = view_1.shape
  r_2, I_2, J_2 = depth_view_trans_targets(view_1, y0, z0, x0, pitch, yaw, roll)
  view_2 = np.ones((output_h, output_w)).astype(np.float32) * 100000
  conf_2 = np.zeros((output_h, output_w)).astype(np.float32)
  r1 = view_1.reshape(-1)
  r2 = r_2.reshape(-1)
  c1 = conf_1.reshape(-1)
  tgt = (I_2.astype(np.int64) * output_w + J_2.astype(np.int64)).reshape(-1)
  v2 = view_2.reshape(-1)
  c2 = conf_2.reshape(-1)
  for n in range(r1.size):  # geometry.py:150-155 (row-major, strict <)
    if r1[n] > 0:
      j = tgt[n]
      if r2[n] < v2[j]:
        v2[j] = r2[n]  # fp64 -> fp32 store, as numba does into the float32 buffer
        c2[j] = c1[n]
  view_2[view_2 == 100000] = 0
  view_2[view_2 > 1000] = 1000
  return view_2, conf_2


# ----------------------------------------------------------------------------
# a8 (tail). per-pair alignment into camera-1's frame  save_output_disparity_stage.py:135-157
# ----------------------------------------------------------------------------


def disp2depth(disp: np.ndarray, conf_map: np.ndarray, cam_pair: str):
  idx = CAM_PAIRS.index(cam_pair)
  depth_l = disp_to_depth(disp, DEEP360_BASELINES[idx])
  if cam_pair == '12':
    return depth_l, conf_map
  if cam_pair == '13':
    return rotate_cassini(depth_l, 0.5 * math.pi, 0, 0), rotate_cassini(conf_map, 0.5 * math.pi, 0, 0)
  if cam_pair == '14':
    return rotate_cassini(depth_l, 0.25 * math.pi, 0, 0), rotate_cassini(conf_map, 0.25 * math.pi, 0, 0)
  if cam_pair == '23':
    return depth_view_trans_with_conf(depth_l, conf_map, 0, -math.sqrt(2) / 2, -math.sqrt(2) / 2, 0.75 * math.pi, 0, 0)
  if cam_pair == '24':
    return depth_view_trans_with_conf(depth_l, conf_map, 0, -1, 0, 0.5 * math.pi, 0, 0)
  if cam_pair == '34':
    return depth_view_trans_with_conf(depth_l, conf_map, 0, 1, 0, 0, 0, 0)
  raise ValueError(cam_pair)


# ----------------------------------------------------------------------------
# a11. cassini2Equirec   utils/geometry.py:7-45
# ----------------------------------------------------------------------------


def cassini2equirec_grid(ca_h: int, ca_w: int) -> np.ndarray:
  """Grid (erp_h=ca_w, erp_w=ca_h, 2) fp32 (geometry.py:16-36)."""
  erp_h, erp_w = ca_w, ca_h
  theta_erp_range = np.arange(np.pi - (np.pi / erp_w), -np.pi, -(2 * np.pi / erp_w))
  theta_erp_map = np.array([theta_erp_range for i in range(erp_h)]).astype(np.float32)
  phi_erp_range = np.arange(0.5 * np.pi - (0.5 * np.pi / erp_h), -0.5 * np.pi, -(np.pi / erp_h))
  phi_erp_map = np.array([phi_erp_range for j in range(erp_w)]).astype(np.float32).T
  theta_cassini_map = np.arctan2(np.tan(phi_erp_map), np.cos(theta_erp_map))
  phi_cassini_map = np.arcsin(np.cos(phi_erp_map) * np.sin(theta_erp_map))
  gx = np.clip(-phi_cassini_map / (0.5 * np.pi), -1, 1).astype(np.float32)
  gy = np.clip(-theta_cassini_map / np.pi, -1, 1).astype(np.float32)
  return np.stack([gx, gy], axis=-1)


def cassini2equirec(cassini: np.ndarray) -> np.ndarray:
  """(ca_h, ca_w[,C]) -> (ca_w, ca_h[,C])."""
  return grid_sample_bilinear_border(cassini, cassini2equirec_grid(cassini.shape[0], cassini.shape[1]))
