"""Stage-boundary geometry on the GPU -- API-compatible with the reference utils/geometry.py
(`cassini2Equirec`, `rotateCassini`, `depthViewTransWithConf`) plus `disp2depth`
(save_output_disparity_stage.py:105-160), backed by libmode_b200 kernels.

Every function accepts either numpy arrays (reference behaviour: result returned as numpy, one H2D + one D2H)
or CUDA tensors (result stays on the device: this is what the in-memory stage boundary uses, so that
disparity/confidence never leave HBM between ModeDisparity and ModeFusion).

All constant sampling grids / trig tables depend only on the image shape and the camera angles; they are
generated once on the host in numpy (fp64 -> fp32 exactly as the reference does) and cached on the device.
"""
from __future__ import annotations

import collections
import math
from functools import lru_cache

import numpy as np
import torch

from .. import ops

CAM_PAIRS = ('12', '13', '14', '23', '24', '34')
# Deep360 rig baselines per pair (save_output_disparity_stage.py:109)
DEEP360_BASELINES = np.array([1, 1, math.sqrt(2), math.sqrt(2), 1, 1]).astype(np.float32)
# pair -> ('rotate', pitch) or ('warp', (y0, z0, x0, pitch))   (save_output_disparity_stage.py:135-157)
PAIR_TRANSFORMS = {
    '12': ('identity', None),
    '13': ('rotate', 0.5 * math.pi),
    '14': ('rotate', 0.25 * math.pi),
    '23': ('warp', (0, -math.sqrt(2) / 2, -math.sqrt(2) / 2, 0.75 * math.pi)),
    '24': ('warp', (0, -1, 0, 0.5 * math.pi)),
    '34': ('warp', (0, 1, 0, 0)),
}


def _rotation(pitch, yaw, roll):
  """R = Rx(roll) . Rz(yaw) . Ry(pitch)  (reference geometry.py:49-55, 95-101)."""
  cr, sr, cy, sy, cp, sp = np.cos(roll), np.sin(roll), np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
  Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
  Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
  Ry = np.array([[cp, 0, -sp], [0, 1, 0], [sp, 0, cp]])
  return np.dot(np.dot(Rx, Rz), Ry)


def _cassini_axes(h, w):
  """Per-row longitude theta (h) and per-column latitude phi (w) of a Cassini image, fp32
  (reference geometry.py:60-70: arange in fp64, cast to fp32)."""
  theta = np.arange(np.pi - (np.pi / h), -np.pi, -(2 * np.pi / h)).astype(np.float32)
  phi = np.arange(0.5 * np.pi - (0.5 * np.pi / w), -0.5 * np.pi, -(np.pi / w)).astype(np.float32)
  return theta, phi


@lru_cache(maxsize=None)
def _rotate_grid_host(h, w, pitch, yaw, roll):
  theta, phi = _cassini_axes(h, w)
  theta_map = np.repeat(theta[:, None], w, axis=1)
  phi_map = np.repeat(phi[None, :], h, axis=0)
  X2 = np.expand_dims(np.dstack((np.sin(phi_map), np.cos(phi_map) * np.sin(theta_map), np.cos(phi_map) * np.cos(theta_map))), axis=-1)
  X1 = np.matmul(np.linalg.inv(_rotation(pitch, yaw, roll)), X2)
  theta1 = np.arctan2(X1[:, :, 1, 0], X1[:, :, 2, 0])
  phi1 = np.arcsin(np.clip(X1[:, :, 0, 0], -1, 1))
  gx = np.clip(-phi1 / (0.5 * np.pi), -1, 1).astype(np.float32)
  gy = np.clip(-theta1 / np.pi, -1, 1).astype(np.float32)
  return np.ascontiguousarray(np.stack([gx, gy], axis=-1))


@lru_cache(maxsize=None)
def _c2e_grid_host(ca_h, ca_w):
  erp_h, erp_w = ca_w, ca_h
  th = np.arange(np.pi - (np.pi / erp_w), -np.pi, -(2 * np.pi / erp_w))
  ph = np.arange(0.5 * np.pi - (0.5 * np.pi / erp_h), -0.5 * np.pi, -(np.pi / erp_h))
  theta_map = np.array([th for _ in range(erp_h)]).astype(np.float32)
  phi_map = np.array([ph for _ in range(erp_w)]).astype(np.float32).T
  theta_c = np.arctan2(np.tan(phi_map), np.cos(theta_map))
  phi_c = np.arcsin(np.cos(phi_map) * np.sin(theta_map))
  gx = np.clip(-phi_c / (0.5 * np.pi), -1, 1).astype(np.float32)
  gy = np.clip(-theta_c / np.pi, -1, 1).astype(np.float32)
  return np.ascontiguousarray(np.stack([gx, gy], axis=-1))


_DEV = collections.OrderedDict()  # constant grids / trig tables on the device, LRU-bounded
_DEV_MAX = 64


def _dev(key, make, device):
  k = (key, str(device))
  t = _DEV.get(k)
  if t is None:
    t = _DEV[k] = torch.from_numpy(np.ascontiguousarray(make())).to(device)
    if not torch.cuda.is_current_stream_capturing():
      torch.cuda.current_stream(device).synchronize()  # built once; complete for every stream that uses it later
    while len(_DEV) > _DEV_MAX:
      _DEV.popitem(last=False)
  else:
    _DEV.move_to_end(k)
  return t


def _as_device(x):
  """-> (tensor on cuda, was_numpy)."""
  if isinstance(x, np.ndarray):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda(), True
  if not x.is_cuda:
    raise NotImplementedError('geometry: CUDA tensors or numpy arrays only')
  return x, False


def _resample(img, grid_key, grid_fn):
  """img: numpy (H,W) / (H,W,C) or tensor (N,C,H,W) -> resampled with a cached constant grid."""
  t, was_np = _as_device(img)
  if was_np:
    dtype = img.dtype
    t = t.float()
    t = t[None, None] if t.dim() == 2 else t.permute(2, 0, 1)[None]
  grid = _dev(grid_key, grid_fn, t.device)
  out = ops.grid_sample_border(t.float(), grid)
  if was_np:
    out = out[0].permute(1, 2, 0).cpu().numpy().astype(dtype)
    return out[:, :, 0] if img.ndim == 2 else out
  return out


def cassini2Equirec(cassini):
  """Cassini (H,W[,C]) numpy or (N,C,H,W) tensor -> equirectangular (W,H[,C]) / (N,H... squeezed as the
  reference does for tensors: (N,W,H) when C == 1)  (reference geometry.py:7-45)."""
  if isinstance(cassini, np.ndarray):
    h, w = cassini.shape[:2]
  else:
    h, w = cassini.shape[-2:]
  out = _resample(cassini, ('c2e', h, w), lambda: _c2e_grid_host(h, w))
  if isinstance(cassini, np.ndarray):
    return out.squeeze()
  return out.squeeze(1)


def rotateCassini(cassini_1, pitch, yaw, roll):
  """Rotate a Cassini map by (pitch, yaw, roll)  (reference geometry.py:48-91)."""
  if isinstance(cassini_1, np.ndarray):
    h, w = cassini_1.shape[:2]
  else:
    h, w = cassini_1.shape[-2:]
  return _resample(cassini_1, ('rot', h, w, float(pitch), float(yaw), float(roll)), lambda: _rotate_grid_host(h, w, float(pitch), float(yaw), float(roll)))


def _warp_tables(h, w, device):
  theta, phi = _cassini_axes(h, w)
  return (_dev(('sp', h, w), lambda: np.sin(phi), device), _dev(('cp', h, w), lambda: np.cos(phi), device), _dev(('st', h, w), lambda: np.sin(theta), device),
          _dev(('ct', h, w), lambda: np.cos(theta), device))


def depthViewTransWithConf(view_1, conf_1, y0, z0, x0, pitch, yaw, roll):
  """Forward-warp a Cassini depth map (+confidence) into another camera's frame with a z-buffer
  (reference geometry.py:94-156; note the reference argument order y0, z0, x0)."""
  d, was_np = _as_device(view_1)
  c, _ = _as_device(conf_1)
  d, c = (d if d.dtype == torch.float64 else d.float()), c.float()  # fp64 depth stays fp64 (numpy promotion)
  squeeze = d.dim() == 2
  if squeeze:
    d, c = d[None], c[None]
  lead = d.shape[:-2]
  h, w = d.shape[-2:]
  R = _rotation(pitch, yaw, roll)
  Rt = [*R.reshape(-1).tolist(), float(x0), float(y0), float(z0)]
  v2, c2 = ops.depth_view_trans(d.reshape(-1, h, w), c.reshape(-1, h, w), *_warp_tables(h, w, d.device), Rt)
  v2, c2 = v2.reshape(*lead, h, w), c2.reshape(*lead, h, w)
  if squeeze:
    v2, c2 = v2[0], c2[0]
  if was_np:
    return v2.cpu().numpy(), c2.cpu().numpy()
  return v2, c2


def disp_to_depth(disp, baseline, want_f64=False):
  """Sine-rule triangulation of a Cassini disparity map (save_output_disparity_stage.py:118-133).
  Returns fp32 depth, or (fp32, fp64) with want_f64 -- the reference computes this in fp64 under NumPy >= 2
  and feeds the un-rounded values to the forward warp."""
  t, was_np = _as_device(disp)
  h, w = t.shape[-2:]
  phi_l = _dev(('phil', w), lambda: np.arange(0.5 * math.pi - (0.5 * math.pi / w), -0.5 * math.pi, -(math.pi / w)).astype(np.float32), t.device)
  out = ops.disp_to_depth(t.float(), phi_l, float(baseline), want_f64)
  if was_np:
    return tuple(o.cpu().numpy() for o in out) if want_f64 else out.cpu().numpy()
  return out


def disp2depth(disp, conf_map, cam_pair, baselines=DEEP360_BASELINES):
  """Disparity+confidence of one camera pair -> depth+confidence in camera 1's frame
  (reference save_output_disparity_stage.py:105-160).  (H,W) numpy arrays or CUDA tensors."""
  if cam_pair not in PAIR_TRANSFORMS:
    raise ValueError('Error! Wrong Cam_pair!')
  kind, arg = PAIR_TRANSFORMS[cam_pair]
  depth = disp_to_depth(disp, baselines[CAM_PAIRS.index(cam_pair)], want_f64=(kind == 'warp'))
  if kind == 'warp':
    depth = depth[1]
  if kind == 'identity':
    return depth, conf_map
  if kind == 'rotate':
    if isinstance(depth, np.ndarray):
      return rotateCassini(depth, arg, 0, 0), rotateCassini(conf_map, arg, 0, 0)
    both = rotateCassini(torch.stack([depth.float(), conf_map.float()])[None], arg, 0, 0)[0]
    return both[0], both[1]
  y0, z0, x0, pitch = arg
  return depthViewTransWithConf(depth, conf_map, y0, z0, x0, pitch, 0, 0)


class StageBoundary:
  """In-memory replacement of the npz/PNG hand-off between the two stages:
  takes the (6,1,H,W) disparity/confidence maps of one frame (pair order CAM_PAIRS) on the device and returns
  the lists ModeFusion.forward expects, without leaving HBM.  `quantise_conf=True` emulates the reference's
  uint8 PNG round trip of the confidence (save_output_disparity_stage.py:199, deep360_loader.py:22-29)."""

  def __init__(self, baselines=DEEP360_BASELINES, quantise_conf=False):
    self.baselines = baselines
    self.quantise_conf = quantise_conf

  def __call__(self, disp6: torch.Tensor, conf6: torch.Tensor):
    depths, confs = [], []
    for i, pair in enumerate(CAM_PAIRS):
      d, c = disp2depth(disp6[i, 0], conf6[i, 0], pair, self.baselines)
      if self.quantise_conf:
        c = torch.round(c * 255).clamp(0, 255) / 255.0  # cv2.imwrite saturate_cast<uchar> (round-half-even), read back /255
      depths.append(d[None, None])
      confs.append(c[None, None])
    return depths, confs
