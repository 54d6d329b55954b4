"""Shared test helpers: deterministic synthetic inputs / weights (identical to the golden generator)."""
import json
import os

import numpy as np
import torch

from oracle import mode_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
MANIFEST = json.load(open(os.path.join(GOLD, 'MANIFEST.json')))
KEY_SHAPES = json.load(open(os.path.join(GOLD, 'mode_disparity_keys.json')))


def synth_inputs(H, W, seed, B=1):
  g = torch.Generator().manual_seed(1000 + seed)
  return torch.randn(B, 3, H, W, generator=g), torch.randn(B, 3, H, W, generator=g)


def synth_geometry_inputs(H, W, seed, maxdisp=192):
  g = np.random.default_rng(seed)
  disp = (g.random((H, W), dtype=np.float32) * (maxdisp - 1)).astype(np.float32)
  disp[g.random((H, W)) < 0.01] = 0.0
  conf = g.random((H, W), dtype=np.float32)
  return disp, conf


def golden_state_dict(name):
  """Synthetic weights of golden config `name` with the calibrated BN statistics stored in the fixture."""
  H, W, D, st, seed = MANIFEST['disp_configs'][name]
  z = np.load(os.path.join(GOLD, f'mode_disparity_{name}.npz'))
  sd = O.synthetic_state_dict(KEY_SHAPES, seed=seed)
  for k in z.files:
    if k.startswith('bn/'):
      sd[k[3:]] = torch.from_numpy(z[k])
  return sd, (H, W, D, st, seed), z


def calibrated_state_dict(H, W, D, st, seed, B=1):
  left, right = synth_inputs(H, W, seed, B)
  sd = O.calibrate_bn(O.synthetic_state_dict(KEY_SHAPES, seed=seed), left, right, D, st)
  return sd, left, right
