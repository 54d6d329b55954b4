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


# ---- training-step fixture (oracle/pin_training_against_reference.py writes it, tests/test_gpu_model.py reads it)
TRAIN_GRAD_KEYS = ['feature_extraction.firstconv.0.0.weight', 'feature_extraction.layer2.0.conv1.0.0.weight', 'feature_extraction.layer4.0.conv1.0.0.weight',
                   'feature_extraction.layer4.0.conv1.0.1.weight', 'feature_extraction.layer4.2.conv2.0.weight', 'feature_extraction.lastconv.4.0.weight',
                   'dres0.0.0.weight', 'dres2.conv1.0.0.weight', 'dres3.conv5.0.weight', 'dres4.conv6.0.weight', 'dres4.conv6.1.bias', 'classif1.2.weight',
                   'classif3.2.weight']


def train_inputs(H, W, D, seed, B=2):
  g = torch.Generator().manual_seed(7000 + seed)
  left, right = torch.randn(B, 3, H, W, generator=g), torch.randn(B, 3, H, W, generator=g)
  disp_true = torch.rand(B, 1, H, W, generator=g) * (D - 1)
  mask = torch.rand(B, 1, H, W, generator=g) < 0.8
  return left, right, disp_true, mask


def train_loss(o1, o2, o3, disp_true, mask):
  """The reference's training loss (train_disparity.py:147-158)."""
  import torch.nn.functional as F
  return 0.5 * F.smooth_l1_loss(o1[mask], disp_true[mask]) + 0.7 * F.smooth_l1_loss(o2[mask], disp_true[mask]) + F.smooth_l1_loss(o3[mask], disp_true[mask])


def grad_sample(g: torch.Tensor, n=8192):
  """Every k-th element of a flattened gradient (<= n values): keeps the fixture small; the relative L2 distance over the
  sample is the statistic compared."""
  f = g.detach().reshape(-1)
  return f[::max(1, (f.numel() + n - 1) // n)]


def fusion_inputs(H, W, seed, B=1):
  """6 depth maps in [0, 20], 6 confidences in [0, 1], 4 RGB images (the shapes ModeFusion.forward takes)."""
  g = torch.Generator().manual_seed(9000 + seed)
  depthes = [torch.rand(B, 1, H, W, generator=g) * 20.0 for _ in range(6)]
  confs = [torch.rand(B, 1, H, W, generator=g) for _ in range(6)]
  rgbs = [torch.randn(B, 3, H, W, generator=g) for _ in range(4)]
  return depthes, confs, rgbs


def numpy_matmul_is_fma_102(n=40):
  """Does THIS host's np.matmul evaluate a stacked (3,3) @ (3,1) product as fma(r2, c, fma(r0, a, r1*b)) per row?  (OpenBLAS' FMA
  micro-kernel on the AVX-512 hosts of this pool does; the forward-warp kernel follows that order, csrc/geometry.cu.)  Checked with
  exact rational arithmetic on a sample."""
  from fractions import Fraction
  rng = np.random.default_rng(12)
  R = rng.standard_normal((3, 3))
  X = rng.standard_normal((n, 3, 1)) * 7
  Y = np.matmul(R, X)[..., 0]
  fma = lambda a, b, c: float(Fraction(a) * Fraction(b) + Fraction(c))
  for i in range(n):
    v = X[i, :, 0]
    for r in range(3):
      if fma(R[r, 2], v[2], fma(R[r, 0], v[0], R[r, 1] * v[1])) != Y[i, r]:
        return False
  return True
