"""CPU: host-side logic of the product -- checkpoint-key parity, bit-exact constant grids, library exports."""
import ctypes
import hashlib
import json
import os
import re

import numpy as np
import pytest
import torch

from oracle import mode_oracle as O
from tests import helpers as Hh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
  from mode_2022_b200 import _lib
  hdr = open(os.path.join(ROOT, 'include', 'mode_b200.h')).read()
  declared = set(re.findall(r'\b(mode_[a-z0-9_]+)\s*\(', hdr))
  assert len(declared) >= 15
  lib = ctypes.CDLL(_lib.LIB_PATH)
  missing = [s for s in sorted(declared) if not hasattr(lib, s)]
  assert not missing, missing
  assert set(_lib.SIGNATURES) | set(_lib.OTHER_SYMBOLS) == declared
  assert _lib.load().mode_b200_version() >= 100


def test_state_dict_keys_match_reference(gold_dir):
  from mode_2022_b200.models import ModeDisparity, ModeFusion
  m = ModeDisparity(192, in_height=1024, in_width=512, sphereType='Cassini', out_conf=True)
  assert {k: list(v.shape) for k, v in m.state_dict().items()} == Hh.KEY_SHAPES
  assert sum(p.numel() for p in m.parameters()) == 5489280
  gold = json.load(open(os.path.join(gold_dir, 'mode_fusion_keys.json')))
  f = ModeFusion(20.0, [32, 64, 128, 256], {'depth': 12, 'rgb': 12})
  assert {k: list(v.shape) for k, v in f.state_dict().items()} == gold
  # DataParallel-style checkpoints ('module.' prefix) load too
  m.load_state_dict({'module.' + k: v for k, v in m.state_dict().items()})


def test_sphere_position_bit_exact(gold_dir):
  from mode_2022_b200.models.sphere_conv import sphere_position_numpy
  gold = json.load(open(os.path.join(gold_dir, 'sphere_position_sha256.json')))
  for k, v in gold.items():
    st, hw = k.split('_')
    h, w = map(int, hw.split('x'))
    assert hashlib.sha256(sphere_position_numpy(min(h, w), max(h, w), st).tobytes()).hexdigest() == v, k


def test_sphere_conv_module_contract():
  from mode_2022_b200.models import SphereConv
  with pytest.raises(AssertionError):
    SphereConv(10, 10, 'Cassini', 1, 1, 3)  # long side must be 2x short side (reference sphere_conv.py:131-133)
  sc = SphereConv(16, 8, 'Cassini', 4, 6, 3, 1, 1, 1)
  assert list(sc.state_dict()) == ['weight'] and sc.weight.shape == (6, 4, 3, 3)
  with pytest.raises(NotImplementedError):
    sc(torch.zeros(1, 4, 16, 8))  # CPU tensors are rejected like the reference (sphere_conv.py:33-34)
  with pytest.raises(ValueError):
    sc(torch.zeros(4, 16, 8))


def test_geometry_host_grids_bit_exact():
  from mode_2022_b200.utils import geometry as G
  for (h, w) in [(64, 32), (128, 64)]:
    for pitch in (0.5 * np.pi, 0.25 * np.pi):
      assert np.array_equal(G._rotate_grid_host(h, w, float(pitch), 0.0, 0.0), O.rotate_cassini_grid(h, w, pitch, 0, 0))
    assert np.array_equal(G._c2e_grid_host(h, w), O.cassini2equirec_grid(h, w))
    th, ph = G._cassini_axes(h, w)
    tm, pm = O._cassini_angles(h, w)
    assert np.array_equal(th, tm[:, 0]) and np.array_equal(ph, pm[0])


def test_cpu_tensors_fail_loudly_in_both_modes():
  """No CPU / PyTorch fallback: the module refuses CPU tensors in training and in eval mode alike."""
  from mode_2022_b200.models import ModeDisparity
  m = ModeDisparity(16, in_height=64, in_width=32)
  for mode in (True, False):
    m.train(mode)
    with pytest.raises(NotImplementedError):
      m(torch.zeros(1, 3, 64, 32), torch.zeros(1, 3, 64, 32))


def test_bench_roofline_report_classes():
  """bench.py's per-kernel roofline bookkeeping (pure host logic): C-ABI argument lists -> algorithmic FLOPs / bytes -> binding bound."""
  import ctypes as C
  import bench

  class Ev:
    def __init__(self, t):
      self.t = t

    def elapsed_time(self, other):
      return other.t - self.t

  vp = lambda: C.c_void_p(1234)
  prof = []
  for _ in range(2):
    prof.append(('mode_conv3d_tc', (vp(), vp(), vp(), vp(), None, None, vp(), None, 6, 32, 32, 48, 256, 128, 0, 1, 1, vp()), Ev(0.0), Ev(0.45)))
    prof.append(('mode_conv3d_tc', (vp(), vp(), vp(), vp(), vp(), None, vp(), None, 6, 64, 32, 24, 128, 64, 2, 1, 1, vp()), Ev(0.0), Ev(0.27)))
    prof.append(('mode_conv3d_classifier_tc', (vp(), vp(), None, vp(), 6, 48, 256, 128, 1, vp()), Ev(0.0), Ev(0.2)))
    prof.append(('mode_sphere_conv_tc', (vp(), vp(), vp(), vp(), vp(), vp(), vp(), 12, 128, 256, 128, 128, 1, 1, vp()), Ev(0.0), Ev(0.2)))
    prof.append(('mode_disp_regress', (vp(), vp(), vp(), 6, 48, 256, 128, 192, 1024, 512, vp()), Ev(0.0), Ev(0.39)))
    prof.append(('mode_costvol_conv_fused', (vp(), vp(), vp(), vp(), vp(), 6, 48, 256, 128, 1, 1, vp()), Ev(0.0), Ev(0.36)))
  r = bench.roofline_report(prof, 2)
  assert r['bound'] == 'tensor' and r['launches_per_step'] == 3 and 0 < r['frac'] < 1.5
  by = {k['kernel']: k for k in r['kernels']}
  s1 = by['conv3d_tc 32->32 s1 @48x256x128']
  assert s1['bound'] == 'tensor' and abs(s1['TFLOPs'] - 2 * 27 * 32 * 32 * 6 * 48 * 256 * 128 / 0.45e-3 / 1e12) < 1.0
  assert by['conv3d_tc 64->32 deconv @24x128x64']['bound'] == 'hbm'           # 1.3 GB moved: HBM-bound, not tensor-bound
  assert by['disp_regress (upsample + softmax + soft-argmin + confidence)']['bound'] == 'mufu_exp'
  assert by['sphere_conv_tc 128->128 @256x128']['launches_per_step'] == 1


def test_batchnorm_modules_are_drop_in_on_the_host():
  """models/batchnorm.py: same parameters, buffers and state-dict keys as torch's modules, `isinstance` still holds (the reference's
  init code and the BN-folding plans test for nn.BatchNorm2d / 3d), and everything the CUDA kernels do not cover -- here: CPU tensors,
  eval mode, momentum=None -- goes through the parent class unchanged.  The kernels themselves are pinned in tests/test_gpu_kernels.py."""
  import torch.nn as nn
  from mode_2022_b200.models.batchnorm import BatchNorm2d, BatchNorm3d
  from mode_2022_b200.models import ModeDisparity
  for cls_m, cls_t, shape in ((BatchNorm2d, nn.BatchNorm2d, (2, 6, 5, 7)), (BatchNorm3d, nn.BatchNorm3d, (2, 4, 3, 5, 6))):
    a, b = cls_m(shape[1]), cls_t(shape[1])
    assert isinstance(a, cls_t) and list(a.state_dict()) == list(b.state_dict())
    b.load_state_dict(a.state_dict())
    x = torch.randn(*shape)
    for mode in ('train', 'train_cumulative', 'eval'):
      a.train(mode != 'eval'), b.train(mode != 'eval')
      a.momentum = b.momentum = None if mode == 'train_cumulative' else 0.1
      assert torch.equal(a(x), b(x))
      assert torch.equal(a.running_mean, b.running_mean) and torch.equal(a.running_var, b.running_var) and int(a.num_batches_tracked) == int(b.num_batches_tracked)
  m = ModeDisparity(32, in_height=64, in_width=32, sphereType='Cassini')
  bns = [mod for mod in m.modules() if isinstance(mod, (nn.BatchNorm2d, nn.BatchNorm3d))]
  assert bns and all(isinstance(mod, (BatchNorm2d, BatchNorm3d)) for mod in bns)
