"""GPU parity of the whole stereo stage: ModeDisparity (B200 kernels) against the golden vectors recorded from
the reference, and against the oracle at a larger size."""
import os

import numpy as np
import pytest
import torch

from oracle import mode_oracle as O
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


def _model(name_or_cfg, precision, sd, H, W, D, st):
  from mode_2022_b200.models import ModeDisparity
  m = ModeDisparity(D, conv='Sphere', in_height=H, in_width=W, sphereType=st, out_conf=True, precision=precision)
  m.load_state_dict(sd)
  return m.cuda().eval()


@pytest.mark.parametrize('name', ['tiny_cassini', 'tiny_erp', 'small_cassini'])
def test_fp32_matches_reference_golden(name):
  torch.backends.cudnn.allow_tf32 = False
  torch.backends.cuda.matmul.allow_tf32 = False
  sd, (H, W, D, st, seed), z = Hh.golden_state_dict(name)
  left, right = Hh.synth_inputs(H, W, seed)
  m = _model(name, 'fp32', sd, H, W, D, st)
  pred, conf = m(left.cuda(), right.cuda())
  pred, conf = pred.cpu().numpy(), conf.cpu().numpy()
  assert pred.shape == z['pred'].shape == (1, 1, H, W)
  rel = np.abs(pred - z['pred']) / np.maximum(np.abs(z['pred']), 1.0)
  assert rel.max() <= 1e-4, rel.max()  # north_star: fp32 disparity within 1e-4 relative
  same_r = np.rint(pred) == np.rint(z['pred'])
  assert same_r.mean() > 0.995
  assert (np.abs(conf - z['conf']) * same_r).max() <= 1e-4
  _, _, stages = m._plan.run(left.cuda(), right.cuda(), return_stages=True)
  assert np.abs(stages['feat'][:1].cpu().numpy() - z['feat_l']).max() <= 1e-4
  assert np.abs(stages['cost3'].cpu().numpy() - z['cost3']).max() <= 2e-4


def test_fp32_batch_and_out_conf_contract():
  sd, (H, W, D, st, seed), z = Hh.golden_state_dict('tiny_cassini')
  from mode_2022_b200.models import ModeDisparity
  m = ModeDisparity(D, in_height=H, in_width=W, sphereType=st, out_conf=False, precision='fp32')
  m.load_state_dict(sd)
  m = m.cuda().eval()
  left, right = Hh.synth_inputs(H, W, seed)
  l3 = torch.cat([left, right, left]).cuda()
  r3 = torch.cat([right, left, right]).cuda()
  out = m(l3, r3)
  assert isinstance(out, torch.Tensor) and out.shape == (3, 1, H, W)  # out_conf=False -> pred3 only
  assert np.abs(out[0].cpu().numpy() - z['pred'][0]).max() <= 1e-3 and torch.equal(out[0], out[2])
  with pytest.raises(ValueError):
    m(l3[:, :, :-8], r3[:, :, :-8])
