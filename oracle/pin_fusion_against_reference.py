"""TEST INFRASTRUCTURE ONLY -- golden vectors for ModeFusion / Baseline, generated from the unmodified reference.

Runs the reference's own fusion networks (models/mode_fusion.py:91-247) on CPU with key-addressed synthetic weights
(oracle.synthetic_state_dict over the reference's state-dict shapes) and seeded inputs, and stores the outputs in
tests/golden/mode_fusion_64x32.npz.  Cannot travel to the GPU box; run here:  python oracle/pin_fusion_against_reference.py
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mode_oracle as O  # noqa: E402
from oracle import pin_against_reference as P  # noqa: E402
from tests import helpers as Hh  # noqa: E402


def main():
  models, SCM, RG = P.import_reference()
  torch.set_num_threads(os.cpu_count())
  H, W, seed = 64, 32, 4
  depthes, confs, rgbs = Hh.fusion_inputs(H, W, seed)
  out = {}
  for name, ctor in (('fusion', lambda: models.ModeFusion(20.0, [32, 64, 128, 256], {'depth': 12, 'rgb': 12})), ('baseline', lambda: models.Baseline(20.0))):
    m = ctor()
    shapes = {k: list(v.shape) for k, v in m.state_dict().items()}
    keyfile = os.path.join(P.GOLD, 'mode_fusion_keys.json' if name == 'fusion' else 'baseline_keys.json')
    if os.path.exists(keyfile):
      assert shapes == json.load(open(keyfile)), name
    else:
      json.dump(shapes, open(keyfile, 'w'), indent=0)
    m.load_state_dict(O.synthetic_state_dict(shapes, seed=seed))
    m.eval()
    with torch.no_grad():
      y = m(depthes, confs, rgbs) if name == 'fusion' else m(depthes)
    out[name] = y.numpy()
    print(name, tuple(y.shape), 'range [%.4f, %.4f] std %.4f' % (y.min().item(), y.max().item(), y.std().item()))
  np.savez_compressed(os.path.join(P.GOLD, 'mode_fusion_64x32.npz'), **out)


if __name__ == '__main__':
  main()
