import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
  config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def pytest_collection_modifyitems(config, items):
  import torch
  if torch.cuda.is_available():
    return
  skip = pytest.mark.skip(reason='no CUDA device')
  for it in items:
    if 'gpu' in it.keywords:
      it.add_marker(skip)


@pytest.fixture(scope='session')
def gold_dir():
  return GOLD


@pytest.fixture(autouse=True)
def _deterministic_rng():
  """Every test starts from the same global RNG state (CPU and CUDA): a few tests draw unseeded tensors, and statistical
  bounds (max over ~1e5 outputs of a random-walk rounding error) must not depend on what ran before."""
  import torch
  torch.manual_seed(20221017)
  yield
