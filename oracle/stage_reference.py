"""Stage the UNMODIFIED reference under baseline/_ref/ref so that it can run on the GPU box (checker + GPU baseline).

TEST / MEASUREMENT INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so the reference's Python packages
(`models/`, `utils/`) are copied, byte for byte, to baseline/_ref/ref/ (git-ignored: it never enters history, but it travels
with the gpurun snapshot) and the reference's own CUDA extension -- compiled from the sources where they lie by
oracle/build_ref.py -- is placed where the reference imports it from (`from . import sphere_conv_cuda`,
models/basic/spherical_conv/sphere_conv.py:12).  Nothing is patched: `reference_package()` imports exactly those files.

  python oracle/stage_reference.py        # in the build container (has /root/reference)

Used by tests/test_gpu_reference.py (whole-model parity against the primary oracle of SURVEY.md section 8c, and the
operator-level drop-in proof of INTEGRATION.md section 2), tools/ref_gpu_bench.py and bench.py's `reference_gpu` block.
"""
from __future__ import annotations

import importlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get('MODE_REFERENCE', '/root/reference')
DST = os.path.join(ROOT, 'baseline', '_ref', 'ref')
OP_DIR = os.path.join(DST, 'models', 'basic', 'spherical_conv')


def stage(verbose: bool = True) -> str | None:
  """Copy models/ + utils/ and the compiled extension; returns the staged root, or None when nothing can be staged."""
  if not os.path.isdir(os.path.join(REF, 'models')):
    return DST if os.path.isdir(os.path.join(DST, 'models')) else None  # GPU box: use what travelled
  for pkg in ('models', 'utils'):
    dst = os.path.join(DST, pkg)
    if os.path.isdir(dst):
      shutil.rmtree(dst)
    shutil.copytree(os.path.join(REF, pkg), dst, ignore=shutil.ignore_patterns('__pycache__', '*.pyc', 'build', '*.so'))
  from oracle import build_ref
  so = build_ref.build(verbose=verbose)
  if so is None:
    raise RuntimeError('oracle/_ref/sphere_conv_cuda.so could not be built')
  for d, _, files in os.walk(DST):  # the reference tree is mounted read-only; the staged copy must stay replaceable
    os.chmod(d, 0o755)
    for f in files:
      os.chmod(os.path.join(d, f), 0o644)
  shutil.copy2(so, os.path.join(OP_DIR, 'sphere_conv_cuda.so'))
  if verbose:
    print('staged', DST)
  return DST


def available() -> bool:
  return os.path.exists(os.path.join(OP_DIR, 'sphere_conv_cuda.so')) and os.path.exists(os.path.join(DST, 'models', 'mode_disparity.py'))


def reference_package(native_op=None):
  """Import the staged reference `models` package (fresh copy of the module objects on every call).

  native_op=None: the reference's own compiled extension.  Otherwise `native_op` is a module object that is installed as
  `models.basic.spherical_conv.sphere_conv_cuda` BEFORE the import, i.e. the reference's unmodified Python runs on top of a
  replacement of its pybind module."""
  if not available():
    return None
  for name in [n for n in sys.modules if n == 'models' or n.startswith('models.')]:
    del sys.modules[name]
  if native_op is not None:
    sys.modules['models.basic.spherical_conv.sphere_conv_cuda'] = native_op
  sys.path.insert(0, DST)
  try:
    pkg = importlib.import_module('models')
  finally:
    sys.path.remove(DST)
    for name in [n for n in sys.modules if n == 'models' or n.startswith('models.')]:
      del sys.modules[name]  # keep the name free for the next variant; the returned module objects stay alive
  return pkg


if __name__ == '__main__':
  sys.path.insert(0, ROOT)
  print(stage())
