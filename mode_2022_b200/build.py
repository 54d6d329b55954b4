"""Build libmode_b200.so (the C-ABI kernel library) in-tree with nvcc for sm_100a.

    python -m mode_2022_b200.build [--force] [--verbose]

Every csrc/*.cu is compiled with `-gencode arch=compute_100a,code=sm_100a -lineinfo` and linked into
mode_2022_b200/lib/libmode_b200.so.  The library depends only on the CUDA runtime (no torch, no Python):
it is the drop-in boundary described in include/mode_b200.h and INTEGRATION.md.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, 'csrc')
OBJ = os.path.join(ROOT, 'build', 'obj')
LIBDIR = os.path.join(PKG, 'lib')
LIB = os.path.join(LIBDIR, 'libmode_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-O3', '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-Xcompiler', '-fPIC', '-Xcompiler', '-O3',
         '--expt-relaxed-constexpr', '-Xptxas', '-v']


def _digest(path: str) -> str:
  h = hashlib.sha256()
  for p in [path] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))) + [os.path.join(ROOT, 'include', 'mode_b200.h')]:
    h.update(open(p, 'rb').read())
  h.update(' '.join(FLAGS).encode())
  return h.hexdigest()


def _compile(src: str, force: bool, verbose: bool):
  name = os.path.splitext(os.path.basename(src))[0]
  obj = os.path.join(OBJ, name + '.o')
  stamp = obj + '.sha'
  dig = _digest(src)
  if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
    return obj, ''
  r = subprocess.run([NVCC, *FLAGS, '-c', src, '-o', obj], capture_output=True, text=True)
  if r.returncode != 0:
    raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
  open(stamp, 'w').write(dig)
  open(os.path.join(OBJ, name + '.ptxas.txt'), 'w').write(r.stderr)
  return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
  os.makedirs(OBJ, exist_ok=True)
  os.makedirs(LIBDIR, exist_ok=True)
  srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))
  if not os.path.exists(NVCC):
    if os.path.exists(LIB):
      return LIB
    raise RuntimeError('nvcc not found and libmode_b200.so is not built')
  with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
    res = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
  objs = [o for o, _ in res]
  if verbose:
    for _, log in res:
      if log:
        print(log)
  newest = max(os.path.getmtime(o) for o in objs)
  if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
    subprocess.check_call([NVCC, '-shared', '-o', LIB, *objs, '-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart', '-lcuda'])
  return LIB


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
