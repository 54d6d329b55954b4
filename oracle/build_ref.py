"""Compile the UNMODIFIED reference spherical-conv CUDA op into oracle/_ref/.

TEST INFRASTRUCTURE.  The two reference source files are compiled where they
lie under /root/reference (nothing is copied into this repo); the only output
is oracle/_ref/sphere_conv_cuda.so (git-ignored, travels to the GPU box with
the gpurun snapshot).  The reference's own build system (setup.py /
torch.utils.cpp_extension, models/basic/spherical_conv/setup.py:4-12) is not
run: this is a plain nvcc + g++ recipe against the torch headers.

The resulting module exposes exactly the reference pybind API
(sphere_conv_cuda.cpp:339-345): sphere_conv_forward_cuda / sphere_conv_backward_cuda.
It is used by tests/test_gpu_sphere_conv.py as the ground truth for oracle
function `sphere_conv` and for the product kernel.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('MODE_REFERENCE', '/root/reference')
SRC = os.path.join(REF, 'models', 'basic', 'spherical_conv', 'src')
OUT = os.path.join(HERE, '_ref')


def build(verbose: bool = True) -> str | None:
  so = os.path.join(OUT, 'sphere_conv_cuda.so')
  cpp = os.path.join(SRC, 'sphere_conv_cuda.cpp')
  cu = os.path.join(SRC, 'sphere_conv_cuda_kernel.cu')
  if not (os.path.exists(cpp) and os.path.exists(cu)):
    return so if os.path.exists(so) else None  # GPU box: use the prebuilt file
  if os.path.exists(so) and os.path.getmtime(so) > max(os.path.getmtime(cpp), os.path.getmtime(cu), os.path.getmtime(__file__)):
    return so
  import torch
  from torch.utils import cpp_extension as CE
  os.makedirs(OUT, exist_ok=True)
  inc = [f'-I{p}' for p in CE.include_paths('cuda')] + [f'-I{sysconfig.get_paths()["include"]}']
  defs = ['-DTORCH_EXTENSION_NAME=sphere_conv_cuda', '-DTORCH_API_INCLUDE_EXTENSION_H', f'-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}']
  o1, o2 = os.path.join(OUT, 'sphere_conv_cuda.o'), os.path.join(OUT, 'sphere_conv_cuda_kernel.o')
  cmds = [
      ['g++', '-O2', '-fPIC', '-std=c++17', '-w', *defs, *inc, '-c', cpp, '-o', o1],
      ['nvcc', '-O2', '-std=c++17', '-w', '-Xcompiler', '-fPIC', '-gencode', 'arch=compute_100a,code=sm_100a', '--expt-relaxed-constexpr', *defs, *inc, '-c', cu, '-o', o2],
      ['g++', '-shared', o1, o2, '-o', so, *[f'-L{p}' for p in CE.library_paths('cuda')], '-lc10', '-lc10_cuda', '-ltorch_cpu', '-ltorch_cuda', '-ltorch', '-ltorch_python', '-lcudart'],
  ]
  for c in cmds:
    if verbose:
      print(' '.join(c[:6]), '...', flush=True)
    subprocess.check_call(c)
  for o in (o1, o2):
    os.remove(o)
  return so


def load():
  """Import oracle/_ref/sphere_conv_cuda.so as a Python module (needs `import torch` first)."""
  import importlib.util
  import torch  # noqa: F401
  so = os.path.join(OUT, 'sphere_conv_cuda.so')
  if not os.path.exists(so):
    return None
  spec = importlib.util.spec_from_file_location('sphere_conv_cuda', so)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


if __name__ == '__main__':
  print(build())
  print(load())
