"""Drop-in replacement of the reference's pybind module `sphere_conv_cuda` on top of libmode_b200.so.

The reference's hot path crosses one native boundary: `models/basic/spherical_conv/src/sphere_conv_cuda.cpp:339-345` exports
`sphere_conv_forward_cuda` and `sphere_conv_backward_cuda`, called from `SphereConvFunction` (`sphere_conv.py:38-54,68-87`).
This module has the same two functions with the same positional signatures and the same ownership rule (the caller allocates
`output` / zero-filled gradients, the op fills / accumulates), so the reference's Python runs UNMODIFIED when this file is
installed as `models/basic/spherical_conv/sphere_conv_cuda.py` (or injected into `sys.modules` under that name, which is what
tests/test_gpu_reference.py does).  Only `ctypes` + the C ABI of include/mode_b200.h are used -- no other part of this package.

Differences from the pybind module, all inherited from the C ABI:
  * `ones` / `columns` scratch tensors are ignored (there is no column buffer);
  * shape errors raise RuntimeError from the library's message instead of TORCH_CHECK (`cpp:40-126`);
  * launch errors are reported (the reference printf()s and swallows them, `sphere_conv_cuda_kernel.cu:286-289`).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_LIB_PATH = os.environ.get('MODE_B200_LIB', os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'lib', 'libmode_b200.so'))
_lib = C.CDLL(_LIB_PATH)
_lib.mode_sphere_conv_f32.restype = C.c_int
_lib.mode_sphere_conv_f32.argtypes = [C.c_void_p] * 7 + [C.c_int] * 8 + [C.c_void_p]
_lib.mode_sphere_conv_backward_f32.restype = C.c_int
_lib.mode_sphere_conv_backward_f32.argtypes = [C.c_void_p] * 7 + [C.c_int] * 7 + [C.c_void_p]
_lib.mode_b200_last_error.restype = C.c_char_p


def _check(rc):
  if rc != 0:
    raise RuntimeError(_lib.mode_b200_last_error().decode())  # reference: TORCH_CHECK -> c10::Error -> RuntimeError


def _ptr(t):
  return C.c_void_p(t.data_ptr())


def _stream(t):
  return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)  # reference: at::cuda::getCurrentCUDAStream(), kernel.cu:280


def _guard(*tensors):
  for t in tensors:
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
      raise RuntimeError('sphere_conv_cuda (libmode_b200): tensors must be contiguous fp32 CUDA tensors')  # cpp:43-60


def sphere_conv_forward_cuda(input, weight, bias, ones, position, output, columns, kH, kW, sH, sW, pH, pW, dH, dW, group, has_bias):
  """sphere_conv_cuda.cpp:129-133.  `output` (B, Co, H, W) is allocated by the caller (sphere_conv.py:35) and overwritten."""
  if (sH, sW) != (1, 1) or group != 1:
    raise RuntimeError('sphere_conv_cuda (libmode_b200): only stride 1 / groups 1 are built (all MODE instantiates, submodule.py:128-130,161)')
  _guard(input, weight, position, output)
  B, Cin, H, W = input.shape
  Co = weight.shape[0]
  with torch.cuda.device(input.device):  # reference: at::DeviceGuard(input.device()), cpp:136
    _check(_lib.mode_sphere_conv_f32(_ptr(input), _ptr(position), _ptr(weight), None, _ptr(bias) if has_bias else None, None, _ptr(output),
                                     B, Cin, H, W, Co, kH, kW, 0, _stream(input)))


def sphere_conv_backward_cuda(input, weight, bias, ones, position, columns, grad_input, grad_weight, grad_bias, grad_output, kH, kW, sH, sW, pH, pW, dH, dW, group,
                              has_bias):
  """sphere_conv_cuda.cpp:213-219.  The caller passes zero-filled gradients (sphere_conv.py:62-64); the op accumulates."""
  if (sH, sW) != (1, 1) or group != 1:
    raise RuntimeError('sphere_conv_cuda (libmode_b200): only stride 1 / groups 1 are built')
  grad_output = grad_output.contiguous()
  _guard(input, weight, position, grad_output, grad_input, grad_weight)
  B, Cin, H, W = input.shape
  Co = weight.shape[0]
  with torch.cuda.device(input.device):
    _check(_lib.mode_sphere_conv_backward_f32(_ptr(input), _ptr(position), _ptr(weight), _ptr(grad_output), _ptr(grad_input), _ptr(grad_weight),
                                              _ptr(grad_bias) if has_bias else None, B, Cin, H, W, Co, kH, kW, _stream(input)))
