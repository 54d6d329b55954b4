"""ctypes binding of libmode_b200.so (C ABI declared in include/mode_b200.h).

There is no CPU or PyTorch fallback: if the library is missing or a call fails, a RuntimeError is raised.
ctypes releases the GIL for the duration of each call (reference: GIL held, sphere_conv_cuda.cpp has no
gil_scoped_release), so per-GPU Python threads can drive separate streams.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('MODE_B200_LIB') or os.path.join(_HERE, 'lib', 'libmode_b200.so')  # override: kernel experiments (tools/build_variant.sh)

_vp, _i, _f = C.c_void_p, C.c_int, C.c_float

# name -> argtypes; every function returns int status (see include/mode_b200.h)
SIGNATURES = {
    'mode_cost_volume_f32': [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    'mode_cost_volume_16': [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    'mode_cost_volume_backward_f32': [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    'mode_conv3d_classifier_tc': [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    'mode_costvol_cols': [_vp, _vp, _i, _i, _i, _vp],
    'mode_costvol_conv_fused': [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    'mode_stem_conv_tc': [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    'mode_disp_regress': [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    'mode_disp_regress_backward': [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    'mode_disp_regress_backward_ws': [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    'mode_sphere_conv_f32': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    'mode_sphere_conv_tc': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    'mode_sphere_conv_pack_weights': [_vp, _vp, _i, _i, _i, _vp],
    'mode_sphere_conv_backward_f32': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    'mode_sphere_conv_backward_det_f32': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    'mode_sphere_conv_build_table': [_vp, _vp, _i, _i, _i, _i, _i, _vp],
    'mode_conv3d_f32': [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    'mode_conv3d_pack_weights': [_vp, _vp, _i, _i, _i, _i, _vp],
    'mode_conv3d_tc': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    'mode_nchw_f32_to_nhwc_16': [_vp, _vp, _i, _i, _i, _i, _vp],
    'mode_nhwc_16_to_nchw_f32': [_vp, _vp, _i, _i, _i, _i, _vp],
    'mode_concat3_nhwc_16': [_vp, _vp, _vp, _vp, C.c_longlong, _i, _i, _i, _vp],
    'mode_disp_to_depth': [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _vp],
    'mode_grid_sample_border': [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp],
    'mode_batchnorm_train_fwd_f32': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_longlong, _i, C.c_longlong, _i, _f, _f, _vp],
    'mode_batchnorm_train_bwd_f32': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, C.c_longlong, _i, C.c_longlong, _i, _vp],
    'mode_depth_view_trans': [_vp, _vp, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_double), _vp, _vp, _vp, _i, _i, _i, _vp],
}
OTHER_SYMBOLS = ['mode_disp_regress_backward_workspace_bytes', 'mode_batchnorm_workspace_bytes', 'mode_sphere_conv_table_bytes', 'mode_sphere_conv_backward_workspace_bytes', 'mode_conv3d_set_debug_buffer', 'mode_b200_version', 'mode_b200_last_error', 'mode_b200_launch_count', 'mode_conv3d_packed_weight_elems']

PENDING = set()
_lib = None


def load() -> C.CDLL:
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise RuntimeError(f'{LIB_PATH} is not built: run `python -m mode_2022_b200.build` (there is no fallback path)')
  lib = C.CDLL(LIB_PATH)
  for name, argtypes in SIGNATURES.items():
    if name in PENDING:
      continue
    fn = getattr(lib, name)
    fn.argtypes = argtypes
    fn.restype = _i
  lib.mode_b200_version.restype = _i
  lib.mode_b200_last_error.restype = C.c_char_p
  lib.mode_b200_launch_count.restype = C.c_ulonglong
  if 'mode_conv3d_packed_weight_elems' not in PENDING:
    lib.mode_conv3d_packed_weight_elems.argtypes = [_i, _i, _i]
    lib.mode_conv3d_packed_weight_elems.restype = C.c_size_t
  lib.mode_sphere_conv_table_bytes.argtypes = [_i, _i, _i, _i]
  lib.mode_disp_regress_backward_workspace_bytes.argtypes = [_i] * 7
  lib.mode_disp_regress_backward_workspace_bytes.restype = C.c_size_t
  lib.mode_batchnorm_workspace_bytes.argtypes = [_i, C.c_longlong, C.c_longlong, _i]
  lib.mode_batchnorm_workspace_bytes.restype = C.c_size_t
  lib.mode_sphere_conv_backward_workspace_bytes.argtypes = [_i] * 7
  lib.mode_sphere_conv_backward_workspace_bytes.restype = C.c_size_t
  lib.mode_sphere_conv_table_bytes.restype = C.c_size_t
  _lib = lib
  return lib


# optional per-call CUDA-event profiler (bench.py roofline leg): list of (name, start_event, end_event)
PROFILE = None


def call(name: str, *args) -> None:
  lib = load()
  if PROFILE is not None:
    import torch
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    rc = getattr(lib, name)(*args)
    b.record()
    PROFILE.append((name, args, a, b))
  else:
    rc = getattr(lib, name)(*args)
  if rc != 0:
    raise RuntimeError(f'{name} failed ({rc}): {lib.mode_b200_last_error().decode()}')


def launch_count() -> int:
  return int(load().mode_b200_launch_count())
