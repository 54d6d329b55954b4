"""torch-facing operators of the MODE hot path, backed by libmode_b200.so through ctypes.

Each op is registered with torch.library (namespace `mode_b200`) with a fake (meta) implementation so it can
be traced / graph-captured, and a CUDA implementation that forwards raw device pointers and the current
stream to the C ABI (include/mode_b200.h).  Outputs are allocated here with the torch caching allocator and
handed to the library (caller-allocates, as in the reference: sphere_conv.py:35).

There is no CPU implementation: calling an op on CPU tensors raises NotImplementedError, like the reference
(`Only support cuda tensor!`, sphere_conv.py:33-34).
"""
from __future__ import annotations

import collections
import ctypes as C
import functools
import os
from typing import Optional

import torch

from . import _lib

_lib.load()  # fail loudly at import time if the extension is missing


def _p(t: Optional[torch.Tensor]):
  return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
  """Current stream of the current device; every op runs under `_device_guard`, which makes the tensors' device current."""
  return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _device_guard(fn):
  """All CUDA tensor arguments must live on ONE device; that device is made current for the duration of the op, so the stream
  handed to the C ABI, the output allocations and the library's per-device kernel attributes all belong to it (a model on
  cuda:1 called while cuda:0 is current would otherwise launch on the wrong device)."""

  @functools.wraps(fn)
  def wrapper(*args, **kwargs):
    dev = None
    for a in list(args) + list(kwargs.values()):
      if isinstance(a, torch.Tensor) and a.is_cuda:
        if dev is None:
          dev = a.device
        elif a.device != dev:
          raise RuntimeError(f'{fn.__name__}: tensor arguments live on different devices ({dev} and {a.device})')
    if dev is None or dev.index == torch.cuda.current_device():
      return fn(*args, **kwargs)
    with torch.cuda.device(dev):
      return fn(*args, **kwargs)

  return wrapper


def _chk(t: torch.Tensor, dtype, name: str):
  if not t.is_cuda:
    raise NotImplementedError(f'{name}: only CUDA tensors are supported (no CPU fallback)')
  if t.dtype != dtype:
    raise TypeError(f'{name}: expected {dtype}, got {t.dtype}')
  return t.contiguous()


def _opt(t: Optional[torch.Tensor], dtype, name: str):
  return None if t is None else _chk(t, dtype, name)


HALF_DTYPES = (torch.bfloat16, torch.float16)


def _fmt(dtype) -> int:
  """MODE_FMT_* code of a 16-bit storage dtype (include/mode_b200.h)."""
  if dtype == torch.bfloat16:
    return 0
  if dtype == torch.float16:
    return 1
  raise TypeError(f'expected torch.bfloat16 or torch.float16, got {dtype}')


# ------------------------------------------------------------------------------------------------
# a4. cost volume
# ------------------------------------------------------------------------------------------------


@torch.library.custom_op('mode_b200::cost_volume', mutates_args=())
@_device_guard
def cost_volume(ref: torch.Tensor, tgt: torch.Tensor, d4: int) -> torch.Tensor:
  """fp32: (B,C,H,W) x2 -> (B,2C,D4,H,W)   [reference layout, models/mode_disparity.py:104-113]
  bf16/fp16: (B,H,W,C) x2 -> (B,D4,H,W,2C)   [NDHWC, consumed by the tensor-core conv3d]"""
  if ref.shape != tgt.shape or ref.dim() != 4:
    raise ValueError('cost_volume: ref/tgt must be 4-D tensors of equal shape')
  if ref.dtype == torch.float32:
    ref, tgt = _chk(ref, torch.float32, 'cost_volume'), _chk(tgt, torch.float32, 'cost_volume')
    B, Cc, H, W = ref.shape
    out = ref.new_empty((B, 2 * Cc, d4, H, W))
    _lib.call('mode_cost_volume_f32', _p(ref), _p(tgt), _p(out), B, Cc, H, W, d4, _stream())
  else:
    _fmt(ref.dtype)
    ref, tgt = _chk(ref, ref.dtype, 'cost_volume'), _chk(tgt, ref.dtype, 'cost_volume')
    B, H, W, Cc = ref.shape
    out = ref.new_empty((B, d4, H, W, 2 * Cc))
    _lib.call('mode_cost_volume_16', _p(ref), _p(tgt), _p(out), B, Cc, H, W, d4, _stream())
  return out


@cost_volume.register_fake
def _(ref, tgt, d4):
  if ref.dtype == torch.float32:
    B, Cc, H, W = ref.shape
    return ref.new_empty((B, 2 * Cc, d4, H, W))
  B, H, W, Cc = ref.shape
  return ref.new_empty((B, d4, H, W, 2 * Cc))


@torch.library.custom_op('mode_b200::cost_volume_backward', mutates_args=())
@_device_guard
def cost_volume_backward(grad_cost: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
  """grad of the fp32 cost volume: (B,2C,D4,H,W) -> (grad_ref, grad_tgt), each (B,C,H,W): gather-sum over the D4 shifts."""
  grad_cost = _chk(grad_cost, torch.float32, 'cost_volume_backward')
  if grad_cost.dim() != 5 or grad_cost.shape[1] % 2:
    raise ValueError('cost_volume_backward: expected a (B,2C,D4,H,W) gradient')
  B, C2, d4, H, W = grad_cost.shape
  gr, gt = grad_cost.new_empty((B, C2 // 2, H, W)), grad_cost.new_empty((B, C2 // 2, H, W))
  _lib.call('mode_cost_volume_backward_f32', _p(grad_cost), _p(gr), _p(gt), B, C2 // 2, H, W, d4, _stream())
  return gr, gt


@cost_volume_backward.register_fake
def _(grad_cost):
  B, C2, d4, H, W = grad_cost.shape
  return grad_cost.new_empty((B, C2 // 2, H, W)), grad_cost.new_empty((B, C2 // 2, H, W))


def _cost_volume_setup(ctx, inputs, output):
  ctx.is_f32 = inputs[0].dtype == torch.float32


def _cost_volume_bwd(ctx, grad):
  if not ctx.is_f32:
    raise NotImplementedError('cost_volume: only the fp32 (NCHW -> NCDHW) layout is differentiable; the 16-bit plans are inference-only')
  gr, gt = cost_volume_backward(grad.contiguous())
  return gr, gt, None


torch.library.register_autograd('mode_b200::cost_volume', _cost_volume_bwd, setup_context=_cost_volume_setup)


@_device_guard
def costvol_conv_weights(weight: torch.Tensor, dtype=torch.bfloat16):
  """dres0[0] weight (32, 64, 3, 3, 3) fp32 -> the two (96, 288) GEMM operands of costvol_conv: row kh*32 + c, column
  (kd*3 + kw)*32 + o, rounded to the 16-bit storage format exactly like the implicit-GEMM weights."""
  weight = _chk(weight, torch.float32, 'costvol_conv_weights')
  if tuple(weight.shape) != (32, 64, 3, 3, 3):
    raise ValueError('costvol_conv_weights: expected (32, 64, 3, 3, 3)')
  halves = []
  for w in (weight[:, :32], weight[:, 32:]):
    halves.append(w.permute(3, 1, 2, 4, 0).reshape(96, 288).to(dtype).contiguous())  # (kh, c, kd, kw, o)
  return halves[0], halves[1]


@torch.library.custom_op('mode_b200::costvol_conv', mutates_args=())
@_device_guard
def costvol_conv(ref: torch.Tensor, tgt: torch.Tensor, wr: torch.Tensor, wt: torch.Tensor, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor], d4: int,
                 relu: bool) -> torch.Tensor:
  """Cost volume + first 3-D conv (64 -> 32) + affine + ReLU without materialising the volume (costvol_conv.cu).
  ref / tgt (B,H,W,32) NHWC 16-bit features, wr / wt from costvol_conv_weights -> (B, d4, H, W, 32) NDHWC 16-bit."""
  fmt = _fmt(ref.dtype)
  ref, tgt = _chk(ref, ref.dtype, 'costvol_conv'), _chk(tgt, ref.dtype, 'costvol_conv')
  wr, wt = _chk(wr, ref.dtype, 'costvol_conv'), _chk(wt, ref.dtype, 'costvol_conv')
  if ref.dim() != 4 or ref.shape != tgt.shape or ref.shape[-1] != 32 or tuple(wr.shape) != (96, 288) or tuple(wt.shape) != (96, 288):
    raise ValueError('costvol_conv: expected (B,H,W,32) features and (96,288) weight matrices')
  B, H, W, _ = ref.shape

  def cols(f):  # (B,H,W,32) -> (B*H*W, 96): channels of rows h-1, h, h+1 (zero rows outside)
    c = torch.empty((B * H * W, 96), dtype=f.dtype, device=f.device)
    _lib.call('mode_costvol_cols', _p(f), _p(c), B, H, W, _stream())
    return c

  ur = torch.mm(cols(ref), wr, out_dtype=torch.float32)  # library GEMMs, fp32 accumulate and output
  ut = torch.mm(cols(tgt), wt, out_dtype=torch.float32)
  out = torch.empty((B, d4, H, W, 32), dtype=ref.dtype, device=ref.device)
  _lib.call('mode_costvol_conv_fused', _p(ur), _p(ut), _p(_opt(scale, torch.float32, 'scale')), _p(_opt(shift, torch.float32, 'shift')), _p(out), B, d4,
            H, W, int(relu), fmt, _stream())
  return out


@costvol_conv.register_fake
def _(ref, tgt, wr, wt, scale, shift, d4, relu):
  return torch.empty((ref.shape[0], d4, ref.shape[1], ref.shape[2], 32), dtype=ref.dtype, device=ref.device)


# ------------------------------------------------------------------------------------------------
# a6/a7. regression + confidence
# ------------------------------------------------------------------------------------------------


@torch.library.custom_op('mode_b200::disp_regress', mutates_args=())
@_device_guard
def disp_regress(cost: torch.Tensor, maxdisp: int, height: int, width: int) -> tuple[torch.Tensor, torch.Tensor]:
  """cost (B,1,D4,H4,W4) or (B,D4,H4,W4) fp32 -> (pred, conf), both (B,1,H,W) fp32
  (models/mode_disparity.py:143-183)."""
  cost = _chk(cost, torch.float32, 'disp_regress')
  if cost.dim() == 5:
    if cost.shape[1] != 1:
      raise ValueError('disp_regress: channel dim must be 1')
    cost = cost[:, 0]
  if cost.dim() != 4:
    raise ValueError('disp_regress: expected a 4-D or 5-D cost tensor')
  B, D4, H4, W4 = cost.shape
  pred = cost.new_empty((B, 1, height, width))
  conf = cost.new_empty((B, 1, height, width))
  _lib.call('mode_disp_regress', _p(cost), _p(pred), _p(conf), B, D4, H4, W4, maxdisp, height, width, _stream())
  return pred, conf


@disp_regress.register_fake
def _(cost, maxdisp, height, width):
  B = cost.shape[0]
  return cost.new_empty((B, 1, height, width)), cost.new_empty((B, 1, height, width))


@torch.library.custom_op('mode_b200::disp_regress_backward', mutates_args=())
@_device_guard
def disp_regress_backward(cost: torch.Tensor, grad_pred: torch.Tensor, maxdisp: int) -> torch.Tensor:
  """d loss / d logits of one soft-argmin head: cost (B,D4,H4,W4) fp32, grad_pred (B,1,H,W) or (B,H,W) -> (B,D4,H4,W4)."""
  cost = _chk(cost, torch.float32, 'disp_regress_backward')
  grad_pred = _chk(grad_pred, torch.float32, 'disp_regress_backward')
  if cost.dim() != 4:
    raise ValueError('disp_regress_backward: expected a (B,D4,H4,W4) cost tensor')
  B, D4, H4, W4 = cost.shape
  H, W = grad_pred.shape[-2:]
  if grad_pred.numel() != B * H * W:
    raise ValueError('disp_regress_backward: grad_pred must be (B,1,H,W)')
  gcost = torch.empty_like(cost)
  nws = _lib.load().mode_disp_regress_backward_workspace_bytes(B, D4, H4, W4, maxdisp, H, W)
  if nws:  # deterministic two-pass path (maxdisp 192, width % 256 == 0): per-pixel dL/dt through a workspace, gather transpose
    ws = torch.empty(nws // 4, dtype=torch.float32, device=cost.device)
    _lib.call('mode_disp_regress_backward_ws', _p(cost), _p(grad_pred), _p(gcost), _p(ws), B, D4, H4, W4, maxdisp, H, W, _stream())
  else:
    _lib.call('mode_disp_regress_backward', _p(cost), _p(grad_pred), _p(gcost), B, D4, H4, W4, maxdisp, H, W, _stream())
  return gcost


@disp_regress_backward.register_fake
def _(cost, grad_pred, maxdisp):
  return torch.empty_like(cost)


def _disp_regress_setup(ctx, inputs, output):
  cost, maxdisp, height, width = inputs
  ctx.save_for_backward(cost)
  ctx.maxdisp = maxdisp


def _disp_regress_bwd(ctx, grad_pred, grad_conf):
  """Gradient flows through the disparity only: the confidence is a rounded-index gather (not differentiable w.r.t. the index; the
  reference computes it in eval mode only, mode_disparity.py:157-183)."""
  (cost,) = ctx.saved_tensors
  shape = cost.shape
  c4 = cost[:, 0] if cost.dim() == 5 else cost
  g = disp_regress_backward(c4.contiguous(), grad_pred.contiguous(), ctx.maxdisp)
  return g.reshape(shape), None, None, None


torch.library.register_autograd('mode_b200::disp_regress', _disp_regress_bwd, setup_context=_disp_regress_setup)


# ------------------------------------------------------------------------------------------------
# f1. training-mode BatchNorm (batch statistics)
# ------------------------------------------------------------------------------------------------
def _bn_layout(x: torch.Tensor, name: str):
  """(x as the kernels read it, N, C, S, channels_last): a dense channels_last / channels_last_3d tensor with a power-of-two channel
  count is processed in place as (N, S, C) -- what cuDNN's tensor-core conv kernels produce and consume without layout transforms;
  everything else as contiguous (N, C, S)."""
  if not x.is_cuda:
    raise NotImplementedError(f'{name}: only CUDA tensors are supported (no CPU fallback)')
  if x.dtype != torch.float32:
    raise TypeError(f'{name}: expected torch.float32, got {x.dtype}')
  if x.dim() < 3:
    raise ValueError(f'{name}: expected (N, C, spatial...) input')
  n, c = x.shape[0], x.shape[1]
  s = x.numel() // (n * c)
  fmt = torch.channels_last if x.dim() == 4 else torch.channels_last_3d if x.dim() == 5 else None
  if fmt is not None and not x.is_contiguous() and x.is_contiguous(memory_format=fmt) and 4 <= c <= 256 and (c & (c - 1)) == 0:
    return x, n, c, s, 1, fmt
  return x.contiguous(), n, c, s, 0, torch.contiguous_format


def _bn_workspace(x: torch.Tensor, n: int, c: int, s: int, cl: int) -> torch.Tensor:
  return torch.empty((_lib.load().mode_batchnorm_workspace_bytes(c, n, s, cl) + 15) // 16 * 2, dtype=torch.float64, device=x.device)


@torch.library.custom_op('mode_b200::batch_norm_train', mutates_args=())
@_device_guard
def batch_norm_train(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor], eps: float) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
  """Training-mode BatchNorm of a (N, C, *spatial) fp32 tensor (nn.BatchNorm2d / 3d as the reference trains them, models/submodule.py:14-30):
  returns (y, batch mean, 1/sqrt(biased var + eps), unbiased batch variance); y keeps x's memory format (contiguous or channels_last).
  Functional: the module (models/batchnorm.py) applies the momentum update of the running statistics from the returned vectors."""
  x, n, c, s, cl, _ = _bn_layout(x, 'batch_norm_train')
  weight, bias = _opt(weight, torch.float32, 'batch_norm_train'), _opt(bias, torch.float32, 'batch_norm_train')
  y = torch.empty_like(x)  # dense input: same strides
  mean, invstd, var_u = x.new_empty(c), x.new_empty(c), x.new_empty(c)
  ws = _bn_workspace(x, n, c, s, cl)
  _lib.call('mode_batchnorm_train_fwd_f32', _p(x), _p(weight), _p(bias), _p(y), _p(mean), _p(invstd), _p(var_u), None, None, _p(ws), n, c, s, cl, float(eps), 0.0, _stream())
  return y, mean, invstd, var_u


@batch_norm_train.register_fake
def _(x, weight, bias, eps):
  c = x.shape[1]
  return torch.empty_like(x), x.new_empty(c), x.new_empty(c), x.new_empty(c)


@torch.library.custom_op('mode_b200::batch_norm_train_backward', mutates_args=())
@_device_guard
def batch_norm_train_backward(x: torch.Tensor, grad_y: torch.Tensor, weight: Optional[torch.Tensor], save_mean: torch.Tensor,
                              save_invstd: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
  """(grad_x, grad_weight, grad_bias) of batch_norm_train."""
  x, n, c, s, cl, fmt = _bn_layout(x, 'batch_norm_train_backward')
  if grad_y.dtype != torch.float32 or grad_y.shape != x.shape:
    raise TypeError('batch_norm_train_backward: grad_y must be an fp32 tensor of x\'s shape')
  grad_y = grad_y.contiguous(memory_format=fmt)  # same memory format as x (no copy when the producer already wrote it that way)
  gx = torch.empty_like(x)
  gw, gb = x.new_empty(c), x.new_empty(c)
  ws = _bn_workspace(x, n, c, s, cl)
  _lib.call('mode_batchnorm_train_bwd_f32', _p(x), _p(grad_y), _p(_opt(weight, torch.float32, 'batch_norm_train_backward')), _p(save_mean), _p(save_invstd), _p(gx), _p(gw),
            _p(gb), _p(ws), n, c, s, cl, _stream())
  return gx, gw, gb


@batch_norm_train_backward.register_fake
def _(x, grad_y, weight, save_mean, save_invstd):
  return torch.empty_like(x), x.new_empty(x.shape[1]), x.new_empty(x.shape[1])


def _bn_setup(ctx, inputs, output):
  x, weight, bias, eps = inputs
  _, mean, invstd, _ = output
  ctx.save_for_backward(x, weight, mean, invstd)
  ctx.has_weight, ctx.has_bias = weight is not None, bias is not None


def _bn_bwd(ctx, grad_y, grad_mean, grad_invstd, grad_var):
  """Gradient through y only: the statistics outputs feed the (non-differentiable) running-stat update."""
  x, weight, mean, invstd = ctx.saved_tensors
  gx, gw, gb = batch_norm_train_backward(x, grad_y, weight, mean, invstd)
  return gx, (gw if ctx.has_weight else None), (gb if ctx.has_bias else None), None


torch.library.register_autograd('mode_b200::batch_norm_train', _bn_bwd, setup_context=_bn_setup)


# ------------------------------------------------------------------------------------------------
# a2. spherical convolution
# ------------------------------------------------------------------------------------------------


@torch.library.custom_op('mode_b200::sphere_conv_f32', mutates_args=())
@_device_guard
def sphere_conv_f32(x: torch.Tensor, pos: torch.Tensor, weight: torch.Tensor, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor],
                    residual: Optional[torch.Tensor], relu: bool) -> torch.Tensor:
  """fp32 NCHW spherical conv with fused per-channel affine (+residual)(+ReLU).
  Plain reference semantics (sphere_conv_cuda.cpp:129-210): scale=None, shift=bias, residual=None, relu=False."""
  if x.dim() != 4:
    raise ValueError('Expected 4D tensor as input, got {}D tensor instead.'.format(x.dim()))  # sphere_conv.py:19-20
  if not x.is_cuda:
    raise NotImplementedError('Only support cuda tensor!')  # sphere_conv.py:33-34
  x = _chk(x, torch.float32, 'sphere_conv')
  weight = _chk(weight, torch.float32, 'sphere_conv')
  pos = _chk(pos, torch.float32, 'sphere_conv')
  B, Cc, H, W = x.shape
  Co, Cw, Kh, Kw = weight.shape
  if Cw != Cc:
    raise RuntimeError(f'Input shape and kernel channels wont match: ({Cc} vs {Cw}).')  # cpp:154-157
  if pos.numel() != 2 * Kh * Kw * H * W:
    raise RuntimeError(f'invalid spatial size of position, expected {2 * Kh * Kw}x{H}x{W}, got {tuple(pos.shape)}')  # cpp:93-101
  out = x.new_empty((B, Co, H, W))
  _lib.call('mode_sphere_conv_f32', _p(x), _p(pos), _p(weight), _p(_opt(scale, torch.float32, 'scale')), _p(_opt(shift, torch.float32, 'shift')),
            _p(_opt(residual, torch.float32, 'residual')), _p(out), B, Cc, H, W, Co, Kh, Kw, int(relu), _stream())
  return out


@sphere_conv_f32.register_fake
def _(x, pos, weight, scale, shift, residual, relu):
  return x.new_empty((x.shape[0], weight.shape[0], x.shape[2], x.shape[3]))


@torch.library.custom_op('mode_b200::sphere_conv_backward', mutates_args=())
@_device_guard
def sphere_conv_backward(x: torch.Tensor, pos: torch.Tensor, weight: torch.Tensor, grad_out: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
  """(grad_input, grad_weight, grad_bias) of the plain fp32 spherical conv, as a traceable op (see sphere_conv_backward_f32)."""
  gi, gw, gb = sphere_conv_backward_f32(x, pos, weight, grad_out, True, True, True)
  return gi, gw, gb


@sphere_conv_backward.register_fake
def _(x, pos, weight, grad_out):
  return torch.empty_like(x), torch.empty_like(weight), x.new_empty((weight.shape[0],))


def _sphere_f32_setup(ctx, inputs, output):
  x, pos, weight, scale, shift, residual, relu = inputs
  ctx.plain = scale is None and residual is None and not relu
  ctx.has_bias = shift is not None
  ctx.save_for_backward(x, pos, weight)


def _sphere_f32_bwd(ctx, grad_out):
  if not ctx.plain:
    raise NotImplementedError('sphere_conv_f32: only the plain form (scale=None, residual=None, relu=False; shift = bias) is differentiable')
  x, pos, weight = ctx.saved_tensors
  gi, gw, gb = sphere_conv_backward(x, pos, weight, grad_out.contiguous())
  return gi, None, gw, None, (gb if ctx.has_bias else None), None, None


torch.library.register_autograd('mode_b200::sphere_conv_f32', _sphere_f32_bwd, setup_context=_sphere_f32_setup)


@_device_guard
def sphere_conv_backward_f32(x: torch.Tensor, pos: torch.Tensor, weight: torch.Tensor, grad_out: torch.Tensor, need_input: bool, need_weight: bool,
                             need_bias: bool):
  """Gradients of sphere_conv_f32 (plain semantics: scale=None, shift=bias): (grad_input, grad_weight, grad_bias), None where
  not requested.  Mirrors sphere_conv_backward_cuda (sphere_conv_cuda.cpp:213-336): zero-filled buffers, op accumulates."""
  if not grad_out.is_cuda:
    raise NotImplementedError  # sphere_conv.py:66-67
  x, pos, weight, grad_out = (_chk(t, torch.float32, 'sphere_conv_backward') for t in (x, pos, weight, grad_out))
  B, Cc, H, W = x.shape
  Co, _, Kh, Kw = weight.shape
  if grad_out.shape != (B, Co, H, W):
    raise RuntimeError(f'invalid batch size / shape of grad_output, expected {(B, Co, H, W)}, got {tuple(grad_out.shape)}')  # cpp:107-123
  gi = torch.zeros_like(x) if need_input else None
  gw = torch.zeros_like(weight) if need_weight else None
  gb = torch.zeros(Co, dtype=torch.float32, device=x.device) if need_bias else None
  if os.environ.get('MODE_B200_NONDETERMINISTIC_BWD'):  # the reference's scheme: fp32 atomicAdd scatter, last bits vary from run to run
    _lib.call('mode_sphere_conv_backward_f32', _p(x), _p(pos), _p(weight), _p(grad_out), _p(gi), _p(gw), _p(gb), B, Cc, H, W, Co, Kh, Kw, _stream())
  else:  # default: order-free 64-bit fixed-point accumulation -> bit-identical gradients from run to run
    ws = torch.empty(_lib.load().mode_sphere_conv_backward_workspace_bytes(B, Cc, H, W, Co, Kh, Kw), dtype=torch.uint8, device=x.device)
    _lib.call('mode_sphere_conv_backward_det_f32', _p(x), _p(pos), _p(weight), _p(grad_out), _p(gi), _p(gw), _p(gb), _p(ws), B, Cc, H, W, Co, Kh, Kw, _stream())
  return gi, gw, gb


@_device_guard
def sphere_conv_pack_weights(weight: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
  """(Co,C,3,3) fp32 -> 16-bit weight slabs [tap][C/64][8][Co][8] streamed by the tensor-core kernel."""
  weight = _chk(weight, torch.float32, 'sphere_conv_pack_weights')
  Co, Cc, Kh, Kw = weight.shape
  if (Kh, Kw) != (3, 3):
    raise ValueError('sphere_conv_pack_weights: 3x3 kernels only')
  out = torch.empty(9 * Cc * Co, dtype=dtype, device=weight.device)
  _lib.call('mode_sphere_conv_pack_weights', _p(weight), _p(out), Cc, Co, _fmt(dtype), _stream())
  return out


_TABLES = collections.OrderedDict()  # (grid storage, shape, version, 16-bit format) -> (table, grid); LRU-bounded
_TABLES_MAX = 16


def _publish(stream_sync: bool = True):
  """A cached constant built on the current stream may be consumed from any other stream later (HostPipeline's compute
  stream, per-GPU threads): make it visible to all of them once, at creation (skipped inside a CUDA-graph capture, where
  the warm-up pass has already built every entry)."""
  if stream_sync and not torch.cuda.is_current_stream_capturing():
    torch.cuda.current_stream().synchronize()


@_device_guard
def sphere_gather_table(pos: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
  """Pre-digested sampling grid for the tensor-core kernel (cached per grid tensor and 16-bit format): see
  mode_sphere_conv_build_table."""
  pos = _chk(pos, torch.float32, 'sphere_gather_table')
  key = (pos.data_ptr(), tuple(pos.shape), pos._version, pos.device.index)  # the table does not depend on the 16-bit format
  hit = _TABLES.get(key)
  if hit is None:
    H, W = pos.shape[-2:]
    if pos.numel() != 18 * H * W:
      raise RuntimeError(f'invalid spatial size of position, expected 18x{H}x{W}, got {tuple(pos.shape)}')
    table = torch.empty(_lib.load().mode_sphere_conv_table_bytes(H, W, 3, 3), dtype=torch.uint8, device=pos.device)
    _lib.call('mode_sphere_conv_build_table', _p(pos), _p(table), H, W, 3, 3, _fmt(dtype), _stream())
    _publish()
    hit = _TABLES[key] = (table, pos)  # keeps `pos` alive while the entry lives, so the data_ptr key stays unique
    while len(_TABLES) > _TABLES_MAX:
      _TABLES.popitem(last=False)
  else:
    _TABLES.move_to_end(key)
  return hit[0]


@torch.library.custom_op('mode_b200::sphere_conv_bf16', mutates_args=())
@_device_guard
def sphere_conv_bf16(x: torch.Tensor, pos: torch.Tensor, w_packed: torch.Tensor, cout: int, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor],
                     residual: Optional[torch.Tensor], relu: bool) -> torch.Tensor:
  """NHWC bf16/fp16 spherical conv on tcgen05 tensor cores (gather fused into operand staging) + affine + residual
  + ReLU.  x (B,H,W,C) -> (B,H,W,cout), same 16-bit dtype."""
  if x.dim() != 4:
    raise ValueError('Expected 4D tensor as input, got {}D tensor instead.'.format(x.dim()))
  fmt = _fmt(x.dtype)
  x = _chk(x, x.dtype, 'sphere_conv_bf16')
  pos = _chk(pos, torch.float32, 'sphere_conv_bf16')
  w_packed = _chk(w_packed, x.dtype, 'sphere_conv_bf16')
  B, H, W, Cc = x.shape
  if pos.numel() != 18 * H * W:
    raise RuntimeError(f'invalid spatial size of position, expected 18x{H}x{W}, got {tuple(pos.shape)}')
  if w_packed.numel() != 9 * Cc * cout:
    raise RuntimeError('sphere_conv_bf16: packed weight size does not match (C, cout)')
  out = torch.empty((B, H, W, cout), dtype=x.dtype, device=x.device)
  if residual is not None and residual.shape != out.shape:
    raise RuntimeError('sphere_conv_bf16: residual shape mismatch')
  _lib.call('mode_sphere_conv_tc', _p(x), _p(sphere_gather_table(pos, x.dtype)), _p(w_packed), _p(_opt(scale, torch.float32, 'scale')), _p(_opt(shift, torch.float32, 'shift')),
            _p(_opt(residual, x.dtype, 'residual')), _p(out), B, Cc, H, W, cout, int(relu), fmt, _stream())
  return out


@sphere_conv_bf16.register_fake
def _(x, pos, w_packed, cout, scale, shift, residual, relu):
  return torch.empty((*x.shape[:3], cout), dtype=x.dtype, device=x.device)


@torch.library.custom_op('mode_b200::stem_conv', mutates_args=())
@_device_guard
def stem_conv(x0: torch.Tensor, x1: Optional[torch.Tensor], weight: torch.Tensor, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor], relu: bool,
              fp16: bool) -> torch.Tensor:
  """firstconv[0] of the MODE feature extractor (3 -> 32, 7x7, stride 2, pad 3) + affine + ReLU on tcgen05 tensor cores.
  x0 (B0,3,H,W) and optional x1 (B1,3,H,W) fp32 NCHW (left / right batches, no concatenation) -> (B0+B1, Ho, Wo, 32) NHWC
  bf16 (fp16 if `fp16`)."""
  if x0.dim() != 4 or x0.shape[1] != 3:
    raise ValueError('stem_conv: expected (B,3,H,W) images')
  x0 = _chk(x0, torch.float32, 'stem_conv')
  if x1 is not None:
    x1 = _chk(x1, torch.float32, 'stem_conv')
    if x1.shape[1:] != x0.shape[1:]:
      raise ValueError('stem_conv: x0 / x1 image shapes differ')
  weight = _chk(weight, torch.float32, 'stem_conv')
  if tuple(weight.shape) != (32, 3, 7, 7):
    raise ValueError('stem_conv: weight must be (32,3,7,7)')
  dtype = torch.float16 if fp16 else torch.bfloat16
  B0, _, H, W = x0.shape
  B1 = 0 if x1 is None else x1.shape[0]
  out = torch.empty((B0 + B1, (H - 1) // 2 + 1, (W - 1) // 2 + 1, 32), dtype=dtype, device=x0.device)
  _lib.call('mode_stem_conv_tc', _p(x0), _p(x1) if x1 is not None else None, _p(weight), _p(_opt(scale, torch.float32, 'scale')), _p(_opt(shift, torch.float32, 'shift')),
            _p(out), B0, B1, H, W, int(relu), _fmt(dtype), _stream())
  return out


@stem_conv.register_fake
def _(x0, x1, weight, scale, shift, relu, fp16):
  B = x0.shape[0] + (0 if x1 is None else x1.shape[0])
  return torch.empty((B, (x0.shape[2] - 1) // 2 + 1, (x0.shape[3] - 1) // 2 + 1, 32), dtype=torch.float16 if fp16 else torch.bfloat16, device=x0.device)


# ------------------------------------------------------------------------------------------------
# a5. conv3d family
# ------------------------------------------------------------------------------------------------

CONV_S1, CONV_S2, DECONV_S2 = 0, 1, 2


def conv3d_out_dims(d, h, w, mode):
  if mode == CONV_S1:
    return d, h, w
  if mode == CONV_S2:
    return (d - 1) // 2 + 1, (h - 1) // 2 + 1, (w - 1) // 2 + 1
  return 2 * d, 2 * h, 2 * w


@torch.library.custom_op('mode_b200::conv3d_f32', mutates_args=())
@_device_guard
def conv3d_f32(x: torch.Tensor, weight: torch.Tensor, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor],
               residual: Optional[torch.Tensor], mode: int, relu: bool) -> torch.Tensor:
  """fp32 NCDHW 3x3x3 conv (mode 0/1) or transposed conv (mode 2) + affine + residual + ReLU
  (models/submodule.py:20-22, models/mode_disparity.py:15-25)."""
  x = _chk(x, torch.float32, 'conv3d')
  weight = _chk(weight, torch.float32, 'conv3d')
  if x.dim() != 5 or weight.dim() != 5 or tuple(weight.shape[2:]) != (3, 3, 3):
    raise ValueError('conv3d: expected 5-D input and a 3x3x3 kernel')
  B, Ci, D, H, W = x.shape
  Co = weight.shape[1] if mode == DECONV_S2 else weight.shape[0]
  if (weight.shape[0] if mode == DECONV_S2 else weight.shape[1]) != Ci:
    raise RuntimeError('conv3d: input channels do not match the kernel')
  Do, Ho, Wo = conv3d_out_dims(D, H, W, mode)
  out = x.new_empty((B, Co, Do, Ho, Wo))
  if residual is not None and residual.shape != out.shape:
    raise RuntimeError('conv3d: residual shape mismatch')
  _lib.call('mode_conv3d_f32', _p(x), _p(weight), _p(_opt(scale, torch.float32, 'scale')), _p(_opt(shift, torch.float32, 'shift')),
            _p(_opt(residual, torch.float32, 'residual')), _p(out), B, Ci, Co, D, H, W, mode, int(relu), _stream())
  return out


@conv3d_f32.register_fake
def _(x, weight, scale, shift, residual, mode, relu):
  B, Ci, D, H, W = x.shape
  Co = weight.shape[1] if mode == DECONV_S2 else weight.shape[0]
  return x.new_empty((B, Co, *conv3d_out_dims(D, H, W, mode)))


@_device_guard
def conv3d_pack_weights(weight: torch.Tensor, mode: int, dtype=torch.bfloat16) -> torch.Tensor:
  """fp32 PyTorch-layout 3x3x3 weights -> per-tap 16-bit tiles resident in shared memory."""
  weight = _chk(weight, torch.float32, 'conv3d_pack_weights')
  Ci, Co = (weight.shape[0], weight.shape[1]) if mode == DECONV_S2 else (weight.shape[1], weight.shape[0])
  n = _lib.load().mode_conv3d_packed_weight_elems(Ci, Co, mode)
  out = torch.empty(n, dtype=dtype, device=weight.device)
  _lib.call('mode_conv3d_pack_weights', _p(weight), _p(out), Ci, Co, mode, _fmt(dtype), _stream())
  return out


@torch.library.custom_op('mode_b200::conv3d_bf16', mutates_args=())
@_device_guard
def conv3d_bf16(x: torch.Tensor, w_packed: torch.Tensor, cout: int, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor],
                residual: Optional[torch.Tensor], mode: int, relu: bool, out_f32: bool) -> torch.Tensor:
  """NDHWC bf16/fp16 3x3x3 conv / strided conv / transposed conv on tcgen05 tensor cores with fused affine +
  residual + ReLU.  x (B,D,H,W,Ci) -> (B,Do,Ho,Wo,cout) in the same 16-bit dtype; with out_f32 (cout <= 16, the
  32->1 classifier) the output and the residual are fp32."""
  fmt = _fmt(x.dtype)
  x = _chk(x, x.dtype, 'conv3d_bf16')
  w_packed = _chk(w_packed, x.dtype, 'conv3d_bf16')
  if x.dim() != 5:
    raise ValueError('conv3d_bf16: expected (B,D,H,W,C) input')
  B, D, H, W, Ci = x.shape
  Do, Ho, Wo = conv3d_out_dims(D, H, W, mode)
  out = torch.empty((B, Do, Ho, Wo, cout), dtype=torch.float32 if out_f32 else x.dtype, device=x.device)
  if residual is not None and residual.shape != out.shape:
    raise RuntimeError('conv3d_bf16: residual shape mismatch')
  res_bf16 = _opt(residual, x.dtype, 'residual') if not out_f32 else None
  res_f32 = _opt(residual, torch.float32, 'residual') if out_f32 else None
  _lib.call('mode_conv3d_tc', _p(x), _p(w_packed), _p(_opt(scale, torch.float32, 'scale')), _p(_opt(shift, torch.float32, 'shift')), _p(res_bf16),
            _p(res_f32), _p(None if out_f32 else out), _p(out if out_f32 else None), B, Ci, cout, D, H, W, mode, int(relu), fmt, _stream())
  return out


@conv3d_bf16.register_fake
def _(x, w_packed, cout, scale, shift, residual, mode, relu, out_f32):
  B, D, H, W, Ci = x.shape
  return torch.empty((B, *conv3d_out_dims(D, H, W, mode), cout), dtype=torch.float32 if out_f32 else x.dtype, device=x.device)


@torch.library.custom_op('mode_b200::conv3d_classifier', mutates_args=())
@_device_guard
def conv3d_classifier(x: torch.Tensor, weight: torch.Tensor, residual: Optional[torch.Tensor]) -> torch.Tensor:
  """The 32 -> 1 classifier conv (3x3x3, pad 1, no bias) + fp32 residual as a pointwise tensor-core GEMM + shifted sum.
  x (B,D,H,W,32) bf16/fp16, weight (1,32,3,3,3) fp32, residual (B,D,H,W) fp32 or None -> (B,D,H,W) fp32."""
  fmt = _fmt(x.dtype)
  x = _chk(x, x.dtype, 'conv3d_classifier')
  weight = _chk(weight, torch.float32, 'conv3d_classifier')
  if x.dim() != 5 or x.shape[-1] != 32 or tuple(weight.shape) != (1, 32, 3, 3, 3):
    raise ValueError('conv3d_classifier: expected x (B,D,H,W,32) and weight (1,32,3,3,3)')
  B, D, H, W, _ = x.shape
  out = torch.empty((B, D, H, W), dtype=torch.float32, device=x.device)
  if residual is not None and residual.shape != out.shape:
    raise RuntimeError('conv3d_classifier: residual shape mismatch')
  _lib.call('mode_conv3d_classifier_tc', _p(x), _p(weight), _p(_opt(residual, torch.float32, 'residual')), _p(out), B, D, H, W, fmt, _stream())
  return out


@conv3d_classifier.register_fake
def _(x, weight, residual):
  return torch.empty(x.shape[:4], dtype=torch.float32, device=x.device)


# ------------------------------------------------------------------------------------------------
# layout helpers
# ------------------------------------------------------------------------------------------------


@_device_guard
def nchw_f32_to_nhwc_bf16(x: torch.Tensor, dtype=torch.bfloat16) -> torch.Tensor:
  """(B,C,*spatial) fp32 -> (B,*spatial,C) bf16 (or fp16 with dtype=torch.float16)."""
  x = _chk(x, torch.float32, 'nchw_f32_to_nhwc_bf16')
  B, Cc = x.shape[:2]
  sp = tuple(x.shape[2:])
  hw = 1
  for s in sp:
    hw *= s
  y = torch.empty((B, *sp, Cc), dtype=dtype, device=x.device)
  _lib.call('mode_nchw_f32_to_nhwc_16', _p(x), _p(y), B, Cc, hw, _fmt(dtype), _stream())
  return y


@_device_guard
def nhwc_bf16_to_nchw_f32(x: torch.Tensor) -> torch.Tensor:
  """(B,*spatial,C) bf16/fp16 -> (B,C,*spatial) fp32."""
  x = _chk(x, x.dtype, 'nhwc_bf16_to_nchw_f32')
  B, Cc = x.shape[0], x.shape[-1]
  sp = tuple(x.shape[1:-1])
  hw = 1
  for s in sp:
    hw *= s
  y = torch.empty((B, Cc, *sp), dtype=torch.float32, device=x.device)
  _lib.call('mode_nhwc_16_to_nchw_f32', _p(x), _p(y), B, Cc, hw, _fmt(x.dtype), _stream())
  return y


@_device_guard
def concat3_nhwc(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor) -> torch.Tensor:
  """Channel concatenation of three NHWC 16-bit maps (B,H,W,Ca|Cb|Cc) -> (B,H,W,Ca+Cb+Cc): a streaming copy kernel."""
  a = _chk(a, a.dtype, 'concat3_nhwc')
  b, c = _chk(b, a.dtype, 'concat3_nhwc'), _chk(c, a.dtype, 'concat3_nhwc')
  if a.shape[:-1] != b.shape[:-1] or a.shape[:-1] != c.shape[:-1]:
    raise ValueError('concat3_nhwc: leading dimensions differ')
  out = torch.empty((*a.shape[:-1], a.shape[-1] + b.shape[-1] + c.shape[-1]), dtype=a.dtype, device=a.device)
  npix = 1
  for s in a.shape[:-1]:
    npix *= s
  _lib.call('mode_concat3_nhwc_16', _p(a), _p(b), _p(c), _p(out), npix, a.shape[-1], b.shape[-1], c.shape[-1], _stream())
  return out


# ------------------------------------------------------------------------------------------------
# geometry (a8-a11)
# ------------------------------------------------------------------------------------------------


@_device_guard
def disp_to_depth(disp: torch.Tensor, phi_l: torch.Tensor, baseline: float, want_f64: bool = False):
  """disp (...,H,W) fp32 -> depth fp32 [, depth fp64 (un-rounded, for the forward warp)]."""
  disp = _chk(disp, torch.float32, 'disp_to_depth')
  phi_l = _chk(phi_l, torch.float32, 'disp_to_depth')
  H, W = disp.shape[-2:]
  if phi_l.numel() != W:
    raise ValueError('disp_to_depth: phi_l must have W entries')
  out = torch.empty_like(disp)
  out64 = torch.empty_like(disp, dtype=torch.float64) if want_f64 else None
  _lib.call('mode_disp_to_depth', _p(disp), _p(phi_l), _p(out), _p(out64), disp.numel() // (H * W), H, W, C.c_float(baseline), _stream())
  return (out, out64) if want_f64 else out


@_device_guard
def grid_sample_border(src: torch.Tensor, grid: torch.Tensor) -> torch.Tensor:
  """src (N,C,Hs,Ws) fp32, grid (Ho,Wo,2) fp32 shared by the batch -> (N,C,Ho,Wo)."""
  src = _chk(src, torch.float32, 'grid_sample_border')
  grid = _chk(grid, torch.float32, 'grid_sample_border')
  if src.dim() != 4 or grid.dim() != 3 or grid.shape[-1] != 2:
    raise ValueError('grid_sample_border: src must be (N,C,Hs,Ws) and grid (Ho,Wo,2)')
  N, Cc, Hs, Ws = src.shape
  Ho, Wo = grid.shape[:2]
  out = src.new_empty((N, Cc, Ho, Wo))
  _lib.call('mode_grid_sample_border', _p(src), _p(grid), _p(out), N, Cc, Hs, Ws, Ho, Wo, _stream())
  return out


@_device_guard
def depth_view_trans(depth: torch.Tensor, conf: torch.Tensor, sin_phi, cos_phi, sin_theta, cos_theta, Rt) -> tuple[torch.Tensor, torch.Tensor]:
  """depth (B,H,W) fp32 or fp64, conf (B,H,W) fp32; tables fp32 on device; Rt: 12 python floats (R row-major, then t)."""
  if depth.dtype not in (torch.float32, torch.float64):
    raise TypeError('depth_view_trans: depth must be fp32 or fp64')
  depth = _chk(depth, depth.dtype, 'depth_view_trans')
  conf = _chk(conf, torch.float32, 'depth_view_trans')
  if depth.shape != conf.shape or depth.dim() != 3:
    raise ValueError('depth_view_trans: depth/conf must be (B,H,W) tensors of equal shape')
  B, H, W = depth.shape
  ws = torch.empty((3, B, H, W), dtype=torch.int32, device=depth.device)
  v2, c2 = torch.empty_like(conf), torch.empty_like(conf)
  rt = (C.c_double * 12)(*[float(v) for v in Rt])
  d32, d64 = (depth, None) if depth.dtype == torch.float32 else (None, depth)
  _lib.call('mode_depth_view_trans', _p(d32), _p(d64), _p(conf), _p(sin_phi), _p(cos_phi), _p(sin_theta), _p(cos_theta), rt, _p(ws), _p(v2), _p(c2), B, H, W,
            _stream())
  return v2, c2
