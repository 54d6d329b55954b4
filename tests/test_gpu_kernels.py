"""GPU parity tests: every libmode_b200 kernel (through the C ABI, via mode_2022_b200.ops) against the CPU oracle
on the same seeded inputs, plus size-independent properties at BASELINE.json's full size."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import mode_oracle as O
from tests import helpers as Hh

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops():
  torch.backends.cudnn.allow_tf32 = False
  torch.backends.cuda.matmul.allow_tf32 = False
  from mode_2022_b200 import ops as _ops
  return _ops


# ---------------------------------------------------------------------------- a4 cost volume
@pytest.mark.parametrize('shape,d4', [((1, 32, 16, 8), 4), ((2, 32, 8, 16), 16), ((1, 8, 4, 4), 1), ((1, 32, 4, 8), 12)])
def test_cost_volume_f32_bit_exact(ops, shape, d4):
  g = torch.Generator().manual_seed(0)
  ref, tgt = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
  out = ops.cost_volume(ref.cuda(), tgt.cuda(), d4).cpu()
  assert torch.equal(out, O.cost_volume(ref, tgt, d4))  # pure copy: bit-exact (d4 > W exercises all-zero planes)


def test_cost_volume_bf16_bit_exact(ops):
  g = torch.Generator().manual_seed(1)
  ref, tgt = torch.randn(2, 32, 16, 8, generator=g).bfloat16(), torch.randn(2, 32, 16, 8, generator=g).bfloat16()
  out = ops.cost_volume(ref.permute(0, 2, 3, 1).contiguous().cuda(), tgt.permute(0, 2, 3, 1).contiguous().cuda(), 4).cpu()
  want = O.cost_volume(ref.float(), tgt.float(), 4).permute(0, 2, 3, 4, 1)  # NCDHW -> NDHWC
  assert torch.equal(out.float(), want)


def test_cost_volume_full_size_properties(ops):
  """1024x512 / D=192 (C1 shape): shift structure checked without a CPU copy of the 403 MB volume."""
  ref, tgt = torch.randn(1, 32, 256, 128, device='cuda'), torch.randn(1, 32, 256, 128, device='cuda')
  c = ops.cost_volume(ref, tgt, 48)
  assert c.shape == (1, 64, 48, 256, 128)
  for i in (0, 1, 17, 47):
    assert torch.equal(c[:, :32, i, :, i:], ref[..., i:]) and torch.equal(c[:, 32:, i, :, i:], tgt[..., :128 - i])
    assert (c[:, :, i, :, :i] == 0).all()


# ---------------------------------------------------------------------------- a6/a7 regression
@pytest.mark.parametrize('d4,h4,w4', [(4, 16, 8), (8, 8, 16), (16, 32, 16), (48, 8, 4)])
def test_disp_regress(ops, d4, h4, w4):
  g = torch.Generator().manual_seed(d4)
  cost = torch.randn(2, 1, d4, h4, w4, generator=g) * 4
  D, H, W = 4 * d4, 4 * h4, 4 * w4
  pred_o, conf_o = O.disparity_regression(cost, D, H, W, want_conf=True)
  pred, conf = ops.disp_regress(cost.cuda(), D, H, W)
  pred, conf = pred.cpu(), conf.cpu()
  # fp32 disparity within 1e-4 relative (north_star) -- measured against max(|d|, 1 px)
  assert ((pred - pred_o).abs() / pred_o.abs().clamp_min(1.0)).max().item() <= 1e-4
  same_r = torch.round(pred) == torch.round(pred_o)
  assert same_r.float().mean().item() > 0.999
  assert ((conf - conf_o).abs() * same_r).max().item() <= 2e-5  # three exp/sum terms, fp32
  assert conf.max().item() <= 2.0 + 1e-5  # border clamp can count a bin twice


@pytest.mark.parametrize('h4,w4,scale', [(8, 64, 4.0), (5, 128, 1.0), (16, 256, 12.0)])
def test_disp_regress_maxdisp192_row_tile_kernel(ops, h4, w4, scale):
  """maxdisp 192 with W % 256 == 0 (every BASELINE shape) runs the row-tile kernel: row-interpolated coarse tile in shared memory,
  logits in registers, confidence planes re-evaluated from the tile.  Same bounds as the generic kernel; `scale` 12 saturates most
  posteriors."""
  g = torch.Generator().manual_seed(h4 + w4)
  cost = torch.randn(2, 1, 48, h4, w4, generator=g) * scale
  D, H, W = 192, 4 * h4, 4 * w4
  pred_o, conf_o = O.disparity_regression(cost, D, H, W, want_conf=True)
  pred, conf = ops.disp_regress(cost.cuda(), D, H, W)
  pred, conf = pred.cpu(), conf.cpu()
  assert torch.isfinite(pred).all() and torch.isfinite(conf).all()
  assert ((pred - pred_o).abs() / pred_o.abs().clamp_min(1.0)).max().item() <= 1e-4
  same_r = torch.round(pred) == torch.round(pred_o)
  assert same_r.float().mean().item() > 0.999
  assert ((conf - conf_o).abs() * same_r).max().item() <= 2e-5
  assert conf.max().item() <= 2.0 + 1e-5
  assert torch.equal(ops.disp_regress(cost.cuda(), D, H, W)[0].cpu(), pred)  # deterministic


def test_disp_regress_one_hot_extremes(ops):
  """Saturated logits pin the disparity to {0, D-1} and the confidence to 2.0 at the border (clamp double count)."""
  cost = torch.full((1, 1, 4, 4, 4), -50.0)
  cost[0, 0, 0, :2] = 50.0
  cost[0, 0, 3, 2:] = 50.0
  pred, conf = ops.disp_regress(cost.cuda(), 16, 16, 16)
  pred_o, conf_o = O.disparity_regression(cost, 16, 16, 16, want_conf=True)
  assert torch.allclose(pred.cpu(), pred_o, atol=1e-4) and torch.allclose(conf.cpu(), conf_o, atol=1e-5)
  assert pred.min().item() < 1e-3 and pred.max().item() > 15 - 1e-3 and conf.max().item() > 1.99


# ---------------------------------------------------------------------------- a2 sphere conv
def _sphere_case(B, C, Co, h, w, st, seed):
  g = torch.Generator().manual_seed(seed)
  H, W = (h, w)
  x = torch.randn(B, C, H, W, generator=g)
  wgt = torch.randn(Co, C, 3, 3, generator=g) / math.sqrt(9 * C)
  pos = torch.from_numpy(O.gen_sphere_position(H, W, st))
  return x, wgt, pos


@pytest.mark.parametrize('B,C,Co,h,w,st', [(1, 1, 1, 5, 10, 'ERP'), (2, 8, 16, 16, 8, 'Cassini'), (1, 64, 128, 32, 16, 'Cassini'), (1, 128, 128, 16, 32, 'ERP'), (1, 5, 33, 16, 8, 'Cassini'),
                                                 (2, 16, 40, 10, 20, 'ERP'), (1, 8, 136, 6, 3, 'Cassini'), (1, 24, 200, 20, 10, 'Cassini')])
def test_sphere_conv_f32_vs_oracle(ops, B, C, Co, h, w, st):
  x, wgt, pos = _sphere_case(B, C, Co, h, w, st, 3)
  want = O.sphere_conv(x, pos, wgt)
  got = ops.sphere_conv_f32(x.cuda(), pos.cuda(), wgt.cuda(), None, None, None, False).cpu()
  assert (got - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())


def test_sphere_conv_f32_fused_epilogue(ops):
  x, wgt, pos = _sphere_case(1, 16, 32, 16, 8, 'Cassini', 4)
  scale, shift, res = torch.rand(32) + 0.5, torch.randn(32), torch.randn(1, 32, 16, 8)
  want = F.relu(O.sphere_conv(x, pos, wgt) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res)
  got = ops.sphere_conv_f32(x.cuda(), pos.cuda(), wgt.cuda(), scale.cuda(), shift.cuda(), res.cuda(), True).cpu()
  assert (got - want).abs().max().item() <= 2e-5


def test_sphere_conv_vs_compiled_reference_op(ops):
  """Ground truth = the UNMODIFIED reference CUDA op compiled into oracle/_ref (sphere_conv_cuda.cpp:129-210):
  pins both the oracle's restatement and the product kernel."""
  from oracle import build_ref
  ref = build_ref.load()
  if ref is None:
    pytest.skip('oracle/_ref/sphere_conv_cuda.so not built')
  for (B, C, Co, h, w, st) in [(2, 64, 128, 64, 32, 'Cassini'), (1, 128, 128, 32, 64, 'ERP')]:
    x, wgt, pos = _sphere_case(B, C, Co, h, w, st, 5)
    xc, wc, pc = x.cuda(), wgt.cuda(), pos.cuda()
    out = xc.new_empty((B, Co, h, w))
    ref.sphere_conv_forward_cuda(xc, wc, xc.new_empty(1), xc.new_empty(0), pc, out, xc.new_empty(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, False)
    torch.cuda.synchronize()
    tol = 2e-5 * max(1.0, out.abs().max().item())
    assert (O.sphere_conv(x, pos, wgt) - out.cpu()).abs().max().item() <= tol  # oracle restatement == reference op
    got = ops.sphere_conv_f32(xc, pc, wc, None, None, None, False)
    assert (got - out).abs().max().item() <= tol  # product kernel == reference op


# ---------------------------------------------------------------------------- a5 conv3d
@pytest.mark.parametrize('mode', [0, 1, 2])
@pytest.mark.parametrize('ci,co,dims', [(64, 32, (4, 8, 8)), (32, 64, (6, 10, 12)), (32, 1, (4, 8, 40)), (3, 5, (5, 7, 9))])
def test_conv3d_f32(ops, mode, ci, co, dims):
  g = torch.Generator().manual_seed(mode * 10 + ci)
  x = torch.randn(2, ci, *dims, generator=g)
  w = torch.randn((ci, co, 3, 3, 3) if mode == 2 else (co, ci, 3, 3, 3), generator=g) / math.sqrt(27 * ci)
  scale, shift = torch.rand(co, generator=g) + 0.5, torch.randn(co, generator=g)
  want = F.conv_transpose3d(x, w, None, 2, 1, 1) if mode == 2 else F.conv3d(x, w, None, mode + 1, 1)
  res = torch.randn(want.shape, generator=g)
  want = F.relu(want * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1) + res)
  got = ops.conv3d_f32(x.cuda(), w.cuda(), scale.cuda(), shift.cuda(), res.cuda(), mode, True).cpu()
  assert got.shape == want.shape
  assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())
  plain = ops.conv3d_f32(x.cuda(), w.cuda(), None, None, None, mode, False).cpu()
  want_plain = F.conv_transpose3d(x, w, None, 2, 1, 1) if mode == 2 else F.conv3d(x, w, None, mode + 1, 1)
  assert (plain - want_plain).abs().max().item() <= 2e-5 * max(1.0, want_plain.abs().max().item())


# ---------------------------------------------------------------------------- layout helpers
def test_layout_round_trip(ops):
  x = torch.randn(2, 37, 5, 9, device='cuda')
  for dt in (torch.bfloat16, torch.float16):
    y = ops.nchw_f32_to_nhwc_bf16(x, dt)
    assert y.shape == (2, 5, 9, 37) and y.dtype == dt and torch.equal(y, x.permute(0, 2, 3, 1).to(dt))
    assert torch.equal(ops.nhwc_bf16_to_nchw_f32(y), x.to(dt).float())


# ---------------------------------------------------------------------------- a8-a11 geometry
def test_disp2depth_all_pairs_vs_golden(ops, gold_dir):
  from mode_2022_b200.utils import geometry as G
  z = np.load(os.path.join(gold_dir, 'geometry_64x32.npz'))
  disp, conf = z['disp'], z['conf']
  for pair in G.CAM_PAIRS:
    d, c = G.disp2depth(torch.from_numpy(disp).cuda(), torch.from_numpy(conf).cuda(), pair)
    d, c = d.cpu().numpy().astype(np.float32), c.cpu().numpy()
    gd, gc = z[f'depth_{pair}'], z[f'conf_{pair}']
    if pair in ('12', '13', '14'):  # triangulation (+ bilinear resample): fp32 rounding of an fp64 expression
      assert np.abs(d - gd).max() <= 1e-6 * max(1.0, np.abs(gd).max()) * 8, pair
      assert np.abs(c - gc).max() <= 1e-6, pair
    else:  # forward warp: integer targets + z-buffer (the kernel follows numpy's FMA order of the rigid transform: identical output)
      bad = (np.abs(d - gd) > 1e-4 * np.maximum(1.0, np.abs(gd))) | (np.abs(c - gc) > 1e-6)
      print(f'disp2depth pair {pair}: {int(bad.sum())} of {bad.size} pixels differ from the reference golden')
      assert bad.sum() == 0 if Hh.numpy_matmul_is_fma_102() else bad.mean() <= 2e-3, (pair, bad.sum())  # exact: integer targets and z-buffer ties included


def test_depth_view_trans_bit_exact_on_identical_depth(ops):
  """Same fp64 depth in -> identical integer targets, depths and confidences out (ties included)."""
  from mode_2022_b200.utils import geometry as G
  rng = np.random.default_rng(3)
  conf = rng.random((64, 32), dtype=np.float32)
  exact_host = Hh.numpy_matmul_is_fma_102()
  for depth in (np.full((64, 32), 2.0), rng.random((64, 32)) * 30, np.where(rng.random((64, 32)) < 0.1, 0, rng.random((64, 32)) * 1000),
                np.where(rng.random((64, 32)) < 0.3, 99999.99999, rng.random((64, 32)) * 2e5)):  # the 100000 sentinel: fp64 below it, fp32 on it
    for dt in (np.float64, np.float32):
      d = depth.astype(dt)
      for pose in [(0, -1, 0, 0.5 * math.pi, 0, 0), (0, 1, 0, 0, 0, 0), (0, -math.sqrt(2) / 2, -math.sqrt(2) / 2, 0.75 * math.pi, 0, 0)]:
        v_o, c_o = O.depth_view_trans_with_conf(d, conf, *pose)
        v, c = G.depthViewTransWithConf(d, conf, *pose)
        mism = (v != v_o) | (c != c_o)
        # exact -- integer targets, depths, confidences, z-buffer ties included -- whenever this host's numpy rounds the rigid
        # transform the way the kernel does (fma(r2,c, fma(r0,a, r1*b)), the OpenBLAS order on this pool's AVX-512 hosts);
        # otherwise the reference's own tie-breaks differ from host to host and only the loose bound is meaningful
        print(f'depth_view_trans {dt.__name__} pose {pose[1]:+.2f},{pose[2]:+.2f}: {int(mism.sum())} mismatching pixels (exact host model: {exact_host})')
        assert mism.sum() == 0 if exact_host else mism.mean() <= 1e-3, (dt, pose, int(mism.sum()))


def test_rotate_and_c2e_numpy_api(ops):
  from mode_2022_b200.utils import geometry as G
  rng = np.random.default_rng(5)
  img = (rng.random((64, 32)) * 100).astype(np.float32)
  assert np.abs(G.rotateCassini(img, 0.5 * math.pi, 0, 0) - O.rotate_cassini(img, 0.5 * math.pi, 0, 0)).max() <= 1e-4
  assert np.abs(G.rotateCassini(img[:, :, None], 0.25 * math.pi, 0, 0)[:, :, 0] - O.rotate_cassini(img, 0.25 * math.pi, 0, 0)).max() <= 1e-4
  erp = G.cassini2Equirec(img)
  assert erp.shape == (32, 64) and np.abs(erp - O.cassini2equirec(img)).max() <= 1e-4
  t = torch.from_numpy(img).cuda()[None, None]
  assert torch.allclose(G.cassini2Equirec(t).cpu()[0], torch.from_numpy(O.cassini2equirec(img)), atol=1e-4)


def test_disp_to_depth_round_trip_full_size(ops):
  """Full-size property: depth -> disparity by the inverse sine rule recovers the input (1024x512)."""
  from mode_2022_b200.utils import geometry as G
  disp = torch.rand(1024, 512, device='cuda') * 190 + 1
  depth = G.disp_to_depth(disp, 1.0).double()
  w = 512
  phi_l = torch.from_numpy(np.arange(0.5 * math.pi - (0.5 * math.pi / w), -0.5 * math.pi, -(math.pi / w))).cuda().float().double()
  ok = (depth < 1000) & (depth > 0)
  # depth*sin(d) = cos(phi_l + d)  =>  tan(d) = cos(phi_l) / (depth + sin(phi_l)),  d = disp*pi/W
  d_rec = torch.atan2(torch.cos(phi_l).expand_as(depth), depth + torch.sin(phi_l)) * w / math.pi
  assert ok.float().mean().item() > 0.7  # phi_r beyond the pole gives a negative range, clipped to 0
  assert ((d_rec - disp.double()).abs()[ok]).max().item() < 2e-2


# ---------------------------------------------------------------------------- a5 conv3d on tensor cores (bf16)
def _tc_case(ops, mode, ci, co, dims, B, seed, relu=True, with_res=True, with_affine=True, dtype=torch.bfloat16):
  g = torch.Generator().manual_seed(seed)
  x = torch.randn(B, ci, *dims, generator=g).to(dtype)
  w = (torch.randn((ci, co, 3, 3, 3) if mode == 2 else (co, ci, 3, 3, 3), generator=g) / math.sqrt(27 * ci))
  wq = w.to(dtype).float()  # the kernel multiplies 16-bit-rounded weights
  scale = (torch.rand(co, generator=g) + 0.5) if with_affine else None
  shift = torch.randn(co, generator=g) if with_affine else None
  want = F.conv_transpose3d(x.float(), wq, None, 2, 1, 1) if mode == 2 else F.conv3d(x.float(), wq, None, mode + 1, 1)
  res = torch.randn(want.shape, generator=g).to(dtype) if with_res else None
  if with_affine:
    want = want * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
  if with_res:
    want = want + res.float()
  if relu:
    want = F.relu(want)
  wp = ops.conv3d_pack_weights(w.cuda(), mode, dtype)
  got = ops.conv3d_bf16(x.permute(0, 2, 3, 4, 1).contiguous().cuda(), wp, co, scale.cuda() if with_affine else None, shift.cuda() if with_affine else None,
                        res.permute(0, 2, 3, 4, 1).contiguous().cuda() if with_res else None, mode, relu, False)
  torch.cuda.synchronize()
  got = got.float().cpu().permute(0, 4, 1, 2, 3)
  assert got.shape == want.shape
  err = (got - want).abs()
  # output rounding (bf16: 2^-9 rel, fp16: 2^-12) + fp32 accumulation-order noise
  tol = (2.0**-7 if dtype == torch.bfloat16 else 2.0**-10) * want.abs().clamp_min(1.0)
  assert (err <= tol).all(), (mode, ci, co, dims, err.max().item(), (err / tol).max().item())


@pytest.mark.parametrize('mode', [0, 1, 2])
@pytest.mark.parametrize('ci,co', [(32, 32), (64, 32), (32, 64), (64, 64)])
def test_conv3d_bf16_tensor_core(ops, mode, ci, co):
  _tc_case(ops, mode, ci, co, (6, 16, 8), 1, seed=mode * 100 + ci + co)          # exactly one tile column
  _tc_case(ops, mode, ci, co, (5, 22, 13), 2, seed=mode * 100 + ci + co + 1)     # ragged tiles, odd dims, batch 2


@pytest.mark.parametrize('mode,ci,co', [(0, 32, 32), (0, 64, 32), (1, 32, 64), (2, 64, 32), (2, 64, 64)])
def test_conv3d_fp16_tensor_core(ops, mode, ci, co):
  _tc_case(ops, mode, ci, co, (5, 22, 13), 2, seed=300 + mode + ci + co, dtype=torch.float16)


def test_conv3d_bf16_plain_and_large(ops):
  _tc_case(ops, 0, 32, 32, (12, 48, 40), 1, seed=7, relu=False, with_res=False, with_affine=False)
  _tc_case(ops, 0, 64, 32, (16, 64, 32), 1, seed=8)
  _tc_case(ops, 1, 32, 64, (16, 64, 32), 1, seed=9)
  _tc_case(ops, 2, 64, 32, (8, 32, 16), 1, seed=10)


def test_conv3d_bf16_classifier_fp32_out(ops):
  g = torch.Generator().manual_seed(11)
  x = torch.randn(2, 32, 6, 20, 12, generator=g).bfloat16()
  w = torch.randn(1, 32, 3, 3, 3, generator=g) / math.sqrt(27 * 32)
  res = torch.randn(2, 1, 6, 20, 12, generator=g)
  want = F.conv3d(x.float(), w.bfloat16().float(), None, 1, 1) + res
  wp = ops.conv3d_pack_weights(w.cuda(), 0)
  got = ops.conv3d_bf16(x.permute(0, 2, 3, 4, 1).contiguous().cuda(), wp, 1, None, None, res.permute(0, 2, 3, 4, 1).contiguous().cuda(), 0, False, True)
  assert got.dtype == torch.float32 and got.shape == (2, 6, 20, 12, 1)
  assert (got.cpu().permute(0, 4, 1, 2, 3) - want).abs().max().item() <= 1e-4


# ---------------------------------------------------------------------------- a2 sphere conv on tensor cores (16-bit)
# Two references.  (1) The exact fp32 oracle on the 16-bit-rounded operands: the distance is the STORAGE FORMAT -- the A operand
# of the GEMM is the bilinear sample rounded to 16 bits (a random-walk error over K = 9*C products whose maximum over ~1e5..1e7
# outputs is what is bounded) plus the output rounding: 2^-6 (bf16) / 2^-8 (fp16) of max(|y|, 1); measured 1.0e-2 / 2.8e-3 at full
# size.  (2) The same oracle with its im2col columns rounded ONCE to the storage format ("ideal 16-bit im2col + GEMM"): the kernel
# is as close to that model (measured 7.9e-3 / 3.0e-3) as the model itself is to the exact result, i.e. it adds nothing to the
# format's own error: bf16 blends in fp32 (one rounding of A), fp16 blends in packed half (three extra 2^-12 roundings of A, still
# below one bf16 rounding).
SPHERE_TC_TOL = {torch.bfloat16: 2.0**-6, torch.float16: 2.0**-8}
SPHERE_TC_TOL_MODEL = SPHERE_TC_TOL


def _sphere_model_16(x, pos, wgt, dtype):
  """Oracle sphere conv whose column buffer is rounded once to `dtype` (fp32 accumulation): the ideal 16-bit im2col + GEMM."""
  B, C, h, w = x.shape
  cols = O.sphere_im2col(x.float(), pos).to(dtype).float().reshape(B, C * 9, h * w)
  return torch.matmul(wgt.to(dtype).float().reshape(wgt.shape[0], -1), cols).reshape(B, wgt.shape[0], h, w)


@pytest.mark.parametrize('B,C,Co,h,w,st', [(1, 64, 128, 16, 8, 'Cassini'), (2, 128, 128, 32, 16, 'Cassini'), (1, 128, 128, 16, 32, 'ERP'), (3, 64, 64, 8, 16, 'ERP'),
                                            (1, 128, 128, 40, 20, 'Cassini'), (2, 128, 128, 64, 32, 'Cassini'), (1, 64, 128, 128, 64, 'Cassini'), (2, 128, 128, 32, 64, 'ERP')])
@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16])
def test_sphere_conv_bf16_tensor_core(ops, B, C, Co, h, w, st, dtype):
  x, wgt, pos = _sphere_case(B, C, Co, h, w, st, 21)
  xq, wq = x.to(dtype), wgt.to(dtype)
  g = torch.Generator().manual_seed(1000 * B + Co + h)  # seeded: the bound below is statistical (max over ~1e5 outputs of a random-walk error)
  scale, shift = torch.rand(Co, generator=g) + 0.5, torch.randn(Co, generator=g)
  res = torch.randn(B, Co, h, w, generator=g).to(dtype)
  want = F.relu(O.sphere_conv(xq.float(), pos, wq.float()) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res.float())
  wp = ops.sphere_conv_pack_weights(wgt.cuda(), dtype)
  got = ops.sphere_conv_bf16(xq.permute(0, 2, 3, 1).contiguous().cuda(), pos.cuda(), wp, Co, scale.cuda(), shift.cuda(), res.permute(0, 2, 3, 1).contiguous().cuda(), True)
  torch.cuda.synchronize()
  got = got.float().cpu().permute(0, 3, 1, 2)
  err = (got - want).abs()
  rel = SPHERE_TC_TOL[dtype]
  tol = rel * want.abs().clamp_min(1.0)
  assert (err <= tol).all(), (err.max().item(), (err / tol).max().item())
  plain = ops.sphere_conv_bf16(xq.permute(0, 2, 3, 1).contiguous().cuda(), pos.cuda(), wp, Co, None, None, None, False).float().cpu().permute(0, 3, 1, 2)
  want_plain = O.sphere_conv(xq.float(), pos, wq.float())
  assert ((plain - want_plain).abs() <= rel * want_plain.abs().clamp_min(1.0)).all()
  model = _sphere_model_16(xq, pos, wgt, dtype)
  e_model = ((plain - model).abs() / model.abs().clamp_min(1.0)).max().item()
  assert e_model <= SPHERE_TC_TOL_MODEL[dtype], e_model


# ---------------------------------------------------------------------------- a1 stem conv (3 -> 32, 7x7, stride 2) on tensor cores
@pytest.mark.parametrize('B0,B1,h,w', [(1, 0, 16, 24), (2, 1, 64, 48), (1, 1, 50, 300), (1, 1, 256, 512), (1, 0, 33, 700)])
@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16])
def test_stem_conv_tensor_core(ops, B0, B1, h, w, dtype):
  """ops.stem_conv against F.conv2d (CPU fp32, the oracle's firstconv[0]) on 16-bit-rounded operands: the kernel multiplies
  exactly those values and accumulates in fp32, so only summation order and the output rounding differ."""
  g = torch.Generator().manual_seed(B0 * 100 + h)
  x0 = torch.randn(B0, 3, h, w, generator=g)
  x1 = torch.randn(B1, 3, h, w, generator=g) if B1 else None
  wt = torch.randn(32, 3, 7, 7, generator=g) / math.sqrt(147)
  sc, sh = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.1
  out = ops.stem_conv(x0.cuda(), None if x1 is None else x1.cuda(), wt.cuda(), sc.cuda(), sh.cuda(), True, dtype == torch.float16)
  x = x0 if x1 is None else torch.cat([x0, x1])
  ref = F.relu(F.conv2d(x.to(dtype).float(), wt.to(dtype).float(), None, 2, 3) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1))
  assert out.shape == (B0 + B1, (h - 1) // 2 + 1, (w - 1) // 2 + 1, 32) and out.dtype == dtype
  got = out.float().cpu().permute(0, 3, 1, 2)
  tol = 2.0 ** (-7 if dtype == torch.bfloat16 else -10)  # one output rounding (+ fp32 summation-order noise)
  err = (got - ref).abs() / (ref.abs() + 1.0)
  assert err.max().item() < tol, err.max().item()
  # no affine / no ReLU variant
  out2 = ops.stem_conv(x0.cuda(), None, wt.cuda(), None, None, False, dtype == torch.float16).float().cpu().permute(0, 3, 1, 2)
  ref2 = F.conv2d(x0.to(dtype).float(), wt.to(dtype).float(), None, 2, 3)
  assert ((out2 - ref2).abs() / (ref2.abs() + 1.0)).max().item() < tol


# ---------------------------------------------------------------------------- a3 sphere conv backward (fp32)
@pytest.mark.parametrize('B,C,Co,h,w,st,bias', [(1, 1, 1, 5, 10, 'ERP', False), (2, 8, 16, 16, 8, 'Cassini', True), (1, 64, 128, 32, 16, 'Cassini', False),
                                                (2, 128, 128, 16, 32, 'ERP', True), (1, 5, 33, 16, 8, 'Cassini', True), (1, 70, 40, 40, 20, 'Cassini', False)])
def test_sphere_conv_backward_vs_oracle_autograd(ops, B, C, Co, h, w, st, bias):
  """grad_input / grad_weight / grad_bias of the product op against torch autograd through the oracle's restatement
  (fp64 on CPU), through the reference-shaped autograd node (SphereConvFunction, reference sphere_conv.py:57-90)."""
  from mode_2022_b200.models.sphere_conv import sphere_conv
  x, wgt, pos = _sphere_case(B, C, Co, h, w, st, 11)
  g = torch.Generator().manual_seed(3)
  bs = torch.randn(Co, generator=g) if bias else None
  gout = torch.randn(B, Co, h, w, generator=g)
  xo, wo = x.double().requires_grad_(), wgt.double().requires_grad_()
  bo = bs.double().requires_grad_() if bias else None
  O.sphere_conv(xo, pos.double(), wo, bo).backward(gout.double())
  xc, wc = x.cuda().requires_grad_(), wgt.cuda().requires_grad_()
  bc = bs.cuda().requires_grad_() if bias else None
  out = sphere_conv(xc, pos.cuda(), wc, bc, 1, 1, 1, 1)
  assert out.requires_grad
  out.backward(gout.cuda())
  for name, got, ref in (('input', xc.grad, xo.grad), ('weight', wc.grad, wo.grad)) + ((('bias', bc.grad, bo.grad),) if bias else ()):
    tol = 3e-5 * max(1.0, ref.abs().max().item())
    err = (got.cpu().double() - ref).abs().max().item()
    assert err <= tol, (name, err, tol)
  # only the requested gradients are produced
  gi, gw, gb = ops.sphere_conv_backward_f32(x.cuda(), pos.cuda(), wgt.cuda(), gout.cuda(), False, True, False)
  assert gi is None and gb is None and (gw - wc.grad).abs().max().item() <= 3e-5 * max(1.0, wc.grad.abs().max().item())


def test_sphere_conv_backward_vs_compiled_reference_op(ops):
  """Ground truth = the UNMODIFIED reference CUDA op compiled into oracle/_ref (sphere_conv_backward_cuda, cpp:213-336)."""
  from oracle import build_ref
  ref = build_ref.load()
  if ref is None:
    pytest.skip('oracle/_ref/sphere_conv_cuda.so not built')
  for (B, C, Co, h, w, st) in [(2, 64, 128, 64, 32, 'Cassini'), (1, 128, 128, 32, 64, 'ERP')]:
    x, wgt, pos = _sphere_case(B, C, Co, h, w, st, 7)
    gout = torch.randn(B, Co, h, w, generator=torch.Generator().manual_seed(9))
    xc, wc, pc, gc = x.cuda(), wgt.cuda(), pos.cuda(), gout.cuda()
    bias = torch.zeros(Co, device='cuda')
    gi, gw, gb = torch.zeros_like(xc), torch.zeros_like(wc), torch.zeros_like(bias)
    ref.sphere_conv_backward_cuda(xc, wc, bias, xc.new_empty(0), pc, xc.new_empty(0), gi, gw, gb, gc, 3, 3, 1, 1, 1, 1, 1, 1, 1, True)
    torch.cuda.synchronize()
    pi, pw, pb = ops.sphere_conv_backward_f32(xc, pc, wc, gc, True, True, True)
    for name, got, want in (('input', pi, gi), ('weight', pw, gw), ('bias', pb, gb)):
      tol = 5e-5 * max(1.0, want.abs().max().item())
      assert (got - want).abs().max().item() <= tol, name


# ---------------------------------------------------------------------------- a5 classifier conv (32 -> 1) as pointwise GEMM + shifted sum
@pytest.mark.parametrize('B,d,h,w', [(1, 4, 16, 8), (2, 5, 20, 12), (1, 48, 32, 24), (1, 3, 40, 72), (2, 17, 16, 8), (1, 1, 16, 8), (3, 2, 24, 8)])
@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16])
def test_conv3d_classifier_tensor_core(ops, B, d, h, w, dtype):
  """ops.conv3d_classifier against F.conv3d (CPU fp32) on the 16-bit-rounded operands (+ fp32 residual): the kernel multiplies
  exactly those values with fp32 accumulation, so only the summation order differs."""
  g = torch.Generator().manual_seed(d * 100 + w)
  x = torch.randn(B, d, h, w, 32, generator=g)
  wt = torch.randn(1, 32, 3, 3, 3, generator=g) / math.sqrt(27 * 32)
  res = torch.randn(B, d, h, w, generator=g)
  xq, wq = x.to(dtype).float(), wt.to(dtype).float()
  want = F.conv3d(xq.permute(0, 4, 1, 2, 3), wq, None, 1, 1)[:, 0]
  got = ops.conv3d_classifier(x.to(dtype).cuda(), wt.cuda(), None).cpu()
  assert got.shape == (B, d, h, w) and got.dtype == torch.float32
  assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())
  got2 = ops.conv3d_classifier(x.to(dtype).cuda(), wt.cuda(), res.cuda()).cpu()
  assert (got2 - (want + res)).abs().max().item() <= 2e-5 * max(1.0, (want + res).abs().max().item())


# ---------------------------------------------------------------------------- a4+a5 cost volume fused into dres0[0]
@pytest.mark.parametrize('B,h,w,d4', [(1, 8, 16, 4), (2, 16, 24, 12), (1, 4, 8, 8), (1, 32, 128, 48), (2, 3, 40, 1), (1, 5, 6, 6)])
@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float16])
def test_costvol_conv_fused(ops, B, h, w, d4, dtype):
  """ops.costvol_conv (no volume materialised) against the literal definition: oracle cost volume + F.conv3d in fp64 on the
  16-bit-rounded operands (exact products, so only fp32 summation order and the output rounding differ), and against the
  product's own two-kernel path (cost_volume + conv3d_bf16)."""
  g = torch.Generator().manual_seed(h * 10 + d4)
  ref, tgt = torch.randn(B, 32, h, w, generator=g), torch.randn(B, 32, h, w, generator=g)
  wt = torch.randn(32, 64, 3, 3, 3, generator=g) / math.sqrt(27 * 64)
  sc, sh = torch.rand(32, generator=g) + 0.5, torch.randn(32, generator=g) * 0.2
  rq, tq, wq = ref.to(dtype).double(), tgt.to(dtype).double(), wt.to(dtype).double()
  cost = O.cost_volume(rq, tq, d4)
  want = F.relu(F.conv3d(cost, wq, None, 1, 1) * sc.double().view(1, -1, 1, 1, 1) + sh.double().view(1, -1, 1, 1, 1))  # (B,32,d4,h,w)
  nhwc = lambda t: t.to(dtype).permute(0, 2, 3, 1).contiguous().cuda()
  wr, wtm = ops.costvol_conv_weights(wt.cuda(), dtype)
  got = ops.costvol_conv(nhwc(ref), nhwc(tgt), wr, wtm, sc.cuda(), sh.cuda(), d4, True)
  assert got.shape == (B, d4, h, w, 32) and got.dtype == dtype
  gotc = got.float().cpu().permute(0, 4, 1, 2, 3).double()
  tol = 2.0 ** (-8 if dtype == torch.bfloat16 else -11)  # one output rounding
  assert ((gotc - want).abs() / (want.abs() + 1.0)).max().item() <= tol
  # the two-kernel product path gives the same values up to fp32 summation order (i.e. at most one 16-bit ulp apart)
  vol = ops.cost_volume(nhwc(ref), nhwc(tgt), d4)
  two = ops.conv3d_bf16(vol, ops.conv3d_pack_weights(wt.cuda(), 0, dtype), 32, sc.cuda(), sh.cuda(), None, 0, True, False)
  assert ((got.float() - two.float()).abs() / (two.float().abs() + 1.0)).max().item() <= 2 * tol


# ---------------------------------------------------------------------------- full-size code paths (BASELINE config[1] shapes), checked on the GPU
FULL_LAYERS = [(0, 64, 32, (48, 256, 128)), (0, 32, 32, (48, 256, 128)), (1, 32, 64, (48, 256, 128)), (0, 64, 64, (24, 128, 64)), (1, 64, 64, (24, 128, 64)),
               (0, 64, 64, (12, 64, 32)), (2, 64, 64, (12, 64, 32)), (2, 64, 32, (24, 128, 64))]


@pytest.mark.parametrize('mode,ci,co,dims', FULL_LAYERS)
@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16])
def test_conv3d_tc_full_size_vs_aten(ops, mode, ci, co, dims, dtype):
  """Every conv3d / deconv3d class of the 3-D stack at its real size (2 CTAs/SM, balanced depth chunks, 64-column blocks, 12x64x32
  tail grids -- paths the small cases never reach) against ATen's fp32 conv on the same 16-bit-rounded operands, on the GPU."""
  g = torch.Generator(device='cuda').manual_seed(mode * 1000 + ci + co + dims[0])
  x = torch.randn(1, ci, *dims, generator=g, device='cuda').to(dtype)
  w = torch.randn((ci, co, 3, 3, 3) if mode == 2 else (co, ci, 3, 3, 3), generator=g, device='cuda') / math.sqrt(27 * ci)
  scale, shift = torch.rand(co, generator=g, device='cuda') + 0.5, torch.randn(co, generator=g, device='cuda')
  wq = w.to(dtype).float()
  want = F.conv_transpose3d(x.float(), wq, None, 2, 1, 1) if mode == 2 else F.conv3d(x.float(), wq, None, mode + 1, 1)
  res = torch.randn(want.shape, generator=g, device='cuda').to(dtype)
  want = F.relu(want * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1) + res.float())
  got = ops.conv3d_bf16(x.permute(0, 2, 3, 4, 1).contiguous(), ops.conv3d_pack_weights(w, mode, dtype), co, scale, shift, res.permute(0, 2, 3, 4, 1).contiguous(), mode, True, False)
  got = got.float().permute(0, 4, 1, 2, 3)
  assert got.shape == want.shape
  tol = (2.0**-7 if dtype == torch.bfloat16 else 2.0**-10) * want.abs().clamp_min(1.0)
  err = (got - want).abs()
  assert (err <= tol).all(), (err.max().item(), (err / tol).max().item())


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16])
def test_conv3d_classifier_full_size_vs_aten(ops, dtype):
  g = torch.Generator(device='cuda').manual_seed(5)
  x = torch.randn(1, 48, 256, 128, 32, generator=g, device='cuda').to(dtype)
  wt = torch.randn(1, 32, 3, 3, 3, generator=g, device='cuda') / math.sqrt(27 * 32)
  res = torch.randn(1, 48, 256, 128, generator=g, device='cuda')
  want = F.conv3d(x.float().permute(0, 4, 1, 2, 3), wt.to(dtype).float(), None, 1, 1)[:, 0] + res
  got = ops.conv3d_classifier(x, wt, res)
  assert (got - want).abs().max().item() <= 2e-5 * max(1.0, want.abs().max().item())


@pytest.mark.parametrize('C,st,h,w', [(128, 'Cassini', 256, 128), (64, 'Cassini', 256, 128), (128, 'ERP', 128, 256)])
@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16])
def test_sphere_conv_tc_full_size_vs_reference_op(ops, C, st, h, w, dtype):
  """layer4's real shape (256x128 feature map: the +-128 px polar reach, the compile-time C=128 path, 2-D tiles + TMA epilogue)
  against the UNMODIFIED reference CUDA op (oracle/_ref) on the same 16-bit-rounded operands; falls back to the oracle on the GPU."""
  from oracle import build_ref
  B, Co = 3, 128
  g = torch.Generator(device='cuda').manual_seed(C + h)
  x = torch.randn(B, C, h, w, generator=g, device='cuda').to(dtype)
  wgt = torch.randn(Co, C, 3, 3, generator=g, device='cuda') / math.sqrt(9 * C)
  scale, shift = torch.rand(Co, generator=g, device='cuda') + 0.5, torch.randn(Co, generator=g, device='cuda')
  res = torch.randn(B, Co, h, w, generator=g, device='cuda').to(dtype)
  pos = torch.from_numpy(O.gen_sphere_position(h, w, st)).cuda()
  ref = build_ref.load()
  if ref is not None:
    conv = x.new_empty((B, Co, h, w), dtype=torch.float32)
    xf = x.float().contiguous()
    ref.sphere_conv_forward_cuda(xf, wgt.to(dtype).float().contiguous(), xf.new_empty(1), xf.new_empty(0), pos, conv, xf.new_empty(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, False)
  else:
    conv = O.sphere_conv(x.float(), pos, wgt.to(dtype).float())
  want = F.relu(conv * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res.float())
  got = ops.sphere_conv_bf16(x.permute(0, 2, 3, 1).contiguous(), pos, ops.sphere_conv_pack_weights(wgt, dtype), Co, scale, shift, res.permute(0, 2, 3, 1).contiguous(), True)
  got = got.float().permute(0, 3, 1, 2)
  rel = SPHERE_TC_TOL[dtype]
  err = (got - want).abs() / want.abs().clamp_min(1.0)
  model = F.relu(_sphere_model_16(x, pos, wgt, dtype) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) + res.float())
  e_model = ((got - model).abs() / model.abs().clamp_min(1.0)).max().item()
  print(f'sphere_conv_tc {C}->{Co} @{h}x{w} {st} {dtype}: max err vs exact {err.max().item():.2e} (tol {rel:.2e}), vs rounded-column model {e_model:.2e} '
        f'(tol {SPHERE_TC_TOL_MODEL[dtype]:.2e}), polar columns {err[..., :2].max().item():.2e}')
  assert err.max().item() <= rel and e_model <= SPHERE_TC_TOL_MODEL[dtype]


# ---------------------------------------------------------------------------- training: backward kernels registered with torch.library
@pytest.mark.parametrize('shape,d4', [((1, 32, 16, 8), 4), ((2, 8, 8, 16), 16), ((1, 4, 4, 12), 12), ((1, 32, 6, 8), 20)])
def test_cost_volume_backward_vs_autograd(ops, shape, d4):
  """Gradient of the fp32 cost volume (gather-sum over the shifts) against autograd through the oracle's slice assignments."""
  g = torch.Generator().manual_seed(d4)
  ref, tgt = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
  gout = torch.randn(shape[0], 2 * shape[1], d4, shape[2], shape[3], generator=g)
  ro, to = ref.double().requires_grad_(), tgt.double().requires_grad_()
  O.cost_volume(ro, to, d4).backward(gout.double())
  rc, tc = ref.cuda().requires_grad_(), tgt.cuda().requires_grad_()
  ops.cost_volume(rc, tc, d4).backward(gout.cuda())
  assert (rc.grad.cpu().double() - ro.grad).abs().max().item() <= 1e-5 * max(1.0, ro.grad.abs().max().item())
  assert (tc.grad.cpu().double() - to.grad).abs().max().item() <= 1e-5 * max(1.0, to.grad.abs().max().item())


@pytest.mark.parametrize('B,d4,h4,w4', [(2, 4, 16, 8), (1, 8, 8, 16), (1, 16, 12, 100), (2, 48, 8, 4), (1, 48, 4, 64), (2, 48, 7, 128)])
def test_disp_regress_backward_vs_autograd(ops, B, d4, h4, w4):
  """Fused soft-argmin head backward against autograd through F.interpolate + softmax + weighted sum (fp64 on CPU)."""
  g = torch.Generator().manual_seed(d4 + w4)
  cost = torch.randn(B, 1, d4, h4, w4, generator=g) * 3
  D, H, W = 4 * d4, 4 * h4, 4 * w4
  gp = torch.randn(B, 1, H, W, generator=g)
  co = cost.double().requires_grad_()
  O.disparity_regression(co, D, H, W).backward(gp.double())
  cc = cost.cuda().requires_grad_()
  pred, conf = ops.disp_regress(cc, D, H, W)
  pred.backward(gp.cuda())
  assert cc.grad.shape == cost.shape
  err = (cc.grad.cpu().double() - co.grad).abs().max().item()
  assert err <= 2e-5 * max(1.0, co.grad.abs().max().item()), err
  assert (pred.detach().cpu() - O.disparity_regression(cost, D, H, W)).abs().max().item() <= 1e-4 * D
  if d4 == 48 and W % 256 == 0:  # two-pass workspace path (gather transpose): bit-identical from run to run
    g2 = ops.disp_regress_backward(cost[:, 0].cuda().contiguous(), gp.cuda(), D)
    assert torch.equal(g2, cc.grad[:, 0]) and torch.equal(ops.disp_regress_backward(cost[:, 0].cuda().contiguous(), gp.cuda(), D), g2)


def test_registered_ops_pass_opcheck(ops):
  """torch.library.opcheck: schema, fake (meta) implementations and autograd registration of the differentiable ops."""
  from torch.library import opcheck
  g = torch.Generator().manual_seed(1)
  ref, tgt = torch.randn(1, 8, 8, 8, generator=g).cuda().requires_grad_(), torch.randn(1, 8, 8, 8, generator=g).cuda().requires_grad_()
  opcheck(torch.ops.mode_b200.cost_volume.default, (ref, tgt, 4))
  cost = (torch.randn(1, 1, 4, 8, 8, generator=g) * 2).cuda().requires_grad_()
  opcheck(torch.ops.mode_b200.disp_regress.default, (cost, 16, 32, 32), test_utils=('test_schema', 'test_faketensor', 'test_autograd_registration'))
  x, wgt, pos = _sphere_case(1, 8, 16, 16, 8, 'Cassini', 2)
  opcheck(torch.ops.mode_b200.sphere_conv_f32.default, (x.cuda().requires_grad_(), pos.cuda(), wgt.cuda().requires_grad_(), None, None, None, False),
          test_utils=('test_schema', 'test_faketensor', 'test_autograd_registration'))
  xb = torch.randn(1, 4, 8, 8, 32, generator=g).half().cuda()
  wp = ops.conv3d_pack_weights((torch.randn(32, 32, 3, 3, 3, generator=g) / 30).cuda(), 0, torch.float16)
  opcheck(torch.ops.mode_b200.conv3d_bf16.default, (xb, wp, 32, None, None, None, 0, True, False), test_utils=('test_schema', 'test_faketensor'))


def test_sphere_conv_backward_is_run_to_run_deterministic(ops):
  """The reference's col2im scatters with fp32 atomicAdd (sphere_conv_cuda_kernel.cu:341-352): grad_input differs in the last bits
  between runs.  The product accumulates in 64-bit fixed point with integer atomics: bit-identical gradients, every run."""
  x, wgt, pos = _sphere_case(2, 128, 128, 64, 32, 'Cassini', 13)
  gout = torch.randn(2, 128, 64, 32, generator=torch.Generator().manual_seed(4)) * 1e-4  # tiny gradients: the scale is derived from the data
  xc, wc, pc, gc = x.cuda(), wgt.cuda(), pos.cuda(), gout.cuda()
  runs = [ops.sphere_conv_backward_f32(xc, pc, wc, gc, True, True, True) for _ in range(3)]
  for r in runs[1:]:
    assert all(torch.equal(a, b) for a, b in zip(runs[0], r))
  # and it is the same gradient as the fp32-atomic path (up to that path's own summation-order noise)
  import os
  os.environ['MODE_B200_NONDETERMINISTIC_BWD'] = '1'
  try:
    ref = ops.sphere_conv_backward_f32(xc, pc, wc, gc, True, True, True)
  finally:
    del os.environ['MODE_B200_NONDETERMINISTIC_BWD']
  for a, b in zip(runs[0], ref):
    assert (a - b).abs().max().item() <= 2e-5 * max(b.abs().max().item(), 1e-12)


# ---------------------------------------------------------------------------- f1 training-mode BatchNorm
@pytest.mark.parametrize('cl', [False, True])
@pytest.mark.parametrize('shape', [(2, 32, 12, 64, 32), (1, 64, 6, 32, 16), (3, 5, 7, 9), (2, 128, 64, 32), (4, 3, 5, 7, 3), (1, 8, 2, 2)])
def test_batch_norm_train_vs_torch(ops, shape, cl):
  """mode_b200::batch_norm_train (+ backward) == F.batch_norm(training=True) of torch on the same GPU: output, saved statistics,
  running-stat update (momentum 0.1, unbiased variance), grad_input / grad_weight / grad_bias.  Shapes cover NCHW and NCDHW, vector
  (S % 4 == 0) and scalar paths, a non-zero mean much larger than the spread (shifted-sum statistics)."""
  from mode_2022_b200.models.batchnorm import BatchNorm2d, BatchNorm3d
  g = torch.Generator().manual_seed(sum(shape))
  C = shape[1]
  x = (torch.randn(*shape, generator=g) * 0.7 + 30.0 * torch.randn(1, C, *([1] * (len(shape) - 2)), generator=g)).cuda()
  gy = torch.randn(*shape, generator=g).cuda()
  if cl:  # channels_last / channels_last_3d tensors are normalised in place as (N, S, C) when C is a power of two >= 4 (else: contiguous copy)
    fmt = torch.channels_last_3d if len(shape) == 5 else torch.channels_last
    x, gy = x.contiguous(memory_format=fmt), gy.contiguous(memory_format=fmt)
  cls_t, cls_m = (torch.nn.BatchNorm3d, BatchNorm3d) if len(shape) == 5 else (torch.nn.BatchNorm2d, BatchNorm2d)
  ref, got = cls_t(C).cuda().train(), cls_m(C).cuda().train()
  with torch.no_grad():
    ref.weight.copy_(torch.rand(C, generator=g) + 0.5), ref.bias.copy_(torch.randn(C, generator=g))
  got.load_state_dict(ref.state_dict())
  assert list(got.state_dict().keys()) == list(ref.state_dict().keys())
  xr, xg = x.clone().requires_grad_(), x.clone().requires_grad_()
  for _ in range(2):  # two steps: the running statistics accumulate
    yr, yg = ref(xr), got(xg)
  yr.backward(gy), yg.backward(gy)
  if not cl or (C >= 4 and C & (C - 1) == 0):
    assert yg.stride() == xg.stride()  # the memory format is kept (channels-last needs a power-of-two C >= 4, else a contiguous copy is normalised)
  x64 = x.double()
  dims = [0] + list(range(2, x.dim()))
  truth = ((x64 - x64.mean(dims, keepdim=True)) / torch.sqrt(x64.var(dims, unbiased=False, keepdim=True) + 1e-5)) * ref.weight.double().view(1, C, *([1] * (x.dim() - 2))) \
      + ref.bias.double().view(1, C, *([1] * (x.dim() - 2)))
  err_ref, err_got = (yr.double() - truth).abs().max().item(), (yg.double() - truth).abs().max().item()
  assert err_got <= max(2 * err_ref, 2e-5), (err_got, err_ref)  # at least as close to the fp64 truth as cuDNN
  assert torch.allclose(got.running_mean, ref.running_mean, rtol=1e-5, atol=1e-5) and torch.allclose(got.running_var, ref.running_var, rtol=1e-4, atol=1e-6)
  assert int(got.num_batches_tracked) == int(ref.num_batches_tracked) == 2
  scale = max(1.0, xr.grad.abs().max().item())
  assert (xg.grad - xr.grad).abs().max().item() <= 5e-4 * scale  # dx is a difference of O(1) terms scaled by 1/sigma: conditioning of the op
  assert torch.allclose(got.weight.grad, ref.weight.grad, rtol=2e-4, atol=2e-3) and torch.allclose(got.bias.grad, ref.bias.grad, rtol=2e-4, atol=2e-3)
  # eval mode, momentum=None (cumulative average: the BN calibration of the fixtures) fall back to torch's implementation
  got.load_state_dict(ref.state_dict())
  got.eval(), ref.eval()
  assert torch.equal(got(x), ref(x))
  got.train(), ref.train()
  got.momentum = ref.momentum = None
  assert torch.equal(got(x), ref(x)) and torch.equal(got.running_var, ref.running_var)


def test_batch_norm_train_deterministic_and_opcheck(ops):
  from torch.library import opcheck
  x = torch.randn(2, 16, 4, 8, 8, device='cuda')
  w, b = torch.rand(16, device='cuda') + 0.5, torch.randn(16, device='cuda')
  a = ops.batch_norm_train(x, w, b, 1e-5)
  for _ in range(3):
    c = ops.batch_norm_train(x, w, b, 1e-5)
    assert all(torch.equal(u, v) for u, v in zip(a, c))
  opcheck(torch.ops.mode_b200.batch_norm_train.default, (x.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_(), 1e-5),
          test_utils=('test_schema', 'test_faketensor', 'test_autograd_registration'))
