"""Per-kernel timing at the C1 shapes (1024x512, D=192): CUDA events, warm-up, L2 flush between iterations.
Prints one JSON line per kernel with achieved GB/s or TFLOP/s and the fraction of the measured peak."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mode_2022_b200 import ops  # noqa: E402

PEAKS = {'hbm_gbs': 6535.7, 'bf16_tflops': 1600.2, 'bf16_tflops_sustained': 1381.5}
try:
  PEAKS.update(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json'))))
except Exception:
  pass

_flush = None


def flush_l2():
  global _flush
  if _flush is None:
    _flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device='cuda')
  _flush.zero_()


def timeit(fn, iters=10, warmup=3):
  for _ in range(warmup):
    fn()
  torch.cuda.synchronize()
  ts = []
  for _ in range(iters):
    flush_l2()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  ts.sort()
  return ts[len(ts) // 2]


def report(name, ms, bytes_=None, flops=None, **kw):
  r = dict(kernel=name, ms=round(ms, 4), **kw)
  if bytes_:
    r['GBps'] = round(bytes_ / ms / 1e6, 1)
    r['frac_hbm'] = round(r['GBps'] / PEAKS['hbm_gbs'], 3)
  if flops:
    r['TFLOPs'] = round(flops / ms / 1e9, 1)
    r['frac_bf16_sustained'] = round(r['TFLOPs'] / PEAKS['bf16_tflops_sustained'], 3)
  print(json.dumps(r), flush=True)
  return r


def main():
  which = sys.argv[1:] or ['cost', 'regress', 'conv3d', 'sphere', 'geometry']
  dev = 'cuda'
  D4, H4, W4 = 48, 256, 128
  if 'cost' in which:
    ref, tgt = torch.randn(1, 32, H4, W4, device=dev), torch.randn(1, 32, H4, W4, device=dev)
    ms = timeit(lambda: ops.cost_volume(ref, tgt, D4))
    report('cost_volume_f32', ms, bytes_=2 * ref.numel() * 4 + 64 * D4 * H4 * W4 * 4)
    rb, tb = ops.nchw_f32_to_nhwc_bf16(ref, torch.float16), ops.nchw_f32_to_nhwc_bf16(tgt, torch.float16)
    ms = timeit(lambda: ops.cost_volume(rb, tb, D4))
    report('cost_volume_bf16', ms, bytes_=2 * ref.numel() * 2 + 64 * D4 * H4 * W4 * 2)
  if 'regress' in which:
    cost = torch.randn(1, 1, D4, H4, W4, device=dev) * 3
    ms = timeit(lambda: ops.disp_regress(cost, 192, 1024, 512))
    report('disp_regress', ms, bytes_=cost.numel() * 4 + 2 * 1024 * 512 * 4, exp_per_s=round(192 * 1024 * 512 / ms * 1e3 / 1e9, 1))
  if 'conv3d' in which:
    layers = [('64->32 s1 @48x256x128', 0, 64, 32, (48, 256, 128)), ('32->32 s1 @48x256x128', 0, 32, 32, (48, 256, 128)), ('32->64 s2 @48x256x128', 1, 32, 64, (48, 256, 128)),
              ('64->64 s1 @24x128x64', 0, 64, 64, (24, 128, 64)), ('64->64 s2 @24x128x64', 1, 64, 64, (24, 128, 64)), ('64->64 s1 @12x64x32', 0, 64, 64, (12, 64, 32)),
              ('64->64 deconv @12x64x32', 2, 64, 64, (12, 64, 32)), ('64->32 deconv @24x128x64', 2, 64, 32, (24, 128, 64)), ('32->1 s1 @48x256x128', 0, 32, 1, (48, 256, 128))]
    only = os.environ.get('LAYER')
    for name, mode, ci, co, dims in layers:
      if only and only not in name:
        continue
      x = torch.randn(1, *dims, ci, device=dev).half()
      w = torch.randn((ci, co, 3, 3, 3) if mode == 2 else (co, ci, 3, 3, 3), device=dev) / math.sqrt(27 * ci)
      wp = ops.conv3d_pack_weights(w, mode, torch.float16)
      scale, shift = (torch.ones(co, device=dev), torch.zeros(co, device=dev)) if co > 1 else (None, None)
      f32 = co == 1
      ms = timeit(lambda: ops.conv3d_bf16(x, wp, co, scale, shift, None, mode, not f32, f32))
      vox_in = dims[0] * dims[1] * dims[2]
      vox_out = vox_in if mode == 0 else (vox_in // 8 if mode == 1 else vox_in)  # deconv counted over input voxels
      flops = 2 * 27 * ci * co * vox_out
      out_vox = vox_in if mode == 0 else (vox_in // 8 if mode == 1 else vox_in * 8)
      report('conv3d_bf16 ' + name, ms, flops=flops, bytes_=vox_in * ci * 2 + out_vox * co * (4 if f32 else 2))
  if 'sphere' in which:
    from mode_2022_b200.models.sphere_conv import sphere_position_numpy
    pos = torch.from_numpy(sphere_position_numpy(128, 256, 'Cassini')).to(dev)
    x = torch.randn(1, 128, 256, 128, device=dev)
    w = torch.randn(128, 128, 3, 3, device=dev) / 34
    ms = timeit(lambda: ops.sphere_conv_f32(x, pos, w, None, None, None, False), iters=5)
    report('sphere_conv_f32 128->128 @256x128', ms, flops=2 * 128 * 128 * 9 * 256 * 128)
    if hasattr(ops, 'sphere_conv_bf16'):
      xb = ops.nchw_f32_to_nhwc_bf16(x, torch.float16)
      wp = ops.sphere_conv_pack_weights(w, torch.float16)
      ms = timeit(lambda: ops.sphere_conv_bf16(xb, pos, wp, 128, None, None, None, False))
      report('sphere_conv_bf16 128->128 @256x128', ms, flops=2 * 128 * 128 * 9 * 256 * 128, bytes_=2 * x.numel() * 2 + pos.numel() * 4)

  if 'geometry' in which:
    # stage boundary of one frame: 6 pairs, disp -> depth (fp64 triangulation), rotations (grid sample) and the z-buffer warps
    from mode_2022_b200.utils.geometry import StageBoundary
    g = torch.Generator().manual_seed(0)
    disp = (torch.rand(6, 1, 1024, 512, generator=g) * 191).to(dev)
    conf = torch.rand(6, 1, 1024, 512, generator=g).to(dev)
    sb = StageBoundary()
    ms = timeit(lambda: sb(disp, conf), iters=5)
    report('stage_boundary 6 pairs @1024x512 (disp2depth + rotate/warp, all launches)', ms, bytes_=6 * 4 * 1024 * 512 * 4)


if __name__ == '__main__':
  main()
