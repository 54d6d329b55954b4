"""The per-kernel bar on the same B200 (VERDICT r01 task 2):

  A. the UNMODIFIED reference (baseline/_ref/ref: its Python + its own compiled CUDA op + cuDNN/cuBLAS) -- whole ModeDisparity
     forward at 1024x512, D=192, CUDA-event timed, with a torch.profiler kernel breakdown;
  B. cuDNN 16-bit channels_last_3d conv3d / conv_transpose3d for every layer class of the 3-D stack, next to mode_conv3d_tc;
  C. the spherical conv as im2col (torch gather) + cuBLAS 16-bit GEMM, and the GEMM alone (the floor of any un-fused design),
     next to mode_sphere_conv_tc.

  D. the unmodified reference's TRAINING step (its modules, its op, cuDNN) next to mode_2022_b200.training.train_step.

    python tools/ref_gpu_bench.py [A] [B] [C] [D]  -> gpurun_out/ref_gpu_bench.json + a table on stdout
"""
import json
import math
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mode_2022_b200 import ops  # noqa: E402
from oracle import mode_oracle as O  # noqa: E402
from oracle import stage_reference as SR  # noqa: E402
from tests import helpers as Hh  # noqa: E402

H, W, D = 1024, 512, 192
OUT = {}


def ev_time(fn, iters=5, warmup=2):
  for _ in range(warmup):
    fn()
  torch.cuda.synchronize()
  ts = []
  for _ in range(iters):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
  ts.sort()
  return ts[len(ts) // 2]


def bucket(name):
  n = name.lower()
  for key, b in (('sphere_im2col', 'sphere im2col (reference kernel)'), ('sphere_col2im', 'sphere col2im'), ('gemm', 'cuBLAS GEMM (addmm of the sphere conv)'),
                 ('cutlass', 'cuBLAS GEMM (addmm of the sphere conv)'), ('upsample', 'upsample_trilinear3d'), ('softmax', 'softmax'), ('grid_sampler', 'grid_sample (confidence)'),
                 ('batch_norm', 'batch norm'), ('bn_fw', 'batch norm'), ('conv', 'cuDNN conv (2-D + 3-D)'), ('xmma', 'cuDNN conv (2-D + 3-D)'), ('cudnn', 'cuDNN conv (2-D + 3-D)'),
                 ('implicit', 'cuDNN conv (2-D + 3-D)'), ('nchw', 'cuDNN layout transforms'), ('nhwc', 'cuDNN layout transforms'), ('memcpy', 'memcpy / memset'),
                 ('memset', 'memcpy / memset'), ('elementwise', 'elementwise / copy / reduce (ATen)'), ('reduce', 'elementwise / copy / reduce (ATen)'),
                 ('copy', 'elementwise / copy / reduce (ATen)'), ('fill', 'elementwise / copy / reduce (ATen)')):
    if key in n:
      return b
  return 'other'


def part_a():
  pkg = SR.reference_package()
  if pkg is None:
    OUT['A'] = {'unavailable': 'baseline/_ref/ref not staged'}
    return
  res = {}
  for tf32_conv, tf32_mm, tag in ((True, False, 'torch defaults (cuDNN TF32 on, matmul fp32)'), (False, False, 'TF32 off (parity setting)'), (True, True, 'TF32 everywhere')):
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32_conv, tf32_mm
    for B in (1, 6):
      m = pkg.ModeDisparity(D, conv='Sphere', in_height=H, in_width=W, sphereType='Cassini', out_conf=True)
      m.load_state_dict(O.synthetic_state_dict(Hh.KEY_SHAPES, seed=0))
      m = m.cuda().eval()
      g = torch.Generator().manual_seed(0)
      left, right = torch.randn(B, 3, H, W, generator=g).cuda(), torch.randn(B, 3, H, W, generator=g).cuda()
      with torch.no_grad():
        ms = ev_time(lambda: m(left, right), iters=4, warmup=2)
        # the same through host buffers: pinned upload + forward + download (what save_output_disparity_stage.py:93-102 does)
        lh, rh = left.cpu().pin_memory(), right.cpu().pin_memory()

        def e2e():
          p, c = m(lh.cuda(non_blocking=True), rh.cuda(non_blocking=True))
          return p.cpu(), c.cpu()
        ms_e2e = ev_time(e2e, iters=3, warmup=1)
      res[f'{tag} | B={B}'] = {'ms_per_forward': round(ms, 2), 'pairs_per_s': round(B / ms * 1e3, 2), 'e2e_pairs_per_s': round(B / ms_e2e * 1e3, 2),
                              'peak_mem_GB': round(torch.cuda.max_memory_allocated() / 2**30, 2)}
      print(f'[A] reference GPU forward, {tag}, B={B}: {ms:.1f} ms -> {B / ms * 1e3:.2f} pairs/s (host-to-host {B / ms_e2e * 1e3:.2f})', flush=True)
      if tf32_conv and not tf32_mm and B == 1:
        from torch.profiler import ProfilerActivity, profile
        with torch.no_grad(), profile(activities=[ProfilerActivity.CUDA]) as prof:
          m(left, right)
          torch.cuda.synchronize()
        by, top = {}, []
        for e in prof.key_averages():
          t = getattr(e, 'device_time_total', None)
          if t is None:
            t = getattr(e, 'cuda_time_total', 0.0)
          if t <= 0:
            continue
          by[bucket(e.key)] = by.get(bucket(e.key), 0.0) + t / 1e3
          top.append((t / 1e3, e.count, e.key[:110]))
        top.sort(reverse=True)
        res['profile (torch defaults, B=1)'] = {'buckets_ms': {k: round(v, 3) for k, v in sorted(by.items(), key=lambda kv: -kv[1])},
                                               'top_kernels': [{'ms': round(t, 3), 'calls': c, 'name': n} for t, c, n in top[:25]]}
        for k, v in sorted(by.items(), key=lambda kv: -kv[1]):
          print(f'    {v:9.3f} ms  {k}')
      del m
      torch.cuda.empty_cache()
      torch.cuda.reset_peak_memory_stats()
  OUT['A'] = res


LAYERS = [('64->32 s1 @48x256x128', 0, 64, 32, (48, 256, 128)), ('32->32 s1 @48x256x128', 0, 32, 32, (48, 256, 128)), ('32->64 s2 @48x256x128', 1, 32, 64, (48, 256, 128)),
          ('64->64 s1 @24x128x64', 0, 64, 64, (24, 128, 64)), ('64->64 s2 @24x128x64', 1, 64, 64, (24, 128, 64)), ('64->64 s1 @12x64x32', 0, 64, 64, (12, 64, 32)),
          ('64->64 deconv @12x64x32', 2, 64, 64, (12, 64, 32)), ('64->32 deconv @24x128x64', 2, 64, 32, (24, 128, 64)), ('32->1 s1 @48x256x128', 0, 32, 1, (48, 256, 128))]


def part_b(B=6):
  torch.backends.cudnn.benchmark = True
  res = {}
  for dtype in (torch.float16, torch.bfloat16):
    for name, mode, ci, co, dims in LAYERS:
      x = torch.randn(B, ci, *dims, device='cuda').to(dtype).contiguous(memory_format=torch.channels_last_3d)
      w = (torch.randn((ci, co, 3, 3, 3) if mode == 2 else (co, ci, 3, 3, 3), device='cuda') / math.sqrt(27 * ci))
      wl = w.to(dtype).contiguous(memory_format=torch.channels_last_3d)
      lib = (lambda: F.conv_transpose3d(x, wl, None, 2, 1, 1)) if mode == 2 else (lambda: F.conv3d(x, wl, None, mode + 1, 1))
      vox_out = B * math.prod(dims) * (8 if mode == 2 else 1) // (8 if mode == 1 else 1)
      flops = 2.0 * 27 * ci * co * (B * math.prod(dims) if mode == 2 else vox_out)  # deconv counted over input voxels
      with torch.no_grad():
        try:
          t_lib = ev_time(lib, iters=5, warmup=3)
        except Exception as e:  # pragma: no cover
          t_lib = float('nan')
          print('cuDNN failed:', e)
        xl = x.permute(0, 2, 3, 4, 1).contiguous()
        if co == 1:
          wc = w.float().contiguous()
          ours = lambda: ops.conv3d_classifier(xl, wc, None)
        else:
          wp = ops.conv3d_pack_weights(w.float().contiguous(), mode, dtype)
          ours = lambda: ops.conv3d_bf16(xl, wp, co, None, None, None, mode, False, False)
        t_ours = ev_time(ours, iters=5, warmup=3)
      key = f'{name} {str(dtype)[6:]}'
      res[key] = {'cudnn_ms': round(t_lib, 4), 'ours_ms': round(t_ours, 4), 'cudnn_TFLOPs': round(flops / t_lib / 1e9, 1), 'ours_TFLOPs': round(flops / t_ours / 1e9, 1),
                  'speedup': round(t_lib / t_ours, 2)}
      print(f'[B] {key:42s} cuDNN channels_last_3d {t_lib:8.3f} ms ({flops / t_lib / 1e9:7.1f} TF)   ours {t_ours:8.3f} ms ({flops / t_ours / 1e9:7.1f} TF)   x{t_lib / t_ours:.2f}', flush=True)
      del x, xl
  OUT['B'] = res


def part_c(B=12, C=128, Co=128, h=256, w=128):
  res = {}
  pos = torch.from_numpy(O.gen_sphere_position(h, w, 'Cassini')).cuda()
  flops = 2.0 * 9 * C * Co * B * h * w
  for dtype in (torch.float16, torch.bfloat16):
    x = torch.randn(B, C, h, w, device='cuda').to(dtype)
    wt = (torch.randn(Co, C, 3, 3, device='cuda') / math.sqrt(9 * C))
    w2 = wt.to(dtype).reshape(Co, C * 9)
    with torch.no_grad():
      cols = O.sphere_im2col(x[:1].float(), pos).to(dtype).reshape(1, C * 9, h * w)
      cols = cols.expand(B, -1, -1).contiguous()
      t_gemm = ev_time(lambda: torch.matmul(w2, cols), iters=5, warmup=3)
      try:
        t_i2c = ev_time(lambda: O.sphere_im2col(x.float(), pos).to(dtype), iters=3, warmup=1)
      except Exception as e:  # pragma: no cover
        t_i2c = float('nan')
        print('torch im2col failed:', e)
      xl = x.permute(0, 2, 3, 1).contiguous()
      wp = ops.sphere_conv_pack_weights(wt.float().contiguous(), dtype)
      t_ours = ev_time(lambda: ops.sphere_conv_bf16(xl, pos, wp, Co, None, None, None, False), iters=5, warmup=3)
    col_bytes = 2.0 * B * C * 9 * h * w
    key = str(dtype)[6:]
    res[key] = {'cublas_gemm_ms': round(t_gemm, 4), 'cublas_gemm_TFLOPs': round(flops / t_gemm / 1e9, 1), 'torch_im2col_ms': round(t_i2c, 3),
                'column_buffer_MB': round(col_bytes / 1e6, 1), 'ideal_im2col_write_plus_gemm_read_ms': round(2 * col_bytes / 6535.7e9 * 1e3, 4),
                'ours_fused_ms': round(t_ours, 4), 'ours_TFLOPs': round(flops / t_ours / 1e9, 1)}
    print(f'[C] sphere conv {C}->{Co} @{h}x{w} B={B} {key}: cuBLAS GEMM alone {t_gemm:.3f} ms ({flops / t_gemm / 1e9:.0f} TF), column buffer {col_bytes / 1e6:.0f} MB '
          f'(write+read at HBM peak {2 * col_bytes / 6535.7e9 * 1e3:.3f} ms), torch im2col {t_i2c:.2f} ms; ours fused {t_ours:.3f} ms ({flops / t_ours / 1e9:.0f} TF)', flush=True)
  # the reference op itself (fp32 im2col kernel + cuBLAS sgemm per batch element), same shape
  from oracle import build_ref
  ref = build_ref.load()
  if ref is not None:
    x = torch.randn(B, C, h, w, device='cuda')
    wt = torch.randn(Co, C, 3, 3, device='cuda') / math.sqrt(9 * C)
    out = x.new_empty((B, Co, h, w))
    for tf32 in (False, True):
      torch.backends.cuda.matmul.allow_tf32 = tf32
      t_ref = ev_time(lambda: ref.sphere_conv_forward_cuda(x, wt, x.new_empty(1), x.new_empty(0), pos, out, x.new_empty(0), 3, 3, 1, 1, 1, 1, 1, 1, 1, False), iters=5, warmup=2)
      res[f'reference op fp32 (matmul tf32={tf32})'] = {'ms': round(t_ref, 3), 'TFLOPs': round(flops / t_ref / 1e9, 1)}
      print(f'[C] reference sphere_conv_forward_cuda fp32 (tf32={tf32}) B={B}: {t_ref:.2f} ms ({flops / t_ref / 1e9:.1f} TF)', flush=True)
    torch.backends.cuda.matmul.allow_tf32 = False
  OUT['C'] = res


def part_d():
  """The UNMODIFIED reference's training step on the same GPU (train_disparity.py:147-163: three heads, smooth-L1 with weights
  0.5 / 0.7 / 1.0, Adam), 1 pair 1024x512 D=192, torch defaults (cuDNN TF32 allowed), next to mode_2022_b200.training.train_step."""
  pkg = SR.reference_package()
  if pkg is None:
    OUT['D'] = {'unavailable': 'baseline/_ref/ref not staged'}
    return
  from mode_2022_b200 import training as T
  from mode_2022_b200.models import ModeDisparity
  torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = True, False
  torch.backends.cudnn.benchmark = os.environ.get('CUDNN_BENCHMARK', '1') != '0'  # train_disparity.py:82 (benchmark = not args.cudnn_deter, default on)
  g = torch.Generator().manual_seed(1)
  res = {'cudnn_benchmark': torch.backends.cudnn.benchmark}
  for B in (1, 2):
    left, right = torch.randn(B, 3, H, W, generator=g).cuda(), torch.randn(B, 3, H, W, generator=g).cuda()
    disp = (torch.rand(B, 1, H, W, generator=g) * (D - 1)).cuda()
    mask = (torch.rand(B, 1, H, W, generator=g) < 0.9).cuda()
    sd = O.synthetic_state_dict(Hh.KEY_SHAPES, seed=0)
    ref = pkg.ModeDisparity(D, conv='Sphere', in_height=H, in_width=W, sphereType='Cassini')
    ref.load_state_dict(sd)
    ref = ref.cuda().train()
    opt_r = torch.optim.Adam(ref.parameters(), lr=1e-3, betas=(0.9, 0.999))

    def ref_step():
      opt_r.zero_grad()
      o1, o2, o3 = ref(left, right)
      loss = 0.5 * F.smooth_l1_loss(o1[mask], disp[mask], reduction='mean') + 0.7 * F.smooth_l1_loss(o2[mask], disp[mask], reduction='mean') \
          + F.smooth_l1_loss(o3[mask], disp[mask], reduction='mean')
      loss.backward()
      opt_r.step()

    try:
      torch.cuda.reset_peak_memory_stats()
      t_ref = ev_time(ref_step, iters=3, warmup=3)
      mem_ref = torch.cuda.max_memory_allocated() / 2**30
    except torch.cuda.OutOfMemoryError:
      t_ref, mem_ref = float('nan'), float('nan')
    del ref, opt_r
    torch.cuda.empty_cache()
    ours = ModeDisparity(D, conv='Sphere', in_height=H, in_width=W, sphereType='Cassini', precision='fp32')
    ours.load_state_dict(sd)
    ours = ours.cuda().train()
    red = T.GradAllReduce(ours.parameters())
    opt_o = torch.optim.Adam(ours.parameters(), lr=1e-3, betas=(0.9, 0.999))
    torch.cuda.reset_peak_memory_stats()
    t_ours = ev_time(lambda: T.train_step(ours, red, opt_o, left, right, disp, mask), iters=3, warmup=3)
    mem_ours = torch.cuda.max_memory_allocated() / 2**30
    del ours, opt_o, red
    torch.cuda.empty_cache()
    res[f'B={B}'] = {'reference_ms': round(t_ref, 1), 'reference_pairs_per_s': round(B / t_ref * 1e3, 2), 'reference_peak_GB': round(mem_ref, 1), 'ours_ms': round(t_ours, 1),
                     'ours_pairs_per_s': round(B / t_ours * 1e3, 2), 'ours_peak_GB': round(mem_ours, 1), 'speedup': round(t_ref / t_ours, 2)}
    print(f'[D] training step B={B}: reference {t_ref:.1f} ms ({mem_ref:.1f} GB), ours {t_ours:.1f} ms ({mem_ours:.1f} GB) -> {t_ref / t_ours:.2f}x', flush=True)
  OUT['D'] = res


def main():
  which = [a for a in sys.argv[1:] if a in ('A', 'B', 'C', 'D')] or ['A', 'B', 'C', 'D']
  t0 = time.time()
  for p in which:
    {'A': part_a, 'B': part_b, 'C': part_c, 'D': part_d}[p]()
  OUT['gpu'] = torch.cuda.get_device_name(0)
  OUT['seconds'] = round(time.time() - t0, 1)
  os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
  json.dump(OUT, open(os.path.join(ROOT, 'gpurun_out', 'ref_gpu_bench_%s.json' % ''.join(which)), 'w'), indent=1)


if __name__ == '__main__':
  main()
