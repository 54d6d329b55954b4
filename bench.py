#!/usr/bin/env python
"""bench.py -- stereo pairs/s of the MODE stereo stage (ModeDisparity forward) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): Deep360-shape stereo-stage inference, 6 camera pairs per step, Cassini
1024x512 (= "512x1024 equirect", SURVEY.md orientation note), maxdisp=192, 16-bit storage (fp16: the 16-bit format that meets
north_star's 0.01 px conv3d budget; bf16 measures 0.03 px), fp32 accumulation / logits / regression,
out_conf=True, synthetic inputs, seeded random-init weights.  One step = ModeDisparity.forward on one batch of 6 pairs; frames
and pairs are independent, so with N GPUs every rank processes its own frames (weak scaling, no data-path
collective).  The all-gather of per-pair maps into the fusion stage belongs to the two-stage layout: `--mode twostage`.

  value  pairs/s with inputs resident in HBM, CUDA-event timed, max over ranks
  e2e    same metric through the public module API with HOST (pinned) inputs and outputs copied back to the host
  roofline   tensor-core conv3d stack (dominant kernel family): algorithmic FLOPs / measured kernel time; `roofline.kernels` lists
             every kernel class of the step against its BINDING bound (tensor pipe, HBM, or the MUFU exp rate)
  reference_gpu  the UNMODIFIED reference (its Python + its own CUDA op + cuDNN/cuBLAS, staged in baseline/_ref/ref) on this GPU
  cpu_baseline  the CPU oracle (torch-CPU restatement of the reference; the reference has no CPU path of its own)
                on a bounded sample, on this box's host cores

--impl reference times the reference's algorithm on the host cores (oracle port), same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

H, W, MAXDISP, PAIRS = 1024, 512, 192, 6
# measured once with ncu (profiles/r01_launch_list_summary.md): 676.1 MB average DRAM traffic per conv3d launch x 27 launches per step;
# the dominant 32->32 layer moves 1.16 GB against 1.21 GB of algorithmic input + output bytes (profiles/r01_conv3d_tc_s1_ncu.md)
CONV3D_DRAM_BYTES_PER_STEP = 18.25e9
CONV3D_GFLOP_PER_PAIR = 1013.8  # 22 conv3d + 6 deconv3d at 1024x512, D=192
WORKLOAD = 'ModeDisparity stereo stage, 6 camera pairs/step, Cassini 1024x512 (=512x1024 ERP), maxdisp=192, out_conf'


def peaks():
  p = {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback (B200_PROFILING.md)'}
  try:
    m = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    p.update({k: m[k] for k in ('hbm_gbs', 'bf16_tflops', 'bf16_tflops_sustained') if k in m})
    p['source'] = 'MEASURED_PEAKS.json'
  except Exception:
    pass
  return p


class ClockSampler(threading.Thread):
  """Samples SM clock / throttle reasons through NVML while the timed region runs."""

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.samples, self.reasons, self.max_mhz, self._stop_evt = index, [], set(), None, threading.Event()
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
    except Exception:
      self.nv = None

  def run(self):
    if self.nv is None:
      return
    nv = self.nv
    names = {'hw_slowdown': nv.nvmlClocksThrottleReasonHwSlowdown, 'hw_thermal_slowdown': nv.nvmlClocksThrottleReasonHwThermalSlowdown,
             'sw_thermal_slowdown': nv.nvmlClocksThrottleReasonSwThermalSlowdown, 'sw_power_cap': nv.nvmlClocksThrottleReasonSwPowerCap}
    while not self._stop_evt.is_set():
      try:
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for k, bit in names.items():
          if r & bit:
            self.reasons.add(k)
      except Exception:
        pass
      time.sleep(0.05)

  def stop(self):
    self._stop_evt.set()
    self.join(timeout=2)
    s = sorted(self.samples)
    return {'sm_mhz': s[len(s) // 2] if s else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


PRECISION = os.environ.get('MODE_B200_BENCH_PRECISION', 'fp16')  # the benchmarked 16-bit storage format


def build_model(device, precision=None):
  precision = precision or PRECISION
  from mode_2022_b200.models import ModeDisparity
  torch.manual_seed(0)
  m = ModeDisparity(MAXDISP, conv='Sphere', in_height=H, in_width=W, sphereType='Cassini', out_conf=True, precision=precision)
  # seeded random init (reference constructor init) with mildly randomised BN statistics so nothing folds to identity
  g = torch.Generator().manual_seed(1)
  for mod in m.modules():
    if isinstance(mod, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
      mod.running_mean.copy_(torch.randn(mod.running_mean.shape, generator=g) * 0.05)
      mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 0.4 + 0.8)
      mod.weight.data.copy_(torch.rand(mod.weight.shape, generator=g) * 0.4 + 0.6)
  return m.to(device).eval()


MUFU_EXP_PER_S = 148 * 16 * 1.965e9  # SFU exp2 rate: 148 SMs x 16 lanes/clk x max SM clock; measured 15.9 lanes/clk/SM = 4.63 T/s (tools/mufu_bench.cu, profiles/r02_mufu_microbench.txt)


def _cval(a):
  return a.value if hasattr(a, 'value') else a


def roofline_report(prof, reps):
  """Per-kernel-class roofline from CUDA-event timings of the C-ABI calls.  For every class: algorithmic FLOPs and bytes (DESIGN.md
  section 3), the two time floors (FLOPs / sustained bf16-fp16 tensor peak, bytes / measured HBM copy bandwidth) and
  frac = binding floor / measured time.  The top-level object is the dominant kernel family (tcgen05 conv3d), as before."""
  pk = peaks()
  tf_peak, hbm_peak = pk['bf16_tflops_sustained'] * 1e12, pk['hbm_gbs'] * 1e9
  cls = {}

  def add(key, ms, flops=0.0, nbytes=0.0, exps=0.0):
    c = cls.setdefault(key, {'ms': 0.0, 'n': 0, 'flops': 0.0, 'bytes': 0.0, 'exps': 0.0})
    c['ms'] += ms
    c['n'] += 1
    c['flops'] += flops
    c['bytes'] += nbytes
    c['exps'] += exps

  for name, a, e0, e1 in prof:
    ms = e0.elapsed_time(e1)
    v = [_cval(x) for x in a]
    if name == 'mode_conv3d_tc':
      B, Ci, Co, D, Hh, Ww, mode = v[8:15]
      vox = B * D * Hh * Ww
      out_vox = vox if mode == 0 else (B * ((D - 1) // 2 + 1) * ((Hh - 1) // 2 + 1) * ((Ww - 1) // 2 + 1) if mode == 1 else vox * 8)
      flops = 2.0 * 27 * Ci * Co * (out_vox if mode != 2 else vox)  # transposed conv counted over input voxels
      has_res = (v[4] is not None) or (v[5] is not None)
      nbytes = vox * Ci * 2 + out_vox * Co * 2 * (2 if has_res else 1) + 27 * Ci * Co * 2
      tag = {0: 's1', 1: 's2', 2: 'deconv'}[mode]
      add(f'conv3d_tc {Ci}->{Co} {tag} @{D}x{Hh}x{Ww}', ms, flops, nbytes)
    elif name == 'mode_conv3d_classifier_tc':
      B, D, Hh, Ww = v[4:8]
      vox = B * D * Hh * Ww
      add(f'conv3d_cls_tc 32->1 @{D}x{Hh}x{Ww}', ms, 2.0 * 27 * 32 * vox, vox * 32 * 2 + vox * 4 * (2 if v[2] is not None else 1))
    elif name == 'mode_sphere_conv_tc':
      B, Cc, Hh, Ww, Co = v[7:12]
      px = B * Hh * Ww
      add(f'sphere_conv_tc {Cc}->{Co} @{Hh}x{Ww}', ms, 2.0 * 9 * Cc * Co * px, px * (Cc + Co * (2 if v[5] is not None else 1)) * 2 + 9 * Hh * Ww * 16 + 9 * Cc * Co * 2)
    elif name == 'mode_costvol_conv_fused':
      B, D4, Hh, Ww = v[5:9]
      add('costvol_conv (cost volume + dres0[0] fused; K=96 GEMMs excluded)', ms, 0.0, 2 * B * Hh * Ww * 288 * 4 + B * D4 * Hh * Ww * 32 * 2)
    elif name == 'mode_costvol_cols':
      B, Hh, Ww = v[2:5]
      add('costvol_cols', ms, 0.0, B * Hh * Ww * (32 + 96) * 2)
    elif name == 'mode_disp_regress':
      B, D4, H4, W4, D, Hh, Ww = v[3:10]
      add('disp_regress (upsample + softmax + soft-argmin + confidence)', ms, 0.0, B * (D4 * H4 * W4 + 2 * Hh * Ww) * 4, exps=float(B) * D * Hh * Ww)
    elif name == 'mode_stem_conv_tc':
      B0, B1, Hh, Ww = v[6:10]
      px_o = (B0 + B1) * ((Hh - 1) // 2 + 1) * ((Ww - 1) // 2 + 1)
      add('stem_conv_tc 3->32 7x7 s2', ms, 2.0 * 147 * 32 * px_o, (B0 + B1) * 3 * Hh * Ww * 4 + px_o * 32 * 2)
    elif name == 'mode_concat3_nhwc_16':
      npix, ca, cb, cc = v[4:8]
      add('concat3_nhwc', ms, 0.0, 2 * npix * (ca + cb + cc) * 2)
    else:
      add(name, ms)
  kernels = []
  for key, c in cls.items():
    t = c['ms'] / reps * 1e-3
    fl, by, ex = c['flops'] / reps, c['bytes'] / reps, c['exps'] / reps
    floors = {'tensor': fl / tf_peak, 'hbm': by / hbm_peak}
    if ex:
      floors['mufu_exp'] = ex / MUFU_EXP_PER_S
    bound = max(floors, key=floors.get)
    ent = {'kernel': key, 'launches_per_step': c['n'] // reps, 'ms_per_step': round(t * 1e3, 4), 'bound': bound, 'frac': round(floors[bound] / t, 3) if t > 0 else None}
    if fl:
      ent['TFLOPs'] = round(fl / t / 1e12, 1)
      ent['frac_tensor'] = round(floors['tensor'] / t, 3)
    if by:
      ent['GBps'] = round(by / t / 1e9, 1)
      ent['frac_hbm'] = round(floors['hbm'] / t, 3)
    if ex:
      ent['Texp_per_s'] = round(ex / t / 1e12, 3)
    kernels.append(ent)
  kernels.sort(key=lambda e: -e['ms_per_step'])
  conv = [c for k, c in cls.items() if k.startswith('conv3d_')]
  t_conv = sum(c['ms'] for c in conv) / reps
  fl_conv = sum(c['flops'] for c in conv) / reps
  ach = fl_conv / (t_conv * 1e-3) / 1e12 if t_conv > 0 else 0.0
  n_conv = sum(c['n'] for c in conv) // reps
  fused_first = any(k.startswith('costvol_conv') for k in cls)
  return {'bound': 'tensor',
          'kernel': f'conv3d_tc_kernel + conv3d_cls_tc_kernel ({n_conv} conv3d/deconv3d/classifier launches per step' + ('; the first layer of the 3-D stack is fused with the cost volume)' if fused_first else ')'),
          'achieved': round(ach, 1), 'peak': pk['bf16_tflops_sustained'], 'unit': 'TFLOP/s', 'frac': round(ach / pk['bf16_tflops_sustained'], 3),
          'traffic': CONV3D_DRAM_BYTES_PER_STEP,
          'traffic_note': 'DRAM read+write bytes of the conv3d launches of one step (6 pairs), ncu launch list under profiles/',
          'peak_source': pk['source'] + ' (sustained 16-bit dense; burst %.0f; HBM %.1f GB/s; MUFU exp %.2f T/s nominal, 4.63 measured)' % (pk['bf16_tflops'], pk['hbm_gbs'], MUFU_EXP_PER_S / 1e12),
          'launches_per_step': n_conv, 'ms_per_step_in_kernel': round(t_conv, 3),
          'kernels': kernels}


def reference_gpu_block(dev):
  """The UNMODIFIED reference on this GPU ('that, not the CPU, is the number to beat'): its own Python, its own
  compiled CUDA op and cuDNN/cuBLAS with torch's default settings, from baseline/_ref/ref (oracle/stage_reference.py).  A bounded
  sample: 1 pair per forward, device-resident inputs (the reference's host-side cost-volume upload is inside its forward)."""
  try:
    from oracle import mode_oracle as O
    from oracle import stage_reference as SR
    from tests.helpers import KEY_SHAPES
    pkg = SR.reference_package()
    if pkg is None:
      return {'unavailable': 'baseline/_ref/ref is not staged on this box (python oracle/stage_reference.py)'}
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = True, False  # torch defaults = what a reference user gets
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):  # the reference's constructors print(); stdout carries exactly one JSON line
      m = pkg.ModeDisparity(MAXDISP, conv='Sphere', in_height=H, in_width=W, sphereType='Cassini', out_conf=True)
    m.load_state_dict(O.synthetic_state_dict(KEY_SHAPES, seed=0))
    m = m.to(dev).eval()
    g = torch.Generator().manual_seed(0)
    l, r = torch.randn(1, 3, H, W, generator=g).to(dev), torch.randn(1, 3, H, W, generator=g).to(dev)
    with torch.no_grad():
      for _ in range(2):
        m(l, r)
      torch.cuda.synchronize()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      n = 4
      a.record()
      for _ in range(n):
        m(l, r)
      b.record()
      torch.cuda.synchronize()
    ms = a.elapsed_time(b) / n
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32
    del m
    torch.cuda.empty_cache()
    return {'value': round(1e3 / ms, 2), 'unit': 'pairs/s', 'ms_per_pair': round(ms, 2), 'n_gpus': 1,
            'sample': f'{n} forwards of 1 pair 1024x512 D=192, fp32 NCHW, cuDNN TF32 allowed (torch default), matmul fp32 (torch default), eval, out_conf',
            'impl': 'unmodified reference ModeDisparity + its compiled sphere_conv_cuda extension (baseline/_ref/ref)'}
  except Exception as e:  # the checker must never take the bench down
    return {'unavailable': f'{type(e).__name__}: {e}'[:300]}


def run_ours(args):
  import torch.distributed as dist
  from mode_2022_b200 import _lib
  rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
  local = int(os.environ.get('LOCAL_RANK', 0))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  model = build_model(dev)
  g = torch.Generator().manual_seed(100 + rank)
  left_h = torch.randn(PAIRS, 3, H, W, generator=g).pin_memory()
  right_h = torch.randn(PAIRS, 3, H, W, generator=g).pin_memory()
  left, right = left_h.to(dev), right_h.to(dev)

  def exchange(pred, conf):
    """Default mode: frames are independent and every rank owns whole frames, so the hot path has no exchange step (the
    all-gather into the fusion stage is measured by --mode twostage, where the 6 pairs of ONE frame are sharded)."""

  # ---- device-resident step, captured in a CUDA graph (static shapes; ~250 launches per step)
  with torch.no_grad():
    for _ in range(2):
      pred, conf = model(left, right)
    torch.cuda.synchronize()
    graph = None
    if not args.no_graph:
      try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
          model(left, right)
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
          g_pred, g_conf = model(left, right)
      except Exception as e:  # pragma: no cover - report and run eagerly
        print(f'[bench] CUDA graph capture failed ({e}); running eagerly', file=sys.stderr)
        graph = None

    def step():
      if graph is not None:
        graph.replay()
        exchange(g_pred, g_conf)
        return g_pred, g_conf
      p_, c_ = model(left, right)
      exchange(p_, c_)
      return p_, c_

    def timed(fn, steps):
      if world > 1:
        dist.barrier()
      torch.cuda.synchronize()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record()
      for _ in range(steps):
        fn()
      b.record()
      torch.cuda.synchronize()
      ms = torch.tensor([a.elapsed_time(b)], device=dev)
      if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
      return ms.item()

    for _ in range(max(args.warmup, 3)):
      step()
    launches0 = _lib.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(step, args.steps)
    clocks = sampler.stop()
    # launches of OUR kernels per step: counted on an eager step (graph replays do not pass through the C ABI)
    l0 = _lib.launch_count()
    model(left, right)
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - l0

    # ---- end to end through the public API with host buffers: mode_2022_b200.pipeline.HostPipeline streams frames host -> device
    # -> host; every step uploads its own 6 pairs from pinned memory and downloads its own disparity + confidence maps inside
    # the timed region, the upload of step i+1 / download of step i-1 overlapping the compute of step i (2 frames in flight).
    from mode_2022_b200.pipeline import HostPipeline
    pipe = HostPipeline(model, PAIRS, H, W, depth=2, use_graph=not args.no_graph, post=exchange)

    def e2e_run(steps):
      if world > 1:
        dist.barrier()
      torch.cuda.synchronize()
      a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      a.record(pipe.s_in)
      tickets = []
      for i in range(steps):
        tickets.append(pipe.submit(left_h, right_h))
        if i >= 1:
          pipe.collect(tickets[i - 1])  # the host consumes frame i-1 while frame i computes
      pipe.collect(tickets[-1])
      b.record(pipe.s_out)
      torch.cuda.synchronize()
      t = torch.tensor([a.elapsed_time(b)], device=dev)
      if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
      return t.item()

    e2e_run(3)
    ms_e2e = e2e_run(args.steps)
    pred_h, conf_h = pipe.h_pred[0], pipe.h_conf[0]

    # ---- roofline: CUDA events around every C-ABI call of an eager pass (the graph replays the same kernels)
    roof, ref_gpu = None, None
    if rank == 0:
      _lib.PROFILE = []
      for _ in range(3):
        model(left, right)
      torch.cuda.synchronize()
      prof, _lib.PROFILE = _lib.PROFILE, None
      roof = roofline_report(prof, 3)
      if world == 1 and not os.environ.get('MODE_B200_BENCH_LIGHT'):  # (LIGHT: profiling runs under ncu skip the checker-side legs)
        ref_gpu = reference_gpu_block(dev)

  pairs = PAIRS * world * args.steps
  out = {
      'metric': 'stereo pairs/s @512x1024 D=192', 'value': round(pairs / (ms * 1e-3), 2), 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps,
      'warmup': max(args.warmup, 3), 'ms_per_step': round(ms / args.steps, 3), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
      'dtype': PRECISION, 'data': 'synthetic (randn images, seeded random-init weights)',
      'config': {'workload': WORKLOAD, 'pairs_per_step_per_gpu': PAIRS, 'parallelism': f'frames (6 pairs each) sharded over {world} GPU(s), weights replicated, no data-path collective',
                 'l2': 'per-step working set (>2 GB of activations) exceeds the 126 MB L2', 'cuda_graph': graph is not None},
      'e2e': {'value': round(pairs / (ms_e2e * 1e-3), 2), 'unit': 'pairs/s', 'h2d_bytes_per_step': int(left_h.numel() * 4 * 2), 'd2h_bytes_per_step': int(pred_h.numel() * 4 * 2),
              'ms_per_step': round(ms_e2e / args.steps, 3),
              'api': 'mode_2022_b200.pipeline.HostPipeline (2 frames in flight: H2D / compute / D2H on separate streams)'},
      'gpu_launches': int(launches_per_step * args.steps), 'gpu_launches_per_step': int(launches_per_step),
      'clocks': clocks,
  }
  if rank == 0:
    out['roofline'] = roof
    if ref_gpu is not None:
      out['reference_gpu'] = ref_gpu
    out['cpu_baseline'] = cpu_baseline(sample_only=True) if not os.environ.get('MODE_B200_BENCH_LIGHT') else None
    print(json.dumps(out), flush=True)
  if world > 1:
    dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# CPU legs: the oracle (test infrastructure) is the only CPU implementation; it is used here as the checker-side
# baseline, never on the product path.
# ------------------------------------------------------------------------------------------------


def _oracle_forward_seconds(h, w, d, n_threads):
  from oracle import mode_oracle as O
  from tests.helpers import KEY_SHAPES
  torch.set_num_threads(n_threads)
  sd = O.synthetic_state_dict(KEY_SHAPES, seed=0)
  g = torch.Generator().manual_seed(0)
  left, right = torch.randn(1, 3, h, w, generator=g), torch.randn(1, 3, h, w, generator=g)
  t0 = time.time()
  O.mode_disparity_forward(sd, left, right, d, 'Cassini', out_conf=True)
  return time.time() - t0


def cpu_baseline(sample_only=False):
  cores = os.cpu_count() or 1
  t = _oracle_forward_seconds(H, W, MAXDISP, cores)
  return {'value': round(1.0 / t, 4), 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
          'sample': f'1 pair 1024x512 D=192 fp32, oracle (torch-CPU restatement of the reference; the reference has no CPU path), {cores} threads, {t:.1f} s'}


def run_reference(args):
  rank = int(os.environ.get('RANK', 0))
  if rank != 0:
    return
  cores = os.cpu_count() or 1
  budget_s = 150.0
  t_first = _oracle_forward_seconds(H, W, MAXDISP, cores)  # warm-up (also sizes the run)
  steps = max(1, min(args.steps, int(budget_s / max(t_first, 1e-3))))
  t0 = time.time()
  for _ in range(steps):
    _oracle_forward_seconds(H, W, MAXDISP, cores)
  dt = time.time() - t0
  v = round(steps / dt, 4)
  print(json.dumps({
      'impl': 'reference', 'metric': 'stereo pairs/s @512x1024 D=192', 'value': v, 'unit': 'pairs/s', 'n_gpus': int(os.environ.get('WORLD_SIZE', 1)), 'steps': steps,
      'warmup': 1, 'ms_per_step': round(dt / steps * 1e3, 1), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
      'data': 'synthetic (randn images, seeded random-init weights)',
      'config': {'workload': WORKLOAD, 'sample': '1 pair per step (1/6 of a frame) on the host cores; steps capped to a ~150 s budget'},
      'cpu_baseline': {'value': v, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
                       'sample': f'{steps} x 1 pair 1024x512 D=192 fp32 through the oracle port (the reference has no CPU path: hard .cuda() calls, SURVEY.md §8c)'},
      'e2e': {'value': v, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
  }), flush=True)


# ------------------------------------------------------------------------------------------------
# Other BASELINE.json configs (not the driver's default line): --mode train | twostage | highres
# ------------------------------------------------------------------------------------------------


def _dist_setup():
  import torch.distributed as dist
  rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
  local = int(os.environ.get('LOCAL_RANK', 0))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  return dist, rank, world, local, dev


def _timed(dist, world, dev, fn, steps):
  if world > 1:
    dist.barrier()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(steps):
    fn()
  b.record()
  torch.cuda.synchronize()
  ms = torch.tensor([a.elapsed_time(b)], device=dev)
  if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  return ms.item()


def run_train(args):
  """BASELINE config[2]: stereo-stage TRAINING step, global batch of 8 pairs 1024x512 D=192, camera-pair data-parallel over N GPUs with
  the bucketed NCCL gradient all-reduce (mode_2022_b200/training.py) overlapped with the backward pass; fp32 like the reference
  (cuDNN TF32 allowed = torch default).  STRONG scaling: the global batch is fixed, every rank takes 8/N pairs."""
  dist, rank, world, local, dev = _dist_setup()
  from mode_2022_b200 import _lib
  from mode_2022_b200 import training as T
  from mode_2022_b200.models import ModeDisparity
  gb = args.pairs or 8
  if gb % world:
    raise SystemExit(f'--mode train: global batch {gb} is not divisible by {world} ranks')
  nb = gb // world
  torch.backends.cudnn.benchmark = True  # as the reference's training script (train_disparity.py:82: benchmark = not args.cudnn_deter)
  torch.manual_seed(0)
  model = ModeDisparity(MAXDISP, conv='Sphere', in_height=H, in_width=W, sphereType='Cassini', precision='fp32').to(dev).train()
  reducer = T.GradAllReduce(model.parameters())
  opt = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999))  # reference train_disparity.py:293
  g = torch.Generator().manual_seed(200 + rank)
  left_h, right_h = torch.randn(nb, 3, H, W, generator=g).pin_memory(), torch.randn(nb, 3, H, W, generator=g).pin_memory()
  disp_h = (torch.rand(nb, 1, H, W, generator=g) * (MAXDISP - 1)).pin_memory()
  mask_h = (torch.rand(nb, 1, H, W, generator=g) < 0.9).pin_memory()
  left, right, disp, mask = left_h.to(dev), right_h.to(dev), disp_h.to(dev), mask_h.to(dev)
  loss_box = [None]

  def step():
    loss_box[0] = T.train_step(model, reducer, opt, left, right, disp, mask)

  def step_e2e():
    l, r = left_h.to(dev, non_blocking=True), right_h.to(dev, non_blocking=True)
    d, m = disp_h.to(dev, non_blocking=True), mask_h.to(dev, non_blocking=True)
    loss_box[0] = T.train_step(model, reducer, opt, l, r, d, m).cpu()

  for _ in range(max(args.warmup, 3)):
    step()
  l0 = _lib.launch_count()
  sampler = ClockSampler(local)
  sampler.start()
  ms = _timed(dist, world, dev, step, args.steps)
  clocks = sampler.stop()
  launches = _lib.launch_count() - l0
  ms_e2e = _timed(dist, world, dev, step_e2e, args.steps)
  # how much of the all-reduce is exposed: time the same step with the collectives skipped (world == 1 semantics) is not available in
  # a multi-rank run, so report the wait inside finish() instead
  torch.cuda.synchronize()
  t0 = time.time()
  reducer.zero_grad()
  outs = model(left, right)
  T.global_masked_loss(outs, disp, mask).backward()
  torch.cuda.synchronize()
  t1 = time.time()
  reducer.finish()
  torch.cuda.synchronize()
  t2 = time.time()
  if rank == 0:
    print(json.dumps({
        'metric': 'stereo training pairs/s @512x1024 D=192 (global batch 8)', 'value': round(gb * args.steps / (ms * 1e-3), 3), 'unit': 'pairs/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': round(ms / args.steps, 2), 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32 (cuDNN TF32 allowed, as the reference trains)', 'data': 'synthetic (randn images, uniform disparities, 90 % mask, seeded random-init weights)',
        'config': {'workload': f'ModeDisparity training step (3 heads, smooth-L1, Adam), global batch {gb} pairs, Cassini 1024x512, maxdisp=192', 'pairs_per_gpu': nb,
                   'parallelism': f'camera pairs sharded over {world} GPU(s); bucketed NCCL gradient all-reduce ({reducer.grad_bytes() / 1e6:.1f} MB in {len(reducer.buckets)} buckets) '
                                  'launched from autograd hooks, overlapped with the backward pass',
                   'custom_kernels': 'sphere conv fwd+bwd, cost volume fwd+bwd, soft-argmin heads fwd+bwd, BatchNorm2d/3d fwd+bwd (libmode_b200); conv2d/conv3d: cuDNN TF32 on channels-last tensors'},
        'e2e': {'value': round(gb * args.steps / (ms_e2e * 1e-3), 3), 'unit': 'pairs/s', 'h2d_bytes_per_step': int(nb * (2 * 3 + 1) * H * W * 4 + nb * H * W),
                'd2h_bytes_per_step': 4, 'ms_per_step': round(ms_e2e / args.steps, 2)},
        'gpu_launches': int(launches), 'clocks': clocks, 'loss': float(loss_box[0]),
        'allreduce': {'bytes': reducer.grad_bytes(), 'buckets': len(reducer.buckets), 'launch_order': reducer.launch_order[:16],
                      'fwd_bwd_ms': round((t1 - t0) * 1e3, 2), 'exposed_wait_ms_after_backward': round((t2 - t1) * 1e3, 3)},
        'peak_mem_GB': round(torch.cuda.max_memory_allocated() / 2**30, 1),
    }), flush=True)
  if world > 1:
    dist.destroy_process_group()


def run_twostage(args):
  """BASELINE config[3]: full two-stage MODE.  The 6 camera pairs of a frame are sharded over the ranks (north_star layout), the
  per-pair disparity / confidence maps are all-gathered (NCCL) into the fusion stage, which warps them into camera 1's frame in
  memory (StageBoundary: disp->depth, rotation / z-buffer forward warp) and runs ModeFusion.  With N ranks a step processes
  F = N / gcd(6, N) frames so that every rank holds the same number of pairs; frame f is fused on rank f mod N."""
  import math as _m
  dist, rank, world, local, dev = _dist_setup()
  from mode_2022_b200.models import ModeFusion
  from mode_2022_b200.sharding import gather_maps, shard_items
  from mode_2022_b200.utils.geometry import StageBoundary
  F_ = world // _m.gcd(6, world)
  n_items = 6 * F_
  mine = shard_items(n_items, rank, world)
  model = build_model(dev)
  from mode_2022_b200.pipeline import FusionStage
  fusion = ModeFusion(20.0, [32, 64, 128, 256], {'depth': 12, 'rgb': 12}, precision='fp16').to(dev).eval()
  stage2 = FusionStage(fusion, H, W, boundary=StageBoundary(), use_graph=not args.no_graph)  # boundary + fusion of one frame = one CUDA graph
  g = torch.Generator().manual_seed(300 + rank)
  left, right = torch.randn(len(mine), 3, H, W, generator=g).to(dev), torch.randn(len(mine), 3, H, W, generator=g).to(dev)
  rgbs = [torch.randn(1, 3, H, W, generator=g).to(dev) for _ in range(4)]
  my_frames = [f for f in range(F_) if f % world == rank]
  out_box = [None]

  def step():
    with torch.no_grad():
      pred, conf = model(left, right)
      both = gather_maps(torch.cat([pred, conf], 1), n_items)  # (6F, 2, H, W) on every rank: the north_star all-gather
      for f in my_frames:
        out_box[0] = stage2(both[6 * f:6 * f + 6, :1], both[6 * f:6 * f + 6, 1:], rgbs)

  for _ in range(max(args.warmup, 3)):
    step()
  sampler = ClockSampler(local)
  sampler.start()
  ms = _timed(dist, world, dev, step, args.steps)
  clocks = sampler.stop()
  # stage shares on rank 0 (CUDA events around each stage of one step)
  ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
  with torch.no_grad():
    ev[0].record()
    pred, conf = model(left, right)
    ev[1].record()
    both = gather_maps(torch.cat([pred, conf], 1), n_items)
    ev[2].record()
    for f in my_frames:
      stage2(both[6 * f:6 * f + 6, :1], both[6 * f:6 * f + 6, 1:], rgbs)
    ev[3].record()
  torch.cuda.synchronize()
  if rank == 0:
    print(json.dumps({
        'metric': 'two-stage MODE frames/s @512x1024 D=192 (6 pairs + fusion per frame)', 'value': round(F_ * args.steps / (ms * 1e-3), 3), 'unit': 'frames/s', 'n_gpus': world,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': round(ms / args.steps, 3), 'higher_is_better': True, 'scaling': 'weak' if F_ > 1 else 'strong',
        'vs_baseline': None, 'dtype': PRECISION + ' stereo stage, fp16 folded ModeFusion plan', 'data': 'synthetic',
        'config': {'workload': f'{F_} frame(s)/step: {n_items} camera pairs sharded over {world} GPU(s) -> NCCL all-gather of disp/conf maps ({n_items * 2 * H * W * 4 / 1e6:.0f} MB) -> '
                               'in-memory stage boundary (disp->depth, rotate / z-buffer warp) -> ModeFusion',
                   'pairs_per_gpu': len(mine), 'frames_fused_on_rank0': len(my_frames)},
        'pairs_per_s': round(n_items * args.steps / (ms * 1e-3), 2),
        'stage_ms_rank0': {'stereo': round(ev[0].elapsed_time(ev[1]), 3), 'all_gather': round(ev[1].elapsed_time(ev[2]), 3),
                           'boundary_plus_fusion': round(ev[2].elapsed_time(ev[3]), 3)},
        'clocks': clocks,
    }), flush=True)
  if world > 1:
    dist.destroy_process_group()


def run_highres(args):
  """BASELINE config[4]: 2048x1024 Cassini (= "1024x2048 equirect"), maxdisp=192, one pair per rank per step."""
  dist, rank, world, local, dev = _dist_setup()
  from mode_2022_b200 import _lib
  from mode_2022_b200.models import ModeDisparity
  Hh, Wh = 2048, 1024
  nb = args.pairs or 1
  torch.manual_seed(0)
  model = ModeDisparity(MAXDISP, conv='Sphere', in_height=Hh, in_width=Wh, sphereType='Cassini', out_conf=True, precision=PRECISION).to(dev).eval()
  g = torch.Generator().manual_seed(400 + rank)
  left, right = torch.randn(nb, 3, Hh, Wh, generator=g).to(dev), torch.randn(nb, 3, Hh, Wh, generator=g).to(dev)
  with torch.no_grad():
    for _ in range(max(args.warmup, 3)):
      model(left, right)
    ms = _timed(dist, world, dev, lambda: model(left, right), args.steps)
    roof = None
    if rank == 0:
      _lib.PROFILE = []
      for _ in range(2):
        model(left, right)
      torch.cuda.synchronize()
      prof, _lib.PROFILE = _lib.PROFILE, None
      roof = roofline_report(prof, 2)
  if rank == 0:
    print(json.dumps({
        'metric': 'stereo pairs/s @1024x2048 D=192', 'value': round(nb * world * args.steps / (ms * 1e-3), 2), 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': round(ms / args.steps, 3), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': PRECISION,
        'data': 'synthetic', 'config': {'workload': f'ModeDisparity stereo stage, {nb} pair(s)/step/GPU, Cassini 2048x1024 (=1024x2048 ERP), maxdisp=192, out_conf'},
        'roofline': roof, 'peak_mem_GB': round(torch.cuda.max_memory_allocated() / 2**30, 1),
    }), flush=True)
  if world > 1:
    dist.destroy_process_group()


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=10)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--no-graph', action='store_true')
  ap.add_argument('--mode', default='stereo', choices=['stereo', 'train', 'twostage', 'highres'],
                  help='stereo = BASELINE config[1] (the default line); train = config[2]; twostage = config[3]; highres = config[4]')
  ap.add_argument('--pairs', type=int, default=0, help='train: global batch (default 8); highres: pairs per step per GPU (default 1)')
  args = ap.parse_args()
  if args.impl == 'reference':
    run_reference(args)
  elif args.mode == 'train':
    run_train(args)
  elif args.mode == 'twostage':
    run_twostage(args)
  elif args.mode == 'highres':
    run_highres(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
