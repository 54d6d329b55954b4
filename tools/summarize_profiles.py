"""Turn gpurun_out/ evidence (tools/collect_profiles.sh) into the tracked summaries under profiles/."""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r01'


def short(name):
  name = re.sub(r'\(anonymous namespace\)::|<unnamed>::|void ', '', name)
  m = re.match(r'([A-Za-z0-9_:]+(<[^(]*>)?)', name)
  return (m.group(1) if m else name)[:110]


def launch_list():
  rows = list(csv.reader(l for l in open(os.path.join(OUT, 'launches.csv' if TAG == 'r01' else f'{TAG}_launches.csv')) if l.startswith('"')))
  h = rows[0]
  ik, im, iv, iid = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('ID')
  iu = h.index('Metric Unit')
  tscale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3}
  bscale = {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}
  per = collections.defaultdict(dict)
  for r in rows[1:]:  # every row carries its own (auto-scaled) unit: normalise to ms / MB
    val = float(r[iv].replace(',', ''))
    per[r[iid]]['k'] = short(r[ik])
    per[r[iid]][r[im]] = val * (tscale[r[iu]] if r[im].startswith('gpu__time') else bscale[r[iu]])
  agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
  for v in per.values():
    a = agg[v['k']]
    a[0] += 1
    a[1] += v.get('gpu__time_duration.sum', 0)
    a[2] += v.get('dram__bytes_read.sum', 0)
    a[3] += v.get('dram__bytes_write.sum', 0)
  tot = sum(a[1] for a in agg.values())
  lines = [f'# {TAG} -- ncu launch list (`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none`,',
           '# `python bench.py --steps 2 --warmup 1 --no-graph`, first launches incl. model setup and warm-up, capped by -c)', '',
           'Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.', '',
           f'total launches captured: {len(per)}; total kernel time {tot:.1f} ms', '', '| kernel | launches | total ms | share | dram read MB / launch | dram write MB / launch |', '|---|---:|---:|---:|---:|---:|']
  for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    lines.append(f'| `{k}` | {a[0]} | {a[1]:.2f} | {100 * a[1] / tot:.1f}% | {a[2] / a[0]:.1f} | {a[3] / a[0]:.1f} |')
  conv = [a for k, a in agg.items() if k.startswith(('conv3d_tc_kernel', 'conv3d_cls_tc_kernel'))]
  n = sum(a[0] for a in conv)
  per_launch = sum(a[2] + a[3] for a in conv) / max(n, 1)
  lines += ['', f'conv3d_tc_kernel + conv3d_cls_tc_kernel: {n} launches, {sum(a[1] for a in conv):.2f} ms, average DRAM traffic {per_launch:.1f} MB / launch '
            f'(x 27 launches per step = {per_launch * 27 / 1e3:.2f} GB / step of 6 pairs)']
  open(os.path.join(PROF, f'{TAG}_launch_list_summary.md'), 'w').write('\n'.join(lines) + '\n')
  return per_launch


KEYS = ['gpu__time_duration.sum', 'sm__cycles_active.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic', 'sm__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg',
        'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__warps_active.avg.per_cycle_active']


def full(rep, title, out):
  path = os.path.join(OUT, rep)
  raw = path.replace('.ncu-rep', '.raw.csv')  # exported on the GPU box (tools/collect_profiles_r02b.sh), the report itself stays there
  if os.path.exists(raw):
    txt = open(raw).read()
  elif os.path.exists(path):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  else:
    return
  r = list(csv.reader(txt.splitlines()))
  h, u, v = r[0], r[1], r[2]
  d = {a.split('TriageCompute.')[-1]: (b, c) for a, b, c in zip(h, u, v)}  # some metrics carry a '<unit>.TriageCompute.' prefix
  lines = [f'# {TAG} -- ncu --set full --clock-control none: {title}', '', f"kernel: `{short(d.get('Kernel Name', ('', ''))[1])}`", '', '| metric | unit | value |', '|---|---|---|']
  for k in KEYS:
    if k in d:
      lines.append(f'| {k} | {d[k][0]} | {d[k][1]} |')
  open(os.path.join(PROF, out), 'w').write('\n'.join(lines) + '\n')


R02 = [('r02_sphere_slab.ncu-rep', 'sphere_conv_slab_kernel<fp16, cassini> 128->128 @256x128, B=12, residual + ReLU (2688 of 3072 tiles)', 'sphere_conv_slab_ncu.md'),
       ('r02_sphere_direct.ncu-rep', 'sphere_conv_tc_kernel<fp16,128> in list mode: the 384 polar tiles of the same layer call', 'sphere_conv_direct_polar_ncu.md'),
       ('r02_conv3d_s1.ncu-rep', 'conv3d_tc_kernel<0,32,fp16,32> 32->32 stride 1 @48x256x128, B=6 (dominant kernel)', 'conv3d_tc_s1_ncu.md'),
       ('r02_deconv.ncu-rep', 'conv3d_tc_kernel<2,32,fp16,64> transposed 64->32 @24x128x64 -> 48x256x128, B=6, residual + ReLU', 'conv3d_tc_deconv_ncu.md'),
       ('r02_conv3d_s2small.ncu-rep', 'conv3d_tc_kernel<1,..> 64->64 stride 2 @24x128x64 -> 12x64x32, B=6 (small-grid tail)', 'conv3d_tc_s2_small_ncu.md'),
       ('r02_cls.ncu-rep', 'conv3d_cls_tc_kernel<fp16> 32->1 classifier, B=6, 48x256x128', 'conv3d_cls_tc_ncu.md'),
       ('r02_costvol.ncu-rep', 'costvol_conv_kernel<fp16> cost volume fused into dres0[0], B=6, 256x128 features, D/4=48', 'costvol_conv_ncu.md'),
       ('r02_regress.ncu-rep', 'disp_regress_192_kernel (row-tile kernel) upsample + softmax + soft-argmin + confidence, 1 pair 1024x512', 'disp_regress_ncu.md'),
       ('r02_cost_volume.ncu-rep', 'cost_volume_bf16_kernel (stand-alone 16-bit NDHWC cost volume), 1 pair', 'cost_volume_ncu.md'),
       ('r02_stem.ncu-rep', 'stem_conv_tc_kernel<fp16> 3->32 7x7 s2 @1024x512, B=12', 'stem_conv_tc_ncu.md'),
       ('r02_warp.ncu-rep', 'warp_scatter_kernel (z-buffer forward warp, pass 0), 1024x512', 'geometry_warp_ncu.md'),
       ('r02_sphere_f32.ncu-rep', 'sphere_conv_f32_tiled_kernel (training / fp32 parity forward: register-tiled SGEMM with fused gather; 512x256 D=96 B=2)', 'sphere_f32_fwd_ncu.md'),
       ('r02_sphere_dgrad.ncu-rep', 'sphere_dgrad_f32_tiled_kernel<deterministic> (training, 512x256 D=96 B=2)', 'sphere_dgrad_ncu.md'),
       ('r02_sphere_wgrad.ncu-rep', 'sphere_wgrad_f32_tiled_kernel<deterministic> (training)', 'sphere_wgrad_ncu.md'),
       ('r02_bn_cl_reduce.ncu-rep', 'bn_cl_reduce_kernel (training BatchNorm3d, channels-last: per-channel sums)', 'bn_cl_reduce_ncu.md'),
       ('r02_bn_cl_apply.ncu-rep', 'bn_cl_apply_kernel (training BatchNorm3d, channels-last: normalise / dx)', 'bn_cl_apply_ncu.md'),
       ('r02_bn_bwd_apply.ncu-rep', 'bn_bwd_apply_kernel (training BatchNorm2d backward, NCHW: dx)', 'bn_bwd_apply_ncu.md'),
       ('r02_regress_bwd.ncu-rep', 'disp_regress_bwd_kernel (training: fused soft-argmin head backward)', 'disp_regress_bwd_ncu.md'),
       ('r02_cost_volume_bwd.ncu-rep', 'cost_volume_bwd_f32_kernel (training: gather-sum over the shifts)', 'cost_volume_bwd_ncu.md')]

if __name__ == '__main__':
  os.makedirs(PROF, exist_ok=True)
  per_launch = launch_list()
  if TAG == 'r01':
    full('conv3d_s1_b6.ncu-rep', 'conv3d_tc_kernel<0,32,bf16,32> 32->32 stride 1 @48x256x128, B=6 (dominant kernel)', f'{TAG}_conv3d_tc_s1_ncu.md')
    full('deconv_b6.ncu-rep', 'conv3d_tc_kernel<2,32,bf16,64> transposed 64->32 @24x128x64 -> 48x256x128, B=6, residual + ReLU', f'{TAG}_conv3d_tc_deconv_ncu.md')
    full('sphere_b12.ncu-rep', 'sphere_conv_tc_kernel<bf16,128> 128->128 @256x128, B=12, residual + ReLU', f'{TAG}_sphere_conv_tc_ncu.md')
    full('costvol.ncu-rep', 'costvol_conv_kernel<bf16> cost volume fused into dres0[0], B=6, 256x128 features, D/4=48', f'{TAG}_costvol_conv_ncu.md')
    full('cls.ncu-rep', 'conv3d_cls_tc_kernel<bf16> 32->1 classifier, B=6, 48x256x128', f'{TAG}_conv3d_cls_tc_ncu.md')
    full('stem.ncu-rep', 'stem_conv_tc_kernel<bf16> 3->32 7x7 s2 @1024x512, B=12', f'{TAG}_stem_conv_tc_ncu.md')
    files = (('kernel_timings.txt', f'{TAG}_kernel_timings_b1.txt'), ('conv3d_layer_timings_b6.txt', f'{TAG}_layer_timings_b6.txt'))
    final = 'bench_final.json'
  else:
    for rep, title, out in R02:
      full(rep, title, f'{TAG}_{out}')
    files = ((f'{TAG}_kernel_timings.txt', f'{TAG}_kernel_timings_b1.txt'), (f'{TAG}_layer_timings_b6.txt', f'{TAG}_layer_timings_b6.txt'))
    final = f'{TAG}_bench_final.json'
  for f, o in files:
    if os.path.exists(os.path.join(OUT, f)):
      keep = [l for l in open(os.path.join(OUT, f)) if l.startswith(('{', 'conv3d_tc', 'sphere_conv_tc', 'direct-gather', 'pointwise', 'implicit-gemm', 'fused', 'cost_volume +'))]
      open(os.path.join(PROF, o), 'w').write(''.join(keep))
  for f, o in ((f'{TAG}_train_profile.txt', f'{TAG}_train_profile.txt'), (f'{TAG}_ref_train.txt', f'{TAG}_reference_train_step.txt'), (f'{TAG}_mufu_bench.txt', f'{TAG}_mufu_microbench.txt'),
               (f'{TAG}_sphere_split.txt', f'{TAG}_sphere_kernel_split.txt')):
    if os.path.exists(os.path.join(OUT, f)):
      keep = [l for l in open(os.path.join(OUT, f)) if 'Warn' not in l and 'warn' not in l and l.strip()]
      open(os.path.join(PROF, o), 'w').write(''.join(keep))
  for mode in ('train', 'twostage', 'highres'):
    f = os.path.join(OUT, f'{TAG}_bench_{mode}.json')
    if os.path.exists(f):
      lines = [l for l in open(f) if l.startswith('{')]
      if lines:
        open(os.path.join(PROF, f'{TAG}_bench_line_{mode}.json'), 'w').write(lines[-1])
  if os.path.exists(os.path.join(OUT, final)):
    line = [l for l in open(os.path.join(OUT, final)) if l.startswith('{')][-1]
    open(os.path.join(PROF, f'{TAG}_bench_line.json'), 'w').write(line)
  print('conv3d_tc average DRAM MB per launch: %.1f' % per_launch)
