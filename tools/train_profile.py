"""Kernel-time breakdown of one full-size training step (1 pair, 1024x512, D=192): torch.profiler, CUDA activities only."""
import os, sys, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mode_2022_b200.models import ModeDisparity
from mode_2022_b200 import training as T
H, W, D, B = 1024, 512, 192, int(os.environ.get('BATCH', 1))
torch.backends.cudnn.benchmark = os.environ.get('CUDNN_BENCHMARK', '1') != '0'  # the reference trains with cudnn.benchmark on (train_disparity.py:82)
torch.manual_seed(0)
m = ModeDisparity(D, in_height=H, in_width=W, sphereType='Cassini', precision='fp32').cuda().train()
red = T.GradAllReduce(m.parameters())
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
g = torch.Generator().manual_seed(1)
left, right = torch.randn(B, 3, H, W, generator=g).cuda(), torch.randn(B, 3, H, W, generator=g).cuda()
disp = (torch.rand(B, 1, H, W, generator=g) * (D - 1)).cuda()
mask = (torch.rand(B, 1, H, W, generator=g) < 0.9).cuda()
for _ in range(3):
  T.train_step(m, red, opt, left, right, disp, mask)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
T.train_step(m, red, opt, left, right, disp, mask)
b.record()
torch.cuda.synchronize()
print('step: %.1f ms for %d pair(s)' % (a.elapsed_time(b), B))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  T.train_step(m, red, opt, left, right, disp, mask)
  torch.cuda.synchronize()
tot = 0
rows = []
for e in prof.key_averages():
  t = getattr(e, 'device_time_total', 0)
  if t > 0:
    rows.append((t, e.count, e.key))
    tot += t
rows.sort(reverse=True)
print('total kernel time %.1f ms' % (tot / 1e3))
for t, c, k in rows[:28]:
  print('%9.2f ms %5.1f%% x%-4d %s' % (t / 1e3, 100 * t / tot, c, k[:110]))
