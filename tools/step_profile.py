"""Kernel-level time breakdown of one bench step (torch.profiler, eager mode): which kernels make up the non-custom part."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device('cuda', 0)
model = bench.build_model(dev)
g = torch.Generator().manual_seed(0)
left = torch.randn(6, 3, 1024, 512, generator=g).to(dev)
right = torch.randn(6, 3, 1024, 512, generator=g).to(dev)
with torch.no_grad():
  for _ in range(3):
    model(left, right)
  torch.cuda.synchronize()
  from torch.profiler import profile, ProfilerActivity
  with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
      model(left, right)
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
  t = getattr(e, 'device_time_total', None) or getattr(e, 'cuda_time_total', 0)
  if t > 0:
    rows.append((t / 3.0 / 1000.0, e.count // 3, e.key[:100]))
rows.sort(reverse=True)
tot = sum(r[0] for r in rows)
print('total kernel ms/step %.2f' % tot)
for t, c, k in rows[:22]:
  print('%7.3f ms  x%-4d %s' % (t, c, k))
