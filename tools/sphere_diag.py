"""Table classes of a sampling grid + per-kernel time split of one spherical-conv layer call (diagnostics)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mode_2022_b200 import ops
from mode_2022_b200.models.sphere_conv import sphere_position_numpy
dev = 'cuda'
H, W, B = 256, 128, 12
dt = torch.float16
pos = torch.from_numpy(sphere_position_numpy(W, H, 'Cassini')).to(dev)
tab = ops.sphere_gather_table(pos, dt)
npos = (H // 16) * (W // 8)
tail = tab[16 * 9 * H * W:].view(torch.int32).cpu()
hdr = tail[:4].tolist()
print('hdr {n_fast, n_rest, TH, TW}:', hdr)
x = torch.randn(B, H, W, 128, device=dev).to(dt)
res = torch.randn(B, H, W, 128, device=dev).to(dt)
w = torch.randn(128, 128, 3, 3, device=dev) / 34
wp = ops.sphere_conv_pack_weights(w, dt)
sc, sh = torch.ones(128, device=dev), torch.zeros(128, device=dev)
f = lambda: ops.sphere_conv_bf16(x, pos, wp, 128, sc, sh, res, True)
for _ in range(3):
  f()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
  for _ in range(5):
    f()
  torch.cuda.synchronize()
for e in prof.key_averages():
  t = getattr(e, 'device_time_total', 0)
  if t > 0:
    print('%8.1f us x%d  %s' % (t / e.count, e.count, e.key[:90]))
