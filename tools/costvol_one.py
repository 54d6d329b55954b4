"""Fused cost volume + dres0[0] at the bench shape (B=6, 256x128 features, D/4=48) vs the two-kernel path."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mode_2022_b200 import ops
dev = 'cuda'
B = int(os.environ.get('BATCH', '6'))
fl = torch.randn(B, 256, 128, 32, device=dev).half()
fr = torch.randn(B, 256, 128, 32, device=dev).half()
w = torch.randn(32, 64, 3, 3, 3, device=dev) / math.sqrt(27 * 64)
sc, sh = torch.ones(32, device=dev), torch.zeros(32, device=dev)
wr, wt = ops.costvol_conv_weights(w, torch.float16)
wp = ops.conv3d_pack_weights(w, 0, torch.float16)
fns = {'fused': lambda: ops.costvol_conv(fl, fr, wr, wt, sc, sh, 48, True),
       'cost_volume + conv3d_tc': lambda: ops.conv3d_bf16(ops.cost_volume(fl, fr, 48), wp, 32, sc, sh, None, 0, True, False)}
for name, f in fns.items():
  for _ in range(3):
    f()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(int(os.environ.get('ITERS', '10'))):
    f()
  b.record()
  torch.cuda.synchronize()
  print('%s: %.1f us' % (name, a.elapsed_time(b) / int(os.environ.get('ITERS', '10')) * 1e3))
