"""One transposed conv3d layer at the bench shape (64->32 @24x128x64 -> 48x256x128, B=6, residual + ReLU): target for ncu."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mode_2022_b200 import ops
dev = 'cuda'
ci, co, d, h, w, mode = [int(v) for v in os.environ.get('CFG', '64,32,24,128,64,2').split(',')]
B = int(os.environ.get('BATCH', '6'))
x = torch.randn(B, d, h, w, ci, device=dev).half()
od = ops.conv3d_out_dims(d, h, w, mode)
res = torch.randn(B, *od, co, device=dev).half() if not os.environ.get('NORES') else None
wt = torch.randn((ci, co, 3, 3, 3) if mode == 2 else (co, ci, 3, 3, 3), device=dev) / math.sqrt(27 * ci)
wp = ops.conv3d_pack_weights(wt, mode, torch.float16)
sc, sh = torch.ones(co, device=dev), torch.zeros(co, device=dev)
f = lambda: ops.conv3d_bf16(x, wp, co, sc, sh, res, mode, True, False)
for _ in range(3):
  f()
torch.cuda.synchronize()
n = int(os.environ.get('ITERS', '10'))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(n):
  f()
b.record()
torch.cuda.synchronize()
print('conv3d_tc CFG=%s B=%d: %.1f us/launch' % (os.environ.get('CFG', '64,32,24,128,64,2'), B, a.elapsed_time(b) / n * 1e3))
