"""One spherical-conv layer at the bench shape (B=12 images, 128->128 @256x128, residual + ReLU): timing, or a target for ncu."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mode_2022_b200 import ops
from mode_2022_b200.models.sphere_conv import sphere_position_numpy
dev = 'cuda'
B = int(os.environ.get('BATCH', '12'))
dt = torch.float16 if os.environ.get('FP16') else torch.bfloat16
pos = torch.from_numpy(sphere_position_numpy(128, 256, 'Cassini')).to(dev)
x = torch.randn(B, 256, 128, 128, device=dev).to(dt)
res = torch.randn(B, 256, 128, 128, device=dev).to(dt)
w = torch.randn(128, 128, 3, 3, device=dev) / 34
wp = ops.sphere_conv_pack_weights(w, dt)
sc, sh = torch.ones(128, device=dev), torch.zeros(128, device=dev)
f = lambda: ops.sphere_conv_bf16(x, pos, wp, 128, sc, sh, res, True)
for _ in range(3):
  f()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = int(os.environ.get('ITERS', '10'))
a.record()
for _ in range(n):
  f()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / n
print('sphere_conv_tc B=%d: %.1f us/launch, %.1f TFLOP/s' % (B, ms * 1e3, 2 * 128 * 128 * 9 * 256 * 128 * B / ms / 1e9))
