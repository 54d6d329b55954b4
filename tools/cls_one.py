"""The 32->1 classifier conv at the bench shape (B=6, 48x256x128): pointwise kernel vs the implicit-GEMM kernel."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mode_2022_b200 import ops
dev = 'cuda'
B = int(os.environ.get('BATCH', '6'))
x = torch.randn(B, 48, 256, 128, 32, device=dev).half()
w = torch.randn(1, 32, 3, 3, 3, device=dev) / math.sqrt(27 * 32)
res = torch.randn(B, 48, 256, 128, device=dev)
wp = ops.conv3d_pack_weights(w, 0, torch.float16)
fns = {'pointwise': lambda: ops.conv3d_classifier(x, w, res), 'implicit-gemm': lambda: ops.conv3d_bf16(x, wp, 1, None, None, res.unsqueeze(-1), 0, False, True)}
for name, f in fns.items():
  for _ in range(3):
    f()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(10):
    f()
  b.record()
  torch.cuda.synchronize()
  ms = a.elapsed_time(b) / 10
  print('%s: %.1f us/launch (%.0f GB/s of input+output)' % (name, ms * 1e3, (x.numel() * 2 + 2 * res.numel() * 4) / ms / 1e6))
