"""One training step of the stereo stage at a small size (target for ncu captures of the backward kernels)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mode_2022_b200.models import ModeDisparity
from mode_2022_b200 import training as T
H, W, D, B = int(os.environ.get('H', 512)), int(os.environ.get('W', 256)), int(os.environ.get('D', 96)), int(os.environ.get('BATCH', 2))
torch.manual_seed(0)
m = ModeDisparity(D, in_height=H, in_width=W, sphereType='Cassini', precision='fp32').cuda().train()
red = T.GradAllReduce(m.parameters())
opt = torch.optim.Adam(m.parameters(), lr=1e-3)
g = torch.Generator().manual_seed(1)
left, right = torch.randn(B, 3, H, W, generator=g).cuda(), torch.randn(B, 3, H, W, generator=g).cuda()
disp = (torch.rand(B, 1, H, W, generator=g) * (D - 1)).cuda()
mask = (torch.rand(B, 1, H, W, generator=g) < 0.9).cuda()
for _ in range(int(os.environ.get('ITERS', 2))):
  loss = T.train_step(m, red, opt, left, right, disp, mask)
torch.cuda.synchronize()
print('train step ok, loss', float(loss))
