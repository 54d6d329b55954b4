"""Stage boundary of one frame at 1024x512 (target for ncu captures of the geometry kernels)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mode_2022_b200.utils.geometry import StageBoundary
g = torch.Generator().manual_seed(0)
disp = (torch.rand(6, 1, 1024, 512, generator=g) * 191).cuda()
conf = torch.rand(6, 1, 1024, 512, generator=g).cuda()
sb = StageBoundary()
for _ in range(3):
  out = sb(disp, conf)
torch.cuda.synchronize()
print('stage boundary ok')
