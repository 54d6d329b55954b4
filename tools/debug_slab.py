"""Debug harness for the slab spherical-conv kernel: one-hot weights isolate the A operand (gather) per tap."""
import math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mode_2022_b200 import ops
from oracle import mode_oracle as O

torch.set_printoptions(linewidth=250, precision=3, sci_mode=False)
dt = torch.float16
for (B, C, Co, h, w, st) in [(1, 64, 128, 16, 8, 'Cassini'), (1, 128, 128, 32, 16, 'Cassini'), (2, 128, 128, 256, 128, 'Cassini'), (1, 128, 128, 16, 32, 'ERP'), (1, 128, 128, 128, 256, 'ERP')]:
  g = torch.Generator().manual_seed(1)
  x = torch.randn(B, C, h, w, generator=g).to(dt)
  pos = torch.from_numpy(O.gen_sphere_position(h, w, st))
  xl = x.permute(0, 2, 3, 1).contiguous().cuda()
  print(f'==== B={B} C={C} Co={Co} {h}x{w} {st}')
  # table header
  tab = ops.sphere_gather_table(pos.cuda(), dt)
  hdr = tab[16 * 9 * h * w:16 * 9 * h * w + 16].view(torch.int32).cpu().tolist()
  print('hdr n_fast, n_rest, TH, TW =', hdr)
  if hdr[0] + hdr[1] > 0:
    npos = hdr[0] + hdr[1]
    info = tab[16 * 9 * h * w + 16:16 * 9 * h * w + 16 + 16 * npos].view(torch.int32).view(npos, 4).cpu()
    print('info[:8] (l0, s0, L, ty<<16|tx):', info[:8].tolist())
    print('L histogram:', torch.bincount(info[:, 2]).tolist())
  for k0 in [0, 4, 8, -1]:
    if k0 >= 0:
      wt = torch.zeros(Co, C, 3, 3)
      for o in range(Co):
        wt[o, o % C, k0 // 3, k0 % 3] = 1.0
    else:
      wt = torch.randn(Co, C, 3, 3, generator=g) / math.sqrt(9 * C)
    want = O.sphere_conv(x.float(), pos, wt.to(dt).float())
    wp = ops.sphere_conv_pack_weights(wt.cuda(), dt)
    got = ops.sphere_conv_bf16(xl, pos.cuda(), wp, Co, None, None, None, False).float().cpu().permute(0, 3, 1, 2)
    torch.cuda.synchronize()
    err = (got - want).abs()
    bad = err > 2e-2 * want.abs().clamp_min(1.0)
    print(f'  tap {k0}: max err {err.max().item():.4f}, bad fraction {bad.float().mean().item():.4f}')
    if bad.any() and h * w <= 512:
      print('   bad pixel map (any channel):')
      print((bad.any(1)[0]).int())
      bc = bad.any(2).any(2)[0].nonzero().flatten().tolist()
      print('   bad channels:', bc[:40], '...' if len(bc) > 40 else '')
      b0 = bad[0].nonzero()[0].tolist()
      print('   first bad (o,h,w):', b0, 'got', got[0, b0[0], b0[1], b0[2]].item(), 'want', want[0, b0[0], b0[1], b0[2]].item())
    elif bad.any():
      pm = bad.any(1)[0]
      print('   bad rows (count per row, first 32):', pm.sum(1)[:32].tolist())
      print('   bad cols (count per col):', pm.sum(0).tolist())
      bc = bad.any(2).any(2)[0].nonzero().flatten().tolist()
      print('   bad channels:', bc[:40], '...' if len(bc) > 40 else '')
