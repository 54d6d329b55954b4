"""Per-CTA timeline of one conv3d_bf16 launch (profiling aid): who ran where, for how long."""
import ctypes, math, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mode_2022_b200 import ops, _lib
lib = _lib.load()
dev = 'cuda'
ci, co, dims, mode = [int(v) for v in os.environ.get('CFG', '32,32,48,256,128,0').split(',')][:2] + [tuple(int(v) for v in os.environ.get('CFG', '32,32,48,256,128,0').split(',')[2:5])] + [int(os.environ.get('CFG', '32,32,48,256,128,0').split(',')[5])]
B = int(os.environ.get('BATCH', '1'))
x = torch.randn(B, *dims, ci, device=dev).bfloat16()
res = None
if os.environ.get('RES'):
  od = ops.conv3d_out_dims(*dims, mode)
  res = torch.randn(B, *od, co, device=dev).bfloat16()
w = torch.randn((ci, co, 3, 3, 3) if mode == 2 else (co, ci, 3, 3, 3), device=dev) / math.sqrt(27 * ci)
wp = ops.conv3d_pack_weights(w, mode)
for _ in range(3):
  ops.conv3d_bf16(x, wp, co, None, None, res, mode, True, False)
NMAX = 1024
dbg = torch.zeros(NMAX * 8 + 4 * 64 + 8 * NMAX, dtype=torch.int64, device=dev)  # room for any grid <= 1024
lib.mode_conv3d_set_debug_buffer(ctypes.c_void_p(dbg.data_ptr()))
torch.cuda.synchronize()
ops.conv3d_bf16(x, wp, co, None, None, res, mode, True, False)
torch.cuda.synchronize()
lib.mode_conv3d_set_debug_buffer(ctypes.c_void_p(0))
allrec = dbg[:NMAX * 8].view(NMAX, 8).cpu()
grid = int((allrec[:, 2] != 0).sum().item())  # CTAs that wrote an end time
d = allrec[:grid]
st = dbg[grid * 8:grid * 8 + 256].view(4, 64).cpu()
t0 = d[:, 1].min().item()
dur = (d[:, 2] - d[:, 1]).float() / 1e3
start = (d[:, 1] - t0).float() / 1e3
print('CTAs', grid, 'distinct SMs', len(set(d[:, 0].tolist())))
print('start offset us: min %.1f max %.1f' % (start.min(), start.max()))
print('duration us: min %.1f median %.1f max %.1f' % (dur.min(), dur.median(), dur.max()))
print('end us: max %.1f' % ((d[:, 2] - t0).float().max() / 1e3))
order = torch.argsort(dur)
for i in list(order[:5]) + list(order[-5:]):
  print('cta %3d sm %3d start %.1f dur %.1f items %d' % (i, d[i, 0], start[i], dur[i], d[i, 3]))
import collections
c = collections.Counter(d[:, 0].tolist())
print('SMs with >1 CTA:', {k: v for k, v in c.items() if v > 1})
print('dur by cta id (us):')
print(' '.join('%d' % round(v) for v in dur.tolist()))
sm = d[:, 0].tolist()
bysm = sorted(zip(sm, dur.tolist()))
print('dur by smid:')
print(' '.join('%d:%d' % (a, round(b)) for a, b in bysm))

print('MMA warp cycles: cta, total, wait_tempty, wait_full, in_mma_groups | epilogue warp0 wait_tfull')
for i in (0, 1, 2, 3, 4, 5, 6, 7, grid - 4, grid - 3, grid - 2, grid - 1):
  print(i, d[i, 6].item(), d[i, 4].item(), d[i, 5].item(), d[i, 3].item(), '|', d[i, 7].item())

for c in range(4):
  ts = st[c, :60].tolist()
  print('cta', c, 'stage deltas (cycles):', ' '.join(str(int(b - a)) for a, b in zip(ts[:-1], ts[1:])))

for c in range(4):
  print('cta', c, 'raw seg:', ' '.join(str(int(v)) for v in st[c, :12].tolist()))
