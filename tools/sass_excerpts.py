"""profiles/r02_sass_excerpts.md: static SASS instruction counts per kernel of the tensor-core objects (cuobjdump -sass build/obj/*.o)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else 'r02'
OBJS = ['sphere_conv_tc', 'conv3d_tc', 'conv3d_cls_tc', 'stem_conv_tc', 'costvol_conv', 'disp_regress', 'sphere_conv_f32', 'sphere_conv_bwd_f32', 'batchnorm']
COLS = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'HMMA', 'MUFU.EX2', 'FFMA', 'ACQBULK|PREEXIT']
lines = [f'# {TAG} -- SASS evidence of the Blackwell-native kernels (`cuobjdump -sass build/obj/<file>.o`, sm_100a; `python tools/sass_excerpts.py`)', '',
         'Mnemonics: `UTCHMMA` = tcgen05.mma (kind::f16), `LDTM` / `STTM` = tcgen05.ld / tcgen05.st (tensor memory), `UTMALDG` / `UTMASTG` = TMA tensor load / store',
         '(cp.async.bulk.tensor), `UBLKCP` = cp.async.bulk (1-D bulk copy), `SYNCS` = mbarrier, `MUFU.EX2` = the exponential of the soft-argmin, `FFMA` = fp32 FMA (the CUDA-core SGEMM',
         'kernels of the fp32 / training path), `ACQBULK|PREEXIT` = griddepcontrol (programmatic dependent launch). Counts are static instruction counts per kernel.', '',
         '| object | kernel | ' + ' | '.join(COLS) + ' |', '|---|---|' + '---:|' * len(COLS)]
for o in OBJS:
  path = os.path.join(ROOT, 'build', 'obj', o + '.o')
  if not os.path.exists(path):
    continue
  txt = subprocess.run(['cuobjdump', '-sass', path], capture_output=True, text=True).stdout
  cur, cnt = None, collections.OrderedDict()
  for l in txt.splitlines():
    m = re.search(r'Function : (\S+)', l)
    if m:
      cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
      cur = re.sub(r'\(anonymous namespace\)::', '', cur)
      cur = re.sub(r'\(.*', '', cur)
      cnt[cur] = collections.Counter()
      continue
    m = re.match(r'\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
    if m and cur:
      op = m.group(1)
      for c in COLS:
        if any(op.startswith(a) for a in c.split('|')):
          cnt[cur][c] += 1
  for k, c in cnt.items():
    if sum(c.values()) == 0:
      continue
    lines.append(f'| {o}.cu | `{k}` | ' + ' | '.join(str(c[x]) for x in COLS) + ' |')
open(os.path.join(ROOT, 'profiles', f'{TAG}_sass_excerpts.md'), 'w').write('\n'.join(lines) + '\n')
print('\n'.join(lines[:14]))
