"""Print the summary metrics (the KEYS of tools/summarize_profiles.py + stall reasons) of one .ncu-rep:  python tools/ncu_keys.py gpurun_out/x.ncu-rep [extra regex]"""
import csv, re, subprocess, sys
sys.path.insert(0, __file__.rsplit('/', 1)[0])
from summarize_profiles import KEYS
txt = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(txt.splitlines()))
h, u, v = r[0], r[1], r[2]
d = {a.split('TriageCompute.')[-1]: (b, c) for a, b, c in zip(h, u, v)}
extra = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
for k in d:
  if k in KEYS or (extra and extra.search(k)) or 'issue_stalled' in k and 'per_issue_active' in k and 'average' in k:
    print(f'{k:100s} {d[k][0]:12s} {d[k][1]}')
