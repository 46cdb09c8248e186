"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time and share
over a window of launches.  usage: launch_list.py <csv> <first> <count> <title>
       or: launch_list.py <csv> after:<kernel substring>:<k> until:<kernel substring> <title>   (window = the launches after the
           k-th occurrence of the first kernel up to and including the next occurrence of the second one: one step)"""
import csv
import re
import sys
from collections import defaultdict

path, a2, a3, title = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
rows = []
with open(path, newline='') as f:
    lines = [ln for ln in f if ln.startswith('"')]
rd = csv.reader(lines)
hdr = next(rd)
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
for r in rd:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(',', ''))
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0, 'nsecond': 1e-3, 'msecond': 1e3}.get(r[iu], 1.0)
    rows.append((r[ik], v))
if a2.startswith('after:'):
    _, name, k = a2.split(':')
    hits = [i for i, (kn, _) in enumerate(rows) if name in kn]
    first = hits[int(k) - 1] + 1
    end = next(i for i in range(first, len(rows)) if a3.split(':', 1)[1] in rows[i][0])
    count = end - first + 1
else:
    first, count = int(a2), int(a3)
win = rows[first:first + count]
agg = defaultdict(lambda: [0, 0.0])
for k, v in win:
    k = re.sub(r'\(.*$', '', k).replace('void ', '')
    k = re.sub(r'at::native::.*?(elementwise|FillFunctor|index|copy).*', r'at::native::\1*', k)
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in win)
print(f'# {title}\n\nlaunches in window: {len(win)}   sum of kernel time: {tot / 1e3:.2f} ms\n')
print('| kernel | launches | ms | share |\n|---|---:|---:|---:|')
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| `{k[:90]}` | {n} | {v / 1e3:.3f} | {100 * v / tot:.1f}% |')
