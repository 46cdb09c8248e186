"""Summarise `ncu --page source --csv` output: per kernel, warp instructions and stall samples by SASS opcode, and
the hottest instructions.  Development aid."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else None
kernels, cur = [], None
for r in rows:
    if r and r[0] == 'Kernel Name':
        cur = {'name': r[1], 'hdr': None, 'data': []}
        kernels.append(cur)
    elif cur is not None and cur['hdr'] is None:
        cur['hdr'] = r
    elif cur is not None and len(r) == len(cur['hdr']):
        cur['data'].append(r)
for k_i, k in enumerate(kernels):
    if which is not None and k_i != which:
        continue
    hdr, data = k['hdr'], k['data']
    iS, iE, iN = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    tot = sum(int(r[iE]) for r in data)
    ts = max(1, sum(int(r[iN]) for r in data))
    print(f'== kernel {k_i}: {k["name"][:110]}\n   warp instructions {tot}, static {len(data)}, samples {ts}')
    ops, samp = collections.Counter(), collections.Counter()
    for r in data:
        t = r[iS].split()
        op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
        ops[op] += int(r[iE])
        samp[op] += int(r[iN])
    for op, c in ops.most_common(24):
        print(f'   {op:12s} {c:10d} {100 * c / tot:5.1f} %   samples {100 * samp[op] / ts:5.1f} %')
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    st = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
    print('   stalls:', {k2: round(100 * v / ts, 1) for k2, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]})
    top = sorted(data, key=lambda r: -int(r[iN]))[:14]
    for r in top:
        print(f'   {int(r[iN]):6d} samples  exec {int(r[iE]):8d}  {r[iS].strip()[:90]}')
