#!/usr/bin/env python
"""Static SASS census of the built library: which kernels carry tcgen05 (UTCHMMA / UTCQMMA, LDTM, UTCBAR), TMA (UTMALDG /
UTMASTG / UTMAREDG), cluster barriers, fp32 atomics and legacy tensor-core instructions (HMMA - there must be none).

    python scripts/sass_summary.py [path/to/libcreamfl_b200.so] > profiles/rNN_sass_summary.md

Runs anywhere cuobjdump is installed (no GPU needed); the mnemonics are the ones /opt/skills/guides/B200_PROFILING.md
lists as proof of tcgen05 / TMA code paths."""
from __future__ import annotations

import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
WATCH = ['UTCHMMA', 'UTCQMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTCCP', 'UTMALDG', 'UTMASTG', 'UTMAREDG', 'UTMAPF',
         'UCGABAR', 'SYNCS', 'RED', 'ATOM', 'HMMA', 'IMMA', 'MUFU', 'LDGSTS']


def demangle(names):
    out = subprocess.run(['cu++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out)) if len(out) == len(names) else {n: n for n in names}


def main():
    lib = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / 'creamfl_b200' / 'libcreamfl_b200.so'
    sass = subprocess.run(['cuobjdump', '-sass', str(lib)], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    arch = set()
    cur = None
    op = re.compile(r'^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)')
    for line in sass.splitlines():
        if line.lstrip().startswith('Function :'):
            cur = line.split(':', 1)[1].strip()
            per[cur] = collections.Counter()
        elif 'arch = ' in line:
            arch.add(line.split('arch = ')[1].strip())
        elif cur is not None:
            m = op.match(line)
            if m:
                per[cur]['_total'] += 1
                base = m.group(1)
                for w in WATCH:
                    if base.startswith(w):
                        per[cur][w] += 1
                        break
    names = demangle(list(per))
    fam = collections.OrderedDict()
    for mangled, c in per.items():
        d = names[mangled]
        d = d.replace('(anonymous namespace)::', '').replace('<unnamed>::', '')
        key = re.sub(r'<.*', '', re.sub(r'^void ', '', d)).replace('cfl::', '').split('(')[0]
        f = fam.setdefault(key, [0, collections.Counter()])
        f[0] += 1
        f[1].update(c)
    cols = ['UTCHMMA', 'LDTM', 'UTCBAR', 'UTMALDG', 'UTMASTG', 'UCGABAR', 'RED', 'ATOM', 'MUFU', 'HMMA']
    tot = collections.Counter()
    for _, c in fam.values():
        tot.update(c)
    print(f'# SASS census of `{lib.relative_to(ROOT) if lib.is_relative_to(ROOT) else lib}`\n')
    print(f'`cuobjdump -sass`, {len(per)} kernels (template instantiations), arch {", ".join(sorted(arch))}; counts are '
          'static instructions summed over the instantiations of a kernel.  UTCHMMA = `tcgen05.mma` (bf16), LDTM = '
          '`tcgen05.ld`, UTCBAR = `tcgen05.commit`, UTMALDG / UTMASTG = TMA bulk-tensor load / store, UCGABAR = cluster '
          'barrier, RED / ATOM = global reductions / atomics, MUFU = special-function unit; HMMA (legacy `mma.sync`) '
          'must be absent.\n')
    print('| kernel | instantiations | instructions | ' + ' | '.join(cols) + ' |')
    print('|---|---:|---:|' + '---:|' * len(cols))
    for key, (n, c) in sorted(fam.items(), key=lambda kv: -kv[1][1]['UTCHMMA'] * 10**6 - kv[1][1]['_total']):
        print(f'| `{key}` | {n} | {c["_total"]} | ' + ' | '.join(str(c[w]) if c[w] else '' for w in cols) + ' |')
    print(f'| **total** | {len(per)} | {tot["_total"]} | ' + ' | '.join(str(tot[w]) for w in cols) + ' |')
    print(f'\nLegacy tensor-core instructions: HMMA {tot["HMMA"]}, IMMA {tot["IMMA"]}; `cp.async` (LDGSTS) {tot["LDGSTS"]}.')


if __name__ == '__main__':
    main()
