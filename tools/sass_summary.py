"""Per-kernel counts of the Blackwell-native SASS mnemonics in the built library (cuobjdump -sass):
UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAPF (TMA load / store / prefetch),
UBLKCP (bulk copy), UTCBAR (tcgen05.commit), USETMAXREG, HMMA (legacy mma.sync: expected 0).
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'vitpose_b200', 'libvitpose_b200.so')
KEYS = ['UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTMAPF', 'UBLKCP', 'UTCBAR', 'USETMAXREG', 'HMMA', 'MUFU.EX2']

out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
demangle = {}
names = re.findall(r'Function : (\S+)', out)
if names:
    dm = subprocess.run(['c++filt'] + names, capture_output=True, text=True).stdout.splitlines()
    demangle = dict(zip(names, dm))
counts, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        op = m.group(1)
        for k in KEYS:
            if op == k or op.startswith(k + '.') or (k == 'MUFU.EX2' and op.startswith('MUFU.EX2')):
                counts[cur][k] += 1
                if k == 'UTCHMMA' and '2CTA' in op:
                    counts[cur]['UTCHMMA.2CTA'] += 1
print(f'# cuobjdump -sass {os.path.relpath(LIB, ROOT)} ({os.path.getsize(LIB)} bytes): mnemonic counts per kernel')
tot = collections.Counter()
for fn, c in counts.items():
    tot.update(c)
    if not any(c[k] for k in ('UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP')):
        continue
    name = demangle.get(fn, fn)
    name = re.sub(r'\(CUtensorMap_st.*', '', name).replace('void vpb::', '')
    print(f'{name}: ' + ' '.join(f'{k}={v}' for k, v in c.items()))
print('TOTAL: ' + ' '.join(f'{k}={tot[k]}' for k in KEYS + ['UTCHMMA.2CTA']))
print(f'kernels in the library: {len(counts)}; legacy tensor-core HMMA instructions: {tot["HMMA"]}')
