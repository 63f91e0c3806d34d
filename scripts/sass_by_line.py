#!/usr/bin/env python
"""Join ncu's per-SASS-instruction counts with nvdisasm line info: instructions per source line.
usage: sass_by_line.py <dis.txt from nvdisasm -g -c> <ncu source csv> <mangled kernel substring> <points>"""
import collections
import csv
import re
import sys

dis, src_csv, kern, npts = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('//---') and '.text.' in l and kern in l)
cur, seq = None, []
for l in lines[start + 1:]:
    if l.startswith('//---') and seq:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        seq.append((m.group(2), cur))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
si, ie, ns = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
seen, inst = set(), []
for r in rows[2:]:
    if len(r) <= ie or not r[ie].isdigit() or r[0] in seen:
        continue
    seen.add(r[0])
    inst.append((int(r[ie]), int(r[ns])))
assert len(inst) == len(seq), (len(inst), len(seq))
by, sm = collections.Counter(), collections.Counter()
for (txt, cur), (n, s) in zip(seq, inst):
    by[cur] += n
    sm[cur] += s
tot = sum(by.values())
import os
_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'grid_vision_b200', 'csrc')
_src = {}
def src_line(k):
    if not k:
        return ''
    f = os.path.join(_root, k[0])
    if k[0] not in _src:
        _src[k[0]] = open(f).read().split('\n') if os.path.exists(f) else []
    t = _src[k[0]]
    return t[k[1] - 1].strip()[:88] if 0 < k[1] <= len(t) else ''
print(f'total warp-inst {tot}, thread-inst/pt {tot * 32 / npts:.1f}, samples {sum(sm.values())}')
for k, v in by.most_common(int(sys.argv[5]) if len(sys.argv) > 5 else 40):
    t = src_line(k)
    print(f'{100 * v / tot:5.1f}% {v * 32 / npts:6.1f}/pt samp {100 * sm[k] / sum(sm.values()):4.1f}% {k[0][:14] if k else None}:{k[1] if k else 0} {t}')
