#!/usr/bin/env python
"""profiles/r02_sass_by_line.md: instruction accounting of the hot kernel before (k_points_col,
start of round 2) and after (k_points_pair), from ncu source-page CSVs.
usage: make_sass_report.py <before src csv> <before points> <after src csv> <after points> <after .dis> <after kernel substr> <out.md>"""
import collections, csv, re, subprocess, sys

b_csv, b_pts, a_csv, a_pts, a_dis, a_kern, out = sys.argv[1:8]
b_pts, a_pts = float(b_pts), float(a_pts)


def load(path):
    rows = list(csv.reader(open(path)))
    name = rows[0][1] if len(rows[0]) > 1 else '?'
    hdr = rows[1]
    si, ie, te, ns = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
    seen, o = set(), []
    for r in rows[2:]:
        if len(r) <= ie or not r[ie].isdigit() or r[0] in seen:
            continue
        seen.add(r[0])
        o.append((r[si].strip(), int(r[ie]), int(r[te]), int(r[ns])))
    return name, o


def tally(ins, pts):
    c = collections.Counter()
    for s, n, t, sa in ins:
        m = re.match(r'(@!?U?P\d\s+)?([A-Z0-9_]+)', s)
        c[m.group(2)] += n * 32 / pts
    return c


GROUPS = collections.OrderedDict([
    ('packed FP32 (FMUL2/FFMA2/FADD2)', ('FMUL2', 'FFMA2', 'FADD2')),
    ('scalar FP32 (FMUL/FADD/FFMA/FMNMX/FSEL/FMNMX3)', ('FMUL', 'FADD', 'FFMA', 'FMNMX', 'FSEL', 'FMNMX3')),
    ('FP64 + conversions (DFMA/F2F/F2I/MUFU)', ('DFMA', 'F2F', 'F2I', 'MUFU', 'DADD', 'DMUL', 'I2F')),
    ('compares + predicate logic (FSETP/ISETP/PLOP3)', ('FSETP', 'ISETP', 'PLOP3', 'UISETP')),
    ('integer / bit / select (IMAD/IADD3/VIADD/LOP3/SHF/SEL/LEA/PRMT/FLO/VIMNMX/MOV/CS2R/HFMA2)',
     ('IMAD', 'IADD3', 'VIADD', 'LOP3', 'SHF', 'SEL', 'LEA', 'PRMT', 'FLO', 'VIMNMX', 'MOV', 'CS2R', 'HFMA2', 'IABS', 'UIADD3', 'UMOV', 'ULEA', 'UIMAD', 'USHF', 'S2UR', 'S2R', 'USEL')),
    ('constant-bank loads (LDC/LDCU)', ('LDC', 'LDCU')),
    ('branches + reconvergence (BRA/BSSY/BSYNC/BREAK/WARPSYNC/NOP/EXIT)', ('BRA', 'BSSY', 'BSYNC', 'BREAK', 'WARPSYNC', 'NOP', 'EXIT', 'BAR')),
    ('memory (LDG/STG/LDS/STS/REDG/LDL/STL)', ('LDG', 'STG', 'LDS', 'STS', 'REDG', 'LDL', 'STL', 'ATOMG')),
])


def table(c):
    tot = sum(c.values())
    rows, used = [], set()
    for g, ops in GROUPS.items():
        v = sum(c[o] for o in ops)
        used.update(ops)
        rows.append((g, v))
    rest = sum(v for k, v in c.items() if k not in used)
    if rest > 0.05:
        rows.append(('other', rest))
    return tot, rows


bn, b = load(b_csv)
an, a = load(a_csv)
bt, br = table(tally(b, b_pts))
at, ar = table(tally(a, a_pts))
md = ['# Round 2 — instruction accounting of the fused point kernel (ncu source page, `--import-source on`)', '',
      'Thread-instructions per point = 32 x warp-level `Instructions Executed` / points of the launch (a warp instruction',
      'counts 32 whatever its active mask: what matters on an issue-bound kernel).', '',
      f'* before: `{bn}` — the kernel at the start of round 2 ({b_pts:.0f} points per launch)',
      f'* after:  `{an}` — this round\'s kernel ({a_pts:.0f} points per launch)', '',
      '| instruction class | before /pt | after /pt |', '|---|---|---|']
bd = dict(br)
for g, v in ar:
    md.append(f'| {g} | {bd.get(g, 0.0):.1f} | {v:.1f} |')
md += [f'| **total** | **{bt:.1f}** | **{at:.1f}** |', '',
       '(Round 1\'s `k_points<1,1,0,0,1>`: 266 thread-instructions per point, `profiles/r01_ncu_c3_4096frames.md`.)', '']
# per source line, after
md += ['## After: `k_points_pair` by source line (nvdisasm -g line table joined with the per-instruction counts)', '',
       '```']
r = subprocess.run([sys.executable, __file__.replace('make_sass_report.py', 'sass_by_line.py'), a_dis, a_csv, a_kern, str(int(a_pts)), '40'],
                   capture_output=True, text=True)
md += r.stdout.rstrip().split('\n')
md += ['```', '']
open(out, 'w').write('\n'.join(md))
print('\n'.join(md[:24]))
