#!/usr/bin/env python3
"""Per-device-function share of warp time / executed instructions for one kernel of an ncu report.
python tools/ncu_by_function.py source.csv dis.txt KERNEL_SUBSTR
  source.csv = ncu -i REP --page source --csv ; dis.txt = nvdisasm -hex of the cubin that ran."""
import csv, re, sys, collections
src, dis, ker = sys.argv[1:4]
# function offset ranges inside the kernel's .text section
funcs = []; in_sec = False
for l in open(dis):
    if l.startswith('//--------------------- .text.'):
        in_sec = ker in l; continue
    if not in_sec: continue
    m = re.match(r'^(\$?_Z[^:]*|[\w$.]+):\s*$', l)
    if m and not m.group(1).startswith('.L'):
        funcs.append([m.group(1).split('$')[-1], None]); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+\S', l)
    if m and funcs and funcs[-1][1] is None:
        funcs[-1][1] = int(m.group(1), 16)
funcs = [f for f in funcs if f[1] is not None]
rows = list(csv.reader(open(src)))
# find the kernel's block of rows
start = None
for i, r in enumerate(rows):
    if r and r[0] == 'Kernel Name' and len(r) > 1 and ker.replace('ILi1E', '<(int)1>').split('<')[0].lstrip('_Z0123456789') in r[1]:
        start = i
        if ker.endswith('ILi1E') and '<(int)1>' not in r[1]: continue
        break
hdr = rows[start + 1]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[start + 2:]:
    if r and r[0] == 'Kernel Name': break
    if len(r) >= len(hdr) and r[ix['# Samples']].strip().isdigit(): data.append(r)
base = int(data[0][ix['Address']], 16)
agg = collections.OrderedDict((f[0], [0, 0, 0]) for f in funcs)
fi = 0
for r in data:
    off = int(r[ix['Address']], 16) - base
    while fi + 1 < len(funcs) and off >= funcs[fi + 1][1]: fi += 1
    a = agg[funcs[fi][0]]
    a[0] += int(r[ix['# Samples']]); a[1] += int(r[ix['Instructions Executed']])
    if 'IMAD.WIDE' in r[ix['Source']]: a[2] += int(r[ix['Instructions Executed']])
S = sum(a[0] for a in agg.values()); E = sum(a[1] for a in agg.values())
print('%-44s %8s %8s %10s' % ('function', 'time%', 'instr%', 'wideMAC%'))
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if a[0] == 0: continue
    print('%-44s %7.2f%% %7.2f%% %9.1f%%' % (k[:44], 100.0 * a[0] / S, 100.0 * a[1] / E, 100.0 * a[2] / max(1, a[1])))
