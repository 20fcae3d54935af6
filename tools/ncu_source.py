#!/usr/bin/env python3
"""Aggregates the ncu source page (ncu -i REP --page source --csv > file): samples by opcode and by
stall reason, plus the hottest contiguous regions.  python tools/ncu_source.py file.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter(); byop = collections.Counter(); execop = collections.Counter()
opstall = collections.defaultdict(collections.Counter)
insts = []
for r in rows[2:]:
    if len(r) < len(hdr) or not r[ix['# Samples']].strip().isdigit(): continue   # kernel/header rows of multi-kernel reports
    src = r[ix['Source']].strip()
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    op = op.rstrip(';')
    base = '.'.join(op.split('.')[:3])
    smp = int(r[ix['# Samples']] or 0); ex = int(r[ix['Instructions Executed']] or 0)
    byop[base] += smp; execop[base] += ex
    for s in stalls:
        v = int(r[ix[s]] or 0); tot[s] += v; opstall[base][s] += v
    insts.append((r[ix['Address']], src, smp, ex))
S = sum(byop.values()); E = sum(execop.values())
print('total samples', S, 'total warp-instr', E)
print('--- stall reasons'); 
for s, v in tot.most_common(12): print('  %-26s %6.2f%%' % (s, 100.0 * v / S))
print('--- by opcode: exec share, sample share, top stalls')
for op, v in byop.most_common(22):
    top = ', '.join('%s %.0f%%' % (s[6:], 100.0 * c / max(1, v)) for s, c in opstall[op].most_common(3))
    print('  %-22s exec %5.1f%%  samples %5.1f%%   %s' % (op, 100.0 * execop[op] / E, 100.0 * v / S, top))
