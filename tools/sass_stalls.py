#!/usr/bin/env python3
"""Static issue-time estimate of one device function from its SASS control words (stall counts):
python tools/sass_stalls.py dis.txt KERNEL_SUBSTR FUNC_SUBSTR   (dis.txt = nvdisasm -hex cubin)"""
import re, sys, collections
lines = open(sys.argv[1]).read().split('\n')
ker, fn = sys.argv[2], sys.argv[3]
start = None
for i, l in enumerate(lines):
    if l.startswith('$') and ker in l and fn in l and l.rstrip().endswith(':'):
        start = i; break
    if fn == '' and l.startswith(ker) and l.rstrip().endswith(':'):
        start = i; break
assert start is not None
ops = collections.Counter(); stall_by = collections.Counter(); n = 0; tot = 0
i = start + 1
hist = collections.Counter()
while i < len(lines):
    l = lines[i]
    if (l.startswith('$') or l.startswith('_Z') or l.startswith('//---')) and l.rstrip().endswith(':') and i > start + 1:
        break
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/', l)
    if m:
        hi = re.search(r'/\* (0x[0-9a-f]+) \*/', lines[i + 1])
        h = int(hi.group(1), 16)
        stall = (h >> 41) & 0xf
        txt = m.group(2).strip()
        op = txt.split()[1] if txt.startswith('@') else txt.split()[0]
        op = '.'.join(op.split('.')[:2])
        ops[op] += 1; stall_by[op] += stall; n += 1; tot += stall
        if op.startswith('IMAD.WIDE'): hist[stall] += 1
        i += 2
    else:
        i += 1
print('instructions', n, 'sum of stall counts', tot)
for op, c in ops.most_common(14):
    print('  %-14s n=%5d  stall sum=%6d  avg=%.2f' % (op, c, stall_by[op], stall_by[op] / c))
print('IMAD.WIDE stall histogram', dict(hist))
