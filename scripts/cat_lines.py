#!/usr/bin/env python
"""Aggregate a by-line table (scripts/bylines.sh) into named source ranges.  usage: cat_lines.py table.txt file:lo-hi=name ..."""
import re, sys
from collections import defaultdict
rows = []
for ln in open(sys.argv[1]).read().split('\n')[1:]:
    m = re.match(r'(\S+)\s*:\s*(\d+)\s+inst\s+([\d.]+)%\s+thr/inst\s+([\d.]+)\s+samples\s+([\d.]+)%', ln)
    if m: rows.append((m.group(1), int(m.group(2)), float(m.group(3)), float(m.group(4)), float(m.group(5))))
rng = []
for a in sys.argv[2:]:
    spec, name = a.split('=')
    f, r = spec.split(':'); lo, hi = r.split('-')
    rng.append((f, int(lo), int(hi), name))
agg = defaultdict(lambda: [0, 0, 0])
for f, l, i, t, s in rows:
    name = 'other:' + f
    for (ff, lo, hi, nm) in rng:
        if f.startswith(ff) and lo <= l <= hi: name = nm; break
    agg[name][0] += i; agg[name][1] += i * t; agg[name][2] += s
for c, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print('%-34s inst %5.1f%%  thr %5.1f  samples %5.1f%%' % (c, v[0], v[1] / max(v[0], 1e-9), v[2]))
