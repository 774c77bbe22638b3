#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS dump with nvdisasm line info and aggregate by source line.
usage: ncu_by_line.py <sass.csv> <nvdisasm -g -c output> <mangled kernel name substring> [top]"""
import csv, re, sys
from collections import defaultdict
sass_csv, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
# nvdisasm: offset -> (file, line)
cur = None; inside = False; loc = {}
for ln in open(dis):
    if ln.startswith('\t.section') or ln.startswith('//-----'):
        inside = kname in ln and '.text.' in ln if '.text.' in ln else inside and False
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m: loc[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > 10]
base = int(data[0][ci['Address']], 16)
n = len(loc)
data = data[:n]      # first launch only
agg = defaultdict(lambda: [0.0, 0.0, 0.0])
tot = [0.0, 0.0, 0.0]
for r in data:
    off = int(r[ci['Address']], 16) - base
    key = loc.get(off, ('?', 0))
    v = [float(r[ci['Instructions Executed']] or 0), float(r[ci['Thread Instructions Executed']] or 0), float(r[ci['# Samples']] or 0)]
    for i in range(3): agg[key][i] += v[i]; tot[i] += v[i]
print('total warp-inst %.3g thread-inst %.3g (avg %.1f thr) samples %d over %d SASS instrs' % (tot[0], tot[1], tot[1] / tot[0], tot[2], len(data)))
src = {}
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    if f not in src:
        try: src[f] = open('/root/repo/signer_b200/csrc/' + f).read().split('\n')
        except Exception: src[f] = []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ''
    print('%-12s:%4d  inst %5.1f%%  thr/inst %5.1f  samples %5.1f%%  | %s' % (f, l, 100 * v[0] / tot[0], v[1] / max(v[0], 1), 100 * v[2] / tot[2], text))
