#!/usr/bin/env python
"""Executed warp instructions and stall samples per CUDA source line of an .ncu-rep (needs -lineinfo)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = None; hdr = None; agg = {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No' and 'Instructions Executed' in r:
        hdr = r; iex = r.index('Instructions Executed'); ismp = r.index('# Samples'); continue
    if hdr is None or len(r) <= iex or not r[0].isdigit():
        continue
    if r[2] != '-':      # SASS sub-rows repeat the counts of the line
        continue
    try:
        ex = int(float(r[iex])); sm = int(float(r[ismp]))
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1]])
    a[0] += ex; a[1] += sm
tot = sum(a[0] for a in agg.values()) or 1; tsm = sum(a[1] for a in agg.values()) or 1
print('total warp instructions %d, samples %d' % (tot, tsm))
top = sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]
for (f, ln), (ex, sm, src) in sorted(top):
    print('%-13s %4d  inst %5.2f%%  samples %5.2f%%  %s' % (f, ln, 100 * ex / tot, 100 * sm / tsm, src.strip()[:100]))
