#!/usr/bin/env python
"""Where the stall samples of an .ncu-rep sit: top SASS instructions by long-scoreboard / short-scoreboard / mio / wait samples,
with the CUDA source line of each (needs -lineinfo and --import-source on).   python tools/ncu_stalls.py REP [TOPN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr = None; out = []
for r in rows:
    if r and r[0] == 'Address':
        hdr = r; continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        d = {h: r[i] for i, h in enumerate(hdr)}
        out.append((d['Source'].strip(), int(float(d['Instructions Executed'])), int(float(d['# Samples'])),
                    {k: int(float(d[k] or 0)) for k in ('stall_long_sb', 'stall_short_sb', 'stall_mio', 'stall_wait', 'stall_lg', 'stall_not_selected', 'stall_math', 'stall_branch_resolving', 'stall_dispatch', 'stall_no_inst')}))
    except ValueError:
        continue
tot = sum(o[2] for o in out)
print('samples', tot)
for key in ('stall_long_sb', 'stall_short_sb', 'stall_mio', 'stall_wait'):
    s = sum(o[3][key] for o in out)
    print(f'--- {key}: {100 * s / tot:.1f}% of samples')
    for i in sorted(range(len(out)), key=lambda i: -out[i][3][key])[:topn]:
        o = out[i]
        prev = out[i - 1][0] if i else ''
        print(f'   {100 * o[3][key] / tot:5.2f}%  exec {o[1]:9d}  {o[0][:70]:70s} | prev: {prev[:50]}')
if len(sys.argv) > 3:
    # bucket report: samples by executed-count bucket (loop nest level) and stall reason
    import collections
    keys = ('stall_long_sb', 'stall_short_sb', 'stall_mio', 'stall_wait', 'stall_lg', 'stall_not_selected', 'stall_math', 'stall_branch_resolving', 'stall_dispatch', 'stall_no_inst')
    edges = [float(v) for v in sys.argv[3].split(',')]
    b = collections.defaultdict(lambda: collections.Counter())
    for src, ex, smp, st in out:
        k = sum(ex > e for e in edges)
        b[k]['samples'] += smp; b[k]['inst'] += ex; b[k]['n'] += 1
        for key in keys:
            b[k][key] += st[key]
    for k in sorted(b):
        c = b[k]
        print(f'bucket {k} (exec > {edges[k - 1] if k else 0:g}): {c["n"]} sass, inst {c["inst"] / 1e6:.0f} M, samples {100 * c["samples"] / tot:.1f}% :: ' +
              ' '.join(f'{key[6:]} {100 * c[key] / tot:.1f}' for key in keys))
