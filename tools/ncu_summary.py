#!/usr/bin/env python
"""Summarise an .ncu-rep: key metrics per kernel + executed-instruction mix by opcode (reads `ncu -i`).

    python tools/ncu_summary.py REP                       # text summary
    python tools/ncu_summary.py REP --traffic KEY         # also record dram read + write bytes of the first kernel of the report in
                                                          # profiles/crop_traffic.json under KEY (bench.py's roofline.traffic reads it)
"""
import collections, csv, io, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps', 'launch__grid_size',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_bytes.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_fmaheavy.sum', 'sm__inst_executed_pipe_fmalite.sum', 'sm__inst_executed_pipe_xu.sum',
        'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.sum']


def run(args):
    return subprocess.run(['ncu', '-i', *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    traffic_key = sys.argv[sys.argv.index('--traffic') + 1] if '--traffic' in sys.argv else None
    rows = list(csv.reader(io.StringIO(run([rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    if traffic_key and len(rows) > 2:
        import json, os
        r = rows[2]
        def gb(name):
            v, u = float(r[hdr.index(name)]), units[hdr.index(name)]
            return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}[u]
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'crop_traffic.json')
        try:
            db = json.load(open(path))
        except Exception:
            db = {}
        db[traffic_key] = {'dram_bytes': gb('dram__bytes_read.sum') + gb('dram__bytes_write.sum'), 'dram_read_bytes': gb('dram__bytes_read.sum'),
                           'dram_write_bytes': gb('dram__bytes_write.sum'), 'kernel': r[hdr.index('Kernel Name')][:80],
                           'source': f'ncu --set full capture {os.path.basename(rep)} (one launch)'}
        json.dump(db, open(path, 'w'), indent=1)
    for r in rows[2:]:
        print('===', r[hdr.index('Kernel Name')][:70])
        for w in WANT:
            if w in hdr:
                print('  %-72s %s %s' % (w, r[hdr.index(w)], units[hdr.index(w)]))
        for i, h in enumerate(hdr):
            if 'warp_issue_stalled' in h and h.endswith('per_warp_active.pct'):
                try:
                    v = float(r[i])
                except ValueError:
                    continue
                if v > 4:
                    print('  stall %-66s %.1f' % (h.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), v))
    rows = list(csv.reader(io.StringIO(run([rep, '--page', 'source', '--csv']))))
    data = [r for r in rows if len(r) > 10]
    hdr = data[0]
    data = data[1:]
    isrc, iex, ismp = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    num = lambda x: int(float(x)) if x.replace('.', '', 1).isdigit() else 0
    tot = sum(num(r[iex]) for r in data) or 1
    ops, smp = collections.Counter(), collections.Counter()
    for r in data:
        toks = r[isrc].split()
        if not toks:
            continue
        op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
        parts = op.split('.')
        key = parts[0] + ('.' + parts[1] if len(parts) > 1 and parts[0] in ('LDS', 'STG', 'LDG', 'STS', 'I2FP', 'F2I', 'I2F', 'LDGSTS') else '')
        ops[key] += num(r[iex]); smp[key] += num(r[ismp])
    print('--- instruction mix (warp instructions executed: %d)' % tot)
    for op, c in ops.most_common(28):
        print('  %-14s %6.2f%%   stall samples %d' % (op, 100 * c / tot, smp[op]))


if __name__ == '__main__':
    main()
