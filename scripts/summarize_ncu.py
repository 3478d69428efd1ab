#!/usr/bin/env python3
"""Turn an ncu report (gpurun_out/*.ncu-rep, --set full) into the tracked summaries under profiles/:

    python scripts/summarize_ncu.py gpurun_out/r1_final.ncu-rep profiles/r1_final

writes <out>_summary.txt (per-kernel metrics + top stall reasons) and updates profiles/traffic.json, the per-launch
DRAM traffic that bench.py reports as roofline.traffic.  Runs in the build container (ncu -i needs no GPU)."""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']


def to_bytes(val, unit):
    scale = {'byte': 1., 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    return float(val) * scale.get(unit, 1.)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, units = rows[0], rows[1]
    stall_cols = [i for i, c in enumerate(h) if c.startswith('smsp__pcsamp_warps_issue_stalled') and not c.endswith('not_issued')]
    traffic = {}
    with open(out + '_summary.txt', 'w') as f:
        f.write('# ncu --set full --clock-control none, source report: %s\n' % os.path.basename(rep))
        for r in rows[2:]:
            name = r[h.index('Kernel Name')]
            short = name.split('(')[0].replace('void ', '')
            f.write('\n== %s\n' % name[:150])
            for k in KEYS:
                if k in h:
                    i = h.index(k)
                    f.write('   %-62s %s %s\n' % (k, r[i], units[i]))
            st = sorted([(float(r[i]), h[i].replace('smsp__pcsamp_warps_issue_stalled_', '')) for i in stall_cols if r[i]], reverse=True)[:6]
            f.write('   top stalls (pc samples): %s\n' % ', '.join('%s=%d' % (n, v) for v, n in st))
            ir, iw = h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
            b = to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw])
            traffic.setdefault(short, []).append(b)
    tj = os.path.join(os.path.dirname(out) or '.', 'traffic.json')
    cur = json.load(open(tj)) if os.path.exists(tj) else {}
    cur.update({k: {'dram_bytes_per_launch': sum(v) / len(v), 'launches_profiled': len(v), 'report': os.path.basename(rep)}
                for k, v in traffic.items()})
    json.dump(cur, open(tj, 'w'), indent=1, sort_keys=True)
    print(open(out + '_summary.txt').read())


if __name__ == '__main__':
    main()
