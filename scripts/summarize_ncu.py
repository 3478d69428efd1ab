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
    # optional: --fp64 <config name> <elements>  -> FP64 flop per element update of that BASELINE config into profiles/fp64_ops.json
    fp64_cfg = None
    if '--fp64' in sys.argv:
        i = sys.argv.index('--fp64')
        fp64_cfg = (sys.argv[i + 1], float(sys.argv[i + 2]))
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
    if fp64_cfg:
        ops = {'dfma': 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'dadd': 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
               'dmul': 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum'}
        flop, seen = 0., []
        for r in rows[2:]:
            name = r[h.index('Kernel Name')].split('(')[0].replace('void ', '')
            if not name.startswith('response_') and not name.startswith('sweep_'):
                continue
            if name in seen:      # one launch per kernel variant
                continue
            seen.append(name)
            for k, col in ops.items():
                if col in h:
                    flop += float(r[h.index(col)].replace(',', '')) * (2. if k == 'dfma' else 1.)
        fj = os.path.join(os.path.dirname(out) or '.', 'fp64_ops.json')
        cur = json.load(open(fj)) if os.path.exists(fj) else {}
        cur[fp64_cfg[0]] = {'flop_per_update': flop / fp64_cfg[1], 'kernels': seen, 'elements': fp64_cfg[1],
                            'source': 'ncu thread-level DFMA x 2 + DADD + DMUL of ' + os.path.basename(rep)}
        json.dump(cur, open(fj, 'w'), indent=1, sort_keys=True)
    tj = os.path.join(os.path.dirname(out) or '.', 'traffic.json')
    cur = json.load(open(tj)) if os.path.exists(tj) else {}
    cur.update({k: {'dram_bytes_per_launch': sum(v) / len(v), 'launches_profiled': len(v), 'report': os.path.basename(rep)}
                for k, v in traffic.items()})
    json.dump(cur, open(tj, 'w'), indent=1, sort_keys=True)
    print(open(out + '_summary.txt').read())


if __name__ == '__main__':
    main()
