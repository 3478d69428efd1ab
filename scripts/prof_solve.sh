# where a load step spends its time: ncu launch list (durations only) of a short Model.solve, aggregated per kernel
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${N:-512}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 6000 --csv --log-file gpurun_out/solve_launches.csv python scripts/solve_demo.py $N 3 0.003 > gpurun_out/solve_prof.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/solve_launches.csv', errors='ignore')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]
ik, iv = h.index('Kernel Name'), h.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.])
for r in rows[hdr + 1:]:
    if len(r) <= iv: continue
    name = r[ik].split('(')[0].replace('void plf::', '')[:60]
    try: v = float(r[iv].replace(',', ''))
    except ValueError: continue
    agg[name][0] += 1; agg[name][1] += v
tot = sum(v[1] for v in agg.values())
unit = rows[hdr + 1][h.index('Metric Unit')] if 'Metric Unit' in h else '?'
print('total', tot, unit, 'over', sum(v[0] for v in agg.values()), 'launches')
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print('%-62s n=%5d  %6.2f %%  avg %.2f' % (k, v[0], 100 * v[1] / tot, v[1] / v[0]))
PY
