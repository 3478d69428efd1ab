"""Top stall instructions of one kernel from an ncu report: python scripts/ncu_hot.py report.ncu-rep kernel_regex [N]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# several kernels may follow each other; take the first block
hdr = None
blocks = []
for r in rows:
    if r and r[0] == 'Kernel Name':
        blocks.append({'name': r[1], 'rows': []}); hdr = None
    elif r and r[0] == 'Address':
        hdr = r
        blocks[-1]['hdr'] = r
    elif hdr and len(r) == len(hdr):
        blocks[-1]['rows'].append(r)
b = blocks[0]
h = b['hdr']
iS, iSrc, iEx = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
stall = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
tot = sum(int(r[iS]) for r in b['rows'])
print(b['name'][:120]); print('total samples', tot, 'instructions', len(b['rows']))
agg = {}
for r in b['rows']:
    for i in stall:
        agg[h[i]] = agg.get(h[i], 0) + int(r[i] or 0)
print('stall totals:', sorted(agg.items(), key=lambda kv: -kv[1])[:8])
idx = sorted(range(len(b['rows'])), key=lambda k: -int(b['rows'][k][iS]))[:N]
for k in sorted(idx):
    r = b['rows'][k]
    top = sorted([(int(r[i] or 0), h[i][6:]) for i in stall], reverse=True)[:2]
    print('%5d %5s  %-70s %s' % (k, r[iS], r[iSrc].strip()[:70], ' '.join('%s=%d' % (n, v) for v, n in top if v)))
