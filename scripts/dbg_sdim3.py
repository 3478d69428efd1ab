import sys, os
ROOT = os.environ.get('GRAFT_REPO_ROOT', '/root/repo')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from pylabfea_b200 import _lib
from helpers import plf_material_from_golden
g = dict(np.load(os.path.join(ROOT, 'tests/golden/response_j2s3.npz')))
ctx = _lib.Context(0)
m = plf_material_from_golden(g); m.kind = _lib.MAT_HILL3
ctx.set_materials([m], planestress=False)
sig, epl, deps = g['pe_sig'], g['pe_epl'], g['pe_deps']
fy, so, dp, ct, ns = ctx.response_batch(sig, epl, deps)
bad = []
for i in range(len(fy)):
    e_ct = np.max(np.abs(ct[i] - g['pe_ctan'][i])) / np.max(np.abs(g['pe_ctan'][i]))
    e_s = np.max(np.abs(so[i] - g['pe_sig_out'][i])) / max(np.max(np.abs(g['pe_sig_out'][i])), 1e-300)
    if e_ct > 1e-6 or e_s > 1e-6:
        bad.append((i, int(ns[i]), int(g['pe_nsteps'][i]), float(e_ct), float(e_s)))
print('bad tuples', len(bad), bad[:10])

import collections
print(collections.Counter((b[1], b[2]) for b in bad))
ok50 = [i for i in range(len(fy)) if ns[i] == 49 and i not in [b[0] for b in bad]]
print('50-step tuples that pass:', len(ok50), 'fail:', len([b for b in bad if b[1] == 49]))
# which failing tuples have fy far above the tolerance at the end (corrections)?
for b in bad[:6]:
    i = b[0]
    print(i, 'fy dev', fy[i], 'ref', g['pe_fy'][i], 'depl diff', np.max(np.abs(dp[i] - g['pe_depl'][i])), 'ct err', b[3])


for dbg in (0, 1, 2, 3):
    md = plf_material_from_golden(g); md.kind = _lib.MAT_HILL3; md.dev_only = dbg
    ctx.set_materials([md], planestress=False)
    fyd, sod, dpd, ctd, nsd = ctx.response_batch(sig, epl, deps)
    ed = [np.max(np.abs(ctd[i] - g['pe_ctan'][i])) / np.max(np.abs(g['pe_ctan'][i])) for i in range(len(fy))]
    print('debug', dbg, 'bad', int(np.sum(np.array(ed) > 1e-6)), 'max', max(ed))
