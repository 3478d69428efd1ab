"""GPU check: every pass A / pass B variant gives the same results as the staged kernels with the general fast mode
(1024^2 planar field, plus a non-planar batch that exercises the back of the two-ended work list)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from pylabfea_b200 import _lib
ctx = _lib.Context(0)
mat, CV = bench.j2_material()
ctx.set_materials([mat], planestress=False)
def run(sig, epl, deps, **opts):
    base = dict(compact=1, passa_direct=0, max_fm=2, passb_recur=0)
    base.update(opts)
    for k, v in base.items():
        ctx.set_option(k, v)
    return ctx.response_batch(sig, epl, deps)
sig, epl, deps = bench.synth_states(1024, 1234)
sig3, epl3, deps3 = (a[:300000].copy() for a in (sig, epl, deps))
rng = np.random.default_rng(1)
odd = rng.random(len(sig3)) < 0.3                       # 30 % of the elements get out-of-plane shear
sig3[odd, 3] = 20. * rng.standard_normal(odd.sum()); sig3[odd, 4] = 10. * rng.standard_normal(odd.sum())
deps3[odd, 3] = 1e-4 * rng.standard_normal(odd.sum())
for name, (a, b, c) in (('planar', (sig, epl, deps)), ('mixed', (sig3, epl3, deps3))):
    ref = run(a, b, c)
    print(name, 'n50 =', int(np.sum(ref[4] == 49)))
    for opts in (dict(passa_direct=1), dict(max_fm=3), dict(max_fm=4), dict(max_fm=4, passa_direct=1), dict(max_fm=4), dict(max_fm=4, passb_recur=1)):
        out = run(a, b, c, **opts)
        ok = np.array_equal(out[4], ref[4])
        err = max(float(np.max(np.abs(x - y)) / np.max(np.abs(y))) for x, y in zip(out[:4], ref[:4]))
        print('  ', opts, 'nsteps equal', ok, 'max rel diff %.2e' % err)
        assert ok and err < 1e-9
print('variants ok')
