import sys, os
sys.path.insert(0, os.environ.get('GRAFT_REPO_ROOT', '/root/repo'))
import numpy as np
import pylabfea_b200 as FE
g = dict(np.load(os.path.join(os.environ.get('GRAFT_REPO_ROOT', '/root/repo'), 'tests/golden/solve_fem2_sdim3.npz')))
mat1 = FE.Material(); mat1.elasticity(E=100.e3, nu=0.35)
mat2 = FE.Material(); mat2.elasticity(E=300.e3, nu=0.3); mat2.plasticity(sy=150., khard=500., sdim=3)
fem2 = FE.Model(dim=2, planestress=False)
fem2.geom([2, 2], LY=4.); fem2.assign([mat1, mat2]); fem2.bcleft(0.); fem2.bcbot(0.); fem2.bcright(0., 'force'); fem2.bctop(0.1 * fem2.leny, 'disp')
fem2.mesh(NX=4, NY=4); fem2.solve(); fem2.calc_global()
print('niter', list(fem2.niter), 'ref', list(g['niter']))
print('glob epl', fem2.glob['epl'], 'ref', g['glob_epl'])
print('glob sig', fem2.glob['sig'], 'ref', g['glob_sig'])
sig = np.array([el.sig for el in fem2.element])
print('max sig diff', np.max(np.abs(sig - g['sig'])), 'max |sig|', np.max(np.abs(g['sig'])))
print('max shear in ref state', np.max(np.abs(g['sig'][:, 5])), 'u diff', np.max(np.abs(fem2.u - g['u'])) / np.max(np.abs(g['u'])))
