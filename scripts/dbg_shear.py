import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+'/tests')
import numpy as np, pylabfea_b200 as FE
g = dict(np.load(R+'/tests/golden/solve_shear_hill3_4x4.npz'))
m = FE.Material(name='hill3s'); m.elasticity(E=200.e3, nu=0.3); m.plasticity(sy=150., hill=[0.7, 1., 1.4], khard=100., sdim=3)
fe = FE.Model(dim=2, planestress=False); fe.geom([4.], LY=4.); fe.assign([m])
fe.bcleft(0., "force"); fe.bcright(0., "force"); fe.bcbot(0., "disp", "x"); fe.bcbot(0., "disp", "y"); fe.bctop(0.006 * fe.leny, "disp", "x"); fe.bctop(0.008 * fe.leny, "disp", "y")
fe.mesh(NX=4, NY=4); fe.solve(min_step=4)
print('nsteps', fe.nsteps, g['nsteps'], fe.niter, list(g['niter']))
for i in range(len(fe.sgl)):
    print(i, 'sgl', np.max(np.abs(fe.sgl[i]-g['sgl'][i])), 'egl', np.max(np.abs(fe.egl[i]-g['egl'][i])), 'epgl', np.max(np.abs(fe.epgl[i]-g['epgl'][i])))
print('sweep stats', fe.sweep_stats[:6])
