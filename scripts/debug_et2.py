import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pylabfea_b200 as FE
mat3 = FE.Material()
mat3.elasticity(E=300.e3, nu=0.3)
mat3.plasticity(sy=150., khard=2000.)
size=2; eps=0.1; u=eps*size
fe = FE.Model(dim=2, planestress=True)
fe.geom([size], LY=size); fe.assign([mat3]); fe.bcleft(0.); fe.bcbot(0.)
fe.bcright(0.4*u,'disp'); fe.bctop(0.4*u,'disp'); fe.mesh(NX=2,NY=2)
try:
    fe.solve(verb=True)
except Exception as e:
    print('EXC', e)
print('solver stats', fe.solver_stats)
for s in fe.sweep_stats: print(s)
print(fe.sgl)
