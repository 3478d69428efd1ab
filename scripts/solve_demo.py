#!/usr/bin/env python3
"""End-to-end Model.solve on the device at sizes the reference cannot allocate (dense K: 140 GB at 256 x 256).

    python scripts/solve_demo.py [N] [min_step] [eps]

examples/inclusion.py pattern (J2 matrix with hardening, soft elastic inclusion in the middle third, uniaxial-y
displacement control, corner pin).  Prints per-increment statistics: inner passes, CG iterations, branch census."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import pylabfea_b200 as FE  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    min_step = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    eps = float(sys.argv[3]) if len(sys.argv) > 3 else 0.004
    n1 = N // 3
    el = np.ones((N, N))
    el[n1:2 * n1, n1:2 * n1] = 2
    mat1 = FE.Material(num=1)
    mat1.elasticity(E=100.e3, nu=0.27)
    mat1.plasticity(sy=150., khard=500., sdim=6)
    mat2 = FE.Material(num=2)
    mat2.elasticity(E=3.e3, nu=0.3)
    fe = FE.Model(dim=2, planestress=False)
    fe.geom(sect=2, LX=4., LY=4.)
    fe.assign([mat1, mat2])
    fe.bcbot(0.)
    fe.bcright(0., 'force')
    fe.bcleft(0., 'force')
    fe.bctop(eps * fe.leny, 'disp')
    t0 = time.time()
    fe.mesh(elmts=el, NX=N, NY=N)
    fe.bcnode(0, 0., 'disp', 'x')
    fe.solver_rtol = float(os.environ.get('PLF_RTOL', '1e-10'))
    fe._dev().set_option('solver_mg', int(os.environ.get('PLF_MG', '1')))   # 0: Jacobi-preconditioned CG
    for key in ('mg_nu', 'mg_coarse_sweeps', 'mg_omega_permille', 'solver_warm', 'mg_fuse_tail', 'skip_unchanged'):
        if 'PLF_' + key.upper() in os.environ:
            fe._dev().set_option(key, int(os.environ['PLF_' + key.upper()]))
    t1 = time.time()
    # wall-clock breakdown of the load sequence (PLF_BREAKDOWN=1): time spent inside the main device calls and in calc_BC
    acc = {}
    if os.environ.get('PLF_BREAKDOWN'):
        def timed(obj, name):
            f = getattr(obj, name)
            def g(*a, **k):
                t = time.perf_counter()
                r = f(*a, **k)
                acc[name] = acc.get(name, 0.) + time.perf_counter() - t
                return r
            setattr(obj, name, g)
        dev = fe._dev()
        for nm in ('solve_linear', 'sweep', 'setupK', 'commit', 'calc_scf', 'reduce_global', 'get_field', 'set_field'):
            if hasattr(dev, nm):
                timed(dev, nm)
        timed(fe, '_calc_BC')
    fe.solve(min_step=min_step)
    t2 = time.time()
    if acc:
        print('breakdown [s]:', {k: round(v, 3) for k, v in sorted(acc.items(), key=lambda kv: -kv[1])})
    its = np.array([s[0] for s in fe.solver_stats])
    print('N=%d  Nel=%d  Ndof=%d  mesh %.2fs  solve %.2fs' % (N, fe.Nel, fe.Ndof, t1 - t0, t2 - t1))
    print('increments %d, inner passes per increment %s' % (fe.nsteps, fe.niter))
    print('linear solves %d, CG iterations: total %d, mean %.0f, max %d' % (len(its), its.sum(), its.mean(), its.max()))
    for s in fe.sweep_stats[:3] + fe.sweep_stats[-3:]:
        print('  increment %d pass %d: change=%s conv=%s max_dstiff=%.3g branches[elastic,1-step,50-sub,corr]=%s' % s)
    print('global stress', np.round(fe.glob['sig'], 3), 'global plastic strain', np.round(fe.glob['epl'], 6))
    print('launches', fe._dev().launch_count())


if __name__ == '__main__':
    main()
