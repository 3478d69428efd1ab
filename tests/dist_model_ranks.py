#!/usr/bin/env python3
"""One rank of a strip-partitioned Model.solve (run by tests/test_gpu_model.py and scripts/; one process per GPU):

    RANK=r WORLD_SIZE=w [LOCAL_RANK=r] python tests/dist_model_ranks.py N idfile [replicated|distributed]

Every rank builds the same global inclusion model (N x N elements), calls Model.distribute() and solves; rank 0 then
solves the same model undistributed on its own GPU and compares: increment / pass counts, the global stress-strain
history and the boundary values must agree to 1e-8 (the linear solves of the two runs use different preconditioners and
stop at |r| <= 1e-11 |b|), and its strip of u must match the global u.  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import pylabfea_b200 as FE  # noqa: E402
from pylabfea_b200 import _lib, shard  # noqa: E402


def build(N, device, plastic=True, eps=0.004):
    n1 = N // 3
    el = np.ones((N, N))
    el[n1:2 * n1, n1:2 * n1] = 2
    mat1 = FE.Material(num=1)
    mat1.elasticity(E=100.e3, nu=0.27)
    if plastic:
        mat1.plasticity(sy=150., khard=500., sdim=6)
    mat2 = FE.Material(num=2)
    mat2.elasticity(E=3.e3, nu=0.3)
    fe = FE.Model(dim=2, planestress=False, device=device)
    fe.geom(sect=2, LX=4., LY=4.)
    fe.assign([mat1, mat2])
    fe.bcbot(0.)
    fe.bcright(0., 'force')
    fe.bcleft(0., 'force')
    fe.bctop(eps * fe.leny, 'disp')
    fe.mesh(elmts=el, NX=N, NY=N)
    fe.bcnode(0, 0., 'disp', 'x')
    fe.solver_rtol = 1.e-11
    return fe


def main():
    N = int(sys.argv[1])
    idfile = sys.argv[2]
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    device = int(os.environ.get('LOCAL_RANK', str(rank)))
    uid = shard.share_bytes(idfile, rank, _lib.comm_unique_id() if rank == 0 else None)
    fe = build(N, device)
    solver = sys.argv[3] if len(sys.argv) > 3 else 'replicated'
    fe.distribute(rank, world, uid, solver=solver)
    if 'PLF_HYBRID' in os.environ:      # 0 (default): every rank solves the gathered system alone; 1: level 0 + CG shared among the ranks
        fe._dev().set_option('solver_hybrid', int(os.environ['PLF_HYBRID']))
    fe._dev().synchronize()   # contexts, communicator and peer mailboxes are set up here, not inside the timed solve
    t0 = time.perf_counter()
    fe.solve(min_step=4)
    dt = time.perf_counter() - t0
    out = {'rank': rank, 'world': world, 'N': N, 'solver': solver, 'seconds': dt, 'nsteps': fe.nsteps, 'niter': list(fe.niter),
           'cg_iterations': int(sum(s[0] for s in fe.solver_stats)), 'solves': len(fe.solver_stats),
           'p2p': fe._dev().comm_p2p()}
    if rank == 0:
        ref = build(N, device)
        t0 = time.perf_counter()
        ref.solve(min_step=4)
        out['seconds_single_gpu'] = time.perf_counter() - t0
        out['cg_iterations_single_gpu'] = int(sum(s[0] for s in ref.solver_stats))

        def rel(a, b):
            return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(np.max(np.abs(np.asarray(b))), 1e-300))
        sh = fe.local_info()
        u_loc = shard.slice_node_vector(ref.u, sh, N)
        own = shard.owned_slice(sh, N)
        keys = ('ebc1', 'sbc1', 'ebc2', 'sbc2', 'ebc12', 'sbc12')
        out.update(same_steps=bool(fe.nsteps == ref.nsteps and list(fe.niter) == list(ref.niter)),
                   err_sgl=rel(fe.sgl, ref.sgl), err_egl=rel(fe.egl, ref.egl),
                   err_epgl=float(np.max(np.abs(fe.epgl - ref.epgl)) / max(np.max(np.abs(ref.epgl)), 1e-12)),
                   err_bc=rel([fe.glob[k] for k in keys], [ref.glob[k] for k in keys]),
                   err_u_owned=rel(fe.u[own], u_loc[own]), err_u_local=rel(fe.u, u_loc))
        # replicated: the same K and the same preconditioner (dot products and calc_scf moments summed in another order);
        # distributed: another preconditioner.  All solves stop at |r| <= 1e-11 |b|
        tol = 1e-8 if solver == 'replicated' else 1e-7
        out['ok'] = bool(out['same_steps'] and max(out['err_sgl'], out['err_egl'], out['err_epgl'], out['err_bc'], out['err_u_owned']) <= tol)
        print(json.dumps(out))
    fe._dev().synchronize()


if __name__ == '__main__':
    main()
