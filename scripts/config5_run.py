#!/usr/bin/env python3
"""BASELINE config 5: 2048 x 2048 mesh, Barlat Yld2004-18p plasticity (examples/train_goss_barlat.py parameters, a = 8),
element columns sharded over the ranks, one process per GPU, NCCL for the exchange step of the solve.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        scripts/config5_run.py [N=2048] [min_step=3] [eps=0.0015]

Model.distribute() cuts the mesh into strips; every rank updates the elements of its strip (+ one halo column),
assembles its K rows, and the ranks solve together (CG: one halo node column per neighbour and iteration, all-reduced
dot products; preconditioner = one multigrid V-cycle per strip).  Rank 0 prints one JSON line.  Works with 1 rank too."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import pylabfea_b200 as FE  # noqa: E402
from pylabfea_b200 import _lib  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    min_step = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    eps = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0015
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', '0'), ('WORLD_SIZE', '1'), ('LOCAL_RANK', '0')))
    g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'barlat_response.npz')))
    mat = FE.Material('Goss-Barlat', num=1)
    mat.elasticity(E=float(g['E']), nu=float(g['nu']))
    mat.plasticity(sy=float(g['sy']), khard=float(os.environ.get('PLF_KHARD', '300.')), barlat=list(g['bp']),
                   barlat_exp=int(g['barlat_exp']), sdim=6)
    soft = FE.Material('inclusion', num=2)
    soft.elasticity(E=3.e3, nu=0.3)
    n1 = N // 3
    el = np.ones((N, N))
    el[n1:2 * n1, n1:2 * n1] = 2
    fe = FE.Model(dim=2, planestress=False, device=local)
    fe.geom(sect=2, LX=4., LY=4.)
    fe.assign([mat, soft])
    fe.bcbot(0.)
    fe.bcright(0., 'force')
    fe.bcleft(0., 'force')
    fe.bctop(eps * fe.leny, 'disp')
    fe.mesh(elmts=el, NX=N, NY=N)
    fe.bcnode(0, 0., 'disp', 'x')
    fe.solver_rtol = 1.e-10
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        t = torch.zeros(_lib.COMM_ID_BYTES, dtype=torch.uint8, device='cuda')
        if rank == 0:
            t = torch.tensor(list(_lib.comm_unique_id()), dtype=torch.uint8, device='cuda')
        dist.broadcast(t, src=0)
        fe.distribute(rank, world, bytes(t.cpu().tolist()))
    ctx = fe._dev()
    if 'PLF_HYBRID' in os.environ:   # 1: level 0 of the V-cycle and the CG vectors shared among the ranks (option solver_hybrid)
        ctx.set_option('solver_hybrid', int(os.environ['PLF_HYBRID']))
    ctx.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    fe.solve(min_step=min_step)
    ctx.synchronize()
    if dist:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        its = int(sum(s[0] for s in fe.solver_stats))
        upd = float(sum(sum(s[5][:3]) for s in fe.sweep_stats))
        print(json.dumps({'config': 'c5', 'mesh': [N, N], 'material': 'Barlat Yld2004-18p a=8 + elastic inclusion', 'n_gpus': world,
                          'solver_hybrid': os.environ.get('PLF_HYBRID', 'default (2: shared when p2p)'), 'p2p': bool(world > 1 and ctx.comm_p2p()),
                          'seconds': dt, 'increments': int(fe.nsteps), 'inner_passes': [int(x) for x in fe.niter],
                          'linear_solves': len(fe.solver_stats), 'cg_iterations': its, 'ms_per_cg_iteration': dt * 1e3 / max(its, 1),
                          'element_updates': upd, 'element_updates_per_s': upd / dt,
                          'branches_last_pass': fe.sweep_stats[-1][5] if fe.sweep_stats else None,
                          'global_stress': [float(x) for x in fe.glob['sig']], 'global_plastic_strain': [float(x) for x in fe.glob['epl']]}))
    if dist:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
