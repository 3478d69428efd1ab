#!/usr/bin/env python3
"""Strip-partitioned linear solve over several GPUs, checked against the single-GPU solve of the same global system.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 \
        scripts/dist_solve.py [NX per rank = 512] [NY = 512]

(the launcher is only used to start one process per GPU with RANK / LOCAL_RANK / WORLD_SIZE set; nothing here imports
torch.)  One (world * NX) x NY mesh, examples/inclusion.py pattern (stiff matrix, soft inclusion in the middle third),
uniaxial-y displacement control.  Every rank
  1. builds the global boundary-condition vectors on the host (Model._calc_BC) and takes the slice of its node grid,
  2. assembles K of its strip (owned element columns + the halo column, pylabfea_b200/shard.py),
  3. joins the NCCL communicator and calls plf_solve_linear -> distributed Jacobi-preconditioned CG,
  4. solves the whole system again alone on its own GPU (same preconditioner) and compares du on its node grid.
Writes gpurun_out/dist_solve_rank<r>.json."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import pylabfea_b200 as FE  # noqa: E402
from pylabfea_b200 import _lib  # noqa: E402
from pylabfea_b200.shard import strip_shard, slice_node_vector, owned_slice, share_bytes  # noqa: E402


def main():
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    NXper = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    NY = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    rtol = float(os.environ.get('PLF_RTOL', '1e-10'))
    NXg = world * NXper

    # ---- the global model on the host (mesh numbering and calc_BC need no device) -------------------------------
    el = np.ones((NXg, NY))
    el[NXg // 3:2 * NXg // 3, NY // 3:2 * NY // 3] = 2
    mat1 = FE.Material(num=1)
    mat1.elasticity(E=100.e3, nu=0.27)
    mat2 = FE.Material(num=2)
    mat2.elasticity(E=3.e3, nu=0.3)
    fe = FE.Model(dim=2, planestress=False)
    fe.geom(sect=2, LX=4. * world, LY=4.)
    fe.assign([mat1, mat2])
    fe.bcbot(0.)
    fe.bcright(0., 'force')
    fe.bcleft(0., 'force')
    fe.bctop(0.004 * fe.leny, 'disp')
    fe.mesh(elmts=el, NX=NXg, NY=NY)
    fe.bcnode(0, 0., 'disp', 'x')
    fixed, du_bc, w, fext = fe._calc_BC(fe.bcl, fe.bcb, fe.bcr, fe.bct, fe.bcn if fe.noset is not None else None)
    descr = [m._descriptor() for m in fe.mat]
    elem_mat = np.asarray(fe._elem_mat, dtype=np.int32).reshape(NXg, NY)

    # ---- this rank's strip --------------------------------------------------------------------------------------
    s = strip_shard(NXg, rank, world)
    first, nx = s['first_col_global'], s['nx_local']
    ctx = _lib.Context(local)
    ctx.set_materials(descr, planestress=False)
    ctx.mesh_structured(nx, NY, np.asarray(fe._dx)[first:first + nx], fe._dy, fe.thick, elem_mat[first:first + nx].ravel())
    ctx.setupK()
    token = '%s_%d' % (os.environ.get('MASTER_PORT', '0'), os.getppid())
    uid = share_bytes('/tmp/plf_nccl_id_' + token, rank, _lib.comm_unique_id() if rank == 0 else None)
    ctx.comm_init(uid, rank, world)
    cols = s['owned_node_cols_local']
    ctx.comm_set_strip(cols[0], cols[-1] + 1)
    red = ctx.comm_allreduce([rank + 1., 1.])
    assert red[0] == world * (world + 1) / 2 and red[1] == world, red

    def loc(v):
        return np.ascontiguousarray(slice_node_vector(v, s, NY))
    ctx.set_option('solver_warm', 0)
    t0 = time.perf_counter()
    it, rr = ctx.solve_linear(loc(fixed), loc(du_bc), loc(w), loc(fext), rtol=rtol)
    t1 = time.perf_counter()
    it2, rr2 = ctx.solve_linear(None, None, None, None, rtol=rtol)   # boundary-condition vectors already on the device
    t2 = time.perf_counter()
    du_loc = ctx.get_field(_lib.F_DU).ravel()

    # ---- the same system solved by this GPU alone ---------------------------------------------------------------
    g = _lib.Context(local)
    g.set_materials(descr, planestress=False)
    g.mesh_structured(NXg, NY, fe._dx, fe._dy, fe.thick, elem_mat.ravel())
    g.setupK()
    g.set_option('solver_warm', 0)
    g.set_option('solver_mg', 0)
    g.solve_linear(fixed, du_bc, w, fext, rtol=rtol)
    t3 = time.perf_counter()
    it_g, rr_g = g.solve_linear(None, None, None, None, rtol=rtol)
    t4 = time.perf_counter()
    du_g = loc(g.get_field(_lib.F_DU).ravel())
    g.set_option('solver_mg', 1)
    g.solve_linear(None, None, None, None, rtol=rtol)
    t5 = time.perf_counter()
    it_mg, _ = g.solve_linear(None, None, None, None, rtol=rtol)
    t6 = time.perf_counter()

    scale = float(np.max(np.abs(du_g)))
    own = owned_slice(s, NY)
    out = {'rank': rank, 'world': world, 'mesh_global': [NXg, NY], 'mesh_local': [nx, NY], 'halo': s['halo'], 'rtol': rtol,
           'dist_iters': it, 'dist_relres': rr, 'dist_first_call_s': t1 - t0, 'dist_solve_s': t2 - t1, 'dist_iters_2nd': it2,
           'single_gpu_jacobi_iters': it_g, 'single_gpu_jacobi_s': t4 - t3, 'single_gpu_mg_iters': it_mg, 'single_gpu_mg_s': t6 - t5,
           'max_diff_owned_rel': float(np.max(np.abs(du_loc[own] - du_g[own])) / scale),
           'max_diff_local_grid_rel': float(np.max(np.abs(du_loc - du_g)) / scale), 'du_max': scale,
           'launches': ctx.launch_count()}
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/dist_solve_rank%d.json' % rank, 'w') as f:
        json.dump(out, f)
    print(json.dumps(out), flush=True)
    ok = out['max_diff_local_grid_rel'] < 1e-5 and abs(it - it_g) <= max(4, it_g // 50)
    ctx.comm_destroy()
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
