"""world_size-N CPU rehearsal of the strip-partitioned CG (launched by tests/test_dist_gloo.py with torchrun, gloo).

Every rank builds the stiffness of ITS strip (owned element columns + halo column, pylabfea_b200/shard.py) with the
oracle's element integration and assembly, takes its slice of the global calc_BC vectors and runs the same sequence of
steps as launch_cg_dist (pylabfea_b200/csrc/model_kernels.cu) -- ghost | fixed mask, halo exchange of p with the
neighbours (send/recv of one node column), q = K p on the owned rows, all-reduced p.q, update, all-reduced (r.z, r.r),
direction -- in numpy, with torch.distributed (gloo) standing in for NCCL.  Each rank then checks du on its local node
grid against the dense solve of the global system.  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import pylabfea_b200 as FE  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from pylabfea_b200.shard import strip_shard, slice_node_vector, owned_slice  # noqa: E402


def strip_stiffness(fe, first, nx, NY, E_by_mat, nu):
    """Dense K of the local structured mesh of element columns [first, first + nx) (local node / element numbering)."""
    nodes = []
    for j in range(nx):
        for k in range(NY):
            n1 = j * (NY + 1) + k
            nodes.append([n1, n1 + 1, n1 + NY + 1, n1 + NY + 2])
    nodes = np.array(nodes, dtype=np.int32)
    Kel = np.zeros((len(nodes), 8, 8))
    emat = np.asarray(fe._elem_mat).reshape(fe._NX, NY)[first:first + nx].ravel()
    dx = np.asarray(fe._dx)[first:first + nx]
    for e in range(len(nodes)):
        m = fe.mat[emat[e]]
        lx, ly = dx[e // NY], fe._dy
        CV = orc.element_CV(m.CV, m.E, m.nu, m.C44, False)
        B4 = orc.gauss_bmats(lx, ly, 0, CV, m.E, m.nu)
        Kel[e] = orc.kel(B4, CV, (lx * ly * fe.thick) * 4.)
    ndof = 2 * (nx + 1) * (NY + 1)
    return orc.setupK_dense(nodes, Kel, ndof)


def main():
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    NXg, NY = 4 * world + 1, 5            # uneven strips
    el = np.ones((NXg, NY))
    el[NXg // 3:2 * NXg // 3, 1:4] = 2
    mat1, mat2 = FE.Material(num=1), FE.Material(num=2)
    mat1.elasticity(E=100.e3, nu=0.27)
    mat2.elasticity(E=3.e3, nu=0.3)
    fe = FE.Model(dim=2, planestress=False)
    fe.geom(sect=2, LX=4., LY=2.)
    fe.assign([mat1, mat2])
    fe.bcbot(0.)
    fe.bcright(10., 'force')             # a force boundary: keeps its global normalisation in the slices
    fe.bcleft(0.)
    fe.bctop(0.004 * fe.leny, 'disp')
    fe.mesh(elmts=el, NX=NXg, NY=NY)
    fixed, du_bc, w, fext = fe._calc_BC(fe.bcl, fe.bcb, fe.bcr, fe.bct, None)

    s = strip_shard(NXg, rank, world)
    first, nx = s['first_col_global'], s['nx_local']
    K = strip_stiffness(fe, first, nx, NY, None, None)
    ncol = 2 * (NY + 1)
    cols = s['owned_node_cols_local']
    own_first, own_end = cols[0], cols[-1] + 1
    fx, dub, wl, fl = (np.array(slice_node_vector(v, s, NY)) for v in (fixed, du_bc, w, fext))
    ghost = np.ones(len(fx), dtype=bool)
    ghost[own_first * ncol:own_end * ncol] = False
    mask = ghost | (fx != 0)

    def allreduce(*vals):
        t = torch.tensor(vals, dtype=torch.float64)
        dist.all_reduce(t)
        return t.tolist()

    def halo(v):
        ops, bufs = [], []
        if own_first > 0:      # left neighbour: send my first owned column, receive its last owned column into column 0
            send = torch.from_numpy(v[own_first * ncol:(own_first + 1) * ncol].copy())
            recv = torch.empty(ncol, dtype=torch.float64)
            ops += [dist.P2POp(dist.isend, send, rank - 1), dist.P2POp(dist.irecv, recv, rank - 1)]
            bufs.append((recv, (own_first - 1) * ncol))
        if own_end < nx + 1:   # right neighbour
            send = torch.from_numpy(v[(own_end - 1) * ncol:own_end * ncol].copy())
            recv = torch.empty(ncol, dtype=torch.float64)
            ops += [dist.P2POp(dist.isend, send, rank + 1), dist.P2POp(dist.irecv, recv, rank + 1)]
            bufs.append((recv, own_end * ncol))
        if ops:
            for r in dist.batch_isend_irecv(ops):
                r.wait()
        for recv, off in bufs:
            v[off:off + ncol] = recv.numpy()

    diag = np.diag(K).copy()
    b = np.where(mask, 0., fl - K @ wl)
    x = np.zeros_like(b)
    r = b.copy()
    z = np.where(mask, 0., r / diag)
    p = z.copy()
    rz, rr = allreduce(float(r @ z), float(r @ r))
    target = (1e-13 ** 2) * rr
    iters = 0
    while rr > target and iters < 5000:
        halo(p)
        q = np.where(mask, 0., K @ p)
        pq, = allreduce(float(p @ q))
        alpha = rz / pq
        x = np.where(mask, x, x + alpha * p)
        r = np.where(mask, r, r - alpha * q)
        z = np.where(mask, 0., r / diag)
        rz_new, rr = allreduce(float(r @ z), float(r @ r))
        p = z + (rz_new / rz) * p
        rz = rz_new
        iters += 1
    du = np.where(fx != 0, dub, x)
    halo(du)

    # the global system, dense, on every rank
    Kg = strip_stiffness(fe, 0, NXg, NY, None, None)
    free = np.nonzero(fixed == 0)[0]
    dug = np.array(du_bc, dtype=float)
    dug[free] = np.linalg.solve(Kg[np.ix_(free, free)], (fext - Kg @ w)[free])
    ref = np.array(slice_node_vector(dug, s, NY))
    err = float(np.max(np.abs(du - ref)) / np.max(np.abs(dug)))
    own = owned_slice(s, NY)
    # owned rows of the strip stiffness are the global rows (what makes the solve communication-free in K)
    gl = slice(first * ncol, (first + nx + 1) * ncol)
    rows_equal = bool(np.array_equal(K[own], Kg[gl, gl][own]))
    # The replicated solve (plf_comm_set_replica): every rank contributes the rows of the node columns it owns, the ranges
    # of all ranks tile the global matrix exactly once, every rank solves the gathered system and keeps its strip of du.
    mine = {'rows': (first * ncol + own.start, first * ncol + own.stop), 'K': K[own]}
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    Kr = np.full_like(Kg, np.nan)
    covered = np.zeros(len(Kg), dtype=int)
    for r_, pt in enumerate(parts):
        sr = strip_shard(NXg, r_, world)
        a, b2 = pt['rows']
        covered[a:b2] += 1
        Kr[a:b2, sr['first_col_global'] * ncol:(sr['first_col_global'] + sr['nx_local'] + 1) * ncol] = pt['K']
    Kr = np.nan_to_num(Kr, nan=0.)        # columns outside a strip's node grid: structurally zero in those rows
    gathered_equal = bool(np.all(covered == 1) and np.array_equal(Kr, Kg))
    dur = np.array(du_bc, dtype=float)
    dur[free] = np.linalg.solve(Kr[np.ix_(free, free)], (fext - Kr @ w)[free])
    err_rep = float(np.max(np.abs(np.array(slice_node_vector(dur, s, NY)) - ref)) / np.max(np.abs(dug)))
    t = torch.tensor([err, float(rows_equal), float(iters), float(gathered_equal), err_rep], dtype=torch.float64)
    gathered = [torch.zeros(5, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, t)
    if rank == 0:
        print(json.dumps({'world': world, 'mesh': [NXg, NY], 'iters': [int(g[2]) for g in gathered],
                          'max_rel_err': max(float(g[0]) for g in gathered),
                          'owned_rows_equal_global': all(bool(g[1]) for g in gathered),
                          'gathered_rows_equal_global': all(bool(g[3]) for g in gathered),
                          'replicated_max_rel_err': max(float(g[4]) for g in gathered)}))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
