"""world_size-N CPU rehearsal of the V-cycle with SHARED finest levels (launched by tests/test_dist_gloo.py, gloo).

The replicated multi-GPU solve shares the finest levels of its multigrid hierarchy among the ranks (mg_vcycle with MgDist,
mg_plan_sharing in pylabfea_b200/csrc/mg_kernels.cu): on a shared level a rank works only on the rows of the node columns it
owns -- [a_l, b_l) with a_{l+1} = ceil(a_l / 2) -- and swaps one ghost column with each neighbour after every step whose
output the next stencil reads; the right-hand side of the first replicated level is gathered.  This program restates that
schedule in numpy / scipy.sparse with gloo standing in for the peer mailboxes, on a small vector problem of the same shape
(2 dofs per node, 9-point stencil, bilinear transfers, Galerkin coarse operators, damped Jacobi), and checks it against the
plain single-process cycle.  Every entry a rank neither owns nor has received is NaN while it works, so a step that reads a
column the schedule did not provide poisons the result: the test proves that the ranges and the ghost swaps of the schedule
are sufficient, for odd column splits and for a middle rank with two neighbours too.  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stiffness_like(NX, NY, rng):
    """SPD sparse matrix on the (NX+1) x (NY+1) node grid, 2 dofs per node, node-column-major numbering (dof = 2 (j (NY+1) + k)
    + c): random positive element weights, every element couples the 8 dofs of its 4 nodes."""
    n = 2 * (NX + 1) * (NY + 1)
    A = sp.lil_matrix((n, n))
    for j in range(NX):
        for k in range(NY):
            n1 = j * (NY + 1) + k
            nodes = [n1, n1 + 1, n1 + NY + 1, n1 + NY + 2]
            dofs = [2 * m + c for m in nodes for c in (0, 1)]
            w = rng.uniform(0.5, 30.)
            M = rng.standard_normal((8, 8))
            Ke = w * (M @ M.T / 8. + np.eye(8))
            for r_, dr in enumerate(dofs):
                for c_, dc in enumerate(dofs):
                    A[dr, dc] += Ke[r_, c_]
    return A.tocsr()


def prolongation(NXc, NYc):
    """Bilinear interpolation from the (NXc+1) x (NYc+1) grid to the (2 NXc+1) x (2 NYc+1) grid, per component."""
    NXf, NYf = 2 * NXc, 2 * NYc
    rows, cols, vals = [], [], []
    for j in range(NXf + 1):
        js = [(j // 2, 1.)] if j % 2 == 0 else [(j // 2, 0.5), (j // 2 + 1, 0.5)]
        for k in range(NYf + 1):
            ks = [(k // 2, 1.)] if k % 2 == 0 else [(k // 2, 0.5), (k // 2 + 1, 0.5)]
            for J, wj in js:
                for K, wk in ks:
                    for c in (0, 1):
                        rows.append(2 * (j * (NYf + 1) + k) + c)
                        cols.append(2 * (J * (NYc + 1) + K) + c)
                        vals.append(wj * wk)
    return sp.csr_matrix((vals, (rows, cols)), shape=(2 * (NXf + 1) * (NYf + 1), 2 * (NXc + 1) * (NYc + 1)))


class Level:
    pass


def build(NX, NY, nlev, rng):
    lv = []
    A = stiffness_like(NX, NY, rng)
    for l in range(nlev):
        L = Level()
        L.NX, L.NY, L.A = NX, NY, A
        L.dinv = 1. / A.diagonal()
        L.n = A.shape[0]
        lv.append(L)
        if l + 1 < nlev:
            L.P = prolongation(NX // 2, NY // 2)
            L.R = L.P.T.tocsr()
            A = (L.R @ A @ L.P).tocsr()
            NX, NY = NX // 2, NY // 2
    return lv


def plan_sharing(NX, world, rank, nshared):
    """Column ranges of mg_plan_sharing: rank r owns node columns [NX r / world, NX (r + 1) / world) (+ the last column on the
    last rank) on level 0 and [ceil(a / 2), ceil(b / 2)) one level down."""
    def cols(r):
        return NX * r // world, NX * (r + 1) // world + (1 if r == world - 1 else 0)
    a, b = [], []
    x, y = cols(rank)
    for _ in range(nshared):
        a.append(x)
        b.append(y)
        x, y = (x + 1) // 2, (y + 1) // 2
    gather = []
    for r in range(world):
        x, y = cols(r)
        for _ in range(nshared):
            x, y = (x + 1) // 2, (y + 1) // 2
        gather.append((x, y))
    return a, b, gather


def vcycle(lv, r_fine, nu, coarse_sweeps, om, share=None):
    """One V(nu, nu) cycle from a zero initial guess, the schedule of mg_vcycle.  share = (rank, world, a, b, gather): the
    first len(a) levels are shared."""
    nl = len(lv)
    S = len(share[2]) if share else 0
    rank, world = (share[0], share[1]) if share else (0, 1)

    def ncol(L):
        return 2 * (L.NY + 1)

    def rows(L, l, ext=0):
        if l >= S:
            return 0, L.n
        a, b = share[2][l], share[3][l]
        return ncol(L) * max(a - ext, 0), ncol(L) * min(b + ext, L.NX + 1)

    def swap(L, l, v):
        if l >= S:
            return
        a, b, w = share[2][l], share[3][l], ncol(L)
        ops, bufs = [], []
        if rank > 0 and a > 0:
            send = torch.from_numpy(v[w * a:w * (a + 1)].copy())
            recv = torch.empty(w, dtype=torch.float64)
            ops += [dist.P2POp(dist.isend, send, rank - 1), dist.P2POp(dist.irecv, recv, rank - 1)]
            bufs.append((recv, w * (a - 1)))
        if rank + 1 < world and b <= L.NX:
            send = torch.from_numpy(v[w * (b - 1):w * b].copy())
            recv = torch.empty(w, dtype=torch.float64)
            ops += [dist.P2POp(dist.isend, send, rank + 1), dist.P2POp(dist.irecv, recv, rank + 1)]
            bufs.append((recv, w * b))
        if ops:
            for q in dist.batch_isend_irecv(ops):
                q.wait()
        for recv, off in bufs:
            v[off:off + w] = recv.numpy()

    def fresh(L):
        return np.full(L.n, np.nan)

    def jacobi(L, l, x):
        r0, r1 = rows(L, l)
        out = fresh(L)
        if x is None:
            out[r0:r1] = om * L.dinv[r0:r1] * L.b[r0:r1]
        else:
            out[r0:r1] = x[r0:r1] + om * L.dinv[r0:r1] * (L.b[r0:r1] - L.A[r0:r1] @ x)
        return out

    def smooth(L, l, sweeps, from_zero, last_needs_ghosts):
        s0 = 0
        if from_zero:
            L.x = jacobi(L, l, None)
            if sweeps > 1 or last_needs_ghosts:
                swap(L, l, L.x)
            s0 = 1
        for s in range(s0, sweeps):
            L.x = jacobi(L, l, L.x)
            if s + 1 < sweeps or last_needs_ghosts:
                swap(L, l, L.x)

    lv[0].b = r_fine
    for l in range(nl - 1):   # down
        L, C = lv[l], lv[l + 1]
        smooth(L, l, nu, True, True)
        r0, r1 = rows(L, l)
        res = fresh(L)
        res[r0:r1] = L.b[r0:r1] - L.A[r0:r1] @ L.x
        if not os.environ.get('PLF_REHEARSAL_SKIP_RESIDUAL_SWAP'):   # (the test's self-check: without this swap the cycle must fail)
            swap(L, l, res)
        C.b = fresh(C)
        if l < S:
            c0, c1 = ncol(C) * ((share[2][l] + 1) // 2), ncol(C) * ((share[3][l] + 1) // 2)
        else:
            c0, c1 = 0, C.n
        C.b[c0:c1] = L.R[c0:c1] @ res
        if l == S - 1:   # first replicated level: its right-hand side is gathered from all ranks
            mine = np.zeros(C.n)
            g0, g1 = (ncol(C) * v for v in share[4][rank])
            assert (g0, g1) == (c0, c1)
            mine[g0:g1] = C.b[g0:g1]
            t = torch.from_numpy(mine)
            dist.all_reduce(t)
            C.b = t.numpy().copy()
    smooth(lv[-1], nl - 1, coarse_sweeps, True, False)
    for l in range(nl - 2, -1, -1):   # up
        L, C = lv[l], lv[l + 1]
        r0, r1 = rows(L, l, ext=1)   # the ghost columns too: no exchange before the first post-smoothing sweep
        L.x[r0:r1] = L.x[r0:r1] + L.P[r0:r1] @ C.x
        smooth(L, l, nu, False, 1 <= l < S)
    return lv[0].x


def main():
    dist.init_process_group('gloo')
    rank, world = dist.get_rank(), dist.get_world_size()
    results = []
    for NX, NY, nlev in ((40, 8, 4), (48, 16, 4), (56, 8, 3)):   # 3 ranks: odd first columns (13, 37) and a two-sided rank
        rng = np.random.default_rng(11)
        lv = build(NX, NY, nlev, rng)
        r = rng.standard_normal(lv[0].n)
        ref = vcycle(lv, r.copy(), 2, 8, 0.7)
        for nshared in (1, 2, 3):
            if nshared >= nlev:
                continue
            a, b, gather = plan_sharing(NX, world, rank, nshared)
            # every rank must keep >= 2 node columns on the coarsest shared level (mg_plan_sharing stops sharing before that)
            widths = []
            for q in range(world):
                aq, bq, _ = plan_sharing(NX, world, q, nshared)
                widths.append(bq[-1] - aq[-1])
            if min(widths) < 2:
                continue
            z = vcycle(lv, r.copy(), 2, 8, 0.7, share=(rank, world, a, b, gather))
            w = 2 * (NY + 1)
            own = slice(w * a[0], w * b[0])
            err = float(np.max(np.abs(z[own] - ref[own])) / np.max(np.abs(ref)))
            t = torch.tensor([err if np.isfinite(err) else 1e300], dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            results.append({'NX': NX, 'NY': NY, 'levels': nlev, 'shared': nshared, 'a0': a[0], 'max_rel_err': float(t.item())})
    if rank == 0:
        print(json.dumps({'world': world, 'cases': results}))
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
