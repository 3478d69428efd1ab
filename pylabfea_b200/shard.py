"""Strip partition of ONE structured mesh over ranks: rank g owns a contiguous range of element columns
-- contiguous in the reference numbering e = j*NY + k, model.py:935 -- and the node columns with the same indices (the
last rank also owns node column NX).  A rank with a left neighbour additionally computes that neighbour's last
element column (the "halo"): node column a then has both of its element columns on the rank, so its rows of K come
out complete, with the reference's ascending-element summation order, and without any communication (redundant work
1 / columns-per-rank).  The rank's mesh is a plain structured mesh of halo + owned element columns; rows of the node
columns listed in `owned_node_cols_local` are bit-identical to the same rows of the global K."""


def strip_shard(NX, rank, world):
    """Element-column range and local mesh of `rank` out of `world` for a mesh with NX element columns."""
    if not (0 <= rank < world) or world > NX:
        raise ValueError('strip_shard: need 0 <= rank < world <= NX')
    a, b = NX * rank // world, NX * (rank + 1) // world
    halo = 1 if rank > 0 else 0
    first = a - halo                       # global index of the local element column 0
    nx_local = b - first
    last = rank == world - 1
    owned_nodes_global = list(range(a, b + (1 if last else 0)))
    return {'rank': rank, 'world': world, 'elem_cols_global': (a, b), 'halo': halo, 'first_col_global': first,
            'nx_local': nx_local, 'owned_elem_cols_local': (halo, nx_local),
            'owned_node_cols_global': owned_nodes_global,
            'owned_node_cols_local': [j - first for j in owned_nodes_global]}


def local_element_ids(shard, NY):
    """Global element ids (e = j*NY + k) of the rank's local mesh, in local element order (halo column first)."""
    first, nx = shard['first_col_global'], shard['nx_local']
    return [j * NY + k for j in range(first, first + nx) for k in range(NY)]


def slice_node_vector(vec, shard, NY, ncomp=2):
    """Entries of a global nodal vector (node id j*(NY+1)+k, `ncomp` values per node: Model.u / f and the calc_BC vectors)
    that belong to the rank's local node grid -- halo and ghost columns included, which is what plf_solve_linear wants on
    a rank of a distributed solve.  Node columns are contiguous, so this is a view."""
    first, nx = shard['first_col_global'], shard['nx_local']
    w = ncomp * (NY + 1)
    return vec[first * w:(first + nx + 1) * w]


def owned_slice(shard, NY, ncomp=2):
    """Index range, inside the rank's local nodal vector, of the node columns it owns."""
    cols = shard['owned_node_cols_local']
    w = ncomp * (NY + 1)
    return slice(cols[0] * w, (cols[-1] + 1) * w)


def share_bytes(path, rank, payload=None, timeout=120., max_age=300.):
    """Side channel for the NCCL unique id when no launcher-provided store is at hand: rank 0 writes `payload` to `path`
    (atomically, via rename), the other ranks wait for the file and read it.  All ranks of one job must use the same path
    (e.g. derived from MASTER_PORT) on a file system they share.  A file older than `max_age` seconds is taken for the
    left-over of a crashed job and ignored (a rank that read a stale id would hang in ncclCommInitRank); rank 0 replaces
    whatever is there."""
    import os
    import time
    if rank == 0:
        tmp = '%s.tmp.%d' % (path, os.getpid())
        with open(tmp, 'wb') as f:
            f.write(payload)
        os.replace(tmp, path)
        return payload
    t0 = time.time()

    def fresh():
        try:
            return time.time() - os.path.getmtime(path) <= max_age
        except OSError:
            return False
    while not fresh():
        if time.time() - t0 > timeout:
            raise TimeoutError('share_bytes: %s did not appear within %g s' % (path, timeout))
        time.sleep(0.02)
    with open(path, 'rb') as f:
        return f.read()
