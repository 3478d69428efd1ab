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
