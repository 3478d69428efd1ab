"""world_size-2 tests of the multi-rank host logic on CPU (gloo), launched exactly like the driver launches the bench
(python -m torch.distributed.run, rendezvous on 127.0.0.1).  No GPU needed."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(args, port, timeout=300, nproc=2, script='bench.py', extra_env=None):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(nproc), '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, script)] + args
    env = dict(os.environ)
    env.update(extra_env or {})
    env.setdefault('OMP_NUM_THREADS', '1')
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith('{')]
    return [json.loads(ln) for ln in lines]


def test_two_ranks_shard_plan_barrier_and_reductions():
    out = _torchrun(['--gpus', '2', '--dist-selftest', '--mesh', '64'], 29631)
    assert len(out) == 1, 'rank 0 alone prints the line'
    d = out[0]
    assert d['world'] == 2 and d['scaling'] == 'weak'
    assert d['max_time'] == 2.0                # MAX over ranks of (rank + 1)
    assert d['total_elements'] == 2 * 64 * 64  # every rank owns a full block (weak scaling)


def test_two_ranks_reference_arm_rank0_only():
    out = _torchrun(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0', '--mesh', '64',
                     '--cpu-sample', '1500'], 29632)
    assert len(out) == 1
    d = out[0]
    assert d['impl'] == 'reference' and d['n_gpus'] == 2 and d['gpu_launches'] == 0
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['value'] > 0


def test_shard_plan_is_per_rank():
    sys.path.insert(0, ROOT)
    import bench
    p0, p1 = bench.shard_plan(128, 0, 2), bench.shard_plan(128, 1, 2)
    assert p0['elements'] == p1['elements'] == 128 * 128 and p0['seed'] != p1['seed']


def test_strip_shards_partition_the_mesh_exactly():
    """Host logic of the strip partition (pylabfea_b200/shard.py): owned element and node columns of all ranks tile the
    global mesh exactly once, ranks > 0 carry one halo element column, local element ids are contiguous runs."""
    sys.path.insert(0, ROOT)
    from pylabfea_b200.shard import strip_shard, local_element_ids
    for NX, NY, world in ((1024, 1024, 8), (37, 5, 4), (8, 3, 8), (10, 2, 3)):
        elem_cols, node_cols = [], []
        for r in range(world):
            s = strip_shard(NX, r, world)
            a, b = s['elem_cols_global']
            elem_cols += list(range(a, b))
            node_cols += s['owned_node_cols_global']
            assert s['halo'] == (1 if r > 0 else 0) and s['nx_local'] == b - a + s['halo']
            ids = local_element_ids(s, NY)
            assert ids == list(range(s['first_col_global'] * NY, b * NY))
            assert s['owned_node_cols_local'][0] == s['halo']     # the halo's left node column is never owned
        assert elem_cols == list(range(NX)) and node_cols == list(range(NX + 1))
    with pytest.raises(ValueError):
        strip_shard(4, 0, 5)


def test_nodal_vector_slices_of_the_strips_reassemble_the_global_vector():
    """What every rank of a distributed solve passes to plf_solve_linear: the slice of the global calc_BC vectors on its
    local node grid (ghost columns included); the owned parts of all ranks tile the global vector exactly once."""
    import numpy as np
    sys.path.insert(0, ROOT)
    from pylabfea_b200.shard import strip_shard, slice_node_vector, owned_slice
    for NX, NY, world in ((12, 3, 4), (37, 5, 3), (8, 2, 1)):
        v = np.arange(2 * (NX + 1) * (NY + 1), dtype=float)
        parts = []
        for r in range(world):
            s = strip_shard(NX, r, world)
            loc = slice_node_vector(v, s, NY)
            assert len(loc) == 2 * (s['nx_local'] + 1) * (NY + 1)
            assert loc[0] == v[2 * s['first_col_global'] * (NY + 1)]
            parts.append(loc[owned_slice(s, NY)])
            # ghost columns: column 0 on ranks with a left neighbour, the last column on ranks with a right neighbour
            cols = s['owned_node_cols_local']
            assert cols[0] == (1 if r > 0 else 0) and cols[-1] == s['nx_local'] - (1 if r < world - 1 else 0)
        assert np.array_equal(np.concatenate(parts), v)


def test_unique_id_side_channel(tmp_path):
    """share_bytes: rank 0 publishes the NCCL unique id through a file, another process picks it up."""
    sys.path.insert(0, ROOT)
    from pylabfea_b200 import _lib
    from pylabfea_b200.shard import share_bytes
    try:
        uid = _lib.comm_unique_id()
    except _lib.PlfError as e:      # image without libnccl.so.2
        pytest.skip(str(e))
    assert len(uid) == _lib.COMM_ID_BYTES and uid != bytes(_lib.COMM_ID_BYTES)
    path = str(tmp_path / 'id')
    reader = subprocess.Popen([sys.executable, '-c', 'import sys; sys.path.insert(0, %r)\n'
                               'from pylabfea_b200.shard import share_bytes\n'
                               'sys.stdout.buffer.write(share_bytes(%r, 1, timeout=30.))' % (ROOT, path)], stdout=subprocess.PIPE)
    assert share_bytes(path, 0, uid) == uid
    out, _ = reader.communicate(timeout=60)
    assert reader.returncode == 0 and out == uid
    with pytest.raises(TimeoutError):
        share_bytes(str(tmp_path / 'missing'), 1, timeout=0.1)


@pytest.mark.parametrize('nproc,port', [(2, 29641), (3, 29642)])
def test_strip_partitioned_cg_rehearsal_over_gloo(nproc, port):
    """The exchange pattern of the distributed solve (launch_cg_dist: ghost | fixed mask, one halo node column per
    neighbour and iteration, all-reduced dot products, final halo of du) rehearsed on the CPU with real ranks over gloo
    and numpy arithmetic (tests/dist_cg_gloo.py): every rank ends with the dense global solution on its node grid, all
    ranks stop at the same iteration, and the owned rows of every strip stiffness equal the global rows.  3 ranks give
    the middle rank ghosts on both sides.  The same ranks also rehearse the replicated solve (plf_comm_set_replica): owned
    rows gathered from all ranks reproduce the global K bit for bit, and the solve on them is the global solve."""
    out = _torchrun([], port, nproc=nproc, script=os.path.join('tests', 'dist_cg_gloo.py'))
    assert len(out) == 1
    d = out[0]
    assert d['world'] == nproc and d['owned_rows_equal_global']
    assert len(set(d['iters'])) == 1 and 10 < d['iters'][0] < 5000
    assert d['max_rel_err'] < 1e-9
    # the replicated solve: the owned row ranges of all ranks tile the global K exactly once, bit for bit
    assert d['gathered_rows_equal_global'] and d['replicated_max_rel_err'] == 0.


@pytest.mark.parametrize('nproc,port', [(2, 29651), (3, 29652)])
def test_shared_multigrid_levels_rehearsal_over_gloo(nproc, port):
    """The V-cycle with its finest levels shared among the ranks (mg_vcycle with MgDist / mg_plan_sharing: the default solver
    of a distributed Model.solve) rehearsed with real ranks over gloo (tests/dist_shared_vcycle_gloo.py): row ranges per
    level, one ghost column swapped after every step whose output the next stencil reads, gathered right-hand side of the
    first replicated level.  Entries a rank neither owns nor received are NaN, so a missing swap or a wrong range poisons
    the result.  1, 2 and 3 shared levels, even and odd column splits, 3 ranks for a rank with two neighbours: the cycle
    equals the single-process cycle."""
    out = _torchrun([], port, nproc=nproc, script=os.path.join('tests', 'dist_shared_vcycle_gloo.py'))
    assert len(out) == 1 and out[0]['world'] == nproc
    cases = out[0]['cases']
    assert {c['shared'] for c in cases} == {1, 2, 3} and len(cases) >= 8
    assert all(c['max_rel_err'] <= 1e-13 for c in cases), cases


def test_shared_multigrid_rehearsal_detects_a_missing_swap():
    """Self-check of the rehearsal: without the ghost swap of the residual the restriction reads columns nobody sent."""
    out = _torchrun([], 29653, nproc=2, script=os.path.join('tests', 'dist_shared_vcycle_gloo.py'),
                    extra_env={'PLF_REHEARSAL_SKIP_RESIDUAL_SWAP': '1'})
    assert any(c['max_rel_err'] > 1. for c in out[0]['cases'])
