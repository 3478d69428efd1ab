"""world_size-2 tests of the multi-rank host logic on CPU (gloo), launched exactly like the driver launches the bench
(python -m torch.distributed.run, rendezvous on 127.0.0.1).  No GPU needed."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(args, port, timeout=300):
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(port), os.path.join(ROOT, 'bench.py')] + args
    env = dict(os.environ)
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
