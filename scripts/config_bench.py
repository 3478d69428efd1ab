#!/usr/bin/env python3
"""Kernel throughput on the other BASELINE.json configurations (bench.py measures the headline 1024 x 1024 J2 case):

    python scripts/config_bench.py j2_256 | hill_512 | svc_512 | barlat_2048 | j2ps_1024 [n_elements]

config 2: 256 x 256 J2; config 3: 512 x 512 Hill (sy=150, hill=[0.7,1,1.4,1,1.2,0.8], khard=100); config 4: 512 x 512
with the RBF-SVC yield function of examples/train_hill.py (1 585 support vectors, the model the reference trained in
the build container: tests/golden/svc_hill.npz); config 5: 2048 x 2048 Barlat Yld2004-18p (examples/train_goss_barlat.py
parameters, a = 8; one GPU's share or a sample of it).  Inputs resident in HBM, CUDA events, L2 flushed between repetitions
for the small meshes.  Prints one JSON line per config."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np  # noqa: E402
import bench  # noqa: E402
from pylabfea_b200 import _lib  # noqa: E402
import pylabfea_b200 as FE  # noqa: E402


def run(name, n_override=None):
    ctx = _lib.Context(0)
    reps = 20
    planestress = False
    if name == 'j2_256':
        N = 256
        mat, _ = bench.j2_material()
        sig, epl, deps = bench.synth_states(N, 1234)
    elif name == 'hill_512':
        N = 512
        m = FE.Material()
        m.elasticity(E=200.e3, nu=0.3)
        m.plasticity(sy=150., hill=[0.7, 1., 1.4, 1., 1.2, 0.8], khard=100., sdim=6)
        mat = m._descriptor()
        sig, epl, deps = bench.synth_states(N, 1234, khard=100.)
    elif name == 'j2ps_1024':   # plane-stress variant of the target workload (not a BASELINE config): block-CV fast mode
        N = 1024
        mat, _ = bench.j2_material()
        sig, epl, deps = bench.synth_states(N, 1234)
        sig[:, 2] = 0.
        planestress = True
    elif name == 'svc_512':
        N = 512
        g = dict(np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'svc_hill.npz')))
        m = FE.Material('ML-Hill-p1')
        m.elasticity(CV=g['CV'])
        m.plasticity(sy=float(g['sy']), sdim=6)
        m.set_svc(g['sv'], g['dual'], float(g['rho']), float(g['gamma']), scale_seq=float(g['scale_seq']))
        mat = m._descriptor()
        sig, epl, deps = bench.synth_states(N, 1234, sy=float(g['sy']), khard=0., deps_scale=1. / 3.)
        reps = 2
    elif name == 'barlat_2048':
        N = 2048
        g = dict(np.load(os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden', 'barlat_response.npz')))
        m = FE.Material('Goss-Barlat')
        m.elasticity(E=float(g['E']), nu=float(g['nu']))
        m.plasticity(sy=float(g['sy']), khard=0., barlat=list(g['bp']), barlat_exp=int(g['barlat_exp']), sdim=6)
        mat = m._descriptor()
        sig, epl, deps = bench.synth_states(N, 1234, sy=float(g['sy']), khard=0., deps_scale=1. / 3.)
        reps = 3
    else:
        raise SystemExit('unknown config ' + name)
    n = N * N if n_override is None else int(n_override)
    sig, epl, deps = sig[:n], epl[:n], deps[:n]
    ctx.set_materials([mat], planestress=planestress)
    lx = 4. / N
    b = ctx.batch_create(n)
    ctx.batch_upload(b, np.ascontiguousarray(sig), np.ascontiguousarray(epl), np.ascontiguousarray(deps))
    full_mesh = (n == N * N)
    if os.environ.get('PLF_MAX_FM'):
        ctx.set_option('max_fm', int(os.environ['PLF_MAX_FM']))
    if full_mesh:
        ctx.mesh_structured(N, N, np.full(N, lx), lx, 1., np.zeros(n, np.int32))
    ctx.batch_run(b)
    ctx.synchronize()
    t_resp, t_asm = [], []
    for i in range(reps):
        ctx.flush_l2()
        ctx.event_record(0)
        ctx.batch_run(b)
        ctx.event_record(1)
        if full_mesh:
            ctx.setupK_from_batch(b)
        ctx.event_record(2)
        ctx.synchronize()
        t_resp.append(ctx.event_elapsed(0, 1))
        t_asm.append(ctx.event_elapsed(1, 2))
    ns = np.empty(n, dtype=np.int32)
    ctx.batch_download(b, nsteps=ns)
    tr, ta = float(np.median(t_resp)), float(np.median(t_asm))
    out = {'config': name, 'elements': n, 'response_ms': tr, 'updates_per_s': n / (tr * 1e-3),
           'response_GBps_algorithmic': 540. * n / (tr * 1e-3) / 1e9, 'substep50_fraction': float(np.mean(ns == 49)),
           'l2': 'flushed between repetitions'}
    if full_mesh:
        out.update(assembly_ms=ta, assembly_elems_per_s=n / (ta * 1e-3))
    print(json.dumps(out))
    ctx.close()


if __name__ == '__main__':
    run(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
