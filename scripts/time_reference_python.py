#!/usr/bin/env python3
"""Time the UNMODIFIED Python reference on a sample of bench.py's workload (build container only: needs /root/reference).

    python scripts/time_reference_python.py [n_sample = 1500] [mesh N = 1024] [j2 | hill]

SURVEY.md 8d / BASELINE.md plan item 2: replay (sig, epl, deps, CV) tuples of the 1024 x 1024 J2 bench workload (the same
seeded generator, bench.synth_states) through the reference's Material.response (material.py:207-346) and through
Element.calc_Kel-style integration + Model.setupK of a small mesh, single core -- the reference is serial pure Python.
The same tuples go through the oracle port (oracle/plf_oracle.c) to show (a) that the port returns what the reference
returns on the bench workload itself and (b) how much faster the C port -- which bench.py's `--impl reference` arm and
`cpu_baseline` have to use on the GPU box, where /root/reference does not exist -- is than the real reference.
Writes profiles/r1_reference_python_timing.json."""
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np  # noqa: E402
import bench  # noqa: E402
from ref_shim import import_reference  # noqa: E402
from oracle import oracle as orc  # noqa: E402

warnings.filterwarnings('ignore')


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    cfg = sys.argv[3] if len(sys.argv) > 3 else 'j2'
    hill, khard = (None, bench.KHARD) if cfg == 'j2' else ([0.7, 1., 1.4, 1., 1.2, 0.8], 100.)   # BASELINE config 3
    FE = import_reference()
    sig, epl, deps = bench.synth_states(N, seed=1234, khard=khard)
    rng = np.random.default_rng(5)
    idx = np.sort(rng.choice(len(sig), n, replace=False))
    sig, epl, deps = sig[idx], epl[idx], deps[idx]
    m = FE.Material()
    m.elasticity(E=bench.E_MOD, nu=bench.NU)
    m.plasticity(sy=bench.SY, khard=khard, hill=hill, sdim=6)
    CV = np.array(m.CV)

    fy = np.zeros(n); so = np.zeros((n, 6)); dp = np.zeros((n, 6)); ct = np.zeros((n, 6, 6)); ns = np.zeros(n, dtype=np.int32)
    t0 = time.perf_counter()
    for i in range(n):
        m.msg['nsteps'] = -1
        fy[i], so[i], dp[i], ct[i] = m.response(sig[i], epl[i], deps[i], CV)
        ns[i] = m.msg['nsteps']
    t_ref = time.perf_counter() - t0

    om = orc.Mat(kind=orc.HILL, E=bench.E_MOD, nu=bench.NU, sy=bench.SY, khard=khard, hill=hill)
    t0 = time.perf_counter()
    ofy, oso, odp, oct_, ons = orc.response_batch([om], None, sig, epl, deps, CV[None])
    t_orc = time.perf_counter() - t0

    def rel(a, b):
        return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
    elastic = int(np.sum((ns == 0) & (np.abs(ct - CV).max(axis=(1, 2)) == 0)))
    sub50 = int(np.sum(ns == 49))

    # K assembly of the reference: Model.setupK() wall time on a 32 x 32 mesh (dense K 38 MB), best of 3
    fe = FE.Model(dim=2, planestress=False)
    fe.geom([4.], LY=4.)
    fe.assign([m])
    fe.mesh(NX=32, NY=32)
    for el in fe.element:
        el.calc_Kel()
    tk = []
    for _ in range(3):
        t0 = time.perf_counter()
        for el in fe.element:
            el.calc_Kel()          # what Model.solve redoes for every element whose tangent changed (model.py:1354)
        fe.setupK()
        tk.append(time.perf_counter() - t0)

    out = {
        'what': 'unmodified pyLabFEA reference (Python, single core) on a seeded sample of the bench workload',
        'host': {'cpu': open('/proc/cpuinfo').read().split('model name')[1].split('\n')[0].strip(': \t'), 'nproc': os.cpu_count(),
                 'where': 'build container (the reference cannot travel to the GPU box)'},
        'workload': '%d elements sampled (seed 5) from bench.synth_states(%d, seed=1234), %s' % (n, N, 'J2 sy=150 khard=500' if cfg == 'j2' else 'Hill-6p [0.7,1,1.4,1,1.2,0.8] sy=150 khard=100 (BASELINE config 3)'),
        'branches': {'elastic': elastic, 'substep50': sub50, 'other': n - elastic - sub50},
        'reference_response_s': t_ref, 'reference_updates_per_s_1core': n / t_ref,
        'oracle_port_response_s': t_orc, 'oracle_port_updates_per_s_1core': n / t_orc,
        'port_over_reference': t_ref / t_orc,
        'oracle_vs_reference_on_this_sample': {'nsteps_equal': bool(np.array_equal(ns, ons)), 'fy_abs_over_sy': float(np.max(np.abs(fy - ofy)) / bench.SY),
                                               'sig_rel': rel(oso, so), 'depl_rel': rel(odp, dp), 'ctan_rel': rel(oct_, ct)},
        'reference_calc_Kel_plus_setupK_32x32_s_best_of_3': min(tk), 'reference_assembly_elements_per_s_1core': 1024 / min(tk),
    }
    with open(os.path.join(ROOT, 'profiles', 'r1_reference_python_timing%s.json' % ('' if cfg == 'j2' else '_' + cfg)), 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
