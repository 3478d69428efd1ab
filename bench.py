#!/usr/bin/env python3
"""bench.py -- headline benchmark of the hot path (BASELINE.json metric) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mesh 1024]

One "step" = one pass of the non-linear inner loop of Model.solve over one mesh: the batched return mapping
(Material.response for every element) followed by the K integration + CSR assembly (Model.setupK) from the freshly
computed tangents.  Workload (config.workload): 1024x1024 Q4 mesh, J2 plasticity (sy=150, khard=500, E=200e3,
nu=0.3, plane strain), synthetic spatially correlated element states (seeded) that populate the elastic / 1-step /
50-sub-step branches of the reference algorithm.  N>1 (torchrun): ONE mesh of N*1024 x 1024 elements cut into strips of
1024 element columns, one strip per rank; a rank with a left neighbour recomputes that neighbour's last element column
(halo), so the K rows of its node columns are complete without communication (pylabfea_b200/shard.py; weak scaling,
no data-path collective -- element updates and element rows are independent).

`value`  : elements/s through (stress update + assembly), inputs resident in HBM, CUDA events, max over ranks.
`e2e`    : same step through the C-ABI call that takes HOST buffers (plf_step_host: states in, response results and
           CSR(K) values out), pinned host memory, H2D + D2H inside the timed region.
`roofline`: return-mapping kernel(s): 540 B/update over the measured launch time vs MEASURED_PEAKS.json.
`cpu_baseline`: the CPU oracle (port of the reference algorithm, oracle/plf_oracle.c) on a bounded sample.
--impl reference: the same step on the host cores only (oracle port, all threads), bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

E_MOD, NU, SY, KHARD = 200.e3, 0.3, 150., 500.
BYTES_PER_UPDATE = 540.    # 144 B in + 4 B id + 392 B out
BYTES_PER_ASSEMBLED = 592.  # 288 B tangent + 16 B connectivity-equivalent + 288 B CSR values


def j2_material():
    from pylabfea_b200 import _lib
    m = _lib.PlfMaterial()
    m.kind = _lib.MAT_HILL
    m.E, m.nu = E_MOD, NU
    hh = E_MOD / ((1. + NU) * (1. - 2. * NU))
    C11, C12, C44 = (1. - NU) * hh, NU * hh, (0.5 - NU) * hh
    CV = np.zeros((6, 6))
    CV[0, 0] = CV[1, 1] = CV[2, 2] = C11
    CV[0, 1] = CV[0, 2] = CV[1, 2] = CV[1, 0] = CV[2, 0] = CV[2, 1] = C12
    CV[3, 3] = CV[4, 4] = CV[5, 5] = C44
    m.CV[:] = list(CV.ravel())
    m.C44 = C44
    m.sy, m.khard = SY, KHARD
    m.hill[:] = [1.] * 6
    m.d[:] = [0.] * 3
    return m, CV


def synth_states(N, seed, correlated=True, sy=SY, khard=KHARD, deps_scale=1.):
    """Element states for an N x N mesh (element id = j*N + k).  Seeded, no file I/O.

    correlated: a stress field as a loaded inclusion problem produces it -- uniaxial-y dominated stress whose level
    (fraction of the flow stress) and direction vary smoothly in space (low-frequency Fourier modes), accumulated
    plastic strain where the level is high, and a smooth strain increment of eps_eq ~ 5e-4 with
    5 % element-wise noise.  Uncorrelated: the same marginals shuffled element-wise (worst case for divergence)."""
    rng = np.random.default_rng(seed)
    j, k = np.meshgrid(np.arange(N), np.arange(N), indexing='ij')
    x, y = (j + .5) / N, (k + .5) / N

    def smooth(nm=4):
        f = np.zeros_like(x)
        for _ in range(nm):
            kx, ky = rng.integers(1, 4, 2)
            f += rng.uniform(.3, 1.) * np.sin(2 * np.pi * (kx * x + rng.uniform())) * np.cos(2 * np.pi * (ky * y + rng.uniform()))
        return f / nm
    level = np.clip(0.53 + 1.9 * smooth(), 0.05, 1.0)            # seq / sflow before the increment (mix ~ 1/3 elastic, 12-15 % 50-sub-step as measured in reference solves)
    peeq = np.where(level > 0.97, 0.01 * (1. + smooth()), 0.)      # yielded zones carry plastic strain
    peeq = np.clip(peeq, 0., None)
    ang = 0.35 * smooth()                                          # in-plane rotation of the principal axis
    n = N * N
    sig = np.zeros((n, 6))
    c, s = np.cos(ang).ravel(), np.sin(ang).ravel()
    s1 = np.ones(n)                                                # uniaxial unit stress along the rotated y axis
    sig[:, 0] = s1 * s * s
    sig[:, 1] = s1 * c * c
    sig[:, 5] = s1 * s * c
    sig[:, 2] = NU * 0.5 * (sig[:, 0] + sig[:, 1]) * (0.5 + 0.5 * smooth().ravel())   # partial plane-strain constraint
    d01, d12, d20 = sig[:, 0] - sig[:, 1], sig[:, 1] - sig[:, 2], sig[:, 2] - sig[:, 0]
    seq = np.sqrt(0.5 * (d01 ** 2 + d12 ** 2 + d20 ** 2 + 6 * sig[:, 5] ** 2))
    sflow = sy + khard * peeq.ravel()
    sig *= (level.ravel() * sflow / seq)[:, None]
    epl = peeq.ravel()[:, None] * np.array([-.5, 1., -.5, 0., 0., 0.])[None, :]
    deps = np.zeros((n, 6))
    eyy = 5.e-4 * (1. + 0.4 * smooth().ravel())
    deps[:, 1] = eyy
    deps[:, 0] = -0.4 * eyy * (1. + 0.3 * smooth().ravel())
    deps[:, 5] = 2.e-4 * smooth().ravel()
    deps *= deps_scale * (1. + 0.05 * rng.standard_normal((n, 1)))
    if not correlated:
        p = rng.permutation(n)
        sig, epl, deps = sig[p], epl[p], deps[p]
    return np.ascontiguousarray(sig), np.ascontiguousarray(epl), np.ascontiguousarray(deps)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.device), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace('.', '').isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith('active'):
                        reasons.add(nme)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def oracle_step(orc, mats, CV, sig, epl, deps, B4, jacw, nthreads):
    """The step on the host: oracle response for every sampled element, then Kel of each new tangent (the CPU side of
    setupK; the O(1) adds of the assembly itself are not included, which favours the CPU)."""
    return orc.step_batch(mats, sig, epl, deps, CV[None], B4, jacw, nthreads)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm on the host cores only (oracle port; the Python reference cannot
    travel to the GPU box and cannot allocate its dense K at this size)."""
    if rank != 0:
        return
    from oracle import oracle as orc
    N = args.mesh
    sample = args.cpu_sample
    sig, epl, deps = synth_states(N, 1234)
    idx = np.random.default_rng(7).choice(len(sig), sample, replace=False)
    sig, epl, deps = sig[idx], epl[idx], deps[idx]
    _, CV = j2_material()
    mats = [orc.Mat(kind=orc.HILL, E=E_MOD, nu=NU, sy=SY, khard=KHARD)]
    lx = 4. / N
    B4 = orc.gauss_bmats(lx, lx, 0, CV, E_MOD, NU)
    jacw = (lx * lx * 1.) * 4.
    cores = host_cores()
    for _ in range(args.warmup):
        oracle_step(orc, mats, CV, sig[:2000], epl[:2000], deps[:2000], B4, jacw, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_step(orc, mats, CV, sig, epl, deps, B4, jacw, cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = sample / dt
    unit = 'elements/s'
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': unit, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(N),
            'cpu_baseline': {'value': val, 'unit': unit, 'cores': cores, 'kind': 'port',
                             'sample': '%d elements sampled (seeded) from the %dx%d workload per step' % (sample, N, N),
                             'note': 'C port of the reference algorithm, all host threads; the unmodified Python reference '
                                     'runs Material.response at 181 updates/s on one core (measured in the build container on a '
                                     'sample of this workload, where the port returns identical results: '
                                     'profiles/r1_reference_python_timing.json)'},
            'e2e': {'value': val, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line))


METRIC = 'Gauss-point stress updates/sec + K-assembly elems/sec (one step = return mapping + setupK over the mesh)'


def workload_config(N):
    return {'workload': '%dx%d Q4 mesh, J2 plasticity (sy=150, khard=500, E=200e3, nu=0.3, plane strain), one inner pass '
                        'of a load step: Material.response per element + Model.setupK into CSR' % (N, N),
            'elements': N * N, 'state': 'synthetic correlated field, seed 1234', 'l2': 'inputs larger than L2 '
            '(%.0f MB read+written per step)' % (N * N * (BYTES_PER_UPDATE + BYTES_PER_ASSEMBLED) / 1e6)}


def shard_plan(N, rank, world):
    """Weak scaling: the global mesh has world*N x N elements; every rank owns a strip of N element columns (N*N
    elements) and, when it has a left neighbour, additionally computes that neighbour's last column (halo) so that its
    K rows need no communication.  The strip's synthetic state is seeded per rank."""
    from pylabfea_b200.shard import strip_shard
    s = strip_shard(world * N, rank, world)
    return {'rank': rank, 'world': world, 'mesh': N, 'elements': N * N, 'seed': 1234 + rank, 'halo': s['halo'],
            'nx_local': s['nx_local'], 'elem_cols_global': s['elem_cols_global']}


class RankGroup:
    """torch.distributed plumbing for the bench: barrier + MAX/SUM of a scalar over the ranks.  NCCL on GPUs
    (device tensors), gloo for the CPU self-test.  With one rank it is a no-op and torch is never imported."""

    def __init__(self, world, local, backend='nccl'):
        self.world, self.dist, self.dev = world, None, 'cpu'
        if world > 1:
            import torch
            import torch.distributed as dist
            self.torch, self.dist = torch, dist
            if backend == 'nccl':
                torch.cuda.set_device(local)
                self.dev = 'cuda'
                dist.init_process_group('nccl', device_id=torch.device('cuda', local))
            else:
                dist.init_process_group(backend)

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _reduce(self, v, op):
        if self.dist is None:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, v):
        return self._reduce(v, self.dist.ReduceOp.MAX if self.dist else None)

    def sum_over_ranks(self, v):
        return self._reduce(v, self.dist.ReduceOp.SUM if self.dist else None)

    def broadcast_bytes(self, payload, n):
        """n bytes from rank 0 to every rank (the NCCL unique id of the solver's own communicator)."""
        if self.dist is None:
            return payload
        t = self.torch.zeros(n, dtype=self.torch.uint8, device=self.dev)
        if payload is not None:
            t = self.torch.tensor(list(payload), dtype=self.torch.uint8, device=self.dev)
        self.dist.broadcast(t, src=0)
        return bytes(t.cpu().tolist())

    def close(self):
        if self.dist is not None:
            self.dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# BASELINE.json configs 1-5 (the headline line is the 1024 x 1024 target; these ride along in its `configs` block, and
# `--config cN` makes one of them the workload of the run, e.g. under ncu)
# ---------------------------------------------------------------------------------------------------------------------
FP64_PEAK_TFLOPS = 148 * 2 * 32 * 2 * 1.965e9 / 1e12   # 2 warp-DFMA / clk / SM (profiles/r1_microbench_dfma_lat.txt) at 1965 MHz
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def config_material(name):
    """(plf_material descriptor, kernel label, synthetic-state keywords, mesh N) of a BASELINE config."""
    import pylabfea_b200 as FE
    if name in ('c2', 'target'):
        mat, _ = j2_material()
        return mat, 'response_passA_kernel + response_passB_kernel', dict(), 256 if name == 'c2' else 1024
    if name == 'c3':
        m = FE.Material()
        m.elasticity(E=200.e3, nu=0.3)
        m.plasticity(sy=150., hill=[0.7, 1., 1.4, 1., 1.2, 0.8], khard=100., sdim=6)
        return m._descriptor(), 'response_passA_kernel + response_passB_kernel (Hill mode)', dict(khard=100.), 512
    if name == 'c4':
        g = dict(np.load(os.path.join(GOLDEN, 'svc_hill.npz')))
        m = FE.Material('ML-Hill-p1')
        m.elasticity(CV=g['CV'])
        m.plasticity(sy=float(g['sy']), sdim=6)
        m.set_svc(g['sv'], g['dual'], float(g['rho']), float(g['gamma']), scale_seq=float(g['scale_seq']))
        return m._descriptor(), 'response_svc_kernel (1585 support vectors)', dict(sy=float(g['sy']), khard=0., deps_scale=1. / 3.), 512
    if name == 'c5':
        g = dict(np.load(os.path.join(GOLDEN, 'barlat_response.npz')))
        m = FE.Material('Goss-Barlat')
        m.elasticity(E=float(g['E']), nu=float(g['nu']))
        m.plasticity(sy=float(g['sy']), khard=0., barlat=list(g['bp']), barlat_exp=int(g['barlat_exp']), sdim=6)
        return m._descriptor(), 'response_barlat kernels', dict(sy=float(g['sy']), khard=0., deps_scale=1. / 3.), 2048
    raise SystemExit('unknown config ' + name)


def profile_json(name):
    try:
        return json.load(open(os.path.join(ROOT, 'profiles', name)))
    except Exception:
        return {}


def run_kernel_config(name, device, reps, peak, opts=()):
    """Return mapping + assembly of one BASELINE config on synthetic states: inputs resident in HBM, CUDA events on the
    context's stream, L2 flushed before every repetition (the small meshes would otherwise sit in the 126 MB L2)."""
    from pylabfea_b200 import _lib
    mat, kernel, kw, N = config_material(name)
    ctx = _lib.Context(device)
    for kv in opts:
        k, v = kv.split('=')
        ctx.set_option(k, int(v))
    n = N * N
    sig, epl, deps = synth_states(N, 1234, **kw)
    ctx.set_materials([mat], planestress=False)
    lx = 4. / N
    ctx.mesh_structured(N, N, np.full(N, lx), lx, 1., np.zeros(n, np.int32))
    b = ctx.batch_create(n)
    ctx.batch_upload(b, sig, epl, deps)
    l0 = ctx.launch_count()
    for _ in range(3 if name not in ('c4',) else 1):
        ctx.batch_run(b)
        ctx.setupK_from_batch(b)
    ctx.synchronize()
    t_resp, t_asm = [], []
    for i in range(reps):
        ctx.flush_l2()
        ctx.event_record(0)
        ctx.batch_run(b)
        ctx.event_record(1)
        ctx.setupK_from_batch(b)
        ctx.event_record(2)
        ctx.synchronize()
        t_resp.append(ctx.event_elapsed(0, 1))
        t_asm.append(ctx.event_elapsed(1, 2))
    launches = ctx.launch_count() - l0
    ns = np.empty(n, dtype=np.int32)
    ct = np.empty((n, 6, 6))
    ctx.batch_download(b, ctan=ct, nsteps=ns)
    cv = ctx.element_CV(0)
    n_el = int(np.sum(np.all(ct.reshape(n, 36) == cv.ravel()[None, :], axis=1)))
    n50 = int(np.sum(ns == 49))
    tr, ta = float(np.median(t_resp)), float(np.median(t_asm))
    ach = BYTES_PER_UPDATE * n / (tr * 1e-3) / 1e9
    ach_a = BYTES_PER_ASSEMBLED * n / (ta * 1e-3) / 1e9
    fp = profile_json('fp64_ops.json').get(name)   # FP64 flop per update counted by ncu on this config (scripts/summarize_ncu.py)
    out = {'workload': '%dx%d Q4 mesh, %s, synthetic correlated states (seed 1234), L2 flushed between repetitions' % (N, N, kernel),
           'elements': n, 'reps': reps, 'response_ms': tr, 'updates_per_s': n / (tr * 1e-3), 'assembly_ms': ta,
           'assembly_elems_per_s': n / (ta * 1e-3), 'ms_per_step': tr + ta, 'elements_per_s': n / ((tr + ta) * 1e-3),
           'branch_mix': {'elastic': n_el / n, 'substep50': n50 / n, 'plastic_1step': 1. - (n_el + n50) / n},
           'roofline': {'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                        'assembly': {'achieved': ach_a, 'frac': ach_a / peak}},
           'gpu_launches': int(launches)}
    if fp:
        tf = fp['flop_per_update'] * n / (tr * 1e-3) / 1e12
        out['roofline_fp64'] = {'bound': 'fp64', 'achieved': tf, 'peak': FP64_PEAK_TFLOPS, 'unit': 'TFLOP/s', 'frac': tf / FP64_PEAK_TFLOPS,
                                'flop_per_update': fp['flop_per_update'], 'source': fp.get('source')}
    ctx.batch_destroy(b)
    ctx.close()
    return out


def inclusion_model(N, eps, device=0):
    """examples/inclusion.py pattern: J2 matrix (sy=150, khard=500) with a soft elastic inclusion in the
    middle third, uniaxial-y displacement control, corner pin."""
    import pylabfea_b200 as FE
    n1 = N // 3
    el = np.ones((N, N))
    el[n1:2 * n1, n1:2 * n1] = 2
    mat1 = FE.Material(num=1)
    mat1.elasticity(E=E_MOD, nu=NU)
    mat1.plasticity(sy=SY, khard=KHARD, sdim=6)
    mat2 = FE.Material(num=2)
    mat2.elasticity(E=3.e3, nu=0.3)
    fe = FE.Model(dim=2, planestress=False, device=device)
    fe.geom(sect=2, LX=4., LY=4.)
    fe.assign([mat1, mat2])
    fe.bcbot(0.)
    fe.bcright(0., 'force')
    fe.bcleft(0., 'force')
    fe.bctop(eps * fe.leny, 'disp')
    fe.mesh(elmts=el, NX=N, NY=N)
    fe.bcnode(0, 0., 'disp', 'x')
    fe.solver_rtol = 1.e-10
    return fe


def run_load_sequence(N, min_step, eps, device=0):
    """Model.solve through the public API: every increment runs on the device (plf_load_increment).  Wall clock around
    solve() (host buffers: only the boundary description goes in and the increment reports come back)."""
    fe = inclusion_model(N, eps, device)
    ctx = fe._dev()
    def launches():   # the strip's context + the context that holds the replicated global system
        rep = getattr(ctx, '_replica', None)
        return ctx.launch_count() + (rep.launch_count() if rep is not None else 0)
    l0 = launches()
    t0 = time.perf_counter()
    fe.solve(min_step=min_step)
    dt = time.perf_counter() - t0
    its = np.array([s[0] for s in fe.solver_stats])
    passes = int(sum(max(n, 0) for n in fe.niter)) + len(fe.niter)
    out = {'workload': '%dx%d inclusion mesh (J2 sy=150 khard=500 + elastic inclusion), uniaxial eps=%g, Model.solve(min_step=%d)'
                       % (N, N, eps, min_step),
           'elements': N * N, 'dofs': int(fe.Ndof), 'increments': int(fe.nsteps), 'inner_passes': [int(x) for x in fe.niter],
           'linear_solves': int(len(its)), 'cg_iterations': int(its.sum()), 'seconds': dt,
           'ms_per_increment': dt * 1e3 / max(fe.nsteps, 1),
           'element_updates_per_s': float(sum(sum(s[5][:3]) for s in fe.sweep_stats) / dt),
           'gpu_launches': int(ctx.launch_count() - l0), 'launches_per_solve': float((ctx.launch_count() - l0) / max(len(its), 1)),
           'global_stress_yy': float(fe.glob['sig'][1]), 'substep50_final': float(fe.sweep_stats[-1][5][2] / max(sum(fe.sweep_stats[-1][5][:3]), 1)) if fe.sweep_stats else 0.}
    return out, fe


def harvested_step(N, min_step, eps, device, peak, steps=30):
    """The hot path on a state HARVESTED from a real solve instead of the synthetic branch mix of the headline.  The load
    sequence is run once more (untimed) with Model.increment_hook saving the committed element state (sig, eps, epl) at the
    start of every increment; afterwards the state at the start of the LAST increment is put back, and one step = plf_sweep
    (strains from the du of the last solve, Material.response of every element, tangents) + plf_setupK -- the work of the last
    inner pass of the sequence, elements crossing the yield surface inside the increment included.  The sweep leaves the
    committed state alone, so the step repeats identically (skip_unchanged is off for it: an unchanged tangent would otherwise
    not be re-assembled).  Wall clock around `steps` steps with a synchronize on both sides; every step includes the sweep's
    flag read-back."""
    from pylabfea_b200 import _lib
    fe = inclusion_model(N, eps, device)
    ctx = fe._dev()
    snap = {}

    def hook(model, il):
        for f in (_lib.F_SIG, _lib.F_EPS, _lib.F_EPL):
            snap[f] = ctx.get_field(f)
    fe.increment_hook = hook
    fe.solve(min_step=min_step)
    last_pass = fe.sweep_stats[-1][5]

    def restore():
        for f, a in snap.items():
            ctx.set_field(f, a)
    # strain increments of that last pass: what plf_commit adds to Element.eps
    restore()
    ctx.sweep(1)
    ctx.commit()
    deps = ctx.get_field(_lib.F_EPS) - snap[_lib.F_EPS]
    restore()
    ctx.set_option('skip_unchanged', 0)
    try:
        # (1) the in-FE step: plf_sweep (strain kernel, update, tangent comparison, statistics, flag read-back) + plf_setupK
        for _ in range(3):
            ctx.sweep(1)
            ctx.setupK()
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            _, _, _, counts = ctx.sweep(1)
            ctx.setupK()
        ctx.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / steps
    finally:
        ctx.set_option('skip_unchanged', 1)
    # (2) the headline kernels (plf_batch_run + plf_setupK_from_batch) on the same states, timed like the headline
    n1 = N // 3
    mid = np.zeros((N, N), np.int32)
    mid[n1:2 * n1, n1:2 * n1] = 1
    b = ctx.batch_create(fe.Nel)
    ctx.batch_upload(b, snap[_lib.F_SIG], snap[_lib.F_EPL], deps, mat_id=mid.ravel())
    ctx.last_kernels()
    for _ in range(3):
        ctx.batch_run(b)
        ctx.setupK_from_batch(b)
    kernels = ctx.last_kernels()
    ctx.synchronize()
    t_resp, t_asm = [], []
    for _ in range(steps):
        ctx.flush_l2()
        ctx.event_record(0)
        ctx.batch_run(b)
        ctx.event_record(1)
        ctx.setupK_from_batch(b)
        ctx.event_record(2)
        ctx.synchronize()
        t_resp.append(ctx.event_elapsed(0, 1))
        t_asm.append(ctx.event_elapsed(1, 2))
    ns = np.empty(fe.Nel, dtype=np.int32)
    ctx.batch_download(b, nsteps=ns)
    ctx.batch_destroy(b)
    tr, ta = float(np.median(t_resp)), float(np.median(t_asm))
    n = float(sum(counts[:3]))
    gbs = (BYTES_PER_UPDATE + BYTES_PER_ASSEMBLED) * fe.Nel / (ms * 1e-3) / 1e9
    ach = BYTES_PER_UPDATE * fe.Nel / (tr * 1e-3) / 1e9
    return {'workload': '%dx%d inclusion load sequence (J2 matrix + elastic inclusion); element state at the start of its last '
                        'increment + strain increments of its last pass' % (N, N),
            'branch_mix': {'elastic': counts[0] / n, 'plastic_1step': counts[1] / n, 'substep50': counts[2] / n,
                           'radial_correction': counts[3] / n,
                           'note': 'of the elements of the plastic material; the %d inclusion elements are elastic' % int(mid.sum())},
            'branch_counts_last_pass_of_the_sequence': [int(c) for c in last_pass],
            'hot_path': {'api': 'plf_batch_run + plf_setupK_from_batch, inputs resident, CUDA events, L2 flushed (as the headline)',
                         'kernels': kernels, 'response_ms': tr, 'assembly_ms': ta, 'ms_per_step': tr + ta,
                         'elements_per_s': fe.Nel / ((tr + ta) * 1e-3), 'updates_per_s': fe.Nel / (tr * 1e-3),
                         'substep50_elements': int(np.sum(ns == 49)),
                         'roofline': {'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                                      'note': 'return mapping alone; with %.0f %% of the elements in the 50-sub-step loop the '
                                              'kernel pair is FP64-bound, not HBM-bound' % (100. * np.mean(ns == 49))}},
            'in_fe_step': {'api': 'plf_sweep (strains from du, update, tangent comparison, statistics, flag read-back) + plf_setupK, '
                                  'wall clock with a synchronize on both sides',
                           'ms_per_step': ms, 'elements_per_s': fe.Nel / (ms * 1e-3), 'steps': steps,
                           'roofline_step': {'bound': 'hbm', 'achieved': gbs, 'peak': peak, 'unit': 'GB/s', 'frac': gbs / peak,
                                             'algorithmic_bytes_per_element': BYTES_PER_UPDATE + BYTES_PER_ASSEMBLED}}}


def run_coupled_solve(group, rank, world, local, N=256, min_step=3, eps=0.002):
    """The part of the path WITH an exchange step: Model.solve of ONE model cut into strips over the ranks
    (Model.distribute, one process per GPU).  Workload = the material of BASELINE config 4 -- RBF-SVC yield function of
    examples/train_hill.py, 1585 support vectors -- on a 256 x 256 mesh with an elastic inclusion under uniaxial
    displacement control (the 512 x 512 run takes 100 s on one GPU: 165 passes at 0.36 M in-FE updates/s; the kernel-level
    config-4 numbers are in `configs.c4`).  STRONG scaling: the
    model is fixed; every rank updates the elements of its strip and assembles its K rows, the owned rows of all ranks are
    gathered (NCCL broadcasts, one contiguous CSR range per rank) into a replicated global matrix; the CG vectors and the
    finest multigrid levels are shared among the ranks when the peer mailboxes are up (plf_comm_p2p; else every rank solves
    the complete system alone); sweep flags, calc_scf moments, global averages and boundary sums
    are all-reduced.  Wall clock between barriers around solve(), max over ranks."""
    import pylabfea_b200 as FE
    from pylabfea_b200 import _lib
    g = dict(np.load(os.path.join(GOLDEN, 'svc_hill.npz')))
    mat1 = FE.Material('ML-Hill-p1', num=1)
    mat1.elasticity(CV=g['CV'])
    mat1.plasticity(sy=float(g['sy']), sdim=6)
    mat1.set_svc(g['sv'], g['dual'], float(g['rho']), float(g['gamma']), scale_seq=float(g['scale_seq']))
    mat2 = FE.Material(num=2)
    mat2.elasticity(E=3.e3, nu=0.3)
    n1 = N // 3
    el = np.ones((N, N))
    el[n1:2 * n1, n1:2 * n1] = 2
    fe = FE.Model(dim=2, planestress=False, device=local)
    fe.geom(sect=2, LX=4., LY=4.)
    fe.assign([mat1, mat2])
    fe.bcbot(0.)
    fe.bcright(0., 'force')
    fe.bcleft(0., 'force')
    fe.bctop(eps * fe.leny, 'disp')
    fe.mesh(elmts=el, NX=N, NY=N)
    fe.bcnode(0, 0., 'disp', 'x')
    fe.solver_rtol = 1.e-10
    if world > 1:
        uid = group.broadcast_bytes(_lib.comm_unique_id() if rank == 0 else None, _lib.COMM_ID_BYTES)
        fe.distribute(rank, world, uid, solver='replicated')
    # untimed first run: NCCL sets up the connections of each collective on first use (seconds with 8 ranks), the SVC kernels
    # and the CG graph are loaded / captured; then the model is meshed again (zero state, same communicator) and timed
    fe.solve(min_step=min_step)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        fe.mesh(elmts=el, NX=N, NY=N)
    fe.solver_stats = []
    ctx = fe._dev()
    ctx.synchronize()
    group.barrier()
    def launches():   # the strip's context + the context that holds the replicated global system
        rep = getattr(ctx, '_replica', None)
        return ctx.launch_count() + (rep.launch_count() if rep is not None else 0)
    l0 = launches()
    t0 = time.perf_counter()
    fe.solve(min_step=min_step)
    ctx.synchronize()
    group.barrier()
    dt = group.max_over_ranks(time.perf_counter() - t0)
    its = int(sum(s[0] for s in fe.solver_stats))
    updates = float(sum(sum(s[5][:3]) for s in fe.sweep_stats))
    return {'workload': '%dx%d mesh, RBF-SVC yield function (1585 support vectors) + elastic inclusion, Model.solve(min_step=%d), '
                        'eps=%g (material of BASELINE config 4)' % (N, N, min_step, eps),
            'scaling': 'strong',
            'solver': 'single GPU' if world == 1 else (
                'replicated K (rows gathered); CG vectors + finest multigrid levels shared among the ranks, ghost columns and dot '
                'products through CUDA-IPC peer mailboxes' if ctx.comm_p2p() else 'replicated (K rows gathered, every rank solves)'),
            'peer_mailboxes': bool(world > 1 and ctx.comm_p2p()),
            'elements': N * N, 'dofs': int(fe.Ndof), 'seconds': dt, 'increments': int(fe.nsteps),
            'inner_passes': [int(x) for x in fe.niter], 'linear_solves': len(fe.solver_stats), 'cg_iterations': its,
            'element_updates': updates, 'element_updates_per_s': updates / dt,
            'collectives_per_pass': 0 if world == 1 else 'K gather (%d broadcasts) + 2 all-reduces of sweep flags' % world,
            'exchanges_per_cg_iteration': 0 if world == 1 else (
                'shared levels: 1 ghost-column swap per smoothing / residual step of every shared level and 1 for p (5-6 per level), '
                '2 scalar all-reduces, 1 gather of the first replicated level (%d broadcasts)' % world if ctx.comm_p2p()
                else 'none (every rank solves the gathered system alone)'),
            'launches_rank0': int(launches() - l0), 'global_stress_yy': float(fe.glob['sig'][1])}


def run_c1(device):
    """BASELINE config 1: 10 x 10 linear-elastic cantilever (plane stress) through Model.setupK + Model.solve."""
    import pylabfea_b200 as FE
    mat = FE.Material()
    mat.elasticity(E=200.e3, nu=0.3)
    fe = FE.Model(dim=2, planestress=True, device=device)
    fe.geom([10.], LY=1.)
    fe.assign([mat])
    fe.bcleft(0., 'disp', 'x')
    fe.bcleft(0., 'disp', 'y')
    fe.bcbot(0., 'force', 'y')
    fe.bctop(0., 'force', 'y')
    fe.bcright(-10., 'force', 'y')
    fe.mesh(NX=10, NY=10)
    fe.solve()            # warm-up (JIT-free, but first launches / graph capture)
    reps = 20
    t0 = time.perf_counter()
    for _ in range(reps):
        fe.mesh(NX=10, NY=10)         # forget the solution: the next solve starts from the zero state again
        fe.solve()
    dt = (time.perf_counter() - t0) / reps
    tip = float(fe.u[2 * fe.noright[-1] + 1])
    return {'workload': '10x10 linear-elastic cantilever, plane stress, Model.solve (setupK + reduced CG solve), wall clock',
            'elements': 100, 'ms_per_mesh_and_solve': dt * 1e3, 'elements_per_s': 100 / dt, 'tip_uy': tip,
            'reference_tip_uy': -0.14455, 'note': 'latency-bound: ~40 kernel launches for 242 dofs'}


def cpu_load_step_estimate(seq):
    """What the same load sequence costs on the host, from measured pieces (labelled an estimate: the reference cannot
    allocate its dense K beyond ~128 x 128): per inner pass one threaded oracle step over all plastic
    elements (C port, all host threads) + one sparse direct solve (scipy SuperLU) of a 2-d elasticity system of the same
    size class, measured here on a 256 x 256 mesh and scaled by dof^1.5 (nested-dissection flop count of a 2-d mesh)."""
    from oracle import oracle as orc
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    cores = host_cores()
    n = 200000
    sig, epl, deps = synth_states(512, 1234)
    _, CV = j2_material()
    mats = [orc.Mat(kind=orc.HILL, E=E_MOD, nu=NU, sy=SY, khard=KHARD)]
    lx = 4. / 512
    B4 = orc.gauss_bmats(lx, lx, 0, CV, E_MOD, NU)
    t0 = time.perf_counter()
    oracle_step(orc, mats, CV, sig[:n], epl[:n], deps[:n], B4, lx * lx * 4., cores)
    per_elem = (time.perf_counter() - t0) / n
    m = 256   # 2-d vector Laplacian stand-in with the 9-point x 2-dof sparsity of the Q4 stiffness
    I = sp.identity(m + 1, format='csr')
    T = sp.diags([-1., 2.4, -1.], [-1, 0, 1], shape=(m + 1, m + 1), format='csr')
    A = sp.kron(sp.kron(T, I) + sp.kron(I, T) + 0.25 * sp.kron(T, T), sp.csr_matrix(np.array([[1., .2], [.2, 1.]])), format='csc')
    rhs = np.ones(A.shape[0])
    t0 = time.perf_counter()
    spla.spsolve(A, rhs)
    t_solve_256 = time.perf_counter() - t0
    scale = (seq['dofs'] / A.shape[0]) ** 1.5
    passes = seq['linear_solves']
    est = passes * (per_elem * seq['elements'] + t_solve_256 * scale)
    return {'kind': 'estimate from measured pieces', 'cores': cores, 'seconds': est, 'ms_per_increment': est * 1e3 / max(seq['increments'], 1),
            'oracle_step_s_per_element': per_elem, 'spsolve_256x256_s': t_solve_256, 'spsolve_scaling': 'dof^1.5 -> x%.1f' % scale,
            'passes': passes}


def configs_block(device, peak, quick=False):
    out = {}
    for name, fn in (('c1', lambda: run_c1(device)),
                     ('c2', lambda: run_kernel_config('c2', device, 20, peak)),
                     ('c3', lambda: run_kernel_config('c3', device, 20, peak)),
                     ('c4', lambda: run_kernel_config('c4', device, 1 if quick else 2, peak)),
                     ('c5', lambda: run_kernel_config('c5', device, 2 if quick else 3, peak))):
        try:
            out[name] = fn()
        except Exception as exc:   # a config must never take the headline down with it
            out[name] = {'error': '%s: %s' % (type(exc).__name__, exc)}
    try:
        seq, _ = run_load_sequence(256, 20, 0.01, device)
        out['c2']['load_sequence'] = seq
    except Exception as exc:
        out['c2']['load_sequence'] = {'error': '%s: %s' % (type(exc).__name__, exc)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--dist-selftest', action='store_true')
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--mesh', type=int, default=1024)
    ap.add_argument('--cpu-sample', type=int, default=1000000)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--compact', type=int, default=1)
    ap.add_argument('--assembly-tiled', type=int, default=2)
    ap.add_argument('--opt', action='append', default=[], help='extra plf_set_option key=value (A/B measurements)')
    ap.add_argument('--config', default='target', help="target (default) | c2 | c3 | c4 | c5: workload of this run; the default run "
                    "carries all BASELINE configs in its `configs` block")
    ap.add_argument('--no-configs', action='store_true', help='skip the configs block and the load-sequence key')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    if args.config != 'target':
        # one BASELINE config as the workload of this run (single GPU; rank 0 only under torchrun)
        if rank == 0:
            peaks = profile_json('../MEASURED_PEAKS.json')
            peak = float(peaks.get('hbm_gbs', 6650.))
            sampler = ClockSampler(local)
            sampler.start()
            c = run_kernel_config(args.config, local, max(args.steps, 1), peak, args.opt)
            line = {'metric': METRIC, 'value': c['elements_per_s'], 'unit': 'elements/s', 'n_gpus': 1, 'steps': c['reps'],
                    'warmup': 3, 'ms_per_step': c['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
                    'dtype': 'f64', 'data': 'synthetic', 'config': {'workload': c['workload'], 'elements': c['elements'],
                                                                   'l2': 'flushed before every repetition'},
                    'roofline': c['roofline'], 'roofline_fp64': c.get('roofline_fp64'), 'branch_mix': c['branch_mix'],
                    'updates_per_s': c['updates_per_s'], 'assembly_elems_per_s': c['assembly_elems_per_s'], 'cpu_baseline': None,
                    'e2e': None, 'gpu_launches': c['gpu_launches'], 'clocks': sampler.stop()}
            print(json.dumps(line))
        return

    plan = shard_plan(args.mesh, rank, world)
    group = RankGroup(world, local, backend='gloo' if args.dist_selftest else 'nccl')
    barrier, max_over_ranks = group.barrier, group.max_over_ranks
    if args.dist_selftest:
        # CPU-only check of the multi-rank plumbing (tests/test_dist_gloo.py): every rank plans its block, the ranks
        # meet at a barrier and reduce a per-rank time with MAX and the element count with SUM.
        barrier()
        t = max_over_ranks(float(rank + 1))
        total = group.sum_over_ranks(float(plan['elements']))
        if rank == 0:
            print(json.dumps({'selftest': True, 'world': world, 'max_time': t, 'total_elements': total,
                              'seed_rank0': plan['seed'], 'scaling': 'weak'}))
        group.close()
        return

    from pylabfea_b200 import _lib
    ctx = _lib.Context(local)
    ctx.set_option('compact', args.compact)
    ctx.set_option('assembly_tiled', args.assembly_tiled)
    for kv in args.opt:
        k, v = kv.split('=')
        ctx.set_option(k, int(v))
    N = args.mesh
    NXl = plan['nx_local']                 # N owned element columns (+ 1 halo column on ranks with a left neighbour)
    Nel = NXl * N                          # elements this rank processes per step
    owned = N * N                          # elements this rank contributes to `value`
    mat, CV = j2_material()
    ctx.set_materials([mat], planestress=False)
    lx = 4. / N
    ctx.mesh_structured(NXl, N, np.full(NXl, lx), lx, 1., np.zeros(Nel, np.int32))
    sig, epl, deps = synth_states(N, plan['seed'])
    if plan['halo']:                       # halo column: any state will do for the timing, re-use the first column's
        sig, epl, deps = (np.ascontiguousarray(np.concatenate([a[:N], a])) for a in (sig, epl, deps))
    b = ctx.batch_create(Nel)
    ctx.batch_upload(b, sig, epl, deps)
    ctx.synchronize()

    def step():
        ctx.batch_run(b)
        ctx.setupK_from_batch(b)

    l0 = ctx.launch_count()
    for _ in range(max(args.warmup, 3)):
        step()
    ctx.synchronize()
    launches_per_step = (ctx.launch_count() - l0) / max(args.warmup, 3)

    # clocks: nvidia-smi needs ~100 ms per sample and the timed region can be shorter than that, so the same step is
    # first run back to back for >= 1.2 s with the sampler on (untimed, identical load), then the timed region follows
    # with the sampler still running; all samples are taken under this kernel mix.
    sampler = ClockSampler(local)
    sampler.start()
    t_probe = time.perf_counter()
    while time.perf_counter() - t_probe < 1.2:
        for _ in range(20):
            step()
        ctx.synchronize()
    K = args.steps
    barrier()
    ctx.synchronize()
    ctx.last_kernels()            # start recording which kernel variants the timed region launches
    t_wall0 = time.perf_counter()
    l0 = ctx.launch_count()
    ctx.event_record(0)
    for i in range(K):
        ctx.event_record(1 + 3 * i)
        ctx.batch_run(b)
        ctx.event_record(2 + 3 * i)
        ctx.setupK_from_batch(b)
        ctx.event_record(3 + 3 * i)
    ctx.event_record(3 * K + 1)
    ctx.synchronize()
    barrier()
    total_ms = ctx.event_elapsed(0, 3 * K + 1)
    t_wall = time.perf_counter() - t_wall0
    gpu_launches = ctx.launch_count() - l0
    launched = sorted(set(ctx.last_kernels()))
    clocks = sampler.stop()
    resp_ms = float(np.mean([ctx.event_elapsed(1 + 3 * i, 2 + 3 * i) for i in range(K)]))
    asm_ms = float(np.mean([ctx.event_elapsed(2 + 3 * i, 3 + 3 * i) for i in range(K)]))
    total_ms = max_over_ranks(total_ms)
    resp_ms_max = max_over_ranks(resp_ms)
    asm_ms_max = max_over_ranks(asm_ms)
    ms_per_step = total_ms / K
    value = world * owned / (ms_per_step * 1e-3)

    # branch census of the workload (one extra, untimed run)
    ns = np.empty(Nel, dtype=np.int32)
    ct0 = np.empty((Nel, 6, 6))
    ctx.batch_download(b, ctan=ct0, nsteps=ns)
    n50 = int(np.sum(ns == 49))
    nel_elastic = int(np.sum(np.all(ct0.reshape(Nel, 36) == CV.ravel()[None, :], axis=1)))
    if world > 1:
        workload_extra = {'partition': 'strips of %d element columns of one %dx%d mesh, 1 halo column on ranks > 0' % (N, world * N, N)}
    else:
        workload_extra = {}
    del ct0

    # ---- e2e through the host-buffer C-ABI calls -----------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        hs, he, hd = (_lib.pinned_empty((Nel, 6)) for _ in range(3))
        hs[:], he[:], hd[:] = sig, epl, deps
        o_fy = _lib.pinned_empty((Nel,))
        o_sig = _lib.pinned_empty((Nel, 6))
        o_dp = _lib.pinned_empty((Nel, 6))
        o_ct = _lib.pinned_empty((Nel, 36))
        o_ns = _lib.pinned_empty((Nel,), np.int32)
        o_val = _lib.pinned_empty((ctx.nnz,))

        o_c21 = _lib.pinned_empty((Nel * 21,))
        o_bits = _lib.pinned_empty(((Nel + 31) // 32,), np.uint32)
        state = {'packed': 0}

        def timed(fn, reps):
            for _ in range(2):
                fn()
            barrier()
            t0 = time.perf_counter()
            ctx.event_record(4000)
            for _ in range(reps):
                fn()
            ctx.event_record(4001)
            ctx.synchronize()
            barrier()
            return max_over_ranks(max(ctx.event_elapsed(4000, 4001), (time.perf_counter() - t0) * 1e3)) / reps

        # headline: one call, host buffers in and out, results in compact form -- what Model.solve's inner pass consumes on
        # the host (fy, sig, depl, nsteps per element; grad_stiff as upper triangles of the non-elastic elements + one
        # elastic bit per element); the assembled K stays in HBM for the on-device solve (north star: the Newton loop never
        # leaves the GPU).  Upload, kernels and download are pipelined over chunks of element columns on three streams.
        def e2e_step():
            state['packed'] = ctx.step_host_packed(b, hs, he, hd, o_c21, o_bits, fy=o_fy, sig_out=o_sig, depl=o_dp, nsteps=o_ns)
        ke = max(3, min(K, 10))
        e2e_ms = timed(e2e_step, ke)
        h2d = Nel * 3 * 48
        d2h = Nel * (8 + 48 + 48 + 4) + state['packed'] * 168 + ((Nel + 31) // 32) * 4
        e2e = {'value': world * owned / (e2e_ms * 1e-3), 'unit': 'elements/s', 'h2d_bytes_per_step': int(h2d),
               'd2h_bytes_per_step': int(d2h), 'ms_per_step': e2e_ms, 'packed_tangents': int(state['packed']),
               'api': 'plf_step_host_packed: (Nel,6) sig/epl/deps in; fy, sig, depl, nsteps, elastic bits and the upper triangles of '
                      'the non-elastic tangents out; K assembled and left on the device for plf_solve_linear (pinned host buffers, '
                      'chunked three-stream pipeline)'}

        # round-1 definition for continuity: everything in full form, CSR(K) values included (plf_step_host)
        def full_step():
            ctx.step_host(b, hs, he, hd, fy=o_fy, sig_out=o_sig, depl=o_dp, ctan=o_ct, nsteps=o_ns, kval=o_val)
        full_ms = timed(full_step, ke)
        e2e['full_copy'] = {'value': world * owned / (full_ms * 1e-3), 'ms_per_step': full_ms, 'h2d_bytes_per_step': int(h2d),
                            'd2h_bytes_per_step': int(Nel * (8 + 48 + 48 + 288 + 4) + ctx.nnz * 8),
                            'note': 'plf_step_host: full 6x6 tangents of all elements and the CSR(K) values copied back (the round-1 e2e figure)'}

        def packed_k_step():
            ctx.step_host_packed(b, hs, he, hd, o_c21, o_bits, fy=o_fy, sig_out=o_sig, depl=o_dp, nsteps=o_ns, kval=o_val)
        pk_ms = timed(packed_k_step, ke)
        e2e['packed_with_K'] = {'value': world * owned / (pk_ms * 1e-3), 'ms_per_step': pk_ms,
                                'd2h_bytes_per_step': int(d2h + ctx.nnz * 8), 'note': 'compact results plus the CSR(K) values'}

    coupled = None
    if not args.no_configs:
        try:
            coupled = run_coupled_solve(group, rank, world, local)
        except Exception as exc:
            coupled = {'error': '%s: %s' % (type(exc).__name__, exc)}

    if rank != 0:
        group.close()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.))

    def ncu_traffic(prefixes):
        """DRAM bytes per launch from the committed ncu capture (profiles/traffic.json, scripts/summarize_ncu.py) of EXACTLY
        the kernel variants this run launched (plf_last_kernels, template arguments included); None when a launched variant
        has no capture or the mesh is not the one the capture was taken at (1024 x 1024)."""
        tj = profile_json('traffic.json')
        names = [k for k in launched if k.split('<')[0] in prefixes]
        if not names or N != 1024:
            return None, names
        tot = 0.
        for k in names:
            if k not in tj:
                return None, names
            tot += tj[k]['dram_bytes_per_launch']
        return tot, names
    achieved = BYTES_PER_UPDATE * Nel / (resp_ms_max * 1e-3) / 1e9
    asm_achieved = BYTES_PER_ASSEMBLED * Nel / (asm_ms_max * 1e-3) / 1e9
    asm_kernel = {0: 'assemble_rows_kernel', 2: 'assemble_rows_staged_kernel'}[args.assembly_tiled]
    resp_traffic, resp_names = ncu_traffic(('response_passA_kernel', 'response_passA_bulk_kernel', 'response_passB_kernel'))
    asm_traffic, asm_names = ncu_traffic(('assemble_rows_staged_kernel',))
    roofline = {'bound': 'hbm',
                'kernel': ' + '.join(resp_names) if resp_names else 'response_generic_kernel',
                'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                'peak_source': 'MEASURED_PEAKS.json hbm_gbs (of measured)' if peaks else 'fallback 6650 GB/s (of fallback)',
                'algorithmic_bytes': BYTES_PER_UPDATE * Nel,
                'traffic': resp_traffic,
                'bytes_per_update': BYTES_PER_UPDATE, 'ms': resp_ms_max,
                'assembly': {'kernel': ' + '.join(asm_names) if asm_names else asm_kernel, 'achieved': asm_achieved, 'frac': asm_achieved / peak,
                             'bytes_per_element': BYTES_PER_ASSEMBLED, 'algorithmic_bytes': BYTES_PER_ASSEMBLED * Nel,
                             'traffic': asm_traffic, 'ms': asm_ms_max}}

    cpu = None
    if not args.no_cpu:
        from oracle import oracle as orc
        sample = args.cpu_sample
        idx = np.random.default_rng(7).choice(Nel, sample, replace=False)
        mats = [orc.Mat(kind=orc.HILL, E=E_MOD, nu=NU, sy=SY, khard=KHARD)]
        B4 = orc.gauss_bmats(lx, lx, 0, CV, E_MOD, NU)
        jacw = (lx * lx * 1.) * 4.
        oracle_step(orc, mats, CV, sig[idx[:1000]], epl[idx[:1000]], deps[idx[:1000]], B4, jacw, 1)
        t0 = time.perf_counter()
        oracle_step(orc, mats, CV, sig[idx], epl[idx], deps[idx], B4, jacw, 1)
        dt = time.perf_counter() - t0
        cpu = {'value': sample / dt, 'unit': 'elements/s', 'cores': 1, 'kind': 'port',
               'sample': '%d elements sampled (seeded) from the same %dx%d workload, oracle/plf_oracle.c, 1 thread, %.1f s'
                         % (sample, N, N, dt)}

    line = {'metric': METRIC, 'value': value, 'unit': 'elements/s', 'n_gpus': world, 'steps': K, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'config': dict(workload_config(N), **workload_extra),
            'updates_per_s': world * owned / (resp_ms_max * 1e-3), 'assembly_elems_per_s': world * owned / (asm_ms_max * 1e-3),
            'branch_mix': {'elastic': nel_elastic / Nel, 'substep50': n50 / Nel, 'plastic_1step': 1. - (nel_elastic + n50) / Nel},
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e, 'gpu_launches': int(gpu_launches),
            'launches_per_step': launches_per_step, 'clocks': clocks, 'wall_s_timed_region': t_wall,
            'device': ctx.device_info()['name']}
    line['coupled_solve'] = coupled
    if world == 1 and not args.no_configs:
        # every BASELINE config next to the headline, and a whole load sequence through Model.solve (the path that keeps
        # an increment on the device: predictor + calc_scf + inner passes + commit per increment)
        ctx.synchronize()
        line['configs'] = configs_block(local, peak, quick=args.steps < 50)
        try:
            seq, fe_seq = run_load_sequence(N, 5, 0.004, local)
            line['e2e_load_step'] = dict(seq, unit='ms per load increment', value=seq['ms_per_increment'],
                                         cpu_reference=cpu_load_step_estimate(seq) if not args.no_cpu else None)
            del fe_seq
            line['harvested_mix'] = harvested_step(N, 5, 0.004, local, peak)
        except Exception as exc:
            line['e2e_load_step'] = {'error': '%s: %s' % (type(exc).__name__, exc)}
    print(json.dumps(line))
    group.close()


if __name__ == '__main__':
    main()
