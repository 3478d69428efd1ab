"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C-ABI, against
 (a) the golden fixtures produced by the unmodified reference and (b) the CPU oracle on fresh seeded inputs.

Tolerances: CSR(K), Kel, B: bit-exact (np.array_equal).  Material.response: 1e-6 relative (BASELINE.json
north_star); measured agreement is ~1e-12 for the analytic models."""
import numpy as np
import pytest

from helpers import plf_material_from_golden, elastic_material, relmax

pytestmark = pytest.mark.gpu

RESP_MATS = ['j2', 'j2pp', 'hill', 'hillrv', 'drucker', 'cubic']
TOL = 1e-6
PASSB_RECUR_DEFAULT = int(__import__('os').environ.get('PLF_PASSB_RECUR', '1'))   # library default of the option


@pytest.fixture(scope='module')
def ctx():
    from pylabfea_b200 import _lib
    c = _lib.Context(0)
    yield c
    c.close()


def check_response(out, g, tag, sy, tol=TOL):
    fy, so, dp, ct, ns = out
    assert np.array_equal(ns, g[tag + '_nsteps'])
    assert np.max(np.abs(fy - g[tag + '_fy'])) / sy < tol
    for i in range(len(fy)):
        assert relmax(so[i], g[tag + '_sig_out'][i]) < tol, i
        assert np.max(np.abs(dp[i] - g[tag + '_depl'][i])) < tol * max(np.max(np.abs(g[tag + '_depl'][i])), 1e-4), i
        assert relmax(ct[i], g[tag + '_ctan'][i]) < tol, i


@pytest.mark.parametrize('compact', [0, 1])
@pytest.mark.parametrize('tag', ['pe', 'ps'])
@pytest.mark.parametrize('name', RESP_MATS)
def test_response_vs_reference_golden(ctx, golden, name, tag, compact):
    g = golden('response_' + name)
    ctx.set_option('compact', compact)
    ctx.set_materials([plf_material_from_golden(g)], planestress=(tag == 'ps'))
    assert np.array_equal(ctx.element_CV(0), g[tag + '_CV'])
    out = ctx.response_batch(g[tag + '_sig'], g[tag + '_epl'], g[tag + '_deps'])
    check_response(out, g, tag, float(g['sy']))
    ctx.set_option('compact', 1)


@pytest.mark.parametrize('tag', ['pe', 'ps'])
@pytest.mark.parametrize('name', RESP_MATS)
def test_primitives_vs_reference_golden(ctx, golden, name, tag):
    g = golden('response_' + name)
    ctx.set_materials([plf_material_from_golden(g)], planestress=(tag == 'ps'))
    sig, epl, deps = g[tag + '_sig'], g[tag + '_epl'], g[tag + '_deps']
    assert relmax(ctx.calc_seq(0, sig), g[tag + '_seq']) < 1e-12
    assert relmax(ctx.calc_yf(0, sig, epl), g[tag + '_yf']) < 1e-11
    assert relmax(ctx.calc_fgrad(0, sig, epl), g[tag + '_fgrad']) < 1e-11
    assert relmax(ctx.epl_dot(0, sig, epl, deps), g[tag + '_epl_dot']) < 1e-10
    assert relmax(ctx.C_tan(0, sig, epl), g[tag + '_C_tan']) < 1e-11


def test_response_mixed_materials_vs_oracle(ctx, golden):
    """Several materials in one batch, ragged size, fresh random states: CUDA vs the CPU oracle."""
    from oracle import oracle as orc
    gs = [golden('response_' + n) for n in ('j2', 'hill', 'drucker')]
    ctx.set_materials([plf_material_from_golden(g) for g in gs], planestress=False)
    rng = np.random.default_rng(123)
    n = 4099   # not a multiple of the block size
    mid = rng.integers(0, 3, n).astype(np.int32)
    sig = np.zeros((n, 6)); epl = np.zeros((n, 6)); deps = np.zeros((n, 6))
    for k, g in enumerate(gs):
        idx = np.nonzero(mid == k)[0]
        src = rng.integers(0, len(g['pe_sig']), len(idx))
        sig[idx] = g['pe_sig'][src] * rng.uniform(0.9, 1.02, (len(idx), 1))
        epl[idx] = g['pe_epl'][src]
        deps[idx] = g['pe_deps'][rng.integers(0, len(g['pe_sig']), len(idx))] * rng.uniform(0.2, 1.5, (len(idx), 1))
    mats = [orc.Mat.from_golden(g) for g in gs]
    CVs = np.array([g['pe_CV'] for g in gs])
    ref = orc.response_batch(mats, mid, sig, epl, deps, CVs)
    for compact in (0, 1):
        ctx.set_option('compact', compact)
        fy, so, dp, ct, ns = ctx.response_batch(sig, epl, deps, mat_id=mid)
        assert np.array_equal(ns, ref[4])
        assert np.max(np.abs(fy - ref[0])) < TOL * 150
        assert relmax(so, ref[1]) < TOL and relmax(ct, ref[3]) < TOL
        assert np.max(np.abs(dp - ref[2])) < TOL * np.max(np.abs(ref[2]))
    ctx.set_option('compact', 1)


def test_response_edge_cases(ctx, golden):
    g = golden('response_j2')
    ctx.set_materials([plf_material_from_golden(g)], planestress=False)
    # empty batch
    out = ctx.response_batch(np.zeros((0, 6)), np.zeros((0, 6)), np.zeros((0, 6)))
    assert out[0].shape == (0,) and out[3].shape == (0, 6, 6)
    # single element, zero increment from zero stress: purely elastic, tangent = CV
    fy, so, dp, ct, ns = ctx.response_batch(np.zeros((1, 6)), np.zeros((1, 6)), np.zeros((1, 6)))
    assert np.array_equal(ct[0], g['pe_CV']) and np.all(so == 0) and np.all(dp == 0) and ns[0] == 0
    assert fy[0] == -float(g['sy'])
    # shape errors raise ValueError like material.py:236-239
    with pytest.raises(ValueError):
        ctx.response_batch(np.zeros((2, 6)), np.zeros((3, 6)), np.zeros((2, 6)))
    # material id out of range is rejected by the mesh call, not silently used
    with pytest.raises(ValueError):
        ctx.mesh_structured(2, 2, [1., 1.], 1., 1., np.full(4, 7))


def test_svc_vs_reference_golden(ctx, golden):
    g = golden('svc_hill')
    ctx.set_materials([plf_material_from_golden(g)], planestress=False)
    yf = ctx.calc_yf(0, g['yf_sig'])
    assert np.max(np.abs(yf - g['yf'])) < 1e-10          # warp-tree sum vs libsvm's sequential sum
    fg = ctx.calc_fgrad(0, g['yf_sig'])
    assert relmax(fg, g['fgrad']) < 1e-10
    n = len(g['ML_full_yf'])
    fyf = ctx.ml_full_yf(0, g['yf_sig'][:n])
    assert np.max(np.abs(fyf - g['ML_full_yf'])) < TOL * float(g['sy'])


@pytest.mark.parametrize('tag', ['pe', 'ps'])
def test_svc_response_vs_reference_golden(ctx, golden, tag):
    g = golden('svc_hill')
    ctx.set_materials([plf_material_from_golden(g)], planestress=(tag == 'ps'))
    out = ctx.response_batch(g[tag + '_sig'], g[tag + '_epl'], g[tag + '_deps'])
    # north-star tolerance 1e-6; measured on these 2 x 48 tuples: 5e-13 (stress) .. 2e-12 (tangent), scripts/svc_deviation.py
    check_response(out, g, tag, float(g['sy']), tol=1e-6)


@pytest.mark.parametrize('tag', ['pe', 'ps'])
def test_svc_sdim3_vs_reference_golden(ctx, golden, tag):
    """sdim = 3 machine-learned yield function (SVC on the cylindrical coordinates of the principal stresses, trained by
    the reference as in tests/test_ml.py:10-28): decision function, ML_full_yf, the Jacobian-chain gradient
    (material.py:780-805), epl_dot / C_tan and Material.response against the reference on shear-free states."""
    g = golden('svc3_mlplast')
    ctx.set_materials([plf_material_from_golden(g)], planestress=(tag == 'ps'))
    assert np.array_equal(ctx.element_CV(0), g[tag + '_CV'])
    sig, epl, deps = g[tag + '_sig'], g[tag + '_epl'], g[tag + '_deps']
    fg = ctx.calc_fgrad(0, sig, epl)
    dev = {'seq': relmax(ctx.calc_seq(0, sig), g[tag + '_seq']),
           'yf': np.max(np.abs(ctx.calc_yf(0, sig, epl) - g[tag + '_yf'])),
           'full_yf': np.max(np.abs(ctx.ml_full_yf(0, sig, epl) - g[tag + '_full_yf'])) / float(g['sy']),
           'fgrad': relmax(fg[:, 0:3], g[tag + '_fgrad3']),
           'epl_dot': relmax(ctx.epl_dot(0, sig, epl, deps), g[tag + '_epl_dot']),
           'C_tan': relmax(ctx.C_tan(0, sig, epl), g[tag + '_C_tan'])}
    print('svc3', tag, {k: float(v) for k, v in dev.items()})
    assert dev['seq'] < 1e-12 and dev['yf'] < 1e-10 and dev['full_yf'] < TOL, dev
    assert dev['fgrad'] < 1e-10 and np.all(fg[:, 3:6] == 0.), dev
    assert dev['epl_dot'] < 1e-8 and dev['C_tan'] < 1e-8, dev
    check_response(ctx.response_batch(sig, epl, deps), g, tag, float(g['sy']), tol=1e-6)


@pytest.mark.parametrize('tag', ['pe', 'ps'])
def test_lhs_asymmetry_device_vs_oracle(ctx, golden, tag):
    """LHS yield asymmetry (d = lhs in calc_seq 646-650 and calc_fgrad 829-830).  The reference itself raises in calc_seq
    for every LHS material (`if self.lhs:` on an array, material.py:646), so there is no fixture: the device is checked
    against the oracle's restatement on the states of the Drucker fixture (same elastic and Hill parameters)."""
    from pylabfea_b200 import _lib
    from oracle import oracle as orc
    g = golden('response_drucker')
    lhs = [0.05, -0.03, 0.08]
    m = plf_material_from_golden(g)
    m.lhs = 1
    m.d[:] = lhs
    ctx.set_materials([m], planestress=(tag == 'ps'))
    om = orc.Mat(kind=orc.HILL, E=float(g['E']), nu=float(g['nu']), sy=float(g['sy']), khard=float(g['khard']), hill=g['hill'], lhs=lhs)
    sig, epl, deps, CV = g[tag + '_sig'], g[tag + '_epl'], g[tag + '_deps'], g[tag + '_CV']
    assert relmax(ctx.calc_seq(0, sig), orc.calc_seq(om, sig)) < 1e-12
    assert relmax(ctx.calc_fgrad(0, sig, epl), orc.calc_fgrad(om, sig, epl)) < 1e-11
    fy, so, dp, ct, ns = ctx.response_batch(sig, epl, deps)
    ofy, oso, odp, oct_, ons = orc.response_batch([om], None, sig, epl, deps, CV[None])
    assert np.array_equal(ns, ons) and int(np.sum(ns == 49)) > 10
    assert np.max(np.abs(fy - ofy)) < TOL * float(g['sy'])
    assert relmax(so, oso) < TOL and relmax(ct, oct_) < TOL
    assert np.max(np.abs(dp - odp)) < TOL * np.max(np.abs(odp))


def barlat_plf_material(g, khard=0.):
    from pylabfea_b200 import _lib
    m = _lib.PlfMaterial()
    m.kind = _lib.MAT_BARLAT
    m.E, m.nu, m.sy, m.khard = float(g['E']), float(g['nu']), float(g['sy']), float(khard)
    m.hill[:] = [1.] * 6
    m.barlat_m1[:] = list(g['Bar_m1'].ravel())
    m.barlat_m2[:] = list(g['Bar_m2'].ravel())
    m.barlat_exp = float(g['barlat_exp'])
    return m


def test_barlat_seq_vs_reference_golden(ctx, golden):
    g = golden('barlat')
    m = barlat_plf_material(g)
    m.CV[:] = list(np.eye(6).ravel())
    ctx.set_materials([m], planestress=False)
    assert relmax(ctx.calc_seq(0, g['sig']), g['seq']) < 1e-11


@pytest.mark.parametrize('kt', ['pp', 'kh'])
@pytest.mark.parametrize('tag', ['pe', 'ps'])
def test_barlat_gradient_and_response_vs_reference_golden(ctx, golden, kt, tag):
    """Yld2004-18p: analytic device gradient (barlat_seq_grad) and return mapping against the reference return mapping
    driven by the numerical derivative of the reference's own calc_seq (the reference has no Barlat gradient,
    material.py:822-823; fixture: tests/golden/make_golden.py BarlatFD).  1e-6 relative as BASELINE.json asks."""
    g = golden('barlat_response')
    t = kt + '_' + tag
    m = barlat_plf_material(g, float(g[kt + '_khard']))
    m.CV[:] = list(np.ascontiguousarray(g[kt + '_pe_CV']).ravel())
    m.C44 = float(g[kt + '_pe_CV'][3, 3])
    ctx.set_materials([m], planestress=(tag == 'ps'))
    assert np.array_equal(ctx.element_CV(0), g[t + '_CV'])
    sig, epl, deps = g[t + '_sig'], g[t + '_epl'], g[t + '_deps']
    assert relmax(ctx.calc_seq(0, sig), g[t + '_seq']) < 1e-11
    assert relmax(ctx.calc_fgrad(0, sig, epl), g[t + '_fgrad']) < 1e-7      # analytic vs 4th-order differences
    assert relmax(ctx.epl_dot(0, sig, epl, deps), g[t + '_epl_dot']) < 1e-6
    assert relmax(ctx.C_tan(0, sig, epl), g[t + '_C_tan']) < 1e-6
    out = ctx.response_batch(sig, epl, deps)
    check_response(out, g, t, float(g['sy']))


def _oracle_sample(mats, CV, sig, epl, deps, idx, threads=8):
    from oracle import oracle as orc
    return orc.response_batch_threads(mats, None, sig[idx], epl[idx], deps[idx], CV[None], threads)


def test_config4_scale_svc_response_vs_oracle(ctx, golden):
    """BASELINE config 4 at scale: 16 384 elements of the 512 x 512 synthetic field through the SVC return mapping (1 585
    support vectors, warp per element, dynamic queue), the CPU oracle on 4 096 of them (seeded sample; the oracle needs
    ~10 ms per update): same branch per element, stress / plastic strain / tangent within the north star's 1e-6, and
    launch-to-launch reproducibility of the queue-scheduled kernel."""
    import bench
    from oracle import oracle as orc
    g = golden('svc_hill')
    ctx.set_materials([plf_material_from_golden(g)], planestress=False)
    CV = ctx.element_CV(0)
    sig, epl, deps = bench.synth_states(512, 1234, sy=float(g['sy']), khard=0., deps_scale=1. / 3.)
    n = 16384
    sig, epl, deps = sig[:n].copy(), epl[:n].copy(), deps[:n].copy()
    out = ctx.response_batch(sig, epl, deps)
    again = ctx.response_batch(sig, epl, deps)
    for a, b in zip(out, again):
        assert np.array_equal(a, b)
    ns = out[4]
    assert np.sum(ns == 49) > 0.05 * n and np.sum(ns == 0) > 0.3 * n
    idx = np.random.default_rng(4).choice(n, 4096, replace=False)
    ref = _oracle_sample([orc.Mat.from_golden(g)], CV, sig, epl, deps, idx)
    fy, so, dp, ct, _ = (o[idx] for o in out)
    assert np.array_equal(ns[idx], ref[4])
    assert np.max(np.abs(fy - ref[0])) < TOL * float(g['sy'])
    assert relmax(so, ref[1]) < TOL and relmax(ct, ref[3]) < TOL
    assert np.max(np.abs(dp - ref[2])) < TOL * np.max(np.abs(ref[2]))


def test_config5_scale_barlat_response_vs_oracle(ctx, golden):
    """BASELINE config 5 at scale: all 4 194 304 elements of the 2048 x 2048 synthetic field through the two-pass Barlat
    return mapping (analytic gradient, whole-number exponent path), against the unsplit kernel on every element and against
    the CPU oracle (numerically differentiated reference yield function, the recipe of the reference fixture) on a seeded
    sample: same branch per element; 1e-6 as BASELINE.json asks."""
    import bench
    from oracle import oracle as orc
    g = golden('barlat_response')
    m = barlat_plf_material(g, 0.)
    CV = np.ascontiguousarray(g['pp_pe_CV'])
    m.CV[:] = list(CV.ravel())
    m.C44 = float(CV[3, 3])
    ctx.set_materials([m], planestress=False)
    N = 2048
    sig, epl, deps = bench.synth_states(N, 1234, sy=float(g['sy']), khard=0., deps_scale=1. / 3.)
    try:
        out = ctx.response_batch(sig, epl, deps)
        ctx.set_option('compact', 0)
        ref_dev = ctx.response_batch(sig[:400000], epl[:400000], deps[:400000])
    finally:
        ctx.set_option('compact', 1)
    ns = out[4]
    assert 0.05 * N * N < np.sum(ns == 49) < 0.5 * N * N
    for a, b in zip(out, ref_dev):          # two passes == one pass, bit for bit (same device functions)
        assert np.array_equal(a[:400000], b)
    idx = np.random.default_rng(6).choice(N * N, 40000, replace=False)
    om = orc.Mat(kind=orc.BARLAT, E=float(g['E']), nu=float(g['nu']), sy=float(g['sy']), khard=0.,
                 barlat=dict(Bar_m1=g['Bar_m1'], Bar_m2=g['Bar_m2'], barlat_exp=g['barlat_exp']))
    ref = _oracle_sample([om], CV, sig, epl, deps, idx)
    fy, so, dp, ct, _ = (o[idx] for o in out)
    assert np.array_equal(ns[idx], ref[4])
    assert np.max(np.abs(fy - ref[0])) < TOL * float(g['sy'])
    assert relmax(so, ref[1]) < TOL and relmax(ct, ref[3]) < 10 * TOL      # tangent: analytic vs 4th-order differences
    assert np.max(np.abs(dp - ref[2])) < TOL * np.max(np.abs(ref[2]))


# ------------------------------------------------------------------------------------------------------------
# K path: bit-exact
# ------------------------------------------------------------------------------------------------------------
KCASES = ['cantilever10', 'laminate16x4', 'inclusion18', 'uneven4x5_plastic', 'hill6_ps_plastic',
          'inclusion9_plastic', 'single']


def setup_mesh_from_golden(ctx, g):
    from pylabfea_b200 import _lib
    nm = len(g['mat_E'])
    mats = []
    for i in range(nm):
        mats.append(elastic_material(g['mat_E'][i], g['mat_nu'][i], g['mat_CV'][i]))
    ps = int(g['planestress'])
    ctx.set_materials(mats, planestress=ps)
    NX, NY = int(g['NX']), int(g['NY'])
    dx = g['lx'].reshape(NX, NY)[:, 0]
    ctx.mesh_structured(NX, NY, dx, float(g['ly'][0]), float(g['thick']), g['mat_id'])
    return NX, NY


@pytest.mark.parametrize('tiled', [2, 0])
@pytest.mark.parametrize('case', KCASES)
def test_K_bit_exact_vs_reference_golden(ctx, golden, case, tiled):
    g = golden('kmat_' + case)
    ctx.set_option('assembly_tiled', tiled)
    NX, NY = setup_mesh_from_golden(ctx, g)
    for e in range(len(g['Bmat'])):
        assert np.array_equal(ctx.get_bmat(e), g['Bmat'][e]), 'B must be bit-exact'
    ctx.set_tangent(g['elstiff'])
    assert np.array_equal(ctx.get_kel(), g['Kel']), 'Kel must be bit-exact'
    ctx.setupK()
    rowptr, col, val = ctx.get_csr()
    Ndof = int(g['Ndof'])
    assert ctx.Ndof == Ndof and rowptr[-1] == ctx.nnz == 4 * (3 * NX + 1) * (3 * NY + 1)
    # CSR well-formed: sorted, unique columns per row
    for r in range(Ndof):
        c = col[rowptr[r]:rowptr[r + 1]]
        assert np.all(np.diff(c) > 0)
    K = np.zeros((Ndof, Ndof))
    rows = np.repeat(np.arange(Ndof), np.diff(rowptr))
    K[rows, col] = val
    Kref = np.zeros((Ndof, Ndof))
    Kref[g['K_row'], g['K_col']] = g['K_val']
    assert np.array_equal(K, Kref), 'assembled K must be bit-exact (max diff %g)' % np.max(np.abs(K - Kref))
    ctx.set_option('assembly_tiled', 2)


@pytest.mark.parametrize('shape', [(256, 256), (37, 95), (130, 3), (1, 70)])
def test_K_tiled_equals_direct_on_ragged_meshes(ctx, golden, shape):
    """Tile-boundary cases (mesh sizes that are not multiples of the 4 x 32 node tile, one-element-wide meshes,
    plane stress): the staged assembly must reproduce the direct row-per-thread kernel bit for bit."""
    g = golden('response_hill')
    NX, NY = shape
    rng = np.random.default_rng(NX * 1000 + NY)
    for ps in (0, 1):
        ctx.set_materials([plf_material_from_golden(g)], planestress=ps)
        ctx.mesh_structured(NX, NY, rng.choice([0.5, 1.0, 1.5], NX), 0.7, 1.3, np.zeros(NX * NY, np.int32))
        tag = 'ps' if ps else 'pe'
        ctx.set_tangent(g[tag + '_ctan'][rng.integers(0, len(g[tag + '_ctan']), NX * NY)])
        vals = []
        for tiled in (0, 2):
            ctx.set_option('assembly_tiled', tiled)
            ctx.setupK()
            vals.append(ctx.get_csr()[2])
        assert np.array_equal(vals[0], vals[1])
    ctx.set_option('assembly_tiled', 2)


def test_K_large_mesh_vs_oracle(ctx, golden):
    """256x256 (BASELINE config 2 size) with random plastic tangents: CSR values bit-exact vs the oracle's
    ascending-element assembly; plus size-independent properties (symmetry to rounding, zero row sums)."""
    from oracle import oracle as orc
    g = golden('response_j2')
    N = 256
    ctx.set_materials([plf_material_from_golden(g)], planestress=False)
    ctx.mesh_structured(N, N, np.full(N, 4. / N), 4. / N, 1., np.zeros(N * N, np.int32))
    rng = np.random.default_rng(5)
    T = g['pe_ctan'][rng.integers(0, len(g['pe_ctan']), N * N)]
    ctx.set_tangent(T)
    ctx.setupK()
    rowptr, col, val = ctx.get_csr()
    CV = g['pe_CV']
    B4 = orc.gauss_bmats(4. / N, 4. / N, 0, CV, float(g['E']), float(g['nu']))
    assert np.array_equal(ctx.get_bmat(0), B4)
    jacw = ((4. / N) * (4. / N) * 1.) * 4.
    idx = rng.integers(0, N * N, 200)
    kel_gpu = ctx.get_kel()
    for e in idx:
        assert np.array_equal(kel_gpu[e], orc.kel(B4, T[e], jacw))
    j, k = np.divmod(np.arange(N * N), N)
    n1 = j * (N + 1) + k
    nodes = np.stack([n1, n1 + 1, n1 + N + 1, n1 + N + 2], axis=1).astype(np.int32)
    ref = orc.setupK_csr(nodes, kel_gpu, rowptr, col)
    assert np.array_equal(val, ref)
    # rigid translation is in the null space of K: row sums over the x (and y) columns vanish to rounding
    rows = np.repeat(np.arange(ctx.Ndof), np.diff(rowptr))
    sx = np.bincount(rows, weights=val * (col % 2 == 0), minlength=ctx.Ndof)
    assert np.max(np.abs(sx)) < 1e-9 * np.max(np.abs(val))


def test_response_full_size_fast_modes_vs_oracle(ctx):
    """BASELINE target size (1024 x 1024 J2 on isotropic elasticity, synthetic correlated field): the production fast
    mode (planar isotropic pass B, two-ended work list) against the general Hill kernels on every element, against the
    CPU oracle on a seeded sample, launch-to-launch reproducibility (the in-kernel counter reset), and a batch in which
    30 % of the elements carry out-of-plane shear (back of the work list)."""
    import bench
    from oracle import oracle as orc
    N = 1024
    mat, CV = bench.j2_material()
    ctx.set_materials([mat], planestress=False)
    sig, epl, deps = bench.synth_states(N, 1234)
    mats = [orc.Mat(kind=orc.HILL, E=bench.E_MOD, nu=bench.NU, sy=bench.SY, khard=bench.KHARD)]

    def run(a, b, c, **opts):
        o = dict(compact=1, max_fm=4, passa_direct=0, passb_recur=0)
        o.update(opts)
        for k, v in o.items():
            ctx.set_option(k, v)
        return ctx.response_batch(a, b, c)

    try:
        ref = run(sig, epl, deps, max_fm=0)                      # general Hill path (any CV, any Hill parameters)
        n50 = int(np.sum(ref[4] == 49))
        assert 0.05 * N * N < n50 < 0.5 * N * N                  # the 50-sub-step pass really ran
        # passb_recur: the planar isotropic sub-step loop with scalar recurrences on its carried chain (hill_fast.cuh)
        for opts in (dict(max_fm=2), dict(max_fm=3), dict(max_fm=4), dict(max_fm=4), dict(max_fm=4, passa_direct=1),
                     dict(max_fm=4, passb_recur=1), dict(max_fm=4, passb_recur=PASSB_RECUR_DEFAULT)):
            out = run(sig, epl, deps, **opts)
            assert np.array_equal(out[4], ref[4]), opts          # same branch per element
            for a, b in zip(out[:4], ref[:4]):
                assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b)), opts
        idx = np.random.default_rng(3).choice(N * N, 20000, replace=False)
        oref = orc.response_batch(mats, None, sig[idx], epl[idx], deps[idx], CV[None])
        fy, so, dp, ct, ns = (o[idx] for o in out)
        assert np.array_equal(ns, oref[4])
        assert np.max(np.abs(fy - oref[0])) < TOL * bench.SY
        assert relmax(so, oref[1]) < TOL and relmax(ct, oref[3]) < TOL
        assert np.max(np.abs(dp - oref[2])) < TOL * np.max(np.abs(oref[2]))
        # non-planar states: 30 % of a 200k-element batch get 23 / 13 shear components
        m = 200000
        s3, e3, d3 = sig[:m].copy(), epl[:m].copy(), deps[:m].copy()
        rng = np.random.default_rng(1)
        odd = rng.random(m) < 0.3
        s3[odd, 3] = 20. * rng.standard_normal(odd.sum())
        s3[odd, 4] = 10. * rng.standard_normal(odd.sum())
        d3[odd, 3] = 1e-4 * rng.standard_normal(odd.sum())
        out = run(s3, e3, d3, passb_recur=PASSB_RECUR_DEFAULT)
        sel = np.concatenate([np.nonzero(odd)[0][:3000], np.nonzero(~odd)[0][:3000]])
        oref = orc.response_batch(mats, None, s3[sel], e3[sel], d3[sel], CV[None])
        assert np.array_equal(out[4][sel], oref[4])
        assert int(np.sum(out[4][sel][:3000] == 49)) > 100       # deferred non-planar elements exist
        assert relmax(out[1][sel], oref[1]) < TOL and relmax(out[3][sel], oref[3]) < TOL
        assert np.max(np.abs(out[2][sel] - oref[2])) < TOL * np.max(np.abs(oref[2]))
    finally:
        for k, v in dict(compact=1, max_fm=4, passa_direct=0, passb_recur=PASSB_RECUR_DEFAULT).items():
            ctx.set_option(k, v)


def FE_elastic(E, nu):
    """plf_material of an isotropic elastic material (3-d stiffness), as bench.j2_material builds its J2 one."""
    from pylabfea_b200 import _lib
    m = _lib.PlfMaterial()
    m.kind = _lib.MAT_ELASTIC
    m.E, m.nu = E, nu
    hh = E / ((1. + nu) * (1. - 2. * nu))
    C11, C12, C44 = (1. - nu) * hh, nu * hh, (0.5 - nu) * hh
    CV = np.zeros((6, 6))
    CV[:3, :3] = C12
    CV[0, 0] = CV[1, 1] = CV[2, 2] = C11
    CV[3, 3] = CV[4, 4] = CV[5, 5] = C44
    m.CV[:] = list(CV.ravel())
    m.C44 = C44
    m.hill[:] = [1.] * 6
    return m


def test_fast_modes_with_an_elastic_second_material(ctx):
    """An FE model's material set -- J2 matrix + elastic inclusion (examples/inclusion.py) -- runs in the J2 fast modes too: the
    elastic elements take pass A's elastic branch, the plastic ones the planar isotropic loops.  Same branch per element and
    results within 1e-10 of the general form (max_fm = 0), and equal to the oracle on a sample."""
    import bench
    from oracle import oracle as orc
    N = 512
    mat, CV = bench.j2_material()
    soft = FE_elastic(3.e3, 0.3)
    ctx.set_materials([mat, soft], planestress=False)
    sig, epl, deps = bench.synth_states(N, 4321)
    n = N * N
    rng = np.random.default_rng(5)
    mid = (rng.random(n) < 0.15).astype(np.int32)
    epl[mid == 1] = 0.
    ctx.last_kernels()
    try:
        ctx.set_option('max_fm', 0)
        ref = ctx.response_batch(sig, epl, deps, mat_id=mid)
        assert 'response_passA_kernel<0, 0, 0>' in ctx.last_kernels()
        ctx.set_option('max_fm', 4)
        out = ctx.response_batch(sig, epl, deps, mat_id=mid)
        names = ctx.last_kernels()
        assert 'response_passA_kernel<0, 4, 0>' in names and 'response_passB_kernel<3, 0, 8>' in names, names
    finally:
        ctx.set_option('max_fm', 4)
    assert np.array_equal(out[4], ref[4])
    assert int(np.sum(out[4] == 49)) > 0.05 * n
    for a, b in zip(out[:4], ref[:4]):
        assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))
    # elastic elements: stress from CV, no plastic strain, tangent = CV (Model.solve never calls response for them, model.py:1341)
    el = mid == 1
    CV1 = ctx.element_CV(1)
    assert np.all(out[4][el] == 0) and np.all(out[2][el] == 0.) and np.array_equal(out[3][el], np.broadcast_to(CV1, (int(el.sum()), 6, 6)))
    assert relmax(out[1][el], sig[el] + deps[el] @ CV1.T) < 1e-13
    idx = rng.choice(np.nonzero(~el)[0], 20000, replace=False)
    mats = [orc.Mat(kind=orc.HILL, E=bench.E_MOD, nu=bench.NU, sy=bench.SY, khard=bench.KHARD), orc.Mat(kind=orc.ELASTIC, E=3.e3, nu=0.3)]
    CVs = np.array([CV, CV1])
    oref = orc.response_batch(mats, mid[idx], sig[idx], epl[idx], deps[idx], CVs)
    assert np.array_equal(out[4][idx], oref[4])
    assert relmax(out[1][idx], oref[1]) < TOL and relmax(out[3][idx], oref[3]) < TOL
    assert np.max(np.abs(out[2][idx] - oref[2])) < TOL * np.max(np.abs(oref[2]))


@pytest.mark.parametrize('shape', [(40, 24), (3, 70), (1, 5)])
def test_step_host_pipelined_equals_unpipelined(ctx, shape):
    """plf_step_host (chunked upload / update + column-range assembly / download on three streams) returns exactly
    what the single-stream sequence batch_upload + batch_run + setupK_from_batch + batch_download + get_csr returns."""
    import bench
    NX, NY = shape
    mat, CV = bench.j2_material()
    ctx.set_materials([mat], planestress=False)
    ctx.mesh_structured(NX, NY, np.full(NX, 4. / NX), 4. / NY, 1., np.zeros(NX * NY, np.int32))
    n = NX * NY
    s0, e0, d0 = bench.synth_states(max(NX, NY), 77)
    sig, epl, deps = s0[:n].copy(), e0[:n].copy(), d0[:n].copy()
    b = ctx.batch_create(n)
    ctx.batch_upload(b, sig, epl, deps)
    ctx.batch_run(b)
    ctx.setupK_from_batch(b)
    ref = [np.empty(n), np.empty((n, 6)), np.empty((n, 6)), np.empty((n, 6, 6)), np.empty(n, np.int32)]
    ctx.batch_download(b, fy=ref[0], sig_out=ref[1], depl=ref[2], ctan=ref[3], nsteps=ref[4])
    kref = ctx.get_csr_values(np.empty(ctx.nnz))
    out = [np.full(n, np.nan), np.full((n, 6), np.nan), np.full((n, 6), np.nan), np.full((n, 6, 6), np.nan),
           np.full(n, -7, np.int32)]
    kval = np.full(ctx.nnz, np.nan)
    for _ in range(2):   # twice: the second call re-uses streams / events and meets a used batch
        ctx.step_host(b, sig, epl, deps, fy=out[0], sig_out=out[1], depl=out[2], ctan=out[3], nsteps=out[4], kval=kval)
        for a, r in zip(out, ref):
            assert np.array_equal(a, r)
        assert np.array_equal(kval, kref)
    ctx.step_host(b, sig, epl, deps)   # all outputs optional
    ctx.batch_destroy(b)


def _unpack_tangents(n, ct21, words, n_packed, CVs, mat_id):
    """Host side of plf_step_host_packed: full (n,6,6) tangents from the elastic bits and the packed upper triangles."""
    bits = np.unpackbits(words.view(np.uint8), bitorder='little')[:n].astype(bool)
    full = np.empty((n, 6, 6))
    full[bits] = CVs[mat_id[bits]]
    iu = np.triu_indices(6)
    tri = ct21[:21 * n_packed].reshape(n_packed, 21)
    assert n_packed == int(np.sum(~bits))
    t = np.zeros((n_packed, 6, 6))
    t[:, iu[0], iu[1]] = tri
    t[:, iu[1], iu[0]] = tri
    full[~bits] = t
    return full, bits


@pytest.mark.parametrize('shape,mixed', [((64, 32), False), ((16, 64), True), ((8, 4), False)])
def test_step_host_packed_equals_step_host(ctx, shape, mixed):
    """plf_step_host_packed (upper triangles of the non-elastic tangents in element order + one elastic bit per element,
    K optionally left on the device) carries exactly the information of plf_step_host: every result array bit-identical
    after unpacking, also on a mixed-material mesh (J2 + elastic inclusion columns + a second plastic material)."""
    import bench
    from helpers import elastic_material
    NX, NY = shape
    mat, CV = bench.j2_material()
    n = NX * NY
    mats, elem_mat = [mat], np.zeros(n, np.int32)
    if mixed:
        soft = elastic_material(3.e3, 0.3, CV * (3.e3 / bench.E_MOD))
        hard, _ = bench.j2_material()
        hard.sy, hard.khard = 120., 2000.
        mats = [mat, soft, hard]
        elem_mat = ((np.arange(n) // NY) % 3).astype(np.int32)
    ctx.set_materials(mats, planestress=False)
    ctx.mesh_structured(NX, NY, np.full(NX, 4. / NX), 4. / NY, 1., elem_mat)
    CVs = np.array([ctx.element_CV(i) for i in range(len(mats))])
    s0, e0, d0 = bench.synth_states(max(NX, NY), 78)
    sig, epl, deps = s0[:n].copy(), e0[:n].copy(), d0[:n].copy()
    b = ctx.batch_create(n)
    ref = [np.empty(n), np.empty((n, 6)), np.empty((n, 6)), np.empty((n, 6, 6)), np.empty(n, np.int32)]
    kref = np.empty(ctx.nnz)
    ctx.step_host(b, sig, epl, deps, fy=ref[0], sig_out=ref[1], depl=ref[2], ctan=ref[3], nsteps=ref[4], kval=kref)
    out = [np.full(n, np.nan), np.full((n, 6), np.nan), np.full((n, 6), np.nan), np.full(n, -7, np.int32)]
    ct21 = np.full(21 * n, np.nan)
    words = np.zeros((n + 31) // 32, np.uint32)
    kval = np.full(ctx.nnz, np.nan)
    for with_k in (True, False):
        npk = ctx.step_host_packed(b, sig, epl, deps, ct21, words, fy=out[0], sig_out=out[1], depl=out[2], nsteps=out[3],
                                   kval=kval if with_k else None)
        full, bits = _unpack_tangents(n, ct21, words, npk, CVs, elem_mat)
        assert np.array_equal(full, ref[3])
        for a, r in zip(out, (ref[0], ref[1], ref[2], ref[4])):
            assert np.array_equal(a, r)
        assert np.array_equal(kval, kref)
        assert np.all(ref[2][bits] == 0.)            # elastic branch: no plastic strain increment
        assert 0 < npk < n
    assert np.array_equal(ctx.get_csr_values(np.empty(ctx.nnz)), kref)   # K left on the device is the same K
    ctx.batch_destroy(b)


def test_response_hill_isotropic_planar_mode_vs_oracle(ctx):
    """BASELINE config 3 (512 x 512, Hill-6p with isotropic hardening on isotropic elasticity, synthetic planar field):
    the Hill fast mode (hilliso_passA + planar 50-sub-step loop) against the general Hill kernels on every element and
    against the CPU oracle on a seeded sample."""
    import bench
    from oracle import oracle as orc
    N = 512
    hill = [0.7, 1., 1.4, 1., 1.2, 0.8]
    mat, CV = bench.j2_material()
    mat.hill[:] = hill
    mat.khard = 100.
    ctx.set_materials([mat], planestress=False)
    sig, epl, deps = bench.synth_states(N, 4321, khard=100.)
    try:
        ctx.set_option('max_fm', 0)
        ref = ctx.response_batch(sig, epl, deps)
        ctx.set_option('max_fm', 4)
        out = ctx.response_batch(sig, epl, deps)
        n50 = int(np.sum(ref[4] == 49))
        assert n50 > 0.03 * N * N
        assert np.array_equal(out[4], ref[4])
        for a, b in zip(out[:4], ref[:4]):
            assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))
        idx = np.random.default_rng(5).choice(N * N, 20000, replace=False)
        mats = [orc.Mat(kind=orc.HILL, E=bench.E_MOD, nu=bench.NU, sy=bench.SY, khard=100., hill=hill)]
        oref = orc.response_batch(mats, None, sig[idx], epl[idx], deps[idx], CV[None])
        assert np.array_equal(out[4][idx], oref[4])
        assert np.max(np.abs(out[0][idx] - oref[0])) < TOL * bench.SY
        assert relmax(out[1][idx], oref[1]) < TOL and relmax(out[3][idx], oref[3]) < TOL
        assert np.max(np.abs(out[2][idx] - oref[2])) < TOL * np.max(np.abs(oref[2]))
    finally:
        ctx.set_option('max_fm', 4)


@pytest.mark.parametrize('world', [2, 3])
def test_strip_sharded_K_rows_equal_global_K(ctx, golden, world):
    """a rank that owns a strip of element columns of ONE mesh and recomputes its left neighbour's last column
    (halo) assembles the rows of its node columns bit-identically to the single-GPU assembly of the whole mesh -- no
    communication, same ascending-element summation order.  Ranks are played one after the other on one GPU."""
    from pylabfea_b200.shard import strip_shard, local_element_ids
    g = golden('response_j2')
    NX, NY = 26, 9
    dx = np.array([0.1, 0.25, 0.15, 0.3])[np.arange(NX) % 4]   # non-uniform columns: four geometry classes
    ctx.set_materials([plf_material_from_golden(g)], planestress=False)
    rng = np.random.default_rng(17)
    T = g['pe_ctan'][rng.integers(0, len(g['pe_ctan']), NX * NY)]
    ctx.mesh_structured(NX, NY, dx, 0.2, 1., np.zeros(NX * NY, np.int32))
    ctx.set_tangent(T)
    ctx.setupK()
    rp_g, col_g, val_g = ctx.get_csr()
    seen = 0
    for r in range(world):
        s = strip_shard(NX, r, world)
        first, nxl = s['first_col_global'], s['nx_local']
        ids = np.array(local_element_ids(s, NY))
        ctx.mesh_structured(nxl, NY, dx[first:first + nxl], 0.2, 1., np.zeros(nxl * NY, np.int32))
        ctx.set_tangent(T[ids])
        ctx.setupK()
        rp, col, val = ctx.get_csr()
        for jg, jl in zip(s['owned_node_cols_global'], s['owned_node_cols_local']):
            for k in range(NY + 1):
                for c in range(2):
                    rg, rl = 2 * (jg * (NY + 1) + k) + c, 2 * (jl * (NY + 1) + k) + c
                    vg, vl = val_g[rp_g[rg]:rp_g[rg + 1]], val[rp[rl]:rp[rl + 1]]
                    if jg == 0 or jg == NX or (jl > 0 and jl < nxl):
                        assert np.array_equal(vg, vl), (r, jg, k, c)
                    else:
                        raise AssertionError('owned node column on a local boundary that is not a global boundary')
                    # column indices agree up to the shift of the node numbering
                    assert np.array_equal(col_g[rp_g[rg]:rp_g[rg + 1]] - 2 * first * (NY + 1), col[rp[rl]:rp[rl + 1]])
                    seen += 1
    assert seen == 2 * (NX + 1) * (NY + 1)


@pytest.mark.parametrize('tag', ['pe', 'ps'])
@pytest.mark.parametrize('name', ['j2s3', 'hill3'])
def test_sdim3_response_vs_reference_golden(ctx, golden, name, tag):
    """sdim = 3 principal-stress formulation (HillPrincModel) against reference outputs on axis-aligned states."""
    from pylabfea_b200 import _lib
    g = golden('response_' + name)
    m = plf_material_from_golden(g)
    m.kind = _lib.MAT_HILL3
    ctx.set_materials([m], planestress=(tag == 'ps'))
    assert np.array_equal(ctx.element_CV(0), g[tag + '_CV'])
    sig, epl, deps = g[tag + '_sig'], g[tag + '_epl'], g[tag + '_deps']
    assert relmax(ctx.calc_seq(0, sig), g[tag + '_seq']) < 1e-12
    assert np.max(np.abs(ctx.calc_yf(0, sig, epl) - g[tag + '_yf'])) < 1e-10
    fg = ctx.calc_fgrad(0, sig, epl)
    assert relmax(fg[:, 0:3], g[tag + '_fgrad3']) < 1e-11 and np.all(fg[:, 3:6] == 0.)
    assert relmax(ctx.epl_dot(0, sig, epl, deps), g[tag + '_epl_dot']) < 1e-10
    assert relmax(ctx.C_tan(0, sig, epl), g[tag + '_C_tan']) < 1e-11
    check_response(ctx.response_batch(sig, epl, deps), g, tag, float(g['sy']))


def test_sdim3_response_on_rotated_planar_states_vs_reference_golden(ctx, golden):
    """sdim = 3 material on stresses with in-plane shear (principal axes rotated about z by any angle): the reference's
    eigenvalue order (basic.py:155-170) is reproduced for every planar state -- all a 2-d model can produce; 240 reference
    tuples, 134 of them 50-sub-step."""
    from pylabfea_b200 import _lib
    g = golden('response_hill3_shear')
    m = plf_material_from_golden(g)
    m.kind = _lib.MAT_HILL3
    ctx.set_materials([m], planestress=False)
    sig, epl, deps = g['pe_sig'], g['pe_epl'], g['pe_deps']
    assert relmax(ctx.calc_seq(0, sig), g['pe_seq']) < 1e-12
    check_response(ctx.response_batch(sig, epl, deps), g, 'pe', float(g['sy']))


def test_sdim3_j2_equals_sdim6_j2_on_axis_aligned_states(ctx, golden):
    """For J2 and stresses without shear the sdim = 3 formulation is mathematically the sdim = 6 one; 20 000 random
    axis-aligned tuples through both kernel families must agree to rounding (also a guard for the code generation of
    the large eigen-decomposition kernels)."""
    from pylabfea_b200 import _lib
    g = golden('response_j2s3')
    rng = np.random.default_rng(99)
    n = 20000
    src = rng.integers(0, len(g['pe_sig']), n)
    sig = g['pe_sig'][src] * rng.uniform(0.8, 1.05, (n, 1))
    epl = g['pe_epl'][src]
    deps = g['pe_deps'][rng.integers(0, len(g['pe_sig']), n)] * rng.uniform(0.2, 1.5, (n, 1))
    outs = []
    for kind in (_lib.MAT_HILL3, _lib.MAT_HILL):
        m = plf_material_from_golden(g)
        m.kind = kind
        ctx.set_materials([m], planestress=False)
        outs.append(ctx.response_batch(sig, epl, deps))
    assert np.array_equal(outs[0][4], outs[1][4])
    assert int(np.sum(outs[0][4] == 49)) > 2000
    for a, b in zip(outs[0][:4], outs[1][:4]):
        assert np.max(np.abs(a - b)) <= 1e-9 * np.max(np.abs(b))


@pytest.mark.parametrize('hill', [None, [0.7, 1., 1.4, 1., 1.2, 0.8]])
def test_response_plane_stress_block_mode_vs_general_and_oracle(ctx, hill):
    """Plane-stress CV (zero third row / column: not isotropic in the 3-d sense) with J2 and Hill-6p, 512 x 512 synthetic
    planar field: the block-CV fast mode (hillblk_passA + planar loop, FM 7) against the general kernels on every element
    and against the CPU oracle on a seeded sample."""
    import bench
    from oracle import oracle as orc
    N = 512
    mat, CV3 = bench.j2_material()
    if hill is not None:
        mat.hill[:] = hill
    mat.khard = 200.
    ctx.set_materials([mat], planestress=True)
    CV = ctx.element_CV(0)
    assert CV[2, 2] == 0. and CV[0, 0] > 0.
    sig, epl, deps = bench.synth_states(N, 777, khard=200.)
    sig[:, 2] = 0.                                   # plane stress
    if hill is not None:
        # synth_states sets the load level with the J2 equivalent stress; keep the same level in the Hill measure
        def seq_of(h):
            d01, d12, d20 = sig[:, 0] - sig[:, 1], sig[:, 1] - sig[:, 2], sig[:, 2] - sig[:, 0]
            return np.sqrt(0.5 * (h[0] * d01 ** 2 + h[1] * d12 ** 2 + h[2] * d20 ** 2) + 3. * h[5] * sig[:, 5] ** 2)
        sig *= (seq_of([1.] * 6) / np.maximum(seq_of(hill), 1e-300))[:, None]
    try:
        ctx.set_option('max_fm', 0)
        ref = ctx.response_batch(sig, epl, deps)
        ctx.set_option('max_fm', 4)
        out = ctx.response_batch(sig, epl, deps)
        assert int(np.sum(ref[4] == 49)) > 0.02 * N * N
        assert np.array_equal(out[4], ref[4])
        for a, b in zip(out[:4], ref[:4]):
            assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))
        idx = np.random.default_rng(8).choice(N * N, 10000, replace=False)
        mats = [orc.Mat(kind=orc.HILL, E=bench.E_MOD, nu=bench.NU, sy=bench.SY, khard=200., hill=hill)]
        oref = orc.response_batch(mats, None, sig[idx], epl[idx], deps[idx], CV[None])
        assert np.array_equal(out[4][idx], oref[4])
        assert relmax(out[1][idx], oref[1]) < TOL and relmax(out[3][idx], oref[3]) < TOL
        assert np.max(np.abs(out[2][idx] - oref[2])) < TOL * np.max(np.abs(oref[2]))
    finally:
        ctx.set_option('max_fm', 4)
