#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ by RUNNING THE REFERENCE (test infrastructure only).

Runs in the build container only (needs /root/reference); the GPU box never executes this file, it only
reads the committed .npz fixtures.  Every array written here is an output of unmodified reference code
(pyLabFEA v4.4.2 @ c488572) on seeded inputs:

  response_<mat>.npz  Material.response (material.py:207-346) on (sig, epl, deps, CV) tuples that populate the
                      elastic / split / 1-step / 50-sub-step / radial-correction branches; plus calc_seq
                      (576-676), calc_yf (348-412), calc_fgrad (704-858), epl_dot (1009-1055), C_tan (1057-1086).
  kmat_<case>.npz     Element.calc_Bmat / calc_Kel (model.py:365-370, 439-512) and Model.setupK (954-977):
                      B matrices, Kel and the dense K (stored as COO of its non-zeros) for elastic and plastic
                      element tangents.
  solve_<case>.npz    Model.solve (model.py:979-1450) end-to-end: u, f, element sig/eps/epl, sgl/egl/epgl, niter.
  svc_<name>.npz      a trained sklearn SVC (support vectors, dual coefs, intercept, gamma, scale_seq) plus
                      calc_yf / calc_fgrad / ML_full_yf / response outputs of the reference on it.
  barlat.npz          Material.calc_seq for Yld2004-18p (material.py:678-702).
  barlat_response.npz Material.response / epl_dot / C_tan of the reference for Yld2004-18p materials.  The reference
                      has no Barlat gradient (calc_fgrad raises, material.py:822-823), so a test-only subclass
                      supplies it as 4th-order central differences of the reference's OWN calc_seq; everything else
                      (return mapping, sub-stepping, tangent) is unmodified reference code (Barlat plan).

usage: python tests/golden/make_golden.py [response] [kmat] [solve] [svc] [barlat] [barlat_response] [all]
"""
import os
import sys
import time
import warnings
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from ref_shim import import_reference  # noqa: E402

FE = import_reference()
warnings.filterwarnings('ignore')


# ----------------------------------------------------------------------------------------------
# materials
# ----------------------------------------------------------------------------------------------
def make_material(name):
    m = FE.Material(name=name)
    if name == 'j2':        # BASELINE config 2 / target
        m.elasticity(E=200.e3, nu=0.3)
        m.plasticity(sy=150., khard=500., sdim=6)
    elif name == 'j2pp':    # perfectly plastic J2
        m.elasticity(E=200.e3, nu=0.3)
        m.plasticity(sy=150., khard=0., sdim=6)
    elif name == 'hill':    # BASELINE config 3, reference tests/test_basic.py:130-133 uses sy=100
        m.elasticity(E=200.e3, nu=0.3)
        m.plasticity(sy=150., hill=[0.7, 1., 1.4, 1., 1.2, 0.8], khard=100., sdim=6)
    elif name == 'hillrv':  # examples/train_hill.py:26-33 reference material of config 4
        m.elasticity(E=200.e3, nu=0.3)
        m.plasticity(sy=50., rv=[1.2, 1.0, 0.8, 1.0, 1.0, 1.0], sdim=6)
    elif name == 'drucker':  # tension-compression asymmetry (never used in notebooks, cheap to pin)
        m.elasticity(E=100.e3, nu=0.27)
        m.plasticity(sy=120., hill=[1.2, 0.9, 1.1, 1.1, 0.9, 1.0], drucker=0.1, khard=300., sdim=6)
    elif name == 'cubic':   # C11/C12/C44 elasticity + J2 hardening
        m.elasticity(C11=170.e3, C12=124.e3, C44=75.e3)
        m.plasticity(sy=200., khard=1000., sdim=6)
    else:
        raise ValueError(name)
    return m


def mat_params(m):
    d = dict(E=m.E, nu=m.nu, C11=m.C11, C12=m.C12, C44=m.C44, CV=np.array(m.CV))
    if m.sy is not None:
        d.update(sy=m.sy, khard=m.khard, hill=np.array(m.hill, dtype=float),
                 drucker=float(m.drucker), sdim=m.sdim)
    return d


def cv_plane_stress(m):
    """Element.__init__ plane-stress CV (model.py:274-283)."""
    hh = m.E / (1 - m.nu * m.nu)
    C12 = m.nu * hh
    C11 = hh
    return np.array([[C11, C12, 0., 0., 0., 0.],
                     [C12, C11, 0., 0., 0., 0.],
                     [0., 0., 0., 0., 0., 0.],
                     [0., 0., 0., 0., 0., 0.],
                     [0., 0., 0., 0., 0., 0.],
                     [0., 0., 0., 0., 0., m.C44]])


def rand_dirs(rng, n, planar):
    d = rng.standard_normal((n, 6))
    if planar:
        d[:, 2:5] = 0.
    return d


def make_tuples(m, rng, n, planar):
    """Random (sig, epl, deps) states: SURVEY.md §8d recipe, extended with large increments so the
    radial-correction branch (material.py:310-342) fires."""
    sig = np.zeros((n, 6))
    epl = np.zeros((n, 6))
    deps = np.zeros((n, 6))
    dirs = rand_dirs(rng, n, planar)
    for i in range(n):
        # plastic strain history: none / uniaxial-like / random deviatoric
        c = i % 3
        if c == 1:
            epl[i] = rng.uniform(0., 0.02) * np.array([1., -.5, -.5, 0., 0., 0.])
        elif c == 2:
            e = rng.standard_normal(6) * 0.005
            e[0:3] -= e[0:3].mean()
            if planar:
                e[3:5] = 0.
            epl[i] = e
        sflow = m.get_sflow(epl[i])
        s = dirs[i]
        seq = m.calc_seq(s)
        # radius: mostly inside, many close to / on the yield surface
        u = rng.uniform()
        if u < 0.35:
            r = rng.uniform(0., 0.9)
        elif u < 0.7:
            r = rng.uniform(0.9, 1.0)
        else:
            r = rng.uniform(0.996, 1.004)
        sig[i] = s / seq * r * sflow
        de = rng.standard_normal(6)
        if planar:
            de[3:5] = 0.
        if i % 7 == 0:  # mostly radial loading direction (elastic strain direction of sig)
            de = np.linalg.lstsq(m.CV, sig[i] + 1e-3 * de, rcond=None)[0]
            if planar:
                de[3:5] = 0.
        mag = [5.e-4, 5.e-4, 2.e-3, 1.e-2, 1.e-4][i % 5]
        deps[i] = de / FE.eps_eq(de) * mag
    return sig, epl, deps


def gen_response():
    for name in ['j2', 'j2pp', 'hill', 'hillrv', 'drucker', 'cubic']:
        m = make_material(name)
        out = dict(mat_params(m))
        t0 = time.time()
        for planar, tag in [(False, 'pe'), (True, 'ps')]:
            rng = np.random.default_rng(42 + planar)
            n = 240
            CV = cv_plane_stress(m) if planar else np.array(m.CV)
            sig, epl, deps = make_tuples(m, rng, n, planar)
            fy = np.zeros(n)
            so = np.zeros((n, 6))
            dp = np.zeros((n, 6))
            ct = np.zeros((n, 6, 6))
            ns = np.zeros(n, dtype=np.int32)
            seq = np.zeros(n)
            yf = np.zeros(n)
            fg = np.zeros((n, 6))
            ed = np.zeros((n, 6))
            ctan = np.zeros((n, 6, 6))
            for i in range(n):
                m.msg['nsteps'] = -1
                r = m.response(sig[i], epl[i], deps[i], CV)
                fy[i], so[i], dp[i], ct[i] = r
                ns[i] = m.msg['nsteps']
                seq[i] = m.calc_seq(sig[i])
                yf[i] = m.calc_yf(sig[i], epl=epl[i])
                fg[i] = m.calc_fgrad(sig[i], epl=epl[i])
                ed[i] = m.epl_dot(sig[i], epl[i], CV, deps[i])
                ctan[i] = m.C_tan(sig[i], CV, epl=epl[i])
            out.update({f'{tag}_CV': CV, f'{tag}_sig': sig, f'{tag}_epl': epl, f'{tag}_deps': deps,
                        f'{tag}_fy': fy, f'{tag}_sig_out': so, f'{tag}_depl': dp, f'{tag}_ctan': ct,
                        f'{tag}_nsteps': ns, f'{tag}_seq': seq, f'{tag}_yf': yf, f'{tag}_fgrad': fg,
                        f'{tag}_epl_dot': ed, f'{tag}_C_tan': ctan})
            el = np.sum((ns == 0) & (np.abs(ct - CV).max(axis=(1, 2)) == 0))
            print(f'  {name}/{tag}: elastic {el}, 50-step {np.sum(ns == 49)}, other {n - el - np.sum(ns == 49)}')
        np.savez_compressed(os.path.join(HERE, f'response_{name}.npz'), **out)
        print(f'response_{name}.npz  {time.time() - t0:.1f}s')


# ----------------------------------------------------------------------------------------------
# Kel / setupK
# ----------------------------------------------------------------------------------------------
def dense_to_coo(K):
    r, c = np.nonzero(K)
    return r.astype(np.int32), c.astype(np.int32), K[r, c]


def model_arrays(fe):
    Nel = fe.Nel
    d = dict(NX=fe.NnodeX - 1, NY=fe.NnodeY - 1, Ndof=fe.Ndof, planestress=int(fe.planestress),
             lenx=fe.lenx, leny=fe.leny, thick=fe.thick,
             nodes=np.array([el.nodes for el in fe.element], dtype=np.int32),
             lx=np.array([el.Lelx for el in fe.element]), ly=np.array([el.Lely for el in fe.element]),
             mat_id=np.array([fe.mat.index(el.Mat) for el in fe.element], dtype=np.int32),
             CV=np.array([el.CV for el in fe.element]),
             npos=np.array(fe.npos))
    d['mat_E'] = np.array([m.E for m in fe.mat])
    d['mat_nu'] = np.array([m.nu for m in fe.mat])
    d['mat_CV'] = np.array([m.CV for m in fe.mat])
    return d


def kmat_case(name, fe, tangents=None):
    """tangents: optional (Nel,6,6) replaces el.elstiff before calc_Kel (what Model.solve does at 1351-1354)."""
    out = model_arrays(fe)
    if tangents is not None:
        for el, t in zip(fe.element, tangents):
            el.elstiff = np.array(t)
            el.calc_Kel()
    out['elstiff'] = np.array([el.elstiff for el in fe.element])
    out['Bmat'] = np.array([el.Bmat for el in fe.element[:min(fe.Nel, 64)]])   # (n,4,6,8)
    out['Kel'] = np.array([el.Kel for el in fe.element])
    K = fe.setupK()
    r, c, v = dense_to_coo(K)
    out['K_row'], out['K_col'], out['K_val'] = r, c, v
    np.savez_compressed(os.path.join(HERE, f'kmat_{name}.npz'), **out)
    print(f'kmat_{name}.npz Nel={fe.Nel} nnz={len(v)}')


def cantilever(N=10):
    """BASELINE config 1."""
    mat = FE.Material()
    mat.elasticity(E=200.e3, nu=0.3)
    fe = FE.Model(dim=2, planestress=True)
    fe.geom([10.], LY=1.)
    fe.assign([mat])
    fe.bcleft(0., 'disp', 'x')
    fe.bcleft(0., 'disp', 'y')
    fe.bcbot(0., 'force', 'y')
    fe.bctop(0., 'force', 'y')
    fe.bcright(-10., 'force', 'y')
    fe.mesh(NX=N, NY=N)
    return fe


def laminate():
    """tests/test_basic.py:84-101 (fem_v), 5 sections, plane stress."""
    fe = FE.Model(dim=2, planestress=True)
    fe.geom([2, 1, 2, 1, 2], LY=4.)
    mat1 = FE.Material()
    mat1.elasticity(E=100.e3, nu=0.35)
    mat2 = FE.Material()
    mat2.elasticity(E=300.e3, nu=0.3)
    fe.assign([mat1, mat2, mat1, mat2, mat1])
    fe.bcleft(0.)
    fe.bcbot(0.)
    fe.bcright(0., 'force')
    fe.bctop(0.1 * fe.leny, 'disp')
    fe.mesh(NX=16, NY=4)
    return fe


def uneven():
    """two sections with different element widths (dx = 1.25 / 0.75), plane strain."""
    fe = FE.Model(dim=2, planestress=False)
    fe.geom([2.5, 1.5], LY=3.)
    mat1 = make_material('j2')
    mat2 = make_material('hill')
    fe.assign([mat1, mat2])
    fe.bcleft(0.)
    fe.bcbot(0.)
    fe.bcright(0., 'force')
    fe.bctop(0.01 * fe.leny, 'disp')
    fe.mesh(NX=4, NY=5)
    return fe


def inclusion(N=18, plastic=False, planestress=False, eps_tot=0.01):
    """tests/test_basic.py:42-76 (elastic) / examples/inclusion.py:31-62 (plastic matrix)."""
    n1 = int(N / 3)
    el = np.ones((N, N))
    el[n1:2 * n1, n1:2 * n1] = 2
    mat1 = FE.Material(num=1)
    mat1.elasticity(E=100.e3, nu=0.27)
    if plastic:
        mat1.plasticity(sy=150., khard=500., sdim=6)
    mat2 = FE.Material(num=2)
    mat2.elasticity(E=3.e3, nu=0.3)
    fe = FE.Model(dim=2, planestress=planestress)
    fe.geom(sect=2, LX=4., LY=4.)
    fe.assign([mat1, mat2])
    fe.bcbot(0.)
    fe.bcright(0., 'force')
    fe.bcleft(0., 'force')
    fe.bctop(eps_tot * fe.leny, 'disp')
    fe.mesh(elmts=el, NX=N, NY=N)
    hh = [no in fe.nobot for no in fe.noleft]
    noc = np.nonzero(hh)[0]
    fe.bcnode(noc, 0., 'disp', 'x')
    return fe


def uniax(matname, N=4, planestress=False, eps=0.01):
    mat = make_material(matname)
    fe = FE.Model(dim=2, planestress=planestress)
    fe.geom([4.], LY=4.)
    fe.assign([mat])
    fe.bcleft(0.)
    fe.bcbot(0.)
    fe.bcright(0., 'force')
    fe.bctop(eps * fe.leny, 'disp')
    fe.mesh(NX=N, NY=N)
    return fe


def plastic_tangents(fe, seed):
    """Element tangents as Material.response returns them (sampled from the response fixtures' recipe)."""
    rng = np.random.default_rng(seed)
    T = []
    for el in fe.element:
        m = el.Mat
        if m.sy is None:
            T.append(np.array(el.CV))
            continue
        planar = fe.planestress
        sig, epl, deps = make_tuples(m, rng, 1, planar)
        r = m.response(sig[0], epl[0], deps[0] * 4., el.CV)
        T.append(r[3])
    return np.array(T)


def gen_kmat():
    kmat_case('cantilever10', cantilever(10))
    kmat_case('laminate16x4', laminate())
    kmat_case('inclusion18', inclusion(18))
    fe = uneven()
    kmat_case('uneven4x5_plastic', fe, plastic_tangents(fe, 7))
    fe = uniax('hill', N=6, planestress=True)
    kmat_case('hill6_ps_plastic', fe, plastic_tangents(fe, 8))
    fe = inclusion(9, plastic=True)
    kmat_case('inclusion9_plastic', fe, plastic_tangents(fe, 9))
    fe = cantilever(1)
    kmat_case('single', fe)


# ----------------------------------------------------------------------------------------------
# end-to-end solves
# ----------------------------------------------------------------------------------------------
def solve_case(name, fe, **kw):
    t0 = time.time()
    fe.solve(**kw)
    fe.calc_global()
    out = model_arrays(fe)
    out.update(u=fe.u, f=fe.f,
               sig=np.array([el.sig for el in fe.element]),
               eps=np.array([el.eps for el in fe.element]),
               epl=np.array([el.epl for el in fe.element]),
               elstiff=np.array([el.elstiff for el in fe.element]),
               sgl=fe.sgl, egl=fe.egl, epgl=fe.epgl, niter=np.array(fe.niter), nsteps=fe.nsteps,
               co_nconv=np.array(fe.co_nconv),
               glob_sig=fe.glob['sig'], glob_eps=fe.glob['eps'], glob_epl=fe.glob['epl'],
               glob_bc=np.array([fe.glob[k] for k in ('ebc1', 'sbc1', 'ebc21', 'sbc21', 'ebc2', 'sbc2', 'ebc12', 'sbc12')]),
               wall_s=time.time() - t0)
    np.savez_compressed(os.path.join(HERE, f'solve_{name}.npz'), **out)
    print(f'solve_{name}.npz  {time.time() - t0:.1f}s  nsteps={fe.nsteps} niter={fe.niter}')


def gen_solve():
    solve_case('cantilever10', cantilever(10))
    solve_case('laminate16x4', laminate())
    solve_case('inclusion18_elastic', inclusion(18))
    solve_case('j2_4x4', uniax('j2', 4), min_step=20)
    solve_case('hill_4x4_ps', uniax('hill', 4, planestress=True), min_step=10)
    solve_case('j2_inclusion6', inclusion(6, plastic=True), min_step=8)


# ----------------------------------------------------------------------------------------------
# SVC
# ----------------------------------------------------------------------------------------------
def svc_arrays(m):
    svm = m.svm_yf
    return dict(sv=np.array(svm.support_vectors_), dual=np.array(svm.dual_coef_[0]),
                rho=float(svm.intercept_[0]), gamma=float(m.gam_yf), scale_seq=float(m.scale_seq),
                dev_only=int(bool(m.dev_only)), sy=float(m.sy), khard=float(m.khard),
                E=m.E, nu=m.nu, CV=np.array(m.CV), sdim=m.sdim, Ndof=m.Ndof,
                hill=np.array(m.hill, dtype=float), drucker=float(m.drucker))


def gen_svc():
    # (1) examples/train_hill.py:26-49 recipe (BASELINE config 4)
    t0 = time.time()
    mat_h = make_material('hillrv')
    m = FE.Material('ML-Hill-p1', num=2)
    m.train_SVC(C=2.0, gamma=1.0, mat_ref=mat_h, Nseq=25, Nlc=300, Fe=0.1, Ce=0.99, gridsearch=False)
    print('trained: nsv =', len(m.svm_yf.support_vectors_), f'{time.time() - t0:.1f}s')
    out = svc_arrays(m)
    rng = np.random.default_rng(5)
    n = 400
    s = rng.standard_normal((n, 6))
    seq = mat_h.calc_seq(s)
    sig = s / seq[:, None] * (rng.uniform(0.05, 1.6, n) * m.sy)[:, None]
    out['yf_sig'] = sig
    out['yf'] = m.calc_yf(sig)                      # batched decision function (material.py:398-405)
    fg = np.zeros((n, 6))
    for i in range(n):
        fg[i] = m.calc_fgrad(sig[i])
    out['fgrad'] = fg
    nf = 120
    fyf = np.zeros(nf)
    for i in range(nf):
        fyf[i] = m.ML_full_yf(sig[i])
    out['ML_full_yf'] = fyf
    # response tuples
    nr = 48
    CVs = {'pe': np.array(m.CV), 'ps': cv_plane_stress(m)}
    for tag, CV in CVs.items():
        planar = tag == 'ps'
        rs, re, rd = make_tuples(mat_h, np.random.default_rng(11 + planar), nr, planar)
        # SVC material has no work hardening features -> perfectly plastic; keep epl small
        fy = np.zeros(nr)
        so = np.zeros((nr, 6))
        dp = np.zeros((nr, 6))
        ct = np.zeros((nr, 6, 6))
        ns = np.zeros(nr, dtype=np.int32)
        kh = np.zeros(nr)
        t1 = time.time()
        for i in range(nr):
            m.khard = 0.
            r = m.response(rs[i], re[i], rd[i], CV)
            fy[i], so[i], dp[i], ct[i] = r
            ns[i] = m.msg['nsteps']
            kh[i] = m.khard
        print(f'  svc response {tag}: {time.time() - t1:.1f}s')
        out.update({f'{tag}_CV': CV, f'{tag}_sig': rs, f'{tag}_epl': re, f'{tag}_deps': rd, f'{tag}_fy': fy,
                    f'{tag}_sig_out': so, f'{tag}_depl': dp, f'{tag}_ctan': ct, f'{tag}_nsteps': ns})
    np.savez_compressed(os.path.join(HERE, 'svc_hill.npz'), **out)
    print(f'svc_hill.npz {time.time() - t0:.1f}s')

    # (2) tests/test_ml.py:170-201 (test_ml_shear): SVC in FE, 6x3 plane stress simple shear
    t0 = time.time()
    mat_h = FE.Material(name='Hill-shear')
    mat_h.elasticity(E=200.e3, nu=0.3)
    mat_h.plasticity(sy=150., hill=[1.4, 1., 0.7, 1.2, .8, 1.], sdim=6)
    m = FE.Material('Hill-ML')
    m.train_SVC(C=2, gamma=0.5, mat_ref=mat_h, Nseq=4, Nlc=300, Fe=0.7, Ce=0.95)
    m.dev_only = False
    out = svc_arrays(m)
    fem = FE.Model(dim=2, planestress=True)
    fem.geom([2], LY=2.)
    fem.assign([m])
    fem.bcbot(0., bctype='disp', bcdir='y')
    fem.bcbot(0., bctype='disp', bcdir='x')
    fem.bcleft(0., bctype='force')
    fem.bcright(0., bctype='force')
    fem.bctop(0.006 * fem.leny, bctype='disp', bcdir='x')
    fem.bctop(0., bctype='disp', bcdir='y')
    fem.mesh(NX=6, NY=3)
    fem.solve()
    fem.calc_global()
    out.update(u=fem.u, f=fem.f, sig=np.array([el.sig for el in fem.element]),
               eps=np.array([el.eps for el in fem.element]), epl=np.array([el.epl for el in fem.element]),
               sgl=fem.sgl, egl=fem.egl, epgl=fem.epgl, niter=np.array(fem.niter), nsteps=fem.nsteps,
               glob_sig=fem.glob['sig'])
    np.savez_compressed(os.path.join(HERE, 'svc_shear.npz'), **out)
    print(f'svc_shear.npz {time.time() - t0:.1f}s glob sig5={fem.glob["sig"][5]} (golden 77.53778881971623)')


def gen_barlat():
    # examples/train_goss_barlat.py:36-41
    bp = None
    src = open('/root/reference/examples/train_goss_barlat.py').read()
    ns = {}
    for line in src.splitlines():
        if line.strip().startswith('bp ='):
            i = src.index(line)
            j = src.index(']', i)
            exec(src[i:j + 1].strip(), {'np': np}, ns)
            bp = ns['bp']
            break
    m = FE.Material('Goss-Barlat')
    m.elasticity(E=151220., nu=0.3)
    m.plasticity(sy=46.76, barlat=bp, barlat_exp=8, sdim=6)
    rng = np.random.default_rng(3)
    sig = rng.standard_normal((300, 6)) * 40.
    sig[::5, 3:5] = 0.
    seq = m.calc_seq(sig)
    np.savez_compressed(os.path.join(HERE, 'barlat.npz'), bp=np.array(bp, dtype=float), barlat_exp=8.,
                        Bar_m1=m.Bar_m1, Bar_m2=m.Bar_m2, sig=sig, seq=seq, sy=46.76, E=151220., nu=0.3)
    print('barlat.npz', seq[:3])


def barlat_material(khard=0.):
    # examples/train_goss_barlat.py:36-41
    bp = None
    src = open('/root/reference/examples/train_goss_barlat.py').read()
    ns = {}
    for line in src.splitlines():
        if line.strip().startswith('bp ='):
            i = src.index(line)
            j = src.index(']', i)
            exec(src[i:j + 1].strip(), {'np': np}, ns)
            bp = ns['bp']
            break

    class BarlatFD(FE.Material):
        """Reference Material whose (missing) Barlat gradient is the numerical derivative of its own calc_seq."""

        def calc_fgrad(self, sig, epl=None, **kw):
            if not self.barlat:
                return super().calc_fgrad(sig, epl=epl, **kw)
            sig = np.asarray(sig, dtype=float)
            if sig.ndim == 2:
                return np.array([self.calc_fgrad(s) for s in sig])
            seq0 = self.calc_seq(sig)
            h = 1.e-3 * max(seq0, 1.e-3)
            g = np.zeros(6)
            for k in range(6):
                f = []
                for off in (1., -1., 2., -2.):
                    x = sig.copy()
                    x[k] += off * h
                    f.append(self.calc_seq(x))
                g[k] = (8. * (f[0] - f[1]) - (f[2] - f[3])) / (12. * h)
            return g

    m = BarlatFD('Goss-Barlat')
    m.elasticity(E=151220., nu=0.3)
    m.plasticity(sy=46.76, khard=khard, barlat=bp, barlat_exp=8, sdim=6)
    return m, bp


def gen_barlat_response():
    out = {}
    t0 = time.time()
    for khard, kt in [(0., 'pp'), (300., 'kh')]:
        m, bp = barlat_material(khard)
        out.update(bp=np.array(bp, dtype=float), barlat_exp=8., Bar_m1=m.Bar_m1, Bar_m2=m.Bar_m2, sy=46.76, E=151220.,
                   nu=0.3)
        out[kt + '_khard'] = khard
        for planar, tag in [(False, 'pe'), (True, 'ps')]:
            rng = np.random.default_rng(7 + planar + int(khard))
            n = 60
            CV = cv_plane_stress(m) if planar else np.array(m.CV)
            sig, epl, deps = make_tuples(m, rng, n, planar)
            fy = np.zeros(n); so = np.zeros((n, 6)); dp = np.zeros((n, 6)); ct = np.zeros((n, 6, 6))
            ns = np.zeros(n, dtype=np.int32); fg = np.zeros((n, 6)); ed = np.zeros((n, 6)); ctan = np.zeros((n, 6, 6))
            seq = np.zeros(n)
            for i in range(n):
                m.msg['nsteps'] = -1
                fy[i], so[i], dp[i], ct[i] = m.response(sig[i], epl[i], deps[i], CV)
                ns[i] = m.msg['nsteps']
                seq[i] = m.calc_seq(sig[i])
                fg[i] = m.calc_fgrad(sig[i], epl=epl[i])
                ed[i] = m.epl_dot(sig[i], epl[i], CV, deps[i])
                ctan[i] = m.C_tan(sig[i], CV, epl=epl[i])
            t = kt + '_' + tag
            out.update({f'{t}_CV': CV, f'{t}_sig': sig, f'{t}_epl': epl, f'{t}_deps': deps, f'{t}_fy': fy,
                        f'{t}_sig_out': so, f'{t}_depl': dp, f'{t}_ctan': ct, f'{t}_nsteps': ns, f'{t}_seq': seq,
                        f'{t}_fgrad': fg, f'{t}_epl_dot': ed, f'{t}_C_tan': ctan})
            el = np.sum((ns == 0) & (np.abs(ct - CV).max(axis=(1, 2)) == 0))
            print(f'  barlat {t}: elastic {el}, 50-step {np.sum(ns == 49)}, other {n - el - np.sum(ns == 49)}, {time.time() - t0:.0f}s')
    np.savez_compressed(os.path.join(HERE, 'barlat_response.npz'), **out)
    print('barlat_response.npz')


def gen_barlat_solve():
    """Model.solve of the reference on a 3x3 plane-strain mesh with the FD-gradient Barlat material (hardening 300):
    uniaxial-y displacement control, 6 increments."""
    mat, bp = barlat_material(300.)
    fe = FE.Model(dim=2, planestress=False)
    fe.geom([4.], LY=4.)
    fe.assign([mat])
    fe.bcleft(0.)
    fe.bcbot(0.)
    fe.bcright(0., 'force')
    fe.bctop(0.002 * fe.leny, 'disp')
    fe.mesh(NX=3, NY=3)
    solve_case('barlat_3x3', fe, min_step=6)
    d = dict(np.load(os.path.join(HERE, 'solve_barlat_3x3.npz')))
    d.update(bp=np.array(bp, dtype=float), Bar_m1=mat.Bar_m1, Bar_m2=mat.Bar_m2, barlat_exp=8., sy=46.76, khard=300.,
             E=151220., nu=0.3)
    np.savez_compressed(os.path.join(HERE, 'solve_barlat_3x3.npz'), **d)


def gen_mlparam():
    """export_MLparam wire format (material.py:2133-2273): CSV (8 numbers per line: header block, dual coefficients,
    support vectors) + JSON metadata, written by the reference for a freshly trained SVC; plus reference calc_yf values
    so that a loader can be checked end to end."""
    t0 = time.time()
    mat_h = FE.Material(name='Hill-shear')
    mat_h.elasticity(E=200.e3, nu=0.3)
    mat_h.plasticity(sy=150., hill=[1.4, 1., 0.7, 1.2, .8, 1.], sdim=6)
    m = FE.Material('Hill-ML')
    m.train_SVC(C=2, gamma=0.5, mat_ref=mat_h, Nseq=4, Nlc=300, Fe=0.7, Ce=0.95)
    cwd = os.getcwd()
    os.chdir(HERE)
    try:
        m.export_MLparam(sname='tests/golden/make_golden.py', source='Hill-shear reference material', file='mlparam_hill',
                         path='./', descr=['sy', 'hill'], param=[150., [1.4, 1., 0.7, 1.2, .8, 1.]])
    finally:
        os.chdir(cwd)
    rng = np.random.default_rng(9)
    s = rng.standard_normal((200, 6))
    sig = s / mat_h.calc_seq(s)[:, None] * (rng.uniform(0.05, 1.6, 200) * m.sy)[:, None]
    out = svc_arrays(m)
    out.update(yf_sig=sig, yf=m.calc_yf(sig), C11=m.C11, C12=m.C12, C44=m.C44)
    np.savez_compressed(os.path.join(HERE, 'mlparam_hill.npz'), **out)
    print(f'mlparam_hill-svm.csv / -svm_meta.json / mlparam_hill.npz  nsv={len(out["dual"])}  {time.time() - t0:.1f}s')


def gen_mlparam3():
    """export_MLparam of the sdim = 3 SVC of tests/test_ml.py:10-28 (Ndof = 2: cylindrical principal-stress features),
    plus reference calc_yf values on principal stresses so that the loader can be checked end to end."""
    t0 = time.time()
    mat_h = FE.Material(name='anisotropic Hill')
    mat_h.elasticity(E=200.e3, nu=0.3)
    mat_h.plasticity(sy=150., hill=[0.7, 1., 1.4], drucker=0., khard=0., sdim=3)
    m = FE.Material(name='ML flow rule')
    m.elasticity(E=200.e3, nu=0.3)
    m.plasticity(sy=150., sdim=3)
    x_train, y_train = m.create_sig_data(36, mat_ref=mat_h, extend=True)
    m.setup_yf_SVM_3D(x_train, y_train, C=10, gamma=4., fs=0.3)
    cwd = os.getcwd()
    os.chdir(HERE)
    try:
        m.export_MLparam(sname='tests/golden/make_golden.py', source='Hill-3p reference material', file='mlparam_svc3',
                         path='./', descr=['sy', 'hill'], param=[150., [0.7, 1., 1.4]])
    finally:
        os.chdir(cwd)
    rng = np.random.default_rng(19)
    sp = rng.standard_normal((200, 3))
    sp = sp / mat_h.calc_seq(sp)[:, None] * (rng.uniform(0.05, 1.6, 200) * m.sy)[:, None]
    out = svc_arrays(m)
    out.update(yf_sp=sp, yf=m.calc_yf(sp))
    np.savez_compressed(os.path.join(HERE, 'mlparam_svc3.npz'), **out)
    print(f'mlparam_svc3-svm.csv / -svm_meta.json / mlparam_svc3.npz  nsv={len(out["dual"])}  {time.time() - t0:.1f}s')


def gen_tresca():
    """Tresca equivalent stress / yield function of the reference (material.py:630-632) on random Voigt stresses."""
    m = FE.Material('Tresca')
    m.elasticity(E=200.e3, nu=0.3)
    m.plasticity(sy=150., khard=200., tresca=True, sdim=6)
    rng = np.random.default_rng(21)
    sig = rng.standard_normal((300, 6)) * 90.
    sig[::6, 3:6] = 0.
    epl = np.zeros((300, 6))
    epl[::2] = rng.uniform(0., 0.02, (150, 1)) * np.array([1., -.5, -.5, 0., 0., 0.])
    seq = m.calc_seq(sig)
    yf = np.array([m.calc_yf(sig[i], epl=epl[i]) for i in range(300)])
    raised = False
    try:
        m.calc_fgrad(sig[0])
    except ValueError:
        raised = True
    np.savez_compressed(os.path.join(HERE, 'tresca.npz'), sig=sig, epl=epl, seq=seq, yf=yf, sy=150., khard=200., E=200.e3, nu=0.3,
                        CV=np.array(m.CV), fgrad_raises=raised)
    print('tresca.npz', seq[:3], 'calc_fgrad raises:', raised)


def gen_sdim3():
    """sdim = 3 principal-stress formulation (material.py:663-673, 836-838, 1045-1048, 1078-1080): Material.response /
    calc_seq / calc_yf / epl_dot / C_tan of the reference for J2 and 3-parameter Hill materials on states whose
    principal axes are the coordinate axes (no shear: what the reference's own sdim = 3 tests produce), and the two
    reference test models that use sdim = 3 (tests/test_basic.py:107-125)."""
    t0 = time.time()
    for name, hill, khard in (('j2s3', None, 500.), ('hill3', [0.7, 1., 1.4], 100.)):
        m = FE.Material(name=name)
        m.elasticity(E=200.e3, nu=0.3)
        m.plasticity(sy=150., hill=hill, khard=khard, sdim=3)
        out = dict(E=200.e3, nu=0.3, sy=150., khard=khard, hill=np.array(m.hill, dtype=float), drucker=0., mat_CV=np.array(m.CV))
        for planar, tag in [(False, 'pe'), (True, 'ps')]:
            rng = np.random.default_rng(77 + planar)
            n = 160
            CV = cv_plane_stress(m) if planar else np.array(m.CV)
            sig = np.zeros((n, 6)); epl = np.zeros((n, 6)); deps = np.zeros((n, 6))
            for i in range(n):
                if i % 3 == 1:
                    epl[i] = rng.uniform(0., 0.02) * np.array([1., -.5, -.5, 0., 0., 0.])
                elif i % 3 == 2:
                    e = rng.standard_normal(3) * 0.005
                    epl[i, 0:3] = e - e.mean()
                d = np.zeros(6)
                d[0:3] = rng.standard_normal(3)
                if planar:
                    d[2] = 0.
                sflow = m.get_sflow(epl[i])
                u = rng.uniform()
                r = rng.uniform(0., 0.9) if u < 0.35 else (rng.uniform(0.9, 1.0) if u < 0.7 else rng.uniform(0.996, 1.004))
                sig[i] = d / m.calc_seq(d) * r * sflow
                de = np.zeros(6)
                de[0:3] = rng.standard_normal(3)
                if planar:
                    de[2] = 0.
                if i % 7 == 0:
                    de = np.linalg.lstsq(m.CV, sig[i] + 1e-3 * de, rcond=None)[0]
                    de[3:6] = 0.
                    if planar:
                        de[2] = 0.
                deps[i] = de / FE.eps_eq(de) * [5.e-4, 5.e-4, 2.e-3, 1.e-2, 1.e-4][i % 5]
            fy = np.zeros(n); so = np.zeros((n, 6)); dp = np.zeros((n, 6)); ct = np.zeros((n, 6, 6)); ns = np.zeros(n, dtype=np.int32)
            seq = np.zeros(n); yf = np.zeros(n); fg = np.zeros((n, 3)); ed = np.zeros((n, 6)); ctan = np.zeros((n, 6, 6))
            for i in range(n):
                m.msg['nsteps'] = -1
                fy[i], so[i], dp[i], ct[i] = m.response(sig[i], epl[i], deps[i], CV)
                ns[i] = m.msg['nsteps']
                seq[i] = m.calc_seq(sig[i])
                yf[i] = m.calc_yf(sig[i], epl=epl[i])
                fg[i] = m.calc_fgrad(FE.sig_princ(sig[i])[0], epl=epl[i, 0:3])
                ed[i] = m.epl_dot(sig[i], epl[i], CV, deps[i])
                ctan[i] = m.C_tan(sig[i], CV, epl=epl[i])
            out.update({f'{tag}_CV': CV, f'{tag}_sig': sig, f'{tag}_epl': epl, f'{tag}_deps': deps, f'{tag}_fy': fy,
                        f'{tag}_sig_out': so, f'{tag}_depl': dp, f'{tag}_ctan': ct, f'{tag}_nsteps': ns, f'{tag}_seq': seq,
                        f'{tag}_yf': yf, f'{tag}_fgrad3': fg, f'{tag}_epl_dot': ed, f'{tag}_C_tan': ctan})
            el = np.sum((ns == 0) & (np.abs(ct - CV).max(axis=(1, 2)) == 0))
            print(f'  {name}/{tag}: elastic {el}, 50-step {np.sum(ns == 49)}, other {n - el - np.sum(ns == 49)}')
        np.savez_compressed(os.path.join(HERE, f'response_{name}.npz'), **out)
    # tests/test_basic.py:107-121: two sections (elastic | J2 sdim = 3), 4 x 4 plane strain, 10 % strain
    mat1 = FE.Material()
    mat1.elasticity(E=100.e3, nu=0.35)
    mat2 = FE.Material()
    mat2.elasticity(E=300.e3, nu=0.3)
    mat2.plasticity(sy=150., khard=500., sdim=3)
    fem2 = FE.Model(dim=2, planestress=False)
    fem2.geom([2, 2], LY=4.)
    fem2.assign([mat1, mat2])
    fem2.bcleft(0.)
    fem2.bcbot(0.)
    fem2.bcright(0., 'force')
    fem2.bctop(0.1 * fem2.leny, 'disp')
    fem2.mesh(NX=4, NY=4)
    solve_case('fem2_sdim3', fem2)
    print(f'sdim3 fixtures {time.time() - t0:.1f}s; fem2 glob epl[1] = {fem2.glob["epl"][1]} (tests/test_basic.py:17: 0.04966042764325635)')


def gen_svc3():
    """sdim = 3 machine-learned yield function (tests/test_ml.py:10-34, test_ml_plasticity): an SVC trained by the
    reference in cylindrical principal-stress coordinates (setup_yf_SVM_3D, material.py:1280-1400; features
    create_scaled_input 2331-2334; gradient through the Jacobian of calc_fgrad 780-805), its yield function / gradient /
    ML_full_yf / response on seeded shear-free states, and the stress-strain data of calc_properties (3091-3166) that the
    reference test pins."""
    t0 = time.time()
    E, nu, sy = 200.e3, 0.3, 150.
    mat_h = FE.Material(name='anisotropic Hill')
    mat_h.elasticity(E=E, nu=nu)
    mat_h.plasticity(sy=sy, hill=[0.7, 1., 1.4], drucker=0., khard=0., sdim=3)
    m = FE.Material(name='ML flow rule')
    m.elasticity(E=E, nu=nu)
    m.plasticity(sy=sy, sdim=3)
    x_train, y_train = m.create_sig_data(36, mat_ref=mat_h, extend=True)
    train_sc, test_sc = m.setup_yf_SVM_3D(x_train, y_train, C=10, gamma=4., fs=0.3)
    out = svc_arrays(m)
    out['train_score'] = float(train_sc)
    print(f'  trained: {len(out["dual"])} support vectors, Ndof {m.Ndof}, scale_seq {m.scale_seq}')
    for planar, tag in [(False, 'pe'), (True, 'ps')]:
        rng = np.random.default_rng(177 + planar)
        n = 70
        CV = cv_plane_stress(m) if planar else np.array(m.CV)
        sig = np.zeros((n, 6)); epl = np.zeros((n, 6)); deps = np.zeros((n, 6))
        for i in range(n):
            d = np.zeros(6)
            d[0:3] = rng.standard_normal(3)
            if planar:
                d[2] = 0.
            u = rng.uniform()
            r = rng.uniform(0.05, 0.85) if u < 0.35 else (rng.uniform(0.85, 1.0) if u < 0.7 else rng.uniform(0.99, 1.02))
            sig[i] = d / mat_h.calc_seq(d) * r * sy          # scaled by the Hill surface the SVC was trained on
            de = np.zeros(6)
            de[0:3] = rng.standard_normal(3)
            if planar:
                de[2] = 0.
            if i % 5 == 0:
                de = np.linalg.lstsq(m.CV, sig[i] + 1e-3 * de, rcond=None)[0]
                de[3:6] = 0.
                if planar:
                    de[2] = 0.
            deps[i] = de / FE.eps_eq(de) * [5.e-4, 5.e-4, 2.e-3, 1.e-4][i % 4]
        fy = np.zeros(n); so = np.zeros((n, 6)); dp = np.zeros((n, 6)); ct = np.zeros((n, 6, 6)); ns = np.zeros(n, dtype=np.int32)
        seq = np.zeros(n); yf = np.zeros(n); full = np.zeros(n); fg = np.zeros((n, 3)); ed = np.zeros((n, 6)); ctan = np.zeros((n, 6, 6))
        for i in range(n):
            m.msg['nsteps'] = -1
            fy[i], so[i], dp[i], ct[i] = m.response(sig[i], epl[i], deps[i], CV)
            ns[i] = m.msg['nsteps']
            seq[i] = m.calc_seq(sig[i])
            yf[i] = m.calc_yf(sig[i], epl=epl[i])
            full[i] = m.ML_full_yf(sig[i], epl=epl[i])
            fg[i] = m.calc_fgrad(FE.sig_princ(sig[i])[0], epl=epl[i, 0:3])
            ed[i] = m.epl_dot(sig[i], epl[i], CV, deps[i])
            ctan[i] = m.C_tan(sig[i], CV, epl=epl[i])
        out.update({f'{tag}_CV': CV, f'{tag}_sig': sig, f'{tag}_epl': epl, f'{tag}_deps': deps, f'{tag}_fy': fy,
                    f'{tag}_sig_out': so, f'{tag}_depl': dp, f'{tag}_ctan': ct, f'{tag}_nsteps': ns, f'{tag}_seq': seq,
                    f'{tag}_yf': yf, f'{tag}_full_yf': full, f'{tag}_fgrad3': fg, f'{tag}_epl_dot': ed, f'{tag}_C_tan': ctan})
        el = np.sum((ns == 0) & (np.abs(ct - CV).max(axis=(1, 2)) == 0))
        print(f'  svc3/{tag}: elastic {el}, 50-step {np.sum(ns == 49)}, other {n - el - np.sum(ns == 49)}  ({time.time() - t0:.0f}s)')
    # the reference test itself (tests/test_ml.py:29-34)
    m.calc_properties(eps=0.01, sigeps=True, min_step=12)
    for sel in ('stx', 'sty', 'et2', 'ect'):
        for key in ('ys', 'seq', 'eeq', 'peeq'):
            out[f'propJ2_{sel}_{key}'] = np.array(m.propJ2[sel][key])
        out[f'sigeps_{sel}_sig'] = np.array(m.sigeps[sel]['sig'])
        out[f'sigeps_{sel}_epl'] = np.array(m.sigeps[sel]['epl'])
    print('  test_ml.py:32-34:', m.propJ2['stx']['ys'] - 149.62302821433968, m.propJ2['sty']['seq'][-1] - 157.25971534002542,
          m.propJ2['ect']['peeq'][-1] - 0.00855380746615942)
    np.savez_compressed(os.path.join(HERE, 'svc3_mlplast.npz'), **out)
    print(f'svc3 fixture {time.time() - t0:.1f}s')


def gen_sdim3_shear():
    """sdim = 3 materials on ROTATED planar states.  sig_princ orders the eigenvalues that LAPACK's dgeev returns by "major
    force axes" (basic.py:155-170); for every stress with s13 = s23 = 0 -- all a 2-d model produces -- that order is the
    eigenvalue of the eigenvector closest to each coordinate axis (checked on 20 000 random planar states against the
    reference; for general 3-d states the reference's result depends on dgeev's internal order and differs from this in 45 %
    of random cases).  Fixtures: Material.response of the reference on planar states with shear, and a sheared model."""
    t0 = time.time()
    m = FE.Material(name='hill3s')
    m.elasticity(E=200.e3, nu=0.3)
    m.plasticity(sy=150., hill=[0.7, 1., 1.4], khard=100., sdim=3)
    out = dict(E=200.e3, nu=0.3, sy=150., khard=100., hill=np.array(m.hill, dtype=float), drucker=0., mat_CV=np.array(m.CV))
    rng = np.random.default_rng(91)
    n = 240
    CV = np.array(m.CV)
    sig = np.zeros((n, 6)); epl = np.zeros((n, 6)); deps = np.zeros((n, 6))
    for i in range(n):
        if i % 3:
            e = rng.standard_normal(3) * 0.004
            epl[i, 0:3] = e - e.mean()
        d = np.zeros(6)
        d[[0, 1, 2, 5]] = rng.standard_normal(4)          # planar, sheared: principal axes rotated about z by any angle
        sflow = m.get_sflow(epl[i])
        u = rng.uniform()
        r = rng.uniform(0., 0.9) if u < 0.35 else (rng.uniform(0.9, 1.0) if u < 0.7 else rng.uniform(0.996, 1.004))
        sig[i] = d / m.calc_seq(d) * r * sflow
        de = np.zeros(6)
        de[[0, 1, 5]] = rng.standard_normal(3)
        deps[i] = de / FE.eps_eq(de) * [5.e-4, 5.e-4, 2.e-3, 1.e-2, 1.e-4][i % 5]
    fy = np.zeros(n); so = np.zeros((n, 6)); dp = np.zeros((n, 6)); ct = np.zeros((n, 6, 6)); ns = np.zeros(n, dtype=np.int32)
    seq = np.zeros(n); sp = np.zeros((n, 3))
    for i in range(n):
        m.msg['nsteps'] = -1
        fy[i], so[i], dp[i], ct[i] = m.response(sig[i], epl[i], deps[i], CV)
        ns[i] = m.msg['nsteps']
        seq[i] = m.calc_seq(sig[i])
        sp[i] = FE.sig_princ(sig[i])[0]
    out.update(pe_CV=CV, pe_sig=sig, pe_epl=epl, pe_deps=deps, pe_fy=fy, pe_sig_out=so, pe_depl=dp, pe_ctan=ct, pe_nsteps=ns,
               pe_seq=seq, pe_sig_princ=sp)
    print(f'  hill3 sheared states: 50-step {np.sum(ns == 49)}, elastic {np.sum(np.abs(ct - CV).max(axis=(1, 2)) == 0)}')
    np.savez_compressed(os.path.join(HERE, 'response_hill3_shear.npz'), **out)
    # tension + shear of a 4 x 4 plane-strain block of the same material: principal axes rotated by 20-35 degrees.  (Pure
    # simple shear puts them at exactly 45 degrees, where argmax |ev| of basic.py:156 is a tie decided by rounding noise of
    # LAPACK's eigenvectors: the reference's own result is then not reproducible -- a first version of this fixture
    # agreed with the device for 7 increments to 1e-12 and then flipped.)
    fe = FE.Model(dim=2, planestress=False)
    fe.geom([4.], LY=4.)
    fe.assign([m])
    fe.bcleft(0., 'force')
    fe.bcright(0., 'force')
    fe.bcbot(0., 'disp', 'x')
    fe.bcbot(0., 'disp', 'y')
    fe.bctop(0.006 * fe.leny, 'disp', 'x')
    fe.bctop(0.008 * fe.leny, 'disp', 'y')
    fe.mesh(NX=4, NY=4)
    solve_case('shear_hill3_4x4', fe, min_step=4)
    print(f'sdim3 shear fixtures {time.time() - t0:.1f}s')


class _RbfSVC:
    """Duck-typed stand-in for a fitted sklearn.svm.SVC(kernel='rbf') with GIVEN support vectors (sklearn cannot be
    constructed from parameters): the attributes and the decision function the reference reads (material.py:401-405,
    798-799), f(x) = sum_i dual_i exp(-gamma |x - sv_i|^2) + intercept (libsvm k_function)."""

    def __init__(self, sv, dual, rho, gamma):
        self.support_vectors_ = np.array(sv)
        self.dual_coef_ = np.array(dual)[None, :]
        self.intercept_ = np.array([rho])
        self.gamma = gamma

    def decision_function(self, x):
        x = np.atleast_2d(x)
        d2 = ((x[:, None, :] - self.support_vectors_[None, :, :]) ** 2).sum(axis=2)
        return np.exp(-self.gamma * d2) @ self.dual_coef_[0] + self.intercept_[0]

    def predict(self, x):
        return np.sign(self.decision_function(x))


def gen_umat_models():
    """The three trained flow rules the reference ships for its Abaqus UMAT (examples/UMAT/models/*-svm.csv, written by
    export_MLparam of pyLabFEA 4.0: Nset at index 16 where 4.4 writes the dev_only flag; ml_umat.f:129-153 reads them with
    dev_only = false).  The reference has no loader (from_MLparam is a stub, material.py:2688-2703): the parameters are put
    on a reference Material by hand and its own calc_yf / calc_fgrad / ML_full_yf / response are evaluated on them."""
    import json
    src = os.path.join(HERE, 'umat_models')
    for nm in ('abq_ML-Hill-p1_C15_G25', 'abq_ML-J2_C15_G25', 'abq_ML-Goss-Barlat_C15_G25'):
        t0 = time.time()
        props = np.loadtxt(os.path.join(src, nm + '-svm.csv'), delimiter=',').ravel()
        meta = json.load(open(os.path.join(src, nm + '-svm_meta.json')))
        names = meta['Data']['Names']
        assert names[16] == 'Nset'   # layout of pyLabFEA 4.0 - 4.3
        nsv, nsd = int(props[0]), int(props[1])
        C11, C12, C44, rho, gamma, epc, scale_seq, scale_wh = props[2:10]
        dual = props[29:29 + nsv]
        sv = props[29 + nsv:29 + nsv + nsv * nsd].reshape(nsv, nsd)
        m = FE.Material(name=nm)
        m.elasticity(C11=C11, C12=C12, C44=C44)
        m.plasticity(sy=scale_seq, sdim=6)
        m.svm_yf = _RbfSVC(sv, dual, rho, gamma)
        m.ML_yf, m.ML_grad = True, False
        m.gam_yf, m.scale_seq, m.scale_wh, m.Ndof, m.dev_only = gamma, scale_seq, scale_wh, nsd, False
        rng = np.random.default_rng(17)
        n = 120
        d = rng.standard_normal((n, 6))
        d[: n // 2, 3:5] = 0.                                   # half of them planar
        seq = np.array([FE.sig_eq_j2(x) for x in d])
        sig = d / seq[:, None] * (scale_seq * rng.uniform(0.2, 1.3, n))[:, None]
        yf = np.array([m.calc_yf(s) for s in sig])
        fg = np.array([m.calc_fgrad(s) for s in sig])
        full = np.array([m.ML_full_yf(s) for s in sig[:40]])
        nr = 24
        CV = np.array(m.CV)
        deps = rng.standard_normal((nr, 6))
        deps = deps / np.array([FE.eps_eq(x) for x in deps])[:, None] * 4.e-4
        rs = np.zeros((nr, 6)); rd = np.zeros((nr, 6)); rc = np.zeros((nr, 6, 6)); rf = np.zeros(nr); rn = np.zeros(nr, dtype=np.int32)
        for i in range(nr):
            m.msg['nsteps'] = -1
            rf[i], rs[i], rd[i], rc[i] = m.response(sig[i], np.zeros(6), deps[i], CV)
            rn[i] = m.msg['nsteps']
        np.savez_compressed(os.path.join(HERE, 'umat_' + nm.split('_')[1] + '.npz'), model=nm, sig=sig, yf=yf, fgrad=fg, full_yf=full,
                            deps=deps, r_fy=rf, r_sig=rs, r_depl=rd, r_ctan=rc, r_nsteps=rn, CV=CV, scale_seq=scale_seq, nsv=nsv)
        print(f'  {nm}: nsv {nsv}, yf in [{yf.min():.3f}, {yf.max():.3f}], 50-step {np.sum(rn == 49)} of {nr}, {time.time() - t0:.0f}s')


def gen_svc_wh():
    """SVC flow rule WITH work-hardening features (Material.whdat): the recipe of examples/train_hardening.py on a reduced
    data set (Hill reference material with linear hardening khard = 1000; 45 load cases, plastic strain up to 2 %), trained
    by the reference's train_SVC.  Feature vector (material.py:2336-2346): 6 stresses / scale_seq, 6 plastic strains /
    scale_wh, accumulated strain, maximum stress / scale_seq, flag -- the last three are 0 in every call of the FE path.
    Outputs of the reference: calc_yf, calc_fgrad (with the khard it leaves on the material, 808-814), ML_full_yf and
    Material.response, each called on a FRESH copy of khard (the reference keeps khard on the Material object, so the
    result of a call depends on which gradient was evaluated last; the fixture pins the per-call behaviour)."""
    from scipy.optimize import fsolve
    t0 = time.time()
    E, nu, sy, khard = 200.e3, 0.3, 50., 1000.
    rv = [1.2, 1.0, 0.8, 1.0, 1.0, 1.0]
    mat_h = FE.Material(name='Hill-reference', num=1)
    mat_h.elasticity(E=E, nu=nu)
    mat_h.plasticity(sy=sy, rv=rv, khard=khard, sdim=6)
    Nlc, epl_max, depl = 45, 0.02, 2.e-3
    nl3d = int(Nlc / 3)
    sunit_all = FE.load_cases(number_3d=nl3d, number_6d=Nlc - nl3d)
    x1 = fsolve(mat_h.find_yloc, np.ones(Nlc) * mat_h.sy, args=(sunit_all,), xtol=1.e-5)
    sig_ideal = sunit_all * x1[:, None]
    SV = np.linalg.inv(mat_h.CV)
    lc = dict()
    for i, st in enumerate(sig_ideal):          # create_data of train_hardening.py:24-82
        epl = np.zeros(6)
        peeq = 0.
        sl, el, tl = [], [], []
        seq = FE.sig_eq_j2(st)
        su = st / seq
        ind = np.where(su > 0., 1, np.where(su < 0., 2, 0))
        key = f'Us_A{ind[0]}B{ind[1]}C{ind[2]}D{ind[3]}E{ind[4]}F{ind[5]}_HI{i:03d}_NNNNN_Tx_NN'
        for j in range(6):
            s_ = su * j * seq / 5
            sl.append(s_); el.append(np.array(epl)); tl.append(SV @ s_)
        while peeq < epl_max:
            peeq = FE.eps_eq(epl) + depl
            s_ = su * (seq + peeq * khard)
            epl = epl + mat_h.calc_fgrad(sig=s_, epl=epl) * depl
            sl.append(s_); el.append(np.array(epl)); tl.append(epl + SV @ s_)
        sl, el, tl = np.array(sl), np.array(el), np.array(tl)
        lc[key] = {'Stress': sl, 'Eq_Stress': FE.sig_eq_j2(sl), 'Strain_Plastic': el, 'Eq_Strain_Plastic': FE.eps_eq(el),
                   'Shifted_Strain_Plastic': None, 'Strain_Total': tl, 'Eq_Strain_Total': FE.eps_eq(tl)}
    data = FE.Data(lc, mat_name='ML_Hill_hardening', epl_start=0.0, epl_crit=0.0, epl_max=epl_max, depl=depl, wh_data=True)
    m = FE.Material(name='ML_Hill_hardening', num=2)
    m.from_data(data.mat_data)
    m.train_SVC(C=2.0, gamma=1.5, Ce=0.99, Fe=0.1, Nseq=8, gridsearch=False)
    assert m.whdat and m.Ndof == 15 and m.ind_wh == 6
    sv = np.array(m.svm_yf.support_vectors_)
    print(f'  trained: {len(sv)} support vectors, scale_seq {m.scale_seq:.3f}, scale_wh {m.scale_wh:.5f}, sy {m.sy:.3f}, {time.time() - t0:.0f}s')
    rng = np.random.default_rng(23)
    n = 150
    d = rng.standard_normal((n, 6))
    d[: n // 2, 3:5] = 0.
    peeq = np.where(rng.uniform(size=n) < 0.3, 0., rng.uniform(0., 0.015, n))
    sig = np.zeros((n, 6)); epl = np.zeros((n, 6))
    for i in range(n):
        su = d[i] / FE.sig_eq_j2(d[i])
        g = mat_h.calc_fgrad(sig=su * sy)
        epl[i] = g / FE.eps_eq(g) * peeq[i]
        sig[i] = su * (sy + khard * peeq[i]) * rng.uniform(0.3, 1.25)
    k0 = float(m.khard)
    yf = np.array([m.calc_yf(sig[i], epl=epl[i]) for i in range(n)])
    fg = np.zeros((n, 6)); kh = np.zeros(n)
    for i in range(n):
        m.khard = k0
        fg[i] = m.calc_fgrad(sig[i], epl=epl[i])
        kh[i] = m.khard
    nf = 40
    full = np.zeros(nf)
    for i in range(nf):
        m.khard = k0
        full[i] = m.ML_full_yf(sig[i], epl[i])
    nr = 36
    CV = np.array(m.CV)
    deps = rng.standard_normal((nr, 6))
    deps = deps / np.array([FE.eps_eq(x) for x in deps])[:, None] * np.array([2.e-4, 5.e-4, 1.e-3])[np.arange(nr) % 3][:, None]
    rs = np.zeros((nr, 6)); rd = np.zeros((nr, 6)); rc = np.zeros((nr, 6, 6)); rf = np.zeros(nr); rn = np.zeros(nr, dtype=np.int32)
    rk = np.zeros(nr)
    for i in range(nr):
        m.khard = k0
        m.msg['nsteps'] = -1
        rf[i], rs[i], rd[i], rc[i] = m.response(sig[i], epl[i], deps[i], CV)
        rn[i] = m.msg['nsteps']
        rk[i] = m.khard
    m.khard = k0
    np.savez_compressed(os.path.join(HERE, 'svc_wh.npz'), sv=sv, dual=np.array(m.svm_yf.dual_coef_[0]), rho=float(m.svm_yf.intercept_[0]),
                        gamma=float(m.gam_yf), scale_seq=float(m.scale_seq), scale_wh=float(m.scale_wh), sy=float(m.sy), khard0=k0,
                        dev_only=int(m.dev_only), CV=CV, sig=sig, epl=epl, yf=yf, fgrad=fg, khard_after_fgrad=kh, full_yf=full,
                        deps=deps, r_fy=rf, r_sig=rs, r_depl=rd, r_ctan=rc, r_nsteps=rn, r_khard_after=rk)
    print(f'  yf in [{yf.min():.3f}, {yf.max():.3f}], khard after fgrad in [{kh.min():.1f}, {kh.max():.1f}], '
          f'response: 50-step {np.sum(rn == 49)}, 1-step {np.sum(rn == 0)} of {nr}; {time.time() - t0:.0f}s')


if __name__ == '__main__':
    what = sys.argv[1:] or ['all']
    if 'all' in what:
        what = ['response', 'kmat', 'solve', 'svc', 'barlat', 'barlat_response', 'barlat_solve', 'mlparam', 'mlparam3', 'tresca', 'sdim3', 'svc3']
    for w in what:
        print('==', w)
        globals()['gen_' + w]()
