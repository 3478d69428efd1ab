"""ctypes wrapper around oracle/libplf_oracle.so -- TEST INFRASTRUCTURE ONLY.

The CPU restatement of the reference hot path (see plf_oracle.c).  Importable from tests/, from
__graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference legs, and from nowhere else: the
product package pylabfea_b200 must never import this module (tests/test_no_fallback.py enforces it).
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, 'libplf_oracle.so')

ELASTIC, HILL, SVC, BARLAT, HILL3, SVC3 = 0, 1, 2, 3, 4, 5

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)


class OrcMat(C.Structure):
    _fields_ = [('kind', C.c_int32), ('nsv', C.c_int32), ('dev_only', C.c_int32), ('tresca', C.c_int32),
                ('lhs', C.c_int32), ('pad_', C.c_int32),
                ('E', C.c_double), ('nu', C.c_double), ('sy', C.c_double), ('khard', C.c_double),
                ('hill', C.c_double * 6), ('d', C.c_double * 3),
                ('bar_m1', C.c_double * 36), ('bar_m2', C.c_double * 36), ('bar_exp', C.c_double),
                ('gamma', C.c_double), ('rho', C.c_double), ('scale_seq', C.c_double),
                ('sv', _dp), ('dual', _dp), ('nfeat', C.c_int32), ('pad2_', C.c_int32), ('scale_wh', C.c_double)]


def build(force=False):
    src = os.path.join(HERE, 'plf_oracle.c')
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', HERE, '-B', 'CC=gcc'], stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        assert _lib.orc_sizeof_mat.restype is not None
        _lib.orc_sizeof_mat.restype = C.c_size_t
        assert _lib.orc_sizeof_mat() == C.sizeof(OrcMat), 'orc_mat layout mismatch'
        for n in ('orc_eps_eq', 'orc_calc_seq', 'orc_calc_yf', 'orc_ml_full_yf', 'orc_svc_decision', 'orc_brentq'):
            getattr(_lib, n).restype = C.c_double
    return _lib


def _a(x, dt=np.float64):
    return np.ascontiguousarray(x, dtype=dt)


def _p(x):
    return x.ctypes.data_as(_dp if x.dtype == np.float64 else _ip)


class Mat:
    """Material descriptor for the oracle.  Keeps numpy arrays alive for the SV table."""

    def __init__(self, kind=HILL, E=0., nu=0., sy=0., khard=0., hill=None, drucker=0., lhs=None, svc=None,
                 barlat=None, tresca=False):
        m = OrcMat()
        m.kind = kind
        m.E, m.nu, m.sy, m.khard = E, nu, sy, khard
        hill = np.ones(6) if hill is None else np.asarray(hill, dtype=float)
        m.hill[:] = (list(hill) + [1., 1., 1.])[:6]
        if lhs is not None:
            m.lhs = 1
            m.d[:] = list(lhs)
        else:
            m.d[:] = [drucker * 1.0] * 3
        m.tresca = int(tresca)
        if svc is not None:
            sv = np.atleast_2d(np.asarray(svc['sv'], dtype=float))
            if sv.shape[1] < 6:   # sdim = 3 models have 2 features: the table always has 6 columns, the rest zero
                sv = np.hstack([sv, np.zeros((len(sv), 6 - sv.shape[1]))])
            if sv.shape[1] == 15:   # work-hardening features (Material.whdat)
                m.nfeat = 15
                m.scale_wh = float(svc['scale_wh'])
            self._sv = _a(sv)
            self._dual = _a(svc['dual'])
            m.nsv = len(self._dual)
            m.sv = _p(self._sv)
            m.dual = _p(self._dual)
            m.gamma, m.rho, m.scale_seq = float(svc['gamma']), float(svc['rho']), float(svc['scale_seq'])
            m.dev_only = int(svc.get('dev_only', 0))
        if barlat is not None:
            m.bar_m1[:] = list(np.asarray(barlat['Bar_m1'], dtype=float).ravel())
            m.bar_m2[:] = list(np.asarray(barlat['Bar_m2'], dtype=float).ravel())
            m.bar_exp = float(barlat['barlat_exp'])
        self.c = m

    @classmethod
    def from_golden(cls, g):
        """From a tests/golden/response_*.npz or svc_*.npz dict."""
        if 'sv' in g:
            return cls(kind=SVC3 if int(g['sdim']) == 3 else SVC, E=float(g['E']), nu=float(g['nu']), sy=float(g['sy']), khard=0.,
                       hill=g['hill'], drucker=float(g['drucker']),
                       svc=dict(sv=g['sv'], dual=g['dual'], gamma=g['gamma'], rho=g['rho'],
                                scale_seq=g['scale_seq'], dev_only=int(g['dev_only'])))
        if 'sy' not in g:
            return cls(kind=ELASTIC, E=float(g['E']), nu=float(g['nu']))
        return cls(kind=HILL, E=float(g['E']), nu=float(g['nu']), sy=float(g['sy']), khard=float(g['khard']),
                   hill=g['hill'], drucker=float(g['drucker']))


def response_batch(mats, mat_id, sig, epl, deps, CVs):
    """Material.response for n independent elements.  mats: list[Mat]; CVs: (nmat,6,6)."""
    L = lib()
    sig, epl, deps = _a(sig), _a(epl), _a(deps)
    n = len(sig)
    arr = (OrcMat * len(mats))(*[m.c for m in mats])
    CVs = _a(CVs).reshape(len(mats), 36)
    mat_id = _a(mat_id if mat_id is not None else np.zeros(n), np.int32)
    fy = np.zeros(n)
    so = np.zeros((n, 6))
    dp = np.zeros((n, 6))
    ct = np.zeros((n, 6, 6))
    ns = np.zeros(n, dtype=np.int32)
    L.orc_response_batch(arr, C.c_int64(n), _p(mat_id), _p(sig), _p(epl), _p(deps), _p(CVs), _p(fy), _p(so), _p(dp),
                         _p(ct), _p(ns))
    return fy, so, dp, ct, ns


def response_batch_threads(mats, mat_id, sig, epl, deps, CVs, nthreads):
    """Same, sharded over host threads (ctypes releases the GIL) -- the multi-core CPU baseline."""
    from concurrent.futures import ThreadPoolExecutor
    n = len(sig)
    mat_id = np.zeros(n, np.int32) if mat_id is None else np.asarray(mat_id, np.int32)
    cuts = np.linspace(0, n, nthreads * 4 + 1).astype(int)
    parts = [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1) if cuts[i + 1] > cuts[i]]
    with ThreadPoolExecutor(nthreads) as ex:
        res = list(ex.map(lambda ab: response_batch(mats, mat_id[ab[0]:ab[1]], sig[ab[0]:ab[1]], epl[ab[0]:ab[1]],
                                                    deps[ab[0]:ab[1]], CVs), parts))
    return tuple(np.concatenate([r[k] for r in res]) for k in range(5))


def step_batch(mats, sig, epl, deps, CVs, B4, jacw, nthreads=1):
    """bench.py's CPU step: Material.response for every element + Kel of the new tangent, sharded over host threads
    (one C call per shard; ctypes releases the GIL).  Returns (fy, sig, depl, ctan, nsteps)."""
    from concurrent.futures import ThreadPoolExecutor
    L = lib()
    L.orc_step_batch.restype = C.c_double
    sig, epl, deps = _a(sig), _a(epl), _a(deps)
    n = len(sig)
    arr = (OrcMat * len(mats))(*[m.c for m in mats])
    CVs = _a(CVs).reshape(len(mats), 36)
    B4 = _a(B4)
    mid = np.zeros(n, dtype=np.int32)
    fy = np.zeros(n)
    so = np.zeros((n, 6))
    dp = np.zeros((n, 6))
    ct = np.zeros((n, 36))
    ns = np.zeros(n, dtype=np.int32)
    cuts = np.linspace(0, n, max(1, nthreads) * 4 + 1).astype(int)

    def run(ab):
        a, b = ab
        if b <= a:
            return 0.
        return L.orc_step_batch(arr, C.c_int64(b - a), _p(mid[a:b]), _p(sig[a:b]), _p(epl[a:b]), _p(deps[a:b]), _p(CVs),
                                _p(B4), C.c_double(jacw), _p(fy[a:b]), _p(so[a:b]), _p(dp[a:b]), _p(ct[a:b]), _p(ns[a:b]), None)
    parts = list(zip(cuts[:-1], cuts[1:]))
    if nthreads > 1:
        with ThreadPoolExecutor(nthreads) as ex:
            list(ex.map(run, parts))
    else:
        for p in parts:
            run(p)
    return fy, so, dp, ct.reshape(n, 6, 6), ns


def calc_seq(mat, sig):
    sig = _a(sig).reshape(-1, 6)
    out = np.zeros(len(sig))
    lib().orc_calc_seq_batch(C.byref(mat.c), C.c_int64(len(sig)), _p(sig), _p(out))
    return out


def calc_yf(mat, sig, epl=None):
    sig = _a(sig).reshape(-1, 6)
    epl = np.zeros_like(sig) if epl is None else _a(epl).reshape(-1, 6)
    out = np.zeros(len(sig))
    lib().orc_calc_yf_batch(C.byref(mat.c), C.c_int64(len(sig)), _p(sig), _p(epl), _p(out))
    return out


def calc_fgrad(mat, sig, epl=None):
    sig = _a(sig).reshape(-1, 6)
    epl = np.zeros_like(sig) if epl is None else _a(epl).reshape(-1, 6)
    out = np.zeros_like(sig)
    lib().orc_calc_fgrad_batch(C.byref(mat.c), C.c_int64(len(sig)), _p(sig), _p(epl), _p(out))
    return out


def ml_full_yf(mat, sig, epl=None):
    sig = _a(sig).reshape(-1, 6)
    out = np.zeros(len(sig))
    ep = None if epl is None else _p(_a(epl).reshape(-1, 6))
    lib().orc_ml_full_yf_batch(C.byref(mat.c), C.c_int64(len(sig)), _p(sig), ep, _p(out))
    return out


def epl_dot(mat, sig, epl, Cel, deps):
    sig, epl, deps, Cel = _a(sig).reshape(-1, 6), _a(epl).reshape(-1, 6), _a(deps).reshape(-1, 6), _a(Cel)
    out = np.zeros_like(sig)
    lib().orc_epl_dot_batch(C.byref(mat.c), C.c_int64(len(sig)), _p(sig), _p(epl), _p(Cel), _p(deps), _p(out))
    return out


def C_tan(mat, sig, epl, Cel):
    sig, epl, Cel = _a(sig).reshape(-1, 6), _a(epl).reshape(-1, 6), _a(Cel)
    out = np.zeros((len(sig), 6, 6))
    lib().orc_C_tan_batch(C.byref(mat.c), C.c_int64(len(sig)), _p(sig), _p(epl), _p(Cel), _p(out))
    return out


def element_CV(matCV, E, nu, C44, planestress):
    out = np.zeros((6, 6))
    m = _a(matCV)
    lib().orc_element_CV(_p(m), C.c_double(E), C.c_double(nu), C.c_double(C44), C.c_int(int(planestress)), _p(out))
    return out


def gauss_bmats(lx, ly, planestress, CV, E, nu):
    out = np.zeros((4, 6, 8))
    CV = _a(CV)
    lib().orc_gauss_bmats(C.c_double(lx), C.c_double(ly), C.c_int(int(planestress)), _p(CV), C.c_double(E),
                          C.c_double(nu), _p(out))
    return out


def kel(B4, elstiff, jacw):
    out = np.zeros((8, 8))
    B4, elstiff = _a(B4), _a(elstiff)
    lib().orc_kel(_p(B4), _p(elstiff), C.c_double(jacw), _p(out))
    return out


def setupK_dense(nodes, Kel, Ndof):
    nodes, Kel = _a(nodes, np.int32), _a(Kel)
    K = np.zeros((Ndof, Ndof))
    lib().orc_setupK_dense(C.c_int64(len(nodes)), _p(nodes), _p(Kel), C.c_int64(Ndof), _p(K))
    return K


def setupK_csr(nodes, Kel, rowptr, col):
    nodes, Kel = _a(nodes, np.int32), _a(Kel)
    rowptr, col = _a(rowptr, np.int32), _a(col, np.int32)
    val = np.zeros(len(col))
    rc = lib().orc_setupK_csr(C.c_int64(len(nodes)), _p(nodes), _p(Kel), _p(rowptr), _p(col), _p(val))
    if rc != 0:
        raise ValueError('CSR pattern does not contain an assembled entry')
    return val


def element_strain(B4, ue):
    out = np.zeros(6)
    B4, ue = _a(B4), _a(ue)
    lib().orc_element_strain(_p(B4), _p(ue), _p(out))
    return out


def brentq(f, xa, xb, xtol=1.e-5, rtol=4 * np.finfo(float).eps, maxiter=100):
    FN = C.CFUNCTYPE(C.c_double, C.c_double, C.c_void_p)
    cb = FN(lambda x, _: float(f(x)))
    conv = C.c_int(0)
    L = lib()
    L.orc_brentq.argtypes = [FN, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_int)]
    r = L.orc_brentq(cb, None, xa, xb, xtol, rtol, maxiter, C.byref(conv))
    return r, bool(conv.value)
