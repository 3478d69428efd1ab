"""Material -- host-side mirror of pyLabFEA's Material class for the stress-update path.

Same method names, argument meaning and error behaviour as the reference (file:line cited per method, paths relative to
/root/reference/src/pylabfea/material.py); every constitutive evaluation runs as a batched sm_100a kernel through the
C-ABI (include/plfb200.h).  There is no CPU path: without libplfb200.so and a CUDA device the calls raise.

Scope: sdim = 6 materials -- J2 / Hill-6p with linear isotropic hardening and Drucker / LHS
terms, the RBF-SVC machine-learned yield function (parameters adopted from a trained scikit-learn SVC or from a
reference Material object), Barlat Yld2004-18p equivalent stress.  Training (train_SVC and friends), plotting and
the ML variant of the sdim = 3 principal-stress formulation stay with the reference package.
"""
import warnings

import numpy as np

from . import _lib
from .basic import eps_eq

_ctx = None


def default_context(device=0):
    """The process-wide device context used by Material calls made outside a Model."""
    global _ctx
    if _ctx is None:
        _ctx = _lib.Context(device)
    return _ctx


class Material(object):
    """Elastic / elastic-plastic material (reference: material.py:57-203 for the attribute set)."""

    def __init__(self, name='Material', num=1):
        self.name = name
        self.num = num
        self.E = self.nu = None
        self.C11 = self.C12 = self.C44 = None
        self.CV = None
        self.sy = None          # elasticity is considered unless sy is set
        self.sy0 = None
        self.khard = None
        self.drucker = None
        self.lhs = None
        self.hill = None
        self.hill_3p = False
        self.hill_6p = False
        self.tresca = False
        self.barlat = False
        self.Bar_m1 = self.Bar_m2 = None
        self.barlat_exp = None
        self.sdim = None
        self.ML_yf = False
        self.ML_grad = False
        self.dev_only = False
        self.whdat = False
        self.txdat = False
        self.Ndof = 2
        self.svm_yf = None
        self.gam_yf = None
        self.scale_seq = None
        self._svc = None        # (support vectors, dual coefs, intercept)
        self.msg = {'yield_fct': None, 'gradient': None, 'nsteps': 0, 'equiv': None}
        self.prop = {k: {'ys': None, 'seq': None, 'eeq': None, 'peeq': None, 'style': None, 'name': None}
                     for k in ('stx', 'sty', 'et2', 'ect')}
        self.propJ2 = {k: {'ys': None, 'seq': None, 'eeq': None, 'peeq': None} for k in ('stx', 'sty', 'et2', 'ect')}
        self.sigeps = {k: {'sig': None, 'eps': None, 'epl': None} for k in ('stx', 'sty', 'et2', 'ect')}

    # ------------------------------------------------------------------------------------------------------
    # parameter setters (host only)
    # ------------------------------------------------------------------------------------------------------
    def elasticity(self, C11=None, C12=None, C44=None, CV=None, E=None, nu=None):
        """Define elastic constants from (E, nu), (C11, C12, C44) or a Voigt matrix CV (material.py:2401-2464)."""
        if E is not None:
            if nu is None:
                raise ValueError('Error: Inconsistent definition of material parameters: Only E provided')
            if C11 is not None or C12 is not None or C44 is not None:
                raise ValueError('Error: Inconsistent definition of material parameters: E provided together with C_ij')
            fac = E / ((1. + nu) * (1. - 2. * nu))
            self.C11 = (1. - nu) * fac
            self.C12 = nu * fac
            self.C44 = (0.5 - nu) * fac
            self.E, self.nu = E, nu
        elif C11 is not None:
            if nu is not None:
                raise ValueError('Error: Inconsistent definition of material parameters: nu provided together with C_ij')
            if C12 is None or C44 is None:
                raise ValueError('Error: Inconsistent definition of material parameters: C_12 or C_44 values missing')
            self.C11, self.C12, self.C44 = C11, C12, C44
            self.nu = C12 / (C11 + C12)
            self.E = 2 * C44 * (1 + self.nu)
        elif CV is not None:
            self.CV = np.array(CV, dtype=float)
            self.C11, self.C12, self.C44 = self.CV[0, 0], self.CV[0, 1], self.CV[3, 3]
            self.nu = self.C12 / (self.C11 + self.C12)
            self.E = 2 * self.C44 * (1 + self.nu)
        else:
            raise ValueError('elasticity: Inconsistent definition of material parameters')
        if CV is None:
            M = np.zeros((6, 6))
            M[0:3, 0:3] = self.C12
            M[0, 0] = M[1, 1] = M[2, 2] = self.C11
            M[3, 3] = M[4, 4] = M[5, 5] = self.C44
            self.CV = M

    def plasticity(self, sy=None, sdim=6, drucker=0., khard=0., tresca=False, barlat=None, barlat_exp=None,
                   hill=None, hill_3p=None, hill_6p=None, rv=None, lhs=None):
        """Define plastic parameters (material.py:2466-2594)."""
        if sy < 0.:
            raise ValueError('Initial yield strength cannot be negative.')
        if khard < 0.:
            warnings.warn('Strain softening not supported. khard is set to 0.')
            khard = 0.
        if lhs is not None and not np.isclose(drucker, 0.0):
            raise ValueError('Parameters for Drucker-Prager and LHS plasticity models cannot be given at the same time.')
        if sdim not in (3, 6):
            raise ValueError('{} in plasticity: sdim must be either 3 or 6'.format(self.name))
        self.sy0 = self.sy = sy
        self.khard = khard
        self.drucker = drucker
        self.lhs = None if lhs is None else np.array(lhs, dtype=float)
        self.sdim = sdim
        if hill is None and rv is None:
            if lhs is not None:
                raise ValueError('LHS parameters for anisotropic yield asymmetry provided, '
                                 'but no anisotropy parameters for plastic yielding have been given.')
            hill = [1.] * sdim
        elif hill is None:
            if len(rv) != sdim:
                raise ValueError(f'plasticity: wrong dimension of yield stress ratios, must be {sdim}')
            r2 = (1. / np.array(rv, dtype=float)) ** 2   # (1/r)^2, same rounding as material.py:2524-2531
            hill = [1.] * sdim
            hill[0] = r2[0] + r2[1] - r2[2]
            hill[1] = r2[1] + r2[2] - r2[0]
            hill[2] = r2[2] + r2[0] - r2[1]
            if sdim == 6:
                hill[3], hill[4], hill[5] = r2[3], r2[4], r2[5]
        elif rv is not None:
            warnings.warn('plasticity: Both, hill and rv, have been provided. Using Hill parameters.')
        hill = list(hill)
        nh = len(hill)
        if hill_6p is None and hill_3p is None:
            hill_6p = (nh == 6)
            hill_3p = not hill_6p
            if hill_3p and hill[0] == 1. and hill[1] == 1. and hill[2] == 1.:
                hill_3p = False
        if hill_6p and nh != 6:
            raise ValueError('plasticity: When hill_6p is set True, 6 Hill parameters must be provided')
        if hill_3p and nh != 3:
            raise ValueError('plasticity: When hill_3p is set True, only 3 Hill parameters can be provided')
        if hill_6p and sdim == 3:
            warnings.warn('plasticity: 6 Hill parameters are provided, but sdim=3; ignoring shear parameters')
            hill_6p, hill_3p, hill = False, True, hill[0:3]
        if hill_3p and sdim == 6:
            warnings.warn('plasticity: 3 Hill parameters are provided, but sdim=6; shear parameters set to 1')
            hill_3p, hill_6p = False, True
            hill = hill + [1., 1., 1.]
        if sdim == 6 and len(hill) == 3:
            hill = hill + [1., 1., 1.]
        self.hill_6p, self.hill_3p = bool(hill_6p), bool(hill_3p)
        self.hill = np.array(hill, dtype=float)
        self.tresca = bool(tresca) if tresca is not None else False
        if barlat is not None:
            b = np.asarray(barlat, dtype=float)
            self.barlat = True

            def bmat(c):
                M = np.zeros((6, 6))
                M[0, 1], M[0, 2] = -c[0], -c[1]
                M[1, 0], M[1, 2] = -c[2], -c[3]
                M[2, 0], M[2, 1] = -c[4], -c[5]
                M[3, 3], M[4, 4], M[5, 5] = c[6], c[7], c[8]
                return M
            self.Bar_m1, self.Bar_m2 = bmat(b[0:9]), bmat(b[9:18])
            self.barlat_exp = barlat_exp
        else:
            self.barlat = False

    def set_svc(self, support_vectors, dual_coef, intercept, gamma, scale_seq=None, dev_only=False, scale_wh=None):
        """Adopt a trained RBF support-vector yield function (what setup_yf_SVM_6D leaves on the Material,
        material.py:1109-1278): sklearn's svm_yf.support_vectors_ (nsv,6), dual_coef_[0], intercept_[0], gamma and the
        feature scale scale_seq (= sy when no microstructure data is attached, 1160-1161).  A flow rule trained on
        work-hardening data (Material.whdat, train_hardening.py) has 15 features per support vector -- 6 stresses,
        6 plastic strains / scale_wh, accumulated strain, maximum stress, flag (material.py:2342-2346, 2667-2669) -- and
        needs scale_wh; its hardening rate khard follows from the gradient of the yield function (808-814)."""
        sv = np.ascontiguousarray(support_vectors, dtype=float)
        if self.sy is None:
            raise ValueError('set_svc: call plasticity(sy=...) first (train_SVC does so, material.py:1606)')
        nfeat = 6 if self.sdim == 6 else 2
        wh = self.sdim == 6 and sv.ndim == 2 and sv.shape[1] == 15
        if wh:
            if scale_wh is None or not scale_wh > 0.:
                raise ValueError('set_svc: support vectors with work-hardening features need scale_wh > 0')
            nfeat = 15
        if sv.ndim != 2 or sv.shape[1] != nfeat:
            raise ValueError('set_svc: support vectors must have shape (nsv, %d) for sdim = %d (features: %s)'
                             % (nfeat, self.sdim, 'scaled Voigt stress' if nfeat == 6 else 'seq_J2 / scale_seq - 1, polar angle / pi'))
        if nfeat == 2:
            # sdim = 3 (setup_yf_SVM_3D, material.py:1280-1400): the device table always has 6 columns, the rest zero
            sv = np.ascontiguousarray(np.hstack([sv, np.zeros((len(sv), 4))]))
        dc = np.ascontiguousarray(np.ravel(dual_coef), dtype=float)
        if len(dc) != len(sv):
            raise ValueError('set_svc: dual_coef must have one entry per support vector')
        self._svc = (sv, dc, float(intercept))
        self.gam_yf = float(gamma)
        self.scale_seq = float(self.sy if scale_seq is None else scale_seq)
        self.dev_only = bool(dev_only)
        self.ML_yf = True
        self.Ndof = nfeat
        self.whdat = bool(wh)
        self.scale_wh = float(scale_wh) if wh else 1.
        self.ind_wh = 6 if wh else None
        if not wh:
            self.khard = 0.   # material.py:812-814: the SVC gradient resets khard when there are no hardening features

    @classmethod
    def from_MLparam(cls, file, path='', sy=None, name=None):
        """Load a trained ML flow rule from the CSV (+ optional JSON metadata) that export_MLparam writes
        (material.py:2133-2273; the reference's own from_MLparam, 2688-2703, is unimplemented).  `file` is the trunk
        of the file name, as in export_MLparam: the CSV is path + file + '-svm.csv'.

        CSV layout (8 numbers per line, flattened): [0] nsv, [1] Ndof, [2:5] C11 C12 C44, [5] intercept (rho), [6] gamma,
        [7] epc, [8] scale_seq, [9] scale_wh, [10:16] C22 C33 C13 C23 C55 C66 (or -1), [16] -1 when dev_only, [17] Nset,
        [18:18+Nset] scale_text, [29:29+nsv] dual coefficients, then the (nsv, Ndof) support vectors.
        The yield strength is not part of the format; sy defaults to scale_seq (what train_SVC uses for scaling when no
        microstructure data is attached, material.py:1160-1161)."""
        import json
        import os
        trunk = os.path.join(path, file)
        props = np.loadtxt(trunk + '-svm.csv', delimiter=',').ravel()
        nsv, ndof = int(round(props[0])), int(round(props[1]))
        if ndof not in (2, 6, 15):
            raise NotImplementedError('from_MLparam: supported feature vectors are Ndof = 6 (sdim = 6), 15 (sdim = 6 with '
                                      'work-hardening features) and 2 (sdim = 3); texture features are not; file has Ndof = %d' % ndof)
        sdim = 3 if ndof == 2 else 6
        if nsv < 1 or len(props) < 29 + nsv * (ndof + 1):
            raise ValueError('from_MLparam: file %s-svm.csv is truncated' % trunk)
        meta = None
        if os.path.exists(trunk + '-svm_meta.json'):
            with open(trunk + '-svm_meta.json') as fp:
                meta = json.load(fp)
        m = cls(name=name or (os.path.basename(file)))
        if props[10] > 0.:
            CV = np.zeros((6, 6))
            CV[0, 0], CV[1, 1], CV[2, 2] = props[2], props[10], props[11]
            CV[0, 1] = CV[1, 0] = props[3]
            CV[0, 2] = CV[2, 0] = props[12]
            CV[1, 2] = CV[2, 1] = props[13]
            CV[3, 3], CV[4, 4], CV[5, 5] = props[4], props[14], props[15]
            iso = (props[10] == props[2] and props[11] == props[2] and props[12] == props[3] and props[13] == props[3]
                   and props[14] == props[4] and props[15] == props[4])
            if iso:
                m.elasticity(C11=props[2], C12=props[3], C44=props[4])
            else:
                m.elasticity(CV=CV)
        else:
            m.elasticity(C11=props[2], C12=props[3], C44=props[4])
        scale_seq = float(props[8])
        m.plasticity(sy=float(scale_seq if sy is None else sy), sdim=sdim)
        dual = props[29:29 + nsv]
        sv = props[29 + nsv:29 + nsv + nsv * ndof].reshape(nsv, ndof)
        m.set_svc(sv, dual, float(props[5]), float(props[6]), scale_seq=scale_seq, dev_only=(props[16] < 0.),
                  scale_wh=float(props[9]) if ndof == 15 else None)
        m.epc, m.scale_wh = float(props[7]), float(props[9])
        m.meta = meta
        return m

    def from_sklearn(self, svc, scale_seq=None, dev_only=False, scale_wh=None):
        """Adopt a fitted sklearn.svm.SVC(kernel='rbf') (binary) as yield function."""
        self.svm_yf = svc
        self.set_svc(svc.support_vectors_, svc.dual_coef_[0], svc.intercept_[0], svc._gamma if hasattr(svc, '_gamma') else svc.gamma,
                     scale_seq=scale_seq, dev_only=dev_only, scale_wh=scale_wh)

    @classmethod
    def from_reference(cls, ref):
        """Copy the parameters of a pyLabFEA Material object (duck typed): elastic constants, plasticity and, when
        ref.ML_yf is set, the trained SVC.  Lets a material trained with the reference's train_SVC drive this path."""
        m = cls(name=ref.name, num=getattr(ref, 'num', 1))
        m.elasticity(CV=np.array(ref.CV)) if ref.E is None else None
        if ref.E is not None:
            m.E, m.nu, m.C11, m.C12, m.C44 = ref.E, ref.nu, ref.C11, ref.C12, ref.C44
            m.CV = np.array(ref.CV, dtype=float)
        if ref.sy is not None:
            if ref.sdim not in (3, 6):
                raise ValueError('sdim must be either 3 or 6')
            m.sy0, m.sy, m.khard, m.sdim = ref.sy0, ref.sy, ref.khard, int(ref.sdim)
            m.drucker = ref.drucker
            m.lhs = None if ref.lhs is None else np.array(ref.lhs, dtype=float)
            m.hill = np.array(ref.hill, dtype=float)
            m.hill_6p, m.hill_3p, m.tresca = bool(ref.hill_6p), bool(ref.hill_3p), bool(ref.tresca)
            if getattr(ref, 'barlat', False):
                m.barlat, m.Bar_m1, m.Bar_m2, m.barlat_exp = True, np.array(ref.Bar_m1), np.array(ref.Bar_m2), ref.barlat_exp
            if getattr(ref, 'ML_yf', False):
                if getattr(ref, 'txdat', False):
                    # the reference's own texture path divides by self.self.scale_seq (material.py:811) and cannot run
                    raise NotImplementedError('SVC features for texture are not supported on the device path')
                m.from_sklearn(ref.svm_yf, scale_seq=ref.scale_seq, dev_only=ref.dev_only,
                               scale_wh=ref.scale_wh if getattr(ref, 'whdat', False) else None)
        return m

    # ------------------------------------------------------------------------------------------------------
    # descriptor crossing the C-ABI
    # ------------------------------------------------------------------------------------------------------
    def _descriptor(self, CV=None):
        """struct plf_material for this material; CV overrides Material.CV (Material.response takes Cel as argument)."""
        if self.CV is None:
            raise ValueError('Material %s: elasticity() must be called first' % self.name)
        d = _lib.PlfMaterial()
        d.E, d.nu, d.C44 = float(self.E), float(self.nu), float(self.C44)
        M = np.ascontiguousarray(self.CV if CV is None else CV, dtype=float)
        if M.shape != (6, 6):
            raise ValueError('Voigt stiffness must have shape (6,6)')
        d.CV[:] = list(M.ravel())
        keep = []
        if self.sy is None:
            d.kind = _lib.MAT_ELASTIC
            d.hill[:] = [1.] * 6
        else:
            if self.sdim != 6 and self.barlat:
                raise NotImplementedError('Material %s: the Barlat yield function needs sdim = 6' % self.name)
            d.sy, d.khard = float(self.sy), float(self.khard)
            d.hill[:] = ([float(x) for x in self.hill] + [1., 1., 1.])[:6]
            d.tresca = int(bool(self.tresca))
            if self.lhs is not None:
                d.lhs = 1
                d.d[:] = [float(x) for x in self.lhs]
            else:
                d.d[:] = [float(self.drucker)] * 3
            if self.ML_yf:
                sv, dc, rho = self._svc
                d.kind = _lib.MAT_SVC if self.sdim == 6 else _lib.MAT_SVC3
                d.nsv = len(dc)
                d.svc_sv = sv.ctypes.data_as(_lib._dp)
                d.svc_dual = dc.ctypes.data_as(_lib._dp)
                d.svc_gamma, d.svc_rho, d.svc_scale_seq = self.gam_yf, rho, self.scale_seq
                d.dev_only = int(self.dev_only)
                if getattr(self, 'whdat', False):
                    d.svc_nfeat, d.svc_scale_wh = 15, float(self.scale_wh)   # khard: the value the material carries now
                else:
                    d.khard = 0.
                keep = [sv, dc]
            elif self.barlat:
                d.kind = _lib.MAT_BARLAT
                d.barlat_m1[:] = list(np.asarray(self.Bar_m1, float).ravel())
                d.barlat_m2[:] = list(np.asarray(self.Bar_m2, float).ravel())
                d.barlat_exp = float(self.barlat_exp)
            else:
                # Tresca: d.tresca switches calc_seq to max - min principal stress.  sdim = 3: equivalent stress and
                # gradient on the ordered principal stresses (material.py:663-673, 1045-1048)
                d.kind = _lib.MAT_HILL if self.sdim == 6 else _lib.MAT_HILL3
        d._keep = keep
        return d

    def _bind(self, CV=None):
        ctx = default_context()
        ctx.set_materials([self._descriptor(CV)], planestress=False)
        return ctx

    def _as_rows(self, sig, what):
        a = np.asarray(sig, dtype=float)
        if a.ndim == 1 and a.shape[0] == 6:
            return a[None, :], True
        if a.ndim == 2 and a.shape[1] == 6:
            return a, False
        if self.sdim == 3 and a.shape[-1] == 3 and a.ndim in (1, 2):
            # principal stresses (material.py:609-614): the Voigt vector [sp, 0, 0, 0]
            rows = np.atleast_2d(a)
            return np.ascontiguousarray(np.concatenate([rows, np.zeros_like(rows)], axis=1)), a.ndim == 1
        raise TypeError('Unknown format of stress in %s (device path needs (6,) or (N,6) Voigt stresses)' % what)

    def _epl_rows(self, epl, n):
        if epl is None:
            return np.zeros((n, 6))
        if isinstance(epl, (float, np.floating)):
            # calc_yf: a PEEQ is converted into a plastic strain tensor (material.py:381-383)
            return np.broadcast_to(float(epl) * np.array([1., -0.5, -0.5, 0., 0., 0.]), (n, 6))
        e = np.asarray(epl, dtype=float)
        if e.shape[-1] == 3:   # sdim = 3: principal plastic strains
            e = np.concatenate([np.atleast_2d(e), np.zeros_like(np.atleast_2d(e))], axis=1)
        return np.broadcast_to(e, (n, 6))

    # ------------------------------------------------------------------------------------------------------
    # constitutive evaluations (device)
    # ------------------------------------------------------------------------------------------------------
    def calc_seq(self, sig):
        """Equivalent stress, (6,) -> float or (N,6) -> (N,) (material.py:576-676; Barlat 678-702)."""
        a, single = self._as_rows(sig, 'calc_seq')
        out = self._bind().calc_seq(0, a)
        self.msg['equiv'] = ('Barlat Yld2004-18p' if self.barlat else
                             '6-parameter Hill, full Voigt stress' if self.sdim == 6 else '3-parameter Hill')
        return float(out[0]) if single else out

    def get_sflow(self, epl):
        """Flow stress sy + khard * PEEQ (material.py:974-1007)."""
        peeq = epl if isinstance(epl, (float, np.floating)) else eps_eq(epl)
        return self.sy + peeq * self.khard

    def calc_yf(self, sig, epl=None, ana=False, pred=False, **_ignored):
        """Yield function (material.py:348-412): SVC decision function for ML materials (sign for pred=True),
        seq - sflow otherwise."""
        a, single = self._as_rows(sig, 'calc_yf')
        if self.ML_yf and ana:
            raise NotImplementedError('calc_yf(ana=True) for an ML material: use the analytic reference material instead')
        out = self._bind().calc_yf(0, a, self._epl_rows(epl, len(a)))
        if self.ML_yf:
            if pred:
                out = np.where(out >= 0., 1., -1.)
            self.msg['yield_fct'] = 'ML_yf-predict' if pred else 'ML_yf-decision-fct'
        else:
            self.msg['yield_fct'] = 'analytical'
        return float(out[0]) if single else out

    def ML_full_yf(self, sig, epl=None, ld=None, verb=True, **_ignored):
        """Distance of a stress to the SVC yield locus along the loading direction (material.py:414-516)."""
        a = np.asarray(sig, dtype=float)
        if a.shape != (6,):
            raise ValueError('Only individual stress tensors supported in material.ML_full_yf. Shape of argument is {}'
                             .format(a.shape))
        out = self._bind().ml_full_yf(0, a[None, :], self._epl_rows(epl, 1), ld)
        return float(out[0])

    def calc_fgrad(self, sig, epl=None, seq=None, ana=False, **_ignored):
        """Gradient of the yield function w.r.t. stress (material.py:704-858).  Raises ValueError for Tresca like the
        reference (824-825).  Deviation: the reference also raises for Barlat (822-823) and therefore cannot run a
        Barlat return mapping at all; here the Yld2004-18p gradient is implemented analytically on the device
        (barlat_seq_grad, chain rule through the two eigen-decompositions), which is what BASELINE config 5 needs."""
        a, single = self._as_rows(sig, 'calc_fgrad')
        if epl is not None and np.shape(epl) != np.shape(sig):
            raise ValueError('Parameter sig and epl must have the same shape.')
        if self.sdim == 3 and np.shape(sig)[-1] != 3:
            raise NotImplementedError('calc_fgrad of an sdim = 3 material takes principal stresses, shape (3,) or (N,3) '
                                      '(what epl_dot / C_tan pass, material.py:1045-1048)')
        out = self._bind().calc_fgrad(0, a, self._epl_rows(epl, len(a)))
        if self.sdim == 3:
            out = out[:, 0:3]
        if self.ML_yf:
            self.khard = self._svc_khard(a, self._epl_rows(epl, len(a))) if getattr(self, 'whdat', False) else 0.
            self.msg['gradient'] = 'gradient to ML_yf'
        elif self.barlat:
            self.msg['gradient'] = 'analytical, Barlat Yld2004-18p, full stress'
        else:
            self.msg['gradient'] = 'analytical, 6-parameter Hill, full stress'
        return out[0] if single else out

    def _svc_khard(self, sig, epl):
        """Material.khard as calc_fgrad leaves it for a work-hardening SVC (material.py:802-814): minus the plastic-strain
        part of the decision function's gradient, scaled by scale_seq / scale_wh, summed over components and averaged over
        the points; never negative.  Host-side NumPy on the support-vector table (the device evaluates the same per element
        inside the return mapping)."""
        sv, dc, _ = self._svc
        x = np.zeros((len(sig), 15))
        s = np.array(sig, dtype=float)
        if self.dev_only:
            s[:, 0:3] -= s[:, 0:3].mean(axis=1, keepdims=True)
        x[:, 0:6] = s / self.scale_seq
        x[:, 6:12] = np.asarray(epl, dtype=float) / self.scale_wh
        hk = 0.
        for xi in x:
            hv = xi[None, :] - sv
            k = np.exp(-self.gam_yf * np.sum(hv * hv, axis=1))
            dKdx = np.sum((dc * k)[:, None] * (-2. * self.gam_yf * hv), axis=0)
            hk -= np.sum(dKdx[6:12]) * self.scale_seq / self.scale_wh
        return max(hk / len(x), 0.)

    def epl_dot(self, sig, epl, Cel, deps, **_ignored):
        """Plastic strain increment (material.py:1009-1055, Crisfield 6.4 / 6.8 / 6.17)."""
        out = self._bind(Cel).epl_dot(0, np.asarray(sig, float)[None, :], np.asarray(epl, float)[None, :],
                                      np.asarray(deps, float)[None, :])
        return out[0]

    def C_tan(self, sig, Cel, epl=None):
        """Tangent stiffness (material.py:1057-1086, Crisfield 6.9 / 6.18)."""
        e = None if epl is None else np.asarray(epl, float)[None, :]
        return self._bind(Cel).C_tan(0, np.asarray(sig, float)[None, :], e)[0]

    def response(self, sig, epl, deps, CV, maxit=50):
        """Material.response (material.py:207-346) for one element: returns (fy1, sig, depl, grad_stiff) and sets
        msg['nsteps'].  For many elements use response_batch."""
        s = np.asarray(sig, dtype=float)
        if s.shape != (6,):
            raise ValueError('Only individual stress tensors supported in material.response. Shape of argument is {}'
                             .format(s.shape))
        if maxit != 50:
            raise NotImplementedError('maxit is fixed to 50 on the device path (PLF_MAXIT)')
        fy, so, dp, ct, ns = self.response_batch(s[None, :], np.asarray(epl, float)[None, :], np.asarray(deps, float)[None, :], CV)
        self.msg['nsteps'] = int(ns[0])
        return float(fy[0]), so[0], dp[0], ct[0]

    def response_batch(self, sig, epl, deps, CV=None):
        """Batched Material.response: (N,6) arrays -> (fy (N,), sig (N,6), depl (N,6), grad_stiff (N,6,6), nsteps (N,))."""
        if self.sy is None:
            raise ValueError('response: material %s has no plasticity parameters' % self.name)
        return self._bind(CV).response_batch(sig, epl, deps)

    # ------------------------------------------------------------------------------------------------------
    # derived properties: the harness behind most golden values of the reference tests
    # ------------------------------------------------------------------------------------------------------
    def calc_properties(self, size=2, Nel=2, verb=False, eps=0.005, min_step=None, sigeps=False,
                        load_cases=('stx', 'sty', 'et2', 'ect')):
        """Stress-strain curves under uniaxial x / y tension, equibiaxial strain and pure shear on an Nel x Nel
        plane-stress model (material.py:3062-3166).  Fills prop / propJ2 / sigeps like the reference."""
        from .basic import sig_eq_j2
        from .model import Model

        u = eps * size
        # boundary values (x: right edge, y: top edge), style and legend name per load case (3125-3153)
        cases = {'stx': ((u, 'disp'), (0., 'force'), '-r', 'uniax-x'),
                 'sty': ((0., 'force'), (u, 'disp'), '-b', 'uniax-y'),
                 'et2': ((0.4 * u, 'disp'), (0.4 * u, 'disp'), '-k', 'equibiax'),
                 'ect': ((-0.8 * u, 'disp'), (0.8 * u, 'disp'), '-m', 'shear')}
        for case in load_cases:
            if case not in cases:
                warnings.warn('calc_properties: Load case not supported: {}'.format(case))
                continue
            (v1, t1), (v2, t2), style, label = cases[case]
            fe = Model(dim=2, planestress=True)
            fe.geom([size], LY=size)
            fe.assign([self])
            fe.bcleft(0.)
            fe.bcbot(0.)
            fe.bcright(v1, t1)
            fe.bctop(v2, t2)
            fe.mesh(NX=Nel, NY=Nel)
            fe.solve(verb=verb, min_step=min_step)
            eeq, peeq = eps_eq(fe.egl), eps_eq(fe.epgl)
            seq = self.calc_seq(fe.sgl)
            # yield strength = equivalent stress at the last history point below the plastic-strain threshold
            # (1e-2 for prop, 1e-6 for propJ2: 3104-3114)
            self.prop[case].update(ys=seq[np.nonzero(peeq < 1.e-2)[0][-1]], seq=seq, eeq=eeq, peeq=peeq,
                                   style=style, name=label)
            seqJ2 = sig_eq_j2(fe.sgl)
            self.propJ2[case].update(ys=seqJ2[np.nonzero(peeq < 1.e-6)[0][-1]], seq=seqJ2, eeq=eeq, peeq=peeq)
            if sigeps:
                self.sigeps[case].update(sig=fe.sgl, eps=fe.egl, epl=fe.epgl)
