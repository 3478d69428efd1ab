"""Shared helpers for the parity tests: golden dict -> plf_material descriptor."""
import numpy as np

from pylabfea_b200 import _lib


def _keep(m, *arrays):
    m._arrays = arrays
    return m


def plf_material_from_golden(g):
    m = _lib.PlfMaterial()
    m.E, m.nu = float(g['E']), float(g['nu'])
    CV = np.ascontiguousarray(g['CV'] if 'CV' in g else g['mat_CV'], dtype=float)
    m.CV[:] = list(CV.ravel())
    m.C44 = float(CV[3, 3])
    if 'sv' in g:
        sv = np.asarray(g['sv'], dtype=float)
        m.kind = _lib.MAT_SVC
        if 'sdim' in g and int(g['sdim']) == 3:   # sdim = 3 model: 2 features, table padded to 6 columns
            m.kind = _lib.MAT_SVC3
            sv = np.hstack([sv, np.zeros((len(sv), 6 - sv.shape[1]))])
        sv = np.ascontiguousarray(sv)
        dual = np.ascontiguousarray(g['dual'], dtype=float)
        m.nsv = len(dual)
        m.svc_sv = sv.ctypes.data_as(_lib._dp)
        m.svc_dual = dual.ctypes.data_as(_lib._dp)
        m.svc_gamma, m.svc_rho, m.svc_scale_seq = float(g['gamma']), float(g['rho']), float(g['scale_seq'])
        m.dev_only = int(g['dev_only'])
        m.sy, m.khard = float(g['sy']), 0.
        m.hill[:] = (list(np.asarray(g['hill'], float)) + [1., 1., 1.])[:6]   # sdim = 3 fixtures carry 3 parameters
        m.d[:] = [float(g['drucker'])] * 3
        return _keep(m, sv, dual)
    if 'sy' in g:
        m.kind = _lib.MAT_HILL
        m.sy, m.khard = float(g['sy']), float(g['khard'])
        m.hill[:] = (list(np.asarray(g['hill'], float)) + [1., 1., 1.])[:6]   # sdim = 3 fixtures carry 3 parameters
        m.d[:] = [float(g['drucker'])] * 3
    else:
        m.kind = _lib.MAT_ELASTIC
        m.hill[:] = [1.] * 6
    return m


def elastic_material(E, nu, CV):
    m = _lib.PlfMaterial()
    m.kind = _lib.MAT_ELASTIC
    m.E, m.nu = float(E), float(nu)
    CV = np.asarray(CV, float)
    m.CV[:] = list(CV.ravel())
    m.C44 = float(CV[3, 3])
    m.hill[:] = [1.] * 6
    return m


def relmax(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)
