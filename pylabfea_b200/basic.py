"""Voigt-tensor helpers of the hot path (host side, numpy).  Same names and conventions as pyLabFEA's basic.py
(Voigt order 11,22,33,23,13,12, engineering shear): yf_tolerance (basic.py:26), sig_dev (304-325), eps_eq (328-360),
sig_eq_j2 (30-65).  Only what the stress-update / assembly path and its callers need; coordinate conversions for
plots and training are out of scope."""
import numpy as np

yf_tolerance = 5.e-3
"""Plastic yielding if yield function > yf_tolerance (relative to the flow stress in Material.response)"""


def _rows(x, what):
    a = np.asarray(x, dtype=float)
    if a.ndim == 1 and a.shape[0] in (3, 6):
        return a[None, :], True
    if a.ndim == 2 and a.shape[1] in (3, 6):
        return a, False
    raise TypeError('%s: format not supported, shape %s' % (what, a.shape))


def sig_dev(sig):
    """Deviatoric part of (3,), (6,), (N,3) or (N,6) stresses."""
    a, single = _rows(sig, 'sig_dev')
    out = a.copy()
    out[:, 0:3] -= a[:, 0:3].sum(axis=1, keepdims=True) / 3.
    return out[0] if single else out


def eps_eq(eps):
    """Equivalent strain sqrt(2/3 (sum e_ii^2 + 1/2 sum gamma_ij^2)) of principal (3) or Voigt (6) strains."""
    a, single = _rows(eps, 'eps_eq')
    s = np.sum(a[:, 0:3] ** 2, axis=1)
    if a.shape[1] == 6:
        s = s + 0.5 * np.sum(a[:, 3:6] ** 2, axis=1)
    out = np.sqrt(2. * s / 3.)
    return out[0] if single else out


def sig_eq_j2(sig):
    """J2 equivalent stress of principal (3) or Voigt (6) stresses."""
    a, single = _rows(sig, 'sig_eq_j2')
    d = (a[:, 0] - a[:, 1]) ** 2 + (a[:, 1] - a[:, 2]) ** 2 + (a[:, 2] - a[:, 0]) ** 2
    if a.shape[1] == 6:
        d = d + 6. * np.sum(a[:, 3:6] ** 2, axis=1)
    out = np.sqrt(0.5 * d)
    return out[0] if single else out
