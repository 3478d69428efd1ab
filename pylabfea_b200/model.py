"""Model -- host-side mirror of pyLabFEA's Model class (model.py) for 2-d structured quadrilateral meshes.

Keeps the reference's API (geom / assign / bcleft / bcright / bcbot / bctop / bcnode / mesh / setupK / solve /
calc_global / bcval and the attributes notebooks read: u, f, glob, sgl, egl, epgl, element[i].sig/.eps/.epl/...,
Nel, Ndof, npos, noleft/noright/nobot/notop, nsteps, niter, co_nconv), but every O(Nel) / O(Ndof) operation of
Model.solve runs on the device through the C-ABI: K integration + CSR assembly, BC-reduced CG solve, the
Material.response sweep with tangent replacement, calc_scf, the state commit and the volume averages.  Only
scalars, the boundary-condition vectors and (once per increment) u and f cross PCIe.

File:line citations refer to /root/reference/src/pylabfea/model.py.
"""
import warnings

import numpy as np

from . import _lib
from .basic import yf_tolerance


class StiffnessCSR(object):
    """What Model.setupK returns: the assembled stiffness in CSR form.  np.asarray(K) / K.toarray() densify it
    (the reference returns the dense matrix, model.py:963-977), K.tocsr() gives a scipy.sparse matrix."""

    def __init__(self, rowptr, col, val, n):
        self.indptr, self.indices, self.data, self.shape = rowptr, col, val, (n, n)

    def toarray(self):
        n = self.shape[0]
        if n > 20000:
            raise MemoryError('dense K with %d dofs requested' % n)
        K = np.zeros(self.shape)
        rows = np.repeat(np.arange(n), np.diff(self.indptr))
        K[rows, self.indices] = self.data
        return K

    def __array__(self, dtype=None, copy=None):
        return self.toarray()

    def tocsr(self):
        from scipy.sparse import csr_matrix
        return csr_matrix((self.data, self.indices, self.indptr), shape=self.shape)

    def __matmul__(self, x):
        x = np.asarray(x, dtype=float)
        prod = self.data * x[self.indices]
        return np.add.reduceat(prod, self.indptr[:-1])


class _ElementView(object):
    """Element i of the model: read access to what Model.Element exposes (model.py:202-260)."""

    def __init__(self, model, i):
        self._m, self._i = model, i

    @property
    def Model(self):
        return self._m

    @property
    def nodes(self):
        return [int(n) for n in self._m._nodes[self._i]]

    @property
    def Mat(self):
        return self._m.mat[self._m._elem_mat[self._i]]

    @property
    def Lelx(self):
        return float(self._m._dx[self._i // self._m._NY])

    @property
    def Lely(self):
        return float(self._m._dy)

    @property
    def Vel(self):
        return self.Lelx * self.Lely * self._m.thick

    @property
    def CV(self):
        return self._m._dev().element_CV(self._m._elem_mat[self._i])

    @property
    def sig(self):
        return self._m._state('sig')[self._i]

    @property
    def eps(self):
        return self._m._state('eps')[self._i]

    @property
    def epl(self):
        return self._m._state('epl')[self._i]

    @property
    def elstiff(self):
        return self._m._state('elstiff')[self._i]

    @property
    def Kel(self):
        return self._m._state('Kel')[self._i]

    @property
    def Bmat(self):
        return list(self._m._dev().get_bmat(self._i))

    def node_num(self):
        out = []
        for n in self.nodes:
            out += [2 * n, 2 * n + 1]
        return out


class _ElementList(object):
    def __init__(self, model):
        self._m = model

    def __len__(self):
        return self._m.Nel

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        i = int(i)
        if i < 0:
            i += len(self)
        if not 0 <= i < len(self):
            raise IndexError(i)
        return _ElementView(self._m, i)

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class Model(object):
    """2-d FE model on a structured Q4 mesh (reference: model.py:24-200 for parameters and attributes)."""

    def __init__(self, dim=1, planestress=False, device=0):
        if dim not in (1, 2):
            raise ValueError('dim must be either 1 or 2')
        if planestress and dim != 2:
            warnings.warn('Warning: Plane stress only defined for 2-d model')
            planestress = False
        self.dim = dim
        self.planestress = bool(planestress)
        self._device = device
        self.bcl, self.bcb, self.bct, self.bcr, self.bcn = (np.zeros(dim) for _ in range(5))
        self.noset = None
        self.ubctop = [False, False]
        self.ubcright = [False, False]
        self.ubcleft = [True, False]
        self.ubcbot = [False, True]
        self.ubcn = [False, False]
        self.nonlin = False
        self.sgl = np.zeros((1, 6))
        self.egl = np.zeros((1, 6))
        self.epgl = np.zeros((1, 6))
        self.u = None
        self.f = None
        self.Nnode = None
        self.glob = {'ebc1': None, 'ebc2': None, 'sbc1': None, 'sbc2': None,
                     'eps': np.zeros(6), 'sig': np.zeros(6), 'epl': np.zeros(6)}
        self._ctx = None
        self._cache = {}
        self.solver_rtol = 1.e-13     # CG stops at |r| <= rtol |b|
        self.solver_maxit = 200000
        self.solver_stats = []        # (iterations, relative residual) of every linear solve
        self._bc_key = None           # boundary values whose BC vectors are on the device (re-used by the next pass)
        self._bc_epoch = 0

    # ------------------------------------------------------------------------------------------------------
    # pre-processing: geometry, materials, boundary conditions (host, O(1))
    # ------------------------------------------------------------------------------------------------------
    def geom(self, sect=1, LX=None, LY=1., LZ=1.):
        """Model dimensions and sections (model.py:514-552)."""
        if type(sect) == list:
            self.Nsec = len(sect)
            self.LS = np.array(sect)
            self.lenx = sum(sect)
        elif type(sect) == int:
            if sect < 1:
                raise ValueError('At least one section must be defined.')
            if LX is None:
                raise ValueError('LX must be given if sect is of type int')
            self.lenx = LX
            self.Nsec = sect
            self.LS = np.ones(sect) * self.lenx / sect
        else:
            raise TypeError('Sect must be either list or int, not {}'.format(type(sect)))
        self.leny = LY
        self.thick = LZ

    def assign(self, mats):
        """One Material per section (model.py:554-577)."""
        if len(mats) != self.Nsec:
            raise ValueError('Numer of materials ({}) does not match number of sections ({})'.format(len(mats), self.Nsec))
        self.mat = mats
        self.nonlin = any(m.sy is not None for m in mats)

    def _set_bc(self, who, store, flags, val, bctype, bcdir, zero_force_only=False, type_error=False):
        if bcdir in (0, 1):
            j = int(bcdir)
        elif isinstance(bcdir, str) and bcdir.lower() in ('x', 'y'):
            j = 0 if bcdir.lower() == 'x' else 1
        else:
            raise ValueError('%s: Unknown value for direction: %s' % (who, bcdir))
        store[j] = val
        t = bctype.lower()
        if t == 'disp':
            flags[j] = True
        elif t == 'force':
            flags[j] = False
            if zero_force_only and np.abs(val) > 1.e-6:
                raise ValueError('Finite force values at %s boundary not supported.' % who[2:])
        elif type_error:
            raise TypeError('%s: Unknown BC: %s' % (who, bctype))
        else:
            raise ValueError('%s: Unknown BC: %s' % (who, bctype))

    def bcleft(self, val=0., bctype='disp', bcdir='x'):
        """Static BC on the left edge (model.py:580-612)."""
        self._set_bc('bcleft', self.bcl, self.ubcleft, val, bctype, bcdir, zero_force_only=True)

    def bcright(self, val, bctype, bcdir='x'):
        """Incremented BC on the right edge (model.py:614-644)."""
        self._set_bc('bcright', self.bcr, self.ubcright, val, bctype, bcdir, type_error=True)

    def bcbot(self, val=0., bctype='disp', bcdir='y'):
        """Static BC on the bottom edge (model.py:646-680)."""
        self._set_bc('bcbot', self.bcb, self.ubcbot, val, bctype, bcdir, zero_force_only=True)

    def bctop(self, val, bctype, bcdir='y'):
        """Incremented BC on the top edge (model.py:682-713)."""
        self._set_bc('bctop', self.bct, self.ubctop, val, bctype, bcdir, type_error=True)

    def bcnode(self, node, val, bctype, bcdir):
        """Incremented BC on a node set; only after meshing (model.py:715-756)."""
        if isinstance(node, (list, tuple, np.ndarray)):
            self.noset = [int(n) for n in np.ravel(node)]
        else:
            self.noset = [int(node)]
        self._set_bc('bcnode', self.bcn, self.ubcn, val, bctype, bcdir, type_error=True)

    # ------------------------------------------------------------------------------------------------------
    # mesh (model.py:758-952), vectorised; the device receives NX, NY, dx[NX], dy and the material map
    # ------------------------------------------------------------------------------------------------------
    def mesh(self, elmts=None, nodes=None, NX=10, NY=1, SF=1):
        if self.dim != 2:
            raise NotImplementedError('only 2-d models are supported on the device path')
        if SF != 1:
            raise NotImplementedError('Error: Quadrilateral elements with quadratic shape function not yet implemented')
        if nodes is not None:
            raise NotImplementedError('user-defined node positions are not supported on the device path')
        self.shapefact = SF
        if elmts is not None:
            el = np.array(elmts, dtype=int)
            if el.ndim != 2:
                raise ValueError('Cannot use a {}-shaped mesh with a {}-dimemsional model'.format(el.shape, self.dim))
            NX, NY = el.shape
        if NX < self.Nsec:
            raise TypeError('Error: Number of elements is smaller than number of sections')
        if self.u is not None:
            warnings.warn('Warning: Solution of previous steps is deleted')
            self.u = None
            self.f = None
        self.NnodeX, self.NnodeY = NX + 1, NY + 1
        self.Nnode = self.NnodeX * self.NnodeY
        self.Ndof = 2 * self.Nnode
        self.Nel = NX * NY
        self._NX, self._NY = NX, NY
        dy = self.leny / NY
        if elmts is None:
            # laminate: elements per section in proportion to the section lengths (826-830)
            nes = [int(x) for x in np.round(self.LS / self.lenx * NX)]
            if sum(nes) != NX:
                im = int(np.argmax(self.LS))
                nes[im] = nes[im] - sum(nes) + NX
            dx = np.concatenate([np.full(n, self.LS[i] / n) for i, n in enumerate(nes) if n > 0])
            col_mat = np.concatenate([np.full(n, i, dtype=np.int32) for i, n in enumerate(nes)])
            elem_mat = np.repeat(col_mat, NY)
            # node x-positions: (j + section start) * dx of the section (847); identical to a running sum only when
            # all dx are equal, so restate the reference's formula
            xs = np.zeros(self.NnodeX)
            start = 0
            for i, n in enumerate(nes):
                if n == 0:
                    continue
                d = self.LS[i] / n
                j0 = 0 if i == 0 else 1
                for j in range(j0, n + 1):
                    xs[start + j] = (j + start) * d
                start += n
        else:
            dx = np.full(NX, self.lenx / NX)
            elem_mat = (el - 1).astype(np.int32).ravel()
            xs = np.arange(self.NnodeX) * (self.lenx / NX)
        if elem_mat.min() < 0 or elem_mat.max() >= len(self.mat):
            raise IndexError('mesh: material number out of range')
        ys = np.arange(self.NnodeY) * dy
        self.npos = np.zeros(self.Ndof)
        self.npos[0::2] = np.repeat(xs, self.NnodeY)
        self.npos[1::2] = np.tile(ys, self.NnodeX)
        nid = np.arange(self.Nnode).reshape(self.NnodeX, self.NnodeY)
        self.noleft = [int(n) for n in nid[0, :]]
        self.noright = [int(n) for n in nid[-1, :]]
        self.nobot = [int(n) for n in nid[:, 0]]
        self.notop = [int(n) for n in nid[:, -1]]
        self.noinner = [int(n) for n in nid[1:-1, 1:-1].ravel()]
        j, k = np.divmod(np.arange(self.Nel), NY)
        n1 = j * self.NnodeY + k
        self._nodes = np.stack([n1, n1 + 1, n1 + self.NnodeY, n1 + self.NnodeY + 1], axis=1).astype(np.int32)
        self._dx, self._dy, self._elem_mat = dx, dy, elem_mat
        self._plastic = np.array([m.sy is not None for m in self.mat])
        self.element = _ElementList(self)
        self._cache = {}
        self._on_device = False

    def _dev(self):
        """The device context with materials and mesh uploaded (created on first use: no CPU fallback exists, so this
        raises PlfError without libplfb200.so or a CUDA device)."""
        if not self._on_device:
            if self._ctx is None:
                self._ctx = _lib.Context(self._device)
            self._ctx.set_materials([m._descriptor() for m in self.mat], planestress=self.planestress)
            self._ctx.mesh_structured(self._NX, self._NY, self._dx, self._dy, self.thick, self._elem_mat)
            self._on_device = True
        return self._ctx

    # ------------------------------------------------------------------------------------------------------
    # lazy host mirrors of the device state
    # ------------------------------------------------------------------------------------------------------
    def _state(self, name):
        if name not in self._cache:
            c = self._dev()
            if name == 'sig':
                a = c.get_field(_lib.F_SIG)
            elif name == 'eps':
                a = c.get_field(_lib.F_EPS)
            elif name == 'epl':
                a = c.get_field(_lib.F_EPL)
            elif name == 'elstiff':
                a = c.get_tangent()
            elif name == 'Kel':
                a = c.get_kel()
            else:
                raise KeyError(name)
            self._cache[name] = a
        return self._cache[name]

    # ------------------------------------------------------------------------------------------------------
    def setupK(self):
        """Integrate B^T C B per element and assemble the system stiffness on the device (model.py:954-977)."""
        if self.Nnode is None:
            raise AttributeError('Attributes for mesh not set, but required by solver.')
        ctx = self._dev()
        ctx.setupK()
        rowptr, col, val = ctx.get_csr()
        return StiffnessCSR(rowptr, col, val, self.Ndof)

    # ------------------------------------------------------------------------------------------------------
    # boundary conditions -> vectors for the reduced solve (model.py:1070-1206)
    # ------------------------------------------------------------------------------------------------------
    def _calc_BC(self, bcl0, bcb0, dbcr, dbct, dbcn):
        """Restates calc_BC: returns fixed (dofs removed from `ind`), du (prescribed increments), w (sum of every BC
        value applied to a dof: what the reference multiplies K's columns with, including its double application at
        shared corner dofs) and f_ext (force-type BC).  df[ind] = f_ext[ind] - (K w)[ind]."""
        n = self.Ndof
        fixed = np.zeros(n, dtype=np.uint8)
        du = np.zeros(n)
        w = np.zeros(n)
        fext = np.zeros(n)

        def disp(nodes, k, val, who, first=False):
            for j in nodes:
                i = 2 * j + k
                if not fixed[i]:
                    fixed[i] = 1
                    du[i] = val
                elif first:
                    raise ValueError('list.remove(x): x not in list')   # what the reference's ind.remove does
                elif du[i] != val:
                    warnings.warn('Inconsistent BC at %s node %d (%g vs %g).' % (who, j, du[i], val))
                w[i] += val

        for k in range(2):                       # left: static (1099-1105)
            if self.ubcleft[k]:
                disp(self.noleft, k, bcl0[k], 'left', first=True)
        for k in range(2):                       # bottom: static (1111-1122)
            if self.ubcbot[k]:
                disp(self.nobot, k, bcb0[k], 'bottom')
        for k in range(2):                       # right (1126-1151)
            if self.ubcright[k]:
                disp(self.noright, k, dbcr[k], 'right')
            else:
                share = 1. / (self.NnodeY - 1)
                for j in self.noright:
                    hy = self.npos[2 * j + 1]
                    fext[2 * j + k] += dbcr[k] * (share * 0.5 if (hy < 1.e-3 or hy > self.leny - 1.e-3) else share)
        for k in range(2):                       # top (1156-1179)
            if self.ubctop[k]:
                disp(self.notop, k, dbct[k], 'top')
            else:
                share = 1. / (self.NnodeX - 1)
                for j in self.notop:
                    hx = self.npos[2 * j]
                    fext[2 * j + k] += dbct[k] * (share * 0.5 if (hx < 1.e-3 or hx > self.lenx - 1.e-3) else share)
        if self.noset is not None:               # node set (1183-1205)
            if dbcn is None:
                raise ValueError('No BC for selected node set given.')
            for k in range(2):
                if self.ubcn[k]:
                    disp(self.noset, k, dbcn[k], 'node set')
                else:
                    for j in self.noset:
                        fext[2 * j + k] += dbcn[k]
        return fixed, du, w, fext

    def _solve_increment(self, bcl0, bcb0, dbcr, dbct, dbcn):
        # the inner passes of an increment usually repeat the boundary values of the previous pass: the BC vectors are
        # then neither rebuilt on the host nor copied to the device again
        key = (repr(np.asarray(bcl0).tolist()), repr(np.asarray(bcb0).tolist()), repr(np.asarray(dbcr).tolist()),
               repr(np.asarray(dbct).tolist()), None if dbcn is None else repr(np.asarray(dbcn).tolist()), self._bc_epoch)
        if key == self._bc_key:
            it, rr = self._dev().solve_linear(None, None, None, None, rtol=self.solver_rtol, max_iter=self.solver_maxit)
        else:
            fixed, du, w, fext = self._calc_BC(bcl0, bcb0, dbcr, dbct, dbcn)
            it, rr = self._dev().solve_linear(fixed, du, w, fext, rtol=self.solver_rtol, max_iter=self.solver_maxit)
            self._bc_key = key
        self.solver_stats.append((it, rr))

    # ------------------------------------------------------------------------------------------------------
    # solve (model.py:979-1450)
    # ------------------------------------------------------------------------------------------------------
    def solve(self, min_step=None, verb=False):
        if self.Nnode is None:
            raise AttributeError('Attributes for mesh not set, but required by solver.')
        ctx = self._dev()
        dim = self.dim
        self._bc_key = None                      # BC types / node sets may have changed since the last call
        if self.u is None:                       # first call: zero state, elstiff = CV (1214-1234)
            ctx.set_field(_lib.F_U, np.zeros(self.Ndof))
            ctx.set_field(_lib.F_F, np.zeros(self.Ndof))
            z = np.zeros((self.Nel, 6))
            for fld in (_lib.F_SIG, _lib.F_EPS, _lib.F_EPL):
                ctx.set_field(fld, z)
            ctx.set_tangent(None)
            self.u = np.zeros(self.Ndof)
            self.f = np.zeros(self.Ndof)
            self.sgl = np.zeros((1, 6))
            self.egl = np.zeros((1, 6))
            self.epgl = np.zeros((1, 6))
            bcr0 = np.zeros(dim)
            bct0 = np.zeros(dim)
            self.bct_mem = np.zeros(dim)
            self.bcr_mem = np.zeros(dim)
            if self.noset is not None:
                bcn0 = np.zeros(dim)
                self.bcn_mem = np.zeros(dim)
        else:
            bcr0 = self.bcr_mem
            bct0 = self.bct_mem
            if self.noset is not None:
                bcn0 = self.bcn_mem
        bcl0, bcb0 = self.bcl, self.bcb
        ctx.setupK()                             # 1242
        # loading direction for the ML yield-point search in calc_scf (1245-1258)
        sld = np.zeros(6)
        if np.abs(self.bcr[0]) > 1.e-6:
            sld[0] = np.sign(self.bcr[0])
        if np.abs(self.bct[1]) > 1.e-6:
            sld[1] = np.sign(self.bct[1])
        if np.abs(self.bcr[1]) > 1.e-6:
            sld[5] = np.sign(self.bcr[1])
        if np.abs(self.bct[0]) > 1.e-6:
            sld[5] = np.sign(self.bct[0])
        if np.linalg.norm(sld) < 1.e-3:
            warnings.warn('solve: inconsistent BC sld={}, bct={}, bcr={}'.format(sld, self.bct, self.bcr))
            sld[0] = 1.

        il = 0
        nit = 0
        niter = []
        co_nconv = []
        bc_inc = True
        nconv = 0
        self.sweep_stats = []
        while bc_inc:                            # load increments (1270)
            max_dbct = self.bct - bct0
            max_dbcr = self.bcr - bcr0
            if min_step is not None:
                sc = np.maximum(1, min_step - il)
                max_dbct = max_dbct / sc
                max_dbcr = max_dbcr / sc
            dbcr = max_dbcr
            dbct = max_dbct
            if self.noset is not None:
                max_dbcn = self.bcn - bcn0
                if min_step is not None:
                    max_dbcn = max_dbcn / np.maximum(1, min_step - il)
                dbcn = max_dbcn
            else:
                max_dbcn = None
                dbcn = None
            # elastic predictor: final solution of a linear model (1290-1291)
            self._solve_increment(bcl0, bcb0, dbcr, dbct, dbcn)
            if self.nonlin:
                scale_bc = ctx.calc_scf(sld) if il < 10 else 1.          # 1296
                dbcr = max_dbcr * scale_bc
                dbct = max_dbct * scale_bc
                nit = 0
                change = True
                conv = False
                if verb:
                    print('***Load step #', il)
                    print('scaling factor', scale_bc)
                while (change or not conv) and nit <= 15:                 # 1305
                    if il < 6 and nit > 1:                                # halve the increment (1308-1330)
                        hs = 0.5

                        def shrink(tgt, reached, cur, mx):
                            for k in range(dim):
                                if mx[k] >= 0:
                                    cur[k] = np.maximum(0.05 * mx[k], np.minimum(tgt[k] - reached[k], cur[k] * hs))
                                else:
                                    cur[k] = np.minimum(0.05 * mx[k], np.maximum(tgt[k] - reached[k], cur[k] * hs))
                        shrink(self.bcr, bcr0, dbcr, max_dbcr)
                        shrink(self.bct, bct0, dbct, max_dbct)
                        if self.noset is not None:
                            shrink(self.bcn, bcn0, dbcn, max_dbcn)
                    ctx.setupK()                                          # 1333
                    self._solve_increment(bcl0, bcb0, dbcr, dbct, dbcn)   # 1334-1335
                    change, conv, max_hh, counts = ctx.sweep(nit)         # 1338-1361 on the device
                    self.sweep_stats.append((il, nit, change, conv, max_hh, counts))
                    if verb:
                        if not conv:
                            print('\n  ###  Warning: No convergence of plasticity algorithm in trial step #', nit)
                        print('+++Inner trial step #', nit, 'branches [elastic, 1 step, 50 sub-steps, corrections]:', counts)
                        print('load increment right:', dbcr)
                        print('load increment top:', dbct)
                        if self.noset is not None:
                            print('load increment set:', dbcn)
                    if not conv:
                        nconv += 1
                    nit += 1
            # commit (1383-1392).  f += K du uses the matrix the increment was solved with: the last sweep may have
            # replaced element tangents, but like in the reference K is the one assembled before that sweep -- it is
            # still in the CSR memory of the device.
            self._commit()
            il += 1                                                       # 1395-1413
            niter.append(nit - 1)
            co_nconv.append(nconv)
            bcr0 += dbcr
            hl0 = np.abs(bcr0[0] - self.bcr[0]) > 1.e-6 and np.abs(self.bcr[0]) > 1.e-9
            hl1 = np.abs(bcr0[1] - self.bcr[1]) > 1.e-6 and np.abs(self.bcr[1]) > 1.e-9
            bct0 += dbct
            hr0 = np.abs(bct0[0] - self.bct[0]) > 1.e-6 and np.abs(self.bct[0]) > 1.e-9
            hr1 = np.abs(bct0[1] - self.bct[1]) > 1.e-6 and np.abs(self.bct[1]) > 1.e-9
            if self.noset is not None:
                bcn0 += dbcn
                hr0 = hr0 or (np.abs(bcn0[0] - self.bcn[0]) > 1.e-6 and np.abs(self.bcn[0]) > 1.e-9)
                hr1 = hr1 or (np.abs(bcn0[1] - self.bcn[1]) > 1.e-6 and np.abs(self.bcn[1]) > 1.e-9)
            bc_inc = bool(hr0 or hr1 or hl0 or hl1)
            self.calc_global()                                            # 1415-1418
            self.sgl = np.append(self.sgl, [self.glob['sig']], axis=0)
            self.egl = np.append(self.egl, [self.glob['eps']], axis=0)
            self.epgl = np.append(self.epgl, [self.glob['epl']], axis=0)
            if verb:
                print('Load increment ', il, 'top ', bct0, '/', self.bct, '; rhs', bcr0, '/', self.bcr)
                print('Global strain: ', np.around(self.glob['eps'], decimals=5))
                print('Global stress: ', np.around(self.glob['sig'], decimals=3))
                print('Global plastic strain: ', np.around(self.glob['epl'], decimals=6))
                print('----------------------------')
        self.bct_mem = bct0
        self.bcr_mem = bcr0
        if self.noset is not None:
            self.bcn_mem = bcn0
        self.nsteps = il
        self.niter = niter
        self.co_nconv = co_nconv

    def _commit(self):
        """u += du; f += K du; element state update (1383-1392) on the device; refresh the host copies of u, f."""
        ctx = self._dev()
        ctx.commit()
        self.u = ctx.get_field(_lib.F_U)
        self.f = ctx.get_field(_lib.F_F)
        self.du = ctx.get_field(_lib.F_DU)
        self._cache = {}

    # ------------------------------------------------------------------------------------------------------
    # post-processing (model.py:1452-1511)
    # ------------------------------------------------------------------------------------------------------
    def bcval(self, nodes):
        """Average displacement and total force of a node list (model.py:1452-1471)."""
        idx = np.asarray(nodes, dtype=int)
        ux, uy = self.u[2 * idx], self.u[2 * idx + 1]
        fx, fy = self.f[2 * idx], self.f[2 * idx + 1]
        n = len(idx)
        return ux.sum() / n, uy.sum() / n, fx.sum(), fy.sum()

    def calc_global(self):
        """Global quantities from boundary nodes and from volume-averaged element solutions (model.py:1473-1511)."""
        uxl, uyl, fxl, fyl = self.bcval(self.noleft)
        uxr, uyr, fxr, fyr = self.bcval(self.noright)
        g = self.glob
        g['ebc1'] = (uxr - uxl) / self.lenx
        g['sbc1'] = 0.5 * (fxr - fxl) / (self.leny * self.thick)
        g['ebc21'] = (uyr - uyl) / self.lenx
        g['sbc21'] = 0.5 * (fyr - fyl) / (self.leny * self.thick)
        uxb, uyb, fxb, fyb = self.bcval(self.nobot)
        uxt, uyt, fxt, fyt = self.bcval(self.notop)
        g['ebc2'] = (uyt - uyb) / self.leny
        g['sbc2'] = 0.5 * (fyt - fyb) / (self.lenx * self.thick)
        g['ebc12'] = (uxt - uxb) / self.leny
        g['sbc12'] = 0.5 * (fxt - fxb) / (self.lenx * self.thick)
        g['sig'], g['eps'], g['epl'] = self._dev().reduce_global()

    def plot(self, *a, **k):
        raise NotImplementedError('Model.plot needs matplotlib and is outside the accelerated path; use the element '
                                  'fields (element[i].sig / .eps / .epl, npos) with your own plotting')
