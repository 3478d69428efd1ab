"""Model -- host-side mirror of pyLabFEA's Model class (model.py) for 2-d structured quadrilateral meshes.

Keeps the reference's API (geom / assign / bcleft / bcright / bcbot / bctop / bcnode / mesh / setupK / solve /
calc_global / bcval and the attributes notebooks read: u, f, glob, sgl, egl, epgl, element[i].sig/.eps/.epl/...,
Nel, Ndof, npos, noleft/noright/nobot/notop, nsteps, niter, co_nconv), but every O(Nel) / O(Ndof) operation of
Model.solve runs on the device through the C-ABI: K integration + CSR assembly, BC-reduced CG solve, the
Material.response sweep with tangent replacement, calc_scf, the state commit and the volume averages; the loop over
load increments itself is the library's (plf_load_increment).  Per increment one report struct crosses the boundary;
u, f and the element fields are copied to the host only when they are read.

File:line citations refer to /root/reference/src/pylabfea/model.py.
"""
import warnings

import numpy as np

from . import _lib


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
        self._solved = False          # u, f, du live on the device; the properties mirror them on demand
        self._part = None             # (rank, world, NCCL unique id) of a strip-partitioned model, see distribute()
        self.Nnode = None
        self.glob = {'ebc1': None, 'ebc2': None, 'sbc1': None, 'sbc2': None,
                     'eps': np.zeros(6), 'sig': np.zeros(6), 'epl': np.zeros(6)}
        self._ctx = None
        self._cache = {}
        self.solver_rtol = 1.e-13     # CG stops at |r| <= rtol |b|
        self.solver_maxit = 200000
        self.solver_stats = []        # (iterations, relative residual) of every linear solve

    # ------------------------------------------------------------------------------------------------------
    # pre-processing: geometry, materials, boundary conditions (host, O(1))
    # ------------------------------------------------------------------------------------------------------
    def geom(self, sect=1, LX=None, LY=1., LZ=1.):
        """Model dimensions and sections (model.py:514-552): either the list of section lengths, or a number of equally
        long sections that share LX."""
        if type(sect) is list:
            lengths, total = np.array(sect), sum(sect)
        elif type(sect) is int:
            if sect < 1:
                raise ValueError('At least one section must be defined.')
            if LX is None:
                raise ValueError('LX must be given if sect is of type int')
            lengths, total = np.ones(sect) * LX / sect, LX
        else:
            raise TypeError('Sect must be either list or int, not {}'.format(type(sect)))
        self.Nsec, self.LS, self.lenx = len(lengths), lengths, total
        self.leny, self.thick = LY, LZ

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
        if self._solved:
            warnings.warn('Warning: Solution of previous steps is deleted')
            self._solved = False
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

    def distribute(self, rank, world, unique_id, solver='replicated'):
        """Make this process one rank of a strip-partitioned solve (one process per GPU).  Every rank builds
        the SAME global model (geom / assign / bc* / mesh with the global NX, NY) and calls distribute() before the first
        device call; its context then holds only the rank's strip of element columns (+ one halo column on the left,
        pylabfea_b200/shard.py), and solve() runs the load increments on all ranks in lockstep: the linear solves
        exchange one halo node column per CG iteration and all-reduce the dot products (NCCL), sweep flags, calc_scf
        and the global averages are combined over the ranks.  glob / sgl / egl / epgl / nsteps / niter are the global
        model's on every rank; u, f and element[i] refer to the rank's LOCAL node grid and elements (local_info()).

        solver = 'replicated' (default): the element work is sharded, the owned K rows of all ranks are gathered and every
        rank solves the complete system with the single-GPU multigrid-CG (du bit-identical to the single-GPU run, no
        communication inside the solve; each rank additionally holds the global K).  solver = 'distributed': the strips
        solve together (halo exchange + all-reduces per CG iteration, one V-cycle per strip as preconditioner): nothing
        global is stored, but the iteration count grows with the number of strips."""
        if self._on_device:
            raise AttributeError('distribute() must be called before the model is used on the device')
        if solver not in ('replicated', 'distributed'):
            raise ValueError("distribute: solver must be 'replicated' or 'distributed'")
        self._part = (int(rank), int(world), bytes(unique_id))
        self._dist_solver = solver

    def local_info(self):
        """Strip of a distributed model held by this rank (shard.strip_shard) or None."""
        if self._part is None:
            return None
        from .shard import strip_shard
        return strip_shard(self._NX, self._part[0], self._part[1])

    def _dev(self):
        """The device context with materials and mesh uploaded (created on first use: no CPU fallback exists, so this
        raises PlfError without libplfb200.so or a CUDA device)."""
        if not self._on_device:
            if self._ctx is None:
                self._ctx = _lib.Context(self._device)
            self._ctx.set_materials([m._descriptor() for m in self.mat], planestress=self.planestress)
            if self._part is None:
                self._ctx.mesh_structured(self._NX, self._NY, self._dx, self._dy, self.thick, self._elem_mat)
            else:
                rank, world, uid = self._part
                sh = self.local_info()
                j0, nxl = sh['first_col_global'], sh['nx_local']
                em = self._elem_mat.reshape(self._NX, self._NY)[j0:j0 + nxl]
                self._ctx.mesh_structured(nxl, self._NY, self._dx[j0:j0 + nxl], self._dy, self.thick, np.ascontiguousarray(em).ravel())
                if not getattr(self, '_comm_up', False):
                    self._ctx.comm_init(uid, rank, world)
                    self._comm_up = True
                cols = sh['owned_node_cols_local']
                self._ctx.comm_set_partition(cols[0], cols[-1] + 1, j0, self._NX, float(self.npos[2 * j0 * self.NnodeY]), self.lenx)
                if self._dist_solver == 'replicated':
                    self._gctx = _lib.Context(self._device)
                    self._gctx.set_materials([m._descriptor() for m in self.mat], planestress=self.planestress)
                    self._gctx.mesh_structured(self._NX, self._NY, self._dx, self._dy, self.thick, self._elem_mat)
                    self._ctx.comm_set_replica(self._gctx)
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
    # boundary conditions (model.py:1070-1206)
    # ------------------------------------------------------------------------------------------------------
    def _bc_descriptor(self):
        """The boundary conditions as the plain struct the library's load-increment controller reads (plf_bc)."""
        bc = _lib.PlfBC()
        for name, vals, flags in (('left', self.bcl, self.ubcleft), ('bot', self.bcb, self.ubcbot),
                                  ('right', self.bcr, self.ubcright), ('top', self.bct, self.ubctop),
                                  ('node', self.bcn, self.ubcn)):
            getattr(bc, name)[:] = [float(v) for v in vals]
            getattr(bc, name + '_disp')[:] = [int(bool(f)) for f in flags]
        if self.noset is not None:
            self._noset_arr = np.ascontiguousarray(self.noset, dtype=np.int32)
            bc.n_noset = len(self._noset_arr)
            bc.noset = _lib._ptr(self._noset_arr)
        return bc

    def _calc_BC(self, bcl0, bcb0, dbcr, dbct, dbcn):
        """Host-side statement of what calc_BC leaves behind, as four dense vectors: `fixed` (dofs taken out of `ind`),
        `du` (their prescribed increments), `w` (the sum of ALL displacement values applied to a dof -- the reference
        multiplies column i of K into df once per edge that lists the dof, also at shared corners) and `f_ext` (force
        conditions), so that df[ind] = f_ext[ind] - (K w)[ind].  The device path builds the same vectors inside
        plf_load_increment; this NumPy version serves the host-only checks and plf_solve_linear callers."""
        n = self.Ndof
        fixed = np.zeros(n, dtype=np.uint8)
        du, w, fext = np.zeros(n), np.zeros(n), np.zeros(n)
        edges = [('left', self.noleft, self.ubcleft, bcl0), ('bottom', self.nobot, self.ubcbot, bcb0),
                 ('right', self.noright, self.ubcright, dbcr), ('top', self.notop, self.ubctop, dbct)]
        if self.noset is not None:
            if dbcn is None:
                raise ValueError('No BC for selected node set given.')
            edges.append(('node set', self.noset, self.ubcn, dbcn))
        for who, nodes, flags, vals in edges:
            nodes = np.asarray(nodes, dtype=np.int64)
            for k in range(2):
                dofs = 2 * nodes + k
                if flags[k]:
                    fresh = fixed[dofs] == 0
                    clash = ~fresh & (du[dofs] != vals[k])
                    if clash.any():
                        j = int(nodes[np.argmax(clash)])
                        warnings.warn('Inconsistent BC at %s node %d (%g vs %g).' % (who, j, du[2 * j + k], vals[k]))
                    du[dofs[fresh]] = vals[k]
                    fixed[dofs] = 1
                    np.add.at(w, dofs, vals[k])
                elif who == 'right' or who == 'top':
                    # a force is shared evenly over the edge's nodes, the two corner nodes carry half a share
                    along = self.npos[2 * nodes + (1 if who == 'right' else 0)]
                    length, nseg = (self.leny, self.NnodeY - 1) if who == 'right' else (self.lenx, self.NnodeX - 1)
                    share = np.where((along < 1.e-3) | (along > length - 1.e-3), 0.5, 1.) / nseg
                    np.add.at(fext, dofs, vals[k] * share)
                elif who == 'node set':
                    np.add.at(fext, dofs, vals[k])
        return fixed, du, w, fext

    # ------------------------------------------------------------------------------------------------------
    # nodal solution vectors: device-resident, mirrored to the host when somebody looks at them
    # ------------------------------------------------------------------------------------------------------
    def _nodal(self, name, field):
        if not self._solved:
            return None
        if name not in self._cache:
            self._cache[name] = self._dev().get_field(field)
        return self._cache[name]

    @property
    def u(self):
        return self._nodal('u', _lib.F_U)

    @u.setter
    def u(self, val):
        if val is not None:
            raise AttributeError('Model.u is computed by solve(); only None (forget the solution) can be assigned')
        self._solved = False

    @property
    def f(self):
        return self._nodal('f', _lib.F_F)

    @f.setter
    def f(self, val):
        if val is not None:
            raise AttributeError('Model.f is computed by solve(); only None can be assigned')

    @property
    def du(self):
        return self._nodal('du', _lib.F_DU)

    # ------------------------------------------------------------------------------------------------------
    # solve (model.py:979-1450)
    # ------------------------------------------------------------------------------------------------------
    def solve(self, min_step=None, verb=False):
        """Load-controlled solution (reference: Model.solve).  The loop over load increments lives in the library
        (plf_load_increment: predictor solve, calc_scf, the inner passes with tangent update, commit, calc_global);
        this method hands over the boundary conditions, keeps the increment history the reference exposes (sgl / egl /
        epgl, glob, nsteps, niter, co_nconv, bcr_mem / bct_mem) and prints the verbose log."""
        if self.Nnode is None:
            raise AttributeError('Attributes for mesh not set, but required by solver.')
        ctx = self._dev()
        first = not self._solved
        state = _lib.PlfLoadState()
        if first:
            self.sgl, self.egl, self.epgl = np.zeros((1, 6)), np.zeros((1, 6)), np.zeros((1, 6))
            self.bcr_mem, self.bct_mem = np.zeros(self.dim), np.zeros(self.dim)
            if self.noset is not None:
                self.bcn_mem = np.zeros(self.dim)
        state.reached_right[:] = list(self.bcr_mem)
        state.reached_top[:] = list(self.bct_mem)
        # The reference starts the node-set history of the first call from a private zero vector and never stores it
        # (model.py:1233-1234, 1446-1447), so the SECOND call starts from zero again; from then on bcn_mem accumulates.
        carry_set = self.noset is not None and not first and hasattr(self, 'bcn_mem')
        if carry_set:
            state.reached_node[:] = list(self.bcn_mem)
        sld = np.zeros(6)
        for val, idx in ((self.bcr[0], 0), (self.bct[1], 1), (self.bcr[1], 5), (self.bct[0], 5)):
            if abs(val) > 1.e-6:
                sld[idx] = np.sign(val)
        if np.linalg.norm(sld) < 1.e-3:
            warnings.warn('solve: inconsistent BC sld={}, bct={}, bcr={}'.format(sld, self.bct, self.bcr))
        ctx.solve_begin(self._bc_descriptor(), reset=first)
        self._solved = True
        self._cache = {}
        niter, co_nconv = [], []
        self.sweep_stats = []
        state.more = 1
        hook = getattr(self, 'increment_hook', None)   # optional callable(model, il), called before every load increment
        while state.more:
            il = state.il
            if hook is not None:
                hook(self, il)
            rep = ctx.load_increment(state, min_step, self.solver_rtol, self.solver_maxit)
            self._cache = {}
            if rep.n_warnings:
                warnings.warn(rep.warning.decode())
            self.solver_stats += [(int(rep.pass_cg_iters[i]), rep.max_relres) for i in range(rep.solves)]
            for nit in range(rep.passes):
                self.sweep_stats.append((il, nit, bool(rep.pass_change[nit]), bool(rep.pass_conv[nit]), rep.pass_max_hh[nit],
                                         list(rep.pass_counts[nit])))
            niter.append(rep.passes - 1)
            co_nconv.append(int(state.nconv))
            self._store_global(rep)
            if verb:
                self._report_increment(il, rep, state)
        self.bcr_mem = np.array(state.reached_right[:])
        self.bct_mem = np.array(state.reached_top[:])
        if carry_set:
            self.bcn_mem = np.array(state.reached_node[:])
        self.nsteps = int(state.il)
        self.niter = niter
        self.co_nconv = co_nconv

    def _store_global(self, rep):
        """calc_global (model.py:1473-1511) from the increment report, appended to the history (1415-1418)."""
        self._set_boundary_globals(np.array([list(row) for row in rep.edge]))
        g = self.glob
        g['sig'], g['eps'], g['epl'] = (np.array(v[:]) for v in (rep.glob_sig, rep.glob_eps, rep.glob_epl))
        self.sgl = np.append(self.sgl, [g['sig']], axis=0)
        self.egl = np.append(self.egl, [g['eps']], axis=0)
        self.epgl = np.append(self.epgl, [g['epl']], axis=0)

    def _set_boundary_globals(self, edge):
        """Boundary strains and stresses from the bcval rows of the left, right, bottom and top nodes (1491-1498)."""
        (uxl, uyl, fxl, fyl), (uxr, uyr, fxr, fyr), (uxb, uyb, fxb, fyb), (uxt, uyt, fxt, fyt) = edge
        g = self.glob
        ax, ay = self.leny * self.thick, self.lenx * self.thick   # cross sections normal to x and to y
        g['ebc1'], g['sbc1'] = (uxr - uxl) / self.lenx, 0.5 * (fxr - fxl) / ax
        g['ebc21'], g['sbc21'] = (uyr - uyl) / self.lenx, 0.5 * (fyr - fyl) / ax
        g['ebc2'], g['sbc2'] = (uyt - uyb) / self.leny, 0.5 * (fyt - fyb) / ay
        g['ebc12'], g['sbc12'] = (uxt - uxb) / self.leny, 0.5 * (fxt - fxb) / ay

    def _report_increment(self, il, rep, state):
        print('== increment %d: predictor scaled by %.6g, %d inner passes, %d linear solves (%d CG iterations)'
              % (il, rep.scale_bc, rep.passes, rep.solves, rep.cg_iterations))
        for nit in range(rep.passes):
            print('   pass %2d: %s, tangents %s, max |dC| %.3g, elements [elastic, 1 step, 50 sub-steps, corrected] = %s, CG %d'
                  % (nit, 'yield condition met' if rep.pass_conv[nit] else 'NOT converged (f > 5e-3)',
                     'changed' if rep.pass_change[nit] else 'unchanged', rep.pass_max_hh[nit], list(rep.pass_counts[nit]),
                     rep.pass_cg_iters[1 + nit]))
        steps = 'right %s, top %s' % (list(rep.step_right), list(rep.step_top))
        if self.noset is not None:
            steps += ', node set %s' % list(rep.step_node)
        print('   applied: ' + steps)
        print('   reached: top %s of %s, right %s of %s' % (list(state.reached_top), list(self.bct), list(state.reached_right),
                                                          list(self.bcr)))
        for label, key, nd in (('strain', 'eps', 5), ('stress', 'sig', 3), ('plastic strain', 'epl', 6)):
            print('   mean %s: %s' % (label, np.around(self.glob[key], decimals=nd)))

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
        ctx = self._dev()
        self._set_boundary_globals(ctx.boundary_values())
        self.glob['sig'], self.glob['eps'], self.glob['epl'] = ctx.reduce_global()

    def plot(self, *a, **k):
        raise NotImplementedError('Model.plot needs matplotlib and is outside the accelerated path; use the element '
                                  'fields (element[i].sig / .eps / .epl, npos) with your own plotting')
