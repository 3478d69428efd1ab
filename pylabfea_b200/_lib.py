"""ctypes binding of libplfb200.so (include/plfb200.h).  No CPU fallback: if the library or a CUDA device is
missing every entry point raises."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libplfb200.so')

PLF_OK, PLF_EINVAL, PLF_ECUDA, PLF_ESTATE, PLF_EUNSUPPORTED, PLF_ENOCONV = 0, -1, -2, -3, -4, -5
MAT_ELASTIC, MAT_HILL, MAT_SVC, MAT_BARLAT, MAT_HILL3, MAT_SVC3 = 0, 1, 2, 3, 4, 5
(F_SIG, F_EPS, F_EPL, F_RES_SIG, F_RES_DEPL, F_U, F_DU, F_F, F_FYLD, F_NSTEPS) = range(10)

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_bp = C.POINTER(C.c_uint8)


class PlfMaterial(C.Structure):
    """struct plf_material (include/plfb200.h)"""
    _fields_ = [('kind', C.c_int32), ('nsv', C.c_int32), ('dev_only', C.c_int32), ('tresca', C.c_int32),
                ('lhs', C.c_int32), ('pad_', C.c_int32),
                ('E', C.c_double), ('nu', C.c_double), ('C44', C.c_double), ('CV', C.c_double * 36),
                ('sy', C.c_double), ('khard', C.c_double), ('hill', C.c_double * 6), ('d', C.c_double * 3),
                ('barlat_m1', C.c_double * 36), ('barlat_m2', C.c_double * 36), ('barlat_exp', C.c_double),
                ('svc_gamma', C.c_double), ('svc_rho', C.c_double), ('svc_scale_seq', C.c_double),
                ('svc_sv', _dp), ('svc_dual', _dp), ('svc_nfeat', C.c_int32), ('pad2_', C.c_int32), ('svc_scale_wh', C.c_double)]


MAX_PASSES = 17


class PlfBC(C.Structure):
    """struct plf_bc (include/plfb200.h): Model.bcl / bcb / bcr / bct / bcn and the ubc* flags as plain data"""
    _fields_ = [('left', C.c_double * 2), ('bot', C.c_double * 2), ('right', C.c_double * 2), ('top', C.c_double * 2),
                ('node', C.c_double * 2), ('left_disp', C.c_int32 * 2), ('bot_disp', C.c_int32 * 2),
                ('right_disp', C.c_int32 * 2), ('top_disp', C.c_int32 * 2), ('node_disp', C.c_int32 * 2),
                ('n_noset', C.c_int32), ('pad_', C.c_int32), ('noset', _ip)]


class PlfLoadState(C.Structure):
    """struct plf_load_state: boundary values reached so far, increment counter, continuation flag"""
    _fields_ = [('reached_right', C.c_double * 2), ('reached_top', C.c_double * 2), ('reached_node', C.c_double * 2),
                ('il', C.c_int32), ('nconv', C.c_int32), ('more', C.c_int32), ('pad_', C.c_int32)]


class PlfIncrementReport(C.Structure):
    """struct plf_increment_report: what one load increment did"""
    _fields_ = [('passes', C.c_int32), ('solves', C.c_int32), ('cg_iterations', C.c_int64), ('max_relres', C.c_double),
                ('scale_bc', C.c_double), ('step_right', C.c_double * 2), ('step_top', C.c_double * 2),
                ('step_node', C.c_double * 2), ('glob_sig', C.c_double * 6), ('glob_eps', C.c_double * 6),
                ('glob_epl', C.c_double * 6), ('edge', (C.c_double * 4) * 4),
                ('pass_change', C.c_int32 * MAX_PASSES), ('pass_conv', C.c_int32 * MAX_PASSES),
                ('pass_max_hh', C.c_double * MAX_PASSES), ('pass_counts', (C.c_int64 * 4) * MAX_PASSES),
                ('pass_cg_iters', C.c_int32 * (MAX_PASSES + 1)), ('n_warnings', C.c_int32), ('warning', C.c_char * 200)]


class PlfError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (building it is __graft_entry__.build()'s / `python -m pylabfea_b200.build`'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PlfError('%s not found: build it with `python -m pylabfea_b200.build` (nvcc, sm_100a). '
                       'pylabfea_b200 has no CPU fallback.' % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.plf_last_error.restype = C.c_char_p
    L.plf_last_error.argtypes = [C.c_void_p]
    L.plf_version.restype = C.c_char_p
    L.plf_launch_count.restype = C.c_int64
    L.plf_launch_count.argtypes = [C.c_void_p]
    L.plf_destroy.restype = None
    L.plf_destroy.argtypes = [C.c_void_p]
    L.plf_batch_destroy.restype = None
    L.plf_batch_destroy.argtypes = [C.c_void_p, C.c_void_p]
    L.plf_pinned_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
    L.plf_pinned_free.argtypes = [C.c_void_p]
    L.plf_sizeof_material.restype = C.c_size_t
    if L.plf_sizeof_material() != C.sizeof(PlfMaterial):
        raise PlfError('plf_material layout mismatch between libplfb200.so and the ctypes binding')
    L.plf_sizeof_increment_report.restype = C.c_size_t
    if L.plf_sizeof_increment_report() != C.sizeof(PlfIncrementReport):
        raise PlfError('plf_increment_report layout mismatch between libplfb200.so and the ctypes binding')
    _lib = L
    return L


def _arr(x, dtype=np.float64):
    return np.ascontiguousarray(x, dtype=dtype)


def _ptr(a):
    if a is None:
        return None
    if not a.flags.c_contiguous:
        raise ValueError('array passed to libplfb200 must be C-contiguous')
    if a.dtype == np.float64:
        return a.ctypes.data_as(_dp)
    if a.dtype == np.int32:
        return a.ctypes.data_as(_ip)
    if a.dtype == np.uint8:
        return a.ctypes.data_as(_bp)
    raise TypeError(a.dtype)


def pinned_empty(shape, dtype=np.float64):
    """numpy array backed by page-locked host memory (freed when the array's base object dies)."""
    L = load()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = C.c_void_p()
    if L.plf_pinned_alloc(C.c_size_t(max(n, 1)), C.byref(p)) != PLF_OK:
        raise PlfError('plf_pinned_alloc failed (no CUDA device?)')

    class _Owner:
        def __init__(self, ptr):
            self.ptr = ptr

        def __del__(self):
            try:
                L.plf_pinned_free(self.ptr)
            except Exception:
                pass
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    buf._owner = _Owner(p)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class Context:
    """One plf_ctx: one CUDA device + stream.  Raises PlfError when no device is available."""

    def __init__(self, device=0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.plf_create(int(device), C.byref(h))
        if rc != PLF_OK:
            raise PlfError(self.L.plf_last_error(None).decode())
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, 'h', None):
            self.L.plf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- errors ------------------------------------------------------------------------------------------
    def _ck(self, rc):
        if rc == PLF_OK:
            return
        msg = self.L.plf_last_error(self.h).decode()
        if rc == PLF_EINVAL:
            raise ValueError(msg)
        if rc == PLF_ESTATE:
            raise AttributeError(msg)
        if rc == PLF_EUNSUPPORTED:
            raise ValueError(msg)   # the reference raises ValueError for missing gradients (material.py:822-825)
        raise PlfError('%s (code %d)' % (msg, rc))

    # ---- misc --------------------------------------------------------------------------------------------
    def device_info(self):
        name = C.create_string_buffer(128)
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        self._ck(self.L.plf_device_info(self.h, name, 128, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return dict(name=name.value.decode(), sm_count=sm.value, cc=(ma.value, mi.value), mem_bytes=mem.value)

    def timer_start(self):
        self._ck(self.L.plf_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(self.L.plf_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def event_record(self, idx):
        self._ck(self.L.plf_event_record(self.h, int(idx)))

    def event_elapsed(self, i0, i1):
        ms = C.c_float()
        self._ck(self.L.plf_event_elapsed(self.h, int(i0), int(i1), C.byref(ms)))
        return ms.value

    def synchronize(self):
        self._ck(self.L.plf_synchronize(self.h))

    def flush_l2(self):
        self._ck(self.L.plf_flush_l2(self.h))

    def launch_count(self):
        return int(self.L.plf_launch_count(self.h))

    def last_kernels(self):
        """Kernel variants (ncu spelling) of the return mapping / assembly launched since the previous call."""
        buf = C.create_string_buffer(256)
        self._ck(self.L.plf_last_kernels(self.h, buf, 256))
        return [k for k in buf.value.decode().split(';') if k]

    def set_option(self, key, value):
        self._ck(self.L.plf_set_option(self.h, key.encode(), int(value)))

    # ---- materials ---------------------------------------------------------------------------------------
    def set_materials(self, mats, planestress):
        """mats: list of PlfMaterial (see material.Material._descriptor)."""
        arr = (PlfMaterial * len(mats))(*mats)
        self._keep = [arr, mats]
        self._ck(self.L.plf_set_materials(self.h, len(mats), arr, int(bool(planestress))))
        self.nmat = len(mats)

    def element_CV(self, mid):
        out = np.zeros((6, 6))
        self._ck(self.L.plf_get_element_CV(self.h, int(mid), _ptr(out)))
        return out

    # ---- batched constitutive calls ----------------------------------------------------------------------
    def response_batch(self, sig, epl, deps, mat_id=None, want_ctan=True):
        sig, epl, deps = _arr(sig).reshape(-1, 6), _arr(epl).reshape(-1, 6), _arr(deps).reshape(-1, 6)
        n = len(sig)
        if epl.shape != sig.shape or deps.shape != sig.shape:
            raise ValueError('sig, epl and deps must have the same shape (n,6)')
        mid = None if mat_id is None else _arr(mat_id, np.int32)
        fy = np.empty(n)
        so = np.empty((n, 6))
        dp = np.empty((n, 6))
        ct = np.empty((n, 6, 6)) if want_ctan else None
        ns = np.empty(n, dtype=np.int32)
        self._ck(self.L.plf_response_batch(self.h, C.c_int64(n), _ptr(mid), _ptr(sig), _ptr(epl), _ptr(deps), _ptr(fy),
                                           _ptr(so), _ptr(dp), _ptr(ct), _ptr(ns)))
        return fy, so, dp, ct, ns

    def _prim(self, fn, mat, sig, width, *extra):
        sig = _arr(sig).reshape(-1, 6)
        n = len(sig)
        out = np.empty((n, width) if width > 1 else n)
        args = [self.h, int(mat), C.c_int64(n), _ptr(sig)]
        keep = []
        for x in extra:
            if x is None:
                args.append(None)
            else:
                a = _arr(x)
                keep.append(a)
                args.append(_ptr(a))
        args.append(_ptr(out))
        self._ck(fn(*args))
        return out

    def calc_seq(self, mat, sig):
        return self._prim(self.L.plf_calc_seq_batch, mat, sig, 1)

    def calc_yf(self, mat, sig, epl=None):
        sig = _arr(sig).reshape(-1, 6)
        epl = np.zeros_like(sig) if epl is None else np.broadcast_to(_arr(epl), sig.shape)
        return self._prim(self.L.plf_calc_yf_batch, mat, sig, 1, epl)

    def calc_fgrad(self, mat, sig, epl=None):
        sig = _arr(sig).reshape(-1, 6)
        epl = np.zeros_like(sig) if epl is None else np.broadcast_to(_arr(epl), sig.shape)
        return self._prim(self.L.plf_calc_fgrad_batch, mat, sig, 6, epl)

    def ml_full_yf(self, mat, sig, epl=None, ld=None):
        sig = _arr(sig).reshape(-1, 6)
        epl = np.zeros_like(sig) if epl is None else np.broadcast_to(_arr(epl), sig.shape)
        return self._prim(self.L.plf_ml_full_yf_batch, mat, sig, 1, epl, ld)

    def epl_dot(self, mat, sig, epl, deps):
        sig = _arr(sig).reshape(-1, 6)
        return self._prim(self.L.plf_epl_dot_batch, mat, sig, 6, np.broadcast_to(_arr(epl), sig.shape),
                          np.broadcast_to(_arr(deps), sig.shape))

    def C_tan(self, mat, sig, epl=None):
        sig = _arr(sig).reshape(-1, 6)
        epl = np.zeros_like(sig) if epl is None else np.broadcast_to(_arr(epl), sig.shape)
        return self._prim(self.L.plf_ctan_batch, mat, sig, 36, epl).reshape(-1, 6, 6)

    # ---- device-resident batch ---------------------------------------------------------------------------
    def batch_create(self, n):
        b = C.c_void_p()
        self._ck(self.L.plf_batch_create(self.h, C.c_int64(n), C.byref(b)))
        return b

    def batch_destroy(self, b):
        self.L.plf_batch_destroy(self.h, b)

    def batch_upload(self, b, sig, epl, deps, mat_id=None):
        mid = None if mat_id is None else _arr(mat_id, np.int32)
        sig, epl, deps = _arr(sig), _arr(epl), _arr(deps)
        if not (sig.size == epl.size == deps.size) or sig.size % 6:
            raise ValueError('batch_upload: sig, epl and deps must all be (n, 6)')
        self._ck(self.L.plf_batch_upload(self.h, b, _ptr(mid), _ptr(sig), _ptr(epl), _ptr(deps)))

    def batch_run(self, b):
        self._ck(self.L.plf_batch_run(self.h, b))

    def batch_download(self, b, fy=None, sig_out=None, depl=None, ctan=None, nsteps=None):
        self._ck(self.L.plf_batch_download(self.h, b, _ptr(fy), _ptr(sig_out), _ptr(depl), _ptr(ctan), _ptr(nsteps)))

    # ---- mesh / stiffness --------------------------------------------------------------------------------
    def mesh_structured(self, NX, NY, dx, dy, thick, elem_mat):
        dx = _arr(dx)
        em = _arr(elem_mat, np.int32).ravel()
        if len(dx) != NX or len(em) != NX * NY:
            raise ValueError('dx must have NX entries and elem_mat NX*NY')
        self._ck(self.L.plf_mesh_structured(self.h, int(NX), int(NY), _ptr(dx), C.c_double(dy), C.c_double(thick), _ptr(em)))
        nel, nn, nd, nnz = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int64()
        self._ck(self.L.plf_mesh_sizes(self.h, C.byref(nel), C.byref(nn), C.byref(nd), C.byref(nnz)))
        self.Nel, self.Nnode, self.Ndof, self.nnz = nel.value, nn.value, nd.value, nnz.value

    def get_bmat(self, e):
        out = np.zeros((4, 6, 8))
        self._ck(self.L.plf_get_bmat(self.h, C.c_int64(e), _ptr(out)))
        return out

    def set_tangent(self, elstiff=None):
        a = None if elstiff is None else _arr(elstiff).reshape(self.Nel, 36)
        self._ck(self.L.plf_set_tangent(self.h, _ptr(a)))

    def get_tangent(self):
        out = np.empty((self.Nel, 6, 6))
        self._ck(self.L.plf_get_tangent(self.h, _ptr(out)))
        return out

    def get_kel(self):
        out = np.empty((self.Nel, 8, 8))
        self._ck(self.L.plf_get_kel(self.h, _ptr(out)))
        return out

    def setupK(self):
        self._ck(self.L.plf_setupK(self.h))

    def setupK_from_batch(self, b):
        self._ck(self.L.plf_setupK_from_batch(self.h, b))

    def step_host(self, b, sig, epl, deps, fy=None, sig_out=None, depl=None, ctan=None, nsteps=None, kval=None):
        """plf_step_host: Material.response for every mesh element + setupK, host buffers in and out, pipelined."""
        self._ck(self.L.plf_step_host(self.h, b, _ptr(sig), _ptr(epl), _ptr(deps), _ptr(fy), _ptr(sig_out), _ptr(depl),
                                      _ptr(ctan), _ptr(nsteps), _ptr(kval)))

    def step_host_packed(self, b, sig, epl, deps, ct21, elastic_words, fy=None, sig_out=None, depl=None, nsteps=None, kval=None):
        """plf_step_host_packed: as step_host, tangents as upper triangles of the non-elastic elements only (ct21, in
        element order) + one elastic bit per element; returns the number of packed tangents.  kval=None: K stays on the
        device."""
        n = C.c_int64()
        for a, need in ((ct21, self.Nel * 21), (elastic_words, (self.Nel + 31) // 32)):
            if not a.flags.c_contiguous or a.size < need:
                raise ValueError('step_host_packed: output buffer too small or not contiguous')
        self._ck(self.L.plf_step_host_packed(self.h, b, _ptr(sig), _ptr(epl), _ptr(deps), _ptr(fy), _ptr(sig_out), _ptr(depl),
                                             _ptr(nsteps), _ptr(ct21), elastic_words.ctypes.data_as(C.c_void_p), C.byref(n),
                                             _ptr(kval)))
        return n.value

    def get_csr_values(self, out):
        self._ck(self.L.plf_get_csr(self.h, None, None, _ptr(out)))
        return out

    def get_csr(self, values=True):
        rowptr = np.empty(self.Ndof + 1, dtype=np.int32)
        col = np.empty(self.nnz, dtype=np.int32)
        val = np.empty(self.nnz) if values else None
        self._ck(self.L.plf_get_csr(self.h, _ptr(rowptr), _ptr(col), _ptr(val)))
        return rowptr, col, val

    # ---- state -------------------------------------------------------------------------------------------
    def get_field(self, field):
        if field in (F_U, F_DU, F_F):
            out = np.empty(self.Ndof)
        elif field == F_FYLD:
            out = np.empty(self.Nel)
        elif field == F_NSTEPS:
            out = np.empty(self.Nel, dtype=np.int32)
        else:
            out = np.empty((self.Nel, 6))
        self._ck(self.L.plf_get_field(self.h, int(field), out.ctypes.data_as(C.c_void_p)))
        return out

    def set_field(self, field, a):
        a = _arr(a, np.int32 if field == F_NSTEPS else np.float64)
        self._ck(self.L.plf_set_field(self.h, int(field), a.ctypes.data_as(C.c_void_p)))

    def sweep(self, nit):
        change, conv, mh = C.c_int32(), C.c_int32(), C.c_double()
        counts = (C.c_int64 * 4)()
        self._ck(self.L.plf_sweep(self.h, int(nit), C.byref(change), C.byref(conv), C.byref(mh), counts))
        return bool(change.value), bool(conv.value), mh.value, list(counts)

    def commit(self):
        self._ck(self.L.plf_commit(self.h))

    def reduce_global(self):
        s, e, p = np.zeros(6), np.zeros(6), np.zeros(6)
        self._ck(self.L.plf_reduce_global(self.h, _ptr(s), _ptr(e), _ptr(p)))
        return s, e, p

    def calc_scf(self, sld):
        sld = _arr(sld)
        out = C.c_double()
        self._ck(self.L.plf_calc_scf(self.h, _ptr(sld), C.byref(out)))
        return out.value

    def solve_linear(self, fixed, du_bc, w, f_ext, rtol=1e-12, max_iter=200000):
        """fixed = du_bc = w = f_ext = None re-uses the boundary-condition vectors of the previous call."""
        if fixed is not None:
            fixed = _arr(fixed, np.uint8)
            du_bc, w, f_ext = _arr(du_bc), _arr(w), _arr(f_ext)
        it, rr = C.c_int(), C.c_double()
        self._ck(self.L.plf_solve_linear(self.h, _ptr(fixed), _ptr(du_bc), _ptr(w), _ptr(f_ext), C.c_double(rtol),
                                         int(max_iter), C.byref(it), C.byref(rr)))
        return it.value, rr.value

    # ---- load-increment controller (plf_solve_begin / plf_load_increment) -----------------------------------
    def solve_begin(self, bc, reset):
        """bc: PlfBC.  reset: zero u, f and the element state (first Model.solve after meshing)."""
        self._ck(self.L.plf_solve_begin(self.h, C.byref(bc), int(bool(reset))))

    def load_increment(self, state, min_step, rtol, max_iter):
        """One load increment (model.py:1270-1418); `state` (PlfLoadState) is updated in place."""
        rep = PlfIncrementReport()
        self._ck(self.L.plf_load_increment(self.h, C.byref(state), int(min_step or 0), C.c_double(rtol), int(max_iter),
                                           C.byref(rep)))
        return rep

    def boundary_values(self):
        """Model.bcval of the left, right, bottom and top nodes: rows (mean ux, mean uy, sum fx, sum fy)."""
        e = ((C.c_double * 4) * 4)()
        self._ck(self.L.plf_boundary_values(self.h, e))
        return np.array([[e[i][j] for j in range(4)] for i in range(4)])

    # ---- strip-partitioned solve over several GPUs (plf_comm_*) --------------------------------------------
    def comm_init(self, unique_id, rank, world):
        """Collective over all ranks: join the NCCL communicator identified by `unique_id` (comm_unique_id() of rank 0)."""
        if len(unique_id) != COMM_ID_BYTES:
            raise ValueError('unique_id must be the %d bytes of comm_unique_id()' % COMM_ID_BYTES)
        buf = (C.c_ubyte * COMM_ID_BYTES).from_buffer_copy(bytes(unique_id))
        self._ck(self.L.plf_comm_init(self.h, buf, int(rank), int(world)))

    def comm_set_strip(self, own_first, own_end):
        """Owned node columns [own_first, own_end) of the local mesh (shard.strip_shard: owned_node_cols_local)."""
        self._ck(self.L.plf_comm_set_strip(self.h, int(own_first), int(own_end)))

    def comm_set_partition(self, own_first, own_end, first_col_global, NX_global, x_offset, lenx_global):
        """comm_set_strip plus the strip's position in the global mesh (needed by solve_begin / load_increment on a rank)."""
        self._ck(self.L.plf_comm_set_partition(self.h, int(own_first), int(own_end), int(first_col_global), int(NX_global),
                                               C.c_double(x_offset), C.c_double(lenx_global)))

    def comm_set_replica(self, global_ctx):
        """Replicated solve: `global_ctx` (a Context with the GLOBAL mesh on the same device, or None) solves the system
        gathered from all ranks; see plf_comm_set_replica."""
        self._replica = global_ctx
        self._ck(self.L.plf_comm_set_replica(self.h, None if global_ctx is None else global_ctx.h))

    def comm_p2p(self):
        """True when halo columns and all-reduce scalars travel through CUDA-IPC peer mailboxes (one kernel per exchange over
        NVLink) rather than NCCL; decided collectively in comm_set_strip."""
        return bool(self.L.plf_comm_p2p(self.h))

    def comm_allreduce(self, vals):
        """Sum of a small float vector over all ranks."""
        a = np.array(vals, dtype=np.float64).ravel()
        self._ck(self.L.plf_comm_allreduce(self.h, _ptr(a), len(a)))
        return a

    def comm_destroy(self):
        self._ck(self.L.plf_comm_destroy(self.h))


COMM_ID_BYTES = 128


def comm_unique_id():
    """ncclGetUniqueId of the calling process (rank 0); hand the bytes to every rank by any side channel."""
    L = load()
    buf = (C.c_ubyte * COMM_ID_BYTES)()
    if L.plf_comm_unique_id(buf) != PLF_OK:
        raise PlfError('plf_comm_unique_id failed: NCCL (libnccl.so.2) could not be loaded')
    return bytes(buf)
