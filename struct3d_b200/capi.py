"""ctypes binding of libs3d_cuda.so (include/s3d_cuda.h).  Thin by design: one Python method per
C entry point, numpy arrays in the reference's host layouts at the boundary, raw device pointers
(ints) in between.  Fails loudly when the library is missing -- there is no fallback path."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libs3d_cuda.so")

SPMV_AUTO, SPMV_MARCH, SPMV_TMA, SPMV_NAIVE, SPMV_PAIR = 0, 1, 2, 3, 4
SGS_LEXICOGRAPHIC, SGS_REDBLACK, SGS_LEXICOGRAPHIC_PLANES, SGS_JACOBI_SWEEPS, SGS_SLABLOCAL, SGS_ASYNC, SGS_OVERLAP, SGS_BLOCKS = 0, 1, 2, 3, 4, 5, 6, 7
SGS_VARIANT_DEFAULT = 4   # warp-per-line streaming kernel (sgs_stream.cu); 0 warp-per-tile, 1 block-per-tile, 3 r1's line experiment
CONV_STRICT, CONV_INITIAL = 1, 2
COMM_ID_BYTES = 128

_dp = C.POINTER(C.c_double)
_vp = C.c_void_p


class Grid(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32), ("vlead", C.c_int32),
                ("vpitch", C.c_int32), ("cpitch", C.c_int32), ("vplane", C.c_int64), ("voff", C.c_int64),
                ("vlen", C.c_int64), ("cplane", C.c_int64), ("cstride", C.c_int64), ("nz_global", C.c_int32),
                ("k0", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32)]


class Scalar(C.Structure):
    _fields_ = [("imm", C.c_double), ("num", _vp), ("den", _vp)]


def scalar(imm=1.0, num=None, den=None):
    return Scalar(float(imm), num, den)


# name -> argtypes (restype is int unless listed in _RESTYPES); also the list of symbols the header declares
_P, _G, _I, _D = _vp, C.POINTER(Grid), C.c_int, C.c_double
_SIGS = {
    "s3d_ctx_create": [_I, _vp, C.POINTER(_vp)], "s3d_ctx_destroy": [_P], "s3d_ctx_sync": [_P],
    "s3d_ctx_device": [_P], "s3d_ctx_stream": [_P], "s3d_error_string": [_P], "s3d_ctx_launch_count": [_P],
    "s3d_timer_start": [_P], "s3d_timer_stop_ms": [_P, _dp], "s3d_tune_set": [_P, C.c_char_p, _I],
    "s3d_grid_init": [_G, _I, _I, _I], "s3d_grid_init_slab": [_G, _I, _I, _I, _I, _I],
    "s3d_vec_alloc": [_P, _G, C.POINTER(_vp)], "s3d_mat_alloc": [_P, _G, C.POINTER(_vp)],
    "s3d_scalars_alloc": [_P, _I, C.POINTER(_vp)], "s3d_free": [_P, _vp],
    "s3d_coef_alloc": [_P, _G, C.POINTER(_vp)], "s3d_coef_slack_upload": [_P, _G, _vp, _I, _I, _vp], "s3d_coef_halo_exchange": [_P, _G, _vp, _I],
    "s3d_sgs_overlap_setup": [_P, _G, _vp, _vp],
    "s3d_sgs_blocks_plan": [_I, _I, _I, _I, _I, _I, _I, _I, _vp, _vp, _vp],
    "s3d_vec_upload": [_P, _G, _vp, _vp], "s3d_vec_download": [_P, _G, _vp, _vp],
    "s3d_mat_upload": [_P, _G, _vp, _I, _vp], "s3d_mat_download": [_P, _G, _vp, _I, _vp],
    "s3d_scalars_download": [_P, _vp, _I, _vp], "s3d_scalars_upload": [_P, _vp, _I, _vp],
    "s3d_host_alloc": [_P, C.c_size_t, C.POINTER(_vp)], "s3d_host_free": [_P, _vp],
    "s3d_spmv": [_P, _G, _vp, _vp, _vp, _vp, _vp, _I, _vp, _vp, _I],
    "s3d_spmv_exchange": [_P, _G, _vp, _vp, _vp, _vp, _vp, _I, _vp, _vp],
    "s3d_spmv_host": [_P, _G, _vp, _vp, _vp, _vp, _vp, _vp, _I],
    "s3d_vecset": [_P, _G, _D, _vp], "s3d_vecassign": [_P, _G, _vp, _vp], "s3d_veccopy_all": [_P, _G, _vp, _vp],
    "s3d_vecaxpy": [_P, _G, Scalar, _vp, _vp], "s3d_vecaxpby": [_P, _G, Scalar, _vp, Scalar, _vp],
    "s3d_vecwaxpby": [_P, _G, Scalar, _vp, Scalar, _vp, _vp, _vp],
    "s3d_vec_multi_axpy": [_P, _G, _I, _vp, C.POINTER(_vp), _vp],
    "s3d_vec_multi_dot": [_P, _G, _I, C.POINTER(_vp), _vp, _vp],
    "s3d_dot": [_P, _G, _vp, _vp, _vp], "s3d_nrm2sq": [_P, _G, _vp, _vp],
    "s3d_normL2sq": [_P, _G, _vp, _vp, _vp, _vp, _vp],
    "s3d_dot_host": [_P, _G, _vp, _vp, _dp], "s3d_nrm2_host": [_P, _G, _vp, _dp],
    "s3d_allreduce_sum": [_P, _vp, _vp, _I],
    "s3d_jacobi_apply": [_P, _G, _vp, _vp, _vp], "s3d_sgs_setup": [_P, _G, _vp, _vp], "s3d_strilu_setup": [_P, _G, _vp, _I, _vp], "s3d_strilu_setup_async": [_P, _G, _vp, _I, _vp],
    "s3d_sgs_rb_pack_size": [_G, C.POINTER(C.c_size_t)], "s3d_sgs_rb_pack_alloc": [_P, _G, C.POINTER(_vp)], "s3d_sgs_rb_pack": [_P, _G, _vp, _vp, _vp], "s3d_sgs_apply_rb_packed": [_P, _G, _vp, _vp, _vp, _vp],
    "s3d_sgs_apply": [_P, _G, _vp, _vp, _vp, _vp, _vp, _I, _I], "s3d_sgs_slab_sweeps_available": [_P, _G, C.POINTER(C.c_int)],
    "s3d_cg_update": [_P, _G, Scalar, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "s3d_cg_direction_jacobi": [_P, _G, Scalar, _vp, _vp, _vp, _vp],
    "s3d_bicg_direction": [_P, _G, Scalar, Scalar, _vp, _vp, _vp, _vp],
    "s3d_bicg_update": [_P, _G, Scalar, _vp, Scalar, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "s3d_converge_test": [_P, _vp, _vp, _D, _I, _I, _vp, _vp, _vp],
    "s3d_scalar_op": [_P, _I, _vp, _vp, _vp, _vp, _vp, _I, _D, _vp],
    "s3d_small_solve": [_P, _G, _vp, _vp, _vp, _vp, _vp, _vp, _I, _I, _D, _I, _vp, _vp, _vp, _I],
    "s3d_small_solve_gcr": [_P, _G, _vp, _vp, _vp, _vp, _vp, C.POINTER(_vp), _I, _I, _D, _I, _vp, _vp, _vp, _I],
    "s3d_ctx_set_gate": [_P, _vp], "s3d_phase_begin": [_P, _I], "s3d_phase_end": [_P, _I], "s3d_phase_total_ms": [_P, _I, _dp],
    "s3d_graph_begin": [_P], "s3d_graph_end": [_P, C.POINTER(_vp)], "s3d_graph_launch": [_P, _vp], "s3d_graph_destroy": [_P, _vp],
    "s3d_ints_snapshot": [_P, _vp, _I, _I], "s3d_ints_snapshot_wait": [_P, _I, _I, _vp],
    "s3d_ints_alloc": [_P, _I, C.POINTER(_vp)], "s3d_ints_download": [_P, _vp, _I, _vp], "s3d_ints_zero": [_P, _vp, _I],
    "s3d_assemble_poisson": [_P, _G, _vp, _vp, _vp, _vp],
    "s3d_assemble_convdiff": [_P, _G, _vp, _vp, _vp, _vp, _D, _vp],
    "s3d_assemble_convdiff_circ": [_P, _G, _vp, _vp, _vp, _D, _vp],
    "s3d_assemble_rhs_separable": [_P, _G, _I, _vp, _vp, _vp, _vp, _vp],
    "s3d_comm_unique_id": [_vp], "s3d_comm_init": [_P, _I, _I, _vp], "s3d_comm_destroy": [_P],
    "s3d_halo_exchange": [_P, _G, _vp],
}
_RESTYPES = {"s3d_ctx_stream": _vp, "s3d_error_string": C.c_char_p, "s3d_ctx_launch_count": C.c_int64}
EXPORTS = sorted(_SIGS)

_LIB = None


def lib():
    """Loads libs3d_cuda.so (no GPU needed to load; a GPU is needed to create a context)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(struct3d_b200 has no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, C.c_int)
        _LIB = L
    return _LIB


def _ptr(a):
    """device pointer (int / c_void_p / None) or numpy array -> c_void_p"""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(_vp)
    if isinstance(a, _vp):
        return a
    return _vp(int(a))


def make_grid(nx, ny, nz, rank=0, nranks=1):
    g = Grid()
    if nranks == 1:
        rc = lib().s3d_grid_init(C.byref(g), nx, ny, nz)
    else:
        rc = lib().s3d_grid_init_slab(C.byref(g), nx, ny, nz, rank, nranks)
    if rc != 0:
        raise ValueError(f"s3d_grid_init failed for {(nx, ny, nz)} rank {rank}/{nranks}")
    return g


def sgs_blocks_plan(nx, ny, nz, blocks, blocks_y, tile_i=16, pencils=2, direction=0):
    """host-side plan of the blocks ordering (no device): ((ntx, nty, ntz), k layers [ntz][4], j layers [nty][4], order [ntiles])"""
    dims = np.zeros(3, dtype=np.int32)
    rc = lib().s3d_sgs_blocks_plan(nx, ny, nz, blocks, blocks_y, tile_i, pencils, direction, dims.ctypes.data_as(_vp), None, None)
    if rc != 0:
        raise ValueError("s3d_sgs_blocks_plan: bad argument")
    ntx, nty, ntz = (int(v) for v in dims)
    layers = np.zeros((ntz + nty, 4), dtype=np.int32)
    order = np.zeros(ntx * nty * ntz, dtype=np.int32)
    lib().s3d_sgs_blocks_plan(nx, ny, nz, blocks, blocks_y, tile_i, pencils, direction, dims.ctypes.data_as(_vp), layers.ctypes.data_as(_vp), order.ctypes.data_as(_vp))
    return (ntx, nty, ntz), layers[:ntz], layers[ntz:], order


class S3dError(RuntimeError):
    pass


class Context:
    """One GPU context.  Methods map 1:1 onto the C ABI; vectors/matrices are raw device pointers."""

    def __init__(self, device=-1, stream=None):
        self.L = lib()
        h = _vp()
        rc = self.L.s3d_ctx_create(device, stream, C.byref(h))
        if rc != 0:
            raise S3dError(f"s3d_ctx_create failed ({rc}): {self.L.s3d_error_string(None).decode()}")
        self.h = h

    @classmethod
    def from_handle(cls, handle):
        """Wraps an existing s3d_ctx* (e.g. the host layer's, hostapi.cuda_ctx()) without owning it."""
        self = cls.__new__(cls)
        self.L = lib()
        self.h = handle if isinstance(handle, _vp) else _vp(handle)
        self.borrowed = True
        return self

    def close(self):
        if getattr(self, "h", None):
            if not getattr(self, "borrowed", False):
                self.L.s3d_ctx_destroy(self.h)
            self.h = None

    def host_alloc(self, nbytes):
        p = _vp()
        self.call("s3d_host_alloc", C.c_size_t(nbytes), C.byref(p))
        return p.value

    def host_free(self, p):
        self.call("s3d_host_free", _ptr(p))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise S3dError(f"libs3d_cuda error {rc}: {self.L.s3d_error_string(self.h).decode()}")

    def call(self, name, *args):
        self._ck(getattr(self.L, name)(self.h, *args))

    # --- storage
    def vec_alloc(self, g):
        p = _vp()
        self.call("s3d_vec_alloc", C.byref(g), C.byref(p))
        return p.value

    def mat_alloc(self, g):
        p = _vp()
        self.call("s3d_mat_alloc", C.byref(g), C.byref(p))
        return p.value

    def coef_alloc(self, g):
        """one array in the coefficient layout (z-slab grids: with slack planes)"""
        p = _vp()
        self.call("s3d_coef_alloc", C.byref(g), C.byref(p))
        return p.value

    def coef_slack_upload(self, g, dC, stencil, side, h_plane):
        h = np.ascontiguousarray(h_plane, dtype=np.float64)
        assert h.size == g.nx * g.ny
        self.call("s3d_coef_slack_upload", C.byref(g), _ptr(dC), int(stencil), int(side), h.ctypes.data_as(_vp))

    def sgs_overlap_setup(self, g, A, dinv):
        self.call("s3d_sgs_overlap_setup", C.byref(g), _ptr(A), _ptr(dinv))

    def scalars_alloc(self, n):
        p = _vp()
        self.call("s3d_scalars_alloc", n, C.byref(p))
        return p.value

    def ints_alloc(self, n):
        p = _vp()
        self.call("s3d_ints_alloc", n, C.byref(p))
        return p.value

    def free(self, p):
        self.call("s3d_free", _ptr(p))

    def vec_upload(self, g, d, h):
        assert h.dtype == np.float64 and h.flags["C_CONTIGUOUS"] and h.size == (g.nx + 2) * (g.ny + 2) * (g.nz + 2)
        self.call("s3d_vec_upload", C.byref(g), _ptr(d), _ptr(h))
        self.sync()

    def vec_download(self, g, d):
        h = np.empty((g.nz + 2, g.ny + 2, g.nx + 2))
        self.call("s3d_vec_download", C.byref(g), _ptr(d), _ptr(h))
        return h

    def vec_from(self, g, h):
        d = self.vec_alloc(g)
        self.vec_upload(g, d, np.ascontiguousarray(h, dtype=np.float64))
        return d

    def mat_upload(self, g, d, A):
        """A: (7, nz, ny, nx) float64"""
        A = np.ascontiguousarray(A, dtype=np.float64)
        assert A.shape == (7, g.nz, g.ny, g.nx)
        for s in range(7):
            self.call("s3d_mat_upload", C.byref(g), _ptr(d), s, _ptr(A[s]))
        self.sync()

    def mat_from(self, g, A):
        d = self.mat_alloc(g)
        self.mat_upload(g, d, A)
        return d

    def mat_download(self, g, d):
        A = np.empty((7, g.nz, g.ny, g.nx))
        for s in range(7):
            self.call("s3d_mat_download", C.byref(g), _ptr(d), s, _ptr(A[s]))
        return A

    def scalars(self, d, n=1):
        h = np.empty(n)
        self.call("s3d_scalars_download", _ptr(d), n, _ptr(h))
        return h

    def scalars_set(self, d, vals):
        h = np.ascontiguousarray(vals, dtype=np.float64)
        self.call("s3d_scalars_upload", _ptr(d), h.size, _ptr(h))

    def ints(self, d, n=1):
        h = np.empty(n, dtype=np.int32)
        self.call("s3d_ints_download", _ptr(d), n, _ptr(h))
        return h

    def sync(self):
        self.call("s3d_ctx_sync")

    def launch_count(self):
        return int(self.L.s3d_ctx_launch_count(self.h))

    def timer_start(self):
        self.call("s3d_timer_start")

    def timer_stop_ms(self):
        ms = C.c_double()
        self.call("s3d_timer_stop_ms", C.byref(ms))
        return ms.value

    def tune(self, key, value):
        self.call("s3d_tune_set", key.encode(), int(value))

    # --- CUDA graphs: record the calls between graph_begin and graph_end, replay them with graph_launch
    def graph_begin(self):
        self.call("s3d_graph_begin")

    def graph_end(self):
        h = _vp()
        self.call("s3d_graph_end", C.byref(h))
        return h.value

    def graph_launch(self, h):
        self.call("s3d_graph_launch", _vp(h))

    def graph_destroy(self, h):
        self.call("s3d_graph_destroy", _vp(h))

    # --- kernels
    def spmv(self, g, A, x, y, b=None, w=None, want_yy=False, red=None, stop=None, variant=SPMV_AUTO):
        self.call("s3d_spmv", C.byref(g), _ptr(A), _ptr(x), _ptr(y), _ptr(b), _ptr(w), int(want_yy), _ptr(red), _ptr(stop), variant)

    def spmv_host(self, g, A, h_x, h_y, d_x, d_y, d_b=None, nchunks=0):
        """h_x / h_y: numpy arrays or raw host pointers in the SVec layout"""
        self.call("s3d_spmv_host", C.byref(g), _ptr(A), _ptr(h_x), _ptr(h_y), _ptr(d_x), _ptr(d_y), _ptr(d_b), int(nchunks))

    def vecset(self, g, a, x):
        self.call("s3d_vecset", C.byref(g), float(a), _ptr(x))

    def vecassign(self, g, x, y):
        self.call("s3d_vecassign", C.byref(g), _ptr(x), _ptr(y))

    def veccopy_all(self, g, x, y):
        self.call("s3d_veccopy_all", C.byref(g), _ptr(x), _ptr(y))

    def vecaxpy(self, g, a, x, y):
        self.call("s3d_vecaxpy", C.byref(g), a if isinstance(a, Scalar) else scalar(a), _ptr(x), _ptr(y))

    def vecaxpby(self, g, a, x, b, y):
        self.call("s3d_vecaxpby", C.byref(g), a if isinstance(a, Scalar) else scalar(a), _ptr(x),
                  b if isinstance(b, Scalar) else scalar(b), _ptr(y))

    def vecwaxpby(self, g, a, x, b, y, w, red=None):
        self.call("s3d_vecwaxpby", C.byref(g), a if isinstance(a, Scalar) else scalar(a), _ptr(x),
                  b if isinstance(b, Scalar) else scalar(b), _ptr(y), _ptr(w), _ptr(red))

    def vec_multi_axpy(self, g, d_a, xs, y):
        n = len(xs)
        self.call("s3d_vec_multi_axpy", C.byref(g), n, _ptr(d_a), (_vp * n)(*[int(x) for x in xs]), _ptr(y))

    def vec_multi_dot(self, g, xs, y, red):
        n = len(xs)
        self.call("s3d_vec_multi_dot", C.byref(g), n, (_vp * n)(*[int(x) for x in xs]), _ptr(y), _ptr(red))

    def dot(self, g, x, y):
        out = C.c_double()
        self.call("s3d_dot_host", C.byref(g), _ptr(x), _ptr(y), C.byref(out))
        return out.value

    def nrm2(self, g, x):
        out = C.c_double()
        self.call("s3d_nrm2_host", C.byref(g), _ptr(x), C.byref(out))
        return out.value

    def dot_dev(self, g, x, y, red):
        self.call("s3d_dot", C.byref(g), _ptr(x), _ptr(y), _ptr(red))

    def nrm2sq_dev(self, g, x, red):
        self.call("s3d_nrm2sq", C.byref(g), _ptr(x), _ptr(red))

    def normL2sq_dev(self, g, coords, x, red):
        cs = [np.ascontiguousarray(c, dtype=np.float64) for c in coords]
        self.call("s3d_normL2sq", C.byref(g), _ptr(cs[0]), _ptr(cs[1]), _ptr(cs[2]), _ptr(x), _ptr(red))

    def allreduce_sum(self, d_local, d_global, n):
        self.call("s3d_allreduce_sum", _ptr(d_local), _ptr(d_global), n)

    def jacobi_apply(self, g, A, r, z):
        self.call("s3d_jacobi_apply", C.byref(g), _ptr(A), _ptr(r), _ptr(z))

    def sgs_setup(self, g, A, dinv):
        self.call("s3d_sgs_setup", C.byref(g), _ptr(A), _ptr(dinv))

    def strilu_setup(self, g, A, nbuildsweeps, dinv):
        self.call("s3d_strilu_setup", C.byref(g), _ptr(A), int(nbuildsweeps), _ptr(dinv))

    def sgs_rb_pack(self, g, A, dinv):
        """colour-packed copy of the coefficients for the red-black SGS; returns the device buffer (free it with ctx.free)"""
        p = _vp()
        self.call("s3d_sgs_rb_pack_alloc", C.byref(g), C.byref(p))
        self.call("s3d_sgs_rb_pack", C.byref(g), _ptr(A), _ptr(dinv), _ptr(p.value))
        return p.value

    def sgs_apply_rb_packed(self, g, packed, r, z, y):
        self.call("s3d_sgs_apply_rb_packed", C.byref(g), _ptr(packed), _ptr(r), _ptr(z), _ptr(y))

    def strilu_setup_async(self, g, A, nbuildsweeps, dinv):
        self.call("s3d_strilu_setup_async", C.byref(g), _ptr(A), int(nbuildsweeps), _ptr(dinv))

    def sgs_apply(self, g, A, dinv, r, z, y, ordering=SGS_LEXICOGRAPHIC, nsweeps=1):
        self.call("s3d_sgs_apply", C.byref(g), _ptr(A), _ptr(dinv), _ptr(r), _ptr(z), _ptr(y), ordering, nsweeps)

    def cg_update(self, g, alpha, p, q, x, r, A, red, stop=None):
        self.call("s3d_cg_update", C.byref(g), alpha, _ptr(p), _ptr(q), _ptr(x), _ptr(r), _ptr(A), _ptr(red), _ptr(stop))

    def cg_direction_jacobi(self, g, beta, A, r, p, stop=None):
        self.call("s3d_cg_direction_jacobi", C.byref(g), beta, _ptr(A), _ptr(r), _ptr(p), _ptr(stop))

    def bicg_direction(self, g, beta, omega, r, v, p, stop=None):
        self.call("s3d_bicg_direction", C.byref(g), beta, omega, _ptr(r), _ptr(v), _ptr(p), _ptr(stop))

    def bicg_update(self, g, alpha, ph, omega, sh, s, t, rhat, x, r, red, stop=None):
        self.call("s3d_bicg_update", C.byref(g), alpha, _ptr(ph), omega, _ptr(sh), _ptr(s), _ptr(t), _ptr(rhat),
                  _ptr(x), _ptr(r), _ptr(red), _ptr(stop))

    def converge_test(self, rr, bb, rtol, maxiter, flags, stop, iters, hist=None):
        self.call("s3d_converge_test", _ptr(rr), _ptr(bb), float(rtol), int(maxiter), int(flags), _ptr(stop), _ptr(iters), _ptr(hist))

    def scalar_op(self, op, out, a, b, c=None, d=None, n=1, scale=1.0, stop=None):
        self.call("s3d_scalar_op", op, _ptr(out), _ptr(a), _ptr(b), _ptr(c), _ptr(d), n, float(scale), _ptr(stop))

    def ints_zero(self, d, n):
        self.call("s3d_ints_zero", _ptr(d), n)

    def assemble_poisson(self, g, coords, A):
        cs = [np.ascontiguousarray(c, dtype=np.float64) for c in coords]
        self.call("s3d_assemble_poisson", C.byref(g), _ptr(cs[0]), _ptr(cs[1]), _ptr(cs[2]), _ptr(A))

    def assemble_convdiff(self, g, coords, vel, mu, A):
        cs = [np.ascontiguousarray(c, dtype=np.float64) for c in coords]
        v = np.ascontiguousarray(vel, dtype=np.float64)
        self.call("s3d_assemble_convdiff", C.byref(g), _ptr(cs[0]), _ptr(cs[1]), _ptr(cs[2]), _ptr(v), float(mu), _ptr(A))

    def assemble_convdiff_circ(self, g, coords, mu, A):
        cs = [np.ascontiguousarray(c, dtype=np.float64) for c in coords]
        self.call("s3d_assemble_convdiff_circ", C.byref(g), _ptr(cs[0]), _ptr(cs[1]), _ptr(cs[2]), float(mu), _ptr(A))

    def assemble_rhs_separable(self, g, tx, ty, tz, f, face=None):
        tx, ty, tz = [np.ascontiguousarray(t, dtype=np.float64) for t in (tx, ty, tz)]
        fc = None if face is None else np.ascontiguousarray(face, dtype=np.float64)
        self.call("s3d_assemble_rhs_separable", C.byref(g), tx.shape[0], _ptr(tx), _ptr(ty), _ptr(tz), _ptr(fc), _ptr(f))

    # --- multi-GPU
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(COMM_ID_BYTES)
        if lib().s3d_comm_unique_id(buf) != 0:
            raise S3dError("s3d_comm_unique_id failed")
        return buf.raw

    def comm_init(self, rank, nranks, id_bytes):
        self.call("s3d_comm_init", rank, nranks, C.c_char_p(id_bytes) if id_bytes else None)

    def halo_exchange(self, g, x):
        self.call("s3d_halo_exchange", C.byref(g), _ptr(x))
