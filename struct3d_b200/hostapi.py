"""ctypes binding of libstruct3d_host.so (include/s3d_host.h): the C++ host classes that mirror the
reference interface (SVec, SMat, createSolver, PDEBase ...), device-resident on the B200.

The class and function names deliberately match oracle/refshim.py (the binding of the REAL
reference), so a parity test is the same code run against both.  No fallback: loading needs the
built libraries, initialising needs a GPU."""
import ctypes as C
import os
import re
import tempfile

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libstruct3d_host.so")
_dp = C.POINTER(C.c_double)
_vp = C.c_void_p
_LIB = None

_PTR_FUNCS = ("s3dh_cuda_ctx", "s3dh_mesh_create", "s3dh_pde_create", "s3dh_pde_vector", "s3dh_pde_vector_sampled",
              "s3dh_pde_lhs", "s3dh_vec_create", "s3dh_vec_clone", "s3dh_vec_device_ptr", "s3dh_mat_create",
              "s3dh_mat_device_ptr", "s3dh_prec_create", "s3dh_solver_create")
EXPORTS = sorted(set(_PTR_FUNCS) | {
    "s3dh_last_error", "s3dh_init", "s3dh_initialized", "s3dh_shutdown", "s3dh_options_clear", "s3dh_options_set",
    "s3dh_options_load", "s3dh_mesh_destroy", "s3dh_mesh_coords", "s3dh_mesh_h", "s3dh_mesh_local", "s3dh_read_control",
    "s3dh_pde_destroy", "s3dh_vec_destroy", "s3dh_vec_len", "s3dh_vec_get", "s3dh_vec_set", "s3dh_vec_peek", "s3dh_vec_poke",
    "s3dh_mat_destroy", "s3dh_mat_len", "s3dh_mat_get", "s3dh_mat_set", "s3dh_apply", "s3dh_apply_res", "s3dh_vecset",
    "s3dh_vecassign", "s3dh_vecaxpy", "s3dh_vec_multi_axpy", "s3dh_norm_L2", "s3dh_norm_vector_l2", "s3dh_inner_vector_l2",
    "s3dh_compute_error_L2", "s3dh_prec_destroy", "s3dh_prec_apply", "s3dh_solver_update", "s3dh_solver_apply",
    "s3dh_solver_history", "s3dh_solver_destroy"})


def lib():
    global _LIB
    if _LIB is None:
        capi.lib()  # libs3d_cuda.so first: the host library's DT_NEEDED entry then resolves to the same instance
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run __graft_entry__.build() (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        for f in _PTR_FUNCS:
            getattr(L, f).restype = _vp
        L.s3dh_last_error.restype = C.c_char_p
        L.s3dh_vec_len.restype = C.c_longlong
        L.s3dh_mat_len.restype = C.c_longlong
        L.s3dh_mesh_h.restype = C.c_double
        L.s3dh_vec_peek.restype = C.c_double
        L.s3dh_vec_peek.argtypes = [_vp, C.c_longlong]
        L.s3dh_vec_poke.argtypes = [_vp, C.c_longlong, C.c_double]
        L.s3dh_vecset.argtypes = [C.c_double, _vp]
        L.s3dh_vecaxpy.argtypes = [C.c_double, _vp, _vp]
        _LIB = L
    return _LIB


class HostError(RuntimeError):
    pass


def _err():
    return HostError(lib().s3dh_last_error().decode())


def _chk(rc):
    if rc != 0:
        raise _err()


def _a(vals, t=C.c_double):
    return (t * len(vals))(*vals)


def init(device=-1, rank=0, nranks=1, nccl_id=None):
    """One process per GPU.  nranks > 1: nccl_id = capi.Context.comm_unique_id() of rank 0, broadcast by the caller."""
    if lib().s3dh_initialized():
        return
    _chk(lib().s3dh_init(device, rank, nranks, nccl_id))


def shutdown():
    lib().s3dh_shutdown()


def cuda_ctx():
    return _vp(lib().s3dh_cuda_ctx())


def launch_count():
    return int(capi.lib().s3d_ctx_launch_count(cuda_ctx()))


def sync():
    capi.lib().s3d_ctx_sync(cuda_ctx())


class Mesh:
    def __init__(self, npdim, rmin=(-1.0, -1.0, -1.0), rmax=(1.0, 1.0, 1.0), grid="uniform"):
        self.npdim = tuple(int(v) for v in npdim)
        self.h_ = _vp(lib().s3dh_mesh_create(_a(self.npdim, C.c_int), int(grid == "chebyshev"), _a(rmin), _a(rmax)))
        if not self.h_.value:
            raise _err()
        nx, ny, nz, k0 = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        lib().s3dh_mesh_local(self.h_, C.byref(nx), C.byref(ny), C.byref(nz), C.byref(k0))
        self.local = (nx.value, ny.value, nz.value)
        self.k0 = k0.value

    @property
    def vshape(self):   # this rank's slab, ghosted
        return (self.local[2] + 2, self.local[1] + 2, self.local[0] + 2)

    @property
    def mshape(self):
        return (self.local[2], self.local[1], self.local[0])

    def coords(self, d):
        c = np.zeros(self.npdim[d])
        lib().s3dh_mesh_coords(self.h_, d, c.ctypes.data_as(_dp))
        return c

    def h(self):
        return lib().s3dh_mesh_h(self.h_)

    def __del__(self):
        try:
            lib().s3dh_mesh_destroy(self.h_)
        except Exception:
            pass


class Vec:
    """Owns an SVec (device-resident).  `.a` downloads a numpy copy in the reference's ghosted layout."""

    def __init__(self, mesh, handle=None):
        self.mesh = mesh
        self.h_ = _vp(handle if handle is not None else lib().s3dh_vec_create(mesh.h_))
        if not self.h_.value:
            raise _err()

    @property
    def a(self):
        out = np.empty(self.mesh.vshape)
        _chk(lib().s3dh_vec_get(self.h_, out.ctypes.data_as(_dp)))
        return out

    def set(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.shape == self.mesh.vshape
        _chk(lib().s3dh_vec_set(self.h_, arr.ctypes.data_as(_dp)))

    def clone(self):
        return Vec(self.mesh, lib().s3dh_vec_clone(self.h_))

    def peek(self, idx):
        return lib().s3dh_vec_peek(self.h_, idx)

    def poke(self, idx, v):
        _chk(lib().s3dh_vec_poke(self.h_, idx, v))

    def device_ptr(self):
        return lib().s3dh_vec_device_ptr(self.h_)

    def __del__(self):
        try:
            lib().s3dh_vec_destroy(self.h_)
        except Exception:
            pass


class Mat:
    def __init__(self, mesh, handle=None):
        self.mesh = mesh
        self.h_ = _vp(handle if handle is not None else lib().s3dh_mat_create(mesh.h_))
        if not self.h_.value:
            raise _err()

    def stacked(self):
        out = np.empty((7,) + self.mesh.mshape)
        for s in range(7):
            _chk(lib().s3dh_mat_get(self.h_, s, out[s].ctypes.data_as(_dp)))
        return out

    def set(self, A):
        A = np.ascontiguousarray(A, dtype=np.float64)
        for s in range(7):
            _chk(lib().s3dh_mat_set(self.h_, s, A[s].ctypes.data_as(_dp)))

    def device_ptr(self):
        return lib().s3dh_mat_device_ptr(self.h_)

    def __del__(self):
        try:
            lib().s3dh_mat_destroy(self.h_)
        except Exception:
            pass


_WHICH = {"test_rhs": 0, "manufactured_solution": 1, "manufactured_rhs": 2}


class Pde:
    def __init__(self, pde="poisson", vel=(0, 0, 0), mu=0.0, dirichlet=(1,) * 6, bvals=(0.0,) * 6):
        self.h_ = _vp(lib().s3dh_pde_create(pde.encode(), _a(vel), C.c_double(mu), _a(dirichlet, C.c_int), _a(bvals)))
        if not self.h_.value:
            raise _err()

    def vector(self, mesh, which="test_rhs", sampled=False):
        fn = lib().s3dh_pde_vector_sampled if sampled else lib().s3dh_pde_vector
        h = fn(self.h_, mesh.h_, _WHICH[which])
        if not h:
            raise _err()
        return Vec(mesh, h)

    def lhs(self, mesh):
        h = lib().s3dh_pde_lhs(self.h_, mesh.h_)
        if not h:
            raise _err()
        return Mat(mesh, h)

    def __del__(self):
        try:
            lib().s3dh_pde_destroy(self.h_)
        except Exception:
            pass


def apply(A, x, y):
    _chk(lib().s3dh_apply(A.h_, x.h_, y.h_))


def apply_res(A, b, x, y):
    _chk(lib().s3dh_apply_res(A.h_, b.h_, x.h_, y.h_))


def vecaxpy(a, x, y):
    _chk(lib().s3dh_vecaxpy(a, x.h_, y.h_))


def vecset(a, x):
    _chk(lib().s3dh_vecset(a, x.h_))


def vecassign(x, y):
    _chk(lib().s3dh_vecassign(x.h_, y.h_))


def vec_multi_axpy(a, xs, y):
    n = len(xs)
    _chk(lib().s3dh_vec_multi_axpy(n, _a(a), (_vp * n)(*[x.h_ for x in xs]), y.h_))


def _scalar(fn, *args):
    out = C.c_double()
    _chk(fn(*args, C.byref(out)))
    return out.value


def norm_L2(x):
    return _scalar(lib().s3dh_norm_L2, x.h_)


def norm_vector_l2(x):
    return _scalar(lib().s3dh_norm_vector_l2, x.h_)


def inner_vector_l2(x, y):
    return _scalar(lib().s3dh_inner_vector_l2, x.h_, y.h_)


def compute_error_L2(x, y):
    return _scalar(lib().s3dh_compute_error_L2, x.h_, y.h_)


class Prec:
    def __init__(self, A, pc, apply_sweeps=1, build_sweeps=1, threaded_apply=False, threaded_build=False, chunk=1024):
        self.A = A
        self.h_ = _vp(lib().s3dh_prec_create(A.h_, pc.encode(), apply_sweeps, build_sweeps, int(threaded_apply), int(threaded_build), chunk))
        if not self.h_.value:
            raise _err()

    def apply(self, r, z):
        _chk(lib().s3dh_prec_apply(self.h_, r.h_, z.h_))

    def __del__(self):
        try:
            lib().s3dh_prec_destroy(self.h_)
        except Exception:
            pass


def set_options(opts):
    lib().s3dh_options_clear()
    for k, v in opts.items():
        lib().s3dh_options_set(k.encode(), str(v).encode())


def load_options(path):
    _chk(lib().s3dh_options_load(path.encode()))


def _capture_stdout(fn):
    import sys
    sys.stdout.flush()
    with tempfile.TemporaryFile(mode="w+b") as tmp:
        saved = os.dup(1)
        os.dup2(tmp.fileno(), 1)
        try:
            res = fn()
        finally:
            C.CDLL(None).fflush(None)
            os.dup2(saved, 1)
            os.close(saved)
        tmp.seek(0)
        return res, tmp.read().decode(errors="replace")


class Solver:
    """createSolver(A) from the options store; the caller must update() before apply(), as with the reference."""

    def __init__(self, A):
        self.A = A
        self.h_ = _vp(lib().s3dh_solver_create(A.h_))
        if not self.h_.value:
            raise _err()

    def update(self):
        _chk(lib().s3dh_solver_update(self.h_))

    def apply(self, b, x, capture=False):
        conv, its, rn, pct = C.c_int(), C.c_int(), C.c_double(), C.c_double()

        def run():
            return lib().s3dh_solver_apply(self.h_, b.h_, x.h_, C.byref(conv), C.byref(its), C.byref(rn), C.byref(pct))

        text = ""
        if capture:
            rc, text = _capture_stdout(run)
        else:
            rc = run()
        _chk(rc)
        n = lib().s3dh_solver_history(self.h_, None, 0)
        hist = np.zeros(max(n, 1))
        lib().s3dh_solver_history(self.h_, hist.ctypes.data_as(_dp), n)
        printed = {int(m.group(1)): float(m.group(2)) for m in re.finditer(r"Step (\d+): Rel res = (\S+)", text)}
        return {"converged": bool(conv.value), "iters": its.value, "resnorm": rn.value, "precapplywtime": pct.value,
                "hist": hist[:n].copy(), "printed": printed}

    def __del__(self):
        try:
            lib().s3dh_solver_destroy(self.h_)
        except Exception:
            pass


def solve(A, b, x, opts, capture=False):
    """The reference-facing call sequence: createSolver -> updateOperator -> apply."""
    set_options(opts)
    s = Solver(A)
    s.update()
    return s.apply(b, x, capture=capture)


def read_control(path):
    npd, rmin, rmax = (C.c_int * 3)(), (C.c_double * 3)(), (C.c_double * 3)()
    cheb, nruns, mu = C.c_int(), C.c_int(), C.c_double()
    pdet = C.create_string_buffer(32)
    vel, bt, bv = (C.c_double * 3)(), (C.c_int * 6)(), (C.c_double * 6)()
    _chk(lib().s3dh_read_control(path.encode(), npd, rmin, rmax, C.byref(cheb), pdet, vel, C.byref(mu), C.byref(nruns), bt, bv))
    return {"npdim": list(npd), "rmin": list(rmin), "rmax": list(rmax), "grid": "chebyshev" if cheb.value else "uniform",
            "pde": pdet.value.decode(), "vel": list(vel), "mu": mu.value, "nruns": nruns.value,
            "btypes": list(bt), "bvals": list(bv)}
