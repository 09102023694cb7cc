"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libs3dref.so -- the reference's own
sources compiled unchanged (oracle/Makefile) behind oracle/ref_shim.cpp.

`available()` is False where the library was never built (no /root/reference at build time and
no prebuilt .so shipped); callers must then fall back to oracle.cport and say so.
"""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libs3dref.so")
_LIB = None
_dp = C.POINTER(C.c_double)
_vp = C.c_void_p


def build(ref="/root/reference"):
    if os.path.isdir(ref):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref", f"REF={ref}"])
        return True
    return os.path.exists(_PATH)


def available():
    return os.path.exists(_PATH)


def use_perf_build():
    """bench.py's timing legs only: switch to the performance build of the same sources (oracle/Makefile PERFFLAGS: -O3, FMA
    contraction, forced SIMD; x86-64-v4 where the host has AVX-512, else v3).  Must be called before the library is first used.
    Returns a description of the build that will be timed."""
    global _PATH
    if _LIB is not None:
        raise RuntimeError("use_perf_build() must come before the first use of the reference library")
    flags = ""
    try:
        with open("/proc/cpuinfo") as fh:
            for line in fh:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    level = "v4" if all(f" {w}" in flags for w in ("avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl")) else "v3"
    for lv in (level, "v3"):
        p = os.path.join(_HERE, "_ref", f"libs3dref_perf_{lv}.so")
        if os.path.exists(p):
            _PATH = p
            return f"-O3 -march=x86-64-{lv} -fopenmp (reference's performance flags; built where the sources are)"
    return "-O3 -march=x86-64-v3 -ffp-contract=off -DNOFORCESIMD (parity build: no performance build found)"


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(_PATH)
        for f in ("refs_mesh_create", "refs_pde_create", "refs_vec_create", "refs_mat_create",
                  "refs_pde_vector", "refs_pde_lhs", "refs_prec_create"):
            getattr(L, f).restype = _vp
        L.refs_vec_data.restype = _dp
        L.refs_mat_data.restype = _dp
        L.refs_vec_len.restype = C.c_longlong
        L.refs_mat_len.restype = C.c_longlong
        L.refs_last_error.restype = C.c_char_p
        for f in ("refs_norm_L2", "refs_norm_vector_l2", "refs_inner_vector_l2", "refs_compute_error_L2",
                  "refs_mesh_h", "refs_time_apply", "refs_time_axpy", "refs_time_dot"):
            getattr(L, f).restype = C.c_double
        _LIB = L
    return _LIB


def _chk(rc):
    if rc != 0:
        raise RuntimeError(lib().refs_last_error().decode())


def _a(vals, t=C.c_double):
    return (t * len(vals))(*vals)


class Mesh:
    def __init__(self, npdim, rmin=(-1.0, -1.0, -1.0), rmax=(1.0, 1.0, 1.0), grid="uniform"):
        self.npdim = tuple(int(v) for v in npdim)
        self.h_ = _vp(lib().refs_mesh_create(_a(self.npdim, C.c_int), int(grid == "chebyshev"), _a(rmin), _a(rmax)))
        if not self.h_.value:
            raise RuntimeError(lib().refs_last_error().decode())

    @property
    def vshape(self):
        return (self.npdim[2], self.npdim[1], self.npdim[0])

    @property
    def mshape(self):
        return (self.npdim[2] - 2, self.npdim[1] - 2, self.npdim[0] - 2)

    def coords(self, d):
        c = np.zeros(self.npdim[d])
        lib().refs_mesh_coords(self.h_, d, c.ctypes.data_as(_dp))
        return c

    def h(self):
        return lib().refs_mesh_h(self.h_)


class Vec:
    """Owns a reference SVec; `.a` is a numpy view of its storage (ghosted, reference layout)."""

    def __init__(self, mesh, handle=None):
        self.mesh = mesh
        self.h_ = _vp(handle if handle is not None else lib().refs_vec_create(mesh.h_))
        n = lib().refs_vec_len(self.h_)
        # the view must keep this object (and with it the C++ SVec) alive: `P.vector(m, w).a` is a common pattern
        buf = (C.c_double * n).from_address(C.addressof(lib().refs_vec_data(self.h_).contents))
        buf._owner = self
        self.a = np.frombuffer(buf, dtype=np.float64).reshape(mesh.vshape)

    def __del__(self):
        try:
            lib().refs_vec_destroy(self.h_)
        except Exception:
            pass


class Mat:
    """Owns a reference SMat; `.a[s]` are numpy views of vals[s]."""

    def __init__(self, mesh, handle=None):
        self.mesh = mesh
        self.h_ = _vp(handle if handle is not None else lib().refs_mat_create(mesh.h_))
        n = lib().refs_mat_len(self.h_)
        self.a = []
        for s in range(7):
            buf = (C.c_double * n).from_address(C.addressof(lib().refs_mat_data(self.h_, s).contents))
            buf._owner = self
            self.a.append(np.frombuffer(buf, dtype=np.float64).reshape(mesh.mshape))

    def stacked(self):
        return np.stack(self.a)

    def __del__(self):
        try:
            lib().refs_mat_destroy(self.h_)
        except Exception:
            pass


class Pde:
    def __init__(self, pde="poisson", vel=(0, 0, 0), mu=0.0, dirichlet=(1,) * 6, bvals=(0.0,) * 6):
        self.h_ = _vp(lib().refs_pde_create(pde.encode(), _a(vel), C.c_double(mu), _a(dirichlet, C.c_int), _a(bvals)))

    def vector(self, mesh, which="test_rhs"):
        w = {"test_rhs": 0, "manufactured_solution": 1, "manufactured_rhs": 2}[which]
        h = lib().refs_pde_vector(self.h_, mesh.h_, w)
        if not h:
            raise RuntimeError(lib().refs_last_error().decode())
        return Vec(mesh, h)

    def lhs(self, mesh):
        h = lib().refs_pde_lhs(self.h_, mesh.h_)
        if not h:
            raise RuntimeError(lib().refs_last_error().decode())
        return Mat(mesh, h)

    def __del__(self):
        try:
            lib().refs_pde_destroy(self.h_)
        except Exception:
            pass


def apply(A, x, y):
    _chk(lib().refs_apply(A.h_, x.h_, y.h_))


def apply_res(A, b, x, y):
    _chk(lib().refs_apply_res(A.h_, b.h_, x.h_, y.h_))


def vecaxpy(a, x, y):
    _chk(lib().refs_vecaxpy(C.c_double(a), x.h_, y.h_))


def vecset(a, x):
    _chk(lib().refs_vecset(C.c_double(a), x.h_))


def vecassign(x, y):
    _chk(lib().refs_vecassign(x.h_, y.h_))


def vec_multi_axpy(a, xs, y):
    n = len(xs)
    _chk(lib().refs_vec_multi_axpy(n, _a(a), (_vp * n)(*[x.h_ for x in xs]), y.h_))


def norm_L2(x):
    return lib().refs_norm_L2(x.h_)


def norm_vector_l2(x):
    return lib().refs_norm_vector_l2(x.h_)


def inner_vector_l2(x, y):
    return lib().refs_inner_vector_l2(x.h_, y.h_)


def compute_error_L2(x, y):
    return lib().refs_compute_error_L2(x.h_, y.h_)


class Prec:
    def __init__(self, A, pc, apply_sweeps=1, build_sweeps=1, threaded_apply=False, threaded_build=False, chunk=1024):
        self.A = A
        self.h_ = _vp(lib().refs_prec_create(A.h_, pc.encode(), apply_sweeps, build_sweeps, int(threaded_apply), int(threaded_build), chunk))
        if not self.h_.value:
            raise RuntimeError(lib().refs_last_error().decode())

    def apply(self, r, z):
        _chk(lib().refs_prec_apply(self.h_, r.h_, z.h_))

    def __del__(self):
        try:
            lib().refs_prec_destroy(self.h_)
        except Exception:
            pass


def set_options(opts):
    """opts: dict of '-key' -> value, the same keys the reference's .perc files carry."""
    lib().refs_options_clear()
    for k, v in opts.items():
        lib().refs_options_set(k.encode(), str(v).encode())


def _capture_stdout(fn):
    """Runs fn() with fd 1 redirected to a temp file; returns (result, text)."""
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


def solve(A, b, x, opts):
    """The reference's own path: createSolver (options DB) -> updateOperator -> apply.
    Returns info dict incl. the printed residual history {step: rel_res}."""
    set_options(opts)
    conv, its = C.c_int(), C.c_int()
    rn, pct, bt, st = C.c_double(), C.c_double(), C.c_double(), C.c_double()

    def run():
        return lib().refs_solve(A.h_, b.h_, x.h_, C.byref(conv), C.byref(its), C.byref(rn), C.byref(pct), C.byref(bt), C.byref(st))

    rc, text = _capture_stdout(run)
    _chk(rc)
    printed = {int(m.group(1)): float(m.group(2)) for m in re.finditer(r"Step (\d+): Rel res = (\S+)", text)}
    return {"converged": bool(conv.value), "iters": its.value, "resnorm": rn.value, "precapplywtime": pct.value,
            "buildtime": bt.value, "solvetime": st.value, "printed": printed}


def harness(A, prec, b, x, ksp="cg", rtol=1e-6, maxiter=1000):
    conv, its, rn = C.c_int(), C.c_int(), C.c_double()
    hist = np.zeros(maxiter + 1)
    fn = lib().refs_harness_cg if ksp == "cg" else lib().refs_harness_bicgstab
    _chk(fn(A.h_, prec.h_, b.h_, x.h_, C.c_double(rtol), int(maxiter), C.byref(conv), C.byref(its), C.byref(rn), hist.ctypes.data_as(_dp)))
    return {"converged": bool(conv.value), "iters": its.value, "resnorm": rn.value}, hist[: its.value].copy()


def time_apply(A, x, y, reps=5):
    return lib().refs_time_apply(A.h_, x.h_, y.h_, reps)


def time_axpy(x, y, reps=5):
    return lib().refs_time_axpy(x.h_, y.h_, reps)


def time_dot(x, y, reps=5):
    s = C.c_double()
    return lib().refs_time_dot(x.h_, y.h_, reps, C.byref(s))


def num_threads():
    return lib().refs_num_threads()


def set_num_threads(n):
    lib().refs_set_num_threads(int(n))


def read_control(path):
    npd, rmin, rmax = (C.c_int * 3)(), (C.c_double * 3)(), (C.c_double * 3)()
    cheb, nruns, mu = C.c_int(), C.c_int(), C.c_double()
    pdet = C.create_string_buffer(32)
    vel, bt, bv = (C.c_double * 3)(), (C.c_int * 6)(), (C.c_double * 6)()
    _chk(lib().refs_read_control(path.encode(), npd, rmin, rmax, C.byref(cheb), pdet, vel, C.byref(mu), C.byref(nruns), bt, bv))
    return {"npdim": list(npd), "rmin": list(rmin), "rmax": list(rmax), "grid": "chebyshev" if cheb.value else "uniform",
            "pde": pdet.value.decode(), "vel": list(vel), "mu": mu.value, "nruns": nruns.value,
            "btypes": list(bt), "bvals": list(bv)}
