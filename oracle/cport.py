"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/s3d_oracle.c (the plain-C restatement).

All arrays are numpy float64 in the REFERENCE's host layouts (see s3d_oracle.h):
vectors ghosted `(np2, np1, np0)` C-order, matrices `(7, nz, ny, nx)` C-order.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PC = {"none": 0, "jacobi": 1, "sgs": 2, "strilu": 3}
PDE = {"poisson": 0, "convdiff": 1, "convdiff_circular": 2}
FUNC = {"test_rhs": 0, "manufactured_solution": 1, "manufactured_rhs": 2}

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class PcParams(C.Structure):
    _fields_ = [("pc", C.c_int), ("napplysweeps", C.c_int), ("nbuildsweeps", C.c_int)]


class SolveInfo(C.Structure):
    _fields_ = [("converged", C.c_int), ("iters", C.c_int), ("resnorm", C.c_double)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libs3d_oracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_mesh_h.restype = C.c_double
        for f in ("orc_norm_vector_l2", "orc_inner_vector_l2", "orc_norm_L2", "orc_compute_error_L2"):
            getattr(L, f).restype = C.c_double
        for f in ("orc_richardson", "orc_gcr", "orc_cg", "orc_bicgstab"):
            getattr(L, f).restype = SolveInfo
        _LIB = L
    return _LIB


def _d(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"], "need contiguous float64"
    return a.ctypes.data_as(_dp)


def _np(npdim):
    return (C.c_int * 3)(*[int(v) for v in npdim])


def _a3(v):
    return (C.c_double * 3)(*[float(t) for t in v])


class Mesh:
    """npdim counts boundary points (reference convention); coords are the three 1-D arrays."""

    def __init__(self, npdim, rmin=(-1.0, -1.0, -1.0), rmax=(1.0, 1.0, 1.0), grid="uniform"):
        self.npdim = tuple(int(v) for v in npdim)
        self.grid = grid
        self.coords = []
        for d in range(3):
            c = np.zeros(self.npdim[d])
            if grid == "uniform":
                lib().orc_coords_uniform(self.npdim[d], C.c_double(rmin[d]), C.c_double(rmax[d]), _d(c))
            else:
                lib().orc_coords_chebyshev(self.npdim[d], C.c_double(rmin[d]), C.c_double(rmax[d]), _d(c))
            self.coords.append(c)

    @property
    def n(self):
        return tuple(v - 2 for v in self.npdim)  # (nx, ny, nz)

    @property
    def vshape(self):
        return (self.npdim[2], self.npdim[1], self.npdim[0])

    @property
    def mshape(self):
        nx, ny, nz = self.n
        return (7, nz, ny, nx)

    def zeros(self):
        return np.zeros(self.vshape)

    def h(self):
        return lib().orc_mesh_h(_np(self.npdim), *[_d(c) for c in self.coords])

    def _c(self):
        return [_d(c) for c in self.coords]


def lhs(mesh, pde="poisson", vel=(0, 0, 0), mu=0.0):
    A = np.zeros(mesh.mshape)
    if pde == "poisson":
        lib().orc_lhs_poisson(_np(mesh.npdim), *mesh._c(), _d(A))
    elif pde == "convdiff_circular":
        lib().orc_lhs_convdiff_circ(_np(mesh.npdim), *mesh._c(), C.c_double(mu), _d(A))
    else:
        lib().orc_lhs_convdiff(_np(mesh.npdim), *mesh._c(), _a3(vel), C.c_double(mu), _d(A))
    return A


def grid_function(mesh, pde="poisson", func="test_rhs", vel=(0, 0, 0), mu=0.0, bvals=None):
    f = mesh.zeros()
    bv = (C.c_double * 6)(*bvals) if bvals is not None else None
    lib().orc_grid_function(_np(mesh.npdim), *mesh._c(), PDE[pde], FUNC[func], _a3(vel), C.c_double(mu), bv, _d(f))
    return f


def apply(mesh, A, x):
    y = mesh.zeros()
    lib().orc_apply(_np(mesh.npdim), _d(A), _d(x), _d(y))
    return y


def apply_res(mesh, A, b, x):
    y = mesh.zeros()
    lib().orc_apply_res(_np(mesh.npdim), _d(A), _d(b), _d(x), _d(y))
    return y


def vecaxpy(mesh, a, x, y):
    lib().orc_vecaxpy(_np(mesh.npdim), C.c_double(a), _d(x), _d(y))


def vecset(mesh, a, x):
    lib().orc_vecset(_np(mesh.npdim), C.c_double(a), _d(x))


def vecassign(mesh, x, y):
    lib().orc_vecassign(_np(mesh.npdim), _d(x), _d(y))


def vec_multi_axpy(mesh, a, xs, y):
    n = len(xs)
    arr = (_dp * n)(*[_d(x) for x in xs])
    lib().orc_vec_multi_axpy(_np(mesh.npdim), n, (C.c_double * n)(*a), arr, _d(y))


def norm_vector_l2(mesh, x):
    return lib().orc_norm_vector_l2(_np(mesh.npdim), _d(x))


def inner_vector_l2(mesh, x, y):
    return lib().orc_inner_vector_l2(_np(mesh.npdim), _d(x), _d(y))


def norm_L2(mesh, x):
    return lib().orc_norm_L2(_np(mesh.npdim), *mesh._c(), _d(x))


def compute_error_L2(mesh, x, y):
    return lib().orc_compute_error_L2(_np(mesh.npdim), *mesh._c(), _d(x), _d(y))


def jacobi_apply(mesh, A, r):
    z = mesh.zeros()
    lib().orc_jacobi_apply(_np(mesh.npdim), _d(A), _d(r), _d(z))
    return z


def sgs_setup(mesh, A):
    d = np.zeros(mesh.mshape[1:])
    lib().orc_sgs_setup(_np(mesh.npdim), _d(A), _d(d))
    return d


def strilu_setup(mesh, A, nbuildsweeps=1):
    d = np.zeros(mesh.mshape[1:])
    lib().orc_strilu_setup(_np(mesh.npdim), _d(A), int(nbuildsweeps), _d(d))
    return d


def sgs_like_apply(mesh, A, diaginv, r, nsweeps=1):
    z = mesh.zeros()
    lib().orc_sgs_like_apply(_np(mesh.npdim), _d(A), _d(diaginv), int(nsweeps), _d(r), _d(z))
    return z


def sgs_forward_inplace(mesh, A, diaginv, r, y):
    """forward half-sweep into y, reading y's ghost entries as they are"""
    lib().orc_sgs_forward_inplace(_np(mesh.npdim), _d(A), _d(diaginv), _d(r), _d(y))
    return y


def sgs_backward_inplace(mesh, A, diaginv, y, z):
    """backward half-sweep into z, reading z's ghost entries as they are"""
    lib().orc_sgs_backward_inplace(_np(mesh.npdim), _d(A), _d(diaginv), _d(y), _d(z))
    return z


def sgs_jacobi_sweeps_apply(mesh, A, diaginv, r, nsweeps):
    z = mesh.zeros()
    lib().orc_sgs_jacobi_sweeps_apply(_np(mesh.npdim), _d(A), _d(diaginv), int(nsweeps), _d(r), _d(z))
    return z


def sgs_redblack_apply(mesh, A, diaginv, r):
    z = mesh.zeros()
    lib().orc_sgs_redblack_apply(_np(mesh.npdim), _d(A), _d(diaginv), _d(r), _d(z))
    return z


def slab_bounds(nz, nslabs):
    """z-slab decomposition of nz planes (the library's s3d_grid_init_slab): slab p owns planes [p*nz/P, (p+1)*nz/P)"""
    return [(p * nz // nslabs, (p + 1) * nz // nslabs) for p in range(nslabs)]


def blocks_by_size(nx, ny, nz):
    """the library's default block counts (sgs.cu, sgs_blocks_run): about 32 planes / rows per block, none above 96 Mi points"""
    if nx * ny * nz > (96 << 20):
        return 1, 1
    return max(1, min(nz // 32, 16)), max(1, min(ny // 32, 8))


def sgs_overlap_apply(mesh, A, diaginv, r, nslabs, overlap=1, blocks=(1, 1)):
    """The overlapping z-slab ordering (S3D_SGS_OVERLAP; no counterpart in the reference, stated ON TOP of its sequential sweep,
    s3d_sgspreconditioners.cpp:53-110): the sequential SGS-like application over every slab extended by `overlap` planes of either
    neighbour (zero inflow beyond them), own planes kept.  overlap = 0 is the slab-local ordering; one slab is the sequential sweep.
    blocks = (planes, rows): every extended slab is cut further into concurrent blocks (sgs_blocks_apply); None: as the library chooses by size."""
    nx, ny, nz = mesh.n
    z = mesh.zeros()
    for (k0, k1) in slab_bounds(nz, nslabs):
        e0, e1 = max(0, k0 - overlap), min(nz, k1 + overlap)
        sm = Mesh((nx + 2, ny + 2, e1 - e0 + 2))            # only its dimensions matter to the sweep
        rs = sm.zeros()
        rs[1:-1] = r[1 + e0:1 + e1]
        bz, by = blocks_by_size(nx, ny, e1 - e0) if blocks is None else blocks
        zs = sgs_blocks_apply(sm, np.ascontiguousarray(A[:, e0:e1]), np.ascontiguousarray(diaginv[e0:e1]), rs, bz, by)
        z[1 + k0:1 + k1] = zs[1 + (k0 - e0):1 + (k1 - e0)]
    return z


def block_bounds(n, nblocks):
    """blocks of planes (rows) of the library's S3D_SGS_BLOCKS ordering (sgs.cu, axis_blocks): at most n // 8 blocks, thickness 7 modulo 8 (a block
    plus its overlap plane fills whole tile layers), the first blocks 8 thicker where that evens out the last one, which takes what is left (>= 8)"""
    B = max(1, min(nblocks, n // 8))
    if B == 1:
        return [(0, n)]
    thick = 8 * max(1, (n + B) // (8 * B)) - 1
    while B > 1 and n - (B - 1) * thick < 8:
        B -= 1
    rem = n - (B - 1) * thick
    nthicker = max(0, min((2 * (rem - n // B) + 8) // 16, B - 1))
    while nthicker > 0 and rem - 8 * nthicker < 8:
        nthicker -= 1
    bounds, k = [], 0
    for p in range(B):
        k1 = n if p == B - 1 else k + thick + (8 if p < nthicker else 0)
        bounds.append((k, k1))
        k = k1
    return bounds


def sgs_blocks_apply(mesh, A, diaginv, r, nblocks, nblocks_y=1):
    """The blocks ordering on one device (S3D_SGS_BLOCKS; no counterpart in the reference, stated ON TOP of its sequential half-sweeps,
    s3d_sgspreconditioners.cpp:53-110): the grid is cut into blocks of planes (and of rows).  Forward, every block is swept preceded by
    one plane and one row of the blocks below it (zero inflow under them), own points kept; backward, followed by one plane and one row
    of the blocks above, whose forward values are the ones their owners have computed, zero inflow over them, own points kept.
    Blocks as the library cuts them (block_bounds); one block is the sequential sweep."""
    nx, ny, nz = mesh.n
    y, z = mesh.zeros(), mesh.zeros()
    kb, jb = block_bounds(nz, nblocks), block_bounds(ny, nblocks_y)
    for (k0, k1) in kb:
        for (j0, j1) in jb:
            ek, ej = (k0 - 1 if k0 > 0 else k0), (j0 - 1 if j0 > 0 else j0)
            sm = Mesh((nx + 2, j1 - ej + 2, k1 - ek + 2))
            rs, ys = sm.zeros(), sm.zeros()
            rs[1:-1, 1:-1] = r[1 + ek:1 + k1, 1 + ej:1 + j1]
            sgs_forward_inplace(sm, np.ascontiguousarray(A[:, ek:k1, ej:j1]), np.ascontiguousarray(diaginv[ek:k1, ej:j1]), rs, ys)
            y[1 + k0:1 + k1, 1 + j0:1 + j1] = ys[1 + (k0 - ek):1 + (k1 - ek), 1 + (j0 - ej):1 + (j1 - ej)]
    for (k0, k1) in kb:
        for (j0, j1) in jb:
            ek, ej = (k1 + 1 if k1 < nz else k1), (j1 + 1 if j1 < ny else j1)
            sm = Mesh((nx + 2, ej - j0 + 2, ek - k0 + 2))
            ys, zs = sm.zeros(), sm.zeros()
            ys[1:-1, 1:-1] = y[1 + k0:1 + ek, 1 + j0:1 + ej]
            sgs_backward_inplace(sm, np.ascontiguousarray(A[:, k0:ek, j0:ej]), np.ascontiguousarray(diaginv[k0:ek, j0:ej]), ys, zs)
            z[1 + k0:1 + k1, 1 + j0:1 + j1] = zs[1:1 + (k1 - k0), 1:1 + (j1 - j0)]
    return z


def solve(mesh, A, b, ksp="richardson", pc="jacobi", rtol=1e-6, maxiter=1000, restart=30,
          apply_sweeps=1, build_sweeps=1, x0=None):
    """Returns (x, info dict, hist ndarray of relative residuals per iteration)."""
    x = mesh.zeros() if x0 is None else x0.copy()
    hist = np.zeros(maxiter + restart + 1)
    pp = PcParams(PC[pc], int(apply_sweeps), int(build_sweeps))
    L = lib()
    n = _np(mesh.npdim)
    if ksp == "richardson":
        info = L.orc_richardson(n, _d(A), pp, _d(b), _d(x), C.c_double(rtol), int(maxiter), _d(hist))
    elif ksp == "gcr":
        info = L.orc_gcr(n, _d(A), pp, _d(b), _d(x), C.c_double(rtol), int(maxiter), int(restart), _d(hist))
    elif ksp == "cg":
        info = L.orc_cg(n, _d(A), pp, _d(b), _d(x), C.c_double(rtol), int(maxiter), _d(hist))
    elif ksp in ("bcgs", "bicgstab"):
        info = L.orc_bicgstab(n, _d(A), pp, _d(b), _d(x), C.c_double(rtol), int(maxiter), _d(hist))
    else:
        raise ValueError("Need a valid solver option!")
    return x, {"converged": bool(info.converged), "iters": info.iters, "resnorm": info.resnorm}, hist[: info.iters].copy()
