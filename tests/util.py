"""Shared helpers for the parity tests: seeded inputs in the reference's host layouts."""
import numpy as np

from oracle import cport

V1 = (0.577350269189626, 0.7, 0.3)
V2 = (-0.577350269189626, 0.577350269189626, 0.577350269189626)


def rand_vec(mesh, seed, ghosts_zero=True):
    rng = np.random.default_rng(seed)
    v = rng.uniform(-1.0, 1.0, size=mesh.vshape)
    if ghosts_zero:
        v[0], v[-1] = 0, 0
        v[:, 0], v[:, -1] = 0, 0
        v[:, :, 0], v[:, :, -1] = 0, 0
    return np.ascontiguousarray(v)


def golden_vec(mesh):
    """x_idx = frac(0.6180339887*(idx+1)) - 0.5 over real points, ghosts 0."""
    nx, ny, nz = mesh.n
    idx = np.arange(1, nx * ny * nz + 1, dtype=np.float64) * 0.6180339887
    v = mesh.zeros()
    v[1:-1, 1:-1, 1:-1] = (idx - np.floor(idx) - 0.5).reshape(nz, ny, nx)
    return v


def problem(npdim, grid="uniform", pde="poisson", vel=(0, 0, 0), mu=0.0):
    m = cport.Mesh(npdim, grid=grid)
    A = cport.lhs(m, pde, vel, mu)
    return m, A


def real(v):
    return v[1:-1, 1:-1, 1:-1]


def relerr(a, b):
    d = np.linalg.norm((a - b).ravel())
    n = np.linalg.norm(b.ravel())
    return d / n if n > 0 else d


def ghosts_untouched(v):
    return (not v[0].any() and not v[-1].any() and not v[:, 0].any() and not v[:, -1].any()
            and not v[:, :, 0].any() and not v[:, :, -1].any())
