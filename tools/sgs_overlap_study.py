"""CPU study (oracle only, no GPU): BiCGSTAB iteration counts of the z-slab SGS orderings on convection-diffusion.
  exact      the lexicographic sweep over the whole grid (what one GPU runs; the reference's sequential SGS)
  slablocal  the exact sweep inside each slab, the slab faces decoupled (zero inflow): the r2 default on z-slabs
  overlap v  every slab sweeps `v` more planes on either side (its neighbours' planes, recomputed from the neighbours' r with
             zero inflow further out) and keeps its own planes: restricted additive Schwarz with SGS as the subdomain solver
Usage: python tools/sgs_overlap_study.py nx ny nz_per_slab slabs [overlaps...]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import cport


def make_pc(m, A, dinv, slabs, ov):
    if slabs == 1:
        return lambda r: cport.sgs_like_apply(m, A, dinv, r, 1)
    return lambda r: cport.sgs_overlap_apply(m, A, dinv, r, slabs, ov)


def bicgstab(m, A, b, pc, rtol=1e-6, maxit=5000):
    dot = lambda u, v: float(np.vdot(u, v))
    x = m.zeros()
    r = b.copy()
    rhat = r.copy()
    bb = dot(b, b)
    rho = dot(rhat, r)
    p = r.copy()
    v = None
    alpha = omega = 1.0
    for it in range(maxit):
        if it > 0:
            rho_new = dot(rhat, r)
            beta = (rho_new / rho) * (alpha / omega)
            rho = rho_new
            p = r + beta * (p - omega * v)
        ph = pc(p)
        v = cport.apply(m, A, ph)
        alpha = rho / dot(rhat, v)
        s = r - alpha * v
        sh = pc(s)
        t = cport.apply(m, A, sh)
        omega = dot(t, s) / dot(t, t)
        x += alpha * ph + omega * sh
        r = s - omega * t
        if np.sqrt(dot(r, r) / bb) <= rtol:
            return it + 1
    return maxit


def main():
    nx, ny, nzs, slabs = (int(v) for v in sys.argv[1:5])
    ovs = [int(v) for v in sys.argv[5:]] or [0, 1, 2, 4]
    vel, mu = (0.577350269189626, 0.7, 0.3), 0.1
    m = cport.Mesh((nx + 2, ny + 2, nzs * slabs + 2))
    A = cport.lhs(m, "convdiff", vel, mu)
    b = cport.grid_function(m, "convdiff", "manufactured_rhs", vel, mu)
    dinv = cport.sgs_setup(m, A)
    exact = bicgstab(m, A, b, make_pc(m, A, dinv, 1, 0))
    print(f"grid {nx} x {ny} x {nzs * slabs}, {slabs} slabs: exact {exact}", flush=True)
    for ov in ovs:
        it = bicgstab(m, A, b, make_pc(m, A, dinv, slabs, ov))
        print(f"  overlap {ov}: {it} iterations ({100.0 * (it - exact) / exact:+.1f} %)", flush=True)


if __name__ == "__main__":
    main()
