#!/usr/bin/env python
"""Element-wise kernels at n^3: grid-stride order (ew_unroll 0) vs contiguous segments with U loads in flight (4, 8)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from struct3d_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ctx = capi.Context(0)
g = capi.make_grid(n, n, n)
x, y, w = (ctx.vec_alloc(g) for _ in range(3))
ctx.vecset(g, 0.5, x); ctx.vecset(g, 0.25, y)
A = ctx.mat_alloc(g); ctx.assemble_poisson(g, [-1.0 + 2.0 * np.arange(n + 2) / (n + 1)] * 3, A)
def t(fn, reps=20):
    for _ in range(3): fn()
    ctx.sync(); ctx.timer_start()
    for _ in range(reps): fn()
    return ctx.timer_stop_ms() / reps
N = n ** 3
p, q, r, v, sv, tv, rh = (ctx.vec_alloc(g) for _ in range(7))
for vec in (p, q, r, v, sv, tv, rh): ctx.vecset(g, 0.125, vec)
red = ctx.scalars_alloc(4)
one = capi.scalar(1e-9)
ops = {"set": (8, lambda: ctx.vecset(g, 0.5, w)), "copy": (16, lambda: ctx.vecassign(g, x, y)), "axpy": (24, lambda: ctx.vecaxpy(g, 1e-9, x, y)),
       "waxpby": (24, lambda: ctx.vecwaxpby(g, 1.0, x, 1e-9, y, w)), "dot": (16, lambda: ctx.dot_dev(g, x, y, red)),
       "nrm2": (8, lambda: ctx.nrm2sq_dev(g, x, red)), "jacobi": (24, lambda: ctx.jacobi_apply(g, A, x, w)),
       "cg_update": (56, lambda: ctx.cg_update(g, one, p, q, x, r, A, red)),
       "cg_dir_jacobi": (32, lambda: ctx.cg_direction_jacobi(g, one, A, r, p)),
       "bicg_dir": (32, lambda: ctx.bicg_direction(g, one, one, r, v, p)),
       "bicg_update": (72, lambda: ctx.bicg_update(g, one, p, one, q, sv, tv, rh, x, r, red))}
print("| ew_unroll | " + " | ".join(ops) + " |"); print("|---|" + "---|" * len(ops))
for U in (-1, 0, 2, 4, 8):
    ctx.tune("ew_unroll", U)
    print(f"| {U} | " + " | ".join(f"{b * N / t(fn) / 1e6:.0f}" for b, fn in ops.values()) + " |", flush=True)
