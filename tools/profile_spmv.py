#!/usr/bin/env python
"""Short, fixed kernel sequence for ncu: every SpMV variant (with and without cache hints), the fused-dot
forms, and the BLAS-1 / CG kernels, on Poisson n^3.  Prints event timings when run without ncu."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from struct3d_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = capi.Context(0)
g = capi.make_grid(n, n, n)
c = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1)] * 3
A = ctx.mat_alloc(g)
ctx.assemble_poisson(g, c, A)
x, y, b, p, q, r = (ctx.vec_alloc(g) for _ in range(6))
for v, a in ((x, 0.5), (b, 0.25), (p, 0.1), (q, 0.2), (r, 0.3)):
    ctx.vecset(g, a, v)
red = ctx.scalars_alloc(8)
ctx.sync()


def t(name, fn):
    fn(); ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    ms = ctx.timer_stop_ms() / reps
    print(f"{name:40s} {ms:8.4f} ms  {72 * n**3 / ms / 1e6:8.1f} GB/s(72B/pt)", flush=True)


if len(sys.argv) > 3 and sys.argv[3] == "ncu":
    # one launch each of the kernels worth a full-section capture
    ctx.tune("spmv_hints", 1)
    ctx.spmv(g, A, x, y, variant=capi.SPMV_NAIVE)
    ctx.spmv(g, A, x, y, variant=capi.SPMV_PAIR)
    ctx.tune("spmv_kchunk", 8)
    ctx.spmv(g, A, x, y, variant=capi.SPMV_MARCH)
    ctx.spmv(g, A, x, y, w=x, red=red, variant=capi.SPMV_MARCH)
    ctx.spmv(g, A, x, y, w=x, red=red, variant=capi.SPMV_PAIR)
    one = capi.scalar(1e-3)
    ctx.vecassign(g, x, y)
    ctx.vecaxpy(g, one, x, y)
    ctx.dot_dev(g, x, y, red)
    ctx.cg_update(g, one, p, q, x, r, A, red)
    ctx.cg_direction_jacobi(g, one, A, r, p)
    ctx.sync()
    print("ncu sequence done")
    sys.exit(0)

for hints in (1, 0):
    ctx.tune("spmv_hints", hints)
    t(f"naive hints={hints}", lambda: ctx.spmv(g, A, x, y, variant=capi.SPMV_NAIVE))
    t(f"pair  hints={hints}", lambda: ctx.spmv(g, A, x, y, variant=capi.SPMV_PAIR))
ctx.tune("spmv_hints", 1)
ctx.tune("spmv_kchunk", 8)
t("march kchunk=8", lambda: ctx.spmv(g, A, x, y, variant=capi.SPMV_MARCH))
t("pair +(x,Ax)", lambda: ctx.spmv(g, A, x, y, w=x, red=red, variant=capi.SPMV_PAIR))
t("naive +(x,Ax)", lambda: ctx.spmv(g, A, x, y, w=x, red=red, variant=capi.SPMV_NAIVE))
t("naive res+yy", lambda: ctx.spmv(g, A, x, y, b=b, want_yy=True, red=red, variant=capi.SPMV_NAIVE))
one = capi.scalar(1e-3)
t("vecaxpy", lambda: ctx.vecaxpy(g, one, x, y))
t("vecassign", lambda: ctx.vecassign(g, x, y))
t("dot", lambda: ctx.dot_dev(g, x, y, red))
t("cg_update", lambda: ctx.cg_update(g, one, p, q, x, r, A, red))
t("cg_direction_jacobi", lambda: ctx.cg_direction_jacobi(g, one, A, r, p))
t("jacobi", lambda: ctx.jacobi_apply(g, A, r, y))
