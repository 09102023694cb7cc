#!/usr/bin/env python
"""Fixed short kernel sequence for the round-2 ncu captures: the default SpMV (kchunk 2), assembly, the reduction tail,
Jacobi, and one SGS apply per dataflow kernel.  python tools/profile_r2.py [n]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from struct3d_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ctx = capi.Context(0)
g = capi.make_grid(n, n, n)
c = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1)] * 3
A = ctx.mat_alloc(g)
for _ in range(2):
    ctx.assemble_convdiff(g, c, (0.577, 0.7, 0.3), 0.1, A)        # axis_table_kernel + fill_stencil_kernel
x, y, z, s = (ctx.vec_alloc(g) for _ in range(4))
ctx.vecset(g, 0.5, x)
red = ctx.scalars_alloc(4)
for _ in range(3):
    ctx.spmv(g, A, x, y)                                            # spmv_march_kernel, default instantiation (kchunk 2)
for _ in range(2):
    ctx.spmv(g, A, x, y, w=x, want_yy=True, red=red)                # fused reductions + final_reduce_kernel
    ctx.dot_dev(g, x, y, red)
    ctx.jacobi_apply(g, A, x, z)
dinv = ctx.scalars_alloc(int(g.cstride)); ctx.sgs_setup(g, A, dinv)
# every build of the exact sweep: tile kernel with mailboxes (the default) and with flags + fences (r1), streaming, register-streaming,
# row-scan; then the asynchronous sweeps
for variant, mail in ((0, 1), (0, 0), (5, 1), (6, 1), (7, 1)):
    ctx.tune("sgs_variant", variant); ctx.tune("sgs_tile_mail", mail)
    for _ in range(2):
        ctx.sgs_apply(g, A, dinv, x, z, s, capi.SGS_LEXICOGRAPHIC, 1)
ctx.tune("sgs_variant", capi.SGS_VARIANT_DEFAULT); ctx.tune("sgs_tile_mail", 1)
for _ in range(2):
    ctx.sgs_apply(g, A, dinv, x, z, s, capi.SGS_ASYNC, 2)
ctx.sync()
print("done", n)
