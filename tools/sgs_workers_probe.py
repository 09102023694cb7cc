import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from struct3d_b200 import capi
ctx = capi.Context(0)
for n in (256, 512):
    g = capi.make_grid(n, n, n)
    c = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1)] * 3
    A = ctx.mat_alloc(g); ctx.assemble_convdiff(g, c, (0.577, 0.7, 0.3), 0.1, A)
    r, z, y = (ctx.vec_alloc(g) for _ in range(3)); ctx.vecset(g, 0.3, r)
    dinv = ctx.scalars_alloc(int(g.cstride)); ctx.sgs_setup(g, A, dinv)
    ctx.tune("sgs_variant", 0)
    for p in (2, 1):
        for wps in (0, 8, 6, 4, 3, 2):
            ctx.tune("sgs_pencils", p); ctx.tune("sgs_warps_per_sm", wps)
            for _ in range(3): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
            ctx.sync(); ctx.timer_start()
            for _ in range(10): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
            print(f"{n}^3 P={p} warps/SM={wps or 'max'}: {ctx.timer_stop_ms()/10:.4f} ms", flush=True)
    for p_ in (A, r, z, y, dinv): ctx.free(p_)
