#!/usr/bin/env python
"""Times the SGS dataflow kernel under different spin strategies / sizes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from struct3d_b200 import capi
ctx = capi.Context(0)
for n in (64, 256, 512):
    g = capi.make_grid(n, n, n)
    c = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1)] * 3
    A = ctx.mat_alloc(g); ctx.assemble_convdiff(g, c, (0.577, 0.7, 0.3), 0.1, A)
    r, z, y = (ctx.vec_alloc(g) for _ in range(3))
    ctx.vecset(g, 0.3, r)
    dinv = ctx.scalars_alloc(int(g.cstride)); ctx.sgs_setup(g, A, dinv)
    for sl in (0, 1, 2, 4, 8, 1|8, 1|2|8, 1|2|4|8):
        ctx.tune("sgs_sleep_ns", sl)
        for _ in range(2): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
        ctx.sync(); ctx.timer_start()
        for _ in range(5): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
        ms = ctx.timer_stop_ms() / 5
        print(f"{n:4d}^3 debug={sl:3d}  {ms:9.4f} ms   {96*n**3/ms/1e6:8.1f} GB/s(96B/pt)", flush=True)
    for p in (A, r, z, y, dinv): ctx.free(p)
