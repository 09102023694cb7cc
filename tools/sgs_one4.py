#!/usr/bin/env python
"""A few exact-lexicographic SGS applies at n^3 with the default (streaming) kernel, for ncu: python tools/sgs_one4.py n"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from struct3d_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
ctx = capi.Context(0)
ctx.tune("sgs_variant", 5)
g = capi.make_grid(n, n, n)
c = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1)] * 3
A = ctx.mat_alloc(g); ctx.assemble_convdiff(g, c, (0.577, 0.7, 0.3), 0.1, A)
r, z, y = (ctx.vec_alloc(g) for _ in range(3))
ctx.vecset(g, 0.3, r)
dinv = ctx.scalars_alloc(int(g.cstride)); ctx.sgs_setup(g, A, dinv)
for _ in range(2): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
ctx.sync(); ctx.timer_start()
for _ in range(2): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
print(f"{n}^3: {ctx.timer_stop_ms() / 2:.4f} ms per apply")
