#!/usr/bin/env python
"""exact sweep and blocks ordering (block counts by size), ms per application: python tools/sgs_quick_probe.py [n ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from struct3d_b200 import capi
ctx = capi.Context(0)
for n in [int(v) for v in sys.argv[1:]] or [64, 128, 256, 512]:
    g = capi.make_grid(n, n, n)
    c = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1)] * 3
    A = ctx.mat_alloc(g); ctx.assemble_convdiff(g, c, (0.577, 0.7, 0.3), 0.1, A)
    r, z, y = (ctx.vec_alloc(g) for _ in range(3))
    ctx.vecset(g, 0.3, r)
    dinv = ctx.scalars_alloc(int(g.cstride)); ctx.sgs_setup(g, A, dinv)
    out = []
    for name, order in (("exact", capi.SGS_LEXICOGRAPHIC), ("blocks", capi.SGS_BLOCKS)):
        for _ in range(3): ctx.sgs_apply(g, A, dinv, r, z, y, order, 1)
        ctx.sync(); ctx.timer_start()
        for _ in range(20): ctx.sgs_apply(g, A, dinv, r, z, y, order, 1)
        out.append(f"{name} {ctx.timer_stop_ms() / 20:.4f} ms")
    print(f"{n}^3: " + ", ".join(out), flush=True)
    for p in (A, r, z, y, dinv): ctx.free(p)
