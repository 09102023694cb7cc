#!/usr/bin/env python
"""Debug aid: the streaming SGS kernel against the oracle on small grids, forward result (scratch y) and final z."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from struct3d_b200 import capi
from oracle import cport
sys.path.insert(0, os.path.join(ROOT, "tests"))
ctx = capi.Context(0)
V1 = (0.577350269189626, -0.7, 0.3)
for npdim in [(3, 3, 3), (4, 3, 3), (3, 4, 3), (3, 3, 4), (10, 3, 3), (11, 3, 3), (5, 6, 7)]:
    for pde in ("poisson", "convdiff"):
        m = cport.Mesh(npdim, grid="chebyshev")
        A = cport.lhs(m, pde, V1, 0.1)
        rng = np.random.default_rng(40)
        r = m.zeros(); r[1:-1, 1:-1, 1:-1] = rng.standard_normal(r[1:-1, 1:-1, 1:-1].shape)
        g = capi.make_grid(*m.n)
        dA, dr, dz, dy = ctx.mat_from(g, A), ctx.vec_from(g, r), ctx.vec_alloc(g), ctx.vec_alloc(g)
        dinv = ctx.scalars_alloc(int(g.cstride))
        ctx.sgs_setup(g, dA, dinv)
        di = cport.sgs_setup(m, A)
        ref = cport.sgs_like_apply(m, A, di, r, 1)
        for variant in (0, 5):
            ctx.tune("sgs_variant", variant)
            ctx.vecset(g, 7.0, dz)
            ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_LEXICOGRAPHIC, 1)
            z, y = ctx.vec_download(g, dz), ctx.vec_download(g, dy)
            print(npdim, pde, "variant", variant, "z equal", np.array_equal(z, ref), "max diff", np.abs(z - ref).max(), flush=True)
            if variant == 0:
                y0 = y
            elif not np.array_equal(y, y0):
                bad = np.argwhere(y != y0)
                print("   forward sweep differs at (k,j,i):", bad[:6].tolist(), "got", y[tuple(bad[0])], "want", y0[tuple(bad[0])])
            elif not np.array_equal(z, ref):
                bad = np.argwhere(z != ref)
                print("   backward sweep differs at (k,j,i):", bad[:6].tolist(), "got", z[tuple(bad[0])], "want", ref[tuple(bad[0])])
        ctx.tune("sgs_variant", capi.SGS_VARIANT_DEFAULT)
        for p in (dA, dr, dz, dy, dinv): ctx.free(p)
