#!/usr/bin/env python
"""Red-black SGS: plain kernels against the colour-packed coefficient layout (bit-identical; timing per apply)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from struct3d_b200 import capi
ctx = capi.Context(0)
for n in [int(a) for a in sys.argv[1:]] or [64, 128, 256, 512]:
    g = capi.make_grid(n, n, n)
    c = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1)] * 3
    A = ctx.mat_alloc(g); ctx.assemble_convdiff(g, c, (0.577, 0.7, 0.3), 0.1, A)
    r, z, y, z2 = (ctx.vec_alloc(g) for _ in range(4)); ctx.vecset(g, 0.3, r)
    dinv = ctx.scalars_alloc(int(g.cstride)); ctx.sgs_setup(g, A, dinv)
    packed = ctx.sgs_rb_pack(g, A, dinv)
    for _ in range(3): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_REDBLACK, 1)
    ctx.sync(); ctx.timer_start()
    for _ in range(10): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_REDBLACK, 1)
    ms0 = ctx.timer_stop_ms() / 10
    for _ in range(3): ctx.sgs_apply_rb_packed(g, packed, r, z2, y)
    ctx.sync(); ctx.timer_start()
    for _ in range(10): ctx.sgs_apply_rb_packed(g, packed, r, z2, y)
    ms1 = ctx.timer_stop_ms() / 10
    same = np.array_equal(ctx.vec_download(g, z), ctx.vec_download(g, z2))
    print(f"{n}^3 red-black SGS apply: plain {ms0:.4f} ms ({96*n**3/ms0/1e6:.0f} GB/s)  packed {ms1:.4f} ms ({96*n**3/ms1/1e6:.0f} GB/s)  bitwise_same={same}", flush=True)
    for p in (A, r, z, y, z2, dinv, packed): ctx.free(p)
