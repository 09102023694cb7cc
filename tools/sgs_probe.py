#!/usr/bin/env python
"""Times the exact-lexicographic SGS apply: streaming warp-per-line kernel (variant 4, the default) against warp-per-tile (variant 0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from struct3d_b200 import capi
ctx = capi.Context(0)
sizes = [int(a) for a in sys.argv[1:]] or [64, 128, 256, 512]
for n in sizes:
    g = capi.make_grid(n, n, n)
    c = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1)] * 3
    A = ctx.mat_alloc(g); ctx.assemble_convdiff(g, c, (0.577, 0.7, 0.3), 0.1, A)
    r, z, y, zr = (ctx.vec_alloc(g) for _ in range(4))
    ctx.vecset(g, 0.3, r)
    dinv = ctx.scalars_alloc(int(g.cstride)); ctx.sgs_setup(g, A, dinv)
    ctx.tune("sgs_variant", 0)
    ctx.sgs_apply(g, A, dinv, r, zr, y, capi.SGS_LEXICOGRAPHIC, 1)
    ref = ctx.vec_download(g, zr)
    for variant, wps, depth, dbgv in ((0, 0, 0, 0), (5, 0, 0, 0), (5, 2, 0, 0), (5, 0, -1, 0)):
        ctx.tune("sgs_sleep_ns", 1 if depth == -1 else 0)   # -1: timing experiment without any hand-off between lines (wrong results)
        if depth == -1: depth = 0
        ctx.tune("sgs_variant", variant); ctx.tune("sgs_warps_per_sm", wps); ctx.tune("sgs_tile_i", depth)
        ctx.tune("sgs_pencils", dbgv)
        for _ in range(2): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
        same = np.array_equal(ctx.vec_download(g, z), ref)
        ctx.sync(); ctx.timer_start()
        for _ in range(5): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
        ms = ctx.timer_stop_ms() / 5
        print(f"{n:4d}^3 variant={variant} warps/SM={wps} tile_i={depth} pencils={dbgv}  {ms:9.4f} ms   {96*n**3/ms/1e6:8.1f} GB/s(96B/pt)  bitwise_same={same}", flush=True)
        if variant == 5 and wps in (0,) and os.environ.get("SGS_PROF"):
            import ctypes
            ctx.tune("sgs_profile", 1)
            ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
            ctx.tune("sgs_profile", 0)
    ctx.tune("sgs_sleep_ns", 0)
    ctx.tune("sgs_variant", capi.SGS_VARIANT_DEFAULT); ctx.tune("sgs_warps_per_sm", 0); ctx.tune("sgs_tile_i", 0); ctx.tune("sgs_pencils", 0)
    for dbg in ():
        ctx.tune("sgs_sleep_ns", dbg)
        ctx.tune("sgs_profile", 1)
        ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
        print("debug", dbg, file=sys.stderr, end=" ", flush=True); ctx.tune("sgs_profile", 0)
    ctx.tune("sgs_sleep_ns", 0)
    for p in (A, r, z, y, zr, dinv): ctx.free(p)
