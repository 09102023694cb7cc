#!/usr/bin/env python
"""Warp-per-tile SGS kernel: hand-off through sentinel mailboxes (sgs_tile_mail 1) against `out` + fences + epoch flags (0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from struct3d_b200 import capi
ctx = capi.Context(0)
for spec in sys.argv[1:] or ["64", "128", "256", "512"]:
    dims = [int(v) for v in spec.split(",")]
    if len(dims) == 1: dims = dims * 3
    nx, ny, nz = dims
    g = capi.make_grid(nx, ny, nz)
    c = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1) for n in (nx, ny, nz)]
    A = ctx.mat_alloc(g); ctx.assemble_convdiff(g, c, (0.577, 0.7, 0.3), 0.1, A)
    r, z, y = (ctx.vec_alloc(g) for _ in range(3))
    rng = np.random.default_rng(5)
    rh = np.zeros((nz + 2, ny + 2, nx + 2)); rh[1:-1, 1:-1, 1:-1] = rng.standard_normal((nz, ny, nx))
    ctx.vec_upload(g, r, rh)
    dinv = ctx.scalars_alloc(int(g.cstride)); ctx.sgs_setup(g, A, dinv)
    ctx.tune("sgs_variant", 0)
    res = {}
    for pencils in (2, 1):
        for mail in (0, 1):
            ctx.tune("sgs_tile_mail", mail); ctx.tune("sgs_pencils", pencils)
            ctx.vecset(g, 7.0, z)
            for _ in range(3): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
            got = ctx.vec_download(g, z)
            ctx.sync(); ctx.timer_start()
            for _ in range(10): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
            ms = ctx.timer_stop_ms() / 10
            res[(pencils, mail)] = got
            same = np.array_equal(got, res[(2, 0)])
            print(f"{spec:>12s} pencils/lane={pencils} mailboxes={mail}  {ms:9.4f} ms  {96*nx*ny*nz/ms/1e6:8.1f} GB/s  same_as_flags={same}", flush=True)
    ctx.tune("sgs_tile_mail", 1); ctx.tune("sgs_pencils", 0)
    ctx.tune("sgs_variant", 5)                       # the streaming kernel, for the size rule of the default
    for _ in range(3): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
    same = np.array_equal(ctx.vec_download(g, z), res[(2, 0)])
    ctx.sync(); ctx.timer_start()
    for _ in range(10): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
    ms = ctx.timer_stop_ms() / 10
    print(f"{spec:>12s} streaming kernel             {ms:9.4f} ms  {96*nx*ny*nz/ms/1e6:8.1f} GB/s  same_as_flags={same}", flush=True)
    ctx.tune("sgs_variant", capi.SGS_VARIANT_DEFAULT)
    for p in (A, r, z, y, dinv): ctx.free(p)
if os.environ.get("SGS_PROF"):
    # per-tile phase cycles (stderr): wait = flag wait (0 with mailboxes), "load" = ticket + polling / halo fetch, "total" = sweep loop, "face" = publish
    for n in (256, 512):
        g = capi.make_grid(n, n, n)
        c = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1)] * 3
        A = ctx.mat_alloc(g); ctx.assemble_convdiff(g, c, (0.577, 0.7, 0.3), 0.1, A)
        r, z, y = (ctx.vec_alloc(g) for _ in range(3)); ctx.vecset(g, 0.3, r)
        dinv = ctx.scalars_alloc(int(g.cstride)); ctx.sgs_setup(g, A, dinv)
        ctx.tune("sgs_variant", 0)
        for mail in (0, 1):
            ctx.tune("sgs_tile_mail", mail)
            for _ in range(2): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
            print(f"profile {n}^3 mailboxes={mail}", file=sys.stderr, flush=True)
            ctx.tune("sgs_profile", 1)
            ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
            ctx.tune("sgs_profile", 0)
        ctx.tune("sgs_tile_mail", 1); ctx.tune("sgs_variant", capi.SGS_VARIANT_DEFAULT)
        for p in (A, r, z, y, dinv): ctx.free(p)
