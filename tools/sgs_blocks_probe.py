#!/usr/bin/env python
"""z-blocks ordering of the warp-per-tile SGS kernel (S3D_SGS_BLOCKS) against the exact sweep: ms per application by block count and
pencils per lane.  Usage: python tools/sgs_blocks_probe.py [n ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from struct3d_b200 import capi
ctx = capi.Context(0)
for spec in sys.argv[1:] or ["64", "128", "256", "384", "512"]:
    dims = [int(v) for v in spec.split(",")]
    if len(dims) == 1: dims = dims * 3
    nx, ny, nz = dims
    g = capi.make_grid(nx, ny, nz)
    c = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1) for n in (nx, ny, nz)]
    A = ctx.mat_alloc(g); ctx.assemble_convdiff(g, c, (0.577, 0.7, 0.3), 0.1, A)
    r, z, y = (ctx.vec_alloc(g) for _ in range(3))
    ctx.vecset(g, 0.3, r)
    dinv = ctx.scalars_alloc(int(g.cstride)); ctx.sgs_setup(g, A, dinv)
    ctx.tune("sgs_variant", 0)
    for pencils in (2, 1):
        ctx.tune("sgs_pencils", pencils)
        for nby in (1, 2, 4, 8):
            if nby > 1 and ny // 8 < nby: continue
            ctx.tune("sgs_blocks_y", nby)
            row = []
            for nb in (1, 2, 4, 8, 16):
                if nb > 1 and nz // 8 < nb: continue
                ctx.tune("sgs_blocks", nb)
                order = capi.SGS_LEXICOGRAPHIC if nb == 1 and nby == 1 else capi.SGS_BLOCKS
                for _ in range(3): ctx.sgs_apply(g, A, dinv, r, z, y, order, 1)
                ctx.sync(); ctx.timer_start()
                for _ in range(10): ctx.sgs_apply(g, A, dinv, r, z, y, order, 1)
                ms = ctx.timer_stop_ms() / 10
                row.append(f"{nb}: {ms:.4f}")
            print(f"{spec:>8s} pencils/lane={pencils}  {nby} blocks of rows  ms per apply by blocks of planes  " + "  ".join(row), flush=True)
    # two staging areas (sgs_dbuf): exact sweep and a few block shapes
    for pencils in (2, 1):
        ctx.tune("sgs_pencils", pencils)
        row = []
        for db in (0, 1):
            ctx.tune("sgs_dbuf", db)
            for nb, nby in ((1, 1), (8, 1), (8, 4)):
                if nz // 8 < nb or ny // 8 < nby: continue
                ctx.tune("sgs_blocks", nb); ctx.tune("sgs_blocks_y", nby)
                order = capi.SGS_LEXICOGRAPHIC if nb == 1 and nby == 1 else capi.SGS_BLOCKS
                for _ in range(3): ctx.sgs_apply(g, A, dinv, r, z, y, order, 1)
                ctx.sync(); ctx.timer_start()
                for _ in range(10): ctx.sgs_apply(g, A, dinv, r, z, y, order, 1)
                ms = ctx.timer_stop_ms() / 10
                row.append(f"dbuf={db} {nb}x{nby}: {ms:.4f}")
        print(f"{spec:>8s} pencils/lane={pencils}  staging areas  " + "  ".join(row), flush=True)
    ctx.tune("sgs_dbuf", 0)
    ctx.tune("sgs_pencils", 0); ctx.tune("sgs_blocks", 0); ctx.tune("sgs_blocks_y", 0); ctx.tune("sgs_variant", capi.SGS_VARIANT_DEFAULT)
    for p in (A, r, z, y, dinv): ctx.free(p)
