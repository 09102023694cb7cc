#!/usr/bin/env python
"""Register-streaming SGS kernel (variant 6): bitwise check against the warp-per-tile kernel and timing per block shape.
usage: sgs_pencil_probe.py [--nolink] [--shapes 42,22,...] n1 n2 ...   (n or nx,ny,nz)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from struct3d_b200 import capi
args = sys.argv[1:]
nolink = "--nolink" in args
if nolink: args.remove("--nolink")
shapes = [42, 22, 41, 21]
variant = 6
if "--variant" in args:
    i = args.index("--variant"); variant = int(args[i + 1]); del args[i:i + 2]
    if variant == 7: shapes = [408, 404, 204]
if "--shapes" in args:
    i = args.index("--shapes"); shapes = [int(v) for v in args[i + 1].split(",")]; del args[i:i + 2]
ctx = capi.Context(0)
for spec in args or ["64", "128", "256"]:
    dims = [int(v) for v in spec.split(",")]
    if len(dims) == 1: dims = dims * 3
    nx, ny, nz = dims
    g = capi.make_grid(nx, ny, nz)
    c = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1) for n in (nx, ny, nz)]
    A = ctx.mat_alloc(g); ctx.assemble_convdiff(g, c, (0.577, 0.7, 0.3), 0.1, A)
    r, z, y, zr = (ctx.vec_alloc(g) for _ in range(4))
    rng = np.random.default_rng(5)
    rh = np.zeros((nz + 2, ny + 2, nx + 2)); rh[1:-1, 1:-1, 1:-1] = rng.standard_normal((nz, ny, nx))
    ctx.vec_upload(g, r, rh)
    dinv = ctx.scalars_alloc(int(g.cstride)); ctx.sgs_setup(g, A, dinv)
    ctx.tune("sgs_variant", 0)
    for _ in range(2): ctx.sgs_apply(g, A, dinv, r, zr, y, capi.SGS_LEXICOGRAPHIC, 1)
    ctx.sync(); ctx.timer_start()
    for _ in range(5): ctx.sgs_apply(g, A, dinv, r, zr, y, capi.SGS_LEXICOGRAPHIC, 1)
    ms0 = ctx.timer_stop_ms() / 5
    ref = ctx.vec_download(g, zr)
    npts = nx * ny * nz
    print(f"{spec:>12s} tile kernel            {ms0:9.4f} ms  {96*npts/ms0/1e6:8.1f} GB/s", flush=True)
    ctx.tune("sgs_variant", variant)
    for shape in shapes:
        ctx.tune("sgs_pencil_shape", shape)
        ctx.tune("sgs_sleep_ns", 1 if nolink else 0)
        ctx.vecset(g, 7.0, z)
        for _ in range(2): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
        got = ctx.vec_download(g, z)
        same = np.array_equal(got, ref)
        rel = float(np.abs(got - ref).max() / np.abs(ref).max())
        ctx.sync(); ctx.timer_start()
        for _ in range(5): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
        ms = ctx.timer_stop_ms() / 5
        extra = ""
        if not same and not nolink:
            bad = np.argwhere(got != ref)
            extra = f" first diff at (k,j,i)={tuple(bad[0])} of {len(bad)}: got {got[tuple(bad[0])]} want {ref[tuple(bad[0])]}"
        print(f"{spec:>12s} pencil shape {shape:4d}{' nolink' if nolink else ''}  {ms:9.4f} ms  {96*npts/ms/1e6:8.1f} GB/s  bitwise_same={same} max_rel_diff={rel:.2e}{extra if variant == 6 else ''}", flush=True)
    ctx.tune("sgs_sleep_ns", 0); ctx.tune("sgs_pencil_shape", 0); ctx.tune("sgs_variant", capi.SGS_VARIANT_DEFAULT)
    for p in (A, r, z, y, zr, dinv): ctx.free(p)
