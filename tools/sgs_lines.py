#!/usr/bin/env python
"""Timing experiments for the streaming SGS kernel on grids of a few lines: python tools/sgs_lines.py nx,ny,nz [nx,ny,nz ...]
One line (ny <= 4, nz <= 8) shows the cost of a step without any hand-off; two lines in j or k show the lag of a hand-off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from struct3d_b200 import capi
ctx = capi.Context(0)
ctx.tune("sgs_variant", 5)
dbg = int(os.environ.get("SGS_DEBUG", "0"))
ctx.tune("sgs_sleep_ns", dbg)
for spec in sys.argv[1:]:
    nx, ny, nz = (int(v) for v in spec.split(","))
    g = capi.make_grid(nx, ny, nz)
    c = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1) for n in (nx, ny, nz)]
    A = ctx.mat_alloc(g); ctx.assemble_convdiff(g, c, (0.577, 0.7, 0.3), 0.1, A)
    r, z, y = (ctx.vec_alloc(g) for _ in range(3))
    ctx.vecset(g, 0.3, r)
    dinv = ctx.scalars_alloc(int(g.cstride)); ctx.sgs_setup(g, A, dinv)
    for _ in range(3): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
    ctx.sync(); ctx.timer_start()
    for _ in range(10): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
    ms = ctx.timer_stop_ms() / 10
    steps = nx + 2 * (min(ny, 4) - 1) + 2 * (min(nz, 8) - 1)
    print(f"{spec:>12s}: {ms * 1e3:9.1f} us per apply = {ms * 1e6 / 2:9.0f} ns per sweep; {ms * 1e6 / 2 * 1.965 / steps:7.1f} cycles per step if it were one line", flush=True)
    ctx.tune("sgs_profile", 1)
    ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
    ctx.tune("sgs_profile", 0)
    for p in (A, r, z, y, dinv): ctx.free(p)
