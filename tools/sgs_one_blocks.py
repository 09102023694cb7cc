#!/usr/bin/env python
"""A few applications of the blocks ordering at n^3 (for ncu): python tools/sgs_one_blocks.py n [blocks of planes] [blocks of rows]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from struct3d_b200 import capi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ctx = capi.Context(0)
if len(sys.argv) > 2: ctx.tune("sgs_blocks", int(sys.argv[2]))
if len(sys.argv) > 3: ctx.tune("sgs_blocks_y", int(sys.argv[3]))
g = capi.make_grid(n, n, n)
c = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1)] * 3
A = ctx.mat_alloc(g); ctx.assemble_convdiff(g, c, (0.577, 0.7, 0.3), 0.1, A)
r, z, y = (ctx.vec_alloc(g) for _ in range(3))
ctx.vecset(g, 0.3, r)
dinv = ctx.scalars_alloc(int(g.cstride)); ctx.sgs_setup(g, A, dinv)
for _ in range(3): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_BLOCKS, 1)
ctx.sync(); ctx.timer_start()
for _ in range(3): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_BLOCKS, 1)
print(f"{n}^3 blocks ordering: {ctx.timer_stop_ms() / 3:.4f} ms per apply")
