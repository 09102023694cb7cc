#!/usr/bin/env python
"""Per-tile timeline of one forward sweep of the warp-per-tile kernel (mailbox hand-off) on a small grid: where a hop's time goes.
S3D_SGS_TILELOG=1 python tools/sgs_tile_log.py [n]   (n^3 grid, at most 800 tiles)"""
import os, sys, subprocess, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("S3D_TILELOG_CHILD"):
    import numpy as np
    from struct3d_b200 import capi
    n = int(sys.argv[1])
    ctx = capi.Context(0)
    g = capi.make_grid(n, n, n)
    c = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1)] * 3
    A = ctx.mat_alloc(g); ctx.assemble_convdiff(g, c, (0.577, 0.7, 0.3), 0.1, A)
    r, z, y = (ctx.vec_alloc(g) for _ in range(3)); ctx.vecset(g, 0.3, r)
    dinv = ctx.scalars_alloc(int(g.cstride)); ctx.sgs_setup(g, A, dinv)
    ctx.tune("sgs_variant", 0)
    for _ in range(3): ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
    ctx.tune("sgs_profile", 1)
    ctx.sgs_apply(g, A, dinv, r, z, y, capi.SGS_LEXICOGRAPHIC, 1)
    ctx.tune("sgs_profile", 0)
    sys.exit(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
env = dict(os.environ, S3D_TILELOG_CHILD="1", S3D_SGS_TILELOG="1")
out = subprocess.run([sys.executable, __file__, str(n)], env=env, capture_output=True, text=True).stderr
tiles = {}
for m in re.finditer(r"\[sgs_tile\] (\d+) (\d+) (\d+) (\d+) (\d+) (\d+)", out):
    t = int(m.group(1)); tiles[t] = [int(m.group(k)) for k in range(2, 7)]
ntx, nty, ntz = (n + 15) // 16, (n + 7) // 8, (n + 7) // 8
if ntx * nty * ntz > 800:
    sys.exit(f"{n}^3 has {ntx * nty * ntz} tiles: the log holds 800 (use n <= 80)")
t0 = min(v[0] for v in tiles.values())
def T(ci, cj, ck): return tiles.get((ck * nty + cj) * ntx + ci)
# walk the critical path backwards from the last tile: at each tile, the predecessor whose faces arrived last
path, cur = [], (ntx - 1, nty - 1, ntz - 1)
while cur is not None:
    path.append(cur)
    ci, cj, ck = cur
    preds = [p for p in ((ci - 1, cj, ck), (ci, cj - 1, ck), (ci, cj, ck - 1)) if min(p) >= 0]
    cur = max(preds, key=lambda p: T(*p)[4]) if preds else None
path.reverse()
print(f"{n}^3: {len(tiles)} tiles, critical path of {len(path)} tiles, sweep {max(v[4] for v in tiles.values()) - t0} ns")
gaps = {"visible+discovery": [], "set-up": [], "sweep": [], "publish": [], "waited in the poll before the producer published": []}
for a, b in zip(path, path[1:]):
    pa, pb = T(*a), T(*b)
    gaps["visible+discovery"].append(pb[1] - pa[4])
    gaps["set-up"].append(pb[2] - pb[1])
    gaps["sweep"].append(pb[3] - pb[2])
    gaps["publish"].append(pb[4] - pb[3])
    gaps["waited in the poll before the producer published"].append(pa[4] - pb[0])
for k, v in gaps.items():
    v = sorted(v)
    print(f"  {k:55s} median {v[len(v)//2]:6d} ns   mean {sum(v)/len(v):8.0f} ns   min {v[0]} max {v[-1]}")
hop = [(T(*b)[4] - T(*a)[4]) for a, b in zip(path, path[1:])]
print(f"  hop (published -> published)  mean {sum(hop)/len(hop):.0f} ns")
