import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from struct3d_b200 import capi
ctx = capi.Context(0)
for n in (14, 32, 64, 256):
    g = capi.make_grid(n, n, n)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(20):
        p = ctx.vec_alloc(g); ctx.free(p)
    ctx.sync()
    print(n, "vec_alloc+free us:", (time.perf_counter() - t0) / 20 * 1e6)
