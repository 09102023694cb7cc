#!/usr/bin/env python
"""PCIe copy rates on the box: contiguous vs pitched (SVec-layout) copies, one direction and both at once."""
import ctypes as C, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
n = 512
nbytes = (n + 2) ** 3 * 8
h1 = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
h2 = torch.empty(nbytes // 8, dtype=torch.float64).pin_memory()
d1 = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
d2 = torch.empty(nbytes // 8, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
dt = t(lambda: d1.copy_(h1, non_blocking=True)); print(f"H2D contiguous pinned  {nbytes/dt/1e9:6.1f} GB/s")
dt = t(lambda: h2.copy_(d2, non_blocking=True)); print(f"D2H contiguous pinned  {nbytes/dt/1e9:6.1f} GB/s")
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
dt = t(both); print(f"H2D + D2H concurrent   {2*nbytes/dt/1e9:6.1f} GB/s total")
from struct3d_b200 import capi
ctx = capi.Context(0)
g = capi.make_grid(n, n, n)
dx = ctx.vec_alloc(g)
hp = ctx.host_alloc(nbytes)
def up(): ctx.call("s3d_vec_upload", C.byref(g), capi._ptr(dx), capi._ptr(hp)); ctx.sync()
def down(): ctx.call("s3d_vec_download", C.byref(g), capi._ptr(dx), capi._ptr(hp))
dt = t(up); print(f"H2D pitched 2-D (s3d_vec_upload)   {nbytes/dt/1e9:6.1f} GB/s")
dt = t(down); print(f"D2H pitched 2-D (s3d_vec_download) {nbytes/dt/1e9:6.1f} GB/s")
