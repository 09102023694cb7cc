#!/usr/bin/env python
"""CG + Jacobi on Poisson n^3 (golden right-hand side): element-wise kernel order A/B in one process (ew_unroll -1 = grid-stride, 0 = segments)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from struct3d_b200 import capi, hostapi as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
H.init(0)
ctx = capi.Context.from_handle(H.cuda_ctx())
m = H.Mesh((n + 2,) * 3)
P = H.Pde("poisson")
A, b, u = P.lhs(m), H.Vec(m), H.Vec(m)
idx = np.arange(1, n ** 3 + 1, dtype=np.float64) * 0.6180339887
f = np.zeros((n + 2,) * 3); f[1:-1, 1:-1, 1:-1] = (idx - np.floor(idx) - 0.5).reshape((n, n, n)); b.set(f)
H.set_options({"-s3d_ksp_type": "cg", "-s3d_pc_type": "jacobi", "-ksp_rtol": 1e-6, "-ksp_max_it": 5000, "-s3d_pc_apply_sweeps": 1, "-s3d_thread_chunk_size": 1024})
s = H.Solver(A); s.update(); H.sync()
for rep in range(3):
    for U in (-1, 0):
        ctx.tune("ew_unroll", U)
        H.vecset(0.0, u)
        t0 = time.perf_counter(); info = s.apply(b, u); H.sync(); dt = time.perf_counter() - t0
        print(f"n={n} ew_unroll={U}: {info['iters']} iterations, {dt:.4f} s, {dt * 1e3 / info['iters']:.4f} ms/iteration", flush=True)
