#!/usr/bin/env python
"""A short BiCGSTAB + SGS run (convection-diffusion n^3) for launch-list profiling."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from struct3d_b200 import hostapi as H
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
order = sys.argv[2] if len(sys.argv) > 2 else "lexicographic"
maxit = int(sys.argv[3]) if len(sys.argv) > 3 else 10
H.init(0)
m = H.Mesh((n + 2,) * 3)
P = H.Pde("convdiff", (0.577350269189626, 0.7, 0.3), 0.1)
A, b, u = P.lhs(m), P.vector(m, "manufactured_rhs"), H.Vec(m)
o = {"-s3d_ksp_type": "bcgs", "-s3d_pc_type": "sgs", "-ksp_rtol": 1e-6, "-ksp_max_it": maxit, "-s3d_pc_apply_sweeps": 1,
     "-s3d_thread_chunk_size": 1024, "-s3d_sgs_ordering": order}
H.set_options(o)
s = H.Solver(A); s.update(); H.sync()
for rep in range(2):
    H.vecset(0.0, u)
    t0 = time.perf_counter(); info = s.apply(b, u); H.sync(); dt = time.perf_counter() - t0
    print(f"{order} n={n} rep {rep}: {info['iters']} its in {dt*1e3:.2f} ms = {dt*1e3/max(1,info['iters']):.3f} ms/it, precapply {info['precapplywtime']*1e3:.2f} ms", flush=True)
