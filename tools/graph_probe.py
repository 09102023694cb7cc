#!/usr/bin/env python
"""Solve time with iteration batches enqueued eagerly vs replayed from a CUDA graph (-s3d_cuda_graph off|on), small to medium grids."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from struct3d_b200 import hostapi as H
H.init(0)
sizes = [int(a) for a in sys.argv[1:]] or [32, 64, 128, 192, 256]
print("| grid | solver | iterations | eager ms/iteration | graph ms/iteration | gain |")
print("|---|---|---|---|---|---|")
for n in sizes:
    m = H.Mesh((n + 2,) * 3)
    for name, pde, ksp, pc, extra in (("CG + Jacobi, Poisson", "poisson", "cg", "jacobi", {}),
                                      ("BiCGSTAB + SGS (lexicographic), convection-diffusion", "convdiff", "bcgs", "sgs", {}),
                                      ("BiCGSTAB + SGS (red-black), convection-diffusion", "convdiff", "bcgs", "sgs", {"-s3d_sgs_ordering": "redblack"}),
                                      ("Richardson + SGS, convection-diffusion", "convdiff", "richardson", "sgs", {})):
        P = H.Pde(pde, (0.577350269189626, 0.7, 0.3), 0.1)
        A, b, u = P.lhs(m), P.vector(m, "manufactured_rhs"), H.Vec(m)
        if pde == "poisson":      # the manufactured right-hand side is an eigenvector of the discrete Laplacian (CG converges in one step)
            idx = np.arange(1, n ** 3 + 1, dtype=np.float64) * 0.6180339887
            f = np.zeros((n + 2,) * 3); f[1:-1, 1:-1, 1:-1] = (idx - np.floor(idx) - 0.5).reshape((n, n, n))
            b.set(f)
        res = {}
        for mode in ("off", "on"):
            o = {"-s3d_ksp_type": ksp, "-s3d_pc_type": pc, "-ksp_rtol": 1e-8, "-ksp_max_it": 400, "-s3d_pc_apply_sweeps": 1,
                 "-s3d_pc_build_sweeps": 1, "-s3d_thread_chunk_size": 1024, "-s3d_cuda_graph": mode}
            o.update(extra)
            H.set_options(o)
            s = H.Solver(A); s.update(); H.sync()
            best = 1e9
            for rep in range(3):
                H.vecset(0.0, u)
                t0 = time.perf_counter(); info = s.apply(b, u); H.sync(); best = min(best, time.perf_counter() - t0)
            res[mode] = (info["iters"], best)
            del s
        it = res["off"][0]
        assert res["on"][0] == it
        e, gph = res["off"][1] * 1e3 / it, res["on"][1] * 1e3 / it
        print(f"| {n}^3 | {name} | {it} | {e:.4f} | {gph:.4f} | {e / gph:.2f}x |", flush=True)
