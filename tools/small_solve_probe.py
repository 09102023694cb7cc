#!/usr/bin/env python
"""Single-launch solver (cluster / cooperative-grid team) against the host-driven loop (CUDA-graph replay): CG+Jacobi Poisson."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from struct3d_b200 import hostapi as H
H.init(0)
for n in [int(a) for a in sys.argv[1:]] or [32, 48, 64, 96, 128, 160]:
    m = H.Mesh((n + 2, n + 2, n + 2)); P = H.Pde("poisson"); A, b = P.lhs(m), H.Vec(m)
    idx = np.arange(1, n ** 3 + 1, dtype=np.float64) * 0.6180339887
    bh = np.zeros((n + 2,) * 3); bh[1:-1, 1:-1, 1:-1] = (idx - np.floor(idx) - 0.5).reshape(n, n, n)
    b.set(bh)
    for ksp in ("cg", "richardson", "gcr"):
        res = {}
        for mode in ("off", "on"):
            best = 1e9
            for rep in range(3):
                x = H.Vec(m)
                o = {"-s3d_ksp_type": ksp, "-s3d_pc_type": "jacobi", "-ksp_rtol": 1e-6, "-ksp_max_it": 400 if ksp != "cg" else 5000, "-s3d_small_solve": mode}
                H.sync(); t0 = time.perf_counter()
                info = H.solve(A, b, x, o, quiet=True) if "quiet" in H.solve.__code__.co_varnames else H.solve(A, b, x, o)
                H.sync(); best = min(best, time.perf_counter() - t0)
            res[mode] = (info["iters"], best)
        (i0, t0_), (i1, t1_) = res["off"], res["on"]
        print(f"{n}^3 {ksp}+jacobi: host-driven {i0} its {t0_*1e3:.3f} ms ({t0_/max(1,i0)*1e6:.1f} us/it)   single launch {i1} its {t1_*1e3:.3f} ms ({t1_/max(1,i1)*1e6:.1f} us/it)", file=sys.stderr, flush=True)
