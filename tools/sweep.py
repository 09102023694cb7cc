#!/usr/bin/env python
"""Kernel bandwidth sweep on one B200: SpMV variants/tilings and the BLAS-1 family, 32^3..1024^3,
timed with CUDA events on the library's stream (warm-up 3, then `reps` back-to-back launches;
working sets >= 256^3 exceed the 126 MB L2, smaller ones are marked l2_resident).
Writes one JSON line per measurement to gpurun_out/sweep.jsonl and prints a table.
GB/s are ALGORITHMIC bytes (SpMV 72 B/pt, residual 80, axpy 24, dot 16, ...)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from struct3d_b200 import capi  # noqa: E402


def uniform_coords(n):
    return [-1.0 + 2.0 * np.arange(n + 2) / (n + 1) for _ in range(3)]


def timeit(ctx, fn, reps, warm=3):
    for _ in range(warm):
        fn()
    ctx.sync()
    ctx.timer_start()
    for _ in range(reps):
        fn()
    return ctx.timer_stop_ms() / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="64,128,256,512")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--full", action="store_true", help="also sweep tilings")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweep.jsonl"))
    a = ap.parse_args()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    ctx = capi.Context(0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    out = open(a.out, "a")
    rows = []

    def rec(n, name, bytes_per_pt, ms, **kw):
        N = n ** 3
        gbs = bytes_per_pt * N / (ms * 1e-3) / 1e9
        r = dict(n=n, kernel=name, ms=ms, bytes_per_pt=bytes_per_pt, gbs=gbs, frac_measured_peak=gbs / peak,
                 frac_8tbs=gbs / 8000.0, l2_resident=bytes_per_pt * N < 126e6, **kw)
        rows.append(r)
        out.write(json.dumps(r) + "\n")
        out.flush()
        print(f"{n:5d}^3 {name:34s} {ms:9.4f} ms {gbs:9.1f} GB/s  {100 * gbs / peak:5.1f}% of measured {peak:.0f}", flush=True)

    for n in [int(s) for s in a.sizes.split(",")]:
        g = capi.make_grid(n, n, n)
        reps = a.reps if n <= 512 else max(3, a.reps // 4)
        dA = ctx.mat_alloc(g)
        ctx.assemble_poisson(g, uniform_coords(n), dA)
        vs = [ctx.vec_alloc(g) for _ in range(6)]
        x, y, b, p, q, r = vs
        ctx.vecset(g, 0.5, x); ctx.vecset(g, 0.25, b); ctx.vecset(g, 0.1, p); ctx.vecset(g, 0.2, q); ctx.vecset(g, 0.3, r)
        red = ctx.scalars_alloc(8)
        ctx.sync()
        ctx.tune("spmv_kchunk", 0); ctx.tune("spmv_minb", 0)
        for rnd in range(2 if a.full else 1):
            rec(n, "spmv auto (march k8)", 72, timeit(ctx, lambda: ctx.spmv(g, dA, x, y), reps), round=rnd)
            rec(n, "spmv naive", 72, timeit(ctx, lambda: ctx.spmv(g, dA, x, y, variant=capi.SPMV_NAIVE), reps), round=rnd)
            rec(n, "spmv pair", 72, timeit(ctx, lambda: ctx.spmv(g, dA, x, y, variant=capi.SPMV_PAIR), reps), round=rnd)
            rec(n, "spmv tma (x planes staged in shared memory by cp.async.bulk)", 72, timeit(ctx, lambda: ctx.spmv(g, dA, x, y, variant=capi.SPMV_TMA), reps), round=rnd)
            rec(n, "spmv auto res (b-Ax)+yy", 80, timeit(ctx, lambda: ctx.spmv(g, dA, x, y, b=b, want_yy=True, red=red), reps), round=rnd)
            rec(n, "spmv auto +(x,Ax)", 72, timeit(ctx, lambda: ctx.spmv(g, dA, x, y, w=x, red=red), reps), round=rnd)
            if a.full:
                for minb in (3, 4, 6):
                    for kc in (2, 4, 8, 16, 32):
                        ctx.tune("spmv_minb", minb); ctx.tune("spmv_kchunk", kc)
                        rec(n, f"spmv march minb={minb} kchunk={kc}", 72, timeit(ctx, lambda: ctx.spmv(g, dA, x, y, variant=capi.SPMV_MARCH), reps), minb=minb, kchunk=kc, round=rnd)
                for minb in (3, 4):
                    for kc in (4, 8, 16):
                        ctx.tune("spmv_minb", minb); ctx.tune("spmv_kchunk", kc)
                        rec(n, f"spmv march +(x,Ax) minb={minb} kchunk={kc}", 72, timeit(ctx, lambda: ctx.spmv(g, dA, x, y, w=x, red=red, variant=capi.SPMV_MARCH), reps), minb=minb, kchunk=kc, round=rnd)
                ctx.tune("spmv_kchunk", 0); ctx.tune("spmv_minb", 0)
        one = capi.scalar(1e-3)
        rec(n, "vecaxpy", 24, timeit(ctx, lambda: ctx.vecaxpy(g, one, x, y), reps))
        rec(n, "vecassign", 16, timeit(ctx, lambda: ctx.vecassign(g, x, y), reps))
        rec(n, "vecset", 8, timeit(ctx, lambda: ctx.vecset(g, 0.5, y), reps))
        rec(n, "dot", 16, timeit(ctx, lambda: ctx.dot_dev(g, x, y, red), reps))
        rec(n, "nrm2sq", 8, timeit(ctx, lambda: ctx.nrm2sq_dev(g, x, red), reps))
        rec(n, "jacobi_apply", 24, timeit(ctx, lambda: ctx.jacobi_apply(g, dA, r, y), reps))
        rec(n, "cg_update (x,r,+2 dots, jacobi)", 56, timeit(ctx, lambda: ctx.cg_update(g, one, p, q, x, r, dA, red), reps))
        rec(n, "cg_direction_jacobi", 32, timeit(ctx, lambda: ctx.cg_direction_jacobi(g, one, dA, r, p), reps))
        rec(n, "assemble_poisson", 56, timeit(ctx, lambda: ctx.assemble_poisson(g, uniform_coords(n), dA), max(3, reps // 4)))
        if n <= 512:
            dinv = ctx.scalars_alloc(int(g.cstride))
            ctx.sgs_setup(g, dA, dinv)
            rec(n, "sgs_apply red-black", 96, timeit(ctx, lambda: ctx.sgs_apply(g, dA, dinv, r, y, q, capi.SGS_REDBLACK, 1), max(3, reps // 4)))
            rec(n, "sgs_apply lexicographic (dataflow)", 96, timeit(ctx, lambda: ctx.sgs_apply(g, dA, dinv, r, y, q, capi.SGS_LEXICOGRAPHIC, 1), max(3, reps // 4)))
            if n <= 128:
                rec(n, "sgs_apply lexicographic (per-plane launches)", 96, timeit(ctx, lambda: ctx.sgs_apply(g, dA, dinv, r, y, q, capi.SGS_LEXICOGRAPHIC_PLANES, 1), 3, warm=1))
            ctx.free(dinv)
        for v in vs:
            ctx.free(v)
        ctx.free(dA); ctx.free(red)
    out.close()


if __name__ == "__main__":
    main()
