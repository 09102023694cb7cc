#!/usr/bin/env python
"""bench.py -- headline benchmark of the Struct3d native structured-grid path on B200.

Contract (driver): `python bench.py --gpus N --steps K --warmup W [--impl reference]`; for N > 1 it is
launched under torchrun, one rank per GPU.  Prints ONE JSON line (rank 0).

Workload (BASELINE.json metric "7-pt SpMV HBM GB/s ... at 512^3"): Poisson 3-D on [-1,1]^3, uniform
grid, fp64, 512^3 unknowns PER GPU (z-slab decomposition: N GPUs solve 512 x 512 x 512N; weak scaling).
One step = one pass of the hot path over that grid: the halo exchange the path needs (N > 1) and
y = A x through SMat::apply, the reference-facing call (SMat::apply, reference src/linalg/matvec.cpp:59).
  value      aggregate algorithmic GB/s (72 B/point), data resident in HBM, CUDA-event
             time over K steps, max over ranks;
  e2e        the same call with HOST vectors: x uploaded from pinned host memory and y downloaded
             inside the timed region, every step;
  roofline   the SpMV kernel against the measured HBM copy peak (MEASURED_PEAKS.json);
  cpu_baseline / --impl reference: the reference's own CPU SMat::apply (oracle/_ref, built from the
             unmodified reference sources; falls back to the C port) on the box's host cores;
  solve      extra, not part of the timed steps, through createSolver()/apply(), the reference's call sequence: CG+Jacobi on the
             same grid; CG+Jacobi on a FIXED 1024^3 grid (strong scaling); one GPU: BiCGSTAB with every SGS ordering on
             convection-diffusion 256^3 and the exact SGS at 512^3; several GPUs: BiCGSTAB + the exact SGS through the z-slabs.
             A failure in this block is reported as solve.error; the line is printed regardless.

Two more modes (single GPU, not used by the driver; both may time oracle/_ref as the CPU baseline):
  --sweep    SpMV / axpy / dot bandwidth 32^3 .. 1024^3 against the roofline, reference CPU kernels beside it (BASELINE configs[4]);
  --solves   whole solves of the reference's as-is cases and the 256^3 cases: the reference on the host cores beside this library.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BYTES_SPMV = 72          # 7 coefficient streams + x + y, 8 B each
BYTES_CG_ITER = 200      # credited unfused CG+Jacobi iteration
_emit = print


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, device copy burst)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def golden_field(n_real_total, shape):
    idx = np.arange(1, n_real_total + 1, dtype=np.float64) * 0.6180339887
    return (idx - np.floor(idx) - 0.5).reshape(shape)


# ---------------------------------------------------------------------------------------------- reference arm

def cpu_spmv(n, reps, threads=None):
    """Times the reference's own CPU SMat::apply (or the C port when oracle/_ref is absent) on Poisson n^3."""
    from oracle import cport, refshim
    npd = (n + 2,) * 3
    if refshim.available():
        if threads:
            refshim.set_num_threads(threads)
        cores = refshim.num_threads()
        _, _ = refshim._capture_stdout(lambda: None)
        m, _t = refshim._capture_stdout(lambda: refshim.Mesh(npd))
        A, x, y = refshim.Mat(m), refshim.Vec(m), refshim.Vec(m)
        h2 = (2.0 / (n + 1)) ** 2
        for s in range(7):                       # uniform Poisson coefficients; the serial reference assembly is not what is timed
            A.a[s][...] = (6.0 if s == 3 else -1.0) / h2
        x.a[1:-1, 1:-1, 1:-1] = golden_field(n ** 3, (n, n, n))
        best = refshim.time_apply(A, x, y, reps)
        kind = "reference"
        chk = float(np.abs(y.a).max())
    else:
        cores = os.cpu_count()
        m = cport.Mesh(npd)
        A = np.empty(m.mshape)
        h2 = (2.0 / (n + 1)) ** 2
        for s in range(7):
            A[s] = (6.0 if s == 3 else -1.0) / h2
        x = m.zeros()
        x[1:-1, 1:-1, 1:-1] = golden_field(n ** 3, (n, n, n))
        cport.apply(m, A, x)
        best = 1e300
        for _ in range(reps):
            t0 = time.perf_counter(); y = cport.apply(m, A, x); best = min(best, time.perf_counter() - t0)
        kind = "port"
        chk = float(np.abs(y).max())
    gbs = BYTES_SPMV * n ** 3 / best / 1e9
    return {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind, "ms": best * 1e3, "check_max_abs_y": chk,
            "sample": f"Poisson {n}^3 fp64, SMat::apply (y = A x), best of {reps} after 1 warm-up, OMP threads = {cores}"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = a.size
    t0 = time.time()
    cb = cpu_spmv(n, max(1, a.steps))
    line = {"impl": "reference", "metric": "spmv_7pt_fp64_hbm_gbs", "value": cb["value"], "unit": "GB/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": cb["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": f"poisson3d_uniform_{n}^3_fp64_7pt_spmv", "grid": f"{n}^3 unknowns (npdim {n + 2})",
                                             "note": "reference CPU path on host cores; the reference is single-process, so every N reports the same single-host number"},
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t0}
    _emit(json.dumps(line))


# ---------------------------------------------------------------------------------------------- bandwidth sweep (BASELINE.json configs[4])

def run_sweep(a):
    """SpMV / axpy / dot bandwidth, 32^3 .. 1024^3 fp64: this library on cuda:0 (CUDA events) against the HBM roofline, and the
    reference's CPU kernels (oracle/_ref, all host threads) on the same grids up to --sweep-cpu-max.  Writes a markdown table."""
    from struct3d_b200 import capi
    from oracle import refshim
    sizes = [int(v) for v in a.sweep_sizes.split(",")]
    peak, _src = peaks()
    ctx = capi.Context(0)
    rows = []
    for n in sizes:
        g = capi.make_grid(n, n, n)
        coords = [-1.0 + 2.0 * np.arange(n + 2) / (n + 1)] * 3
        dA = ctx.mat_alloc(g); ctx.assemble_poisson(g, coords, dA)
        dx, dy = ctx.vec_alloc(g), ctx.vec_alloc(g)
        ctx.vecset(g, 0.5, dx); ctx.vecset(g, 0.25, dy)
        red = ctx.scalars_alloc(2)
        reps = 20 if n >= 256 else 200
        def t(fn):
            for _ in range(3): fn()
            ctx.sync(); ctx.timer_start()
            for _ in range(reps): fn()
            return ctx.timer_stop_ms() / reps
        gpu = {"spmv": BYTES_SPMV * n ** 3 / t(lambda: ctx.spmv(g, dA, dx, dy)) / 1e6,
               "axpy": 24 * n ** 3 / t(lambda: ctx.vecaxpy(g, 1e-9, dx, dy)) / 1e6,
               "dot": 16 * n ** 3 / t(lambda: ctx.dot_dev(g, dx, dy, red)) / 1e6}
        for p_ in (dA, dx, dy, red): ctx.free(p_)
        cpu = {"spmv": None, "axpy": None, "dot": None}
        cores = None
        if refshim.available() and n <= a.sweep_cpu_max:
            cores = refshim.num_threads()
            m, _t = refshim._capture_stdout(lambda: refshim.Mesh((n + 2,) * 3))
            A, x, y = refshim.Mat(m), refshim.Vec(m), refshim.Vec(m)
            h2 = (2.0 / (n + 1)) ** 2
            for s_ in range(7):
                A.a[s_][...] = (6.0 if s_ == 3 else -1.0) / h2
            x.a[1:-1, 1:-1, 1:-1] = 0.5; y.a[1:-1, 1:-1, 1:-1] = 0.25
            r = 3 if n >= 256 else 20
            cpu["spmv"] = BYTES_SPMV * n ** 3 / refshim.time_apply(A, x, y, r) / 1e9
            cpu["axpy"] = 24 * n ** 3 / refshim.time_axpy(x, y, r) / 1e9
            cpu["dot"] = 16 * n ** 3 / refshim.time_dot(x, y, r) / 1e9
            del A, x, y, m
        rows.append((n, gpu, cpu, cores))
        print(f"[sweep] {n}^3 gpu {gpu} cpu {cpu}", file=sys.stderr, flush=True)
    out = ["| grid | working set | SpMV GB/s (% of measured peak) | axpy GB/s (%) | dot GB/s (%) | reference CPU SpMV / axpy / dot GB/s (threads) |", "|---|---|---|---|---|---|"]
    for n, gpu, cpu, cores in rows:
        ws = (7 * 8 + 2 * 8) * n ** 3
        note = "L2-resident" if ws < 100e6 else "streams from HBM"
        cp = " / ".join("-" if cpu[k] is None else f"{cpu[k]:.1f}" for k in ("spmv", "axpy", "dot")) + (f" ({cores})" if cores else "")
        out.append(f"| {n}^3 | {ws / 1e6:.1f} MB, {note} | " + " | ".join(f"{gpu[k]:.0f} ({100 * gpu[k] / peak:.0f} %)" for k in ("spmv", "axpy", "dot")) + f" | {cp} |")
    text = "\n".join(out)
    os.makedirs(os.path.dirname(a.sweep_out), exist_ok=True)
    with open(a.sweep_out, "w") as fh:
        fh.write(text + "\n")
    print(text, file=sys.stderr)
    _emit(json.dumps({"sweep": [{"n": n, "gpu_gbs": gpu, "cpu_gbs": cpu, "cpu_threads": cores} for n, gpu, cpu, cores in rows], "peak_gbs": peak}))


# ---------------------------------------------------------------------------------------------- solve comparison (SURVEY 8d "CPU baseline")

def run_solves(a):
    """Whole solves, createSolver -> updateOperator -> apply, on the reference's as-is cases and on the 256^3 benchmark cases: the
    reference itself on the host cores (oracle/_ref; CG/BiCGSTAB = the harness loops over reference primitives, SGS sequential) beside
    this library on cuda:0.  Same grids, same right-hand sides, same options; iteration counts must agree (BiCGSTAB within 5 %)."""
    from struct3d_b200 import hostapi as H
    from oracle import refshim as R
    H.init(0)
    V = (0.577350269189626, 0.7, 0.3)
    cases = [("test/input/poisson.control as-is", (16, 16, 16), "chebyshev", "poisson", V, 0.0, "test_rhs", "richardson", "jacobi"),
             ("test/input/poisson.control as-is", (16, 16, 16), "chebyshev", "poisson", V, 0.0, "test_rhs", "gcr", "jacobi"),
             ("structsolvers/input/poisson.control", (26, 26, 26), "uniform", "poisson", V, 0.0, "manufactured_rhs", "richardson", "jacobi"),
             ("structsolvers/input/convdiff.control", (48, 48, 48), "uniform", "convdiff", V, 0.1, "manufactured_rhs", "gcr", "jacobi"),
             ("structsolvers/input/convdiff.control", (48, 48, 48), "uniform", "convdiff", V, 0.1, "manufactured_rhs", "richardson", "sgs"),
             ("structsolvers/input/convdiff-large.control", (64, 64, 64), "uniform", "convdiff", V, 0.1, "manufactured_rhs", "gcr", "sgs"),
             ("structsolvers/input/convdiff-large.control", (64, 64, 64), "uniform", "convdiff", V, 0.1, "manufactured_rhs", "gcr", "strilu"),
             ("Poisson 64^3 (configs[0]), golden rhs", (66, 66, 66), "uniform", "poisson", V, 0.0, "golden", "cg", "jacobi"),
             ("convection-diffusion 64^3", (66, 66, 66), "uniform", "convdiff", V, 0.1, "manufactured_rhs", "bcgs", "sgs")]
    if a.solves_large:
        n = a.solves_large + 2
        cases += [(f"Poisson {a.solves_large}^3 (configs[1]), golden rhs", (n, n, n), "uniform", "poisson", V, 0.0, "golden", "cg", "jacobi"),
                  (f"convection-diffusion {a.solves_large}^3 (configs[2])", (n, n, n), "uniform", "convdiff", V, 0.1, "manufactured_rhs", "bcgs", "sgs")]
    cores = R.num_threads() if R.available() else 0
    out = [f"| case | grid | solver | reference CPU: iterations, s ({cores} threads; SGS/StrILU sequential) | this library, 1 B200: iterations, s | speed-up |", "|---|---|---|---|---|---|"]
    rows = []
    for name, npd, grid, pde, vel, mu, rhs, ksp, pc in cases:
        opts = {"-s3d_ksp_type": ksp, "-s3d_pc_type": pc, "-ksp_rtol": 1e-6, "-ksp_max_it": 5000, "-s3d_pc_apply_sweeps": 1, "-s3d_pc_build_sweeps": 1,
                "-s3d_thread_chunk_size": 1024, "-s3d_pc_use_threaded_apply": "false", "-s3d_pc_use_threaded_build": "false", "-ksp_restart": 30}
        nreal = tuple(v - 2 for v in npd)
        gold = np.zeros(npd[::-1]); gold[1:-1, 1:-1, 1:-1] = golden_field(nreal[0] * nreal[1] * nreal[2], nreal[::-1])
        # --- reference
        ref = None
        if R.available():
            m, _ = R._capture_stdout(lambda: R.Mesh(npd, grid=grid))
            P = R.Pde(pde, vel, mu)
            (A, b), _ = R._capture_stdout(lambda: (P.lhs(m), P.vector(m, "test_rhs" if rhs == "golden" else rhs)))
            if rhs == "golden":
                b.a[...] = gold
            x = R.Vec(m)
            t0 = time.perf_counter()
            if ksp in ("cg", "bcgs"):
                pr = R.Prec(A, pc)
                info, _h = R.harness(A, pr, b, x, "cg" if ksp == "cg" else "bicgstab", 1e-6, 5000)
            else:
                info = R.solve(A, b, x, opts)
            ref = (info["iters"], time.perf_counter() - t0, info["converged"])
            del A, b, x, P, m
        # --- ours
        m = H.Mesh(npd, grid=grid)
        P = H.Pde(pde, vel, mu)
        A = P.lhs(m)
        if rhs == "golden":
            b = H.Vec(m); b.set(gold)
        else:
            b = P.vector(m, rhs)
        x = H.Vec(m)
        H.set_options(opts)
        sv = H.Solver(A); sv.update(); H.sync()
        best, info = 1e30, None
        for _ in range(2):
            H.vecset(0.0, x)
            t0 = time.perf_counter(); info = sv.apply(b, x); H.sync(); best = min(best, time.perf_counter() - t0)
        del sv
        rows.append({"case": name, "npdim": list(npd), "ksp": ksp, "pc": pc, "reference": ref, "ours": (info["iters"], best, info["converged"])})
        r = "-" if ref is None else f"{ref[0]}, {ref[1]:.4f}"
        sp = "-" if ref is None else f"{ref[1] / best:.1f}x"
        out.append(f"| {name} | {nreal[0]}x{nreal[1]}x{nreal[2]} {grid} | {ksp} + {pc} | {r} | {info['iters']}, {best:.4f} | {sp} |")
        print(out[-1], file=sys.stderr, flush=True)
    text = "\n".join(out)
    os.makedirs(os.path.dirname(a.solves_out), exist_ok=True)
    with open(a.solves_out, "w") as fh:
        fh.write(text + "\n")
    _emit(json.dumps({"solves": rows, "cpu_threads": cores}))


# ---------------------------------------------------------------------------------------------- our arm

def run_ours(a):
    import torch
    import torch.distributed as dist
    from struct3d_b200 import capi, hostapi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    nccl_id = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        obj = [capi.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        nccl_id = obj[0]
    hostapi.init(local, rank, world, nccl_id)
    ctx = capi.Context.from_handle(hostapi.cuda_ctx())
    H = hostapi
    ctx.tune("halo_overlap", a.halo_overlap)
    ctx.tune("halo_p2p", a.halo_p2p)

    def barrier():
        H.sync()
        if world > 1:
            dist.barrier()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = a.size
    nzg = n * world                                   # weak scaling: 512 planes per GPU
    null = open(os.devnull, "w")
    mesh, _ = H._capture_stdout(lambda: H.Mesh((n + 2, n + 2, nzg + 2)))
    pde, _ = H._capture_stdout(lambda: H.Pde("poisson"))
    t0 = time.perf_counter()
    A = pde.lhs(mesh)
    H.sync()
    t_assemble = time.perf_counter() - t0
    x, y = H.Vec(mesh), H.Vec(mesh)
    nx, ny, nzl = mesh.local
    N_local = nx * ny * nzl
    N_total = nx * ny * nzg
    xh = np.zeros(mesh.vshape)
    xh[1:-1, 1:-1, 1:-1] = golden_field(N_total, (nzg, ny, nx))[mesh.k0:mesh.k0 + nzl]
    x.set(xh)
    g = capi.make_grid(nx, ny, nzg, rank, world) if world > 1 else capi.make_grid(nx, ny, nzl)

    # ---- device-resident steps
    for _ in range(max(3, a.warmup)):
        H.apply(A, x, y)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(a.steps):
        H.apply(A, x, y)
    ms_total = ctx.timer_stop_ms()
    launches = ctx.launch_count() - l0
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = max_over_ranks(ms_total / a.steps)
    value = BYTES_SPMV * N_total / (ms_step * 1e-3) / 1e9

    # kernel-only duration on this rank (no halo exchange): the roofline numerator
    dA, dx, dy = A.device_ptr(), x.device_ptr(), y.device_ptr()
    for _ in range(3):
        ctx.spmv(g, dA, dx, dy)
    H.sync()
    ctx.timer_start()
    for _ in range(a.steps):
        ctx.spmv(g, dA, dx, dy)
    ms_kernel = ctx.timer_stop_ms() / a.steps
    peak, peak_src = peaks()
    achieved = BYTES_SPMV * N_local / (ms_kernel * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get("spmv_march_512", {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "spmv_march_kernel<BY=8,MINB=3,plain> (2-plane z-march, 2 points/thread)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": BYTES_SPMV * N_local, "avg_launch_ms": ms_kernel}

    # ---- end to end: host vectors in, host vectors out, every step
    nbytes = xh.nbytes
    hx_p, hy_p = ctx.host_alloc(nbytes), ctx.host_alloc(nbytes)
    import ctypes as C
    hx = np.ctypeslib.as_array(C.cast(hx_p, C.POINTER(C.c_double)), shape=(xh.size,)).reshape(xh.shape)
    hy = np.ctypeslib.as_array(C.cast(hy_p, C.POINTER(C.c_double)), shape=(xh.size,)).reshape(xh.shape)
    hx[...] = xh

    def e2e_step():
        if world == 1:
            # SMat::apply for host vectors: upload of x, multiply and download of y pipelined in z-chunks (s3d_spmv_host)
            ctx.spmv_host(g, dA, hx_p, hy_p, dx, dy, nchunks=a.e2e_chunks)
        else:
            ctx.call("s3d_vec_upload", C.byref(g), capi._ptr(dx), capi._ptr(hx_p))
            H.apply(A, x, y)
            ctx.call("s3d_vec_download", C.byref(g), capi._ptr(dy), capi._ptr(hy_p))

    x.device_ptr()
    for _ in range(2):
        e2e_step()
    barrier()
    e2e_steps = max(2, min(a.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    H.sync()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
    e2e = {"value": BYTES_SPMV * N_total / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(nbytes),
           "ms_per_step": e2e_ms, "steps": e2e_steps, "check_max_abs_y": float(np.abs(hy).max()),
           "note": "per step: x from pinned host memory -> y = A x -> y into pinned host memory; single GPU: pipelined in z-chunks over three streams (s3d_spmv_host), multi-GPU: upload, halo exchange + SMat::apply, download"}
    ctx.host_free(hx_p); ctx.host_free(hy_p)

    # ---- extras: the solvers through the reference's call sequence (not part of the timed steps)
    solve = {}
    if not a.no_solve:
        try:
            b = H.Vec(mesh)
            b.set(xh)
            u = H.Vec(mesh)
            opts = {"-s3d_ksp_type": "cg", "-s3d_pc_type": "jacobi", "-ksp_rtol": 1e-6, "-ksp_max_it": a.cg_max_it}
            H.set_options(opts)
            (s, _) = H._capture_stdout(lambda: H.Solver(A))
            s.update()
            barrier()
            t0 = time.perf_counter()
            info, _ = H._capture_stdout(lambda: s.apply(b, u))
            H.sync()
            dt = max_over_ranks(time.perf_counter() - t0)
            solve["cg_jacobi_poisson"] = {"grid": f"{nx}x{ny}x{nzg}", "rhs": "golden non-eigenvector field (SURVEY 8d)", "rtol": 1e-6, "seconds": dt,
                                          "iters": info["iters"], "converged": info["converged"], "rel_res": info["resnorm"] / H.norm_vector_l2(b),
                                          "ms_per_iter": dt * 1e3 / max(1, info["iters"]),
                                          "credited_gbs": BYTES_CG_ITER * N_total * info["iters"] / dt / 1e9, "moved_bytes_per_point_iter": 160}
            del s, u, b
            # north_star's scaling target: CG+Jacobi Poisson on a FIXED 1024^3 grid (strong scaling over the z-slabs),
            # a fixed number of iterations so that every N does the same work; the driver/judge derives the efficiency
            if a.strong > 0 and a.strong >= world:
                ns = a.strong
                del y
                ms_, _ = H._capture_stdout(lambda: H.Mesh((ns + 2, ns + 2, ns + 2)))
                As = pde.lhs(ms_)
                bs_, us_ = H.Vec(ms_), H.Vec(ms_)
                nzs = ms_.local[2]
                chunk = np.zeros((nzs + 2, ns + 2, ns + 2))
                chunk[1:-1, 1:-1, 1:-1] = golden_field(ns * ns * nzs, (nzs, ns, ns))   # any non-eigenvector field will do for timing
                bs_.set(chunk)
                del chunk
                def timed_solve(its):
                    H.set_options({"-s3d_ksp_type": "cg", "-s3d_pc_type": "jacobi", "-ksp_rtol": 1e-300, "-ksp_max_it": its})
                    (ss, _) = H._capture_stdout(lambda: H.Solver(As))
                    ss.update()
                    H.vecset(0.0, us_)
                    barrier()
                    t0 = time.perf_counter()
                    info_, _ = H._capture_stdout(lambda: ss.apply(bs_, us_))
                    H.sync()
                    return max_over_ranks(time.perf_counter() - t0), info_["iters"]

                timed_solve(3)                                        # warm-up (kernel loading, work-vector allocation)
                ta, ia = timed_solve(a.strong_iters // 2)
                tb, ib = timed_solve(a.strong_iters + a.strong_iters // 2)
                per_iter = (tb - ta) / max(1, ib - ia)                # set-up and the initial residual cancel in the difference
                solve["cg_jacobi_poisson_strong"] = {"grid_total": f"{ns}^3", "grid_per_gpu": f"{ns}x{ns}x{nzs}", "scaling": "strong",
                                                     "ms_per_iter": per_iter * 1e3, "iters": [ia, ib], "seconds": [ta, tb],
                                                     "credited_gbs": BYTES_CG_ITER * float(ns) ** 3 / per_iter / 1e9,
                                                     "moved_gbs": 160 * float(ns) ** 3 / per_iter / 1e9,
                                                     "note": "fixed iteration counts (rtol disabled); ms_per_iter = (t(ib) - t(ia)) / (ib - ia), identical work at every GPU count"}
                del us_, bs_, As
            if world > 1 and n >= 256:
                # BiCGSTAB + exact (lexicographic) SGS THROUGH the z-slabs (peer-memory pipeline): 256 x 256 x 256 N convection-diffusion
                mx, _ = H._capture_stdout(lambda: H.Mesh((258, 258, 256 * world + 2)))
                px, _ = H._capture_stdout(lambda: H.Pde("convdiff", (0.577350269189626, 0.7, 0.3), 0.1))
                Ax, bx, ux = px.lhs(mx), px.vector(mx, "manufactured_rhs"), H.Vec(mx)
                H.set_options({"-s3d_ksp_type": "bcgs", "-s3d_pc_type": "sgs", "-ksp_rtol": 1e-6, "-ksp_max_it": 5000, "-s3d_pc_apply_sweeps": 1,
                               "-s3d_thread_chunk_size": 1024, "-s3d_pc_use_threaded_apply": "false"})
                (sx, _) = H._capture_stdout(lambda: H.Solver(Ax))
                sx.update()
                barrier()
                t0 = time.perf_counter()
                ix, _ = H._capture_stdout(lambda: sx.apply(bx, ux))
                H.sync()
                dtx = max_over_ranks(time.perf_counter() - t0)
                solve["bicgstab_sgs_lexicographic_slabs_convdiff256perGPU"] = {"grid": f"256x256x{256 * world}", "seconds": dtx, "iters": ix["iters"],
                                                                               "converged": ix["converged"], "ms_per_iter": dtx * 1e3 / max(1, ix["iters"]),
                                                                               "ordering": "lexicographic through the slabs" if a.halo_p2p else "red-black (peer memory switched off)",
                                                                               "precapply_seconds": ix["precapplywtime"]}
                del sx, Ax, bx, ux
            if world == 1 and n >= 256:
                m2, _ = H._capture_stdout(lambda: H.Mesh((258, 258, 258)))
                p2, _ = H._capture_stdout(lambda: H.Pde("convdiff", (0.577350269189626, 0.7, 0.3), 0.1))
                A2, b2, u2 = p2.lhs(m2), p2.vector(m2, "manufactured_rhs"), H.Vec(m2)
                for pc, extra in (("sgs", {"-s3d_sgs_ordering": "lexicographic"}), ("sgs", {"-s3d_sgs_ordering": "redblack"}),
                                  ("sgs", {"-s3d_sgs_ordering": "sweeps", "-s3d_pc_apply_sweeps": 2}),
                                  ("sgs", {"-s3d_sgs_ordering": "sweeps", "-s3d_pc_apply_sweeps": 4}), ("jacobi", {})):
                    o = {"-s3d_ksp_type": "bcgs", "-s3d_pc_type": pc, "-ksp_rtol": 1e-6, "-ksp_max_it": 5000, "-s3d_pc_apply_sweeps": 1,
                         "-s3d_thread_chunk_size": 1024, "-s3d_pc_use_threaded_apply": "false"}
                    o.update(extra)
                    H.set_options(o)
                    H.vecset(0.0, u2)
                    (s2, _) = H._capture_stdout(lambda: H.Solver(A2))
                    s2.update()
                    H.sync()
                    t0 = time.perf_counter()
                    i2, _ = H._capture_stdout(lambda: s2.apply(b2, u2))
                    H.sync()
                    dt2 = time.perf_counter() - t0
                    key = "bicgstab_" + pc + ("_" + extra["-s3d_sgs_ordering"] if extra else "") + (str(extra["-s3d_pc_apply_sweeps"]) if "-s3d_pc_apply_sweeps" in extra else "") + "_convdiff256"
                    solve[key] = {"seconds": dt2, "iters": i2["iters"], "converged": i2["converged"], "precapply_seconds": i2["precapplywtime"]}
                    del s2
                del A2, b2, u2
                if n >= 512:
                    # the metric's own size (BASELINE.json: "CG/BiCGSTAB solve s at 512^3"): BiCGSTAB + exact lexicographic SGS
                    m3, _ = H._capture_stdout(lambda: H.Mesh((514, 514, 514)))
                    A3, b3, u3 = p2.lhs(m3), p2.vector(m3, "manufactured_rhs"), H.Vec(m3)
                    H.set_options({"-s3d_ksp_type": "bcgs", "-s3d_pc_type": "sgs", "-ksp_rtol": 1e-6, "-ksp_max_it": 5000, "-s3d_pc_apply_sweeps": 1,
                                   "-s3d_thread_chunk_size": 1024, "-s3d_pc_use_threaded_apply": "false", "-s3d_sgs_ordering": "lexicographic"})
                    (s3, _) = H._capture_stdout(lambda: H.Solver(A3))
                    s3.update()
                    H.sync()
                    t0 = time.perf_counter()
                    i3, _ = H._capture_stdout(lambda: s3.apply(b3, u3))
                    H.sync()
                    dt3 = time.perf_counter() - t0
                    solve["bicgstab_sgs_lexicographic_convdiff512"] = {"seconds": dt3, "iters": i3["iters"], "converged": i3["converged"],
                                                                       "precapply_seconds": i3["precapplywtime"], "ms_per_iter": dt3 * 1e3 / max(1, i3["iters"])}
                    del s3, A3, b3, u3
        except Exception as e:  # the headline numbers above are measured already: report the failure instead of losing the line
            solve["error"] = f"{type(e).__name__}: {e}"[:400]

    cb = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cb = cpu_spmv(n, 5)
    if rank == 0:
        line = {"metric": "spmv_7pt_fp64_hbm_gbs", "value": value, "unit": "GB/s", "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"poisson3d_uniform_{n}^3_fp64_7pt_spmv", "grid_per_gpu": f"{nx}x{ny}x{nzl}", "grid_total": f"{nx}x{ny}x{nzg}",
                           "decomposition": (f"z-slabs x{world}, one-plane halo " + ("pushed into the neighbours' peer-memory windows (CUDA IPC over NVLink) and read by the SpMV itself" if a.halo_p2p else "by NCCL send/recv")) if world > 1 else "single GPU",
                           "l2": f"inputs exceed L2: {BYTES_SPMV * N_local / 1e9:.2f} GB streamed per step vs 126 MB L2 (no flush needed)",
                           "step": "halo exchange (N>1) + SMat::apply (y = A x)", "bytes_per_point": BYTES_SPMV},
                "value_frac_of_measured_peak": value / (peak * world), "value_frac_of_nominal_8TBs": value / (8000.0 * world),
                "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "assembly_seconds": t_assemble, "solve": solve}
        if cb is not None:
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        _emit(json.dumps(line))
    barrier()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512, help="unknowns per axis (per GPU along z)")
    ap.add_argument("--strong", type=int, default=1024, help="grid size of the strong-scaling CG block (0 disables)")
    ap.add_argument("--strong-iters", type=int, default=40)
    ap.add_argument("--halo-overlap", type=int, default=0, help="0: exchange halos first, then multiply (A/B switch)")
    ap.add_argument("--halo-p2p", type=int, default=1, help="1: halo planes through peer-memory windows (CUDA IPC), 0: NCCL send/recv (A/B switch)")
    ap.add_argument("--e2e-chunks", type=int, default=32, help="z-chunks of the pipelined host-buffer SpMV (e2e)")
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--cg-max-it", type=int, default=20000, help="iteration cap of the CG solve of the solve block (A/B runs shorten it)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--solves", action="store_true", help="whole solves: the reference on the host cores beside this library (SURVEY 8d)")
    ap.add_argument("--solves-large", type=int, default=256, help="grid size of the two benchmark cases in --solves (0 = skip them)")
    ap.add_argument("--solves-out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpurun_out", "solves.md"))
    ap.add_argument("--sweep", action="store_true", help="bandwidth sweep over grid sizes instead of the bench line (BASELINE.json configs[4])")
    ap.add_argument("--sweep-sizes", default="32,64,128,256,512,1024")
    ap.add_argument("--sweep-cpu-max", type=int, default=512, help="largest grid the reference CPU kernels are timed on")
    ap.add_argument("--sweep-out", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpurun_out", "bw_sweep.md"))
    a = ap.parse_args()
    # stdout carries exactly ONE line, the JSON record: everything else (NCCL's version banner, the host classes'
    # progress prints) goes to stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    global _emit
    _emit = lambda line: os.write(real_stdout, (line + "\n").encode())
    if a.sweep:
        run_sweep(a)
    elif a.solves:
        run_solves(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
