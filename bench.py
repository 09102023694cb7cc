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


def golden_slab(nx, ny, k0, nzl):
    """Planes [k0, k0 + nzl) of the golden field of an nx x ny x * grid: the same values whatever the decomposition."""
    idx = np.arange(k0 * nx * ny + 1, (k0 + nzl) * nx * ny + 1, dtype=np.float64) * 0.6180339887
    return (idx - np.floor(idx) - 0.5).reshape(nzl, ny, nx)


# ---------------------------------------------------------------------------------------------- reference arm

def host_threads():
    """Threads the reference's OpenMP path may use on this box: the cores this process is allowed on.  torchrun exports
    OMP_NUM_THREADS=1 to its workers; that must not turn 'the reference on the host cores' into one core."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def bench_config(n, world):
    """The workload, described the same way by both arms (identical keys and values)."""
    return {"workload": f"poisson3d_uniform_{n}^3_fp64_7pt_spmv", "grid_per_gpu": f"{n}x{n}x{n}", "grid_total": f"{n}x{n}x{n * world}",
            "step": "SMat::apply (y = A x) over the whole grid, incl. the one-plane halo exchange where the grid is split into z-slabs",
            "bytes_per_point": BYTES_SPMV,
            "l2": f"inputs exceed L2: {BYTES_SPMV * n ** 3 / 1e9:.2f} GB streamed per step and GPU vs 126 MB L2 (no flush needed)"}


def cpu_spmv(n, reps, threads=None, perf=True):
    """Times the reference's own CPU SMat::apply (or the C port when oracle/_ref is absent) on Poisson n^3, on `threads` host threads."""
    from oracle import cport, refshim
    npd = (n + 2,) * 3
    threads = threads or host_threads()
    if refshim.available():
        build = refshim.use_perf_build() if perf else "parity build"
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
        build = "oracle/s3d_oracle.c (C port; oracle/_ref was not built)"
        os.environ["OMP_NUM_THREADS"] = str(threads)
        cores = threads
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
    return {"value": gbs, "unit": "GB/s", "cores": cores, "kind": kind, "ms": best * 1e3, "check_max_abs_y": chk, "build": build,
            "sample": f"one {n}^3 slab of the grid (Poisson, fp64), SMat::apply (y = A x), best of {reps} after 1 warm-up, {cores} OpenMP threads "
                      f"(all cores this process may use); GB/s per point is independent of the slab count; build: {build}"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = a.size
    t0 = time.time()
    cb = cpu_spmv(n, max(1, a.steps))
    line = {"impl": "reference", "metric": "spmv_7pt_fp64_hbm_gbs", "value": cb["value"], "unit": "GB/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": cb["ms"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": bench_config(n, a.gpus),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference's native path is single-process (src/cartmesh.cpp:107-109 refuses MPI size > 1): one host, all its cores, at every N",
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
        def tg(fn):
            # the same call replayed from a captured CUDA graph of 20 launches: how the solver loops issue their kernels on grids up to
            # 4 Mi points (one launch from Python costs ~7 us of host time, more than these kernels run)
            for _ in range(3): fn()
            ctx.graph_begin()
            for _ in range(20): fn()
            h = ctx.graph_end()
            ctx.graph_launch(h)
            ctx.sync(); ctx.timer_start()
            for _ in range(10): ctx.graph_launch(h)
            ms = ctx.timer_stop_ms() / 200
            ctx.graph_destroy(h)
            return ms
        gpu = {"spmv": BYTES_SPMV * n ** 3 / t(lambda: ctx.spmv(g, dA, dx, dy)) / 1e6,
               "axpy": 24 * n ** 3 / t(lambda: ctx.vecaxpy(g, 1e-9, dx, dy)) / 1e6,
               "dot": 16 * n ** 3 / t(lambda: ctx.dot_dev(g, dx, dy, red)) / 1e6}
        if n <= 128:
            gpu["spmv_graph"] = BYTES_SPMV * n ** 3 / tg(lambda: ctx.spmv(g, dA, dx, dy)) / 1e6
            gpu["axpy_graph"] = 24 * n ** 3 / tg(lambda: ctx.vecaxpy(g, 1e-9, dx, dy)) / 1e6
            gpu["dot_graph"] = 16 * n ** 3 / tg(lambda: ctx.dot_dev(g, dx, dy, red)) / 1e6
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
    out = ["| grid | working set | SpMV GB/s (% of measured peak) | axpy GB/s (%) | dot GB/s (%) | same three, replayed from a CUDA graph | reference CPU SpMV / axpy / dot GB/s (threads) |", "|---|---|---|---|---|---|---|"]
    for n, gpu, cpu, cores in rows:
        ws = (7 * 8 + 2 * 8) * n ** 3
        note = "L2-resident" if ws < 100e6 else "streams from HBM"
        cp = " / ".join("-" if cpu[k] is None else f"{cpu[k]:.1f}" for k in ("spmv", "axpy", "dot")) + (f" ({cores})" if cores else "")
        gr = " / ".join(f"{gpu[k + '_graph']:.0f}" for k in ("spmv", "axpy", "dot")) if "spmv_graph" in gpu else "-"
        out.append(f"| {n}^3 | {ws / 1e6:.1f} MB, {note} | " + " | ".join(f"{gpu[k]:.0f} ({100 * gpu[k] / peak:.0f} %)" for k in ("spmv", "axpy", "dot")) + f" | {gr} | {cp} |")
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
    import ctypes as C
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
    for kv in a.tune:                                   # A/B switches: --tune halo_push_side=0 --tune allreduce_p2p=0 ...
        k, v = kv.split("=")
        ctx.tune(k, int(v))
    V = (0.577350269189626, 0.7, 0.3)

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

    def quiet(fn):
        return H._capture_stdout(fn)[0]

    n = a.size
    nzg = n * world                                   # weak scaling: 512 planes per GPU
    mesh = quiet(lambda: H.Mesh((n + 2, n + 2, nzg + 2)))
    pde = quiet(lambda: H.Pde("poisson"))
    t0 = time.perf_counter()
    A = pde.lhs(mesh)
    H.sync()
    t_assemble = time.perf_counter() - t0
    x, y = H.Vec(mesh), H.Vec(mesh)
    nx, ny, nzl = mesh.local
    N_local = nx * ny * nzl
    N_total = nx * ny * nzg
    gold = golden_field(N_total, (nzg, ny, nx))
    xh = np.zeros(mesh.vshape)
    xh[1:-1, 1:-1, 1:-1] = gold[mesh.k0:mesh.k0 + nzl]
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

    dA, dx, dy = A.device_ptr(), x.device_ptr(), y.device_ptr()
    # ---- N > 1 self-check of the exchange: every rank recomputes its slab with the ghost planes filled from the KNOWN global field,
    #      through the plain single-GPU kernel, and compares bit for bit with what the exchanging path produced
    halo_check = None
    if world > 1:
        xg = xh.copy()
        if mesh.k0 > 0:
            xg[0, 1:-1, 1:-1] = gold[mesh.k0 - 1]
        if mesh.k0 + nzl < nzg:
            xg[-1, 1:-1, 1:-1] = gold[mesh.k0 + nzl]
        x2, y2 = H.Vec(mesh), H.Vec(mesh)
        x2.set(xg)
        H.apply(A, x, y)
        g1 = capi.Grid.from_buffer_copy(g)    # the slab's own layout (slab grids carry slack planes in cstride), seen as an undecomposed grid
        g1.rank, g1.nranks, g1.k0, g1.nz_global = 0, 1, 0, nzl
        ctx.spmv(g1, dA, x2.device_ptr(), y2.device_ptr())
        ctx.vecaxpy(g1, -1.0, dy, y2.device_ptr())
        diff = ctx.nrm2(g1, y2.device_ptr())
        ynorm = ctx.nrm2(g1, dy)
        bad = max_over_ranks(0.0 if (diff == 0.0 and ynorm > 0.0) else 1.0)
        halo_check = "bit-identical" if bad == 0.0 else "MISMATCH"
        del x2, y2, xg
    del gold

    # kernel-only duration on this rank (no halo exchange): the roofline numerator
    for _ in range(3):
        ctx.spmv(g, dA, dx, dy)
    H.sync()
    ctx.timer_start()
    for _ in range(a.steps):
        ctx.spmv(g, dA, dx, dy)
    ms_kernel = ctx.timer_stop_ms() / a.steps
    peak, peak_src = peaks()
    achieved = BYTES_SPMV * N_local / (ms_kernel * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        ent = tj.get("spmv_march_512_default") or tj.get("spmv_march_512", {})
        traffic, traffic_src = ent.get("dram_bytes_per_launch"), ent.get("source")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "spmv_march_kernel<BY=8,MINB=3,plain> (2-plane z-march, 2 points/thread)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "frac_of_nominal_8TBs": achieved / 8000.0, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": BYTES_SPMV * N_local, "avg_launch_ms": ms_kernel}

    # ---- end to end: host vectors in, host vectors out, every step
    nbytes = xh.nbytes
    hx_p, hy_p = ctx.host_alloc(nbytes), ctx.host_alloc(nbytes)
    hx = np.ctypeslib.as_array(C.cast(hx_p, C.POINTER(C.c_double)), shape=(xh.size,)).reshape(xh.shape)
    hy = np.ctypeslib.as_array(C.cast(hy_p, C.POINTER(C.c_double)), shape=(xh.size,)).reshape(xh.shape)
    hx[...] = xh

    def e2e_step():
        # SMat::apply for host vectors: upload of x, (halo planes first on a z-slab), multiply and download of y pipelined in z-chunks
        ctx.spmv_host(g, dA, hx_p, hy_p, dx, dy, nchunks=a.e2e_chunks)

    for _ in range(2):
        e2e_step()
    e2e_same = None
    if world > 1:
        H.apply(A, x, y)                                  # device-resident result of the same x
        ref_y = y.a.copy()
        e2e_step()
        e2e_same = bool(np.array_equal(hy[1:-1], ref_y[1:-1]))
        del ref_y
    barrier()
    e2e_steps = max(2, min(a.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    H.sync()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e2e_steps)
    e2e = {"value": BYTES_SPMV * N_total / (e2e_ms * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(nbytes), "d2h_bytes_per_step": int(nbytes),
           "ms_per_step": e2e_ms, "steps": e2e_steps, "check_max_abs_y": float(np.abs(hy).max()),
           "note": "per step and rank: x from pinned host memory -> y = A x -> y into pinned host memory, pipelined in z-chunks over three streams (s3d_spmv_host); "
                   "on z-slabs the two boundary planes go up first and travel through the neighbours' peer-memory windows"}
    if e2e_same is not None:
        e2e["equals_device_resident_result"] = e2e_same
    ctx.host_free(hx_p); ctx.host_free(hy_p)
    del hx, hy

    # ---- extras: the solvers through the reference's call sequence (not part of the timed steps)
    solve, summary = {}, {}

    def timed_solve(Am, bv, uv, opts, repeats=1):
        """createSolver -> updateOperator -> apply; wall time of apply, max over ranks; best of `repeats` (each from a zero guess)."""
        H.set_options(opts)
        sv = quiet(lambda: H.Solver(Am))
        sv.update()
        best, info = None, None
        for _ in range(repeats):
            H.vecset(0.0, uv)
            barrier()
            t0 = time.perf_counter()
            info = quiet(lambda: sv.apply(bv, uv))
            H.sync()
            dt = max_over_ranks(time.perf_counter() - t0)
            best = dt if best is None else min(best, dt)
        del sv
        return best, info

    SGS = {"-s3d_pc_apply_sweeps": 1, "-s3d_thread_chunk_size": 1024, "-s3d_pc_use_threaded_apply": "false"}
    if not a.no_solve:
        try:
            # (1) weak: CG + Jacobi on the bench grid (512^3 per GPU), to rtol 1e-6
            b = H.Vec(mesh); b.set(xh)
            u = H.Vec(mesh)
            dt, info = timed_solve(A, b, u, {"-s3d_ksp_type": "cg", "-s3d_pc_type": "jacobi", "-ksp_rtol": 1e-6, "-ksp_max_it": a.cg_max_it})
            solve["cg_jacobi_poisson"] = {"grid": f"{nx}x{ny}x{nzg}", "rhs": "golden non-eigenvector field (SURVEY 8d)", "rtol": 1e-6, "seconds": dt,
                                          "iters": info["iters"], "converged": info["converged"], "rel_res": info["resnorm"] / H.norm_vector_l2(b),
                                          "ms_per_iter": dt * 1e3 / max(1, info["iters"]),
                                          "credited_gbs": BYTES_CG_ITER * N_total * info["iters"] / dt / 1e9, "moved_bytes_per_point_iter": 160}
            del u, b
            del y, x, A
            # (2) BASELINE's metric: solves to rtol 1e-6 on a FIXED 512^3 grid at every GPU count (strong scaling of a whole solve)
            nf = a.fixed
            if nf > 0 and nf >= world:
                mf = quiet(lambda: H.Mesh((nf + 2, nf + 2, nf + 2)))
                Af = pde.lhs(mf)
                nzf = mf.local[2]
                bf, uf = H.Vec(mf), H.Vec(mf)
                gf = golden_field(nf ** 3, (nf, nf, nf))
                ch = np.zeros((nzf + 2, nf + 2, nf + 2)); ch[1:-1, 1:-1, 1:-1] = gf[mf.k0:mf.k0 + nzf]
                bf.set(ch); del ch, gf
                dt, info = timed_solve(Af, bf, uf, {"-s3d_ksp_type": "cg", "-s3d_pc_type": "jacobi", "-ksp_rtol": 1e-6, "-ksp_max_it": a.cg_max_it})
                solve["cg_jacobi_poisson_fixed"] = {"grid_total": f"{nf}^3", "grid_per_gpu": f"{nf}x{nf}x{nzf}", "rtol": 1e-6, "seconds": dt, "iters": info["iters"],
                                                    "converged": info["converged"], "ms_per_iter": dt * 1e3 / max(1, info["iters"])}
                summary[f"cg_jacobi_{nf}^3_solve_s"] = dt
                summary[f"cg_jacobi_{nf}^3_iters"] = info["iters"]
                del Af, bf, uf
                pcd = quiet(lambda: H.Pde("convdiff", V, 0.1))
                Ac, bc, uc = pcd.lhs(mf), pcd.vector(mf, "manufactured_rhs"), H.Vec(mf)
                o = {"-s3d_ksp_type": "bcgs", "-s3d_pc_type": "sgs", "-ksp_rtol": 1e-6, "-ksp_max_it": 5000}
                o.update(SGS)
                dt, info = timed_solve(Ac, bc, uc, o)
                solve["bicgstab_sgs_convdiff_fixed"] = {"grid_total": f"{nf}^3", "grid_per_gpu": f"{nf}x{nf}x{nzf}", "rtol": 1e-6, "seconds": dt, "iters": info["iters"],
                                                        "converged": info["converged"], "ms_per_iter": dt * 1e3 / max(1, info["iters"]),
                                                        "precapply_seconds": info["precapplywtime"],
                                                        "ordering": "lexicographic (the reference's sweep, exact)" if world == 1 else "overlapping slabs (every slab + one plane of either neighbour, cut further into concurrent blocks of planes and rows by size; no cross-rank wait)"}
                summary[f"bicgstab_sgs_{nf}^3_solve_s"] = dt
                summary[f"bicgstab_sgs_{nf}^3_iters"] = info["iters"]
                del Ac, bc, uc, mf
            # (3) north_star's scaling target: CG+Jacobi Poisson on a FIXED 1024^3 grid, a fixed number of iterations so that every N does the
            #     same work: one warm-up solve, then `strong_repeats` timed solves of `strong_iters` iterations; min and median per iteration
            if a.strong > 0 and a.strong >= world:
                ns = a.strong
                ms_ = quiet(lambda: H.Mesh((ns + 2, ns + 2, ns + 2)))
                As = pde.lhs(ms_)
                bs_, us_ = H.Vec(ms_), H.Vec(ms_)
                nzs = ms_.local[2]
                chunk = np.zeros((nzs + 2, ns + 2, ns + 2))
                chunk[1:-1, 1:-1, 1:-1] = golden_slab(ns, ns, ms_.k0, nzs)              # the global golden field: the same system at every GPU count
                bs_.set(chunk)
                del chunk
                so = {"-s3d_ksp_type": "cg", "-s3d_pc_type": "jacobi", "-ksp_rtol": 1e-300, "-ksp_max_it": a.strong_iters}
                timed_solve(As, bs_, us_, dict(so, **{"-ksp_max_it": 5}))                 # warm-up (kernel loading, work-vector allocation)
                H.set_options(so)
                ss = quiet(lambda: H.Solver(As))
                ss.update()
                per, hist = [], None
                for _ in range(a.strong_repeats):
                    H.vecset(0.0, us_)
                    barrier()
                    t0 = time.perf_counter()
                    inf = quiet(lambda: ss.apply(bs_, us_))
                    H.sync()
                    per.append(max_over_ranks(time.perf_counter() - t0) * 1e3 / max(1, inf["iters"]))
                    hist = np.asarray(inf["hist"])
                del ss
                # north_star: "matching the reference's residual history".  The reference cannot run 1024^3 CG (it has no CG and its native
                # path is one process); the stored history is this library's own one-GPU run of the same system (tests/golden/), whose
                # arithmetic is pinned to the oracle at 256^3 (tests/test_fullsize_gpu.py).  Every GPU count must reproduce it to 1e-6.
                hpath = os.path.join(ROOT, "tests", "golden", f"cg{ns}_history.json")
                hcheck = None
                if os.environ.get("S3D_WRITE_HISTORY") and world == 1 and rank == 0:
                    with open(os.environ["S3D_WRITE_HISTORY"], "w") as fh:
                        json.dump({"grid": f"{ns}^3", "rhs": "golden field (SURVEY 8d)", "solver": "cg + jacobi", "n_gpus": 1, "history": [float(v) for v in hist]}, fh)
                if os.path.exists(hpath) and hist is not None:
                    ref_h = np.asarray(json.load(open(hpath))["history"])
                    k = min(len(ref_h), len(hist))
                    rel = float(np.max(np.abs(hist[:k] - ref_h[:k]) / ref_h[:k])) if k else None
                    hcheck = {"entries": int(k), "max_rel_diff": rel, "ok": bool(k > 0 and rel <= 1e-6), "against": f"tests/golden/cg{ns}_history.json (1 GPU)"}
                per.sort()
                solve["cg_jacobi_poisson_strong"] = {"grid_total": f"{ns}^3", "grid_per_gpu": f"{ns}x{ns}x{nzs}", "scaling": "strong", "iters_per_solve": a.strong_iters,
                                                     "repeats": a.strong_repeats, "ms_per_iter_min": per[0], "ms_per_iter_median": per[len(per) // 2], "ms_per_iter_all": per,
                                                     "credited_gbs": BYTES_CG_ITER * float(ns) ** 3 / (per[0] * 1e-3) / 1e9,
                                                     "moved_gbs": 160 * float(ns) ** 3 / (per[0] * 1e-3) / 1e9,
                                                     "residual_history_check": hcheck,
                                                     "note": "fixed iteration count (rtol disabled), whole solve / iterations, incl. set-up of the loop and the initial residual; identical work at every GPU count"}
                summary[f"cg_jacobi_{ns}^3_history_matches_1gpu"] = None if hcheck is None else hcheck["ok"]
                summary[f"cg_jacobi_{ns}^3_strong_ms_per_iter_min"] = per[0]
                summary[f"cg_jacobi_{ns}^3_strong_ms_per_iter_median"] = per[len(per) // 2]
                del us_, bs_, As, ms_
            # (4) BiCGSTAB + SGS on convection-diffusion, 256 x 256 x 256 N (weak), every ordering that applies
            if n >= 256:
                mx = quiet(lambda: H.Mesh((258, 258, 256 * world + 2)))
                px = quiet(lambda: H.Pde("convdiff", V, 0.1))
                Ax, bx, ux = px.lhs(mx), px.vector(mx, "manufactured_rhs"), H.Vec(mx)
                if world == 1:
                    cases = [("sgs", {"-s3d_sgs_ordering": "lexicographic"}), ("sgs", {"-s3d_sgs_ordering": "blocks"}),
                             ("sgs", {"-s3d_sgs_ordering": "blocks", "-s3d_sgs_blocks": 4, "-s3d_sgs_blocks_y": 1}), ("sgs", {"-s3d_sgs_ordering": "redblack"}),
                             ("sgs", {"-s3d_sgs_ordering": "sweeps", "-s3d_pc_apply_sweeps": 2}), ("jacobi", {})]
                else:
                    cases = [("sgs", {"-s3d_sgs_ordering": "overlap"}), ("sgs", {"-s3d_sgs_ordering": "slablocal"}), ("sgs", {"-s3d_sgs_ordering": "lexicographic"}), ("sgs", {"-s3d_sgs_ordering": "redblack"})]
                for pc, extra in cases:
                    o = {"-s3d_ksp_type": "bcgs", "-s3d_pc_type": pc, "-ksp_rtol": 1e-6, "-ksp_max_it": 5000}
                    o.update(SGS); o.update(extra)
                    dt2, i2 = timed_solve(Ax, bx, ux, o)
                    key = ("bicgstab_" + pc + ("_" + extra["-s3d_sgs_ordering"] if extra else "") + (str(extra["-s3d_pc_apply_sweeps"]) if "-s3d_pc_apply_sweeps" in extra else "")
                           + (str(extra["-s3d_sgs_blocks"]) if "-s3d_sgs_blocks" in extra else "") + "_convdiff256perGPU")
                    solve[key] = {"grid": f"256x256x{256 * world}", "seconds": dt2, "iters": i2["iters"], "converged": i2["converged"],
                                  "ms_per_iter": dt2 * 1e3 / max(1, i2["iters"]), "precapply_seconds": i2["precapplywtime"],
                                  "precapply_ms_per_apply": i2["precapplywtime"] * 1e3 / max(1, 2 * i2["iters"])}
                k0_ = "bicgstab_sgs_" + ("lexicographic" if world == 1 else "overlap") + "_convdiff256perGPU"
                if world > 1:   # the iteration gate (north_star: a parallel SGS within +5 % of the reference ordering's iterations), same grid
                    ex_, ov_ = solve["bicgstab_sgs_lexicographic_convdiff256perGPU"]["iters"], solve[k0_]["iters"]
                    summary["sgs_gate_iters_overlap_vs_exact"] = [ov_, ex_, round(100.0 * (ov_ - ex_) / max(1, ex_), 2)]
                summary["bicgstab_sgs_256perGPU_ms_per_iter"] = solve[k0_]["ms_per_iter"]
                summary["bicgstab_sgs_256perGPU_iters"] = solve[k0_]["iters"]
                summary["sgs_apply_256perGPU_ms"] = solve[k0_]["precapply_ms_per_apply"]
                if world == 1:
                    # the blocks ordering (concurrent blocks of planes and rows, one plane / row of overlap): a parallel SGS under north_star's gate
                    kb_ = "bicgstab_sgs_blocks_convdiff256perGPU"
                    summary["bicgstab_sgs_blocks_256_ms_per_iter"] = solve[kb_]["ms_per_iter"]
                    summary["bicgstab_sgs_blocks_256_iters"] = solve[kb_]["iters"]
                    summary["bicgstab_sgs_blocks_256_solve_s"] = solve[kb_]["seconds"]
                    summary["sgs_blocks_apply_256_ms"] = solve[kb_]["precapply_ms_per_apply"]
                    summary["sgs_blocks_iters_vs_exact_pct"] = round(100.0 * (solve[kb_]["iters"] - solve[k0_]["iters"]) / max(1, solve[k0_]["iters"]), 2)
                del Ax, bx, ux, mx
        except Exception as e:  # the headline numbers above are measured already: report the failure instead of losing the line
            solve["error"] = f"{type(e).__name__}: {e}"[:400]

    cb = None
    if rank == 0 and world == 1 and not a.no_cpu:
        cb = cpu_spmv(n, 5)
    if rank == 0:
        e2e["solves"] = summary      # whole solves through createSolver()/apply(): the figures BASELINE.json's metric names beside the SpMV rate
        line = {"metric": "spmv_7pt_fp64_hbm_gbs", "value": value, "unit": "GB/s", "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": bench_config(n, world),
                "decomposition": (f"z-slabs x{world}, one-plane halo " + ("pushed into the neighbours' peer-memory windows (CUDA IPC over NVLink) and read by the SpMV itself" if a.halo_p2p else "by NCCL send/recv")) if world > 1 else "single GPU",
                "value_frac_of_measured_peak": value / (peak * world), "value_frac_of_nominal_8TBs": value / (8000.0 * world),
                "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "assembly_seconds": t_assemble, "solve": solve}
        if halo_check is not None:
            line["halo_check"] = halo_check
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
    ap.add_argument("--strong-iters", type=int, default=200, help="iterations of each timed strong-scaling solve")
    ap.add_argument("--strong-repeats", type=int, default=3)
    ap.add_argument("--fixed", type=int, default=512, help="grid size of the fixed-grid solves to rtol at every GPU count (BASELINE's 'solve s at 512^3'; 0 disables)")
    ap.add_argument("--halo-overlap", type=int, default=0, help="0: exchange halos first, then multiply (A/B switch)")
    ap.add_argument("--halo-p2p", type=int, default=1, help="1: halo planes through peer-memory windows (CUDA IPC), 0: NCCL send/recv (A/B switch)")
    ap.add_argument("--tune", action="append", default=[], help="library tuning key=value (A/B switches), may be repeated")
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
