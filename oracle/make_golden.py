#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY: generates tests/golden/ from the REAL reference.

Runs the reference's own sources (compiled unchanged into oracle/_ref/libs3dref.so by
oracle/Makefile) on the reference's own test configurations and stores
  tests/golden/solves.json   -- iteration counts, residual norms, printed "Rel res" histories,
                                ||b||, L2 errors, ||A*1|| known answers;
  tests/golden/fields_*.npz  -- full coefficient arrays and vectors for three small grids.
Needs /root/reference (so it only runs in the build container); the fixtures it writes are what
travels.  Re-run:  python oracle/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
V1 = (0.577350269189626, 0.7, 0.3)               # test/structsolvers/input/convdiff-large.control
V2 = (-0.577350269189626, 0.577350269189626, 0.577350269189626)  # test/structsolvers/input/convdiff.control

BASE = {"-ksp_rtol": 1e-6, "-ksp_max_it": 5000, "-s3d_pc_apply_sweeps": 1, "-s3d_pc_build_sweeps": 1,
        "-s3d_thread_chunk_size": 1024, "-s3d_pc_use_threaded_apply": "false",
        "-s3d_pc_use_threaded_build": "false"}


def golden_x(n):
    """The survey's deterministic non-eigenvector field: frac(phi*(idx+1)) - 0.5."""
    idx = np.arange(1, n + 1, dtype=np.float64)
    v = idx * 0.6180339887
    return v - np.floor(v) - 0.5


def fill_real(mesh, vec):
    nz, ny, nx = mesh.mshape
    vec.a[...] = 0
    vec.a[1:-1, 1:-1, 1:-1] = golden_x(nx * ny * nz).reshape(nz, ny, nx)


def main():
    R.build()
    R.set_num_threads(1)  # pinned oracle configuration
    os.makedirs(OUT, exist_ok=True)
    solves = []

    def solve_case(name, src, npdim, grid, pde, vel, mu, rhs, combos):
        m = R.Mesh(npdim, grid=grid)
        P = R.Pde(pde, vel, mu)
        A = P.lhs(m)
        b = P.vector(m, rhs)
        uex = P.vector(m, "manufactured_solution")
        for ksp, pc, extra in combos:
            opts = dict(BASE, **{"-s3d_ksp_type": ksp, "-s3d_pc_type": pc})
            opts.update(extra)
            x = R.Vec(m)
            if ksp in ("richardson", "gcr"):
                info = R.solve(A, b, x, opts)
                printed = {str(k): v for k, v in sorted(info["printed"].items())}
                hist = None
            else:
                pr = R.Prec(A, pc, int(opts["-s3d_pc_apply_sweeps"]), int(opts["-s3d_pc_build_sweeps"]))
                info, h = R.harness(A, pr, b, x, ksp, 1e-6, int(opts["-ksp_max_it"]))
                printed, hist = {}, [float(v) for v in h]
            solves.append({
                "case": name, "source": src, "npdim": list(npdim), "grid": grid, "pde": pde, "vel": list(vel),
                "mu": mu, "rhs": rhs, "ksp": ksp, "pc": pc,
                "apply_sweeps": int(opts["-s3d_pc_apply_sweeps"]), "build_sweeps": int(opts["-s3d_pc_build_sweeps"]),
                "origin": "reference createSolver" if hist is None else "harness loop over reference primitives",
                "converged": info["converged"], "iters": info["iters"], "resnorm": info["resnorm"],
                "bnorm": R.norm_vector_l2(b), "xnorm": R.norm_vector_l2(x),
                "l2_error_vs_manufactured": R.compute_error_L2(x, uex),
                "printed_rel_res": printed, "hist": hist})
            print(name, ksp, pc, info["iters"], info["converged"], flush=True)

    RJ, GJ = ("richardson", "jacobi", {}), ("gcr", "jacobi", {})
    solve_case("poisson16_cheb", "test/input/poisson.control", (16, 16, 16), "chebyshev", "poisson", (0, 0, 0), 0.0,
               "test_rhs", [RJ, GJ, ("gcr", "sgs", {}), ("richardson", "sgs", {}), ("richardson", "strilu", {"-s3d_pc_apply_sweeps": 2})])
    solve_case("poisson26", "test/structsolvers/input/poisson.control", (26, 26, 26), "uniform", "poisson", (0, 0, 0), 0.0,
               "manufactured_rhs", [RJ, ("cg", "jacobi", {}), ("cg", "sgs", {}), ("gcr", "sgs", {})])
    solve_case("convdiff48", "test/structsolvers/input/convdiff.control", (48, 48, 48), "uniform", "convdiff", V2, 0.1,
               "manufactured_rhs", [RJ, GJ, ("richardson", "sgs", {"-s3d_pc_apply_sweeps": 4}), ("bcgs", "sgs", {}), ("bcgs", "jacobi", {})])
    solve_case("convdiff64", "test/structsolvers/input/convdiff-large.control", (64, 64, 64), "uniform", "convdiff", V1, 0.1,
               "manufactured_rhs", [("gcr", "sgs", {}), ("richardson", "sgs", {}), ("gcr", "strilu", {}), ("bcgs", "sgs", {})])
    solve_case("poisson64_test_rhs", "SURVEY Appendix B (npdim 64 uniform, test_rhs)", (64, 64, 64), "uniform", "poisson",
               (0, 0, 0), 0.0, "test_rhs", [RJ, GJ, ("gcr", "sgs", {}), ("cg", "jacobi", {})])

    # non-eigenvector CG case: b = golden field, Poisson uniform
    for n in (34, 66):
        m = R.Mesh((n, n, n))
        P = R.Pde("poisson")
        A = P.lhs(m)
        b = R.Vec(m)
        fill_real(m, b)
        for pc in ("jacobi", "sgs"):
            x = R.Vec(m)
            pr = R.Prec(A, pc, 1, 1)
            info, h = R.harness(A, pr, b, x, "cg", 1e-6, 2000)
            solves.append({"case": f"poisson{n}_goldenrhs", "source": "SURVEY 8d synthetic", "npdim": [n] * 3, "grid": "uniform",
                           "pde": "poisson", "vel": [0, 0, 0], "mu": 0.0, "rhs": "golden", "ksp": "cg", "pc": pc,
                           "apply_sweeps": 1, "build_sweeps": 1, "origin": "harness loop over reference primitives",
                           "converged": info["converged"], "iters": info["iters"], "resnorm": info["resnorm"],
                           "bnorm": R.norm_vector_l2(b), "xnorm": R.norm_vector_l2(x), "l2_error_vs_manufactured": None,
                           "printed_rel_res": {}, "hist": [float(v) for v in h]})
            print(f"poisson{n}_goldenrhs cg", pc, info["iters"], flush=True)

    # ||A*1||_2 known answers + matvectest defects (test/matvectest.cpp:57-92)
    known = {"A_times_one_norm": {}, "matvectest": []}
    for n in (26, 34, 64, 66, 130):
        m = R.Mesh((n, n, n))
        A = R.Pde("poisson").lhs(m)
        x, y = R.Vec(m), R.Vec(m)
        R.vecset(1.0, x)
        R.apply(A, x, y)
        known["A_times_one_norm"][str(n)] = R.norm_vector_l2(y)
    for n in (18, 36, 72):
        m = R.Mesh((n, n, n))
        P = R.Pde("poisson")
        A, b, u = P.lhs(m), P.vector(m, "manufactured_rhs"), P.vector(m, "manufactured_solution")
        res, t = R.Vec(m), R.Vec(m)
        R.apply_res(A, b, u, res)
        R.apply(A, u, t)
        R.vecaxpy(-1.0, b, t)
        R.vecaxpy(1.0, res, t)
        known["matvectest"].append({"npdim": n, "apply_vs_apply_res": R.norm_vector_l2(t),
                                    "defect": R.norm_vector_l2(res) / np.sqrt(float(n) ** 3)})
    with open(os.path.join(OUT, "solves.json"), "w") as f:
        json.dump({"generator": "oracle/make_golden.py", "reference_build": "oracle/Makefile (unchanged sources, 1 thread, -DNOFORCESIMD, -ffp-contract=off)",
                   "solves": solves, "known": known}, f, indent=1)

    # full fields on three small grids
    write_fields((("fields_poisson16_cheb", (16, 16, 16), "chebyshev", "poisson", (0, 0, 0), 0.0),
                  ("fields_convdiff_20x14x18", (20, 14, 18), "uniform", "convdiff", V1, 0.1),
                  ("fields_convdiff_18x22x12_cheb", (18, 22, 12), "chebyshev", "convdiff", V2, 0.05)))
    write_circ()


def write_circ():
    """The third PDE of src/pde (convdiff_circular) on a ragged Chebyshev grid.  Its own file-name prefix: the
    fixture tests that loop over fields_*.npz assemble with the Poisson / convection-diffusion kernels."""
    write_fields((("circ_17x21x13_cheb", (17, 21, 13), "chebyshev", "convdiff_circular", (1.3, 0, 0), 0.07),))


def write_fields(specs):
    for name, npdim, grid, pde, vel, mu in specs:
        m = R.Mesh(npdim, grid=grid)
        P = R.Pde(pde, vel, mu)
        A = P.lhs(m)
        x = R.Vec(m)
        fill_real(m, x)
        b = P.vector(m, "manufactured_rhs")
        y, res, zj, zs, zi, ax = R.Vec(m), R.Vec(m), R.Vec(m), R.Vec(m), R.Vec(m), R.Vec(m)
        R.apply(A, x, y)
        R.apply_res(A, b, x, res)
        R.Prec(A, "jacobi").apply(b, zj)
        R.Prec(A, "sgs", 1, 1).apply(b, zs)
        R.Prec(A, "strilu", 1, 1).apply(b, zi)
        ax.a[...] = b.a
        R.vecaxpy(0.37, x, ax)
        trhs, msol = P.vector(m, "test_rhs"), P.vector(m, "manufactured_solution")   # keep the owners alive: .a is a view
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"),
                            npdim=np.array(npdim), grid=grid, pde=pde, vel=np.array(vel), mu=mu,
                            cx=m.coords(0), cy=m.coords(1), cz=m.coords(2), A=A.stacked(), x=x.a,
                            test_rhs=trhs.a.copy(), manufactured_rhs=b.a.copy(),
                            manufactured_solution=msol.a.copy(),
                            Ax=y.a, b_minus_Ax=res.a, jacobi_b=zj.a, sgs_b=zs.a, strilu_b=zi.a, b_plus_037x=ax.a,
                            dot_x_b=R.inner_vector_l2(x, b), norm_x=R.norm_vector_l2(x), normL2_x=R.norm_L2(x),
                            errL2_x_b=R.compute_error_L2(x, b), h=m.h())
        print("fields", name, flush=True)


if __name__ == "__main__":
    if "--circ-only" in sys.argv:
        write_circ()
    else:
        main()
