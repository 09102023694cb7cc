"""GPU tests of the C++ host classes (struct3d_b200/host: SVec, SMat, createSolver, PDEBase ...)
driven through include/s3d_host.h.  They replay the reference's own call sequences
(createSolver -> updateOperator -> apply; test/matvectest.cpp; nativevectorassemblytest.cpp's ghost and
closed-form checks; gridconv.cpp) and compare with
  - tests/golden/solves.json: iteration counts, residual norms and residual histories produced by
    the REAL reference (Richardson, GCR) and by harness loops over reference primitives (CG, BiCGSTAB);
  - the CPU oracle on the same inputs.
Bars: Richardson/GCR/CG iteration counts EQUAL the reference's, residual histories agree to 1e-6
relative (they differ only through the summation order of the dot products); BiCGSTAB within 5%
iterations (it amplifies rounding); red-black SGS within +5% iterations of lexicographic SGS where
north_star demands it, reported where it does not hold."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import cport
from tests.util import V1, V2, ghosts_untouched, golden_vec

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOLVES = json.load(open(os.path.join(ROOT, "tests", "golden", "solves.json")))
LIBDIR = os.path.join(ROOT, "struct3d_b200", "lib")
INPUT = os.path.join(ROOT, "tests", "input")


@pytest.fixture(scope="module")
def H():
    from struct3d_b200 import hostapi
    hostapi.init(0)
    return hostapi


def _opts(rec, **extra):
    o = {"-s3d_ksp_type": rec["ksp"], "-s3d_pc_type": rec["pc"], "-ksp_rtol": 1e-6, "-ksp_max_it": 5000,
         "-s3d_pc_apply_sweeps": rec["apply_sweeps"], "-s3d_pc_build_sweeps": rec["build_sweeps"],
         "-s3d_thread_chunk_size": 1024, "-s3d_pc_use_threaded_apply": "false"}
    o.update(extra)
    return o


def _setup(H, rec):
    m = H.Mesh(rec["npdim"], grid=rec["grid"])
    P = H.Pde(rec["pde"], rec["vel"], rec["mu"])
    A = P.lhs(m)
    if rec["rhs"] == "golden":
        b = H.Vec(m)
        b.set(golden_vec(cport.Mesh(rec["npdim"], grid=rec["grid"])))
    else:
        b = P.vector(m, rec["rhs"])
    return m, P, A, b


RUNS = [r for r in SOLVES["solves"] if r["pc"] in ("jacobi", "sgs", "strilu", "none")]


@pytest.mark.parametrize("rec", RUNS, ids=lambda r: f"{r['case']}-{r['ksp']}-{r['pc']}")
def test_solver_matches_reference_run(H, rec):
    m, P, A, b = _setup(H, rec)
    x = H.Vec(m)
    info = H.solve(A, b, x, _opts(rec), capture=True)
    assert info["converged"] == rec["converged"]
    assert abs(H.norm_vector_l2(b) - rec["bnorm"]) <= 1e-12 * rec["bnorm"]
    if rec["ksp"] in ("bcgs", "bicgstab"):
        assert abs(info["iters"] - rec["iters"]) <= max(2, rec["iters"] // 20)
        assert info["resnorm"] <= 1e-6 * rec["bnorm"]
        return
    assert info["iters"] == rec["iters"], f"iterations {info['iters']} vs reference {rec['iters']}"
    if rec["iters"] > 1:
        assert abs(info["resnorm"] - rec["resnorm"]) <= 1e-6 * rec["resnorm"]
        hist = info["hist"]
        assert len(hist) == rec["iters"]
        # the reference's printed "Step n: Rel res = %g" lines (6 significant digits)
        for step, val in rec["printed_rel_res"].items():
            assert abs(hist[int(step)] - val) <= 2e-5 * val
        # our driver prints the same lines
        for step, val in info["printed"].items():
            if str(step) in rec["printed_rel_res"]:
                assert abs(val - rec["printed_rel_res"][str(step)]) <= 2e-5 * val
        if rec["hist"]:
            h = np.array(rec["hist"])
            assert np.max(np.abs(hist - h) / h) <= 1e-6
        assert abs(H.norm_vector_l2(x) - rec["xnorm"]) <= 1e-7 * rec["xnorm"]
        if rec["l2_error_vs_manufactured"] is not None and rec["rhs"] == "manufactured_rhs":
            uex = P.vector(m, "manufactured_solution")
            assert abs(H.compute_error_L2(x, uex) - rec["l2_error_vs_manufactured"]) <= 1e-6 * rec["l2_error_vs_manufactured"]
    assert ghosts_untouched(x.a)


GRAPH_RUNS = [r for r in RUNS if (r["ksp"], r["pc"]) in (("cg", "jacobi"), ("cg", "sgs"), ("bcgs", "sgs"), ("bcgs", "jacobi"),
                                                         ("richardson", "sgs"), ("richardson", "strilu"), ("gcr", "sgs"))][:14]


@pytest.mark.parametrize("rec", GRAPH_RUNS, ids=lambda r: f"{r['case']}-{r['ksp']}-{r['pc']}")
@pytest.mark.parametrize("ordering", ["lexicographic", "redblack", "sweeps"])
def test_cuda_graph_replay_is_identical(H, rec, ordering):
    """Iteration batches replayed from a captured CUDA graph must give bit-identical iterates, histories and counts to
    the eagerly enqueued loop (same kernels, same order; only the launch mechanism differs)."""
    if rec["pc"] not in ("sgs", "strilu") and ordering != "lexicographic":
        pytest.skip("ordering only matters for the SGS-like preconditioners")
    out = {}
    for mode in ("off", "on"):
        m, P, A, b = _setup(H, rec)
        x = H.Vec(m)
        extra = {"-s3d_cuda_graph": mode, "-s3d_sgs_ordering": ordering}
        if ordering == "sweeps":
            extra["-s3d_pc_apply_sweeps"] = 3
        info = H.solve(A, b, x, _opts(rec, **extra), capture=True)
        out[mode] = (info["iters"], info["converged"], np.array(info["hist"]), x.a.copy())
    H.set_options({})
    assert out["on"][0] == out["off"][0] and out["on"][1] == out["off"][1]
    assert np.array_equal(out["on"][2], out["off"][2]), "residual histories differ between graph replay and eager launches"
    assert np.array_equal(out["on"][3], out["off"][3]), "solutions differ between graph replay and eager launches"
    assert out["on"][0] > 0


def test_asynchronous_sgs_sweeps_iteration_bounds(H):
    """the reference's threaded, asynchronous SGS application (s3d_sgspreconditioners.cpp:30-112) is not
    reproducible from run to run, so it is tested on bounds, as the reference itself can only be: on the convdiff-large case the
    reference needs 75 GCR iterations with the sequential sweep and 84 / 81 with 8 asynchronous threads and 1 / 3 sweeps.
    Here thousands of warps start at once where the reference has 8 threads, so a single asynchronous sweep is closer to a Jacobi
    sweep (measured: 116 iterations), three sweeps give 78 and eight the exact sweep's 75.  Bounds: the solve converges, never needs
    noticeably fewer iterations than the exact sweep nor more than twice as many (1.2 times from three sweeps on), and more sweeps
    per application never cost more iterations."""
    rec = next(r for r in SOLVES["solves"] if r["case"] == "convdiff64" and r["ksp"] == "gcr" and r["pc"] == "sgs")
    counts = {}
    for sweeps in (1, 3, 8):
        m, P, A, b = _setup(H, rec)
        x = H.Vec(m)
        got = H.solve(A, b, x, _opts(rec, **{"-s3d_sgs_ordering": "async", "-s3d_pc_apply_sweeps": sweeps, "-s3d_thread_chunk_size": 4}))
        assert got["converged"], (sweeps, got)
        counts[sweeps] = got["iters"]
    H.set_options({})
    exact = rec["iters"]
    print(f"[async-sgs] gcr + asynchronous SGS, convdiff 62^3: {counts} iterations (exact sweep: {exact})")
    assert all(exact - 3 <= n <= 2 * exact for n in counts.values()), (counts, exact)
    assert counts[3] <= 1.2 * exact and counts[8] <= exact + 5, (counts, exact)
    assert counts[8] <= counts[3] + 3 <= counts[1] + 6, (counts, exact)


def test_asynchronous_strilu_build_iteration_bounds(H):
    """The reference's threaded StrILU build (s3d_ilu.cpp:28-32): asynchronous sweeps from d = diag(A).  GCR + StrILU on the convdiff-large
    case needs 61 iterations with the exact factorisation (golden); with asynchronously built factors and asynchronous application it
    converges within 1.5 times that from three sweeps each on, and reproduces the 61 (give or take rounding) with many sweeps."""
    rec = next(r for r in SOLVES["solves"] if r["case"] == "convdiff64" and r["ksp"] == "gcr" and r["pc"] == "strilu")
    counts = {}
    for sweeps in (3, 40):
        m, P, A, b = _setup(H, rec)
        x = H.Vec(m)
        got = H.solve(A, b, x, _opts(rec, **{"-s3d_sgs_ordering": "async", "-s3d_pc_apply_sweeps": sweeps, "-s3d_pc_build_sweeps": sweeps,
                                             "-s3d_pc_use_threaded_build": "true", "-s3d_thread_chunk_size": 4}))
        assert got["converged"], (sweeps, got)
        counts[sweeps] = got["iters"]
    H.set_options({})
    print(f"[async-strilu] gcr + asynchronously built and applied StrILU, convdiff 62^3: {counts} iterations (exact: {rec['iters']})")
    assert counts[3] <= 1.5 * rec["iters"] and abs(counts[40] - rec["iters"]) <= 3, (counts, rec["iters"])


def _gate_rows(H, case, ordering):
    rows = []
    for rec in [r for r in SOLVES["solves"] if r["case"] == case and r["pc"] == "sgs"]:
        m, P, A, b = _setup(H, rec)
        x = H.Vec(m)
        extra = {"-s3d_sgs_ordering": ordering} if ordering else {}
        got = H.solve(A, b, x, _opts(rec, **extra))
        rows.append((rec["ksp"], rec["iters"], got["iters"], got["converged"], got["resnorm"] / rec["bnorm"]))
    H.set_options({})
    return rows


@pytest.mark.parametrize("case", ["convdiff48", "convdiff64", "poisson26", "poisson64_test_rhs"])
def test_default_sgs_iteration_gate(H, case):
    """north_star: the GPU SGS must reach the same final residual within 5% more iterations than the reference's (lexicographic)
    SGS.  ASSERTED for the ordering the library uses by default on one GPU (the exact lexicographic sweep executed as a dataflow:
    same operator, so the counts are the reference's own) for every solver the reference goldens hold with SGS."""
    for ksp, ref_it, it, conv, rel in _gate_rows(H, case, None):
        print(f"[sgs-gate] {case:20s} {ksp:10s} reference {ref_it:4d}  default ordering {it:4d}  ({100.0 * (it - ref_it) / ref_it:+.1f}%)  rel.res {rel:.3e}")
        assert conv and rel <= 1e-6
        assert it <= ref_it * 1.05 + 0.5, f"{ksp}: {it} iterations against the reference's {ref_it}: more than +5%"


@pytest.mark.parametrize("case", ["convdiff48", "convdiff64", "poisson64_test_rhs"])
def test_blocks_sgs_iteration_gate(H, case):
    """The same gate for the z-blocks ordering (-s3d_sgs_ordering blocks: concurrent sweeps over blocks of planes with one plane of
    overlap, the one-device counterpart of the overlapping slabs): ASSERTED with two blocks on every solver the reference goldens
    hold with SGS; the counts with four blocks of planes and with 2 x 2 blocks of planes and rows are printed beside them."""
    for nb, nby in ((2, 1), (4, 1), (2, 2)):
        for rec in [r for r in SOLVES["solves"] if r["case"] == case and r["pc"] == "sgs"]:
            m, P, A, b = _setup(H, rec)
            x = H.Vec(m)
            got = H.solve(A, b, x, _opts(rec, **{"-s3d_sgs_ordering": "blocks", "-s3d_sgs_blocks": nb, "-s3d_sgs_blocks_y": nby}))
            ref_it, it, rel = rec["iters"], got["iters"], got["resnorm"] / rec["bnorm"]
            print(f"[sgs-gate] {case:20s} {rec['ksp']:10s} reference {ref_it:4d}  {nb} x {nby} blocks {it:4d}  ({100.0 * (it - ref_it) / ref_it:+.1f}%)  rel.res {rel:.3e}")
            assert got["converged"] and rel <= 1e-6
            if nb == 2 and nby == 1:
                assert it <= ref_it * 1.05 + 0.5, f"{rec['ksp']}: {it} iterations against the reference's {ref_it}: more than +5%"
        H.set_options({})


@pytest.mark.parametrize("case", ["convdiff48", "convdiff64"])
@pytest.mark.xfail(strict=True, reason="measured: the red-black (two-colour) ordering needs +30..40% iterations on upwinded convection-diffusion "
                                       "(265 vs 193 at 256^3, driver BENCH r1); it stays selectable but is not the default")
def test_redblack_sgs_iteration_gate(H, case):
    """The same gate for the two-colour ordering north_star suggests: it FAILS on the convection-diffusion cases (strict xfail, so
    that an unexpected pass is noticed); the numbers are printed."""
    worst = 0.0
    for ksp, ref_it, it, conv, rel in _gate_rows(H, case, "redblack"):
        print(f"[sgs-gate] {case:20s} {ksp:10s} reference {ref_it:4d}  red-black {it:4d}  ({100.0 * (it - ref_it) / ref_it:+.1f}%)  rel.res {rel:.3e}")
        assert conv and rel <= 1e-6
        if ksp != "richardson":
            worst = max(worst, (it - ref_it) / ref_it)
    assert worst <= 0.05


@pytest.mark.parametrize("case", ["poisson26", "poisson64_test_rhs"])
def test_redblack_sgs_converges_on_poisson(H, case):
    for ksp, ref_it, it, conv, rel in _gate_rows(H, case, "redblack"):
        print(f"[sgs-gate] {case:20s} {ksp:10s} reference {ref_it:4d}  red-black {it:4d}  ({100.0 * (it - ref_it) / ref_it:+.1f}%)  rel.res {rel:.3e}")
        assert conv and rel <= 1e-6


def test_matvec_suite_like_reference(H):
    """test/matvectest.cpp: apply vs apply_res consistency and second-order defect, 18^3 -> 36^3 -> 72^3,
    with the reference's own numbers from the goldens."""
    P = H.Pde("poisson")
    defects = []
    for rec in SOLVES["known"]["matvectest"]:
        n = rec["npdim"]
        m = H.Mesh((n, n, n))
        A, b, u = P.lhs(m), P.vector(m, "manufactured_rhs"), P.vector(m, "manufactured_solution")
        res, t = H.Vec(m), H.Vec(m)
        H.apply_res(A, b, u, res)
        H.apply(A, u, t)
        H.vecaxpy(-1.0, b, t)
        H.vecaxpy(1.0, res, t)
        assert H.norm_vector_l2(t) < 1e-14
        d = H.norm_vector_l2(res) / np.sqrt(float(n) ** 3)
        assert abs(d - rec["defect"]) <= 1e-12 * rec["defect"]
        defects.append(d)
    for i in (1, 2):
        slope = (np.log10(defects[i]) - np.log10(defects[i - 1])) / np.log10(0.5)
        assert 1.9 <= slope <= 2.1


def test_assembly_like_reference_test(H):
    """test/nativevectorassemblytest.cpp: every ghost entry is zero; uniform Poisson coefficients are
    6/dx^2 and -1/dx^2 within 100 eps; plus: the separable GPU path equals the host-sampled path bitwise."""
    for npdim, grid in (((12, 12, 12), "uniform"), ((13, 9, 11), "chebyshev")):
        m = H.Mesh(npdim, grid=grid)
        cm = cport.Mesh(npdim, grid=grid)
        for pde, vel, mu in (("poisson", (0, 0, 0), 0.0), ("convdiff", V1, 0.1), ("convdiff_circular", (1.3, 0, 0), 0.07)):
            P = H.Pde(pde, vel, mu)
            for which in ("test_rhs", "manufactured_solution", "manufactured_rhs"):
                f = P.vector(m, which).a
                assert ghosts_untouched(f)
                assert np.array_equal(f, P.vector(m, which, sampled=True).a)
                assert np.array_equal(f, cport.grid_function(cm, pde, which, vel, mu))
            assert np.array_equal(P.lhs(m).stacked(), cport.lhs(cm, pde, vel, mu))
        if grid == "uniform":
            A = H.Pde("poisson").lhs(m).stacked()
            dx = cm.coords[0][2] - cm.coords[0][1]
            eps = np.finfo(float).eps
            assert np.all(np.abs(A[3] - 6.0 / dx**2) <= 100 * eps * (6.0 / dx**2))
            for s in (0, 1, 2, 4, 5, 6):
                assert np.all(np.abs(A[s] + 1.0 / dx**2) <= 100 * eps / dx**2)


def test_dirichlet_boundary_values_enter_the_rhs(H):
    bv = (0.5, -1.0, 2.0, 0.25, -0.75, 1.5)
    npdim = (9, 8, 10)
    m, cm = H.Mesh(npdim, grid="chebyshev"), cport.Mesh(npdim, grid="chebyshev")
    P = H.Pde("poisson", bvals=bv)
    ref = cport.grid_function(cm, "poisson", "manufactured_rhs", bvals=bv)
    assert np.array_equal(P.vector(m, "manufactured_rhs").a, ref)
    assert np.array_equal(P.vector(m, "manufactured_rhs", sampled=True).a, ref)
    with pytest.raises(H.HostError, match="Invalid BC type at j\\+"):
        H.Pde("poisson", dirichlet=(1, 1, 1, 0, 1, 1)).vector(m, "test_rhs")


def test_host_mirror_is_coherent(H):
    """`vals[...]` of the reference is public; ours is a coherent view of device memory."""
    m = H.Mesh((8, 7, 6))
    cm = cport.Mesh((8, 7, 6))
    x, y = H.Vec(m), H.Vec(m)
    H.vecset(2.0, x)
    idx = (3 * 7 + 2) * 8 + 4          # localFlattenedIndexAll(k=3, j=2, i=4)
    assert x.peek(idx) == 2.0 and x.peek(0) == 0.0
    x.poke(idx, -5.0)                  # host write ...
    H.vecaxpy(1.0, x, y)               # ... must be visible to the next kernel
    ya = y.a
    assert ya[3, 2, 4] == -5.0 and ya[3, 2, 3] == 2.0
    A = H.Pde("poisson").lhs(m)
    As = A.stacked()
    As[3] *= 2.0
    A.set(As)
    H.apply(A, x, y)
    assert np.array_equal(y.a, cport.apply(cm, As, x.a))
    z = x.clone()
    H.vecset(0.0, x)
    assert z.peek(idx) == -5.0 and x.peek(idx) == 0.0


def test_blas1_free_functions(H):
    npdim = (11, 13, 9)
    m, cm = H.Mesh(npdim, grid="chebyshev"), cport.Mesh(npdim, grid="chebyshev")
    rng = np.random.default_rng(3)
    def rv():
        v = cm.zeros(); v[1:-1, 1:-1, 1:-1] = rng.uniform(-1, 1, cm.mshape[1:]); return v
    xs = [rv() for _ in range(4)]
    hx = []
    for v in xs:
        h = H.Vec(m); h.set(v); hx.append(h)
    y = H.Vec(m); y.set(xs[3])
    ref = xs[3].copy()
    a = [0.5, -0.25, 3.0]
    cport.vec_multi_axpy(cm, a, xs[:3], ref)
    H.vec_multi_axpy(a, hx[:3], y)
    assert np.array_equal(y.a, ref)
    assert abs(H.inner_vector_l2(hx[0], y) - cport.inner_vector_l2(cm, xs[0], ref)) <= 1e-13 * np.linalg.norm(xs[0]) * np.linalg.norm(ref)
    assert abs(H.norm_vector_l2(y) - cport.norm_vector_l2(cm, ref)) <= 1e-13 * np.linalg.norm(ref)
    assert abs(H.norm_L2(y) - cport.norm_L2(cm, ref)) <= 1e-13 * cport.norm_L2(cm, ref)
    assert abs(H.compute_error_L2(hx[0], y) - cport.compute_error_L2(cm, xs[0], ref)) <= 1e-13 * cport.compute_error_L2(cm, xs[0], ref)
    H.vecassign(hx[1], y)
    assert np.array_equal(y.a, xs[1])


def test_error_behaviour_matches_reference(H):
    m1, m2 = H.Mesh((8, 8, 8)), H.Mesh((8, 8, 8))
    A = H.Pde("poisson").lhs(m1)
    x1, y2, b1 = H.Vec(m1), H.Vec(m2), H.Vec(m1)
    with pytest.raises(H.HostError, match="same mesh"):
        H.apply(A, x1, y2)                       # matvec.cpp:61
    with pytest.raises(H.HostError, match="same mesh"):
        H.apply_res(A, b1, x1, y2)               # matvec.cpp:105
    with pytest.raises(H.HostError, match="same mesh"):
        H.vecaxpy(1.0, x1, y2)                   # matvec.cpp:150
    with pytest.raises(H.HostError, match="same mesh"):
        H.inner_vector_l2(x1, y2)                # matvec.cpp:259
    H.set_options({"-s3d_ksp_type": "nonsense", "-s3d_pc_type": "jacobi", "-ksp_rtol": 1e-6, "-ksp_max_it": 10})
    with pytest.raises(H.HostError, match="Need a valid solver option"):
        H.Solver(A)                              # solverfactory.cpp:96
    H.set_options({"-s3d_ksp_type": "cg"})
    with pytest.raises(H.HostError, match="Could not find -ksp_rtol"):
        H.Solver(A)                              # common_utils.cpp:73
    with pytest.raises(H.HostError, match="not actually uniform|at least one interior"):
        H.Mesh((2, 8, 8))
    with pytest.raises(H.HostError, match="PDE type not recognized"):
        H.Pde("stokes")
    # a solve that cannot converge reports it through SolveInfo, not an exception
    H.set_options({"-s3d_ksp_type": "richardson", "-s3d_pc_type": "jacobi", "-ksp_rtol": 1e-12, "-ksp_max_it": 3})
    b = H.Pde("poisson").vector(m1, "manufactured_rhs")
    s = H.Solver(A); s.update()
    info = s.apply(b, H.Vec(m1))
    assert not info["converged"] and info["iters"] == 3
    H.set_options({"-s3d_ksp_type": "richardson", "-s3d_pc_type": "jacobi", "-ksp_rtol": 1e-6, "-ksp_max_it": 0})
    s = H.Solver(A); s.update()
    info = s.apply(b, H.Vec(m1))
    assert not info["converged"] and info["iters"] == 0


def test_zero_rhs_and_nosolver(H):
    m = H.Mesh((10, 10, 10))
    A = H.Pde("poisson").lhs(m)
    b = H.Pde("poisson").vector(m, "test_rhs")
    x = H.Vec(m)
    # no preconditioner option at all: createSolver falls back to the identity, like the reference
    info = H.solve(A, b, x, {"-s3d_ksp_type": "cg", "-ksp_rtol": 1e-10, "-ksp_max_it": 500})
    assert info["converged"]
    cm = cport.Mesh((10, 10, 10))
    xo, oi, _ = cport.solve(cm, cport.lhs(cm), cport.grid_function(cm, "poisson", "test_rhs"), "cg", "none", 1e-10, 500)
    assert info["iters"] == oi["iters"] and np.abs(x.a - xo).max() <= 1e-12 * np.abs(xo).max()


# option sets given on the command line (the drivers take "-key value" like PetscInitialize does): Richardson + StrILU as in the
# reference's discretization-verification runs (sequential build and apply, 1 build sweep, 2 apply sweeps), and Richardson + SGS
STRILU_SEQ = ["-s3d_ksp_type", "richardson", "-ksp_rtol", "1e-6", "-ksp_max_it", "1000", "-s3d_pc_type", "strilu", "-s3d_pc_use_threaded_build", "false", "-s3d_pc_use_threaded_apply", "false", "-s3d_pc_build_sweeps", "1", "-s3d_pc_apply_sweeps", "2", "-s3d_thread_chunk_size", "1024"]
SGS_SEQ = ["-s3d_ksp_type", "richardson", "-ksp_rtol", "1e-6", "-ksp_max_it", "1000", "-s3d_pc_type", "sgs", "-s3d_pc_use_threaded_apply", "false", "-s3d_pc_build_sweeps", "1", "-s3d_pc_apply_sweeps", "1", "-s3d_thread_chunk_size", "1024"]


@pytest.mark.parametrize("args", [
    ["testmatvec", os.path.join(INPUT, "poisson_conv.control")],
    ["gridconv", os.path.join(INPUT, "poisson_conv.control"), "-options_file", os.path.join(INPUT, "cg_jacobi.perc")],
    ["gridconv", os.path.join(INPUT, "convdiff_conv.control"), "-options_file", os.path.join(INPUT, "bcgs_sgs.perc"), "-ksp_rtol", "1e-8"],
    ["gridconv", os.path.join(INPUT, "poisson_conv.control")] + STRILU_SEQ,
    ["gridconv", os.path.join(INPUT, "convdiff_conv.control")] + STRILU_SEQ,
    ["gridconv", os.path.join(INPUT, "cdcirc_conv.control")] + STRILU_SEQ,
    ["runcase", os.path.join(INPUT, "convdiff64.control")] + SGS_SEQ,
    ["runcase", os.path.join(INPUT, "poisson16_cheb.control"), "-options_file", os.path.join(INPUT, "gcr_jacobi.perc")],
    ["scaling", os.path.join(INPUT, "convdiff48.control"), "-options_file", os.path.join(INPUT, "scaling_gpus.perc")],
])
def test_cpp_drivers(args):
    """The C++ drivers (counterparts of the reference's testmatvec / gridconv / runcase) run as processes."""
    exe = os.path.join(LIBDIR, args[0])
    out = subprocess.run([exe] + args[1:], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    text = out.stdout.decode()
    assert out.returncode == 0, text[-3000:]
    if args[0] == "runcase":
        want = {"convdiff64.control": 464, "poisson16_cheb.control": 17}[os.path.basename(args[1])]
        # NB: runcase uses test_rhs like the reference's driver, the goldens for these two cases use other RHS;
        # so only check convergence and the report format here
        assert "Avg solver iters:" in text and "Time taken by native linear solver" in text
    elif args[0] == "scaling":
        # the reference's scaling-output.txt layout (scaling_openmp.cpp:185-211): header of the reference run, column titles,
        # the reference line, then one line per (build sweeps, apply sweeps) pair for this GPU count
        rep = open("/tmp/s3d_scaling-output.txt").read().splitlines()
        assert rep[0].startswith("Reference run: Threads 1, build sweeps 1, apply sweeps 4")
        assert rep[4].split() == ["b.swps.", "a.swps.", "threads", "iters", "dev.iters", "b.spdp", "a.spdp", "spdp", "dev.b.time", "dev.a.time", "dev.resnorm"]
        lines = [ln.split() for ln in rep[6:] if ln.strip()]
        assert [ln[:3] for ln in lines] == [["1", "4", "1"], ["1", "3", "1"], ["1", "6", "1"]]
        assert all(len(ln) == 11 and int(ln[3]) > 0 for ln in lines)
        assert int(lines[2][3]) <= int(lines[0][3]) <= int(lines[1][3])      # more sweeps per factor: never more iterations
    else:
        assert "PASSED" in text


@pytest.mark.parametrize("npdim,grid", [((16, 16, 16), "chebyshev"), ((26, 26, 26), "uniform"), ((34, 30, 22), "uniform"), ((3, 3, 3), "uniform"),
                                        ((50, 40, 34), "uniform"), ((66, 66, 66), "uniform")])   # the last two: the cooperative-grid team
@pytest.mark.parametrize("ksp,pc", [("richardson", "jacobi"), ("cg", "jacobi"), ("cg", "none"), ("richardson", "none"), ("gcr", "jacobi"),
                                    ("gcr", "none"), ("gcr3", "jacobi")])
def test_single_launch_solver_equals_the_host_driven_loop(H, npdim, grid, ksp, pc):
    """Small grids: the whole solve in one kernel launch (s3d_small_solve, -s3d_small_solve on) against the host-driven loop of the
    same solver (off): same iteration count, same residual history and solution up to the summation order of the dot products."""
    if ksp == "cg" and grid != "uniform":
        pytest.skip("the Poisson matrix is not symmetric on a Chebyshev grid (SURVEY 0.6)")
    if ksp == "richardson" and pc == "none":
        maxit = 50        # plain Richardson diverges: compare the first 50 iterates
    else:
        maxit = 3000
    out = {}
    for mode in ("off", "on"):
        m = H.Mesh(npdim, grid=grid)
        P = H.Pde("poisson")
        A, b = P.lhs(m), H.Vec(m)
        n = [v - 2 for v in npdim]
        idx = np.arange(1, n[0] * n[1] * n[2] + 1, dtype=np.float64) * 0.6180339887
        bh = np.zeros(npdim[::-1]); bh[1:-1, 1:-1, 1:-1] = (idx - np.floor(idx) - 0.5).reshape(n[::-1])   # not an eigenvector of A
        b.set(bh)
        x = H.Vec(m)
        o = {"-s3d_ksp_type": ksp.rstrip("3"), "-ksp_rtol": 1e-8, "-ksp_max_it": maxit, "-s3d_small_solve": mode}
        if ksp == "gcr3":
            o["-ksp_restart"] = 3        # several restart cycles
        if pc != "none":
            o["-s3d_pc_type"] = pc
        info = H.solve(A, b, x, o, capture=True)
        out[mode] = (info, x.a.copy())
    H.set_options({})
    a, c = out["off"][0], out["on"][0]
    assert a["converged"] == c["converged"]
    ha, hc = np.asarray(a["hist"]), np.asarray(c["hist"])
    if ksp.startswith("gcr"):
        assert abs(a["iters"] - c["iters"]) <= max(1, a["iters"] // 50), (a["iters"], c["iters"])
        n = min(len(ha), len(hc)); ha, hc = ha[:n], hc[:n]
    else:
        assert a["iters"] == c["iters"], (a["iters"], c["iters"])
    assert len(ha) == len(hc)
    if len(ha) and np.all(np.isfinite(ha)):
        if ksp.startswith("gcr"):
            # the Gram-Schmidt coefficients amplify the different summation order of the dot products as the cycle goes on (measured:
            # 1e-16 in the first iterations, 1e-2 after 200 on the unpreconditioned Chebyshev case): the first ten entries tightly,
            # the rest loosely, the solution to what two solves converged to rtol 1e-8 can differ by
            assert np.all(np.abs(ha - hc)[:10] <= 1e-9 * ha[:10] + 1e-300)
            assert np.all(np.abs(ha - hc) <= 0.1 * ha + 1e-300)
            assert np.abs(out["on"][1] - out["off"][1]).max() <= 1e-5 * max(1e-300, np.abs(out["off"][1]).max())
        else:
            assert np.all(np.abs(ha - hc) <= 1e-9 * ha + 1e-300)
            assert np.abs(out["on"][1] - out["off"][1]).max() <= 1e-11 * max(1e-300, np.abs(out["off"][1]).max())
