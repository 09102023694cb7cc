"""GPU parity tests proper: every kernel of libs3d_cuda.so, called through the C ABI, against the
CPU oracle (oracle/s3d_oracle.c, pinned to the reference) on the same seeded inputs, plus the
golden fixtures generated from the reference itself (tests/golden/).

Tolerances: SpMV, Jacobi, SGS(lexicographic), assembly and vector updates are computed with the
reference's rounding sequence (no FMA contraction) and must be BIT-EXACT; reductions sum in a
different (deterministic) order and must agree to 1e-13 relative; north_star's bound is 1e-12."""
import glob
import os

import numpy as np
import pytest

from oracle import cport
from struct3d_b200 import capi
from tests.util import V1, V2, ghosts_untouched, golden_vec, problem, rand_vec, real, relerr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")

# reference MatVec sizes (test/matvectest.cpp: 18^3 -> 36^3 -> 72^3), ragged/odd shapes, tiny and thin grids
SHAPES = [(18, 18, 18), (36, 36, 36), (72, 72, 72), (20, 14, 18), (19, 23, 11), (3, 3, 3), (4, 9, 3), (131, 7, 5), (67, 66, 35)]
VARIANTS = [capi.SPMV_PAIR, capi.SPMV_MARCH, capi.SPMV_NAIVE, capi.SPMV_TMA]


def grid_of(m):
    nx, ny, nz = m.n
    return capi.make_grid(nx, ny, nz)


@pytest.mark.parametrize("npdim", SHAPES)
@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("pde", ["poisson", "convdiff"])
def test_spmv_apply_bit_exact(ctx, npdim, variant, pde):
    m, A = problem(npdim, "chebyshev" if npdim[0] % 2 else "uniform", pde, V1, 0.1)
    g = grid_of(m)
    x = rand_vec(m, 1)
    dA, dx, dy = ctx.mat_from(g, A), ctx.vec_from(g, x), ctx.vec_alloc(g)
    ctx.spmv(g, dA, dx, dy, variant=variant)
    y = ctx.vec_download(g, dy)
    ref = cport.apply(m, A, x)
    assert np.array_equal(y, ref), f"max diff {np.abs(y - ref).max()}"
    assert ghosts_untouched(y)
    for p in (dA, dx, dy):
        ctx.free(p)


@pytest.mark.parametrize("npdim", SHAPES[:6])
@pytest.mark.parametrize("variant", VARIANTS)
def test_spmv_apply_res_and_fused_dots(ctx, npdim, variant):
    m, A = problem(npdim, "uniform", "convdiff", V2, 0.05)
    g = grid_of(m)
    x, b, w = rand_vec(m, 2), rand_vec(m, 3), rand_vec(m, 4)
    dA, dx, db, dw, dy = ctx.mat_from(g, A), ctx.vec_from(g, x), ctx.vec_from(g, b), ctx.vec_from(g, w), ctx.vec_alloc(g)
    red = ctx.scalars_alloc(2)
    ctx.spmv(g, dA, dx, dy, b=db, w=dw, want_yy=True, red=red, variant=variant)
    y = ctx.vec_download(g, dy)
    ref = cport.apply_res(m, A, b, x)
    assert np.array_equal(y, ref)
    wy, yy = ctx.scalars(red, 2)
    assert abs(wy - cport.inner_vector_l2(m, w, ref)) <= 1e-13 * np.linalg.norm(w) * np.linalg.norm(ref)
    assert abs(yy - cport.inner_vector_l2(m, ref, ref)) <= 1e-13 * yy
    # CG's (p, Ap): w aliases x
    ctx.spmv(g, dA, dx, dy, w=dx, red=red, variant=variant)
    ax = cport.apply(m, A, x)
    assert abs(ctx.scalars(red, 1)[0] - cport.inner_vector_l2(m, x, ax)) <= 1e-13 * np.linalg.norm(x) * np.linalg.norm(ax)
    # reductions are deterministic: the same launch gives the same bits
    a1 = ctx.scalars(red, 1)[0]
    ctx.spmv(g, dA, dx, dy, w=dx, red=red, variant=variant)
    assert ctx.scalars(red, 1)[0] == a1
    for p in (dA, dx, db, dw, dy, red):
        ctx.free(p)


@pytest.mark.parametrize("npdim,nchunks", [((20, 14, 18), 4), ((19, 23, 11), 9), ((36, 36, 36), 1), ((12, 12, 40), 0)])
def test_spmv_host_buffers_pipelined(ctx, npdim, nchunks):
    """SMat::apply for host vectors: chunked upload / multiply / download equals the resident path bit for bit."""
    m, A = problem(npdim, "chebyshev", "convdiff", V1, 0.1)
    g = grid_of(m)
    x, b = rand_vec(m, 6), rand_vec(m, 7)
    dA, dx, dy, db = ctx.mat_from(g, A), ctx.vec_alloc(g), ctx.vec_alloc(g), ctx.vec_from(g, b)
    hy = np.full(m.vshape, 123.0)
    ctx.spmv_host(g, dA, x, hy, dx, dy, nchunks=nchunks)
    assert np.array_equal(hy, cport.apply(m, A, x))
    hy[...] = -7.0
    ctx.spmv_host(g, dA, x, hy, dx, dy, d_b=db, nchunks=nchunks)
    assert np.array_equal(hy, cport.apply_res(m, A, b, x))
    for p in (dA, dx, dy, db):
        ctx.free(p)


def test_spmv_stop_gate(ctx):
    m, A = problem((12, 12, 12))
    g = grid_of(m)
    x = rand_vec(m, 5)
    dA, dx, dy = ctx.mat_from(g, A), ctx.vec_from(g, x), ctx.vec_alloc(g)
    stop = ctx.ints_alloc(1)
    ctx.spmv(g, dA, dx, dy, stop=stop)
    assert np.array_equal(ctx.vec_download(g, dy), cport.apply(m, A, x))
    ctx.vecset(g, 7.0, dy)
    # use converge_test to raise the flag on the device: rr = 0 <= rtol
    rr = ctx.scalars_alloc(2)
    ctx.scalars_set(rr, [0.0, 1.0])
    iters = ctx.ints_alloc(1)
    ctx.converge_test(rr, rr + 8, 1e-6, 10, capi.CONV_INITIAL, stop, iters)
    assert ctx.ints(stop)[0] == 1
    ctx.spmv(g, dA, dx, dy, stop=stop)
    assert np.all(real(ctx.vec_download(g, dy)) == 7.0), "a stopped launch must not write"


@pytest.mark.parametrize("npdim", [(18, 18, 18), (19, 23, 11), (3, 3, 3), (131, 7, 5)])
def test_blas1(ctx, npdim):
    m, A = problem(npdim)
    g = grid_of(m)
    x, y = rand_vec(m, 10), rand_vec(m, 11)
    # ghosts of y non-zero on purpose: real-point-only operations must not touch them
    yg = rand_vec(m, 12, ghosts_zero=False)
    dx, dy, dyg, dw = ctx.vec_from(g, x), ctx.vec_from(g, y), ctx.vec_from(g, yg), ctx.vec_alloc(g)

    ref = y.copy(); cport.vecaxpy(m, 0.37, x, ref)
    ctx.vecaxpy(g, 0.37, dx, dy)
    assert np.array_equal(ctx.vec_download(g, dy), ref)

    ref = yg.copy(); cport.vecassign(m, x, ref)
    ctx.vecassign(g, dx, dyg)
    assert np.array_equal(ctx.vec_download(g, dyg), ref)          # ghosts of yg preserved
    ref2 = ref.copy(); cport.vecset(m, -2.5, ref2)
    ctx.vecset(g, -2.5, dyg)
    assert np.array_equal(ctx.vec_download(g, dyg), ref2)
    ctx.veccopy_all(g, dx, dyg)
    assert np.array_equal(ctx.vec_download(g, dyg), x)

    assert abs(ctx.dot(g, dx, dy) - cport.inner_vector_l2(m, x, ctx.vec_download(g, dy))) <= 1e-13 * np.linalg.norm(x) * np.linalg.norm(y) * 2
    assert abs(ctx.nrm2(g, dx) - cport.norm_vector_l2(m, x)) <= 1e-13 * np.linalg.norm(x)
    red = ctx.scalars_alloc(4)
    ctx.normL2sq_dev(g, m.coords, dx, red)
    assert abs(np.sqrt(ctx.scalars(red)[0]) - cport.norm_L2(m, x)) <= 1e-13 * cport.norm_L2(m, x)

    # w = a x + b y with fused norm; y = a x + b y; y = x + b y
    yh = ctx.vec_download(g, dy)
    ctx.vecwaxpby(g, 1.0, dx, -0.25, dy, dw, red)
    ref = x.copy(); cport.vecaxpy(m, -0.25, yh, ref)
    w = ctx.vec_download(g, dw)
    assert np.array_equal(real(w), real(ref))
    assert abs(ctx.scalars(red)[0] - cport.inner_vector_l2(m, ref, ref)) <= 1e-13 * cport.inner_vector_l2(m, ref, ref)
    ctx.vecaxpby(g, 1.0, dx, 0.5, dy)     # y = x + 0.5 y
    ref = x.copy(); cport.vecaxpy(m, 0.5, yh, ref)
    assert np.array_equal(real(ctx.vec_download(g, dy)), real(ref))

    # multi-axpy (one pass == the reference's n passes, same order) and multi-dot
    xs = [rand_vec(m, 20 + l) for l in range(5)]
    dxs = [ctx.vec_from(g, v) for v in xs]
    a = np.array([0.3, -1.2, 0.01, 2.0, -0.7])
    da = ctx.scalars_alloc(5); ctx.scalars_set(da, a)
    y0 = ctx.vec_download(g, dy)
    ref = y0.copy(); cport.vec_multi_axpy(m, a, xs, ref)
    ctx.vec_multi_axpy(g, da, dxs, dy)
    assert np.array_equal(ctx.vec_download(g, dy), ref)
    red8 = ctx.scalars_alloc(8)
    ctx.vec_multi_dot(g, dxs, dy, red8)
    got = ctx.scalars(red8, 5)
    for l in range(5):
        want = cport.inner_vector_l2(m, xs[l], ref)
        assert abs(got[l] - want) <= 1e-13 * np.linalg.norm(xs[l]) * np.linalg.norm(ref)


@pytest.mark.parametrize("npdim", [(18, 18, 18), (19, 23, 11), (4, 9, 3)])
def test_device_scalars(ctx, npdim):
    """axpy with alpha = -(num/den) taken from device memory equals the host-scalar result bitwise."""
    m, _ = problem(npdim)
    g = grid_of(m)
    x, y = rand_vec(m, 30), rand_vec(m, 31)
    dx, dy = ctx.vec_from(g, x), ctx.vec_from(g, y)
    s = ctx.scalars_alloc(2); ctx.scalars_set(s, [3.0, 7.0])
    ctx.vecaxpy(g, capi.scalar(-1.0, s, s + 8), dx, dy)
    ref = y.copy(); cport.vecaxpy(m, -(3.0 / 7.0), x, ref)
    assert np.array_equal(ctx.vec_download(g, dy), ref)


# sizes chosen to hit full tiles (32x8x8), ragged tile edges in every direction, and grids smaller than one tile
@pytest.mark.parametrize("npdim", [(18, 18, 18), (19, 23, 11), (3, 3, 3), (36, 20, 28), (34, 10, 10), (67, 19, 27), (99, 12, 9)])
@pytest.mark.parametrize("pde", ["poisson", "convdiff"])
def test_jacobi_and_sgs(ctx, npdim, pde):
    m, A = problem(npdim, "chebyshev", pde, V1, 0.1)
    g = grid_of(m)
    r = rand_vec(m, 40)
    dA, dr, dz, dy = ctx.mat_from(g, A), ctx.vec_from(g, r), ctx.vec_alloc(g), ctx.vec_alloc(g)
    ctx.jacobi_apply(g, dA, dr, dz)
    assert np.array_equal(ctx.vec_download(g, dz), cport.jacobi_apply(m, A, r))
    dinv = ctx.scalars_alloc(int(g.cstride))
    ctx.sgs_setup(g, dA, dinv)
    di = cport.sgs_setup(m, A)
    ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_LEXICOGRAPHIC, 1)
    z = ctx.vec_download(g, dz)
    ref = cport.sgs_like_apply(m, A, di, r, 1)
    assert np.array_equal(z, ref), f"lexicographic (dataflow) SGS max diff {np.abs(z - ref).max()}"
    assert ghosts_untouched(z)
    ctx.vecset(g, 9.0, dz)
    ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_LEXICOGRAPHIC, 3)      # sweeps > 1 are idempotent
    assert np.array_equal(ctx.vec_download(g, dz), ref)
    ctx.vecset(g, -9.0, dz)
    ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_LEXICOGRAPHIC_PLANES, 1)
    assert np.array_equal(ctx.vec_download(g, dz), ref), "per-hyperplane SGS differs"
    # synchronous Jacobi sweeps: equal to the oracle's statement sweep for sweep, and to the exact sweep once n >= nx+ny+nz-2
    nx, ny, nz = m.n
    for n in (1, 2, 3, nx + ny + nz - 2, nx + ny + nz - 1):
        ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_JACOBI_SWEEPS, n)
        assert np.array_equal(ctx.vec_download(g, dz), cport.sgs_jacobi_sweeps_apply(m, A, di, r, n)), f"{n} Jacobi sweeps differ"
    assert np.array_equal(ctx.vec_download(g, dz), cport.sgs_like_apply(m, A, di, r, 1)), "enough sweeps must reproduce the exact solve"
    # asynchronous sweeps (the reference's threaded apply, s3d_sgspreconditioners.cpp:30-112): timing-dependent, so no bitwise statement.
    #  - ONE chunk holds every row: a single warp walks them in order, which IS the sequential sweep up to the rounding of the row scan;
    #  - small chunks, many warps: every further sweep can only bring the result closer, and after nx+ny+nz sweeps the triangular
    #    systems are solved whatever the order of the updates was.
    scale = np.abs(ref).max()
    try:
        ctx.tune("sgs_async_chunk", ny * nz)
        ctx.vecset(g, 5.0, dz)
        ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_ASYNC, 1)
        za = ctx.vec_download(g, dz)
        assert ghosts_untouched(za)
        assert np.abs(za - ref).max() <= 1e-12 * scale, f"one-chunk asynchronous sweep: {np.abs(za - ref).max() / scale}"
        ctx.tune("sgs_async_chunk", 2)
        ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_ASYNC, nx + ny + nz)
        za = ctx.vec_download(g, dz)
        assert np.abs(za - ref).max() <= 1e-12 * scale, f"asynchronous sweeps at their fixed point: {np.abs(za - ref).max() / scale}"
        ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_ASYNC, 1)
        z1 = ctx.vec_download(g, dz)
        assert np.all(np.isfinite(z1)) and ghosts_untouched(z1)
    finally:
        ctx.tune("sgs_async_chunk", 4)
    # asynchronous StrILU build (the reference's threaded build, s3d_ilu.cpp:28-32): from d = diag(A), any order of updates reaches the
    # exact ILU(0) diagonal after at most nx+ny+nz sweeps; fewer sweeps give something between the diagonal and it
    dinv3 = ctx.scalars_alloc(int(g.cstride))
    try:
        want = cport.strilu_setup(m, A, 1)
        ctx.tune("sgs_async_chunk", 3)
        ctx.strilu_setup_async(g, dA, nx + ny + nz, dinv3)
        ctx.sgs_apply(g, dA, dinv3, dr, dz, dy, capi.SGS_LEXICOGRAPHIC, 1)
        zs = ctx.vec_download(g, dz)
        refs = cport.sgs_like_apply(m, A, want, r, 1)
        assert np.abs(zs - refs).max() <= 1e-12 * np.abs(refs).max(), "asynchronous StrILU build at its fixed point differs from the exact factorisation"
        ctx.strilu_setup_async(g, dA, 1, dinv3)
        ctx.sgs_apply(g, dA, dinv3, dr, dz, dy, capi.SGS_LEXICOGRAPHIC, 1)
        assert np.all(np.isfinite(ctx.vec_download(g, dz)))
    finally:
        ctx.tune("sgs_async_chunk", 4)
        ctx.free(dinv3)
    # the row-scan kernel (sgs_variant 7): the lexicographic sweep with every row solved by an affine warp scan -- the reference's
    # result to rounding, not bit for bit
    try:
        for shape in (0, 101, 204, 404):
            ctx.tune("sgs_variant", 7); ctx.tune("sgs_pencil_shape", shape)
            for rep in range(2):
                ctx.vecset(g, 2.0 + rep, dz)
                ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_LEXICOGRAPHIC, 1)
                zr = ctx.vec_download(g, dz)
                assert ghosts_untouched(zr)
                assert np.abs(zr - ref).max() <= 1e-13 * scale, f"row-scan sweep (shape {shape}, rep {rep}): {np.abs(zr - ref).max() / scale}"
    finally:
        ctx.tune("sgs_variant", capi.SGS_VARIANT_DEFAULT); ctx.tune("sgs_pencil_shape", 0)
    # StrILU: the ILU(0) diagonal (same dataflow kernel, divisions) and its application
    dinv2 = ctx.scalars_alloc(int(g.cstride))
    for sweeps in (1, 3, 0):
        ctx.strilu_setup(g, dA, sweeps, dinv2)
        want = cport.strilu_setup(m, A, sweeps)
        ctx.sgs_apply(g, dA, dinv2, dr, dz, dy, capi.SGS_LEXICOGRAPHIC, 1)
        assert np.array_equal(ctx.vec_download(g, dz), cport.sgs_like_apply(m, A, want, r, 1)), f"StrILU({sweeps} build sweeps) differs"
    ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_REDBLACK, 1)
    z = ctx.vec_download(g, dz)
    ref = cport.sgs_redblack_apply(m, A, di, r)
    assert np.array_equal(z, ref), f"red-black SGS max diff {np.abs(z - ref).max()}"
    assert ghosts_untouched(z)
    packed = ctx.sgs_rb_pack(g, dA, dinv)                      # the same on colour-packed coefficients
    ctx.vecset(g, 4.0, dz)
    ctx.sgs_apply_rb_packed(g, packed, dr, dz, dy)
    zp = ctx.vec_download(g, dz)
    ctx.free(packed)
    assert np.array_equal(zp, ref), f"red-black SGS on packed coefficients: max diff {np.abs(zp - ref).max()}"
    assert ghosts_untouched(zp)


# every build of the exact dataflow sweep: block-per-tile (variant 1), warp-per-tile (variant 0) with 16- or 32-long tiles and one
# or two pencils per lane, streaming (5), register-streaming (6, three block shapes); full, ragged and sub-tile grids; forward/backward sweep and the StrILU recurrence
@pytest.mark.parametrize("npdim", [(18, 18, 18), (35, 11, 20), (3, 3, 3), (50, 23, 14), (131, 9, 6), (21, 38, 27)])
@pytest.mark.parametrize("variant,tile_i,pencils", [(5, 0, 0), (1, 0, 0), (0, 16, 1), (0, 16, 2), (0, 32, 1), (0, 32, 2), (3, 0, 0), (6, 0, 0), (6, 22, 0),
                                                    (6, 11, 0)])
def test_sgs_dataflow_variants(ctx, npdim, variant, tile_i, pencils):
    m, A = problem(npdim, "chebyshev", "convdiff", V1, 0.1)
    g = grid_of(m)
    r = rand_vec(m, 41)
    dA, dr, dz, dy = ctx.mat_from(g, A), ctx.vec_from(g, r), ctx.vec_alloc(g), ctx.vec_alloc(g)
    dinv = ctx.scalars_alloc(int(g.cstride))
    try:
        if variant == 6:   # the register-streaming kernel: tile_i carries its block shape (warps as WJ*10 + WK, 0 = default 4 x 2)
            ctx.tune("sgs_variant", 6); ctx.tune("sgs_pencil_shape", tile_i)
        else:
            ctx.tune("sgs_variant", variant); ctx.tune("sgs_tile_i", tile_i); ctx.tune("sgs_pencils", pencils)
        ctx.sgs_setup(g, dA, dinv)
        di = cport.sgs_setup(m, A)
        ref = cport.sgs_like_apply(m, A, di, r, 1)
        for rep in range(2):                                    # twice: the flags of the first sweep must not satisfy the second
            ctx.vecset(g, 3.0 + rep, dz)
            ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_LEXICOGRAPHIC, 1)
            z = ctx.vec_download(g, dz)
            assert np.array_equal(z, ref), f"max diff {np.abs(z - ref).max()} (rep {rep})"
            assert ghosts_untouched(z)
        ctx.strilu_setup(g, dA, 1, dinv)
        want = cport.strilu_setup(m, A, 1)
        ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_LEXICOGRAPHIC, 1)
        assert np.array_equal(ctx.vec_download(g, dz), cport.sgs_like_apply(m, A, want, r, 1)), "StrILU differs"
    finally:
        ctx.tune("sgs_variant", capi.SGS_VARIANT_DEFAULT); ctx.tune("sgs_tile_i", 0); ctx.tune("sgs_pencils", 0); ctx.tune("sgs_pencil_shape", 0)
        for p_ in (dA, dr, dz, dy, dinv):
            ctx.free(p_)


# The warp-per-tile kernel with TWO staging areas (tuning key sgs_dbuf: the next tile's copies are issued before this tile's sweep):
# same results bit for bit, for the exact sweep, the blocks ordering and the StrILU build.
@pytest.mark.parametrize("npdim", [(18, 18, 18), (50, 23, 14), (131, 9, 6), (21, 38, 27), (3, 3, 3)])
@pytest.mark.parametrize("tile_i,pencils", [(16, 1), (16, 2), (32, 1), (32, 2)])
def test_sgs_double_buffered_staging(ctx, npdim, tile_i, pencils):
    m, A = problem(npdim, "chebyshev", "convdiff", V1, 0.1)
    g = grid_of(m)
    r = rand_vec(m, 45)
    di = cport.sgs_setup(m, A)
    exact = cport.sgs_like_apply(m, A, di, r, 1)
    blocks = cport.sgs_blocks_apply(m, A, di, r, 2, 2)
    dA, dr, dz, dy = ctx.mat_from(g, A), ctx.vec_from(g, r), ctx.vec_alloc(g), ctx.vec_alloc(g)
    dinv = ctx.scalars_alloc(int(g.cstride))
    try:
        ctx.tune("sgs_variant", 0); ctx.tune("sgs_tile_i", tile_i); ctx.tune("sgs_pencils", pencils); ctx.tune("sgs_dbuf", 1)
        ctx.tune("sgs_blocks", 2); ctx.tune("sgs_blocks_y", 2)
        ctx.sgs_setup(g, dA, dinv)
        for rep in range(2):
            ctx.vecset(g, 1.0 + rep, dz)
            ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_LEXICOGRAPHIC, 1)
            z = ctx.vec_download(g, dz)
            assert np.array_equal(z, exact), f"exact sweep, two staging areas, application {rep}: max diff {np.abs(z - exact).max()}"
            assert ghosts_untouched(z)
            ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_BLOCKS, 1)
            assert np.array_equal(ctx.vec_download(g, dz), blocks), f"blocks, two staging areas, application {rep}"
        ctx.strilu_setup(g, dA, 1, dinv)
        ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_LEXICOGRAPHIC, 1)
        assert np.array_equal(ctx.vec_download(g, dz), cport.sgs_like_apply(m, A, cport.strilu_setup(m, A, 1), r, 1)), "StrILU build, two staging areas"
    finally:
        ctx.tune("sgs_variant", capi.SGS_VARIANT_DEFAULT); ctx.tune("sgs_tile_i", 0); ctx.tune("sgs_pencils", 0); ctx.tune("sgs_dbuf", 0)
        ctx.tune("sgs_blocks", 0); ctx.tune("sgs_blocks_y", 0)
        for p_ in (dA, dr, dz, dy, dinv):
            ctx.free(p_)


# Blocks on one device (S3D_SGS_BLOCKS): concurrent sweeps over blocks of planes (and of rows) with one plane / row of overlap, through
# the layer tables of the warp-per-tile kernel.  Bit-identical to the oracle's statement on top of the reference's sequential half-sweeps;
# ragged grids, blocks of unequal thickness, layers that do not fill a tile, more blocks asked for than the grid allows (clamped).
@pytest.mark.parametrize("npdim,nblocks,nblocks_y", [((37, 22, 41), 3, 1), ((19, 35, 18), 2, 2), ((52, 27, 34), 4, 3), ((20, 12, 74), 8, 1), ((34, 34, 34), 16, 16),
                                                     ((67, 19, 27), 2, 1), ((23, 66, 12), 1, 4), ((40, 51, 47), 3, 5)])
@pytest.mark.parametrize("tile_i,pencils", [(16, 2), (16, 1), (32, 2)])
def test_sgs_blocks_on_one_device(ctx, npdim, nblocks, nblocks_y, tile_i, pencils):
    m, A = problem(npdim, "chebyshev", "convdiff", V1, 0.1)
    g = grid_of(m)
    r = rand_vec(m, 44)
    di = cport.sgs_setup(m, A)
    want = cport.sgs_blocks_apply(m, A, di, r, nblocks, nblocks_y)
    exact = cport.sgs_like_apply(m, A, di, r, 1)
    assert not np.array_equal(want, exact)
    dA, dr, dz, dy = ctx.mat_from(g, A), ctx.vec_from(g, r), ctx.vec_alloc(g), ctx.vec_alloc(g)
    dinv = ctx.scalars_alloc(int(g.cstride))
    try:
        ctx.tune("sgs_variant", 0); ctx.tune("sgs_tile_i", tile_i); ctx.tune("sgs_pencils", pencils)
        ctx.tune("sgs_blocks", nblocks); ctx.tune("sgs_blocks_y", nblocks_y)
        ctx.sgs_setup(g, dA, dinv)
        for rep in range(3):                                 # repeated: the mailboxes are clean after every sweep
            ctx.vecset(g, 3.0 + rep, dz)
            ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_BLOCKS, 1)
            z = ctx.vec_download(g, dz)
            assert np.array_equal(z, want), f"{nblocks} x {nblocks_y} blocks, application {rep}: max diff {np.abs(z - want).max()}"
            assert ghosts_untouched(z)
        # the exact sweep right after it, and the blocks again: the two tilings share mailboxes and dispatch state
        ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_LEXICOGRAPHIC, 1)
        assert np.array_equal(ctx.vec_download(g, dz), exact)
        ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_BLOCKS, 1)
        assert np.array_equal(ctx.vec_download(g, dz), want)
        ctx.tune("sgs_blocks", 1); ctx.tune("sgs_blocks_y", 1)   # one block IS the exact sweep
        ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_BLOCKS, 1)
        assert np.array_equal(ctx.vec_download(g, dz), exact)
    finally:
        ctx.tune("sgs_variant", capi.SGS_VARIANT_DEFAULT); ctx.tune("sgs_tile_i", 0); ctx.tune("sgs_pencils", 0); ctx.tune("sgs_blocks", 0); ctx.tune("sgs_blocks_y", 0)
        for p_ in (dA, dr, dz, dy, dinv):
            ctx.free(p_)


# The overlapping z-slab ordering (S3D_SGS_OVERLAP), every slab of a decomposition on ONE device: a z-slab grid descriptor without a
# communicator, the neighbours' planes put in place by hand (ghost planes of r by the upload, slack planes of A and dinv by
# s3d_coef_slack_upload) -- exactly what the halo exchanges do between ranks (tests/mgpu_check.py covers those).  Bit-identical to the
# oracle's statement: the sequential sweep over the slab extended by one plane of either neighbour, own planes kept.
@pytest.mark.parametrize("npdim,nslabs,blocks", [((37, 22, 41), 3, (1, 1)), ((19, 35, 18), 2, (1, 1)), ((52, 27, 34), 4, (1, 1)), ((20, 12, 10), 8, (1, 1)),
                                                 ((37, 22, 70), 3, (2, 2)), ((21, 50, 41), 2, (2, 3))])
@pytest.mark.parametrize("tile_i,pencils,mail", [(16, 2, 1), (16, 1, 1), (32, 2, 1), (16, 2, 0)])
def test_sgs_overlap_slabs_on_one_device(ctx, npdim, nslabs, blocks, tile_i, pencils, mail):
    if blocks != (1, 1) and not mail:
        pytest.skip("blocks inside the slabs need the mailbox hand-off")
    m, A = problem(npdim, "chebyshev", "convdiff", V1, 0.1)
    nx, ny, nz = m.n
    r = rand_vec(m, 43)
    di = cport.sgs_setup(m, A)
    want = cport.sgs_overlap_apply(m, A, di, r, nslabs, 1, blocks)      # blocks: every extended slab cut further into concurrent blocks
    local = cport.sgs_overlap_apply(m, A, di, r, nslabs, 0)
    assert not np.array_equal(want, local) and (blocks == (1, 1) or not np.array_equal(want, cport.sgs_overlap_apply(m, A, di, r, nslabs, 1)))
    try:
        ctx.tune("sgs_variant", 0); ctx.tune("sgs_tile_i", tile_i); ctx.tune("sgs_pencils", pencils); ctx.tune("sgs_tile_mail", mail)
        ctx.tune("sgs_blocks", blocks[0]); ctx.tune("sgs_blocks_y", blocks[1])
        for rk, (k0, k1) in enumerate(cport.slab_bounds(nz, nslabs)):
            g = capi.make_grid(nx, ny, nz, rk, nslabs)
            assert (g.k0, g.nz) == (k0, k1 - k0) and g.cstride >= g.cplane * (g.nz + 2)
            dA = ctx.mat_from(g, A[:, k0:k1])
            dinv = ctx.coef_alloc(g)
            ctx.sgs_setup(g, dA, dinv)
            for side, kk in ((0, k0 - 1), (1, k1)):
                if 0 <= kk < nz:
                    for st in range(7):
                        ctx.coef_slack_upload(g, dA, st, side, A[st, kk])
                    ctx.coef_slack_upload(g, dinv, 0, side, di[kk])
            dr, dz, dy = ctx.vec_from(g, r[k0:k1 + 2]), ctx.vec_alloc(g), ctx.vec_alloc(g)
            for rep in range(2):                             # twice: the mailboxes / flags of the first application are clean
                ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_OVERLAP, 1)
                z = ctx.vec_download(g, dz)
                assert np.array_equal(z[1:-1], want[1 + k0:1 + k1]), f"slab {rk} of {nslabs}, application {rep}: max diff {np.abs(z[1:-1] - want[1 + k0:1 + k1]).max()}"
                assert not z[1:-1, 0].any() and not z[1:-1, -1].any() and not z[1:-1, :, 0].any() and not z[1:-1, :, -1].any()
            # the same slab without the overlap: the slab-local ordering
            ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_SLABLOCAL, 1)
            assert np.array_equal(ctx.vec_download(g, dz)[1:-1], local[1 + k0:1 + k1]), f"slab-local, slab {rk}"
            for p_ in (dA, dinv, dr, dz, dy):
                ctx.free(p_)
    finally:
        ctx.tune("sgs_variant", capi.SGS_VARIANT_DEFAULT); ctx.tune("sgs_tile_i", 0); ctx.tune("sgs_pencils", 0); ctx.tune("sgs_tile_mail", 1)
        ctx.tune("sgs_blocks", 0); ctx.tune("sgs_blocks_y", 0)


@pytest.mark.parametrize("npdim,grid", [((18, 18, 18), "uniform"), ((16, 16, 16), "chebyshev"), ((19, 23, 11), "chebyshev"), ((3, 5, 4), "uniform")])
def test_assembly_bit_exact(ctx, npdim, grid):
    m = cport.Mesh(npdim, grid=grid)
    g = grid_of(m)
    dA = ctx.mat_alloc(g)
    ctx.assemble_poisson(g, m.coords, dA)
    assert np.array_equal(ctx.mat_download(g, dA), cport.lhs(m, "poisson"))
    for vel, mu in ((V1, 0.1), (V2, 0.05), ((0.0, -1.0, 2.0), 1.0)):
        ctx.assemble_convdiff(g, m.coords, vel, mu, dA)
        assert np.array_equal(ctx.mat_download(g, dA), cport.lhs(m, "convdiff", vel, mu)), (vel, mu)
    for mu in (0.07, 1.0):
        ctx.assemble_convdiff_circ(g, m.coords, mu, dA)
        assert np.array_equal(ctx.mat_download(g, dA), cport.lhs(m, "convdiff_circular", (1.0, 0, 0), mu)), ("circ", mu)
    # uniform Poisson closed form (test/nativevectorassemblytest.cpp:117-123): diag 6/dx^2, off -1/dx^2 within 100 eps
    if grid == "uniform" and npdim[0] == npdim[1] == npdim[2]:
        ctx.assemble_poisson(g, m.coords, dA)
        A = ctx.mat_download(g, dA)
        dx = m.coords[0][2] - m.coords[0][1]
        assert np.all(np.abs(A[3] - 6.0 / dx**2) <= 100 * np.finfo(float).eps * 6.0 / dx**2 + 100 * np.finfo(float).eps)
        for s in (0, 1, 2, 4, 5, 6):
            assert np.all(np.abs(A[s] + 1.0 / dx**2) <= 100 * np.finfo(float).eps * (1.0 / dx**2) + 100 * np.finfo(float).eps)


def _tables(m, pde, func, vel, mu):
    """Separable tables with the reference's rounding sequence folded into the X table."""
    PI = 3.141592653589793238
    cx, cy, cz = m.coords
    if func == "test_rhs":
        return [np.sin(PI * cx)], [np.sin(PI * cy)], [np.sin(PI * cz)]
    s = [np.sin(2 * PI * c) for c in (cx, cy, cz)]
    c = [np.cos(2 * PI * c_) for c_ in (cx, cy, cz)]
    if func == "manufactured_solution":
        return [s[0]], [s[1]], [s[2]]
    if pde == "poisson":
        return [12 * PI * PI * s[0]], [s[1]], [s[2]]
    tx, ty, tz = [mu * 12 * PI * PI * s[0]], [s[1]], [s[2]]
    for d in range(3):
        f = [c[e] if e == d else s[e] for e in range(3)]
        tx.append(2 * PI * vel[d] * f[0]); ty.append(f[1]); tz.append(f[2])
    return tx, ty, tz


@pytest.mark.parametrize("npdim,grid", [((18, 18, 18), "uniform"), ((19, 23, 11), "chebyshev")])
def test_rhs_separable_bit_exact(ctx, npdim, grid):
    m = cport.Mesh(npdim, grid=grid)
    g = grid_of(m)
    df = ctx.vec_alloc(g)
    for pde, func in (("poisson", "test_rhs"), ("poisson", "manufactured_solution"), ("poisson", "manufactured_rhs"), ("convdiff", "manufactured_rhs")):
        tx, ty, tz = _tables(m, pde, func, V1, 0.1)
        ctx.assemble_rhs_separable(g, np.stack(tx), np.stack(ty), np.stack(tz), df)
        got = ctx.vec_download(g, df)
        ref = cport.grid_function(m, pde, func, V1, 0.1)
        assert np.array_equal(got, ref), (pde, func, np.abs(got - ref).max())
        assert ghosts_untouched(got)


def test_fused_cg_and_bicg_steps(ctx):
    m, A = problem((20, 14, 18), "uniform", "poisson")
    g = grid_of(m)
    p, q, x, r = (rand_vec(m, 50 + i) for i in range(4))
    dA = ctx.mat_from(g, A)
    dp, dq, dx, dr = (ctx.vec_from(g, v) for v in (p, q, x, r))
    red = ctx.scalars_alloc(4)
    s = ctx.scalars_alloc(2); ctx.scalars_set(s, [2.0, 5.0])
    alpha = 2.0 / 5.0
    ctx.cg_update(g, capi.scalar(1.0, s, s + 8), dp, dq, dx, dr, dA, red)
    xr, rr_ = x.copy(), r.copy()
    cport.vecaxpy(m, alpha, p, xr); cport.vecaxpy(m, -alpha, q, rr_)
    assert np.array_equal(ctx.vec_download(g, dx), xr) and np.array_equal(ctx.vec_download(g, dr), rr_)
    got = ctx.scalars(red, 2)
    z = cport.jacobi_apply(m, A, rr_)
    assert abs(got[0] - cport.inner_vector_l2(m, rr_, rr_)) <= 1e-13 * got[0]
    assert abs(got[1] - cport.inner_vector_l2(m, rr_, z)) <= 1e-13 * abs(got[1])
    # p = r/diag + beta p
    ctx.cg_direction_jacobi(g, capi.scalar(0.75), dA, dr, dp)
    pref = z.copy(); cport.vecaxpy(m, 0.75, p, pref)
    assert np.array_equal(real(ctx.vec_download(g, dp)), real(pref))
    # BiCGSTAB direction and update
    v = rand_vec(m, 60); dv = ctx.vec_from(g, v)
    pn = ctx.vec_download(g, dp)
    ctx.bicg_direction(g, capi.scalar(0.3), capi.scalar(1.7), dr, dv, dp)
    t = pn.copy(); cport.vecaxpy(m, -1.7, v, t)
    ref = rr_.copy(); cport.vecaxpy(m, 0.3, t, ref)
    assert np.array_equal(real(ctx.vec_download(g, dp)), real(ref))
    ph, sh, sv, tv, rh = (rand_vec(m, 70 + i) for i in range(5))
    dph, dsh, dsv, dtv, drh = (ctx.vec_from(g, u) for u in (ph, sh, sv, tv, rh))
    ctx.bicg_update(g, capi.scalar(0.4), dph, capi.scalar(-0.9), dsh, dsv, dtv, drh, dx, dr, red)
    xref = xr.copy(); cport.vecaxpy(m, 0.4, ph, xref); cport.vecaxpy(m, -0.9, sh, xref)
    rref = sv.copy(); cport.vecaxpy(m, 0.9, tv, rref)
    assert np.array_equal(real(ctx.vec_download(g, dx)), real(xref)) and np.array_equal(real(ctx.vec_download(g, dr)), real(rref))
    got = ctx.scalars(red, 2)
    assert abs(got[0] - cport.inner_vector_l2(m, rref, rref)) <= 1e-13 * got[0]
    assert abs(got[1] - cport.inner_vector_l2(m, rh, rref)) <= 1e-13 * np.linalg.norm(rh) * np.linalg.norm(rref)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "fields_*.npz")) + glob.glob(os.path.join(GOLD, "circ_*.npz"))))
def test_against_reference_fixtures(ctx, path):
    """Golden vectors produced by the REAL reference (oracle/make_golden.py): coefficients, y = Ax,
    b - Ax, Jacobi, sequential SGS, axpy, norms."""
    G = np.load(path)
    npdim = tuple(int(v) for v in G["npdim"])
    g = capi.make_grid(npdim[0] - 2, npdim[1] - 2, npdim[2] - 2)
    coords = (G["cx"], G["cy"], G["cz"])
    dA = ctx.mat_alloc(g)
    if str(G["pde"]) == "poisson":
        ctx.assemble_poisson(g, coords, dA)
    elif str(G["pde"]) == "convdiff_circular":
        ctx.assemble_convdiff_circ(g, coords, float(G["mu"]), dA)
    else:
        ctx.assemble_convdiff(g, coords, G["vel"], float(G["mu"]), dA)
    assert np.array_equal(ctx.mat_download(g, dA), G["A"])
    dx, db, dy, dz, ds = ctx.vec_from(g, G["x"]), ctx.vec_from(g, G["manufactured_rhs"]), ctx.vec_alloc(g), ctx.vec_alloc(g), ctx.vec_alloc(g)
    ctx.spmv(g, dA, dx, dy)
    assert np.array_equal(ctx.vec_download(g, dy), G["Ax"])
    ctx.spmv(g, dA, dx, dy, b=db)
    assert np.array_equal(ctx.vec_download(g, dy), G["b_minus_Ax"])
    ctx.jacobi_apply(g, dA, db, dz)
    assert np.array_equal(ctx.vec_download(g, dz), G["jacobi_b"])
    dinv = ctx.scalars_alloc(int(g.cstride))
    ctx.sgs_setup(g, dA, dinv)
    ctx.sgs_apply(g, dA, dinv, db, dz, ds, capi.SGS_LEXICOGRAPHIC, 1)
    assert np.array_equal(ctx.vec_download(g, dz), G["sgs_b"])
    ctx.strilu_setup(g, dA, 1, dinv)
    ctx.sgs_apply(g, dA, dinv, db, dz, ds, capi.SGS_LEXICOGRAPHIC, 1)
    assert np.array_equal(ctx.vec_download(g, dz), G["strilu_b"])
    ctx.vecassign(g, db, dy)
    ctx.vecaxpy(g, 0.37, dx, dy)
    assert np.array_equal(real(ctx.vec_download(g, dy)), real(G["b_plus_037x"]))
    assert abs(ctx.dot(g, dx, db) - float(G["dot_x_b"])) <= 1e-13 * np.linalg.norm(G["x"]) * np.linalg.norm(G["manufactured_rhs"])
    assert abs(ctx.nrm2(g, dx) - float(G["norm_x"])) <= 1e-13 * float(G["norm_x"])
    red = ctx.scalars_alloc(1)
    ctx.normL2sq_dev(g, coords, dx, red)
    assert abs(np.sqrt(ctx.scalars(red)[0]) - float(G["normL2_x"])) <= 1e-13 * float(G["normL2_x"])


@pytest.mark.parametrize("n", [130, 258])
def test_sgs_dataflow_full_size(ctx, n):
    """At sizes where the CPU oracle is slow: the dataflow kernel must equal the per-hyperplane kernel bitwise
    (both are pinned to the oracle at small sizes), repeatedly (no stale flags between epochs), and M z = r must
    hold: r - (D+L) D^-1 (D+U) z == 0 is checked through A: for SGS, A = M - L D^-1 U, so only equality is used."""
    m = cport.Mesh((n, n, n))
    g = grid_of(m)
    dA = ctx.mat_alloc(g)
    ctx.assemble_convdiff(g, m.coords, V1, 0.1, dA)
    dr, dz, dz2, dy = (ctx.vec_alloc(g) for _ in range(4))
    ctx.vec_upload(g, dr, golden_vec(m))
    dinv = ctx.scalars_alloc(int(g.cstride))
    ctx.sgs_setup(g, dA, dinv)
    ctx.sgs_apply(g, dA, dinv, dr, dz2, dy, capi.SGS_LEXICOGRAPHIC_PLANES, 1)
    for rep in range(3):
        ctx.vecset(g, float(rep), dz)
        ctx.sgs_apply(g, dA, dinv, dr, dz, dy, capi.SGS_LEXICOGRAPHIC, 1)
        ctx.vecaxpy(g, -1.0, dz2, dz)
        assert ctx.nrm2(g, dz) == 0.0, f"dataflow SGS differs from the hyperplane SGS (rep {rep})"
    for p in (dA, dr, dz, dz2, dy, dinv):
        ctx.free(p)


@pytest.mark.parametrize("n", [258, 514])
def test_spmv_full_size_properties(ctx, n):
    """At BASELINE sizes the oracle is too slow to run per test; use size-independent properties:
    (1) ||A*1||_2 known answers from the reference, (2) linearity,
    (3) apply and apply_res are consistent (test/matvectest.cpp:71), (4) all variants agree bitwise."""
    m = cport.Mesh((n, n, n))
    g = grid_of(m)
    dA = ctx.mat_alloc(g)
    ctx.assemble_poisson(g, m.coords, dA)
    dx, dy, dy2, db = (ctx.vec_alloc(g) for _ in range(4))
    ctx.vecset(g, 1.0, dx)
    ctx.spmv(g, dA, dx, dy)
    known = {258: 10434905.903215431, 514: 82834297.872071102}[n]
    assert abs(ctx.nrm2(g, dy) - known) <= 1e-12 * known
    x = golden_vec(m)
    ctx.vec_upload(g, dx, x)
    ctx.spmv(g, dA, dx, dy, variant=capi.SPMV_MARCH)
    ctx.spmv(g, dA, dx, dy2, variant=capi.SPMV_NAIVE)
    ctx.vecaxpy(g, -1.0, dy, dy2)
    assert ctx.nrm2(g, dy2) == 0.0, "MARCH and NAIVE variants must agree bitwise"
    # apply_res consistency: (b - Ax) + Ax - b == 0 to rounding
    ctx.vecset(g, 0.25, db)
    ctx.spmv(g, dA, dx, dy2, b=db)
    ctx.vecaxpy(g, 1.0, dy, dy2)
    ctx.vecaxpy(g, -1.0, db, dy2)
    assert ctx.nrm2(g, dy2) <= 1e-12 * ctx.nrm2(g, dy)
    for p in (dA, dx, dy, dy2, db):
        ctx.free(p)
