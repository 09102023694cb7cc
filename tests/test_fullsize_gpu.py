"""Parity at BASELINE.json's sizes, field by field against the CPU oracle (not against another CUDA path):
256^3 -- SpMV (y = A x and b - A x), Jacobi, the exact SGS sweep with both dataflow kernels, for both PDEs, bit-exact; the CG + Jacobi
solve of BASELINE config 2 with the oracle's iteration count (460) and residual history; 512^3 -- the headline SpMV on the golden x,
bit-exact.  The oracle takes ~1 s per 256^3 field, ~35 s for the 256^3 CG solve and ~15 s for the 512^3 operator."""
import os

import numpy as np
import pytest

from oracle import cport
from struct3d_b200 import capi
from tests.util import V1, golden_vec

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _grid(m):
    return capi.make_grid(*m.n)


@pytest.mark.parametrize("pde", ["poisson", "convdiff"])
def test_256_fields_bit_exact_against_the_oracle(ctx, pde):
    m = cport.Mesh((258, 258, 258))
    A = cport.lhs(m, pde, V1, 0.1)
    g = _grid(m)
    x = golden_vec(m)
    b = x[::-1, ::-1, ::-1].copy()
    dA = ctx.mat_alloc(g)
    if pde == "poisson":
        ctx.assemble_poisson(g, m.coords, dA)
    else:
        ctx.assemble_convdiff(g, m.coords, V1, 0.1, dA)
    assert np.array_equal(ctx.mat_download(g, dA), A), "GPU assembly differs from the oracle at 256^3"
    dx, db, dy, dz, ds = ctx.vec_from(g, x), ctx.vec_from(g, b), ctx.vec_alloc(g), ctx.vec_alloc(g), ctx.vec_alloc(g)
    ctx.spmv(g, dA, dx, dy)
    assert np.array_equal(ctx.vec_download(g, dy), cport.apply(m, A, x)), "y = A x differs from the oracle at 256^3"
    ctx.spmv(g, dA, dx, dy, b=db)
    assert np.array_equal(ctx.vec_download(g, dy), cport.apply_res(m, A, b, x)), "b - A x differs from the oracle at 256^3"
    ctx.jacobi_apply(g, dA, dx, dz)
    assert np.array_equal(ctx.vec_download(g, dz), cport.jacobi_apply(m, A, x)), "Jacobi differs from the oracle at 256^3"
    dinv = ctx.scalars_alloc(int(g.cstride))
    ctx.sgs_setup(g, dA, dinv)
    ref = cport.sgs_like_apply(m, A, cport.sgs_setup(m, A), x, 1)
    try:
        for variant in (0, 5, 6):       # warp-per-tile dataflow kernel, streaming kernel, register-streaming kernel
            ctx.tune("sgs_variant", variant)
            ctx.vecset(g, 1.0, dz)
            ctx.sgs_apply(g, dA, dinv, dx, dz, ds, capi.SGS_LEXICOGRAPHIC, 1)
            assert np.array_equal(ctx.vec_download(g, dz), ref), f"exact SGS (kernel variant {variant}) differs from the oracle's sequential sweep at 256^3"
    finally:
        ctx.tune("sgs_variant", capi.SGS_VARIANT_DEFAULT)
        for p in (dA, dx, db, dy, dz, ds, dinv):
            ctx.free(p)


def test_256_cg_jacobi_matches_the_oracle_solve():
    """BASELINE config 2: Poisson 256^3, CG + Jacobi, golden right-hand side, rtol 1e-6: the oracle (the C restatement of the harness loop
    over reference primitives) needs 460 iterations; same count, same residual history, same solution."""
    from struct3d_b200 import hostapi as H
    H.init(0)
    cm = cport.Mesh((258, 258, 258))
    Ac = cport.lhs(cm, "poisson")
    bh = golden_vec(cm)
    xo, oi, hist = cport.solve(cm, Ac, bh, "cg", "jacobi", 1e-6, 2000)
    assert oi["iters"] == 460 and oi["converged"]
    m = H.Mesh((258, 258, 258))
    A = H.Pde("poisson").lhs(m)
    b, x = H.Vec(m), H.Vec(m)
    b.set(bh)
    info = H.solve(A, b, x, {"-s3d_ksp_type": "cg", "-s3d_pc_type": "jacobi", "-ksp_rtol": 1e-6, "-ksp_max_it": 2000}, capture=True)
    H.set_options({})
    assert info["converged"] and info["iters"] == 460
    h = np.asarray(info["hist"])
    assert len(h) == len(hist) and np.max(np.abs(h - hist) / hist) <= 1e-6, "residual history differs from the oracle's"
    assert np.abs(x.a - xo).max() <= 1e-9 * np.abs(xo).max()


def test_512_spmv_bit_exact_against_the_oracle(ctx):
    """The headline operation at the headline size: Poisson 512^3, x = the golden field, y = A x against the oracle."""
    m = cport.Mesh((514, 514, 514))
    A = cport.lhs(m, "poisson")
    g = _grid(m)
    x = golden_vec(m)
    dA, dx, dy = ctx.mat_alloc(g), ctx.vec_from(g, x), ctx.vec_alloc(g)
    ctx.assemble_poisson(g, m.coords, dA)
    try:
        for variant in (capi.SPMV_AUTO, capi.SPMV_TMA):
            ctx.vecset(g, 3.0, dy)
            ctx.spmv(g, dA, dx, dy, variant=variant)
            got = ctx.vec_download(g, dy)
            want = cport.apply(m, A, x)
            assert np.array_equal(got, want), f"y = A x (variant {variant}) differs from the oracle at 512^3"
            del got, want
    finally:
        for p in (dA, dx, dy):
            ctx.free(p)
