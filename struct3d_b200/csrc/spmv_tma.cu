// TMA-staged SpMV variant (S3D_SPMV_TMA):  y = A x  and  y = b - A x, reference SMat::apply / apply_res, src/linalg/matvec.cpp:59-146.
//
// The x values of a 64 (i) x 8 (j) tile reach shared memory through the TMA engine: per z-plane one cp.async.bulk of a 544-byte
// row piece (columns i0-2 .. i0+65) for each of the ten rows j0-1 .. j0+8, completion counted on an mbarrier, a ring of four planes
// so that the plane two steps ahead is in flight while the current one is multiplied.  Every x value is read from HBM/L2 once per
// tile column and all seven uses (i+-1, j+-1, k+-1, centre) come from shared memory; the seven coefficient streams, b and y are
// read / written directly with evict-first 128-bit accesses exactly as in the MARCH variant.  Same arithmetic, same roundings:
// bit-identical results.  It exists to MEASURE the shared-memory/TMA staging north_star names against the register/L1 march
// (profiles/r1_sweep_kernels.md): the march is the default because it is faster.
#include "common.cuh"

namespace {

constexpr int TX = 64, TY = 8, ROWS = TY + 2, RW = TX + 4, NSTAGE = 4;   // RW: columns i0-2 .. i0+65 (16-byte aligned piece)

__device__ __forceinline__ double stencil7(double a0, double a1, double a2, double a3, double a4, double a5, double a6,
                                           double x0, double x1, double x2, double x3, double x4, double x5, double x6)
{
	double s = __dmul_rn(a0, x0);
	s = __dadd_rn(s, __dmul_rn(a1, x1));
	s = __dadd_rn(s, __dmul_rn(a2, x2));
	s = __dadd_rn(s, __dmul_rn(a3, x3));
	s = __dadd_rn(s, __dmul_rn(a4, x4));
	s = __dadd_rn(s, __dmul_rn(a5, x5));
	s = __dadd_rn(s, __dmul_rn(a6, x6));
	return s;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned a, unsigned count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned a, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned a, unsigned parity)
{
	unsigned ok;
	asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned mbar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

// grid: (ceil(nx/64), ceil(ny/8), ceil(nz/kchunk)); block: (32, 8)
template <bool RES, bool DOTW, bool DOTYY>
__global__ void __launch_bounds__(256)
spmv_tma_kernel(const GridDev g, const double *__restrict__ A, const double *__restrict__ x, double *__restrict__ y,
                const double *__restrict__ b, const double *__restrict__ w, const RedOut ro, const int *stop, const int kchunk)
{
	if (s3d_stopped(stop)) return;
	__shared__ __align__(128) double sx[NSTAGE][ROWS][RW];
	__shared__ __align__(8) unsigned long long bar[NSTAGE];
	const int lane = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + lane;
	const int i0t = blockIdx.x * TX, j0 = blockIdx.y * TY;
	const int kbeg = blockIdx.z * kchunk, kend = min(g.nz, kbeg + kchunk);
	const int i0 = i0t + 2 * lane, j = j0 + ty;
	const bool active = (i0 < g.nx) && (j < g.ny);
	const int nrows = min(ROWS, g.ny - j0 + 2);                              // rows j0-1 .. min(j0+8, ny): the ghost row ny is the last one that exists
	const int cols = min(RW, (int)g.vpitch - 14 - i0t);                      // a piece never runs past the end of its row (even, >= nx - i0t + 3)
	const int nplanes = kend - kbeg + 2;                                     // planes kbeg-1 .. kend
	double acc[2] = {0.0, 0.0};

	auto issue_plane = [&](int n) {                                          // thread 0: plane kbeg-1+n into stage n % NSTAGE
		const unsigned ba = smem_u32(&bar[n % NSTAGE]);
		mbar_expect_tx(ba, (unsigned)(nrows * cols * 8));
		const double *src = x + g.voff + (long long)(kbeg - 1 + n) * g.vplane + (long long)(j0 - 1) * g.vpitch + (i0t - 2);
		for (int r = 0; r < nrows; r++) bulk_g2s(smem_u32(&sx[n % NSTAGE][r][0]), src + (long long)r * g.vpitch, (unsigned)(cols * 8), ba);
	};
	auto wait_plane = [&](int n) {
		const unsigned ba = smem_u32(&bar[n % NSTAGE]);
		const unsigned parity = (unsigned)(n / NSTAGE) & 1u;
		while (!mbar_try_wait(ba, parity)) { }
	};

	if (tid == 0) {
		for (int s = 0; s < NSTAGE; s++) mbar_init(smem_u32(&bar[s]), 1);
		asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
	}
	__syncthreads();
	if (tid == 0) { issue_plane(0); issue_plane(1); issue_plane(2); }       // nplanes >= 3
	wait_plane(0);
	wait_plane(1);

	const long long voffs = g.voff + (long long)kbeg * g.vplane + (long long)j * g.vpitch + i0;
	double *yp = y + voffs;
	const double *bp = RES ? b + voffs : nullptr;
	const double *wp = DOTW ? w + voffs : nullptr;
	const double *ap = A + (long long)kbeg * g.cplane + (long long)j * g.cpitch + i0;
	const long long cs = g.cstride;
	const bool pair = (i0 + 1 < g.nx);
	const int li = 2 * lane + 2, r = ty + 1;                                 // this thread's first point inside a staged plane

	for (int k = kbeg; k < kend; ++k) {
		const int n = k - kbeg;                                             // planes n (k-1), n+1 (k), n+2 (k+1)
		__syncthreads();                                                    // everybody is done with plane n-1: its stage can be refilled
		if (tid == 0 && n + 3 < nplanes) {
			asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
			issue_plane(n + 3);
		}
		wait_plane(n + 2);
		if (active) {
			const double (*pm)[RW] = sx[n % NSTAGE], (*pc)[RW] = sx[(n + 1) % NSTAGE], (*pn)[RW] = sx[(n + 2) % NSTAGE];
			const double2 a0 = ld_stream2(ap);
			const double2 a1 = ld_stream2(ap + cs);
			const double2 a2 = ld_stream2(ap + 2 * cs);
			const double2 a3 = ld_stream2(ap + 3 * cs);
			const double2 a4 = ld_stream2(ap + 4 * cs);
			const double2 a5 = ld_stream2(ap + 5 * cs);
			const double2 a6 = ld_stream2(ap + 6 * cs);
			double2 bv = make_double2(0.0, 0.0);
			if (RES) bv = ld_stream2(bp);
			const double2 xm = *reinterpret_cast<const double2 *>(&pm[r][li]);
			const double2 xc = *reinterpret_cast<const double2 *>(&pc[r][li]);
			const double2 xn = *reinterpret_cast<const double2 *>(&pn[r][li]);
			const double2 xjm = *reinterpret_cast<const double2 *>(&pc[r - 1][li]);
			const double2 xjp = *reinterpret_cast<const double2 *>(&pc[r + 1][li]);
			const double xl = pc[r][li - 1], xr = pc[r][li + 2];
			double y0 = stencil7(a0.x, a1.x, a2.x, a3.x, a4.x, a5.x, a6.x, xl, xjm.x, xm.x, xc.x, xc.y, xjp.x, xn.x);
			double y1 = stencil7(a0.y, a1.y, a2.y, a3.y, a4.y, a5.y, a6.y, xc.x, xjm.y, xm.y, xc.y, xr, xjp.y, xn.y);
			if (RES) { y0 = __dsub_rn(bv.x, y0); y1 = __dsub_rn(bv.y, y1); }
			if (pair) st_stream2(yp, make_double2(y0, y1));
			else *yp = y0;
			if (DOTW) {
				const double2 wv = ld2(wp);
				acc[0] += wv.x * y0;
				if (pair) acc[0] += wv.y * y1;
			}
			if (DOTYY) {
				acc[1] += y0 * y0;
				if (pair) acc[1] += y1 * y1;
			}
			yp += g.vplane; ap += g.cplane;
			if (RES) bp += g.vplane;
			if (DOTW) wp += g.vplane;
		}
	}
	if (DOTW || DOTYY) grid_reduce<2, true>(acc, ro);
}

}  // namespace

int s3d_spmv_tma(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *x, double *y, const double *b,
                 const double *w, bool yy, double *d_red, const int *stop)
{
	if (ctx->red_defer) return s3d_fail(ctx, S3D_ERR_ARG, "S3D_SPMV_TMA does not take part in split launches");
	const GridDev gd = to_dev(g);
	const int kchunk = std::min<int>(g->nz, 16);
	dim3 block(32, TY);
	dim3 grid((g->nx + TX - 1) / TX, (g->ny + TY - 1) / TY, (g->nz + kchunk - 1) / kchunk);
	const bool red = (w != nullptr) || yy;
	const size_t nblk = (size_t)grid.x * grid.y * grid.z;
	if (red) {
		const int rc = s3d_ensure_partials(ctx, nblk * 2 + 256);
		if (rc != S3D_OK) return rc;
	}
	RedOut ro{ctx->partials, nullptr, d_red};
#define GO(RES, DW, DY) spmv_tma_kernel<RES, DW, DY><<<grid, block, 0, ctx->stream>>>(gd, A, x, y, b, w, ro, stop ? stop : ctx->gate, kchunk)
	const int sel = (b ? 4 : 0) | (w ? 2 : 0) | (yy ? 1 : 0);
	switch (sel) {
	case 0: GO(false, false, false); break;
	case 1: GO(false, false, true); break;
	case 2: GO(false, true, false); break;
	case 3: GO(false, true, true); break;
	case 4: GO(true, false, false); break;
	case 5: GO(true, false, true); break;
	case 6: GO(true, true, false); break;
	default: GO(true, true, true); break;
	}
#undef GO
	S3D_LAUNCHED(ctx);
	if (red) return s3d_final_reduce(ctx, 2, (unsigned)nblk, d_red);
	return S3D_OK;
}
