// 7-point structured SpMV for sm_100a:  y = A x  and  y = b - A x
// (reference: SMat::apply / SMat::apply_res, src/linalg/matvec.cpp:59-146).
//
// HBM-bound: 72 B/point algorithmic (7 coefficient streams + x + y; 80 with b), 13 flop/point,
// so no tensor cores.  What matters is (1) every coefficient byte is read exactly once with
// 128-bit, 128-byte-aligned, evict-first loads, (2) x is fetched from HBM once although each value
// is used by seven points, (3) enough bytes in flight per SM to cover HBM latency.
//
// MARCH variant: a 32 x BY thread block owns a (64 i) x (BY j) column of the grid and marches
// through a chunk of z-planes.  Each thread owns two consecutive i-points (one double2), keeps its
// own x[k-1], x[k], x[k+1] in registers (so the k-neighbours cost no loads), gets x[i-1]/x[i+2] from
// the neighbouring lanes by warp shuffle, and reads the j-1 / j+1 rows with aligned double2 loads
// that hit L1/L2 (the rows were loaded as the centre rows of the previous plane step).  No shared
// memory and no block barrier in the streaming loop: 10 independent 128-bit loads per thread per plane.
//
// Arithmetic: the seven products are rounded separately and added in stencil order 0..6
// (__dmul_rn / __dadd_rn, never contracted to FMA), which is bit-for-bit the reference's expression
// when the reference is compiled without FMA contraction (oracle/Makefile).
#include "common.cuh"

namespace {

struct Tune { int kchunk = 0; int variant = S3D_SPMV_MARCH; int hints = 1; int minb = 0; };
Tune g_tune;

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

// grid: (ceil(nx/64), ceil(ny/BY), ceil(nz/kchunk)); block: (32, BY)
// The fused reductions only do stage 1 (per-block partials) in this kernel; the fixed-order sum over blocks
// is the separate s3d_final_reduce launch.  Inlining stage 2 here cost 102 registers (2 blocks/SM).
// HALO = true (z-slab runs, see s3d_spmv_exchange): the planes below k = 0 and above k = nz-1 are not x's ghost planes but the
// slots of this rank's peer-memory window, which the neighbours' push kernels fill while this kernel already runs.  The two
// k-chunks that touch them are dispatched LAST (block order remapped) and only their blocks wait for the window flags, so the
// exchange overlaps the interior planes tile by tile and needs no launch of its own on the receiving side.
template <int BY, int MINB, bool RES, bool DOTW, bool DOTYY, bool HALO = false>
__global__ void __launch_bounds__(32 * BY, MINB)
spmv_march_kernel(const GridDev g, const double *__restrict__ A, const double *__restrict__ x, double *__restrict__ y,
                  const double *__restrict__ b, const double *__restrict__ w, const RedOut ro, const int *stop,
                  const int kchunk, const s3d_halo_view hv = s3d_halo_view())
{
	if (s3d_stopped(stop)) return;
	const int lane = threadIdx.x;
	const int i0 = (blockIdx.x * 32 + lane) * 2;
	const int j = blockIdx.y * BY + threadIdx.y;
	int kz = blockIdx.z;
	if (HALO && gridDim.z >= 3) {   // chunk 0 and the last chunk go to the end of the dispatch order
		const int nz_ = (int)gridDim.z;
		kz = (kz < nz_ - 2) ? kz + 1 : (kz == nz_ - 2 ? 0 : nz_ - 1);
	}
	const int kbeg = kz * kchunk;
	const int kend = min(g.nz, kbeg + kchunk);
	if (HALO) {
		const bool need_lo = hv.lo && kbeg == 0, need_hi = hv.hi && kend == g.nz;
		if (need_lo || need_hi) {
			if (threadIdx.x == 0 && threadIdx.y == 0) {
				const long long t0 = clock64();
				while ((need_lo && *(const volatile int *)hv.flag_lo != hv.epoch) || (need_hi && *(const volatile int *)hv.flag_hi != hv.epoch))
					if (clock64() - t0 > (20ll << 30)) __trap();   // ~10 s: a neighbour died; fail loudly instead of hanging
				__threadfence_system();
			}
			__syncthreads();
		}
	}
	const bool active = (i0 < g.nx) && (j < g.ny);
	const unsigned mask = __ballot_sync(0xffffffffu, active);
	double acc[2] = {0.0, 0.0};

	if (active) {
		const bool pair = (i0 + 1 < g.nx);                                    // false only for the last point of an odd row
		const bool edge_lo = (lane == 0);
		const bool edge_hi = (lane == 31) || (((mask >> (lane + 1)) & 1u) == 0u);
		const long long voffs = g.voff + (long long)kbeg * g.vplane + (long long)j * g.vpitch + i0;
		const double *xp = x + voffs;
		double *yp = y + voffs;
		const double *bp = RES ? b + voffs : nullptr;
		const double *wp = DOTW ? w + voffs : nullptr;
		const double *ap = A + (long long)kbeg * g.cplane + (long long)j * g.cpitch + i0;
		const long long cs = g.cstride;

		const long long inplane = (g.voff - g.vplane) + (long long)j * g.vpitch + i0;   // the same point inside a window slot
		double2 xm = (HALO && hv.lo && kbeg == 0) ? __ldcg(reinterpret_cast<const double2 *>(hv.lo + inplane)) : ld2(xp - g.vplane);
		double2 xc = ld2(xp);
		auto plane = [&](const int k) {
			// the seven coefficient streams: read once, evict-first
			const double2 a0 = ld_stream2(ap);
			const double2 a1 = ld_stream2(ap + cs);
			const double2 a2 = ld_stream2(ap + 2 * cs);
			const double2 a3 = ld_stream2(ap + 3 * cs);
			const double2 a4 = ld_stream2(ap + 4 * cs);
			const double2 a5 = ld_stream2(ap + 5 * cs);
			const double2 a6 = ld_stream2(ap + 6 * cs);
			// x: the next plane (first touch of these values), and the two j-neighbour rows of this plane
			const double2 xn = (HALO && hv.hi && k == g.nz - 1) ? __ldcg(reinterpret_cast<const double2 *>(hv.hi + inplane)) : ld2(xp + g.vplane);
			const double2 xjm = ld2(xp - g.vpitch);
			const double2 xjp = ld2(xp + g.vpitch);
			double2 bv = make_double2(0.0, 0.0);
			if (RES) bv = ld_stream2(bp);
			// i-neighbours from the adjacent lanes; the two warp-edge lanes read them from memory
			double xl = __shfl_up_sync(mask, xc.y, 1);
			double xr = __shfl_down_sync(mask, xc.x, 1);
			if (edge_lo) xl = xp[-1];
			if (edge_hi) xr = xp[2];

			double y0 = stencil7(a0.x, a1.x, a2.x, a3.x, a4.x, a5.x, a6.x, xl, xjm.x, xm.x, xc.x, xc.y, xjp.x, xn.x);
			double y1 = stencil7(a0.y, a1.y, a2.y, a3.y, a4.y, a5.y, a6.y, xc.x, xjm.y, xm.y, xc.y, xr, xjp.y, xn.y);
			if (RES) { y0 = __dsub_rn(bv.x, y0); y1 = __dsub_rn(bv.y, y1); }
			if (pair) st_stream2(yp, make_double2(y0, y1));
			else *yp = y0;
			if (DOTW) {
				const double2 wv = (wp == xp) ? xc : ld2(wp);
				acc[0] += wv.x * y0;
				if (pair) acc[0] += wv.y * y1;
			}
			if (DOTYY) {
				acc[1] += y0 * y0;
				if (pair) acc[1] += y1 * y1;
			}
			xm = xc; xc = xn;
			xp += g.vplane; yp += g.vplane; ap += g.cplane;
			if (RES) bp += g.vplane;
			if (DOTW) wp += g.vplane;
		};
		if constexpr (DOTW || DOTYY) {
#pragma unroll 1   // unrolled, the fused-reduction forms take >100 registers
			for (int k = kbeg; k < kend; ++k) plane(k);
		} else {
			for (int k = kbeg; k < kend; ++k) plane(k);
		}
	}
	if (DOTW || DOTYY) grid_reduce<2, true>(acc, ro);
}


// PAIR variant: no marching.  Each thread computes two consecutive i-points of ONE plane; all five
// x rows it needs (centre, j-1, j+1, k-1, k+1) are aligned double2 loads that hit L1/L2 after their
// first touch, the i-neighbours come from the adjacent lanes.  12 independent 128-bit loads per thread,
// few registers, maximal parallelism: x-reuse is left to the 126 MB L2 (three planes = 6 MB at 512^2),
// which is protected by the evict-first policy of the coefficient / y streams.
// grid: (ceil(nx/64), ceil(ny/8), nz); block: (32, 8)
template <bool RES, bool DOTW, bool DOTYY>
__global__ void __launch_bounds__(256)
spmv_pair_kernel(const GridDev g, const double *__restrict__ A, const double *__restrict__ x, double *__restrict__ y,
                 const double *__restrict__ b, const double *__restrict__ w, const RedOut ro, const int *stop, const int hints)
{
	if (s3d_stopped(stop)) return;
	const int lane = threadIdx.x;
	const int i0 = (blockIdx.x * 32 + lane) * 2;
	const int j = blockIdx.y * 8 + threadIdx.y;
	const int k = blockIdx.z;
	const bool active = (i0 < g.nx) && (j < g.ny);
	const unsigned mask = __ballot_sync(0xffffffffu, active);
	double acc[2] = {0.0, 0.0};
	if (active) {
		const bool pair = (i0 + 1 < g.nx);
		const bool edge_lo = (lane == 0);
		const bool edge_hi = (lane == 31) || (((mask >> (lane + 1)) & 1u) == 0u);
		const long long voffs = g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + i0;
		const double *xp = x + voffs;
		const double *ap = A + (long long)k * g.cplane + (long long)j * g.cpitch + i0;
		const long long cs = g.cstride;
		const double2 a0 = hints ? ld_stream2(ap) : ld2(ap);
		const double2 a1 = hints ? ld_stream2(ap + cs) : ld2(ap + cs);
		const double2 a2 = hints ? ld_stream2(ap + 2 * cs) : ld2(ap + 2 * cs);
		const double2 a3 = hints ? ld_stream2(ap + 3 * cs) : ld2(ap + 3 * cs);
		const double2 a4 = hints ? ld_stream2(ap + 4 * cs) : ld2(ap + 4 * cs);
		const double2 a5 = hints ? ld_stream2(ap + 5 * cs) : ld2(ap + 5 * cs);
		const double2 a6 = hints ? ld_stream2(ap + 6 * cs) : ld2(ap + 6 * cs);
		const double2 xc = ld2(xp);
		const double2 xjm = ld2(xp - g.vpitch);
		const double2 xjp = ld2(xp + g.vpitch);
		const double2 xkm = ld2(xp - g.vplane);
		const double2 xkp = ld2(xp + g.vplane);
		double2 bv = make_double2(0.0, 0.0);
		if (RES) bv = ld_stream2(b + voffs);
		double xl = __shfl_up_sync(mask, xc.y, 1);
		double xr = __shfl_down_sync(mask, xc.x, 1);
		if (edge_lo) xl = xp[-1];
		if (edge_hi) xr = xp[2];
		double y0 = stencil7(a0.x, a1.x, a2.x, a3.x, a4.x, a5.x, a6.x, xl, xjm.x, xkm.x, xc.x, xc.y, xjp.x, xkp.x);
		double y1 = stencil7(a0.y, a1.y, a2.y, a3.y, a4.y, a5.y, a6.y, xc.x, xjm.y, xkm.y, xc.y, xr, xjp.y, xkp.y);
		if (RES) { y0 = __dsub_rn(bv.x, y0); y1 = __dsub_rn(bv.y, y1); }
		if (pair) { if (hints) st_stream2(y + voffs, make_double2(y0, y1)); else st2(y + voffs, make_double2(y0, y1)); }
		else y[voffs] = y0;
		if (DOTW) {
			const double2 wv = (w == x) ? xc : ld2(w + voffs);
			acc[0] = wv.x * y0;
			if (pair) acc[0] += wv.y * y1;
		}
		if (DOTYY) {
			acc[1] = y0 * y0;
			if (pair) acc[1] += y1 * y1;
		}
	}
	if (DOTW || DOTYY) grid_reduce<2, true>(acc, ro);
}

// The first correct version: one thread per point, all neighbours straight from global memory.
template <bool RES, bool DOTW, bool DOTYY>
__global__ void __launch_bounds__(256)
spmv_naive_kernel(const GridDev g, const double *__restrict__ A, const double *__restrict__ x, double *__restrict__ y,
                  const double *__restrict__ b, const double *__restrict__ w, const RedOut ro, const int *stop, const int hints)
{
	if (s3d_stopped(stop)) return;
	const int i = blockIdx.x * 64 + threadIdx.x;
	const int j = blockIdx.y * 4 + threadIdx.y;
	const int k = blockIdx.z;
	double acc[2] = {0.0, 0.0};
	if (i < g.nx && j < g.ny) {
		const long long v = g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + i;
		const long long c = (long long)k * g.cplane + (long long)j * g.cpitch + i;
		double a[7];
#pragma unroll
		for (int q = 0; q < 7; q++) a[q] = hints ? ld_stream(A + c + q * g.cstride) : A[c + q * g.cstride];
		double s = stencil7(a[0], a[1], a[2], a[3], a[4], a[5], a[6],
		                    x[v - 1], x[v - g.vpitch], x[v - g.vplane], x[v], x[v + 1], x[v + g.vpitch], x[v + g.vplane]);
		if (RES) s = __dsub_rn(ld_stream(b + v), s);
		if (hints) __stcs(y + v, s); else y[v] = s;
		if (DOTW) acc[0] = w[v] * s;
		if (DOTYY) acc[1] = s * s;
	}
	if (DOTW || DOTYY) grid_reduce<2, true>(acc, ro);
}

template <int BY, int MINB>
int launch_march(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *x, double *y, const double *b,
                 const double *w, bool yy, double *d_red, const int *stop, int kchunk, const s3d_halo_view *hv = nullptr)
{
	const GridDev gd = to_dev(g);
	dim3 block(32, BY);
	dim3 grid((g->nx + 63) / 64, (g->ny + BY - 1) / BY, (g->nz + kchunk - 1) / kchunk);
	const bool red = (w != nullptr) || yy;
	const size_t nblk = (size_t)grid.x * grid.y * grid.z;
	const size_t first = ctx->red_defer ? ctx->red_blocks : 0;   // deferred mode: append behind earlier launches' partials
	if (red) {
		const int rc = s3d_ensure_partials(ctx, (first + nblk) * 2 + 256);   // + room for s3d_final_reduce's second level
		if (rc != S3D_OK) return rc;
	}
	RedOut ro{ctx->partials + first * 2, nullptr, d_red};
#define GO(RES, DW, DY)                                                                                                        \
	do {                                                                                                                       \
		if (MINB == 3 && hv) spmv_march_kernel<BY, 3, RES, DW, DY, true><<<grid, block, 0, ctx->stream>>>(gd, A, x, y, b, w, ro, stop, kchunk, *hv); \
		else spmv_march_kernel<BY, MINB, RES, DW, DY><<<grid, block, 0, ctx->stream>>>(gd, A, x, y, b, w, ro, stop, kchunk);      \
	} while (0)
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
	if (red && ctx->red_defer) { ctx->red_blocks += nblk; return S3D_OK; }
	if (red) return s3d_final_reduce(ctx, 2, (unsigned)nblk, d_red);
	return S3D_OK;
}

int launch_naive(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *x, double *y, const double *b,
                 const double *w, bool yy, double *d_red, const int *stop)
{
	const GridDev gd = to_dev(g);
	dim3 block(64, 4);
	dim3 grid((g->nx + 63) / 64, (g->ny + 3) / 4, g->nz);
	const bool red = (w != nullptr) || yy;
	const size_t nblk = (size_t)grid.x * grid.y * grid.z;
	const size_t first = ctx->red_defer ? ctx->red_blocks : 0;   // deferred mode: append behind earlier launches' partials
	if (red) {
		const int rc = s3d_ensure_partials(ctx, (first + nblk) * 2 + 256);   // + room for s3d_final_reduce's second level
		if (rc != S3D_OK) return rc;
	}
	RedOut ro{ctx->partials + first * 2, nullptr, d_red};   // very many blocks: partials only, summed by s3d_final_reduce
#define GO(RES, DW, DY) spmv_naive_kernel<RES, DW, DY><<<grid, block, 0, ctx->stream>>>(gd, A, x, y, b, w, ro, stop, g_tune.hints)
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
	if (red && ctx->red_defer) { ctx->red_blocks += nblk; return S3D_OK; }
	if (red) return s3d_final_reduce(ctx, 2, (unsigned)nblk, d_red);
	return S3D_OK;
}


int launch_pair(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *x, double *y, const double *b,
                const double *w, bool yy, double *d_red, const int *stop)
{
	const GridDev gd = to_dev(g);
	dim3 block(32, 8);
	dim3 grid((g->nx + 63) / 64, (g->ny + 7) / 8, g->nz);
	const bool red = (w != nullptr) || yy;
	if (red) {
		const int rc = s3d_ensure_partials(ctx, (size_t)grid.x * grid.y * grid.z * 2 + 256);   // + room for s3d_final_reduce's second level
		if (rc != S3D_OK) return rc;
	}
	RedOut ro{ctx->partials, nullptr, d_red};   // very many blocks: partials only, summed by s3d_final_reduce
#define GO(RES, DW, DY) spmv_pair_kernel<RES, DW, DY><<<grid, block, 0, ctx->stream>>>(gd, A, x, y, b, w, ro, stop, g_tune.hints)
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
	if (red) return s3d_final_reduce(ctx, 2, grid.x * grid.y * grid.z, d_red);
	return S3D_OK;
}

// z-chunk length and occupancy, from the in-run sweep on B200 (profiles/sweep_r1e.md): the SHORTEST march wins
// (2 planes: 7074 GB/s at 512^3, 7106 at 1024^3; 8 planes 6993/6800; 32 planes 6640) because parallelism and
// bytes in flight matter more than saving the k-neighbour loads, which hit L2 anyway; compiling for 3 blocks/SM
// (<= 80 registers, all loads of a plane in flight) beats 6 blocks/SM by ~0.3%.  The fused-reduction forms use
// 4 planes so the per-block reduction is amortised over 8 points per thread.
int pick_kchunk(const s3d_ctx *, const s3d_grid *g, bool fused)
{
	if (g_tune.kchunk > 0) return min(g_tune.kchunk, g->nz);
	return min(fused ? 4 : 2, g->nz);
}

}  // namespace

int s3d_spmv_tma(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *x, double *y, const double *b,
                 const double *w, bool yy, double *d_red, const int *stop);

extern "C" S3D_HIDDEN void s3d_internal_set_ew_unroll(int u);
extern "C" S3D_HIDDEN int s3d_internal_halo_push(s3d_ctx *ctx, const s3d_grid *g, const double *d_x, s3d_halo_view *hv);
extern "C" S3D_HIDDEN int s3d_internal_halo_join(s3d_ctx *ctx);

extern "C" {

int s3d_tune_set(s3d_ctx *ctx, const char *key, int value)
{
	if (!key) return S3D_ERR_ARG;
	if (!strcmp(key, "spmv_minb")) {
		S3D_REQUIRE(ctx, value == 0 || value == 3 || value == 4 || value == 6, "spmv_minb must be 0 (auto), 3, 4 or 6");
		g_tune.minb = value;
	} else if (!strcmp(key, "spmv_kchunk")) g_tune.kchunk = value;
	else if (!strcmp(key, "spmv_hints")) g_tune.hints = value;
	else if (!strcmp(key, "sgs_sleep_ns")) ctx->sgs_sleep_ns = value;
	else if (!strcmp(key, "sgs_variant")) ctx->sgs_variant = value;
	else if (!strcmp(key, "sgs_warps_per_sm")) ctx->sgs_warps_per_sm = value;
	else if (!strcmp(key, "sgs_pencil_shape")) ctx->sgs_pencil_shape = value;
	else if (!strcmp(key, "sgs_async_chunk")) ctx->sgs_async_chunk = value;
	else if (!strcmp(key, "sgs_tile_mail")) ctx->sgs_tile_mail = value;
	else if (!strcmp(key, "sgs_blocks")) ctx->sgs_blocks = value;
	else if (!strcmp(key, "sgs_dbuf")) ctx->sgs_dbuf = value;
	else if (!strcmp(key, "sgs_blocks_y")) ctx->sgs_blocks_y = value;
	else if (!strcmp(key, "small_team")) ctx->small_team = value;
	else if (!strcmp(key, "ew_unroll")) s3d_internal_set_ew_unroll(value);
	else if (!strcmp(key, "sgs_stream_max_points")) ctx->sgs_stream_max_points = (long long)value;
	else if (!strcmp(key, "sgs_tile_i")) ctx->sgs_tile_i = value;
	else if (!strcmp(key, "sgs_pencils")) ctx->sgs_pencils = value;
	else if (!strcmp(key, "sgs_profile")) {
		// 1: start accumulating per-phase cycles in the warp dataflow kernel; 0: print and stop (timing experiments only)
		if (value && !ctx->sgs_prof) {
			S3D_CUDA(ctx, cudaMalloc(&ctx->sgs_prof, (8 + 4096) * sizeof(long long)));
			S3D_CUDA(ctx, cudaMemset(ctx->sgs_prof, 0, (8 + 4096) * sizeof(long long)));
		} else if (!value && ctx->sgs_prof) {
			static long long h[8 + 4096];
			S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
			S3D_CUDA(ctx, cudaMemcpy(h, ctx->sgs_prof, sizeof h, cudaMemcpyDeviceToHost));
			if (getenv("S3D_SGS_TILELOG")) {   // warp-per-tile kernel with mailboxes: begin / data seen / sweep start / sweep end / published, ns
				for (int t = 0; t < 800; t++)
					if (h[8 + 5 * t]) fprintf(stderr, "[sgs_tile] %d %lld %lld %lld %lld %lld\n", t, h[8 + 5 * t], h[9 + 5 * t], h[10 + 5 * t], h[11 + 5 * t], h[12 + 5 * t]);
			}
			if (getenv("S3D_SGS_LINELOG")) {
				long long t0 = 0;
				for (int l = 0; l < 2048; l++) if (h[8 + 2 * l] && (!t0 || h[8 + 2 * l] < t0)) t0 = h[8 + 2 * l];
				for (int l = 0; l < 2048; l++) if (h[8 + 2 * l]) fprintf(stderr, "[sgs_line] %d start %.2f us end %.2f us\n", l, (h[8 + 2 * l] - t0) * 1e-3, (h[9 + 2 * l] - t0) * 1e-3);
			}
			const double t = h[4] ? (double)h[4] : 1.0;
			fprintf(stderr, "[sgs_profile] workers %lld  cycles/worker: wait %.0f  total %.0f (%.1f per step, %.1f without the waits)  helper: load %.0f face %.0f retire %.0f passes %.0f\n", h[4], h[0] / t,
			        h[2] / t, h[5] ? (double)h[2] / h[5] : 0.0, h[5] ? (double)(h[2] - h[0]) / h[5] : 0.0, h[1] / t, h[3] / t, h[6] / t, h[7] / t);
			S3D_CUDA(ctx, cudaFree(ctx->sgs_prof));
			ctx->sgs_prof = nullptr;
		}
	}
	else if (!strcmp(key, "halo_overlap")) ctx->no_overlap = (value == 0);
	else if (!strcmp(key, "halo_p2p")) ctx->p2p.mode = value ? 1 : 0;
	else if (!strcmp(key, "halo_push_side")) ctx->push_side = value ? 1 : 0;
	else if (!strcmp(key, "allreduce_p2p")) ctx->arwin.mode = value ? 1 : 0;
	else if (!strcmp(key, "spmv_variant")) {
		S3D_REQUIRE(ctx, value >= S3D_SPMV_MARCH && value <= S3D_SPMV_PAIR, "unknown SpMV variant");
		g_tune.variant = value;
	}
	else return s3d_fail(ctx, S3D_ERR_ARG, "unknown tuning key %s", key);
	return S3D_OK;
}

int s3d_spmv(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *d_x, double *d_y, const double *d_b,
             const double *d_w, int want_yy, double *d_red, const int *d_stop, int variant)
{
	S3D_REQUIRE(ctx, g && A && d_x && d_y, "null argument");
	S3D_REQUIRE(ctx, d_x != d_y, "x and y must not alias");
	S3D_REQUIRE(ctx, !(d_w || want_yy) || d_red, "fused reductions need d_red");
	if (variant == S3D_SPMV_AUTO) variant = g_tune.variant;
	if (!d_stop) d_stop = ctx->gate;
	switch (variant) {
	case S3D_SPMV_MARCH: {
		const bool yy = want_yy != 0;
		const int kc = pick_kchunk(ctx, g, d_w != nullptr || yy);
		// resident blocks per SM the kernel is compiled for: 6 (<= 40 registers), 4 (<= 64) or 3 (<= 80)
		int minb = g_tune.minb;
		if (minb == 0) minb = 3;
		if (minb >= 6) return launch_march<8, 6>(ctx, g, A, d_x, d_y, d_b, d_w, yy, d_red, d_stop, kc);
		if (minb >= 4) return launch_march<8, 4>(ctx, g, A, d_x, d_y, d_b, d_w, yy, d_red, d_stop, kc);
		return launch_march<8, 3>(ctx, g, A, d_x, d_y, d_b, d_w, yy, d_red, d_stop, kc);
	}
	case S3D_SPMV_PAIR:
		return launch_pair(ctx, g, A, d_x, d_y, d_b, d_w, want_yy != 0, d_red, d_stop);
	case S3D_SPMV_NAIVE:
		return launch_naive(ctx, g, A, d_x, d_y, d_b, d_w, want_yy != 0, d_red, d_stop);
	case S3D_SPMV_TMA:
		return s3d_spmv_tma(ctx, g, A, d_x, d_y, d_b, d_w, want_yy != 0, d_red, d_stop);
	default:
		return s3d_fail(ctx, S3D_ERR_ARG, "unknown SpMV variant %d", variant);
	}
}

// SMat::apply / apply_res on a z-slab: the halo exchange and the multiply in one call, overlapped.
//   comm stream : NCCL send/recv of the two boundary planes of x (starts when x is ready on the compute stream)
//   compute     : the interior planes [1, nz-1), which need no halo, run meanwhile; then, once the halos have landed,
//                 the two boundary planes.  The fused reductions of the three launches share one final sum.
// One rank (or a slab too thin to split): plain s3d_spmv.  Only the MARCH variant is split.
int s3d_spmv_exchange(s3d_ctx *ctx, const s3d_grid *g, const double *A, double *d_x, double *d_y, const double *d_b,
                      const double *d_w, int want_yy, double *d_red, const int *d_stop)
{
	S3D_REQUIRE(ctx, g && A && d_x && d_y, "null argument");
	if (g->nranks == 1) return s3d_spmv(ctx, g, A, d_x, d_y, d_b, d_w, want_yy, d_red, d_stop, S3D_SPMV_AUTO);
	if (ctx->no_overlap && g_tune.variant == S3D_SPMV_MARCH && (g_tune.minb == 0 || g_tune.minb == 3) && ctx->comm) {
		// peer-memory exchange fused into the multiply: push my planes, then a SpMV that reads the neighbours' planes from the
		// window and whose two boundary chunks (dispatched last) are the only blocks that wait for them
		s3d_halo_view hv;
		const int rc = s3d_internal_halo_push(ctx, g, d_x, &hv);
		if (rc != S3D_OK) return rc;
		if (hv.epoch) {
			S3D_REQUIRE(ctx, d_x != d_y, "x and y must not alias");
			const bool yy = want_yy != 0;
			S3D_REQUIRE(ctx, !(d_w || yy) || d_red, "fused reductions need d_red");
			const int kc = pick_kchunk(ctx, g, d_w != nullptr || yy);
			const int rcm = launch_march<8, 3>(ctx, g, A, d_x, d_y, d_b, d_w, yy, d_red, d_stop ? d_stop : ctx->gate, kc, &hv);
			if (rcm != S3D_OK) return rcm;
			return s3d_internal_halo_join(ctx);   // the push ran beside the multiply on its own stream
		}
		// no peer memory: the exchange below
	}
	if (g->nz < 3 || g_tune.variant != S3D_SPMV_MARCH || ctx->no_overlap) {
		const int rc = s3d_halo_exchange(ctx, g, d_x);
		if (rc != S3D_OK) return rc;
		return s3d_spmv(ctx, g, A, d_x, d_y, d_b, d_w, want_yy, d_red, d_stop, S3D_SPMV_AUTO);
	}
	if (!ctx->comm_stream) {
		S3D_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
		S3D_CUDA(ctx, cudaEventCreateWithFlags(&ctx->comm_ev[0], cudaEventDisableTiming));
		S3D_CUDA(ctx, cudaEventCreateWithFlags(&ctx->comm_ev[1], cudaEventDisableTiming));
	}
	// 1. halo exchange on the comm stream, after everything already enqueued on the compute stream (x is final there)
	S3D_CUDA(ctx, cudaEventRecord(ctx->comm_ev[0], ctx->stream));
	S3D_CUDA(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->comm_ev[0], 0));
	cudaStream_t compute = ctx->stream;
	ctx->stream = ctx->comm_stream;
	int rc = s3d_halo_exchange(ctx, g, d_x);
	ctx->stream = compute;
	if (rc != S3D_OK) return rc;
	S3D_CUDA(ctx, cudaEventRecord(ctx->comm_ev[1], ctx->comm_stream));
	// 2. interior planes meanwhile; 3. boundary planes after the halos
	const bool red = (d_w != nullptr) || want_yy;
	if (red) {
		// size the partials buffer for all three launches up front: a reallocation between them would lose partials
		const int kc = pick_kchunk(ctx, g, true);
		const size_t tiles = (size_t)((g->nx + 63) / 64) * ((g->ny + 7) / 8);
		rc = s3d_ensure_partials(ctx, tiles * ((g->nz + kc - 1) / kc + 3) * 2 + 256);
		if (rc != S3D_OK) return rc;
	}
	ctx->red_defer = red;
	ctx->red_blocks = 0;
	auto part = [&](int k0, int nzp) -> int {
		s3d_grid sub = *g;
		sub.nz = nzp;
		const size_t vo = (size_t)k0 * g->vplane, co = (size_t)k0 * g->cplane;
		return s3d_spmv(ctx, &sub, A + co, d_x + vo, d_y + vo, d_b ? d_b + vo : nullptr, d_w ? d_w + vo : nullptr, want_yy, d_red, d_stop,
		                S3D_SPMV_MARCH);
	};
	rc = part(1, g->nz - 2);
	if (rc == S3D_OK) {
		cudaError_t e = cudaStreamWaitEvent(ctx->stream, ctx->comm_ev[1], 0);
		if (e != cudaSuccess) { ctx->red_defer = false; S3D_CUDA(ctx, e); }
		rc = part(0, 1);
	}
	if (rc == S3D_OK) rc = part(g->nz - 1, 1);
	ctx->red_defer = false;
	if (rc != S3D_OK) return rc;
	if (red) return s3d_final_reduce(ctx, 2, (unsigned)ctx->red_blocks, d_red);
	return S3D_OK;
}

// y = A x (or b - A x with a device-resident b) for HOST vectors, pipelined in z-chunks: chunk c+1 of x is on its way up
// (H2D stream) while chunk c is multiplied (compute stream) and chunk c-1 of y goes down (D2H stream), so the two PCIe
// directions run concurrently and the kernel hides behind them.  h_x / h_y: host arrays in the SVec layout (pinned memory for
// full speed); d_x / d_y: device vectors used as staging (contents are overwritten; d_y's ghost planes must be zero, as
// after s3d_vec_alloc).  Blocks until h_y is complete.
// On a z-slab (g->nranks > 1, collective over the ranks) every rank passes its own slab: its first and last real plane of x go up
// first and are pushed into the neighbours' peer-memory windows, the interior chunks follow as on one GPU, and the two chunks that
// touch a neighbour's plane are multiplied last, reading it from the window (they are the only blocks that may wait).  The host
// ghost planes of x at interior cuts are ignored (the neighbour owns those values); without peer memory the planes travel by NCCL.
int s3d_spmv_host(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *h_x, double *h_y, double *d_x, double *d_y,
                  const double *d_b, int nchunks)
{
	S3D_REQUIRE(ctx, g && A && h_x && h_y && d_x && d_y, "null argument");
	const bool slab = g->nranks > 1;
	S3D_REQUIRE(ctx, !slab || ctx->comm, "a z-slab needs s3d_comm_init");
	if (nchunks <= 0) nchunks = 16;
	if (nchunks > g->nz) nchunks = g->nz;
	if (slab && nchunks < 3 && g->nz >= 3) nchunks = 3;
	if (!ctx->up_stream) {
		S3D_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
		S3D_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking));
	}
	const int nev = 2 * nchunks + 2;
	if (ctx->chunk_ev_cap < nev) {
		cudaEvent_t *ne = (cudaEvent_t *)realloc(ctx->chunk_ev, nev * sizeof(cudaEvent_t));
		if (!ne) return s3d_fail(ctx, S3D_ERR_ARG, "out of host memory");
		ctx->chunk_ev = ne;
		for (int i = ctx->chunk_ev_cap; i < nev; i++) {
			S3D_CUDA(ctx, cudaEventCreateWithFlags(&ne[i], cudaEventDisableTiming));
			ctx->chunk_ev_cap = i + 1;
		}
	}
	const size_t hrow = (size_t)(g->nx + 2) * sizeof(double), drow = (size_t)g->vpitch * sizeof(double);
	const size_t rows_per_plane = (size_t)(g->ny + 2);
	const size_t hplane = (size_t)(g->nx + 2) * rows_per_plane;     // doubles per host plane
	const size_t glen = hplane * (g->nz + 2 + (slab ? 2 : 0));
	if (ctx->stage_cap < glen) {
		S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		if (ctx->stage) S3D_CUDA(ctx, cudaFree(ctx->stage));
		ctx->stage = nullptr; ctx->stage_cap = 0;
		S3D_CUDA(ctx, cudaMalloc(&ctx->stage, glen * sizeof(double)));
		ctx->stage_cap = glen;
	}
	if (ctx->stage_y_cap < glen) {
		S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		if (ctx->stage_y) S3D_CUDA(ctx, cudaFree(ctx->stage_y));
		ctx->stage_y = nullptr; ctx->stage_y_cap = 0;
		S3D_CUDA(ctx, cudaMalloc(&ctx->stage_y, glen * sizeof(double)));
		ctx->stage_y_cap = glen;
	}
	// the copy streams must not start before earlier work on the compute stream (previous users of d_x / d_y) is done
	cudaEvent_t start = ctx->chunk_ev[2 * nchunks];
	S3D_CUDA(ctx, cudaEventRecord(start, ctx->stream));
	S3D_CUDA(ctx, cudaStreamWaitEvent(ctx->up_stream, start, 0));
	S3D_CUDA(ctx, cudaStreamWaitEvent(ctx->down_stream, start, 0));
	// repack ghosted planes [from, to) of the staging buffer into the padded device layout (compute stream)
	auto repack = [&](const double *src, int from, int to) -> cudaError_t {
		return cudaMemcpy2DAsync(d_x + (size_t)from * g->vplane + g->vlead - 1, drow, src, hrow, hrow, rows_per_plane * (to - from),
		                         cudaMemcpyDeviceToDevice, ctx->stream);
	};
	s3d_halo_view hv;
	const bool has_lo = slab && g->rank > 0, has_hi = slab && g->rank + 1 < g->nranks;
	if (slab) {
		// my two boundary planes first: the neighbours need them before anything else
		double *bst = ctx->stage + hplane * (g->nz + 2);
		S3D_CUDA(ctx, cudaMemcpyAsync(bst, h_x + hplane, hplane * sizeof(double), cudaMemcpyHostToDevice, ctx->up_stream));
		S3D_CUDA(ctx, cudaMemcpyAsync(bst + hplane, h_x + hplane * g->nz, hplane * sizeof(double), cudaMemcpyHostToDevice, ctx->up_stream));
		cudaEvent_t bev = ctx->chunk_ev[2 * nchunks + 1];
		S3D_CUDA(ctx, cudaEventRecord(bev, ctx->up_stream));
		S3D_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, bev, 0));
		S3D_CUDA(ctx, repack(bst, 1, 2));
		S3D_CUDA(ctx, repack(bst + hplane, g->nz, g->nz + 1));
		int rc = s3d_internal_halo_push(ctx, g, d_x, &hv);
		if (rc != S3D_OK) return rc;
		rc = s3d_internal_halo_join(ctx);   // the repacks below rewrite the boundary planes (with the same values): no overlap here
		if (rc != S3D_OK) return rc;
		if (!hv.epoch) {
			rc = s3d_halo_exchange(ctx, g, d_x);      // NCCL: fills the ghost planes of d_x, which the repacks below leave alone
			if (rc != S3D_OK) return rc;
		}
	}
	auto multiply = [&](int c, int k0, int k1) -> int {
		s3d_grid sub = *g;
		sub.nz = k1 - k0;
		const size_t voff = (size_t)k0 * g->vplane, coff = (size_t)k0 * g->cplane;
		int rc;
		if (hv.epoch && ((k0 == 0 && hv.lo) || (k1 == g->nz && hv.hi))) {
			s3d_halo_view h = hv;
			if (k0 != 0) { h.lo = nullptr; h.flag_lo = nullptr; }
			if (k1 != g->nz) { h.hi = nullptr; h.flag_hi = nullptr; }
			rc = launch_march<8, 3>(ctx, &sub, A + coff, d_x + voff, d_y + voff, d_b ? d_b + voff : nullptr, nullptr, false, nullptr, ctx->gate,
			                        pick_kchunk(ctx, &sub, false), &h);
		} else {
			rc = s3d_spmv(ctx, &sub, A + coff, d_x + voff, d_y + voff, d_b ? d_b + voff : nullptr, nullptr, 0, nullptr, nullptr, S3D_SPMV_AUTO);
		}
		if (rc != S3D_OK) return rc;
		// y: the chunk's real planes, plus the bottom / top ghost plane with the first / last chunk (ghosted planes [lo, hi)):
		// packed into the host layout on the device, then ONE contiguous D2H copy (pitched copies run slower while the other PCIe
		// direction is busy)
		const int lo = (c == 0) ? 0 : k0 + 1, hi = (c == nchunks - 1) ? g->nz + 2 : k1 + 1;
		S3D_CUDA(ctx, cudaMemcpy2DAsync(ctx->stage_y + (size_t)lo * hplane, hrow, d_y + (size_t)lo * g->vplane + g->vlead - 1, drow, hrow,
		                                rows_per_plane * (hi - lo), cudaMemcpyDeviceToDevice, ctx->stream));
		S3D_CUDA(ctx, cudaEventRecord(ctx->chunk_ev[2 * c + 1], ctx->stream));
		S3D_CUDA(ctx, cudaStreamWaitEvent(ctx->down_stream, ctx->chunk_ev[2 * c + 1], 0));
		S3D_CUDA(ctx, cudaMemcpyAsync(h_y + (size_t)lo * hplane, ctx->stage_y + (size_t)lo * hplane, (size_t)(hi - lo) * hplane * sizeof(double),
		                              cudaMemcpyDeviceToHost, ctx->down_stream));
		return S3D_OK;
	};
	const bool defer_first = hv.epoch && has_lo && nchunks > 1, defer_last = hv.epoch && has_hi && nchunks > 1;
	int up_to = 0;   // ghosted planes [0, up_to) of x are enqueued for upload
	for (int c = 0; c < nchunks; c++) {
		const int k0 = (int)((long long)c * g->nz / nchunks), k1 = (int)((long long)(c + 1) * g->nz / nchunks);   // real planes [k0, k1)
		// x planes needed: ghosted [k0, k1 + 2)
		const int need = k1 + 2;
		const int from = up_to;
		if (need > up_to) {
			// contiguous H2D copy at full PCIe rate (a pitched 2-D upload runs at ~31 GB/s instead of ~55 GB/s) ...
			S3D_CUDA(ctx, cudaMemcpyAsync(ctx->stage + (size_t)up_to * hplane, h_x + (size_t)up_to * hplane,
			                              (size_t)(need - up_to) * hplane * sizeof(double), cudaMemcpyHostToDevice, ctx->up_stream));
			up_to = need;
		}
		S3D_CUDA(ctx, cudaEventRecord(ctx->chunk_ev[2 * c], ctx->up_stream));
		S3D_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->chunk_ev[2 * c], 0));
		// ... then repacked into the padded row layout on the device, at HBM speed (ghost planes at interior cuts stay as they are)
		if (up_to > from) {
			const int rf = (has_lo && from == 0) ? 1 : from, rt = (has_hi && up_to == g->nz + 2) ? g->nz + 1 : up_to;
			if (rt > rf) S3D_CUDA(ctx, repack(ctx->stage + (size_t)rf * hplane, rf, rt));
		}
		if ((c == 0 && defer_first) || (c == nchunks - 1 && defer_last)) continue;
		const int rc = multiply(c, k0, k1);
		if (rc != S3D_OK) return rc;
	}
	if (defer_first) {
		const int rc = multiply(0, 0, (int)((long long)g->nz / nchunks));
		if (rc != S3D_OK) return rc;
	}
	if (defer_last) {
		const int rc = multiply(nchunks - 1, (int)((long long)(nchunks - 1) * g->nz / nchunks), g->nz);
		if (rc != S3D_OK) return rc;
	}
	S3D_CUDA(ctx, cudaStreamSynchronize(ctx->down_stream));
	S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	return S3D_OK;
}

}  // extern "C"
