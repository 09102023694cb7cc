// ASYNCHRONOUS (chaotic) multi-sweep application of (D+L) D^-1 (D+U): the reference's threaded apply on the GPU
// (SGS_like_preconditioner::apply with threadedapply = true, src/linalg/s3d_sgspreconditioners.cpp:30-112).
//
// The reference opens ONE parallel region and lets every thread run `napplysweeps` sweeps over ITS rows without ever waiting for the
// other threads (`#pragma omp for collapse(2) nowait schedule(static, chunk)`, cpp:53 and :86): a thread reads whatever its neighbours
// have written so far, values of this sweep, of an earlier one, or the initial zero.  The same structure here, with warps in the role of
// the threads:
//   * the (k, j) rows are numbered k ny + j and cut into chunks of `chunk` consecutive rows (-s3d_thread_chunk_size); chunk q belongs
//     to warp q mod W of the W persistent warps of the launch, and every warp walks its chunks in sweep order, sweep after sweep, with
//     no barrier anywhere -- one kernel launch per triangular factor;
//   * inside a row the i-recurrence is EXACT, as it is for the reference's thread that walks the row: the row is taken in segments of 32
//     columns (coalesced 256-byte loads of the five inputs and the two neighbour rows), y_i = a_i + b_i y_{i-1} is solved for a segment
//     by an affine prefix scan over the lanes (five shuffle steps), the carry goes on to the next segment;
//   * the j- and k-neighbour rows are read from L2 (`ld.global.cg`) and the results stored there (`st.global.cg`), so a value is
//     visible to the other warps as soon as it is written, never held back in an L1.
// Like the reference's threaded mode the result depends on timing (it is not reproducible from run to run) and approaches the exact
// lexicographic result as sweeps are added: any order of updates of a triangular system reaches its fixed point after at most
// nx + ny + nz sweeps.  The scan adds the terms of the i-recurrence in another order than a sequential walk, so even the fixed point
// agrees with the lexicographic kernels to rounding only (tests/test_kernels_gpu.py: 1e-12).  Not available across z-slabs other than
// slab by slab (what lies in the neighbouring slabs is taken as zero, as with S3D_SGS_SLABLOCAL).
#include <algorithm>
#include "common.cuh"

namespace {

__device__ __forceinline__ double ldcg1(const double *p) { return __ldcg(p); }
__device__ __forceinline__ void stcg1(double *p, double v) { __stcg(p, v); }

// BWD = false: y = (r - A0 y[i-1] - A1 y[j-1] - A2 y[k-1]) dinv, rows upwards, columns upwards
// BWD = true : z = y - dinv (A4 z[i+1] + A5 z[j+1] + A6 z[k+1]), rows downwards, columns downwards
template <bool BWD>
__global__ void __launch_bounds__(256)
sgs_async_kernel(const GridDev g, const double *__restrict__ c0, const double *__restrict__ c1, const double *__restrict__ c2,
                 const double *__restrict__ dinv, const double *__restrict__ rhs, double *out, int nsweeps, int chunk, int kzero_lo,
                 int kzero_hi, const int *stop)
{
	if (s3d_stopped(stop)) return;
	constexpr unsigned FULL = 0xffffffffu;
	const int lane = threadIdx.x & 31;
	const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nwarps = (int)((gridDim.x * blockDim.x) >> 5);
	const int nrows = g.ny * g.nz;
	const int nchunks = (nrows + chunk - 1) / chunk;
	const int nseg = (g.nx + 31) / 32;
	for (int swp = 0; swp < nsweeps; swp++) {
		for (int q = warp; q < nchunks; q += nwarps) {
			const int qq = BWD ? nchunks - 1 - q : q;
			const int r0 = qq * chunk, r1 = min(nrows, r0 + chunk);
			for (int rr = 0; rr < r1 - r0; rr++) {
				const int row = BWD ? r1 - 1 - rr : r0 + rr;
				const int k = row / g.ny, j = row - k * g.ny;
				const long long vrow = g.voff + (long long)k * g.vplane + (long long)j * g.vpitch;
				const long long crow = (long long)k * g.cplane + (long long)j * g.cpitch;
				const long long dj = BWD ? g.vpitch : -(long long)g.vpitch, dk = BWD ? g.vplane : -g.vplane;
				// what lies in the neighbouring slab is taken as zero (the ghost plane may hold an exchanged iterate of something else)
				const bool kz = BWD ? (kzero_hi && k == g.nz - 1) : (kzero_lo && k == 0);
				double carry = ldcg1(out + vrow + (BWD ? g.nx : -1));           // the ghost column
				for (int s = 0; s < nseg; s++) {
					const int i = BWD ? (nseg - 1 - s) * 32 + (31 - lane) : s * 32 + lane;   // lane 0 comes first in sweep order
					const bool in = i < g.nx;
					double a = 0.0, b = 1.0;                                  // columns outside the row: the identity, the carry passes through
					if (in) {
						const double nj = ldcg1(out + vrow + dj + i), nk = kz ? 0.0 : ldcg1(out + vrow + dk + i);
						const double x0 = ld_stream(c0 + crow + i), x1 = ld_stream(c1 + crow + i), x2 = ld_stream(c2 + crow + i);
						const double dv = dinv[crow + i], f = rhs[vrow + i];
						if (!BWD) {
							a = __dmul_rn(__dsub_rn(__dsub_rn(f, __dmul_rn(x1, nj)), __dmul_rn(x2, nk)), dv);
							b = -__dmul_rn(x0, dv);
						} else {
							a = __dsub_rn(f, __dmul_rn(dv, __dadd_rn(__dmul_rn(x1, nj), __dmul_rn(x2, nk))));
							b = -__dmul_rn(dv, x0);
						}
					}
					// inclusive scan of the affine maps v -> a + b v over the lanes: afterwards value(lane) = a + b * carry
#pragma unroll
					for (int d = 1; d < 32; d <<= 1) {
						const double ap = __shfl_up_sync(FULL, a, d), bp = __shfl_up_sync(FULL, b, d);
						if (lane >= d) { a = fma(b, ap, a); b = b * bp; }
					}
					const double v = fma(b, carry, a);
					if (in) stcg1(out + vrow + i, v);
					// the carry: the value behind the last lane (lanes outside the row hand the value of the last column on)
					carry = __shfl_sync(FULL, v, 31);
				}
			}
		}
	}
}

// ASYNCHRONOUS build of the StrILU diagonal (StrILU_preconditioner::updateOperator with threadedbuild = true, src/linalg/s3d_ilu.cpp:17-64):
// d starts as the matrix diagonal and every thread of the reference re-evaluates d = A3 - P0 / d[i-1] - P1 / d[j-1] - P2 / d[k-1] over its chunks of
// rows, `nbuildsweeps` times, without ever waiting for the others (`omp for nowait schedule(dynamic, chunk)`).  Here a LANE walks a row (the
// recurrence along i is not linear: no scan), the rows of a chunk are dealt to the lanes of the warp the chunk belongs to, warps never wait;
// neighbour rows are read from and results written to L2.  P_s = A_s A_{s+4}[neighbour] are the exact build's products (zero where the neighbour
// does not exist), D is a vector-layout array whose ghosts are 1.  Set-up code: uncoalesced and latency-bound, run once per operator.
__global__ void __launch_bounds__(256)
strilu_async_kernel(const GridDev g, const double *__restrict__ a3, const double *__restrict__ P, double *D, int nsweeps, int chunk, const int *stop)
{
	if (s3d_stopped(stop)) return;
	const int lane = threadIdx.x & 31;
	const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nwarps = (int)((gridDim.x * blockDim.x) >> 5);
	const int nrows = g.ny * g.nz;
	const int nchunks = (nrows + chunk - 1) / chunk;
	const long long cs = g.cstride;
	for (int swp = 0; swp < nsweeps; swp++)
		for (int q = warp; q < nchunks; q += nwarps) {
			const int r0 = q * chunk, r1 = min(nrows, r0 + chunk);
			for (int row = r0 + lane; row < r1; row += 32) {
				const int k = row / g.ny, j = row - k * g.ny;
				const long long vrow = g.voff + (long long)k * g.vplane + (long long)j * g.vpitch;
				const long long crow = (long long)k * g.cplane + (long long)j * g.cpitch;
				double prev = 1.0;                                          // the ghost column (P0 is zero there)
				for (int i = 0; i < g.nx; i++) {
					const double dj = __ldcg(D + vrow - g.vpitch + i), dk = __ldcg(D + vrow - g.vplane + i);
					double v = __dsub_rn(a3[crow + i], __ddiv_rn(P[crow + i], prev));
					v = __dsub_rn(v, __ddiv_rn(P[cs + crow + i], dj));
					v = __dsub_rn(v, __ddiv_rn(P[2 * cs + crow + i], dk));
					__stcg(D + vrow + i, v);
					prev = v;
				}
			}
		}
}

__global__ void __launch_bounds__(256) strilu_init_kernel(const GridDev g, const double *__restrict__ a3, double *D)
{
	const long long total = (long long)g.nx * g.ny * g.nz;
	for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
		const int i = (int)(t % g.nx);
		const long long row = t / g.nx;
		const int j = (int)(row % g.ny), k = (int)(row / g.ny);
		D[g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + i] = a3[(long long)k * g.cplane + (long long)j * g.cpitch + i];
	}
}

__global__ void __launch_bounds__(256) zero_real_kernel(const GridDev g, double *x, const int *stop)
{
	if (s3d_stopped(stop)) return;
	const long long total = (long long)g.nx * g.ny * g.nz;
	for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
		const int i = (int)(t % g.nx);
		const long long row = t / g.nx;
		const int j = (int)(row % g.ny), k = (int)(row / g.ny);
		x[g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + i] = 0.0;
	}
}

}  // namespace

// D (vector layout, every entry 1 on entry: ghosts stay 1) <- the diagonal after nsweeps asynchronous sweeps; a3: the matrix diagonal,
// P: the three product arrays (coefficient layout, stride g->cstride).
extern "C" S3D_HIDDEN int s3d_internal_strilu_async(s3d_ctx *ctx, const s3d_grid *g, const double *a3, const double *P, double *D, int nsweeps)
{
	const GridDev gd = to_dev(g);
	const int nrows = g->ny * g->nz;
	int chunk = std::max(1, ctx->sgs_async_chunk);
	const int want_chunks = ctx->sm_count * 16;
	if (nrows >= 4 * want_chunks && nrows / chunk < want_chunks) chunk = std::max(1, nrows / want_chunks);
	const int nchunks = (nrows + chunk - 1) / chunk;
	const unsigned nb = (unsigned)std::max(1, std::min((nchunks + 7) / 8, ctx->sm_count * 8));
	strilu_init_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(gd, a3, D);
	S3D_LAUNCHED(ctx);
	strilu_async_kernel<<<nb, 256, 0, ctx->stream>>>(gd, a3, P, D, nsweeps, chunk, nullptr);
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}

// d_y: scratch vector (forward result), d_z: output.  Both start from zero, like the reference's y (cpp:25-28).
extern "C" S3D_HIDDEN int s3d_internal_sgs_async(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *d_dinv, const double *d_r,
                                                 double *d_z, double *d_y, int nsweeps)
{
	const GridDev gd = to_dev(g);
	const long long cs = g->cstride;
	const int nrows = g->ny * g->nz;
	// the reference's chunk is sized for 8-64 CPU threads (run-defaults.perc: 1024 rows): on grids that can feed the whole GPU a chunk
	// is at most what leaves ~16 warps per SM something to do; small grids keep the chunk they were given
	int chunk = std::max(1, ctx->sgs_async_chunk);
	const int want_chunks = ctx->sm_count * 16;
	if (nrows >= 4 * want_chunks && nrows / chunk < want_chunks) chunk = std::max(1, nrows / want_chunks);
	const int nchunks = (nrows + chunk - 1) / chunk;
	const unsigned nb = (unsigned)std::max(1, std::min((nchunks + 7) / 8, ctx->sm_count * 8));   // 8 warps per block, 64 warps per SM
	const int klo = g->nranks > 1 && g->rank > 0, khi = g->nranks > 1 && g->rank < g->nranks - 1;
	const unsigned zb = (unsigned)ctx->sm_count * 8;
	zero_real_kernel<<<zb, 256, 0, ctx->stream>>>(gd, d_y, ctx->gate);
	S3D_LAUNCHED(ctx);
	zero_real_kernel<<<zb, 256, 0, ctx->stream>>>(gd, d_z, ctx->gate);
	S3D_LAUNCHED(ctx);
	sgs_async_kernel<false><<<nb, 256, 0, ctx->stream>>>(gd, A, A + cs, A + 2 * cs, d_dinv, d_r, d_y, nsweeps, chunk, klo, khi, ctx->gate);
	S3D_LAUNCHED(ctx);
	sgs_async_kernel<true><<<nb, 256, 0, ctx->stream>>>(gd, A + 4 * cs, A + 5 * cs, A + 6 * cs, d_dinv, d_y, d_z, nsweeps, chunk, klo, khi, ctx->gate);
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}
