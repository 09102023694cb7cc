// Symmetric Gauss-Seidel preconditioner application  z = (D+U)^-1 D (D+L)^-1 r  for sm_100a
// (reference: SGS_like_preconditioner::apply, src/linalg/s3d_sgspreconditioners.cpp:14-115).
//
// The reference sweeps the grid lexicographically: forward  y = dinv (r - A0 y[i-1] - A1 y[j-1] - A2 y[k-1]),
// backward z = y - dinv (A4 z[i+1] + A5 z[j+1] + A6 z[k+1]).  Two GPU orderings are offered:
//
//  S3D_SGS_LEXICOGRAPHIC  the reference's ordering, executed as a wavefront: all points on a hyperplane
//      i+j+k = const are independent once the previous hyperplane is done, and each point evaluates the
//      very same expression, so the result equals the sequential sweep bit for bit (same roundings:
//      __dmul_rn/__dsub_rn, no FMA contraction).  The iteration count of any solver using it therefore
//      equals the reference's.  (Round-1 implementation: one launch per hyperplane; the tile-level
//      dataflow version is the next kernel, see DESIGN.md.)
//
//  S3D_SGS_REDBLACK       the same product M = (D+L) D^-1 (D+U) with the unknowns ordered red-then-black
//      ((i+j+k) parity).  Every half-sweep is fully parallel: three streaming launches, no dependencies.
//      It changes the iteration (different M), which is why it is an option and not the default.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------- lexicographic wavefront (exact)

// forward hyperplane h: points with i+j+k == h; threads enumerate (j,k), i = h-j-k
__global__ void __launch_bounds__(256)
sgs_fwd_plane_kernel(const GridDev g, const double *__restrict__ A, const double *__restrict__ dinv,
                     const double *__restrict__ r, double *y, int h, const int *stop)
{
	if (s3d_stopped(stop)) return;
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	const int k = blockIdx.y * blockDim.y + threadIdx.y;
	if (j >= g.ny || k >= g.nz) return;
	const int i = h - j - k;
	if (i < 0 || i >= g.nx) return;
	const long long v = g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + i;
	const long long c = (long long)k * g.cplane + (long long)j * g.cpitch + i;
	double t = __dsub_rn(r[v], __dmul_rn(A[c], y[v - 1]));
	t = __dsub_rn(t, __dmul_rn(A[c + g.cstride], y[v - g.vpitch]));
	t = __dsub_rn(t, __dmul_rn(A[c + 2 * g.cstride], y[v - g.vplane]));
	y[v] = __dmul_rn(t, dinv[c]);
}

__global__ void __launch_bounds__(256)
sgs_bwd_plane_kernel(const GridDev g, const double *__restrict__ A, const double *__restrict__ dinv,
                     const double *__restrict__ y, double *z, int h, const int *stop)
{
	if (s3d_stopped(stop)) return;
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	const int k = blockIdx.y * blockDim.y + threadIdx.y;
	if (j >= g.ny || k >= g.nz) return;
	const int i = h - j - k;
	if (i < 0 || i >= g.nx) return;
	const long long v = g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + i;
	const long long c = (long long)k * g.cplane + (long long)j * g.cpitch + i;
	double s = __dmul_rn(A[c + 4 * g.cstride], z[v + 1]);
	s = __dadd_rn(s, __dmul_rn(A[c + 5 * g.cstride], z[v + g.vpitch]));
	s = __dadd_rn(s, __dmul_rn(A[c + 6 * g.cstride], z[v + g.vplane]));
	z[v] = __dsub_rn(y[v], __dmul_rn(dinv[c], s));
}

// ---------------------------------------------------------------- red-black

// phase 0: y_R = r_R dinv                       (red:   (i+j+k) even in the reference's 1-based indices = odd+... see parity())
// phase 1: y_B = (r_B - sum_6 A_s y_s) dinv ;  z_B = y_B
// phase 2: z_R = y_R - dinv sum_6 A_s z_s
__device__ __forceinline__ int parity(int i, int j, int k, int k0) { return (i + j + k + k0 + 3) & 1; }   // 1-based (i+1)+(j+1)+(k+1)

template <int PHASE>
__global__ void __launch_bounds__(256)
sgs_rb_kernel(const GridDev g, int k0, const double *__restrict__ A, const double *__restrict__ dinv,
              const double *__restrict__ r, double *y, double *z, const int *stop)
{
	if (s3d_stopped(stop)) return;
	const unsigned nx2 = (unsigned)(g.nx + 1) >> 1;
	const unsigned total = nx2 * (unsigned)g.ny * (unsigned)g.nz;
	constexpr int COLOUR = (PHASE == 1) ? 1 : 0;   // 0 = red (even), 1 = black
	for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
		const unsigned row = it / nx2;
		const int k = (int)(row / (unsigned)g.ny);
		const int j = (int)(row - (unsigned)k * (unsigned)g.ny);
		int i = 2 * (int)(it - row * nx2);
		if (parity(i, j, k, k0) != COLOUR) i += 1;   // exactly one point of each aligned pair has the colour
		if (i >= g.nx) continue;
		const long long v = g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + i;
		const long long c = (long long)k * g.cplane + (long long)j * g.cpitch + i;
		const long long cs = g.cstride;
		if (PHASE == 0) {
			y[v] = __dmul_rn(r[v], dinv[c]);
		} else if (PHASE == 1) {
			double t = __dsub_rn(r[v], __dmul_rn(A[c], y[v - 1]));
			t = __dsub_rn(t, __dmul_rn(A[c + cs], y[v - g.vpitch]));
			t = __dsub_rn(t, __dmul_rn(A[c + 2 * cs], y[v - g.vplane]));
			t = __dsub_rn(t, __dmul_rn(A[c + 4 * cs], y[v + 1]));
			t = __dsub_rn(t, __dmul_rn(A[c + 5 * cs], y[v + g.vpitch]));
			t = __dsub_rn(t, __dmul_rn(A[c + 6 * cs], y[v + g.vplane]));
			t = __dmul_rn(t, dinv[c]);
			y[v] = t;
			z[v] = t;
		} else {
			double s = __dmul_rn(A[c], z[v - 1]);
			s = __dadd_rn(s, __dmul_rn(A[c + cs], z[v - g.vpitch]));
			s = __dadd_rn(s, __dmul_rn(A[c + 2 * cs], z[v - g.vplane]));
			s = __dadd_rn(s, __dmul_rn(A[c + 4 * cs], z[v + 1]));
			s = __dadd_rn(s, __dmul_rn(A[c + 5 * cs], z[v + g.vpitch]));
			s = __dadd_rn(s, __dmul_rn(A[c + 6 * cs], z[v + g.vplane]));
			z[v] = __dsub_rn(y[v], __dmul_rn(dinv[c], s));
		}
	}
}

unsigned ew_blocks(const s3d_ctx *ctx, const s3d_grid *g)
{
	const unsigned long long total = (unsigned long long)((g->nx + 1) / 2) * g->ny * g->nz;
	unsigned long long b = (total + 255) / 256;
	const unsigned long long cap = (unsigned long long)ctx->sm_count * 16;
	if (b > cap) b = cap;
	return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

int s3d_sgs_apply(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *d_dinv, const double *d_r,
                  double *d_z, double *d_y, int ordering, int nsweeps)
{
	S3D_REQUIRE(ctx, g && A && d_dinv && d_r && d_z && d_y, "null argument");
	S3D_REQUIRE(ctx, d_r != d_z && d_r != d_y && d_z != d_y, "r, z and the scratch vector must be distinct");
	S3D_REQUIRE(ctx, nsweeps >= 1, "need at least one sweep");
	const GridDev gd = to_dev(g);
	if (ordering == S3D_SGS_REDBLACK) {
		const unsigned nb = ew_blocks(ctx, g);
		sgs_rb_kernel<0><<<nb, 256, 0, ctx->stream>>>(gd, g->k0, A, d_dinv, d_r, d_y, d_z, ctx->gate);
		S3D_LAUNCHED(ctx);
		int rc = s3d_halo_exchange(ctx, g, d_y);
		if (rc != S3D_OK) return rc;
		sgs_rb_kernel<1><<<nb, 256, 0, ctx->stream>>>(gd, g->k0, A, d_dinv, d_r, d_y, d_z, ctx->gate);
		S3D_LAUNCHED(ctx);
		rc = s3d_halo_exchange(ctx, g, d_z);
		if (rc != S3D_OK) return rc;
		sgs_rb_kernel<2><<<nb, 256, 0, ctx->stream>>>(gd, g->k0, A, d_dinv, d_r, d_y, d_z, ctx->gate);
		S3D_LAUNCHED(ctx);
		return S3D_OK;
	}
	if (ordering != S3D_SGS_LEXICOGRAPHIC) return s3d_fail(ctx, S3D_ERR_ARG, "unknown SGS ordering %d", ordering);
	if (g->nranks > 1)
		return s3d_fail(ctx, S3D_ERR_ARG, "lexicographic SGS serialises across z-slabs; use S3D_SGS_REDBLACK on a decomposed grid");
	// the scratch vector's real entries are all overwritten and its ghosts stay zero (they are never
	// written), so the per-call zero fill of the reference (s3d_sgspreconditioners.cpp:25-28) is implicit.
	// Sequential sweeps are idempotent, so nsweeps > 1 repeats identical work; do one.
	dim3 block(32, 8);
	dim3 grid((g->ny + 31) / 32, (g->nz + 7) / 8);
	const int hmax = g->nx + g->ny + g->nz - 3;
	for (int h = 0; h <= hmax; h++) {
		sgs_fwd_plane_kernel<<<grid, block, 0, ctx->stream>>>(gd, A, d_dinv, d_r, d_y, h, ctx->gate);
		ctx->launches++;
	}
	for (int h = hmax; h >= 0; h--) {
		sgs_bwd_plane_kernel<<<grid, block, 0, ctx->stream>>>(gd, A, d_dinv, d_y, d_z, h, ctx->gate);
		ctx->launches++;
	}
	S3D_CUDA(ctx, cudaGetLastError());
	return S3D_OK;
}

}  // extern "C"
