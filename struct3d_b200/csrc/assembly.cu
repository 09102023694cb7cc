// GPU assembly of the 7-point stencil coefficients and of separable right-hand sides
// (reference: PDEBase::computeLHS / computeVector, src/pde/pdebase.cpp:50-72,123-150;
//  Poisson::lhsmat_kernel poisson.cpp:28-52; ConvDiff::lhsmat_kernel convdiff.cpp:48-88).
//
// The reference evaluates every row from the 1-D coordinate arrays, serially on the host
// (11.8 s at 512^3).  Each coefficient depends on ONE axis index only, so a
// tiny kernel first builds per-axis tables with exactly the reference's rounding sequence, and a
// streaming kernel then writes the 7 arrays: 56 B/point of pure, 128-bit coalesced stores.
#include <cmath>
#include <cstdlib>
#include "common.cuh"

namespace {

// per-axis table layout inside ctx->tab: for axis d with n_d real points, 4 arrays of n_d doubles
//   lo[m]  = coefficient of the (m-1) neighbour      up[m] = coefficient of the (m+1) neighbour
//   dg[m]  = this axis' diffusion contribution to the diagonal
//   ad[m]  = this axis' signed advection contribution to the diagonal (0 for Poisson)
struct AxisTab { const double *lo, *up, *dg, *ad; };

// c: device copy of the axis coordinates incl. boundary points; real index m <-> coordinate index m+1+shift
__global__ void axis_table_kernel(const double *c, int n, int shift, double bvel, double mu, int convdiff,
                                  double *lo, double *up, double *dg, double *ad, double *o_drm, double *o_drp)
{
	const int m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= n) return;
	const int I = m + 1 + shift;
	const double drm = __dsub_rn(c[I], c[I - 1]);
	const double drp = __dsub_rn(c[I + 1], c[I]);
	const double drs = __dsub_rn(c[I + 1], c[I - 1]);
	// poisson.cpp:46-50:  v[j] = -1/(drm*0.5*drs);  v[3] += 2/drs*(1/drp + 1/drm);  v[4+j] = -1/(drp*0.5*drs)
	double l = __ddiv_rn(-1.0, __dmul_rn(__dmul_rn(drm, 0.5), drs));
	double u = __ddiv_rn(-1.0, __dmul_rn(__dmul_rn(drp, 0.5), drs));
	const double d = __dmul_rn(__ddiv_rn(2.0, drs), __dadd_rn(__ddiv_rn(1.0, drp), __ddiv_rn(1.0, drm)));
	double a = 0.0;
	if (o_drm) { o_drm[m] = drm; o_drp[m] = drp; }
	if (convdiff == 2) {   // point-dependent velocity (ConvDiffCirc): only the mu scaling is separable
		l = __dmul_rn(l, mu);
		u = __dmul_rn(u, mu);
	} else if (convdiff) {
		// convdiff.cpp:74-87: all seven entries times mu, then first-order upwinding (b == 0 takes the else branch)
		l = __dmul_rn(l, mu);
		u = __dmul_rn(u, mu);
		if (bvel > 0) {
			const double t = __ddiv_rn(bvel, drm);
			l = __dsub_rn(l, t);
			a = t;
		} else {
			const double t = __ddiv_rn(bvel, drp);
			u = __dadd_rn(u, t);
			a = -t;
		}
	}
	lo[m] = l; up[m] = u; dg[m] = d; ad[m] = a;
}

// one thread per aligned pair of i-points; rows beyond nx are padding and are left untouched
__global__ void __launch_bounds__(256)
fill_stencil_kernel(const GridDev g, const AxisTab tx, const AxisTab ty, const AxisTab tz, double mu, int convdiff,
                    double *__restrict__ A)
{
	const unsigned nx2 = (unsigned)(g.nx + 1) >> 1;
	const unsigned total = nx2 * (unsigned)g.ny * (unsigned)g.nz;
	for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
		const unsigned row = it / nx2;
		const int i = 2 * (int)(it - row * nx2);
		const unsigned k = row / (unsigned)g.ny;
		const unsigned j = row - k * (unsigned)g.ny;
		const long long c = (long long)k * g.cplane + (long long)j * g.cpitch + i;
		const double dyz_j = __ldg(ty.dg + j), dyz_k = __ldg(tz.dg + k);
		const double ay = __ldg(ty.ad + j), az = __ldg(tz.ad + k);
		const double loy = __ldg(ty.lo + j), upy = __ldg(ty.up + j), loz = __ldg(tz.lo + k), upz = __ldg(tz.up + k);
		const int np = (i + 1 < g.nx) ? 2 : 1;
		double v[7][2];
		for (int e = 0; e < np; e++) {
			// diagonal: ((0 + dgx) + dgy) + dgz, [times mu], then the advection terms in axis order
			double dsum = __dadd_rn(__dadd_rn(__ldg(tx.dg + i + e), dyz_j), dyz_k);
			if (convdiff) {
				dsum = __dmul_rn(dsum, mu);
				dsum = __dadd_rn(dsum, __ldg(tx.ad + i + e));
				dsum = __dadd_rn(dsum, ay);
				dsum = __dadd_rn(dsum, az);
			}
			v[0][e] = __ldg(tx.lo + i + e); v[1][e] = loy; v[2][e] = loz; v[3][e] = dsum;
			v[4][e] = __ldg(tx.up + i + e); v[5][e] = upy; v[6][e] = upz;
		}
#pragma unroll
		for (int s = 0; s < 7; s++) {
			double *p = A + (long long)s * g.cstride + c;
			if (np == 2) st_stream2(p, make_double2(v[s][0], v[s][1]));
			else *p = v[s][0];
		}
	}
}

// ConvDiffCirc (reference src/pde/convdiff_circular.cpp:59-102): mu * diffusion from the per-axis tables, then first-order upwinding
// with the point-dependent velocity b = (2y(1-x^2), -2x(1-y^2), sin(pi z)); b_d >= 0 takes the backward difference.  Reproduced as the
// reference has it (its own TODO notes the convection part is not final; bmag does not enter the matrix).  sin(pi z_k) comes from a host
// libm table so the coefficients are bit-identical to the reference's.
struct CircTab { const double *x, *y, *sinz; const double *drm[3], *drp[3]; };

__global__ void __launch_bounds__(256)
fill_stencil_circ_kernel(const GridDev g, const AxisTab tx, const AxisTab ty, const AxisTab tz, const CircTab ct, double mu,
                         double *__restrict__ A)
{
	const unsigned total = (unsigned)g.nx * (unsigned)g.ny * (unsigned)g.nz;
	for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
		const int i = it % g.nx, j = (it / g.nx) % g.ny, k = it / (g.nx * g.ny);
		double v[7];
		v[0] = __ldg(tx.lo + i); v[1] = __ldg(ty.lo + j); v[2] = __ldg(tz.lo + k);
		v[4] = __ldg(tx.up + i); v[5] = __ldg(ty.up + j); v[6] = __ldg(tz.up + k);
		v[3] = __dmul_rn(__dadd_rn(__dadd_rn(__ldg(tx.dg + i), __ldg(ty.dg + j)), __ldg(tz.dg + k)), mu);
		const double x = __ldg(ct.x + i), y = __ldg(ct.y + j);
		double b[3];
		b[0] = __dmul_rn(__dmul_rn(2.0, y), __dsub_rn(1.0, __dmul_rn(x, x)));
		b[1] = __dmul_rn(__dmul_rn(-2.0, x), __dsub_rn(1.0, __dmul_rn(y, y)));
		b[2] = __ldg(ct.sinz + k);
		const int idx[3] = {i, j, k};
#pragma unroll
		for (int d = 0; d < 3; d++) {
			if (b[d] >= 0) {
				const double t = __ddiv_rn(b[d], __ldg(ct.drm[d] + idx[d]));
				v[d] = __dsub_rn(v[d], t);
				v[3] = __dadd_rn(v[3], t);
			} else {
				const double t = __ddiv_rn(b[d], __ldg(ct.drp[d] + idx[d]));
				v[d + 4] = __dadd_rn(v[d + 4], t);
				v[3] = __dsub_rn(v[3], t);
			}
		}
		const long long c = (long long)k * g.cplane + (long long)j * g.cpitch + i;
#pragma unroll
		for (int s = 0; s < 7; s++) A[(long long)s * g.cstride + c] = v[s];
	}
}

struct SepTab { const double *tx, *ty, *tz; int px, py, pz; };   // row pitch of each table = n_axis + 2

__global__ void __launch_bounds__(256)
rhs_separable_kernel(const GridDev g, const SepTab T, int nterms, int k0, int nzg, const double *face, double *__restrict__ f)
{
	const unsigned nx2 = (unsigned)(g.nx + 1) >> 1;
	const unsigned total = nx2 * (unsigned)g.ny * (unsigned)g.nz;
	for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
		const unsigned row = it / nx2;
		const int i = 2 * (int)(it - row * nx2);
		const int k = (int)(row / (unsigned)g.ny);
		const int j = (int)(row - (unsigned)k * (unsigned)g.ny);
		const long long v = g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + i;
		const int np = (i + 1 < g.nx) ? 2 : 1;
		double out[2] = {0.0, 0.0};
		for (int e = 0; e < np; e++) {
			double val = 0.0;
			for (int t = 0; t < nterms; t++) {
				const double term = __dmul_rn(__dmul_rn(__ldg(T.tx + t * T.px + i + e + 1), __ldg(T.ty + t * T.py + j + 1)),
				                              __ldg(T.tz + t * T.pz + k0 + k + 1));
				val = (t == 0) ? term : __dadd_rn(val, term);
			}
			if (face) {   // Poisson::rhs_kernel face terms, in the reference's order i-,i+,j-,j+,k-,k+
				if (i + e == 0) val = __dadd_rn(val, face[0]);
				if (i + e == g.nx - 1) val = __dadd_rn(val, face[1]);
				if (j == 0) val = __dadd_rn(val, face[2]);
				if (j == g.ny - 1) val = __dadd_rn(val, face[3]);
				if (k0 + k == 0) val = __dadd_rn(val, face[4]);
				if (k0 + k == nzg - 1) val = __dadd_rn(val, face[5]);
			}
			out[e] = val;
		}
		if (np == 2) st2(f + v, make_double2(out[0], out[1]));
		else f[v] = out[0];
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

int assemble(s3d_ctx *ctx, const s3d_grid *g, const double *h_cx, const double *h_cy, const double *h_cz,
             const double *vel, double mu, int convdiff, double *A)
{
	S3D_REQUIRE(ctx, g && h_cx && h_cy && h_cz && A, "null argument");
	const int n[3] = {g->nx, g->ny, g->nz};
	const int ncoord[3] = {g->nx + 2, g->ny + 2, g->nz_global + 2};
	const double *hc[3] = {h_cx, h_cy, h_cz};
	size_t need = 0;
	for (int d = 0; d < 3; d++) need += (size_t)ncoord[d] + 4 * (size_t)n[d];
	const int rc = s3d_ensure_tab(ctx, need);
	if (rc != S3D_OK) return rc;
	AxisTab T[3];
	double *p = ctx->tab;
	for (int d = 0; d < 3; d++) {
		double *dc = p; p += ncoord[d];
		S3D_CUDA(ctx, cudaMemcpyAsync(dc, hc[d], ncoord[d] * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
		double *lo = p, *up = p + n[d], *dg = p + 2 * n[d], *ad = p + 3 * n[d];
		p += 4 * (size_t)n[d];
		const int shift = (d == 2) ? g->k0 : 0;
		axis_table_kernel<<<(n[d] + 127) / 128, 128, 0, ctx->stream>>>(dc, n[d], shift, vel ? vel[d] : 0.0, mu, convdiff, lo, up, dg, ad, nullptr, nullptr);
		S3D_LAUNCHED(ctx);
		T[d] = AxisTab{lo, up, dg, ad};
	}
	// the host coordinate arrays may be freed by the caller as soon as we return
	S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	fill_stencil_kernel<<<ew_blocks(ctx, g), 256, 0, ctx->stream>>>(to_dev(g), T[0], T[1], T[2], mu, convdiff, A);
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}

int assemble_circ(s3d_ctx *ctx, const s3d_grid *g, const double *h_cx, const double *h_cy, const double *h_cz, double mu, double *A)
{
	S3D_REQUIRE(ctx, g && h_cx && h_cy && h_cz && A, "null argument");
	const int n[3] = {g->nx, g->ny, g->nz};
	const int ncoord[3] = {g->nx + 2, g->ny + 2, g->nz_global + 2};
	const double *hc[3] = {h_cx, h_cy, h_cz};
	size_t need = 0;
	for (int d = 0; d < 3; d++) need += (size_t)ncoord[d] + 6 * (size_t)n[d];
	need += (size_t)n[2];
	const int rc = s3d_ensure_tab(ctx, need);
	if (rc != S3D_OK) return rc;
	AxisTab T[3];
	CircTab ct;
	double *p = ctx->tab;
	const double *dcoord[3];
	for (int d = 0; d < 3; d++) {
		double *dc = p; p += ncoord[d];
		dcoord[d] = dc;
		S3D_CUDA(ctx, cudaMemcpyAsync(dc, hc[d], ncoord[d] * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
		double *lo = p, *up = p + n[d], *dg = p + 2 * n[d], *ad = p + 3 * n[d], *drm = p + 4 * n[d], *drp = p + 5 * n[d];
		p += 6 * (size_t)n[d];
		const int shift = (d == 2) ? g->k0 : 0;
		axis_table_kernel<<<(n[d] + 127) / 128, 128, 0, ctx->stream>>>(dc, n[d], shift, 0.0, mu, 2, lo, up, dg, ad, drm, drp);
		S3D_LAUNCHED(ctx);
		T[d] = AxisTab{lo, up, dg, ad};
		ct.drm[d] = drm; ct.drp[d] = drp;
	}
	// sin(pi z) on the host (libm), for this rank's planes
	double *hs = (double *)malloc(sizeof(double) * n[2]);
	for (int k = 0; k < n[2]; k++) hs[k] = sin(3.141592653589793238 * h_cz[g->k0 + k + 1]);
	double *dsin = p;
	cudaError_t e = cudaMemcpyAsync(dsin, hs, n[2] * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	free(hs);
	S3D_CUDA(ctx, e);
	ct.x = dcoord[0] + 1; ct.y = dcoord[1] + 1; ct.sinz = dsin;   // real index m <-> coordinate index m + 1
	const unsigned long long total = (unsigned long long)g->nx * g->ny * g->nz;
	unsigned long long nb = (total + 255) / 256;
	if (nb > (unsigned long long)ctx->sm_count * 16) nb = (unsigned long long)ctx->sm_count * 16;
	fill_stencil_circ_kernel<<<(unsigned)nb, 256, 0, ctx->stream>>>(to_dev(g), T[0], T[1], T[2], ct, mu, A);
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}

}  // namespace

extern "C" {

int s3d_assemble_convdiff_circ(s3d_ctx *ctx, const s3d_grid *g, const double *h_cx, const double *h_cy, const double *h_cz,
                               double mu, double *A)
{
	return assemble_circ(ctx, g, h_cx, h_cy, h_cz, mu, A);
}


int s3d_assemble_poisson(s3d_ctx *ctx, const s3d_grid *g, const double *h_cx, const double *h_cy, const double *h_cz,
                         double *A)
{
	return assemble(ctx, g, h_cx, h_cy, h_cz, nullptr, 1.0, 0, A);
}

int s3d_assemble_convdiff(s3d_ctx *ctx, const s3d_grid *g, const double *h_cx, const double *h_cy, const double *h_cz,
                          const double vel[3], double mu, double *A)
{
	S3D_REQUIRE(ctx, vel, "null velocity");
	return assemble(ctx, g, h_cx, h_cy, h_cz, vel, mu, 1, A);
}

int s3d_assemble_rhs_separable(s3d_ctx *ctx, const s3d_grid *g, int nterms, const double *h_tx, const double *h_ty,
                               const double *h_tz, const double *h_face, double *d_f)
{
	S3D_REQUIRE(ctx, g && h_tx && h_ty && h_tz && d_f && nterms >= 1 && nterms <= 16, "bad argument");
	const int px = g->nx + 2, py = g->ny + 2, pz = g->nz_global + 2;
	const size_t need = (size_t)nterms * (px + py + pz) + 8;
	const int rc = s3d_ensure_tab(ctx, need);
	if (rc != S3D_OK) return rc;
	double *dx = ctx->tab, *dy = dx + (size_t)nterms * px, *dz = dy + (size_t)nterms * py, *df = dz + (size_t)nterms * pz;
	S3D_CUDA(ctx, cudaMemcpyAsync(dx, h_tx, (size_t)nterms * px * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	S3D_CUDA(ctx, cudaMemcpyAsync(dy, h_ty, (size_t)nterms * py * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	S3D_CUDA(ctx, cudaMemcpyAsync(dz, h_tz, (size_t)nterms * pz * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	if (h_face) S3D_CUDA(ctx, cudaMemcpyAsync(df, h_face, 6 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	SepTab T{dx, dy, dz, px, py, pz};
	rhs_separable_kernel<<<ew_blocks(ctx, g), 256, 0, ctx->stream>>>(to_dev(g), T, nterms, g->k0, g->nz_global,
	                                                                  h_face ? df : nullptr, d_f);
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}

}  // extern "C"
