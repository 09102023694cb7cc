/* TEST INFRASTRUCTURE ONLY -- see s3d_oracle.h for the rules on who may load this.
 *
 * Plain-C restatement of the Struct3d native path.  Written from the reference's behaviour,
 * not transcribed: one flat index convention, one loop macro, C arrays instead of the C++
 * classes.  Citations are relative to /root/reference/.  Parity: PINNED (header).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "s3d_oracle.h"

#define ORC_PI 3.141592653589793238 /* src/struct3d_config.h:19 */

typedef long long i64;

/* ghosted ("All") and compact ("Real") flat indices, src/cartmesh.hpp:141-154 */
#define GI(i, j, k) ((i64)(k) * np1 * np0 + (i64)(j) * np0 + (i))
#define RI(i, j, k) ((i64)(k) * ny * nx + (i64)(j) * nx + (i))
#define DIMS                                                                 \
	const int np0 = npdim[0], np1 = npdim[1], np2 = npdim[2];                \
	const int nx = np0 - 2, ny = np1 - 2, nz = np2 - 2;                      \
	const i64 N = (i64)nx * ny * nz;                                         \
	(void)np2; (void)nz; (void)N
/* iterate the real points in the reference's order: k slowest, i fastest, 1-based ghosted indices */
#define FOR_REAL                                                             \
	for (int k = 1; k <= nz; k++)                                            \
		for (int j = 1; j <= ny; j++)                                        \
			for (int i = 1; i <= nx; i++)

/* ------------------------------------------------------------------ mesh */

/* src/cartmesh.cpp:187-192 */
void orc_coords_uniform(int n, double rmin, double rmax, double *c)
{
	for (int i = 0; i < n; i++) c[i] = rmin + (rmax - rmin) * i / (n - 1);
}

/* src/cartmesh.cpp:164-172 */
void orc_coords_chebyshev(int n, double rmin, double rmax, double *c)
{
	const double theta = ORC_PI / (n - 1);
	for (int i = 0; i < n; i++) c[i] = (rmax + rmin) * 0.5 + (rmax - rmin) * 0.5 * cos(ORC_PI - i * theta);
}

/* src/cartmesh.cpp:198-208: returns 1 where the reference would throw "not actually uniform" */
int orc_uniform_selfcheck(int n, const double *c)
{
	const double dr = c[1] - c[0];
	for (int j = 0; j < n - 1; j++)
		if (fabs(c[j + 1] - c[j] - dr) > 2.220446049250313e-16) return 1;
	return 0;
}

/* src/cartmesh.cpp:10-33.  The reference scans all cells; the cell diagonal is monotone in each
 * 1-D spacing, so the maximum is attained at the three per-axis maxima (same rounding sequence). */
double orc_mesh_h(const int npdim[3], const double *cx, const double *cy, const double *cz)
{
	const double *c[3] = {cx, cy, cz};
	double hmax[3] = {0, 0, 0};
	for (int d = 0; d < 3; d++)
		for (int i = 0; i < npdim[d] - 1; i++) {
			const double h = c[d][i + 1] - c[d][i];
			if (h > hmax[d]) hmax[d] = h;
		}
	double diam = 0;
	for (int d = 0; d < 3; d++) diam += hmax[d] * hmax[d];
	return sqrt(diam);
}

/* -------------------------------------------------------------- assembly */

/* the three spacings around 1-D node I: src/pde/poisson.cpp:34-42 */
static void spacings(const double *c, int I, double *drm, double *drp, double *drs)
{
	*drm = c[I] - c[I - 1];
	*drp = c[I + 1] - c[I];
	*drs = c[I + 1] - c[I - 1];
}

/* diffusion part of one row, src/pde/poisson.cpp:44-51 (same loop in convdiff.cpp:65-72) */
static void diffusion_row(const double drm[3], const double drp[3], const double drs[3], double v[7])
{
	v[3] = 0;
	for (int d = 0; d < 3; d++) {
		v[d] = -1.0 / (drm[d] * 0.5 * drs[d]);
		v[3] += 2.0 / drs[d] * (1.0 / drp[d] + 1.0 / drm[d]);
		v[4 + d] = -1.0 / (drp[d] * 0.5 * drs[d]);
	}
}

/* src/pde/pdebase.cpp:123-150 + src/pde/poisson.cpp:28-52 */
void orc_lhs_poisson(const int npdim[3], const double *cx, const double *cy, const double *cz, double *A)
{
	DIMS;
	FOR_REAL {
		double drm[3], drp[3], drs[3], v[7];
		spacings(cx, i, &drm[0], &drp[0], &drs[0]);
		spacings(cy, j, &drm[1], &drp[1], &drs[1]);
		spacings(cz, k, &drm[2], &drp[2], &drs[2]);
		diffusion_row(drm, drp, drs, v);
		const i64 r = RI(i - 1, j - 1, k - 1);
		for (int s = 0; s < 7; s++) A[s * N + r] = v[s];
	}
}

/* src/pde/pdebase.cpp:123-150 + src/pde/convdiff.cpp:48-88.  Note b_d == 0 takes the else branch. */
void orc_lhs_convdiff(const int npdim[3], const double *cx, const double *cy, const double *cz,
                      const double b[3], double mu, double *A)
{
	DIMS;
	FOR_REAL {
		double drm[3], drp[3], drs[3], v[7];
		spacings(cx, i, &drm[0], &drp[0], &drs[0]);
		spacings(cy, j, &drm[1], &drp[1], &drs[1]);
		spacings(cz, k, &drm[2], &drp[2], &drs[2]);
		diffusion_row(drm, drp, drs, v);
		for (int s = 0; s < 7; s++) v[s] *= mu;
		for (int d = 0; d < 3; d++) {
			if (b[d] > 0) { v[d] -= b[d] / drm[d]; v[3] += b[d] / drm[d]; }
			else          { v[d + 4] += b[d] / drp[d]; v[3] -= b[d] / drp[d]; }
		}
		const i64 r = RI(i - 1, j - 1, k - 1);
		for (int s = 0; s < 7; s++) A[s * N + r] = v[s];
	}
}

/* src/pde/convdiff_circular.cpp:20-27 */
static void circ_velocity(const double r[3], double v[3])
{
	v[0] = 2 * r[1] * (1.0 - r[0] * r[0]);
	v[1] = -2 * r[0] * (1.0 - r[1] * r[1]);
	v[2] = sin(ORC_PI * r[2]);
}

/* src/pde/pdebase.cpp:123-150 + src/pde/convdiff_circular.cpp:59-102 (reproduced as it is, including the reference's own
 * "TODO: Correct convection part": bmag is not applied to the matrix, and b_d == 0 takes the FIRST branch here). */
void orc_lhs_convdiff_circ(const int npdim[3], const double *cx, const double *cy, const double *cz, double mu, double *A)
{
	DIMS;
	FOR_REAL {
		double drm[3], drp[3], drs[3], v[7], b[3];
		spacings(cx, i, &drm[0], &drp[0], &drs[0]);
		spacings(cy, j, &drm[1], &drp[1], &drs[1]);
		spacings(cz, k, &drm[2], &drp[2], &drs[2]);
		diffusion_row(drm, drp, drs, v);
		for (int s = 0; s < 7; s++) v[s] *= mu;
		const double r[3] = {cx[i], cy[j], cz[k]};
		circ_velocity(r, b);
		for (int d = 0; d < 3; d++) {
			if (b[d] >= 0) { v[d] -= b[d] / drm[d]; v[3] += b[d] / drm[d]; }
			else           { v[d + 4] += b[d] / drp[d]; v[3] -= b[d] / drp[d]; }
		}
		const i64 q = RI(i - 1, j - 1, k - 1);
		for (int s = 0; s < 7; s++) A[s * N + q] = v[s];
	}
}

/* the source / solution functions: src/pde/poisson.cpp:15-26, src/pde/convdiff.cpp:20-46 */
static double eval_function(int pde, int func, const double b[3], double mu, const double r[3])
{
	if (func == ORC_F_TEST_RHS)
		return sin(ORC_PI * r[0]) * sin(ORC_PI * r[1]) * sin(ORC_PI * r[2]);
	if (func == ORC_F_MANUFACTURED_SOLUTION)
		return sin(2 * ORC_PI * r[0]) * sin(2 * ORC_PI * r[1]) * sin(2 * ORC_PI * r[2]);
	if (pde == ORC_PDE_POISSON)
		return 12 * ORC_PI * ORC_PI * sin(2 * ORC_PI * r[0]) * sin(2 * ORC_PI * r[1]) * sin(2 * ORC_PI * r[2]);
	double val = mu * 12 * ORC_PI * ORC_PI * sin(2 * ORC_PI * r[0]) * sin(2 * ORC_PI * r[1]) * sin(2 * ORC_PI * r[2]);
	if (pde == ORC_PDE_CONVDIFF_CIRC) {   /* src/pde/convdiff_circular.cpp:37-49; b[0] carries bmag */
		double vel[3];
		circ_velocity(r, vel);
		for (int d = 0; d < 3; d++) {
			double term = 2 * ORC_PI * b[0] * vel[d];
			for (int e = 0; e < 3; e++) term *= (d == e) ? cos(2 * ORC_PI * r[e]) : sin(2 * ORC_PI * r[e]);
			val += term;
		}
		return val;
	}
	for (int d = 0; d < 3; d++) {
		double term = 2 * ORC_PI * b[d];
		for (int e = 0; e < 3; e++) term *= (d == e) ? cos(2 * ORC_PI * r[e]) : sin(2 * ORC_PI * r[e]);
		val += term;
	}
	return val;
}

/* src/pde/pdebase.cpp:50-72 with rhs_kernel of src/pde/poisson.cpp:54-122 (Dirichlet face terms) or
 * src/pde/convdiff.cpp:90-123 (no face terms).  bvals == NULL means all-zero boundary values.
 * Returns 0; the reference throws on non-Dirichlet faces, which this restatement does not model. */
int orc_grid_function(const int npdim[3], const double *cx, const double *cy, const double *cz,
                      int pde, int func, const double b[3], double mu, const double bvals[6], double *f)
{
	DIMS;
	memset(f, 0, sizeof(double) * (size_t)np0 * np1 * np2);
	const double *c[3] = {cx, cy, cz};
	FOR_REAL {
		const double r[3] = {cx[i], cy[j], cz[k]};
		double val = eval_function(pde, func, b, mu, r);
		if (pde == ORC_PDE_POISSON && bvals) {
			const int idx[3] = {i, j, k};
			for (int d = 0; d < 3; d++) {
				double drm, drp, drs;
				spacings(c[d], idx[d], &drm, &drp, &drs);
				if (idx[d] == 1) val += bvals[2 * d] / (drm * 0.5 * drs);
				if (idx[d] == npdim[d] - 2) val += bvals[2 * d + 1] / (drp * 0.5 * drs);
			}
		}
		f[GI(i, j, k)] = val;
	}
	return 0;
}

/* ------------------------------------------------------- matvec + BLAS-1 */

/* src/linalg/matvec.cpp:59-101: seven products summed in stencil order 0..6 */
void orc_apply(const int npdim[3], const double *A, const double *x, double *y)
{
	DIMS;
#pragma omp parallel for collapse(2)
	for (int k = 1; k <= nz; k++)
		for (int j = 1; j <= ny; j++)
			for (int i = 1; i <= nx; i++) {
				const i64 r = RI(i - 1, j - 1, k - 1);
				y[GI(i, j, k)] = A[0 * N + r] * x[GI(i - 1, j, k)] + A[1 * N + r] * x[GI(i, j - 1, k)]
				               + A[2 * N + r] * x[GI(i, j, k - 1)] + A[3 * N + r] * x[GI(i, j, k)]
				               + A[4 * N + r] * x[GI(i + 1, j, k)] + A[5 * N + r] * x[GI(i, j + 1, k)]
				               + A[6 * N + r] * x[GI(i, j, k + 1)];
			}
}

/* src/linalg/matvec.cpp:103-146: y = b, then y -= (the same seven-term sum) */
void orc_apply_res(const int npdim[3], const double *A, const double *b, const double *x, double *y)
{
	DIMS;
#pragma omp parallel for collapse(2)
	for (int k = 1; k <= nz; k++)
		for (int j = 1; j <= ny; j++)
			for (int i = 1; i <= nx; i++) {
				const i64 r = RI(i - 1, j - 1, k - 1);
				const double ax = A[0 * N + r] * x[GI(i - 1, j, k)] + A[1 * N + r] * x[GI(i, j - 1, k)]
				                + A[2 * N + r] * x[GI(i, j, k - 1)] + A[3 * N + r] * x[GI(i, j, k)]
				                + A[4 * N + r] * x[GI(i + 1, j, k)] + A[5 * N + r] * x[GI(i, j + 1, k)]
				                + A[6 * N + r] * x[GI(i, j, k + 1)];
				y[GI(i, j, k)] = b[GI(i, j, k)] - ax;
			}
}

/* src/linalg/matvec.cpp:188-201 */
void orc_vecset(const int npdim[3], double a, double *x)
{
	DIMS;
	FOR_REAL x[GI(i, j, k)] = a;
}

/* src/linalg/matvec.cpp:203-219 */
void orc_vecassign(const int npdim[3], const double *x, double *y)
{
	DIMS;
#pragma omp parallel for collapse(2)
	for (int k = 1; k <= nz; k++)
		for (int j = 1; j <= ny; j++)
			for (int i = 1; i <= nx; i++) y[GI(i, j, k)] = x[GI(i, j, k)];
}

/* src/linalg/matvec.cpp:148-164 */
void orc_vecaxpy(const int npdim[3], double a, const double *x, double *y)
{
	DIMS;
#pragma omp parallel for collapse(2)
	for (int k = 1; k <= nz; k++)
		for (int j = 1; j <= ny; j++)
			for (int i = 1; i <= nx; i++) y[GI(i, j, k)] += a * x[GI(i, j, k)];
}

/* src/linalg/matvec.cpp:166-186: one full pass per vector, in order */
void orc_vec_multi_axpy(const int npdim[3], int n, const double *a, const double *const *x, double *y)
{
	for (int l = 0; l < n; l++) orc_vecaxpy(npdim, a[l], x[l], y);
}

/* src/linalg/matvec.cpp:240-255 */
double orc_norm_vector_l2(const int npdim[3], const double *x)
{
	DIMS;
	double s = 0;
#pragma omp parallel for collapse(2) reduction(+ : s)
	for (int k = 1; k <= nz; k++)
		for (int j = 1; j <= ny; j++)
			for (int i = 1; i <= nx; i++) s += x[GI(i, j, k)] * x[GI(i, j, k)];
	return sqrt(s);
}

/* src/linalg/matvec.cpp:257-279 (no square root) */
double orc_inner_vector_l2(const int npdim[3], const double *x, const double *y)
{
	DIMS;
	double s = 0;
#pragma omp parallel for collapse(2) reduction(+ : s)
	for (int k = 1; k <= nz; k++)
		for (int j = 1; j <= ny; j++)
			for (int i = 1; i <= nx; i++) s += x[GI(i, j, k)] * y[GI(i, j, k)];
	return s;
}

/* src/linalg/matvec.cpp:221-238: dual-cell volume weights from the 1-D coordinates */
double orc_norm_L2(const int npdim[3], const double *cx, const double *cy, const double *cz, const double *x)
{
	DIMS;
	double s = 0;
#pragma omp parallel for collapse(2) reduction(+ : s)
	for (int k = 1; k <= nz; k++)
		for (int j = 1; j <= ny; j++)
			for (int i = 1; i <= nx; i++) {
				const double vol = 1.0 / 8.0 * (cx[i + 1] - cx[i - 1]) * (cy[j + 1] - cy[j - 1]) * (cz[k + 1] - cz[k - 1]);
				s += x[GI(i, j, k)] * x[GI(i, j, k)] * vol;
			}
	return sqrt(s);
}

/* src/linalg/matvec.cpp:281-288: diff = y - x over real points (ghosts copied from y), then norm_L2 */
double orc_compute_error_L2(const int npdim[3], const double *cx, const double *cy, const double *cz,
                            const double *x, const double *y)
{
	const size_t G = (size_t)npdim[0] * npdim[1] * npdim[2];
	double *diff = (double *)malloc(G * sizeof(double));
	memcpy(diff, y, G * sizeof(double));
	orc_vecaxpy(npdim, -1.0, x, diff);
	const double e = orc_norm_L2(npdim, cx, cy, cz, diff);
	free(diff);
	return e;
}

/* ------------------------------------------------------- preconditioners */

/* src/linalg/s3d_jacobi.cpp:14-29: a true division by the diagonal */
void orc_jacobi_apply(const int npdim[3], const double *A, const double *r, double *z)
{
	DIMS;
#pragma omp parallel for collapse(2)
	for (int k = 1; k <= nz; k++)
		for (int j = 1; j <= ny; j++)
			for (int i = 1; i <= nx; i++) z[GI(i, j, k)] = r[GI(i, j, k)] / A[3 * N + RI(i - 1, j - 1, k - 1)];
}

/* src/linalg/s3d_sgspreconditioners.cpp:122-135 */
void orc_sgs_setup(const int npdim[3], const double *A, double *diaginv)
{
	DIMS;
	for (i64 r = 0; r < N; r++) diaginv[r] = 1.0 / A[3 * N + r];
}

/* src/linalg/s3d_ilu.cpp:17-64, sequential build (threadedbuild = false) */
void orc_strilu_setup(const int npdim[3], const double *A, int nbuildsweeps, double *diaginv)
{
	DIMS;
	for (i64 r = 0; r < N; r++) diaginv[r] = A[3 * N + r];
	for (int swp = 0; swp < nbuildsweeps; swp++)
		for (int k = 0; k < nz; k++)
			for (int j = 0; j < ny; j++)
				for (int i = 0; i < nx; i++) {
					const i64 r = RI(i, j, k);
					diaginv[r] = A[3 * N + r];
					if (i > 0) diaginv[r] -= A[0 * N + r] * A[4 * N + RI(i - 1, j, k)] / diaginv[RI(i - 1, j, k)];
					if (j > 0) diaginv[r] -= A[1 * N + r] * A[5 * N + RI(i, j - 1, k)] / diaginv[RI(i, j - 1, k)];
					if (k > 0) diaginv[r] -= A[2 * N + r] * A[6 * N + RI(i, j, k - 1)] / diaginv[RI(i, j, k - 1)];
				}
	for (i64 r = 0; r < N; r++) diaginv[r] = 1.0 / diaginv[r];
}

/* src/linalg/s3d_sgspreconditioners.cpp:14-115 with threadedapply = false (the pinned oracle
 * configuration): zeroed scratch y; nsweeps forward lexicographic sweeps
 * y = (r - A0 y[i-1] - A1 y[j-1] - A2 y[k-1]) * dinv; nsweeps backward sweeps
 * z = y - dinv * (A4 z[i+1] + A5 z[j+1] + A6 z[k+1]). */
void orc_sgs_like_apply(const int npdim[3], const double *A, const double *diaginv, int nsweeps,
                        const double *r, double *z)
{
	DIMS;
	double *y = (double *)calloc((size_t)np0 * np1 * np2, sizeof(double));
	for (int swp = 0; swp < nsweeps; swp++)
		FOR_REAL {
			const i64 q = RI(i - 1, j - 1, k - 1);
			double t = r[GI(i, j, k)] - A[0 * N + q] * y[GI(i - 1, j, k)] - A[1 * N + q] * y[GI(i, j - 1, k)]
			         - A[2 * N + q] * y[GI(i, j, k - 1)];
			t *= diaginv[q];
			y[GI(i, j, k)] = t;
		}
	for (int swp = 0; swp < nsweeps; swp++)
		for (int k = nz; k >= 1; k--)
			for (int j = ny; j >= 1; j--)
				for (int i = nx; i >= 1; i--) {
					const i64 q = RI(i - 1, j - 1, k - 1);
					z[GI(i, j, k)] = y[GI(i, j, k)]
					    - diaginv[q] * (A[4 * N + q] * z[GI(i + 1, j, k)] + A[5 * N + q] * z[GI(i, j + 1, k)]
					                    + A[6 * N + q] * z[GI(i, j, k + 1)]);
				}
	free(y);
}

/* The two half-sweeps of the apply above on their own, reading whatever the GHOST entries of the output hold instead of zeros.
 * Not a separate function in the reference (s3d_sgspreconditioners.cpp:53-111 runs both inside apply on a zeroed scratch); split out
 * to state the z-slab pipeline on the CPU: a slab whose bottom ghost plane holds the forward result of the slab below (top ghost plane:
 * the backward result of the slab above) reproduces the global sweep restricted to the slab. */
void orc_sgs_forward_inplace(const int npdim[3], const double *A, const double *diaginv, const double *r, double *y)
{
	DIMS;
	FOR_REAL {
		const i64 q = RI(i - 1, j - 1, k - 1);
		double t = r[GI(i, j, k)] - A[0 * N + q] * y[GI(i - 1, j, k)] - A[1 * N + q] * y[GI(i, j - 1, k)]
		         - A[2 * N + q] * y[GI(i, j, k - 1)];
		t *= diaginv[q];
		y[GI(i, j, k)] = t;
	}
}

void orc_sgs_backward_inplace(const int npdim[3], const double *A, const double *diaginv, const double *y, double *z)
{
	DIMS;
	for (int k = nz; k >= 1; k--)
		for (int j = ny; j >= 1; j--)
			for (int i = nx; i >= 1; i--) {
				const i64 q = RI(i - 1, j - 1, k - 1);
				z[GI(i, j, k)] = y[GI(i, j, k)]
				    - diaginv[q] * (A[4 * N + q] * z[GI(i + 1, j, k)] + A[5 * N + q] * z[GI(i, j + 1, k)]
				                    + A[6 * N + q] * z[GI(i, j, k + 1)]);
			}
}

/* Synchronous multi-sweep variant of the same apply: every sweep recomputes ALL points from the previous sweep's values
 * (Jacobi iterations on the two triangular systems, starting from zero):
 *   forward  y <- dinv (r - A0 y[i-1] - A1 y[j-1] - A2 y[k-1]),   backward  z <- y - dinv (A4 z[i+1] + A5 z[j+1] + A6 z[k+1]).
 * This is the deterministic counterpart of the reference's ASYNCHRONOUS threaded sweeps (s3d_sgspreconditioners.cpp:32-112 with
 * threadedapply = true, napplysweeps = n), which update in place with no ordering between threads.  After nx+ny+nz-2 sweeps it
 * equals the sequential sweep exactly. */
void orc_sgs_jacobi_sweeps_apply(const int npdim[3], const double *A, const double *diaginv, int nsweeps,
                                 const double *r, double *z)
{
	DIMS;
	const size_t G = (size_t)np0 * np1 * np2;
	double *y0 = (double *)calloc(G, sizeof(double)), *y1 = (double *)calloc(G, sizeof(double));
	double *z0 = (double *)calloc(G, sizeof(double));
	for (int swp = 0; swp < nsweeps; swp++) {
		FOR_REAL {
			const i64 q = RI(i - 1, j - 1, k - 1);
			double t = r[GI(i, j, k)] - A[0 * N + q] * y0[GI(i - 1, j, k)] - A[1 * N + q] * y0[GI(i, j - 1, k)]
			         - A[2 * N + q] * y0[GI(i, j, k - 1)];
			y1[GI(i, j, k)] = t * diaginv[q];
		}
		double *tmp = y0; y0 = y1; y1 = tmp;
	}
	/* y0 holds the forward result; reuse y1 / z0 as the backward ping-pong pair */
	memset(y1, 0, G * sizeof(double));
	for (int swp = 0; swp < nsweeps; swp++) {
		FOR_REAL {
			const i64 q = RI(i - 1, j - 1, k - 1);
			z0[GI(i, j, k)] = y0[GI(i, j, k)]
			    - diaginv[q] * (A[4 * N + q] * y1[GI(i + 1, j, k)] + A[5 * N + q] * y1[GI(i, j + 1, k)]
			                    + A[6 * N + q] * y1[GI(i, j, k + 1)]);
		}
		double *tmp = z0; z0 = y1; y1 = tmp;
	}
	FOR_REAL z[GI(i, j, k)] = y1[GI(i, j, k)];
	free(y0); free(y1); free(z0);
}

/* NOT in the reference.  The same M = (D+L) D^-1 (D+U) product, but with the unknowns ordered
 * red ((i+j+k) even, 1-based) before black.  Every neighbour of a red point is black and vice
 * versa, so under that ordering L couples black rows to all six red neighbours and U couples red
 * rows to all six black neighbours:
 *   forward : y_R = dinv r_R ;  y_B = dinv (r_B - sum_6 A_s y_s)
 *   backward: z_B = y_B      ;  z_R = y_R - dinv sum_6 A_s z_s
 * Term order inside the sums follows the stencil order 0,1,2,4,5,6. */
void orc_sgs_redblack_apply(const int npdim[3], const double *A, const double *diaginv,
                            const double *r, double *z)
{
	DIMS;
	double *y = (double *)calloc((size_t)np0 * np1 * np2, sizeof(double));
	FOR_REAL if (((i + j + k) & 1) == 0) y[GI(i, j, k)] = r[GI(i, j, k)] * diaginv[RI(i - 1, j - 1, k - 1)];
	FOR_REAL if (((i + j + k) & 1) == 1) {
		const i64 q = RI(i - 1, j - 1, k - 1);
		double t = r[GI(i, j, k)] - A[0 * N + q] * y[GI(i - 1, j, k)] - A[1 * N + q] * y[GI(i, j - 1, k)]
		         - A[2 * N + q] * y[GI(i, j, k - 1)] - A[4 * N + q] * y[GI(i + 1, j, k)]
		         - A[5 * N + q] * y[GI(i, j + 1, k)] - A[6 * N + q] * y[GI(i, j, k + 1)];
		y[GI(i, j, k)] = t * diaginv[q];
	}
	FOR_REAL if (((i + j + k) & 1) == 1) z[GI(i, j, k)] = y[GI(i, j, k)];
	FOR_REAL if (((i + j + k) & 1) == 0) {
		const i64 q = RI(i - 1, j - 1, k - 1);
		z[GI(i, j, k)] = y[GI(i, j, k)]
		    - diaginv[q] * (A[0 * N + q] * z[GI(i - 1, j, k)] + A[1 * N + q] * z[GI(i, j - 1, k)]
		                    + A[2 * N + q] * z[GI(i, j, k - 1)] + A[4 * N + q] * z[GI(i + 1, j, k)]
		                    + A[5 * N + q] * z[GI(i, j + 1, k)] + A[6 * N + q] * z[GI(i, j, k + 1)]);
	}
	free(y);
}

/* --------------------------------------------------------------- solvers */

typedef struct {
	const int *npdim;
	const double *A;
	orc_pc_params p;
	double *diaginv; /* SGS / StrILU */
	size_t G;
} pc_ctx;

/* what createSolver + updateOperator do for the preconditioner: src/linalg/solverfactory.cpp:39-84 */
static pc_ctx pc_setup(const int npdim[3], const double *A, orc_pc_params p)
{
	pc_ctx c = {npdim, A, p, NULL, (size_t)npdim[0] * npdim[1] * npdim[2]};
	const size_t N = (size_t)(npdim[0] - 2) * (npdim[1] - 2) * (npdim[2] - 2);
	if (p.pc == ORC_PC_SGS) { c.diaginv = (double *)malloc(N * sizeof(double)); orc_sgs_setup(npdim, A, c.diaginv); }
	if (p.pc == ORC_PC_STRILU) { c.diaginv = (double *)malloc(N * sizeof(double)); orc_strilu_setup(npdim, A, p.nbuildsweeps, c.diaginv); }
	return c;
}
static void pc_free(pc_ctx *c) { free(c->diaginv); }

static void pc_apply(const pc_ctx *c, const double *r, double *z)
{
	switch (c->p.pc) {
	case ORC_PC_JACOBI: orc_jacobi_apply(c->npdim, c->A, r, z); break;
	case ORC_PC_SGS:
	case ORC_PC_STRILU: orc_sgs_like_apply(c->npdim, c->A, c->diaginv, c->p.napplysweeps, r, z); break;
	default: memcpy(z, r, c->G * sizeof(double)); /* NoSolver copies ghosts too: s3d_solverbase.cpp:21-33 */
	}
}

static double *zeros(size_t G) { return (double *)calloc(G, sizeof(double)); }

/* src/linalg/s3d_solverbase.cpp:45-80 */
orc_solve_info orc_richardson(const int npdim[3], const double *A, orc_pc_params pc, const double *b,
                              double *x, double rtol, int maxiter, double *hist)
{
	pc_ctx P = pc_setup(npdim, A, pc);
	double *res = zeros(P.G), *dx = zeros(P.G);
	const double bnorm = orc_norm_vector_l2(npdim, b);
	orc_apply_res(npdim, A, b, x, res);
	double resnorm = orc_norm_vector_l2(npdim, res);
	int step = 0;
	while (resnorm / bnorm > rtol && step < maxiter) {
		pc_apply(&P, res, dx);
		orc_vecaxpy(npdim, 1.0, dx, x);
		orc_apply_res(npdim, A, b, x, res);
		resnorm = orc_norm_vector_l2(npdim, res);
		if (hist) hist[step] = resnorm / bnorm;
		step++;
	}
	orc_solve_info info = {resnorm / bnorm <= rtol, step, resnorm};
	free(res); free(dx); pc_free(&P);
	return info;
}

/* src/linalg/s3d_gcr.cpp:14-86: restarted flexible GCR, classical Gram-Schmidt.  The inner loop
 * does not test maxiter, so up to maxiter + restart - 1 iterations (and hist entries) can occur. */
orc_solve_info orc_gcr(const int npdim[3], const double *A, orc_pc_params pc, const double *b, double *x,
                       double rtol, int maxiter, int restart, double *hist)
{
	pc_ctx P = pc_setup(npdim, A, pc);
	const int north = restart;
	double *res = zeros(P.G), *z = zeros(P.G);
	double **p = (double **)malloc(sizeof(double *) * north), **q = (double **)malloc(sizeof(double *) * north);
	for (int i = 0; i < north; i++) { p[i] = zeros(P.G); q[i] = zeros(P.G); }
	double *beta = (double *)malloc(sizeof(double) * (north + 1));
	const double bnorm = orc_norm_vector_l2(npdim, b);
	double resnorm = 1.0;
	int step = 0;
	while (step < maxiter) {
		orc_apply_res(npdim, A, b, x, res);
		resnorm = orc_norm_vector_l2(npdim, res);
		pc_apply(&P, res, p[0]);
		orc_apply(npdim, A, p[0], q[0]);
		for (int k = 0; k < north; k++) {
			const double alpha = orc_inner_vector_l2(npdim, res, q[k]) / orc_inner_vector_l2(npdim, q[k], q[k]);
			orc_vecaxpy(npdim, alpha, p[k], x);
			orc_vecaxpy(npdim, -alpha, q[k], res);
			resnorm = orc_norm_vector_l2(npdim, res);
			if (hist) hist[step] = resnorm / bnorm;
			step++;
			if (resnorm / bnorm < rtol) break;
			if (k == north - 1) break;
			pc_apply(&P, res, z);
			orc_apply(npdim, A, z, q[k + 1]);
			orc_vecassign(npdim, z, p[k + 1]);
			for (int i = 0; i < k + 1; i++)
				beta[i] = -orc_inner_vector_l2(npdim, q[k + 1], q[i]) / orc_inner_vector_l2(npdim, q[i], q[i]);
			orc_vec_multi_axpy(npdim, k + 1, beta, (const double *const *)p, p[k + 1]);
			orc_vec_multi_axpy(npdim, k + 1, beta, (const double *const *)q, q[k + 1]);
		}
		if (resnorm / bnorm < rtol) break;
	}
	orc_solve_info info = {resnorm / bnorm <= rtol, step, resnorm};
	for (int i = 0; i < north; i++) { free(p[i]); free(q[i]); }
	free(p); free(q); free(beta); free(res); free(z); pc_free(&P);
	return info;
}

/* Preconditioned CG.  NOT in the reference; same conventions as Richardson above
 * (loop while rel > rtol and step < maxiter; converged means rel <= rtol; recurrence residual). */
orc_solve_info orc_cg(const int npdim[3], const double *A, orc_pc_params pc, const double *b, double *x,
                      double rtol, int maxiter, double *hist)
{
	pc_ctx P = pc_setup(npdim, A, pc);
	double *r = zeros(P.G), *z = zeros(P.G), *p = zeros(P.G), *q = zeros(P.G);
	const double bnorm = orc_norm_vector_l2(npdim, b);
	orc_apply_res(npdim, A, b, x, r);
	double resnorm = orc_norm_vector_l2(npdim, r);
	pc_apply(&P, r, z);
	orc_vecassign(npdim, z, p);
	double rz = orc_inner_vector_l2(npdim, r, z);
	int step = 0;
	while (resnorm / bnorm > rtol && step < maxiter) {
		orc_apply(npdim, A, p, q);
		const double alpha = rz / orc_inner_vector_l2(npdim, p, q);
		orc_vecaxpy(npdim, alpha, p, x);
		orc_vecaxpy(npdim, -alpha, q, r);
		resnorm = orc_norm_vector_l2(npdim, r);
		if (hist) hist[step] = resnorm / bnorm;
		step++;
		if (!(resnorm / bnorm > rtol)) break;
		pc_apply(&P, r, z);
		const double rznew = orc_inner_vector_l2(npdim, r, z);
		const double beta = rznew / rz;
		rz = rznew;
		orc_vecassign(npdim, z, q);      /* p <- z + beta p via the reference primitives */
		orc_vecaxpy(npdim, beta, p, q);
		orc_vecassign(npdim, q, p);
	}
	orc_solve_info info = {resnorm / bnorm <= rtol, step, resnorm};
	free(r); free(z); free(p); free(q); pc_free(&P);
	return info;
}

/* Right-preconditioned BiCGSTAB.  NOT in the reference; one test per iteration on the recurrence
 * residual; breakdown (rho = 0, <rhat,v> = 0, omega = 0) stops with converged = 0. */
orc_solve_info orc_bicgstab(const int npdim[3], const double *A, orc_pc_params pc, const double *b,
                            double *x, double rtol, int maxiter, double *hist)
{
	pc_ctx P = pc_setup(npdim, A, pc);
	double *r = zeros(P.G), *rhat = zeros(P.G), *p = zeros(P.G), *v = zeros(P.G), *s = zeros(P.G),
	       *t = zeros(P.G), *ph = zeros(P.G), *sh = zeros(P.G), *tmp = zeros(P.G);
	const double bnorm = orc_norm_vector_l2(npdim, b);
	orc_apply_res(npdim, A, b, x, r);
	orc_vecassign(npdim, r, rhat);
	double resnorm = orc_norm_vector_l2(npdim, r);
	double rho = 1, alpha = 1, omega = 1;
	int step = 0, breakdown = 0;
	while (resnorm / bnorm > rtol && step < maxiter) {
		const double rhonew = orc_inner_vector_l2(npdim, rhat, r);
		if (rhonew == 0) { breakdown = 1; break; }
		if (step == 0) orc_vecassign(npdim, r, p);
		else {
			const double beta = (rhonew / rho) * (alpha / omega);
			orc_vecaxpy(npdim, -omega, v, p);
			orc_vecassign(npdim, r, tmp);
			orc_vecaxpy(npdim, beta, p, tmp);
			orc_vecassign(npdim, tmp, p);
		}
		rho = rhonew;
		pc_apply(&P, p, ph);
		orc_apply(npdim, A, ph, v);
		const double rv = orc_inner_vector_l2(npdim, rhat, v);
		if (rv == 0) { breakdown = 1; break; }
		alpha = rho / rv;
		orc_vecassign(npdim, r, s);
		orc_vecaxpy(npdim, -alpha, v, s);
		pc_apply(&P, s, sh);
		orc_apply(npdim, A, sh, t);
		const double tt = orc_inner_vector_l2(npdim, t, t);
		if (tt == 0) {
			orc_vecaxpy(npdim, alpha, ph, x);
			orc_vecassign(npdim, s, r);
			resnorm = orc_norm_vector_l2(npdim, r);
			if (hist) hist[step] = resnorm / bnorm;
			step++;
			break;
		}
		omega = orc_inner_vector_l2(npdim, t, s) / tt;
		orc_vecaxpy(npdim, alpha, ph, x);
		orc_vecaxpy(npdim, omega, sh, x);
		orc_vecassign(npdim, s, r);
		orc_vecaxpy(npdim, -omega, t, r);
		resnorm = orc_norm_vector_l2(npdim, r);
		if (hist) hist[step] = resnorm / bnorm;
		step++;
		if (omega == 0) { breakdown = 1; break; }
	}
	orc_solve_info info = {!breakdown && resnorm / bnorm <= rtol, step, resnorm};
	free(r); free(rhat); free(p); free(v); free(s); free(t); free(ph); free(sh); free(tmp); pc_free(&P);
	return info;
}
