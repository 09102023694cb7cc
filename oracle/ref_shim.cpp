/* TEST INFRASTRUCTURE ONLY -- never linked into the product.
 *
 * C-callable shim over the UNMODIFIED reference classes (compiled from /root/reference/src by
 * oracle/Makefile into oracle/_ref/libs3dref.so).  Used by
 *   - oracle/make_golden.py   to generate tests/golden/ fixtures,
 *   - tests/                  to pin oracle/s3d_oracle.c against the real reference,
 *   - bench.py --impl reference / cpu_baseline  to time the reference's own CPU path.
 *
 * Every function here only forwards to reference symbols:
 *   SVec / SMat / free functions            src/linalg/matvec.hpp:23-115
 *   createSolver                            src/linalg/solverfactory.hpp:13
 *   Jacobi / SGS / StrILU preconditioners   src/linalg/s3d_jacobi.hpp, s3d_sgspreconditioners.hpp, s3d_ilu.hpp
 *   construct_pde / PDEBase                 src/pde/pdefactory.hpp:11, src/pde/pdebase.hpp:14-62
 *   CartMesh                                src/cartmesh.hpp:67-154
 * The CG and BiCGSTAB loops at the bottom are NOT in the reference; they
 * are harness loops assembled only from reference primitives (SMat::apply, inner_vector_l2,
 * vecaxpy, vecassign, norm_vector_l2, the preconditioners' apply), as SURVEY.md 8(c) prescribes.
 */
#include <vector>
#include <string>
#include <cstring>
#include <omp.h>

#include "cartmesh.hpp"
#include "case.hpp"
#include "common_utils.hpp"
#include "linalg/matvec.hpp"
#include "linalg/s3d_solverbase.hpp"
#include "linalg/s3d_jacobi.hpp"
#include "linalg/s3d_sgspreconditioners.hpp"
#include "linalg/s3d_ilu.hpp"
#include "linalg/s3d_gcr.hpp"
#include "linalg/solverfactory.hpp"
#include "pde/pdefactory.hpp"

#define REFS_TRY try {
#define REFS_CATCH } catch (const std::exception &e) { std::snprintf(g_err, sizeof g_err, "%s", e.what()); return -1; } return 0;

static char g_err[512] = "";

extern "C" {

const char *refs_last_error() { return g_err; }
int refs_num_threads() { return omp_get_max_threads(); }
void refs_set_num_threads(int n) { omp_set_num_threads(n); }

/* ---- options database (stub-backed) ---- */
void refs_options_clear() { s3d_stub_options_clear(); }
void refs_options_set(const char *key, const char *val) { PetscOptionsSetValue(NULL, key, val); }
int refs_options_load(const char *path)
{
	REFS_TRY
	int argc = 3;
	char a0[] = "refs", a1[] = "-options_file";
	std::string p(path);
	char *argv[] = {a0, a1, &p[0], nullptr};
	char **av = argv;
	PetscInitialize(&argc, &av, NULL, NULL);
	REFS_CATCH
}

/* ---- mesh ---- */
void *refs_mesh_create(const int npdim[3], int chebyshev, const double rmin[3], const double rmax[3])
{
	try {
		CartMesh *m = new CartMesh();
		m->createMesh(npdim);
		if (chebyshev) m->generateMesh_ChebyshevDistribution(rmin, rmax);
		else           m->generateMesh_UniformDistribution(rmin, rmax);
		return m;
	} catch (const std::exception &e) { std::snprintf(g_err, sizeof g_err, "%s", e.what()); return nullptr; }
}
void refs_mesh_destroy(void *m) { delete static_cast<CartMesh *>(m); }
void refs_mesh_coords(void *m, int dim, double *out)
{
	const CartMesh *mm = static_cast<CartMesh *>(m);
	for (int i = 0; i < mm->gnpoind(dim); i++) out[i] = mm->gcoords(dim, i);
}
double refs_mesh_h(void *m) { return static_cast<CartMesh *>(m)->gh(); }

/* reads a control file with the reference's reader (src/case.cpp:8-88) and reports what it parsed */
int refs_read_control(const char *path, int npdim[3], double rmin[3], double rmax[3], int *chebyshev,
                      char pdetype[32], double vel[3], double *mu, int *nruns, int btypes[6], double bvals[6])
{
	REFS_TRY
	FILE *fp = std::fopen(path, "r");
	if (!fp) throw std::runtime_error("cannot open control file");
	const CaseData c = readCtrl(fp);
	std::fclose(fp);
	for (int i = 0; i < 3; i++) { npdim[i] = c.npdim[i]; rmin[i] = c.rmin[i]; rmax[i] = c.rmax[i]; }
	*chebyshev = c.gridtype == S3D_CHEBYSHEV;
	std::snprintf(pdetype, 32, "%s", c.pdetype.c_str());
	vel[0] = vel[1] = vel[2] = 0; *mu = 0;
	if (c.pdetype == "convdiff") { for (int i = 0; i < 3; i++) vel[i] = c.vel[i]; *mu = c.diffcoeff; }
	if (c.pdetype == "convdiff_circular") { vel[0] = c.vel[0]; *mu = c.diffcoeff; }
	*nruns = c.nruns;
	for (int i = 0; i < 6; i++) { btypes[i] = c.btypes[i] == S3D_DIRICHLET; bvals[i] = c.bvals[i]; }
	REFS_CATCH
}

/* ---- pde ---- */
void *refs_pde_create(const char *type, const double vel[3], double mu, const int dirichlet[6], const double bvals[6])
{
	CaseData c;
	c.pdetype = type;
	for (int i = 0; i < 3; i++) c.vel[i] = vel[i];
	c.diffcoeff = mu;
	for (int i = 0; i < 6; i++) { c.btypes[i] = dirichlet[i] ? S3D_DIRICHLET : S3D_EXTRAPOLATION; c.bvals[i] = bvals[i]; }
	return construct_pde(c);
}
void refs_pde_destroy(void *p) { delete static_cast<PDEBase *>(p); }

/* ---- vectors and matrices (handles own the reference objects; numpy wraps the raw storage) ---- */
void *refs_vec_create(void *mesh) { return new SVec(static_cast<CartMesh *>(mesh)); }
void refs_vec_destroy(void *v) { delete static_cast<SVec *>(v); }
double *refs_vec_data(void *v) { return static_cast<SVec *>(v)->vals.data(); }
long long refs_vec_len(void *v) { return (long long)static_cast<SVec *>(v)->vals.size(); }

void *refs_mat_create(void *mesh) { return new SMat(static_cast<CartMesh *>(mesh)); }
void refs_mat_destroy(void *a) { delete static_cast<SMat *>(a); }
double *refs_mat_data(void *a, int s) { return static_cast<SMat *>(a)->vals[s].data(); }
long long refs_mat_len(void *a) { return (long long)static_cast<SMat *>(a)->vals[0].size(); }

/* which: 0 test_rhs, 1 manufactured solution, 2 manufactured rhs */
void *refs_pde_vector(void *pde, void *mesh, int which)
{
	try {
		const PDEBase *p = static_cast<PDEBase *>(pde);
		const CartMesh *m = static_cast<CartMesh *>(mesh);
		if (which == 0) return new SVec(p->computeVector(m, p->test_rhs()));
		return new SVec(p->computeVector(m, p->manufactured_solution()[which == 1 ? 0 : 1]));
	} catch (const std::exception &e) { std::snprintf(g_err, sizeof g_err, "%s", e.what()); return nullptr; }
}
void *refs_pde_lhs(void *pde, void *mesh)
{
	try {
		return new SMat(static_cast<PDEBase *>(pde)->computeLHS(static_cast<CartMesh *>(mesh)));
	} catch (const std::exception &e) { std::snprintf(g_err, sizeof g_err, "%s", e.what()); return nullptr; }
}

#define V(p) (*static_cast<SVec *>(p))
#define M(p) (*static_cast<SMat *>(p))

int refs_apply(void *A, void *x, void *y) { REFS_TRY M(A).apply(V(x), V(y)); REFS_CATCH }
int refs_apply_res(void *A, void *b, void *x, void *y) { REFS_TRY M(A).apply_res(V(b), V(x), V(y)); REFS_CATCH }
int refs_vecset(double a, void *x) { REFS_TRY vecset(a, V(x)); REFS_CATCH }
int refs_vecassign(void *x, void *y) { REFS_TRY vecassign(V(x), V(y)); REFS_CATCH }
int refs_vecaxpy(double a, void *x, void *y) { REFS_TRY vecaxpy(a, V(x), V(y)); REFS_CATCH }
int refs_vec_multi_axpy(int n, const double *a, void **xs, void *y)
{
	REFS_TRY
	std::vector<SVec> x(n);
	for (int l = 0; l < n; l++) { x[l].init(V(xs[l]).m); x[l].vals = V(xs[l]).vals; }
	vec_multi_axpy(n, a, x.data(), V(y));
	REFS_CATCH
}
double refs_norm_L2(void *x) { return norm_L2(V(x)); }
double refs_norm_vector_l2(void *x) { return norm_vector_l2(V(x)); }
double refs_inner_vector_l2(void *x, void *y) { return inner_vector_l2(V(x), V(y)); }
double refs_compute_error_L2(void *x, void *y) { return compute_error_L2(V(x), V(y)); }

/* ---- preconditioners ---- */
static SolverBase *make_prec(const SMat &A, const char *pc, int applysweeps, int buildsweeps,
                             int threaded_apply, int threaded_build, int chunk)
{
	const std::string s(pc);
	PreconParams pp;
	pp.nbuildsweeps = buildsweeps; pp.napplysweeps = applysweeps; pp.thread_chunk_size = chunk;
	pp.threadedbuild = threaded_build; pp.threadedapply = threaded_apply;
	if (s == "jacobi") return new JacobiPreconditioner(A);
	if (s == "sgs") return new SGS_preconditioner(A, pp);
	if (s == "strilu") return new StrILU_preconditioner(A, pp);
	if (s == "none") return new NoSolver(A);
	throw std::runtime_error("refs: unknown preconditioner");
}

void *refs_prec_create(void *A, const char *pc, int applysweeps, int buildsweeps, int threaded_apply,
                       int threaded_build, int chunk)
{
	try {
		SolverBase *p = make_prec(M(A), pc, applysweeps, buildsweeps, threaded_apply, threaded_build, chunk);
		p->updateOperator();
		return p;
	} catch (const std::exception &e) { std::snprintf(g_err, sizeof g_err, "%s", e.what()); return nullptr; }
}
void refs_prec_destroy(void *p) { delete static_cast<SolverBase *>(p); }
int refs_prec_apply(void *p, void *r, void *z) { REFS_TRY static_cast<SolverBase *>(p)->apply(V(r), V(z)); REFS_CATCH }

/* ---- the reference's own solve path: createSolver -> updateOperator -> apply (runcase(), scaling_openmp.cpp:84-140) ---- */
int refs_solve(void *A, void *b, void *x, int *converged, int *iters, double *resnorm, double *pctime,
               double *buildtime, double *solvetime)
{
	REFS_TRY
	SolverBase *solver = createSolver(M(A));
	double t0 = MPI_Wtime();
	solver->updateOperator();
	*buildtime = MPI_Wtime() - t0;
	t0 = MPI_Wtime();
	const SolveInfo info = solver->apply(V(b), V(x));
	*solvetime = MPI_Wtime() - t0;
	std::fflush(stdout);
	delete solver;
	*converged = info.converged; *iters = info.iters; *resnorm = info.resnorm; *pctime = info.precapplywtime;
	REFS_CATCH
}

/* ---- harness Krylov loops from reference primitives only (NOT reference code) ----
 * Conventions follow the native solvers (s3d_solverbase.cpp:45-80): zero-or-given initial guess,
 * loop while ||r||/||b|| > rtol and step < maxiter, converged flag is <= rtol, hist[step] is the
 * relative residual after step+1 iterations.  hist may be NULL; at most maxiter entries are written. */
int refs_harness_cg(void *Ap, void *prec, void *bp, void *xp, double rtol, int maxiter, int *converged,
                    int *iters, double *resnorm_out, double *hist)
{
	REFS_TRY
	const SMat &A = M(Ap); const SVec &b = V(bp); SVec &x = V(xp);
	SolverBase *P = static_cast<SolverBase *>(prec);
	SVec r(A.m), z(A.m), p(A.m), q(A.m);
	const sreal bnorm = norm_vector_l2(b);
	A.apply_res(b, x, r);
	sreal resnorm = norm_vector_l2(r);
	P->apply(r, z);
	vecassign(z, p);
	sreal rz = inner_vector_l2(r, z);
	int step = 0;
	while (resnorm / bnorm > rtol && step < maxiter) {
		A.apply(p, q);
		const sreal alpha = rz / inner_vector_l2(p, q);
		vecaxpy(alpha, p, x);
		vecaxpy(-alpha, q, r);
		resnorm = norm_vector_l2(r);
		if (hist) hist[step] = resnorm / bnorm;
		step++;
		if (!(resnorm / bnorm > rtol)) break;
		P->apply(r, z);
		const sreal rznew = inner_vector_l2(r, z);
		const sreal beta = rznew / rz;
		rz = rznew;
		/* p <- z + beta p, written with reference primitives: q is free here, use it as scratch */
		vecassign(z, q);
		vecaxpy(beta, p, q);
		vecassign(q, p);
	}
	*converged = resnorm / bnorm <= rtol; *iters = step; *resnorm_out = resnorm;
	REFS_CATCH
}

/* right-preconditioned BiCGSTAB, one convergence test per iteration on the recurrence residual */
int refs_harness_bicgstab(void *Ap, void *prec, void *bp, void *xp, double rtol, int maxiter, int *converged,
                          int *iters, double *resnorm_out, double *hist)
{
	REFS_TRY
	const SMat &A = M(Ap); const SVec &b = V(bp); SVec &x = V(xp);
	SolverBase *P = static_cast<SolverBase *>(prec);
	SVec r(A.m), rhat(A.m), p(A.m), v(A.m), s(A.m), t(A.m), ph(A.m), sh(A.m), tmp(A.m);
	const sreal bnorm = norm_vector_l2(b);
	A.apply_res(b, x, r);
	vecassign(r, rhat);
	sreal resnorm = norm_vector_l2(r);
	sreal rho = 1, alpha = 1, omega = 1;
	int step = 0;
	bool breakdown = false;
	while (resnorm / bnorm > rtol && step < maxiter) {
		const sreal rhonew = inner_vector_l2(rhat, r);
		if (rhonew == 0) { breakdown = true; break; }
		if (step == 0) vecassign(r, p);
		else {
			const sreal beta = (rhonew / rho) * (alpha / omega);
			/* p <- r + beta (p - omega v) */
			vecaxpy(-omega, v, p);
			vecassign(r, tmp);
			vecaxpy(beta, p, tmp);
			vecassign(tmp, p);
		}
		rho = rhonew;
		P->apply(p, ph);
		A.apply(ph, v);
		const sreal rv = inner_vector_l2(rhat, v);
		if (rv == 0) { breakdown = true; break; }
		alpha = rho / rv;
		vecassign(r, s);
		vecaxpy(-alpha, v, s);
		P->apply(s, sh);
		A.apply(sh, t);
		const sreal tt = inner_vector_l2(t, t);
		if (tt == 0) { /* s is already the exact update */
			vecaxpy(alpha, ph, x);
			vecassign(s, r);
			resnorm = norm_vector_l2(r);
			if (hist) hist[step] = resnorm / bnorm;
			step++;
			break;
		}
		omega = inner_vector_l2(t, s) / tt;
		vecaxpy(alpha, ph, x);
		vecaxpy(omega, sh, x);
		vecassign(s, r);
		vecaxpy(-omega, t, r);
		resnorm = norm_vector_l2(r);
		if (hist) hist[step] = resnorm / bnorm;
		step++;
		if (omega == 0) { breakdown = true; break; }
	}
	*converged = !breakdown && resnorm / bnorm <= rtol; *iters = step; *resnorm_out = resnorm;
	REFS_CATCH
}

/* ---- timing loops for the CPU baseline (best of reps after one warm-up), seconds ---- */
double refs_time_apply(void *A, void *x, void *y, int reps)
{
	M(A).apply(V(x), V(y));
	double best = 1e300;
	for (int r = 0; r < reps; r++) {
		const double t0 = MPI_Wtime();
		M(A).apply(V(x), V(y));
		const double t = MPI_Wtime() - t0;
		if (t < best) best = t;
	}
	return best;
}
double refs_time_axpy(void *x, void *y, int reps)
{
	vecaxpy(1e-3, V(x), V(y));
	double best = 1e300;
	for (int r = 0; r < reps; r++) {
		const double t0 = MPI_Wtime();
		vecaxpy(1e-3, V(x), V(y));
		const double t = MPI_Wtime() - t0;
		if (t < best) best = t;
	}
	return best;
}
double refs_time_dot(void *x, void *y, int reps, double *sink)
{
	double acc = inner_vector_l2(V(x), V(y));
	double best = 1e300;
	for (int r = 0; r < reps; r++) {
		const double t0 = MPI_Wtime();
		acc += inner_vector_l2(V(x), V(y));
		const double t = MPI_Wtime() - t0;
		if (t < best) best = t;
	}
	*sink = acc;
	return best;
}

} /* extern "C" */
