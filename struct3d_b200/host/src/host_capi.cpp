// Flat C view of the host classes (include/s3d_host.h).  One forwarding call per function.
#include <vector>
#include "case.hpp"
#include "common_utils.hpp"
#include "linalg/s3d_ilu.hpp"
#include "linalg/s3d_jacobi.hpp"
#include "linalg/s3d_sgspreconditioners.hpp"
#include "linalg/solverfactory.hpp"
#include "pde/pdefactory.hpp"
#include "s3d_host.h"

static char g_err[512] = "";

#define TRY try {
#define CATCH(ret) } catch (const std::exception &e) { std::snprintf(g_err, sizeof g_err, "%s", e.what()); return ret; }
#define V(p) (*static_cast<SVec *>(p))
#define M(p) (*static_cast<SMat *>(p))

extern "C" {

const char *s3dh_last_error(void) { return g_err; }

static bool g_init = false;
int s3dh_init(int device, int rank, int nranks, const void *nccl_id)
{
	TRY
	if (device < 0) {
		const char *lr = std::getenv("LOCAL_RANK");
		device = lr ? std::atoi(lr) : 0;
	}
	s3d::Device::initialize(device, rank, nranks, nccl_id);
	g_init = true;
	return 0;
	CATCH(-1)
}
int s3dh_initialized(void) { return g_init; }
void s3dh_shutdown(void) { s3d::Device::shutdown(); g_init = false; }
void *s3dh_cuda_ctx(void)
{
	TRY return s3d::Device::get().ctx(); CATCH(nullptr)
}

void s3dh_options_clear(void) { s3d::options_clear(); }
void s3dh_options_set(const char *key, const char *value) { s3d::options_set(key, value ? value : ""); }
int s3dh_options_load(const char *path)
{
	TRY s3d::options_load_file(path); return 0; CATCH(-1)
}

void *s3dh_mesh_create(const int npdim[3], int chebyshev, const double rmin[3], const double rmax[3])
{
	TRY
	CartMesh *m = new CartMesh();
	try {
		m->createMesh(npdim);
		if (chebyshev) m->generateMesh_ChebyshevDistribution(rmin, rmax);
		else m->generateMesh_UniformDistribution(rmin, rmax);
	} catch (...) { delete m; throw; }
	return m;
	CATCH(nullptr)
}
void s3dh_mesh_destroy(void *mesh) { delete static_cast<CartMesh *>(mesh); }
void s3dh_mesh_coords(void *mesh, int dim, double *out)
{
	const CartMesh *m = static_cast<CartMesh *>(mesh);
	for (int i = 0; i < m->gnpoind(dim); i++) out[i] = m->gcoords(dim, i);
}
double s3dh_mesh_h(void *mesh) { return static_cast<CartMesh *>(mesh)->gh(); }
void s3dh_mesh_local(void *mesh, int *nx, int *ny, int *nz_local, int *k0)
{
	const s3d_grid &g = static_cast<CartMesh *>(mesh)->device_grid();
	*nx = g.nx; *ny = g.ny; *nz_local = g.nz; *k0 = g.k0;
}

int s3dh_read_control(const char *path, int npdim[3], double rmin[3], double rmax[3], int *chebyshev,
                      char pdetype[32], double vel[3], double *mu, int *nruns, int btypes[6], double bvals[6])
{
	TRY
	FILE *fp = std::fopen(path, "r");
	if (!fp) throw std::runtime_error("cannot open control file");
	const CaseData c = readCtrl(fp);
	std::fclose(fp);
	for (int i = 0; i < 3; i++) { npdim[i] = c.npdim[i]; rmin[i] = c.rmin[i]; rmax[i] = c.rmax[i]; vel[i] = c.vel[i]; }
	*chebyshev = c.gridtype == S3D_CHEBYSHEV;
	std::snprintf(pdetype, 32, "%s", c.pdetype.c_str());
	*mu = c.diffcoeff;
	*nruns = c.nruns;
	for (int i = 0; i < 6; i++) { btypes[i] = c.btypes[i] == S3D_DIRICHLET; bvals[i] = c.bvals[i]; }
	return 0;
	CATCH(-1)
}

void *s3dh_pde_create(const char *type, const double vel[3], double mu, const int dirichlet[6], const double bvals[6])
{
	TRY
	CaseData c;
	c.pdetype = type;
	if (c.pdetype != "poisson" && c.pdetype != "convdiff" && c.pdetype != "convdiff_circular") throw std::runtime_error("PDE type not recognized!");
	for (int i = 0; i < 3; i++) c.vel[i] = vel[i];
	c.diffcoeff = mu;
	for (int i = 0; i < 6; i++) { c.btypes[i] = dirichlet[i] ? S3D_DIRICHLET : S3D_EXTRAPOLATION; c.bvals[i] = bvals[i]; }
	return construct_pde(c);
	CATCH(nullptr)
}
void s3dh_pde_destroy(void *pde) { delete static_cast<PDEBase *>(pde); }

static GridFunction pick(const PDEBase *p, int which)
{
	if (which == 0) return p->test_rhs();
	return p->manufactured_solution()[which == 1 ? 0 : 1];
}
void *s3dh_pde_vector(void *pde, void *mesh, int which)
{
	TRY
	const PDEBase *p = static_cast<PDEBase *>(pde);
	return new SVec(p->computeVector(static_cast<CartMesh *>(mesh), pick(p, which)));
	CATCH(nullptr)
}
void *s3dh_pde_vector_sampled(void *pde, void *mesh, int which)
{
	TRY
	const PDEBase *p = static_cast<PDEBase *>(pde);
	const std::function<sreal(const sreal[NDIM])> f = pick(p, which);   // drops the separable description
	return new SVec(p->computeVector(static_cast<CartMesh *>(mesh), GridFunction(f)));
	CATCH(nullptr)
}
void *s3dh_pde_lhs(void *pde, void *mesh)
{
	TRY return new SMat(static_cast<PDEBase *>(pde)->computeLHS(static_cast<CartMesh *>(mesh))); CATCH(nullptr)
}

void *s3dh_vec_create(void *mesh)
{
	TRY return new SVec(static_cast<CartMesh *>(mesh)); CATCH(nullptr)
}
void *s3dh_vec_clone(void *vec)
{
	TRY return new SVec(V(vec)); CATCH(nullptr)
}
void s3dh_vec_destroy(void *vec) { delete static_cast<SVec *>(vec); }
long long s3dh_vec_len(void *vec) { return (long long)V(vec).vals.size(); }
int s3dh_vec_get(void *vec, double *out)
{
	TRY V(vec).vals.copy_to(out); return 0; CATCH(-1)
}
int s3dh_vec_set(void *vec, const double *in)
{
	TRY V(vec).vals.assign(in); return 0; CATCH(-1)
}
double s3dh_vec_peek(void *vec, long long idx)
{
	TRY return static_cast<const SVec *>(vec)->vals[(size_t)idx]; CATCH(0.0)
}
int s3dh_vec_poke(void *vec, long long idx, double v)
{
	TRY V(vec).vals[(size_t)idx] = v; return 0; CATCH(-1)
}
void *s3dh_vec_device_ptr(void *vec)
{
	TRY return V(vec).dev_rw(); CATCH(nullptr)
}
void *s3dh_mat_create(void *mesh)
{
	TRY return new SMat(static_cast<CartMesh *>(mesh)); CATCH(nullptr)
}
void s3dh_mat_destroy(void *mat) { delete static_cast<SMat *>(mat); }
long long s3dh_mat_len(void *mat) { return (long long)M(mat).vals[0].size(); }
int s3dh_mat_get(void *mat, int s, double *out)
{
	TRY M(mat).vals[s].copy_to(out); return 0; CATCH(-1)
}
int s3dh_mat_set(void *mat, int s, const double *in)
{
	TRY M(mat).vals[s].assign(in); return 0; CATCH(-1)
}
void *s3dh_mat_device_ptr(void *mat)
{
	TRY return M(mat).dev_w(); CATCH(nullptr)
}

int s3dh_apply(void *A, void *x, void *y) { TRY M(A).apply(V(x), V(y)); return 0; CATCH(-1) }
int s3dh_apply_res(void *A, void *b, void *x, void *y) { TRY M(A).apply_res(V(b), V(x), V(y)); return 0; CATCH(-1) }
int s3dh_vecset(double a, void *x) { TRY vecset(a, V(x)); return 0; CATCH(-1) }
int s3dh_vecassign(void *x, void *y) { TRY vecassign(V(x), V(y)); return 0; CATCH(-1) }
int s3dh_vecaxpy(double a, void *x, void *y) { TRY vecaxpy(a, V(x), V(y)); return 0; CATCH(-1) }
int s3dh_vec_multi_axpy(int n, const double *a, void **xs, void *y)
{
	TRY
	std::vector<SVec> x;
	x.reserve(n);
	for (int l = 0; l < n; l++) x.emplace_back(V(xs[l]));   // device copies; the reference API takes an array of SVec
	vec_multi_axpy(n, a, x.data(), V(y));
	return 0;
	CATCH(-1)
}
int s3dh_norm_L2(void *x, double *out) { TRY *out = norm_L2(V(x)); return 0; CATCH(-1) }
int s3dh_norm_vector_l2(void *x, double *out) { TRY *out = norm_vector_l2(V(x)); return 0; CATCH(-1) }
int s3dh_inner_vector_l2(void *x, void *y, double *out) { TRY *out = inner_vector_l2(V(x), V(y)); return 0; CATCH(-1) }
int s3dh_compute_error_L2(void *x, void *y, double *out) { TRY *out = compute_error_L2(V(x), V(y)); return 0; CATCH(-1) }

void *s3dh_prec_create(void *A, const char *pc, int applysweeps, int buildsweeps, int threaded_apply, int threaded_build, int chunk)
{
	TRY
	const std::string s(pc);
	PreconParams pp;
	pp.nbuildsweeps = buildsweeps; pp.napplysweeps = applysweeps; pp.thread_chunk_size = chunk;
	pp.threadedbuild = threaded_build; pp.threadedapply = threaded_apply;
	SolverBase *p = nullptr;
	if (s == "jacobi") p = new JacobiPreconditioner(M(A));
	else if (s == "sgs") p = new SGS_preconditioner(M(A), pp);
	else if (s == "strilu") p = new StrILU_preconditioner(M(A), pp);
	else if (s == "none") p = new NoSolver(M(A));
	else throw std::runtime_error("unknown preconditioner");
	p->updateOperator();
	return p;
	CATCH(nullptr)
}
void s3dh_prec_destroy(void *prec) { delete static_cast<SolverBase *>(prec); }
int s3dh_prec_apply(void *prec, void *r, void *z)
{
	TRY static_cast<SolverBase *>(prec)->apply(V(r), V(z)); return 0; CATCH(-1)
}

void *s3dh_solver_create(void *A)
{
	TRY return createSolver(M(A)); CATCH(nullptr)
}
int s3dh_solver_update(void *solver)
{
	TRY static_cast<SolverBase *>(solver)->updateOperator(); return 0; CATCH(-1)
}
int s3dh_solver_apply(void *solver, void *b, void *x, int *converged, int *iters, double *resnorm, double *pctime)
{
	TRY
	const SolveInfo info = static_cast<SolverBase *>(solver)->apply(V(b), V(x));
	std::fflush(stdout);
	*converged = info.converged; *iters = info.iters; *resnorm = info.resnorm; *pctime = info.precapplywtime;
	return 0;
	CATCH(-1)
}
int s3dh_solver_history(void *solver, double *out, int cap)
{
	const std::vector<sreal> &h = static_cast<SolverBase *>(solver)->residual_history();
	const int n = std::min((int)h.size(), cap);
	for (int i = 0; i < n; i++) out[i] = h[i];
	return (int)h.size();
}
void s3dh_solver_destroy(void *solver) { delete static_cast<SolverBase *>(solver); }

}  // extern "C"
