/* s3d_host.h -- flat C view of the C++ host classes in struct3d_b200/host (libstruct3d_host.so).
 *
 * The drop-in boundary of this path is a C++ class interface, not an FFI: callers of
 * the reference use SVec / SMat / createSolver / PDEBase directly, and struct3d_b200/host/include
 * provides exactly those headers.  This file only exists so that non-C++ callers (the Python tests,
 * bench.py, __graft_entry__.smoke) can drive the SAME host classes: every function forwards to one
 * C++ call and converts exceptions to a status code.  It mirrors, function for function, the shim
 * the tests use to drive the reference itself (oracle/ref_shim.cpp), so parity tests read the same
 * against both.  Handles are opaque pointers to the C++ objects; host arrays are in the reference's
 * host layouts (include/s3d_cuda.h).
 *
 * Returns: 0 on success, -1 on a C++ exception (message via s3dh_last_error), unless noted.
 */
#ifndef S3D_HOST_H
#define S3D_HOST_H

#ifdef __cplusplus
extern "C" {
#endif

const char *s3dh_last_error(void);

/* s3d::Device::initialize: device < 0 -> $LOCAL_RANK; nranks > 1 needs the S3D_COMM_ID_BYTES NCCL id of rank 0 */
int   s3dh_init(int device, int rank, int nranks, const void *nccl_id);
int   s3dh_initialized(void);
void  s3dh_shutdown(void);
void *s3dh_cuda_ctx(void);                       /* the underlying s3d_ctx* (for timers / launch counts) */

/* options store: petscoptions_* keys (src/common_utils.hpp:36-48) */
void  s3dh_options_clear(void);
void  s3dh_options_set(const char *key, const char *value);
int   s3dh_options_load(const char *path);

/* CartMesh (src/cartmesh.hpp:67-154) */
void *s3dh_mesh_create(const int npdim[3], int chebyshev, const double rmin[3], const double rmax[3]);
void  s3dh_mesh_destroy(void *mesh);
void  s3dh_mesh_coords(void *mesh, int dim, double *out);
double s3dh_mesh_h(void *mesh);
void  s3dh_mesh_local(void *mesh, int *nx, int *ny, int *nz_local, int *k0);

/* readCtrl (src/case.hpp:34) */
int   s3dh_read_control(const char *path, int npdim[3], double rmin[3], double rmax[3], int *chebyshev,
                        char pdetype[32], double vel[3], double *mu, int *nruns, int btypes[6], double bvals[6]);

/* construct_pde / PDEBase (src/pde/pdefactory.hpp:11, pdebase.hpp:14-62); which: 0 test_rhs, 1 solution, 2 rhs */
void *s3dh_pde_create(const char *type, const double vel[3], double mu, const int dirichlet[6], const double bvals[6]);
void  s3dh_pde_destroy(void *pde);
void *s3dh_pde_vector(void *pde, void *mesh, int which);
void *s3dh_pde_vector_sampled(void *pde, void *mesh, int which);   /* forces the host-sampled std::function path */
void *s3dh_pde_lhs(void *pde, void *mesh);

/* SVec / SMat (src/linalg/matvec.hpp:23-89) */
void *s3dh_vec_create(void *mesh);
void *s3dh_vec_clone(void *vec);
void  s3dh_vec_destroy(void *vec);
long long s3dh_vec_len(void *vec);               /* host-layout length of this rank's slab */
int   s3dh_vec_get(void *vec, double *out);      /* device -> host (ghosted layout) */
int   s3dh_vec_set(void *vec, const double *in); /* host -> device */
double s3dh_vec_peek(void *vec, long long idx);  /* vals[idx] through the coherent host mirror */
int   s3dh_vec_poke(void *vec, long long idx, double v);
void *s3dh_vec_device_ptr(void *vec);
void *s3dh_mat_create(void *mesh);
void  s3dh_mat_destroy(void *mat);
long long s3dh_mat_len(void *mat);
int   s3dh_mat_get(void *mat, int stencil, double *out);
int   s3dh_mat_set(void *mat, int stencil, const double *in);
void *s3dh_mat_device_ptr(void *mat);

int   s3dh_apply(void *A, void *x, void *y);
int   s3dh_apply_res(void *A, void *b, void *x, void *y);
int   s3dh_vecset(double a, void *x);
int   s3dh_vecassign(void *x, void *y);
int   s3dh_vecaxpy(double a, void *x, void *y);
int   s3dh_vec_multi_axpy(int n, const double *a, void **xs, void *y);
int   s3dh_norm_L2(void *x, double *out);
int   s3dh_norm_vector_l2(void *x, double *out);
int   s3dh_inner_vector_l2(void *x, void *y, double *out);
int   s3dh_compute_error_L2(void *x, void *y, double *out);

/* preconditioners constructed directly, then updateOperator() */
void *s3dh_prec_create(void *A, const char *pc, int applysweeps, int buildsweeps, int threaded_apply,
                       int threaded_build, int chunk);
void  s3dh_prec_destroy(void *prec);
int   s3dh_prec_apply(void *prec, void *r, void *z);

/* createSolver(A) -> updateOperator() -> apply(b, x)  (runcase(), src/perftesting/scaling_openmp.cpp:84-140) */
void *s3dh_solver_create(void *A);
int   s3dh_solver_update(void *solver);
int   s3dh_solver_apply(void *solver, void *b, void *x, int *converged, int *iters, double *resnorm, double *pctime);
int   s3dh_solver_history(void *solver, double *out, int cap);   /* returns the number of entries */
void  s3dh_solver_destroy(void *solver);

#ifdef __cplusplus
}
#endif
#endif
