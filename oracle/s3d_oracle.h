/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the Struct3d native structured-grid path.
 *
 * Plain-C restatement of the reference algorithms (each function cites the reference file:line
 * it follows, relative to /root/reference/).  Parity status: PINNED -- checked against the
 * reference's own sources compiled unchanged (oracle/_ref/libs3dref.so, see oracle/Makefile and
 * tests/test_oracle_vs_reference.py) and against fixtures generated from that build
 * (tests/golden/, made by oracle/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (struct3d_b200/) never links, loads or calls it.
 *
 * Layouts are the REFERENCE's host layouts (src/linalg/matvec.hpp:23-89, src/cartmesh.hpp:141-154):
 *   vector  : (np0)(np1)(np2) doubles incl. one ghost layer, index k*np1*np0 + j*np0 + i,
 *             real points 1..np-2 per axis, ghosts never written;
 *   matrix  : 7 arrays, stencil-major {i-1, j-1, k-1, diag, i+1, j+1, k+1}, each (np0-2)(np1-2)(np2-2)
 *             doubles, index k*ny*nx + j*nx + i (no ghosts), stored contiguously A[s*N + idx].
 * npdim[] counts the two boundary points per axis, exactly as in the reference's control files.
 */
#ifndef S3D_ORACLE_H
#define S3D_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_PC_NONE = 0, ORC_PC_JACOBI = 1, ORC_PC_SGS = 2, ORC_PC_STRILU = 3 };
enum { ORC_PDE_POISSON = 0, ORC_PDE_CONVDIFF = 1, ORC_PDE_CONVDIFF_CIRC = 2 };
enum { ORC_F_TEST_RHS = 0, ORC_F_MANUFACTURED_SOLUTION = 1, ORC_F_MANUFACTURED_RHS = 2 };

typedef struct {
	int    pc;            /* ORC_PC_* */
	int    napplysweeps;  /* -s3d_pc_apply_sweeps  (sequential apply: extra sweeps are idempotent) */
	int    nbuildsweeps;  /* -s3d_pc_build_sweeps  (StrILU only) */
} orc_pc_params;

typedef struct {
	int    converged;
	int    iters;
	double resnorm;       /* absolute 2-norm of the final residual */
} orc_solve_info;

/* mesh: src/cartmesh.cpp */
void   orc_coords_uniform(int n, double rmin, double rmax, double *c);
void   orc_coords_chebyshev(int n, double rmin, double rmax, double *c);
int    orc_uniform_selfcheck(int n, const double *c);
double orc_mesh_h(const int npdim[3], const double *cx, const double *cy, const double *cz);

/* assembly: src/pde */
void   orc_lhs_poisson(const int npdim[3], const double *cx, const double *cy, const double *cz, double *A);
void   orc_lhs_convdiff(const int npdim[3], const double *cx, const double *cy, const double *cz,
                        const double b[3], double mu, double *A);
void   orc_lhs_convdiff_circ(const int npdim[3], const double *cx, const double *cy, const double *cz, double mu, double *A);
int    orc_grid_function(const int npdim[3], const double *cx, const double *cy, const double *cz,
                         int pde, int func, const double b[3], double mu, const double bvals[6], double *f);

/* matvec + BLAS-1: src/linalg/matvec.cpp */
void   orc_apply(const int npdim[3], const double *A, const double *x, double *y);
void   orc_apply_res(const int npdim[3], const double *A, const double *b, const double *x, double *y);
void   orc_vecset(const int npdim[3], double a, double *x);
void   orc_vecassign(const int npdim[3], const double *x, double *y);
void   orc_vecaxpy(const int npdim[3], double a, const double *x, double *y);
void   orc_vec_multi_axpy(const int npdim[3], int n, const double *a, const double *const *x, double *y);
double orc_norm_vector_l2(const int npdim[3], const double *x);
double orc_inner_vector_l2(const int npdim[3], const double *x, const double *y);
double orc_norm_L2(const int npdim[3], const double *cx, const double *cy, const double *cz, const double *x);
double orc_compute_error_L2(const int npdim[3], const double *cx, const double *cy, const double *cz,
                            const double *x, const double *y);

/* preconditioners: src/linalg/s3d_jacobi.cpp, s3d_sgspreconditioners.cpp, s3d_ilu.cpp */
void   orc_jacobi_apply(const int npdim[3], const double *A, const double *r, double *z);
void   orc_sgs_setup(const int npdim[3], const double *A, double *diaginv);
void   orc_strilu_setup(const int npdim[3], const double *A, int nbuildsweeps, double *diaginv);
void   orc_sgs_like_apply(const int npdim[3], const double *A, const double *diaginv, int nsweeps,
                          const double *r, double *z);
/* the forward / backward half of orc_sgs_like_apply on its own, honouring the ghost entries of the output (the CPU statement of the
 * sweep THROUGH z-slabs: the ghost plane holds the neighbouring slab's result) */
void   orc_sgs_forward_inplace(const int npdim[3], const double *A, const double *diaginv, const double *r, double *y);
void   orc_sgs_backward_inplace(const int npdim[3], const double *A, const double *diaginv, const double *y, double *z);
/* synchronous Jacobi sweeps on the two triangular systems: NOT in the reference as such; the deterministic counterpart of its
 * asynchronous threaded sweeps, used to pin the GPU's S3D_SGS_JACOBI_SWEEPS ordering. */
void   orc_sgs_jacobi_sweeps_apply(const int npdim[3], const double *A, const double *diaginv, int nsweeps,
                                   const double *r, double *z);
/* red-black (2-colour) variant of the same M = (D+L) D^-1 (D+U) sweeps: NOT in the reference; it is
 * the CPU statement of the multicolour ordering the GPU path offers, used to pin that kernel. */
void   orc_sgs_redblack_apply(const int npdim[3], const double *A, const double *diaginv,
                              const double *r, double *z);

/* solvers: src/linalg/s3d_solverbase.cpp (Richardson), s3d_gcr.cpp (GCR); CG/BiCGSTAB are harness
 * loops with the same conventions (the reference has none).
 * hist (may be NULL) receives the relative residual after each iteration, at most maxiter entries. */
orc_solve_info orc_richardson(const int npdim[3], const double *A, orc_pc_params pc, const double *b,
                              double *x, double rtol, int maxiter, double *hist);
orc_solve_info orc_gcr(const int npdim[3], const double *A, orc_pc_params pc, const double *b, double *x,
                       double rtol, int maxiter, int restart, double *hist);
orc_solve_info orc_cg(const int npdim[3], const double *A, orc_pc_params pc, const double *b, double *x,
                      double rtol, int maxiter, double *hist);
orc_solve_info orc_bicgstab(const int npdim[3], const double *A, orc_pc_params pc, const double *b,
                            double *x, double rtol, int maxiter, double *hist);

#ifdef __cplusplus
}
#endif
#endif
