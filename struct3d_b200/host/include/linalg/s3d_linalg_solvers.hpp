// Every solver and preconditioner class of the native path in ONE header.
//
// The reference spreads these declarations over src/linalg/s3d_solverbase.hpp:11-94, s3d_gcr.hpp, s3d_jacobi.hpp,
// s3d_sgspreconditioners.hpp and s3d_ilu.hpp; callers include those names, so headers of the same names exist next to this file
// and simply forward here.  Class names, constructor and method signatures and the SolveParams / PreconParams / SolveInfo fields
// are the reference's (that is the drop-in boundary); everything behind them is device-resident.
#ifndef STRUCT3D_B200_LINALG_SOLVERS_HPP
#define STRUCT3D_B200_LINALG_SOLVERS_HPP

#include <vector>
#include "cartmesh.hpp"
#include "matvec.hpp"

// ---------------------------------------------------------------- parameter / result types, the abstract solver, identity and Richardson
// Solver / preconditioner base types (reference: src/linalg/s3d_solverbase.hpp:11-94).

struct SolveParams {
	sreal rtol;
	int maxiter;
	int restart;
};

struct PreconParams {
	int nbuildsweeps;
	int napplysweeps;
	int thread_chunk_size;     ///< rows per chunk of the asynchronous sweeps (-s3d_sgs_ordering async); the deterministic orderings ignore it
	bool threadedbuild;
	bool threadedapply;        ///< reference: chaotic threaded sweeps.  GPU: the sweep is deterministic whatever this says (a warning is printed when it is asked for)
};

struct SolveInfo {
	bool converged;            ///< relative residual <= rtol
	int iters;
	sreal resnorm;             ///< absolute 2-norm of the final residual
	double precapplywtime;     ///< seconds spent applying the preconditioner (GPU time from CUDA events)
};

/// Abstract solver over a Cartesian mesh.  Holds the matrix by reference (not owned) and OWNS the
/// preconditioner.  Callers must call updateOperator() before apply(), as with the reference.
class SolverBase
{
public:
	SolverBase(const SMat &lhs, SolverBase *const precond);
	virtual ~SolverBase();
	virtual void updateOperator() = 0;
	/// Solvers: solve A x = b from the initial guess in x.  Preconditioners: x = M^-1 b (enqueued, not awaited).
	virtual SolveInfo apply(const SVec &b, SVec &x) const = 0;

	/// relative residual after each iteration of the last apply() (B200 addition; the reference only prints)
	const std::vector<sreal> &residual_history() const { return history; }

protected:
	const SMat &A;
	SolverBase *const prec;
	mutable std::vector<sreal> history;
	/// Work vectors of apply(), allocated on first use and kept for the life of the solver (the reference
	/// allocates them inside every apply(); on the GPU a cudaMalloc/cudaFree pair per vector per solve costs
	/// far more than the solve's first iterations).  Consequence: apply() is not re-entrant per object.
	SVec &work(size_t i) const;
	mutable std::vector<SVec *> workspace;
};

/// Identity: copies b into x, ghosts included.
class NoSolver : public SolverBase
{
public:
	NoSolver(const SMat &lhs, SolverBase *const precond = nullptr);
	void updateOperator();
	SolveInfo apply(const SVec &b, SVec &x) const;
};

/// x += M^-1 (b - A x) until ||r||/||b|| <= rtol or maxiter.
class Richardson : public SolverBase
{
public:
	Richardson(const SMat &lhs, SolverBase *const precond, const SolveParams params);
	void updateOperator();
	SolveInfo apply(const SVec &b, SVec &x) const;

protected:
	SolveParams sparams;
};

// ---------------------------------------------------------------- restarted flexible GCR
// Restarted flexible GCR (reference: src/linalg/s3d_gcr.hpp).

class GCR : public SolverBase
{
public:
	GCR(const SMat &lhs, SolverBase *const precond, const SolveParams params);
	void updateOperator();
	SolveInfo apply(const SVec &b, SVec &x) const;

protected:
	SolveParams sparams;
};

// ---------------------------------------------------------------- Jacobi
// Jacobi preconditioner z = r / diag(A)  (reference: src/linalg/s3d_jacobi.hpp).

class JacobiPreconditioner : public SolverBase
{
public:
	JacobiPreconditioner(const SMat &lhs);
	void updateOperator();
	SolveInfo apply(const SVec &r, SVec &z) const;
};

// ---------------------------------------------------------------- preconditioners of the form (D+L) D^-1 (D+U)
// Preconditioners of the form (D+L) D^-1 (D+U)  (reference: src/linalg/s3d_sgspreconditioners.hpp).

class SGS_like_preconditioner : public SolverBase
{
public:
	SGS_like_preconditioner(const SMat &lhs, const PreconParams parms);
	~SGS_like_preconditioner();
	/// z = (D+U)^-1 D (D+L)^-1 r.  Ordering: the reference's lexicographic sweep executed as a dataflow (bit-identical to its
	/// sequential sweep) on one GPU, the same sweep inside every z-slab on several (slablocal), unless the option
	/// -s3d_sgs_ordering {lexicographic|slablocal|redblack|sweeps} says otherwise (sweeps: params.napplysweeps synchronous Jacobi
	/// sweeps per triangular factor, the deterministic counterpart of the reference's asynchronous threaded sweeps).
	SolveInfo apply(const SVec &r, SVec &z) const;
	/// S3D_SGS_LEXICOGRAPHIC, S3D_SGS_SLABLOCAL, S3D_SGS_REDBLACK or S3D_SGS_JACOBI_SWEEPS
	int ordering() const { return order; }

protected:
	double *diaginv;          ///< device array in the coefficient layout (cstride doubles)
	double *rbpack = nullptr; ///< red-black ordering: colour-packed copy of the off-diagonal coefficients and diaginv (made by updateOperator)
	void pack_if_redblack();  ///< to be called at the end of updateOperator()
	void overlap_setup();     ///< overlapping ordering on z-slabs: neighbours' boundary planes of A and diaginv into the slack planes; end of updateOperator()
	const PreconParams params;
	int order;
	mutable SVec scratch;     ///< the reference allocates and zeroes this per call; here once
};

class SGS_preconditioner : public SGS_like_preconditioner
{
public:
	SGS_preconditioner(const SMat &lhs, const PreconParams parms);
	/// diaginv = 1/diag(A)
	void updateOperator();
};

// ---------------------------------------------------------------- structured ILU(0) diagonal
// Structured ILU(0)-diagonal preconditioner (reference: src/linalg/s3d_ilu.hpp, s3d_ilu.cpp:17-64).

/// (D+L) D^-1 (D+U) with D the ILU(0) diagonal of the 7-point matrix instead of the matrix diagonal.
/// The application is SGS_like_preconditioner::apply; only the build differs.
class StrILU_preconditioner : public SGS_like_preconditioner
{
public:
	StrILU_preconditioner(const SMat &lhs, const PreconParams params);
	/// Sequential (lexicographic) build executed as a tile-dataflow wavefront on the GPU: one sweep gives the exact
	/// recurrence, so params.nbuildsweeps > 1 reproduces it.  The reference's threaded (chaotic) build is replaced by it.
	void updateOperator();
};

// ---------------------------------------------------------------- CG and BiCGSTAB (new: the reference's native path has only Richardson and GCR)
// Conjugate gradients and BiCGSTAB as SolverBase subclasses.  NEW: the reference's native path has
// only Richardson and GCR (solverfactory.cpp:86-97); these follow the same conventions (zero-or-given
// initial guess, loop while ||r||/||b|| > rtol and iters < maxiter, converged means <= rtol).

/// Preconditioned CG.  With a Jacobi preconditioner the iteration is three fused kernels
/// (SpMV + (p,Ap); x/r update + (r,r) + (r,z); p = r/diag + beta p) and z is never stored.
class CG : public SolverBase
{
public:
	CG(const SMat &lhs, SolverBase *const precond, const SolveParams params);
	void updateOperator();
	SolveInfo apply(const SVec &b, SVec &x) const;

protected:
	SolveParams sparams;
};

/// Right-preconditioned BiCGSTAB, one convergence test per iteration on the recurrence residual.
class BiCGSTAB : public SolverBase
{
public:
	BiCGSTAB(const SMat &lhs, SolverBase *const precond, const SolveParams params);
	void updateOperator();
	SolveInfo apply(const SVec &b, SVec &x) const;

protected:
	SolveParams sparams;
};

#endif
