// Solver / preconditioner base types (reference: src/linalg/s3d_solverbase.hpp:11-94).
#ifndef STRUCT3D_B200_SOLVERBASE_HPP
#define STRUCT3D_B200_SOLVERBASE_HPP

#include <vector>
#include "cartmesh.hpp"
#include "matvec.hpp"

struct SolveParams {
	sreal rtol;
	int maxiter;
	int restart;
};

struct PreconParams {
	int nbuildsweeps;
	int napplysweeps;
	int thread_chunk_size;     ///< CPU-thread chunking of the reference; accepted and ignored on the GPU
	bool threadedbuild;
	bool threadedapply;        ///< reference: chaotic threaded sweeps.  GPU: true selects the multicolour (red-black) ordering
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

#endif
