// Conjugate gradients and BiCGSTAB as SolverBase subclasses.  NEW: the reference's native path has
// only Richardson and GCR (solverfactory.cpp:86-97); these follow the same conventions (zero-or-given
// initial guess, loop while ||r||/||b|| > rtol and iters < maxiter, converged means <= rtol).
#ifndef STRUCT3D_B200_KRYLOV_HPP
#define STRUCT3D_B200_KRYLOV_HPP

#include "s3d_solverbase.hpp"

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
