// Preconditioners of the form (D+L) D^-1 (D+U)  (reference: src/linalg/s3d_sgspreconditioners.hpp).
#ifndef STRUCT3D_B200_SGS_HPP
#define STRUCT3D_B200_SGS_HPP

#include "s3d_solverbase.hpp"

class SGS_like_preconditioner : public SolverBase
{
public:
	SGS_like_preconditioner(const SMat &lhs, const PreconParams parms);
	~SGS_like_preconditioner();
	/// z = (D+U)^-1 D (D+L)^-1 r.  Ordering: lexicographic wavefront (bit-identical to the reference's
	/// sequential sweep) unless params.threadedapply selects the red-black multicolour ordering, or the
	/// option -s3d_sgs_ordering {lexicographic|redblack|sweeps} says otherwise (sweeps: params.napplysweeps synchronous Jacobi
	/// sweeps per triangular factor, the deterministic counterpart of the reference's asynchronous threaded sweeps).
	SolveInfo apply(const SVec &r, SVec &z) const;
	/// S3D_SGS_LEXICOGRAPHIC or S3D_SGS_REDBLACK
	int ordering() const { return order; }

protected:
	double *diaginv;          ///< device array in the coefficient layout (cstride doubles)
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

#endif
