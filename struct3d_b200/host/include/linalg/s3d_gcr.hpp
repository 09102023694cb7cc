// Restarted flexible GCR (reference: src/linalg/s3d_gcr.hpp).
#ifndef STRUCT3D_B200_GCR_HPP
#define STRUCT3D_B200_GCR_HPP

#include "s3d_solverbase.hpp"

class GCR : public SolverBase
{
public:
	GCR(const SMat &lhs, SolverBase *const precond, const SolveParams params);
	void updateOperator();
	SolveInfo apply(const SVec &b, SVec &x) const;

protected:
	SolveParams sparams;
};

#endif
