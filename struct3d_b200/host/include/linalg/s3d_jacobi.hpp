// Jacobi preconditioner z = r / diag(A)  (reference: src/linalg/s3d_jacobi.hpp).
#ifndef STRUCT3D_B200_JACOBI_HPP
#define STRUCT3D_B200_JACOBI_HPP

#include "s3d_solverbase.hpp"

class JacobiPreconditioner : public SolverBase
{
public:
	JacobiPreconditioner(const SMat &lhs);
	void updateOperator();
	SolveInfo apply(const SVec &r, SVec &z) const;
};

#endif
