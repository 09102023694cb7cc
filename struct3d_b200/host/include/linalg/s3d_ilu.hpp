// Structured ILU(0)-diagonal preconditioner (reference: src/linalg/s3d_ilu.hpp, s3d_ilu.cpp:17-64).
#ifndef STRUCT3D_B200_ILU_HPP
#define STRUCT3D_B200_ILU_HPP

#include "s3d_sgspreconditioners.hpp"

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

#endif
