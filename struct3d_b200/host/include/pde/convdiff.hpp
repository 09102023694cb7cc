// b.grad u - mu div grad u = f, first-order upwind, constant b (reference: src/pde/convdiff.hpp).
#ifndef STRUCT3D_B200_CONVDIFF_HPP
#define STRUCT3D_B200_CONVDIFF_HPP

#include "pdebase.hpp"

class ConvDiff : public PDEBase
{
public:
	ConvDiff(const std::array<BCType, 6> &bc_types, const std::array<sreal, 6> &bc_vals,
	         const std::array<sreal, NDIM> advel, const sreal diffusion_coeff);
	std::array<GridFunction, 2> manufactured_solution() const;
	GridFunction test_rhs() const;

protected:
	const sreal mu;
	const std::array<sreal, NDIM> b;
	void assemble_lhs(const CartMesh *const m, SMat &A) const;
	bool boundary_terms(const CartMesh *const m, sreal face[6]) const;
};

#endif
