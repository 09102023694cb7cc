// b(r).grad u - mu div grad u = f with the rotational velocity b = (2y(1-x^2), -2x(1-y^2), sin(pi z))
// (reference: src/pde/convdiff_circular.hpp).  The matrix is reproduced exactly as the reference assembles it,
// including that `bmag` only enters the manufactured right-hand side.
#ifndef STRUCT3D_B200_CONVDIFF_CIRCULAR_HPP
#define STRUCT3D_B200_CONVDIFF_CIRCULAR_HPP

#include "pdebase.hpp"

class ConvDiffCirc : public PDEBase
{
public:
	ConvDiffCirc(const std::array<BCType, 6> &bc_types, const std::array<sreal, 6> &bc_vals,
	             const sreal advelmag, const sreal diffusion_coeff);
	/// {sin(2 pi x) sin(2 pi y) sin(2 pi z), matching source}.  The source is not separable: it is sampled on the host.
	std::array<GridFunction, 2> manufactured_solution() const;
	GridFunction test_rhs() const;

protected:
	const sreal mu;
	const sreal bmag;
	static std::array<sreal, NDIM> advectionVel(const sreal r[NDIM]);
	void assemble_lhs(const CartMesh *const m, SMat &A) const;
	bool boundary_terms(const CartMesh *const m, sreal face[6]) const;
};

#endif
