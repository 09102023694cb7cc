// -div grad u = f with Dirichlet boundaries (reference: src/pde/poisson.hpp).
#ifndef STRUCT3D_B200_POISSON_HPP
#define STRUCT3D_B200_POISSON_HPP

#include "pdebase.hpp"

class Poisson : public PDEBase
{
public:
	Poisson(const std::array<BCType, 6> &bc_types, const std::array<sreal, 6> &bc_vals);
	/// u = sin(2 pi x) sin(2 pi y) sin(2 pi z), f = 12 pi^2 u
	std::array<GridFunction, 2> manufactured_solution() const;
	/// sin(pi x) sin(pi y) sin(pi z)
	GridFunction test_rhs() const;

protected:
	void assemble_lhs(const CartMesh *const m, SMat &A) const;
	bool boundary_terms(const CartMesh *const m, sreal face[6]) const;
};

#endif
