// PDE discretisations on Cartesian grids (reference: src/pde/pdebase.hpp:14-62, native half).
#ifndef STRUCT3D_B200_PDEBASE_HPP
#define STRUCT3D_B200_PDEBASE_HPP

#include <array>
#include <functional>
#include <vector>
#include "cartmesh.hpp"
#include "linalg/matvec.hpp"

/// A function of space handed to computeVector.  Any std::function<sreal(const sreal[NDIM])> converts
/// to it (then it is sampled on the host, as the reference does, and uploaded).  The PDE classes'
/// own test/manufactured functions additionally carry their separable form
/// f = sum_t X_t(x) Y_t(y) Z_t(z), which the GPU assembles from three 1-D tables per term.
class GridFunction
{
public:
	typedef std::function<sreal(sreal)> F1;
	struct Term { F1 fx, fy, fz; };

	GridFunction() {}
	template <class Callable, class = decltype(std::declval<Callable>()((const sreal *)nullptr))>
	GridFunction(Callable f) : point(f) {}
	GridFunction(std::function<sreal(const sreal[NDIM])> f, std::vector<Term> separable_terms)
		: point(f), terms(separable_terms) {}

	sreal operator()(const sreal r[NDIM]) const { return point(r); }
	operator std::function<sreal(const sreal[NDIM])>() const { return point; }
	bool separable() const { return !terms.empty(); }
	const std::vector<Term> &separable_terms() const { return terms; }

private:
	std::function<sreal(const sreal[NDIM])> point;
	std::vector<Term> terms;
};

class PDEBase
{
public:
	PDEBase(const std::array<BCType, 6> &bc_types, const std::array<sreal, 6> &bc_vals);
	virtual ~PDEBase() {}

	/// Grid vector of `func` at the real points (plus the PDE's boundary terms); ghosts stay zero.
	SVec computeVector(const CartMesh *const m, const GridFunction &func) const;
	/// The 7-point left-hand-side operator, assembled on the GPU.
	SMat computeLHS(const CartMesh *const m) const;

	/// {solution, right-hand side}
	virtual std::array<GridFunction, 2> manufactured_solution() const = 0;
	virtual GridFunction test_rhs() const = 0;

protected:
	std::array<BCType, 6> bctypes;
	std::array<sreal, 6> bvals;

	/// fills A on the device
	virtual void assemble_lhs(const CartMesh *const m, SMat &A) const = 0;
	/// six additive boundary terms for the boundary-adjacent layers (i-,i+,j-,j+,k-,k+); returns false when there are none
	virtual bool boundary_terms(const CartMesh *const m, sreal face[6]) const = 0;
};

#endif
