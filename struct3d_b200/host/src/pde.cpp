// PDE assembly front-end: builds the per-axis inputs on the host and lets the GPU assemble
// (reference: src/pde/pdebase.cpp:50-72,123-150, poisson.cpp, convdiff.cpp, pdefactory.cpp).
#include "common_utils.hpp"
#include "pde/convdiff.hpp"
#include "pde/convdiff_circular.hpp"
#include "pde/pdefactory.hpp"
#include "pde/poisson.hpp"

using s3d::Device;

PDEBase::PDEBase(const std::array<BCType, 6> &bc_types, const std::array<sreal, 6> &bc_vals) : bctypes(bc_types), bvals(bc_vals)
{
	if (Device::get().rank() == 0) {
		std::printf("Boundary conditions:\n");
		for (int f = 0; f < 6; f++) std::printf("  %c : %f\n", bctypes[f] == S3D_DIRICHLET ? 'D' : 'O', bvals[f]);
	}
}

SVec PDEBase::computeVector(const CartMesh *const m, const GridFunction &func) const
{
	const Device &d = Device::get();
	SVec f(m);
	sreal face[6];
	const bool has_face = boundary_terms(m, face);
	const s3d_grid &g = m->device_grid();

	if (func.separable()) {
		// three 1-D tables per term, evaluated on the host with the same libm the reference uses, so each
		// table entry is the reference's factor bit for bit; the GPU forms ((X*Y)*Z) and adds terms in order
		const std::vector<GridFunction::Term> &T = func.separable_terms();
		const int nt = (int)T.size();
		const int n[3] = {m->gnpoind(0), m->gnpoind(1), m->gnpoind(2)};
		std::vector<sreal> tab[3];
		for (int a = 0; a < 3; a++) tab[a].resize((size_t)nt * n[a]);
		for (int t = 0; t < nt; t++) {
			for (int i = 0; i < n[0]; i++) tab[0][(size_t)t * n[0] + i] = T[t].fx(m->gcoords(0, i));
			for (int j = 0; j < n[1]; j++) tab[1][(size_t)t * n[1] + j] = T[t].fy(m->gcoords(1, j));
			for (int k = 0; k < n[2]; k++) tab[2][(size_t)t * n[2] + k] = T[t].fz(m->gcoords(2, k));
		}
		d.check(s3d_assemble_rhs_separable(d.ctx(), &g, nt, tab[0].data(), tab[1].data(), tab[2].data(),
		                                   has_face ? face : nullptr, f.dev_rw()), "s3d_assemble_rhs_separable");
		return f;
	}

	// general callable: sample it on the host over this rank's real points, as the reference does, then upload
	const sint np0 = g.nx + 2, np1 = g.ny + 2;
	std::vector<sreal> host((size_t)np0 * np1 * (g.nz + 2), 0.0);
	for (sint k = 1; k <= g.nz; k++)
		for (sint j = 1; j <= g.ny; j++)
			for (sint i = 1; i <= g.nx; i++) {
				const sint K = k + g.k0;
				const sreal r[NDIM] = {m->gcoords(0, i), m->gcoords(1, j), m->gcoords(2, K)};
				sreal v = func(r);
				if (has_face) {
					if (i == 1) v += face[0];
					if (i == g.nx) v += face[1];
					if (j == 1) v += face[2];
					if (j == g.ny) v += face[3];
					if (K == 1) v += face[4];
					if (K == g.nz_global) v += face[5];
				}
				host[(size_t)k * np1 * np0 + (size_t)j * np0 + i] = v;
			}
	f.vals.assign(host.data());
	return f;
}

SMat PDEBase::computeLHS(const CartMesh *const m) const
{
	SMat A(m);
	assemble_lhs(m, A);
	return A;
}

// ------------------------------------------------------------------ Poisson

Poisson::Poisson(const std::array<BCType, 6> &bc_types, const std::array<sreal, 6> &bc_vals) : PDEBase(bc_types, bc_vals) {}

std::array<GridFunction, 2> Poisson::manufactured_solution() const
{
	auto s2 = [](sreal x) { return std::sin(2 * PI * x); };
	GridFunction u([](const sreal r[NDIM]) { return std::sin(2 * PI * r[0]) * std::sin(2 * PI * r[1]) * std::sin(2 * PI * r[2]); },
	               {{s2, s2, s2}});
	GridFunction f([](const sreal r[NDIM]) { return 12 * PI * PI * std::sin(2 * PI * r[0]) * std::sin(2 * PI * r[1]) * std::sin(2 * PI * r[2]); },
	               {{[](sreal x) { return 12 * PI * PI * std::sin(2 * PI * x); }, s2, s2}});
	return {u, f};
}

GridFunction Poisson::test_rhs() const
{
	auto s1 = [](sreal x) { return std::sin(PI * x); };
	return GridFunction([](const sreal r[NDIM]) { return std::sin(PI * r[0]) * std::sin(PI * r[1]) * std::sin(PI * r[2]); }, {{s1, s1, s1}});
}

void Poisson::assemble_lhs(const CartMesh *const m, SMat &A) const
{
	const Device &d = Device::get();
	const sreal *const *c = m->pointer_coords();
	d.check(s3d_assemble_poisson(d.ctx(), &m->device_grid(), c[0], c[1], c[2], A.dev_w()), "s3d_assemble_poisson");
}

// Dirichlet data enter the right-hand side of the boundary-adjacent layers as bval/(dr * 0.5 drs)
// (poisson.cpp:71-119).  The divisor depends only on the face, so each face contributes one scalar.
bool Poisson::boundary_terms(const CartMesh *const m, sreal face[6]) const
{
	static const char *names[6] = {"i-", "i+", "j-", "j+", "k-", "k+"};
	for (int f = 0; f < 6; f++)
		if (bctypes[f] != S3D_DIRICHLET) throw std::runtime_error(std::string("Invalid BC type at ") + names[f] + "!");
	for (int a = 0; a < NDIM; a++) {
		const sint n = m->gnpoind(a);
		{   // lower face: node 1
			const sreal drm = m->gcoords(a, 1) - m->gcoords(a, 0), drs = m->gcoords(a, 2) - m->gcoords(a, 0);
			face[2 * a] = bvals[2 * a] / (drm * 0.5 * drs);
		}
		{   // upper face: node n-2
			const sreal drp = m->gcoords(a, n - 1) - m->gcoords(a, n - 2), drs = m->gcoords(a, n - 1) - m->gcoords(a, n - 3);
			face[2 * a + 1] = bvals[2 * a + 1] / (drp * 0.5 * drs);
		}
	}
	return true;
}

// ------------------------------------------------------------------ ConvDiff

ConvDiff::ConvDiff(const std::array<BCType, 6> &bc_types, const std::array<sreal, 6> &bc_vals,
                   const std::array<sreal, NDIM> advel, const sreal diffc)
	: PDEBase(bc_types, bc_vals), mu{diffc}, b(advel)
{
	if (Device::get().rank() == 0) std::printf("ConvDiff: Using b = (%f,%f,%f), mu = %f.\n", b[0], b[1], b[2], mu);
}

std::array<GridFunction, 2> ConvDiff::manufactured_solution() const
{
	const sreal munum = mu;
	const std::array<sreal, NDIM> advec = b;
	auto s2 = [](sreal x) { return std::sin(2 * PI * x); };
	auto c2 = [](sreal x) { return std::cos(2 * PI * x); };
	GridFunction u([](const sreal r[NDIM]) { return std::sin(2 * PI * r[0]) * std::sin(2 * PI * r[1]) * std::sin(2 * PI * r[2]); },
	               {{s2, s2, s2}});

	auto point = [munum, advec](const sreal r[NDIM]) {
		sreal val = munum * 12 * PI * PI * std::sin(2 * PI * r[0]) * std::sin(2 * PI * r[1]) * std::sin(2 * PI * r[2]);
		for (int i = 0; i < NDIM; i++) {
			sreal term = 2 * PI * advec[i];
			for (int j = 0; j < NDIM; j++) term *= (i == j) ? std::cos(2 * PI * r[j]) : std::sin(2 * PI * r[j]);
			val += term;
		}
		return val;
	};
	// the same sum as four separable terms; constants are folded into the x-factor in evaluation order
	std::vector<GridFunction::Term> terms;
	terms.push_back({[munum](sreal x) { return munum * 12 * PI * PI * std::sin(2 * PI * x); }, s2, s2});
	for (int i = 0; i < NDIM; i++) {
		const sreal bi = advec[i];
		GridFunction::Term t;
		t.fx = (i == 0) ? GridFunction::F1([bi](sreal x) { return 2 * PI * bi * std::cos(2 * PI * x); })
		                : GridFunction::F1([bi](sreal x) { return 2 * PI * bi * std::sin(2 * PI * x); });
		t.fy = (i == 1) ? GridFunction::F1(c2) : GridFunction::F1(s2);
		t.fz = (i == 2) ? GridFunction::F1(c2) : GridFunction::F1(s2);
		terms.push_back(t);
	}
	return {u, GridFunction(point, terms)};
}

GridFunction ConvDiff::test_rhs() const
{
	auto s1 = [](sreal x) { return std::sin(PI * x); };
	return GridFunction([](const sreal r[NDIM]) { return std::sin(PI * r[0]) * std::sin(PI * r[1]) * std::sin(PI * r[2]); }, {{s1, s1, s1}});
}

void ConvDiff::assemble_lhs(const CartMesh *const m, SMat &A) const
{
	const Device &d = Device::get();
	const sreal *const *c = m->pointer_coords();
	const double vel[3] = {b[0], b[1], b[2]};
	d.check(s3d_assemble_convdiff(d.ctx(), &m->device_grid(), c[0], c[1], c[2], vel, mu, A.dev_w()), "s3d_assemble_convdiff");
}

// the reference's boundary terms for this PDE are commented out (convdiff.cpp:96-121): none
bool ConvDiff::boundary_terms(const CartMesh *const, sreal[6]) const { return false; }

// ------------------------------------------------------------------ ConvDiffCirc

ConvDiffCirc::ConvDiffCirc(const std::array<BCType, 6> &bc_types, const std::array<sreal, 6> &bc_vals, const sreal advel, const sreal diffc)
	: PDEBase(bc_types, bc_vals), mu{diffc}, bmag(advel)
{
	if (Device::get().rank() == 0) std::printf("ConvDiff: Using |b| = %f, mu = %f.\n", bmag, mu);
}

std::array<sreal, NDIM> ConvDiffCirc::advectionVel(const sreal r[NDIM])
{
	std::array<sreal, NDIM> v;
	v[0] = 2 * r[1] * (1.0 - r[0] * r[0]);
	v[1] = -2 * r[0] * (1.0 - r[1] * r[1]);
	v[2] = std::sin(PI * r[2]);
	return v;
}

std::array<GridFunction, 2> ConvDiffCirc::manufactured_solution() const
{
	const sreal munum = mu, bmagnum = bmag;
	auto s2 = [](sreal x) { return std::sin(2 * PI * x); };
	GridFunction u([](const sreal r[NDIM]) { return std::sin(2 * PI * r[0]) * std::sin(2 * PI * r[1]) * std::sin(2 * PI * r[2]); },
	               {{s2, s2, s2}});
	GridFunction f([munum, bmagnum](const sreal r[NDIM]) {
		sreal val = munum * 12 * PI * PI * std::sin(2 * PI * r[0]) * std::sin(2 * PI * r[1]) * std::sin(2 * PI * r[2]);
		const std::array<sreal, NDIM> vel = ConvDiffCirc::advectionVel(r);
		for (int i = 0; i < NDIM; i++) {
			sreal term = 2 * PI * bmagnum * vel[i];
			for (int j = 0; j < NDIM; j++) term *= (i == j) ? std::cos(2 * PI * r[j]) : std::sin(2 * PI * r[j]);
			val += term;
		}
		return val;
	});
	return {u, f};
}

GridFunction ConvDiffCirc::test_rhs() const
{
	auto s1 = [](sreal x) { return std::sin(PI * x); };
	return GridFunction([](const sreal r[NDIM]) { return std::sin(PI * r[0]) * std::sin(PI * r[1]) * std::sin(PI * r[2]); }, {{s1, s1, s1}});
}

void ConvDiffCirc::assemble_lhs(const CartMesh *const m, SMat &A) const
{
	const Device &d = Device::get();
	const sreal *const *c = m->pointer_coords();
	d.check(s3d_assemble_convdiff_circ(d.ctx(), &m->device_grid(), c[0], c[1], c[2], mu, A.dev_w()), "s3d_assemble_convdiff_circ");
}

bool ConvDiffCirc::boundary_terms(const CartMesh *const, sreal[6]) const { return false; }   // "TODO: Add BC" in the reference: none

// ------------------------------------------------------------------ factory

PDEBase *construct_pde(const CaseData &cdata)
{
	if (cdata.pdetype == "poisson") return new Poisson(cdata.btypes, cdata.bvals);
	if (cdata.pdetype == "convdiff") return new ConvDiff(cdata.btypes, cdata.bvals, cdata.vel, cdata.diffcoeff);
	if (cdata.pdetype == "convdiff_circular") return new ConvDiffCirc(cdata.btypes, cdata.bvals, cdata.vel[0], cdata.diffcoeff);
	std::printf("PDE type not recognized!\n");
	std::abort();
}
