// testmatvec <control file>
// The reference's MatVec test (test/matvectest.cpp) against the GPU path: on nruns meshes refined by 2,
// (1) apply and apply_res must agree: || A u - b + (b - A u) ||_2 < 1e-14, and
// (2) the defect || b - A u_exact || / sqrt(N) must shrink with slope ~2 (Poisson) or ~1 (convection-diffusion).
#undef NDEBUG
#include <cassert>
#include <cmath>
#include <vector>
#include "case.hpp"
#include "common_utils.hpp"
#include "linalg/matvec.hpp"
#include "pde/pdefactory.hpp"

int main(int argc, char *argv[])
{
	if (argc < 2) { std::printf("Please specify a control file.\n"); return 2; }
	FILE *conf = std::fopen(argv[1], "r");
	if (!conf) { std::printf("cannot open %s\n", argv[1]); return 2; }
	const CaseData cdata = readCtrl(conf);
	std::fclose(conf);
	const int nmesh = cdata.nruns;
	const PDEBase *const pde = construct_pde(cdata);
	std::vector<sreal> h(nmesh), ress(nmesh);
	for (int imesh = 0; imesh < nmesh; imesh++) {
		sint np[NDIM];
		for (int d = 0; d < NDIM; d++) np[d] = (sint)(cdata.npdim[d] * std::pow(2.0, imesh));
		CartMesh m;
		m.createMesh(np);
		if (cdata.gridtype == S3D_CHEBYSHEV) throw std::runtime_error("Chebyshev mesh not supported for grid convergence!");
		m.generateMesh_UniformDistribution(cdata.rmin, cdata.rmax);

		const SVec b = pde->computeVector(&m, pde->manufactured_solution()[1]);
		const SVec uexact = pde->computeVector(&m, pde->manufactured_solution()[0]);
		const SMat A = pde->computeLHS(&m);
		SVec res(&m), tempv(&m);
		A.apply_res(b, uexact, res);
		A.apply(uexact, tempv);
		vecaxpy(-1.0, b, tempv);
		vecaxpy(1.0, res, tempv);
		const sreal diffnorm = norm_vector_l2(tempv);
		std::printf("  Difference between apply_res and apply = %g.\n", diffnorm);
		assert(diffnorm < 1e-14);
		ress[imesh] = norm_vector_l2(res) / std::sqrt((sreal)m.gnpointotal());
		h[imesh] = 1.0 / std::pow(2.0, imesh);
		std::printf("Defect = %f\nMesh size = %f\n--Mesh %d--\n", ress[imesh], h[imesh], imesh + 1);
	}
	delete pde;
	sreal slope = 0;
	for (int i = 1; i < nmesh; i++) {
		slope = (std::log10(ress[i]) - std::log10(ress[i - 1])) / (std::log10(h[i]) - std::log10(h[i - 1]));
		std::printf("Slope %d = %f\n", i, slope);
	}
	if (cdata.pdetype == "poisson") assert(slope >= 1.9 && slope <= 2.1);
	else assert(slope >= 0.9 && slope <= 1.1);
	std::printf("testmatvec: PASSED\n");
	return 0;
}
