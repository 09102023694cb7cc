// gridconv <control file> -options_file <perc>
// Order-of-accuracy check by grid refinement (the reference's test/discretization-verification/gridconv.cpp):
// solve with createSolver on nruns meshes, measure the L2 error against the manufactured solution and
// the defect, and assert slopes ~2 (Poisson) / ~1 (upwinded convection-diffusion).
#undef NDEBUG
#include <cassert>
#include <cmath>
#include <vector>
#include "case.hpp"
#include "common_utils.hpp"
#include "linalg/solverfactory.hpp"
#include "pde/pdefactory.hpp"

int main(int argc, char *argv[])
{
	if (argc < 3) { std::printf("Please specify a control file and an options file.\n"); return 2; }
	s3d::options_from_args(argc, argv);
	FILE *conf = std::fopen(argv[1], "r");
	if (!conf) { std::printf("cannot open %s\n", argv[1]); return 2; }
	const CaseData cdata = readCtrl(conf);
	char tmp[100];
	int nrefinedirs = 3;
	if (std::fscanf(conf, "%99s", tmp) != 1 || std::fscanf(conf, "%d", &nrefinedirs) != 1) {
		std::printf("Could not read number of refinement directions, using 3.\n");
		nrefinedirs = 3;
	}
	std::fclose(conf);
	const int nmesh = cdata.nruns;
	assert(nmesh >= 2);
	assert(nrefinedirs >= 1 && nrefinedirs <= 3);
	const PDEBase *const pde = construct_pde(cdata);
	std::vector<sreal> h(nmesh), errors(nmesh), ress(nmesh);
	for (int imesh = 0; imesh < nmesh; imesh++) {
		sint np[NDIM];
		for (int d = 0; d < NDIM; d++) np[d] = d < nrefinedirs ? (sint)(cdata.npdim[d] * std::pow(2.0, imesh)) : cdata.npdim[d];
		CartMesh m;
		m.createMesh(np);
		if (cdata.gridtype == S3D_CHEBYSHEV) m.generateMesh_ChebyshevDistribution(cdata.rmin, cdata.rmax);
		else m.generateMesh_UniformDistribution(cdata.rmin, cdata.rmax);
		const SVec b = pde->computeVector(&m, pde->manufactured_solution()[1]);
		const SVec uexact = pde->computeVector(&m, pde->manufactured_solution()[0]);
		const SMat A = pde->computeLHS(&m);
		SVec res(&m);
		A.apply_res(b, uexact, res);
		ress[imesh] = norm_vector_l2(res) / std::sqrt((sreal)m.gnpointotal());
		SolverBase *const solver = createSolver(A);
		solver->updateOperator();
		SVec u(&m);
		const SolveInfo sinfo = solver->apply(b, u);
		errors[imesh] = compute_error_L2(u, uexact);
		h[imesh] = 1.0 / std::pow(2.0, imesh);
		std::printf("Mesh size = %f\n  Converged in %d itrs.\n  Defect = %f\n  Error = %f\n", h[imesh], sinfo.iters, ress[imesh], errors[imesh]);
		delete solver;
	}
	delete pde;
	sreal slope = 0, resslope = 0;
	for (int i = 1; i < nmesh; i++) {
		slope = (std::log10(errors[i]) - std::log10(errors[i - 1])) / (std::log10(h[i]) - std::log10(h[i - 1]));
		resslope = (std::log10(ress[i]) - std::log10(ress[i - 1])) / (std::log10(h[i]) - std::log10(h[i - 1]));
		std::printf("Error slope %d = %f\nDefect slope %d = %f\n", i, slope, i, resslope);
	}
	const bool poisson = cdata.pdetype == "poisson";
	assert(poisson ? (resslope >= 1.9 && resslope <= 2.1) : (resslope >= 0.9 && resslope <= 1.1));
	assert(poisson ? (slope >= 1.9 && slope <= 2.1) : (slope >= 0.9 && slope <= 1.1));
	std::printf("gridconv: PASSED\n");
	return 0;
}
