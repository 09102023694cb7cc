// runcase <control file> [-options_file x.perc] [-key value ...]
// GPU counterpart of the reference's native driver (src/perftesting/runcase.cpp:21-103 with
// runcase() of scaling_openmp.cpp:84-140): read the case, assemble b = test_rhs and A on the GPU,
// then nruns times { createSolver, updateOperator, zero guess, apply }, and report averages.
// One process per GPU: launched alone it uses one GPU; under torchrun (--no-python) every rank
// takes a z-slab (s3d::Device::initialize_from_env).
#include <chrono>
#include <cmath>
#include <vector>
#include "case.hpp"
#include "common_utils.hpp"
#include "linalg/solverfactory.hpp"
#include "pde/pdefactory.hpp"

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <class T> static double mean(const std::vector<T> &v)
{
	double s = 0;
	for (const T &x : v) s += x;
	return s / v.size();
}
template <class T> static double reldev(const std::vector<T> &v, double avg)
{
	double s = 0;
	for (const T &x : v) s += (x - avg) * (x - avg);
	return std::sqrt(s / v.size()) / avg;
}

int main(int argc, char *argv[])
{
	if (argc < 2) {
		std::printf("Please specify a control file (and -options_file <perc file>).\n");
		return 0;
	}
	try {
		s3d::Device::initialize_from_env();
		const s3d::Device &dev = s3d::Device::get();
		const bool talk = dev.rank() == 0;
		s3d::options_from_args(argc, argv);
		if (talk) std::printf("Number of GPUs (ranks) = %d.\n", dev.nranks());

		FILE *conf = std::fopen(argv[1], "r");
		if (!conf) throw std::runtime_error("cannot open control file");
		const CaseData cdata = readCtrl(conf);
		std::fclose(conf);
		if (talk) std::printf("PDE: %s\n", cdata.pdetype.c_str());
		const PDEBase *const pde = construct_pde(cdata);

		CartMesh m;
		m.createMesh(cdata.npdim);
		if (cdata.gridtype == S3D_CHEBYSHEV) m.generateMesh_ChebyshevDistribution(cdata.rmin, cdata.rmax);
		else m.generateMesh_UniformDistribution(cdata.rmin, cdata.rmax);

		double t0 = now();
		const SVec b = pde->computeVector(&m, pde->test_rhs());
		const SMat A = pde->computeLHS(&m);
		dev.sync();
		if (talk) std::printf("Assembly (RHS + LHS on the GPU) took %f s\n", now() - t0);

		const int nruns = cdata.nruns;
		std::vector<double> wtime(nruns), build(nruns), pcapply(nruns), res(nruns);
		std::vector<int> iters(nruns);
		for (int irun = 0; irun < nruns; irun++) {
			if (talk) std::printf("    Test run %d:\n", irun);
			SolverBase *const solver = createSolver(A);
			t0 = now();
			solver->updateOperator();
			dev.sync();
			build[irun] = now() - t0;
			SVec u(&m);
			t0 = now();
			const SolveInfo info = solver->apply(b, u);
			wtime[irun] = build[irun] + (now() - t0);
			pcapply[irun] = info.precapplywtime;
			delete solver;
			if (!info.converged) throw std::runtime_error("Solver did not converge!");
			iters[irun] = info.iters;
			res[irun] = info.resnorm;
		}
		if (talk) {
			std::printf("Time taken by preconditioner build =            %f\n", mean(build));
			std::printf("Time taken by all preconditioner applications = %f\n", mean(pcapply));
			std::printf("Time taken by native linear solver =            %f\n\n", mean(wtime));
			std::printf("Deviation in preconditioner build time =            %f\n", reldev(build, mean(build)));
			std::printf("Deviation in all preconditioner applications time = %f\n\n", reldev(pcapply, mean(pcapply)));
			std::printf("Avg solver iters: %d.\n", (int)mean(iters));
			std::printf("Deviation in solver iters = %f\n", reldev(iters, mean(iters)));
			std::printf("Final residual norm = %.16g\n\n", mean(res));
		}
		delete pde;
		s3d::Device::shutdown();
	} catch (const std::exception &e) {
		std::printf("runcase: %s\n", e.what());
		return 1;
	}
	return 0;
}
