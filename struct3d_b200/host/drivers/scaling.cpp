// scaling <control file> -options_file x.perc
// GPU-count counterpart of the reference's OpenMP scaling study (src/perftesting/scaling_openmp.cpp:28-82 runtests_openmp,
// :185-211 report format).  The reference varies the number of THREADS inside one process and, for every pair of
// (build sweeps, apply sweeps) of its asynchronous preconditioners, reports iterations and the speed-up of the preconditioner build
// and apply times against a reference run.  Here the unit of parallelism is the GPU and one process drives one GPU, so one
// invocation measures ONE GPU count (its number of ranks; launch it under torchrun --no-python for more than one):
//   - ranks == -openmptest_ref_threads : the reference run; writes the report header (same text as the reference's) and remembers
//     the reference times in <output file>.ref;
//   - ranks listed in -openmptest_threads_list : appends one report line per sweep pair, same columns, "threads" = GPUs.
// tools/scaling_report.sh runs the reference count first and then every listed count.  Same option keys as
// test/input/scaling_openmp.perc; the sweeps act through -s3d_pc_build_sweeps / -s3d_pc_apply_sweeps exactly as in setSweeps().
#include <chrono>
#include <cmath>
#include <vector>
#include "case.hpp"
#include "common_utils.hpp"
#include "linalg/solverfactory.hpp"
#include "pde/pdefactory.hpp"

namespace {

double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct CaseInfo {   // ThreadCaseInfo of the reference (scaling_openmp.hpp)
	int nthreads, nbuildsweeps, napplysweeps, avg_niter;
	double dev_niter, avg_residual, dev_residual, avg_precbuildtime, dev_precbuildtime, avg_precapplytime, dev_precapplytime;
};

template <class T> double mean(const std::vector<T> &v)
{
	double s = 0;
	for (const T &x : v) s += x;
	return s / v.size();
}
template <class T> double reldev(const std::vector<T> &v, double avg)
{
	double s = 0;
	for (const T &x : v) s += (x - avg) * (x - avg);
	return std::sqrt(s / v.size()) / avg;
}

void setSweeps(int nb, int na)
{
	s3d::options_set("-s3d_pc_build_sweeps", std::to_string(nb));
	s3d::options_set("-s3d_pc_apply_sweeps", std::to_string(na));
	if (petscoptions_get_int("-s3d_pc_build_sweeps") != nb || petscoptions_get_int("-s3d_pc_apply_sweeps") != na)
		throw std::runtime_error("sweeps not set properly!");
}

CaseInfo runcase(const CartMesh &m, const SMat &A, const SVec &b, int nruns, int ngpus)
{
	const s3d::Device &dev = s3d::Device::get();
	std::vector<double> build(nruns), apply(nruns), res(nruns);
	std::vector<int> iters(nruns);
	for (int irun = 0; irun < nruns; irun++) {
		SolverBase *const solver = createSolver(A);
		const double t0 = now();
		solver->updateOperator();
		dev.sync();
		build[irun] = now() - t0;
		SVec u(&m);
		const SolveInfo info = solver->apply(b, u);
		apply[irun] = info.precapplywtime;
		iters[irun] = info.iters;
		res[irun] = info.resnorm;
		delete solver;
		if (!info.converged) throw std::runtime_error("Solver did not converge!");
	}
	CaseInfo c;
	c.nthreads = ngpus;
	c.nbuildsweeps = petscoptions_get_int("-s3d_pc_build_sweeps");
	c.napplysweeps = petscoptions_get_int("-s3d_pc_apply_sweeps");
	c.avg_niter = (int)mean(iters); c.dev_niter = reldev(iters, mean(iters));
	c.avg_residual = mean(res); c.dev_residual = reldev(res, mean(res));
	c.avg_precbuildtime = mean(build); c.dev_precbuildtime = reldev(build, mean(build));
	c.avg_precapplytime = mean(apply); c.dev_precapplytime = reldev(apply, mean(apply));
	return c;
}

void writeReportLine(const CaseInfo &run, const CaseInfo &ref, FILE *fp)
{
	const double total = (ref.avg_precbuildtime + ref.avg_precapplytime) / (run.avg_precbuildtime + run.avg_precapplytime);
	std::fprintf(fp, "%8d %8d %8d %6d %10f %8f %8f %8f %12f %12f %12f\n", run.nbuildsweeps, run.napplysweeps, run.nthreads, run.avg_niter, run.dev_niter,
	             ref.avg_precbuildtime / run.avg_precbuildtime, ref.avg_precapplytime / run.avg_precapplytime, total, run.dev_precbuildtime,
	             run.dev_precapplytime, run.dev_residual);
}

void writeReportHeader(const CaseInfo &ref, FILE *fp)
{
	std::fprintf(fp, "Reference run: Threads %d, build sweeps %d, apply sweeps %d\n", ref.nthreads, ref.nbuildsweeps, ref.napplysweeps);
	std::fprintf(fp, "              iterations %d, iterations dev. %f, residual norm %g, res norm dev. %f\n", ref.avg_niter, ref.dev_niter, ref.avg_residual,
	             ref.dev_residual);
	std::fprintf(fp, "              build time %f, build time dev. %f, apply time %f, apply time dev. %f\n", ref.avg_precbuildtime, ref.dev_precbuildtime,
	             ref.avg_precapplytime, ref.dev_precapplytime);
	std::fprintf(fp, "\n");
	std::fprintf(fp, "%8s %8s %8s %6s %10s %8s %8s %8s %12s %12s %12s\n", "b.swps.", "a.swps.", "threads", "iters", "dev.iters", "b.spdp", "a.spdp", "spdp",
	             "dev.b.time", "dev.a.time", "dev.resnorm");
	std::fprintf(fp, "---------------------------------------------------------------------------------");
	std::fprintf(fp, "-----------------------------\n");
	writeReportLine(ref, ref, fp);
}

}  // namespace

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
		const int ngpus = dev.nranks();
		s3d::options_from_args(argc, argv);
		// the preconditioner stopwatch has to run: keep the iteration batches out of CUDA graphs
		if (!s3d::options_has("-s3d_cuda_graph")) s3d::options_set("-s3d_cuda_graph", "off");

		FILE *conf = std::fopen(argv[1], "r");
		if (!conf) throw std::runtime_error("cannot open control file");
		const CaseData cdata = readCtrl(conf);
		std::fclose(conf);
		const PDEBase *const pde = construct_pde(cdata);
		CartMesh m;
		m.createMesh(cdata.npdim);
		if (cdata.gridtype == S3D_CHEBYSHEV) m.generateMesh_ChebyshevDistribution(cdata.rmin, cdata.rmax);
		else m.generateMesh_UniformDistribution(cdata.rmin, cdata.rmax);
		const SVec b = pde->computeVector(&m, pde->test_rhs());
		const SMat A = pde->computeLHS(&m);

		const int refgpus = petscoptions_get_int("-openmptest_ref_threads");
		const int refb = petscoptions_get_int("-openmptest_ref_build_sweeps"), refa = petscoptions_get_int("-openmptest_ref_apply_sweeps");
		const int nrepeats = petscoptions_get_int("-openmptest_num_repeats"), nrefrepeats = petscoptions_get_int("-openmptest_ref_num_repeats");
		const std::string out = petscoptions_get_string("-openmptest_output_file", 200);
		const std::vector<int> gpulist = petscoptions_get_array_int("-openmptest_threads_list", 30);
		const std::vector<int> nbs = petscoptions_get_array_int("-openmptest_build_sweeps", 30);
		const std::vector<int> nas = petscoptions_get_array_int("-openmptest_apply_sweeps", 30);
		if (nbs.size() != nas.size()) throw std::runtime_error("Sizes of build sweeps list and apply sweeps list must be the same!");
		const std::string side = out + ".ref";

		CaseInfo ref{};
		if (ngpus == refgpus) {
			setSweeps(refb, refa);
			ref = runcase(m, A, b, nrefrepeats, ngpus);
			if (talk) {
				FILE *fp = std::fopen(out.c_str(), "w");
				if (!fp) throw std::runtime_error("Output file could not be opened!");
				writeReportHeader(ref, fp);
				std::fclose(fp);
				FILE *fs = std::fopen(side.c_str(), "w");
				if (!fs) throw std::runtime_error("Reference side file could not be opened!");
				std::fprintf(fs, "%d %d %d %d %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", ref.nthreads, ref.nbuildsweeps, ref.napplysweeps, ref.avg_niter,
				             ref.dev_niter, ref.avg_residual, ref.dev_residual, ref.avg_precbuildtime, ref.dev_precbuildtime, ref.avg_precapplytime,
				             ref.dev_precapplytime);
				std::fclose(fs);
			}
		} else {
			FILE *fs = std::fopen(side.c_str(), "r");
			if (!fs) throw std::runtime_error("run the reference GPU count (-openmptest_ref_threads) first: " + side + " is missing");
			if (std::fscanf(fs, "%d %d %d %d %lf %lf %lf %lf %lf %lf %lf", &ref.nthreads, &ref.nbuildsweeps, &ref.napplysweeps, &ref.avg_niter, &ref.dev_niter,
			                &ref.avg_residual, &ref.dev_residual, &ref.avg_precbuildtime, &ref.dev_precbuildtime, &ref.avg_precapplytime,
			                &ref.dev_precapplytime) != 11)
				throw std::runtime_error("cannot parse " + side);
			std::fclose(fs);
		}
		bool listed = false;
		for (int n : gpulist) listed = listed || (n == ngpus);
		if (listed) {
			for (size_t isw = 0; isw < nbs.size(); isw++) {
				setSweeps(nbs[isw], nas[isw]);
				if (talk) std::printf(" Sweeps = %d,%d; num GPUs = %d.\n", nbs[isw], nas[isw], ngpus);
				const CaseInfo c = runcase(m, A, b, nrepeats, ngpus);
				if (talk) {
					FILE *fp = std::fopen(out.c_str(), "a");
					if (!fp) throw std::runtime_error("Output file could not be opened!");
					writeReportLine(c, ref, fp);
					std::fclose(fp);
				}
			}
		} else if (talk && ngpus != refgpus) {
			std::printf("%d GPUs are neither the reference count nor in -openmptest_threads_list: nothing to do.\n", ngpus);
		}
		delete pde;
		s3d::Device::shutdown();
	} catch (const std::exception &e) {
		std::printf("scaling: %s\n", e.what());
		return 1;
	}
	return 0;
}
