// Host-driven solver loops over device-resident data (reference: src/linalg/s3d_solverbase.cpp,
// s3d_jacobi.cpp, s3d_sgspreconditioners.cpp, s3d_gcr.cpp, solverfactory.cpp; CG/BiCGSTAB are new).
//
// How a loop runs on the GPU.  The host only ENQUEUES kernels; every Krylov coefficient is a device
// scalar (s3d_scalar operands: value = imm * *num / *den), every reduction lands in device memory,
// and the convergence test is a one-thread kernel that raises a device flag.  While the loop runs,
// that flag is the context-wide gate: once it is set, every later launch is a no-op, so x is exactly
// the iterate at which the reference's test `resnorm/bnorm > rtol` first failed.  The host polls the
// flag once per batch of iterations, so the only host<->device round trips are those polls.
#include <algorithm>
#include <climits>
#include <cstring>
#include "common_utils.hpp"
#include "linalg/s3d_gcr.hpp"
#include "linalg/s3d_ilu.hpp"
#include "linalg/s3d_jacobi.hpp"
#include "linalg/s3d_krylov.hpp"
#include "linalg/s3d_sgspreconditioners.hpp"
#include "linalg/solverfactory.hpp"

using s3d::Device;

namespace {

s3d_scalar imm(double a) { return s3d_scalar{a, nullptr, nullptr}; }
s3d_scalar ratio(double sign, const double *num, const double *den) { return s3d_scalar{sign, num, den}; }

// CUDA graphs for the iteration batches: -s3d_cuda_graph on|off|auto (default auto = on for grids up to 4 Mi points on one GPU,
// where an iteration is bound by launch latency rather than by HBM; measured in profiles/r1_graphs.md).
bool g_time_phases = true;   // preconditioner stopwatch; off while batches are replayed from a graph (events cannot be timed there)

bool want_graphs(const Device &d, const s3d_grid &g)
{
	if (d.nranks() > 1) return false;                       // the NCCL exchange stays outside captures
	std::string v = "auto";
	if (s3d::options_has("-s3d_cuda_graph")) v = petscoptions_get_string("-s3d_cuda_graph", 10);
	if (v == "on" || v == "1" || v == "true") return true;
	if (v == "off" || v == "0" || v == "false") return false;
	return (long long)g.nx * g.ny * g.nz <= (4ll << 20);
}

// Device-side loop bookkeeping: stop flag, iteration counter, residual history.
struct DeviceLoop {
	const Device &d;
	s3d_ctx *ctx;
	int *flags;        // [0] stop, [1] iterations
	double *hist;
	int hist_cap;
	int printed = 0;
	int every;         // print "Step n: Rel res" every this many steps, like the reference
	std::vector<sreal> &history;
	bool use_graph = false;  // batches after the first are captured once and replayed
	s3d_graph *graph = nullptr;
	int nbatch = 0;

	DeviceLoop(int cap, int print_every, std::vector<sreal> &out)
		: d(Device::get()), ctx(d.ctx()), flags(d.ints(2)), hist(nullptr), hist_cap(cap), every(print_every), history(out)
	{
		// the device writes one history entry per iteration without a bounds check of its own: an absurd iteration cap (1e9 ...) gets
		// no history instead of gigabytes of it
		if (cap < 0 || cap > (1 << 22)) hist_cap = 0;
		hist = hist_cap ? d.scalars(hist_cap) : nullptr;
		history.clear();
		g_time_phases = true;
	}
	~DeviceLoop()
	{
		if (graph) s3d_graph_destroy(ctx, graph);
		g_time_phases = true;
		s3d_ctx_set_gate(ctx, nullptr);
		s3d_free(ctx, flags);
		if (hist) s3d_free(ctx, hist);
	}
	int waits = 0, first_batch_iters = 0;   // iterations the first (eager, timed) batch completed: scales its preconditioner time
	void graphs(const s3d_grid &g) { use_graph = want_graphs(d, g); }
	// With graphs only the first, eagerly enqueued batch can carry the preconditioner stopwatch (events cannot be timed inside a
	// replayed capture): SolveInfo::precapplywtime is then that batch's time scaled by the iteration count -- an estimate, and said so.
	double phase_scale(int iterations) const
	{
		return (use_graph && first_batch_iters > 0 && iterations > first_batch_iters) ? (double)iterations / first_batch_iters : 1.0;
	}
	// One batch of iterations.  `enqueue` must issue the same launches every time it is called from the second batch on (an even
	// number of iterations when the loop ping-pongs its scalars).  The first batch runs eagerly (first-use allocations, the special
	// first iteration), the second is captured and launched, later ones replay the capture.
	template <class F> void batch(F &&enqueue)
	{
		const int index = nbatch++;
		if (!use_graph || index == 0) {
			enqueue();
			if (use_graph) g_time_phases = false;   // from here on the batches are captured / replayed
			return;
		}
		if (!graph) {
			d.check(s3d_graph_begin(ctx), "s3d_graph_begin");
			try {
				enqueue();
			} catch (...) {
				// leave the stream usable: close the capture, drop what was recorded, let the caller see the error
				s3d_graph *dead = nullptr;
				if (s3d_graph_end(ctx, &dead) == S3D_OK) s3d_graph_destroy(ctx, dead);
				throw;
			}
			d.check(s3d_graph_end(ctx, &graph), "s3d_graph_end");
		}
		d.check(s3d_graph_launch(ctx, graph), "s3d_graph_launch");
	}
	int *stop() const { return flags; }
	int *iters() const { return flags + 1; }
	void gate() { d.check(s3d_ctx_set_gate(ctx, flags), "s3d_ctx_set_gate"); }
	void test(const double *rr, const double *bb, double rtol, int maxiter, int how)
	{
		d.check(s3d_converge_test(ctx, rr, bb, rtol, maxiter, how, stop(), iters(), hist), "s3d_converge_test");
	}
	// Pipelined polling.  snapshot(s) enqueues an asynchronous copy of {stop, iterations} into pinned slot s;
	// wait(s) spins until that copy has landed.  A loop snapshots after batch k, enqueues batch k+1 and only then
	// waits for snapshot k: the GPU never idles while the host looks at the flags, and the host never sits in a
	// blocking driver wait (measured: up to 2.8 ms of wake-up latency per poll on the virtualised GPU hosts).
	// Batches enqueued after the stop flag rose are no-ops (the gate), so the state of snapshot k is final.
	void snapshot(int slot) { d.check(s3d_ints_snapshot(ctx, flags, 2, slot), "s3d_ints_snapshot"); }
	int wait(int slot, int &iterations)
	{
		int h[2] = {0, 0};
		d.check(s3d_ints_snapshot_wait(ctx, slot, 2, h), "s3d_ints_snapshot_wait");
		iterations = h[1];
		if (waits++ == 1) first_batch_iters = h[1];   // the second snapshot is the state behind the first batch
		return h[0];
	}
	// after the loop: fetch the residual history recorded on the device and print the reference's progress lines
	void collect(int iterations)
	{
		const int have = std::min(iterations, hist_cap);
		if (have > (int)history.size()) {
			const int old = (int)history.size();
			history.resize(have);
			d.check(s3d_scalars_download(ctx, hist + old, have - old, history.data() + old), "s3d_scalars_download");
		}
		if (d.rank() == 0)
			for (; printed < (int)history.size(); printed++)
				if (printed % every == 0) {
					std::printf("      Step %d: Rel res = %g\n", printed, history[printed]);
					std::fflush(stdout);
				}
	}
};

// how many iterations to enqueue between polls: aim for ~2 ms of GPU work per batch
int batch_size(const s3d_grid &g, double bytes_per_point_per_iter)
{
	const double t = (double)g.nx * g.ny * g.nz * bytes_per_point_per_iter / 6.0e12 + 15e-6;
	const int b = (int)(2.0e-3 / t);
	return std::max(1, std::min(64, b));
}

// Device scalars of a solver loop.  Kernels write their reductions into the LOCAL copy (`loc`), the
// allreduce sums them into the GLOBAL copy (`S`, what every consumer reads).  Out of place, so that once the
// device-side gate has turned the kernels into no-ops a further allreduce just reproduces the same global
// value (an in-place NCCL sum would keep multiplying it by the rank count).  One rank: both are the same memory.
struct Scalars {
	const Device &d;
	double *S;      // global values
	long off;       // loc(i) = S + off + i
	Scalars(const Device &dev, int n) : d(dev), S(dev.scalars(2 * n)), off(dev.nranks() > 1 ? n : 0) {}
	~Scalars() { s3d_free(d.ctx(), S); }
	double *loc(int i) const { return S + off + i; }
	double *glob(int i) const { return S + i; }
	void allreduce(int i, int n) const
	{
		if (off) d.check(s3d_allreduce_sum(d.ctx(), loc(i), glob(i), n), "s3d_allreduce_sum");
	}
	// for slots addressed through a pointer into S (ping-pong pairs)
	double *loc_of(double *g) const { return g + off; }
	void allreduce_at(double *g, int n) const
	{
		if (off) d.check(s3d_allreduce_sum(d.ctx(), g + off, g, n), "s3d_allreduce_sum");
	}
};

SolveInfo finish(const Device &d, const double *rr, const double *bb, double rtol, int iterations, int stopcode, int phase_slot, double phase_scale = 1.0)
{
	double h[2];
	d.check(s3d_scalars_download(d.ctx(), rr, 1, &h[0]), "s3d_scalars_download");
	d.check(s3d_scalars_download(d.ctx(), bb, 1, &h[1]), "s3d_scalars_download");
	const sreal resnorm = std::sqrt(h[0]), bnorm = std::sqrt(h[1]);
	double ms = 0;
	d.check(s3d_phase_total_ms(d.ctx(), phase_slot, &ms), "s3d_phase_total_ms");
	SolveInfo info;
	info.converged = (stopcode != S3D_STOP_BREAKDOWN) && (resnorm / bnorm <= rtol);
	info.iters = iterations;
	info.resnorm = resnorm;
	info.precapplywtime = ms * 1e-3 * phase_scale;
	if (phase_scale != 1.0 && ms > 0 && d.rank() == 0)
		std::printf("      (preconditioner time: %.3g s measured over the first batch, scaled by %.3g to all %d iterations: CUDA-graph replay carries no stopwatch)\n",
		            ms * 1e-3, phase_scale, iterations);
	return info;
}

struct TimedPrec {   // brackets a preconditioner application with the GPU stopwatch (SolveInfo::precapplywtime)
	const Device &d;
	explicit TimedPrec(const Device &dev) : d(dev) { if (g_time_phases) d.check(s3d_phase_begin(d.ctx(), 0), "s3d_phase_begin"); }
	~TimedPrec() { if (g_time_phases) s3d_phase_end(d.ctx(), 0); }
};

// Small grids: the whole solve in ONE kernel launch (s3d_small_solve), for Richardson, CG and GCR with the Jacobi or the identity
// preconditioner on one GPU.  -s3d_small_solve on|off|auto (auto: up to 32 Ki points -- one cluster of 8 blocks -- where a host-driven
// iteration costs 25-35 us of launches against a few microseconds of work, and up to 128 Ki points for Richardson and CG -- a cooperative
// grid of one block per SM; measured in profiles/r2_small_solves.md).
bool want_small_solve(const Device &d, const s3d_grid &g, const SolverBase *prec, int method)
{
	if (d.nranks() > 1) return false;
	const long long n = (long long)g.nx * g.ny * g.nz;
	if (n > S3D_SMALL_MAX_POINTS) return false;
	if (prec && !dynamic_cast<const JacobiPreconditioner *>(prec) && !dynamic_cast<const NoSolver *>(prec)) return false;
	std::string v = "auto";
	if (s3d::options_has("-s3d_small_solve")) v = petscoptions_get_string("-s3d_small_solve", 10);
	if (v == "on" || v == "1" || v == "true") return true;
	if (v == "off" || v == "0" || v == "false") return false;
	// measured (tools/small_solve_probe.py): the cluster team (up to 32 Ki points) wins for every method; the cooperative-grid team that
	// takes over above wins up to ~48^3 for Richardson and CG (15.0 vs 21.9 and 37.7 vs 43.2 us per iteration at 48^3, a draw at 64^3)
	// and loses for GCR, whose Gram-Schmidt step is a grid barrier per group of eight dot products
	return n <= (32ll << 10) || (method != 2 && n <= (128ll << 10));
}

SolveInfo small_solve(int method, const SMat &A, const SolverBase *prec, const SVec &b, SVec &x, SVec &w0, SVec *w1, SVec *w2,
                      const SolveParams &sp, int print_every, std::vector<sreal> &history, double *const *pq = nullptr, int north = 0)
{
	const Device &d = Device::get();
	s3d_ctx *ctx = d.ctx();
	// result buffers live as long as the process: a cudaMalloc / cudaFree pair per solve would cost more than these solves take
	constexpr int HCAP = 1 << 16;
	static int *flags = nullptr;
	static double *scal = nullptr;
	static s3d_ctx *owner = nullptr;          // the buffers belong to a context: a re-initialised device gets new ones
	if (owner != ctx) { flags = d.ints(2); scal = d.scalars(2 + HCAP); owner = ctx; }
	const int pc = dynamic_cast<const JacobiPreconditioner *>(prec) ? 1 : 0;
	if (method == 2)
		d.check(s3d_small_solve_gcr(ctx, &A.grid(), A.dev(), b.dev(), x.dev_rw(), w0.dev_w(), w1->dev_w(), pq, north, pc, sp.rtol, sp.maxiter,
		                            flags, scal, scal + 2, HCAP), "s3d_small_solve_gcr");
	else
		d.check(s3d_small_solve(ctx, &A.grid(), A.dev(), b.dev(), x.dev_rw(), w0.dev_w(), w1 ? w1->dev_w() : nullptr, w2 ? w2->dev_w() : nullptr,
		                        method, pc, sp.rtol, sp.maxiter, flags, scal, scal + 2, HCAP), "s3d_small_solve");
	int hf[2];
	double hs[2];
	d.check(s3d_ints_download(ctx, flags, 2, hf), "s3d_ints_download");
	d.check(s3d_scalars_download(ctx, scal, 2, hs), "s3d_scalars_download");
	const int iterations = hf[1];
	history.assign(std::min(iterations, HCAP), 0.0);
	if (!history.empty()) d.check(s3d_scalars_download(ctx, scal + 2, (int)history.size(), history.data()), "s3d_scalars_download");
	if (d.rank() == 0)
		for (int i = 0; i < (int)history.size(); i += print_every) {
			std::printf("      Step %d: Rel res = %g\n", i, history[i]);
			std::fflush(stdout);
		}
	const sreal resnorm = std::sqrt(hs[0]), bnorm = std::sqrt(hs[1]);
	SolveInfo info;
	info.converged = (hf[0] != S3D_STOP_BREAKDOWN) && (resnorm / bnorm <= sp.rtol);   // the reference's closing test (s3d_gcr.cpp:80)
	info.iters = iterations;
	info.resnorm = resnorm;
	info.precapplywtime = 0.0;   // the preconditioner is fused into the single launch: there is nothing separate to time
	return info;
}

}  // namespace

// ------------------------------------------------------------------ base + identity

SolverBase::SolverBase(const SMat &lhs, SolverBase *const precond) : A(lhs), prec{precond} {}

SolverBase::~SolverBase()
{
	for (SVec *w : workspace) delete w;
	delete prec;
}

SVec &SolverBase::work(size_t i) const
{
	if (workspace.size() <= i) workspace.resize(i + 1, nullptr);
	if (!workspace[i]) workspace[i] = new SVec(A.m);   // zero-filled, ghosts included; every use overwrites the real points
	return *workspace[i];
}

NoSolver::NoSolver(const SMat &lhs, SolverBase *const) : SolverBase(lhs, nullptr) {}

void NoSolver::updateOperator() {}

SolveInfo NoSolver::apply(const SVec &b, SVec &x) const
{
	if (b.m != x.m) throw std::runtime_error("Arguments must be defined over a common mesh!");
	const Device &d = Device::get();
	d.check(s3d_veccopy_all(d.ctx(), &b.grid(), b.dev(), x.dev_w()), "s3d_veccopy_all");
	return {false, 1, 1.0, 0.0};
}

// ------------------------------------------------------------------ preconditioners

JacobiPreconditioner::JacobiPreconditioner(const SMat &lhs) : SolverBase(lhs, nullptr) {}

void JacobiPreconditioner::updateOperator() {}

SolveInfo JacobiPreconditioner::apply(const SVec &r, SVec &z) const
{
	const Device &d = Device::get();
	d.check(s3d_jacobi_apply(d.ctx(), &A.grid(), A.dev(), r.dev(), z.dev_rw()), "s3d_jacobi_apply");
	return {false, 1, 1.0, 0.0};
}

SGS_like_preconditioner::SGS_like_preconditioner(const SMat &lhs, const PreconParams parms)
	: SolverBase(lhs, nullptr), diaginv{nullptr}, params(parms), order{S3D_SGS_LEXICOGRAPHIC}, scratch(lhs.m)
{
	const Device &d = Device::get();
	d.check(s3d_coef_alloc(d.ctx(), &A.grid(), &diaginv), "s3d_coef_alloc");
	// On z-slabs the default is the OVERLAPPING ordering: the reference's lexicographic sweep over every slab extended by one plane of
	// either neighbour, own planes kept (restricted additive Schwarz).  Like the slab-local ordering (`-s3d_sgs_ordering slablocal`: no
	// coupling across the slab faces at all) it has no cross-rank wait, so it scales with the GPU count; unlike it, it stays within a
	// few per cent of the exact sweep's iteration count (profiles/r2_sgs_overlap.md).  `-s3d_sgs_ordering lexicographic` asks for the
	// exact sweep THROUGH the slabs (a pipeline).
	if (d.nranks() > 1) order = S3D_SGS_OVERLAP;
	if (s3d::options_has("-s3d_sgs_ordering")) {
		const std::string o = petscoptions_get_string("-s3d_sgs_ordering", 40);
		if (o == "redblack" || o == "multicolour" || o == "multicolor") order = S3D_SGS_REDBLACK;
		else if (o == "lexicographic") order = S3D_SGS_LEXICOGRAPHIC;
		else if (o == "slablocal") order = S3D_SGS_SLABLOCAL;
		else if (o == "overlap") order = S3D_SGS_OVERLAP;
		else if (o == "blocks") order = S3D_SGS_BLOCKS;
		else if (o == "sweeps" || o == "jacobi_sweeps") order = S3D_SGS_JACOBI_SWEEPS;
		else if (o == "async" || o == "threaded") order = S3D_SGS_ASYNC;
		else throw std::runtime_error("-s3d_sgs_ordering must be lexicographic, blocks, overlap, slablocal, redblack, sweeps or async");
	}
	// The reference's threaded apply / build (s3d_sgspreconditioners.cpp:32-53, s3d_ilu.cpp:28-32) runs asynchronous, unordered sweeps
	// over CPU threads.  Its GPU counterpart exists (-s3d_sgs_ordering async: S3D_SGS_ASYNC, chunks of -s3d_thread_chunk_size rows dealt
	// to warps that never wait for one another), but is opt-in: the default stays the deterministic exact sweep, which needs the fewest
	// iterations and gives the same result on every run.  Asked for through the reference's own switches, say so once.
	if (order == S3D_SGS_ASYNC) d.check(s3d_tune_set(d.ctx(), "sgs_async_chunk", std::max(1, parms.thread_chunk_size)), "s3d_tune_set");
	// z-blocks on one device (-s3d_sgs_ordering blocks): concurrent sweeps over -s3d_sgs_blocks blocks of planes x -s3d_sgs_blocks_y blocks of rows (default 0 = by size: about 32 planes / rows each) with one plane / row of
	// overlap; opt-in -- another operator than the reference's sweep, within a few per cent of its iteration counts.  On z-slabs the ranks ARE the blocks.
	if (order == S3D_SGS_BLOCKS && d.nranks() > 1) order = S3D_SGS_OVERLAP;
	if (order == S3D_SGS_BLOCKS || order == S3D_SGS_OVERLAP) {   // (overlapping slabs are cut further into blocks the same way; 1 and 1: one sweep per slab)
		d.check(s3d_tune_set(d.ctx(), "sgs_blocks", s3d::options_has("-s3d_sgs_blocks") ? std::max(0, petscoptions_get_int("-s3d_sgs_blocks")) : 0), "s3d_tune_set");
		d.check(s3d_tune_set(d.ctx(), "sgs_blocks_y", s3d::options_has("-s3d_sgs_blocks_y") ? std::max(0, petscoptions_get_int("-s3d_sgs_blocks_y")) : 0), "s3d_tune_set");
	}
	static bool warned = false;
	const bool asked = order != S3D_SGS_ASYNC && ((s3d::options_has("-s3d_pc_use_threaded_apply") && parms.threadedapply) || (s3d::options_has("-s3d_pc_use_threaded_build") && parms.threadedbuild));
	if (asked && !warned && d.rank() == 0) {
		std::printf(" WARNING: -s3d_pc_use_threaded_apply / -s3d_pc_use_threaded_build ask for the reference's asynchronous CPU-thread sweeps; "
		            "the GPU sweeps are deterministic unless asked otherwise (ordering: %s). -s3d_sgs_ordering async runs -s3d_pc_apply_sweeps asynchronous "
		            "sweeps (the threaded apply's GPU counterpart; with it -s3d_pc_use_threaded_build true builds StrILU by asynchronous sweeps as well), -s3d_sgs_ordering sweeps as many synchronous ones.\n",
		            order == S3D_SGS_LEXICOGRAPHIC ? "lexicographic" : order == S3D_SGS_SLABLOCAL ? "slablocal" : order == S3D_SGS_OVERLAP ? "overlap" : order == S3D_SGS_BLOCKS ? "blocks" : order == S3D_SGS_REDBLACK ? "redblack"
		            : order == S3D_SGS_ASYNC ? "async" : "sweeps");
		warned = true;
	}
	if (d.nranks() > 1 && order == S3D_SGS_LEXICOGRAPHIC) {
		// the exact sweep can run through the slabs as a pipeline over peer memory (same iterations as one GPU); without the
		// peer-memory windows the red-black ordering takes over
		int yes = 0;
		d.check(s3d_sgs_slab_sweeps_available(d.ctx(), &A.grid(), &yes), "s3d_sgs_slab_sweeps_available");
		if (!yes) {
			if (d.rank() == 0) std::printf(" WARNING: SGS: no peer-memory windows for lexicographic sweeps across z-slabs; using the red-black ordering.\n");
			order = S3D_SGS_REDBLACK;
		}
	}
}

SGS_like_preconditioner::~SGS_like_preconditioner()
{
	if (diaginv) s3d_free(Device::get().ctx(), diaginv);
	if (rbpack) s3d_free(Device::get().ctx(), rbpack);
}

// Red-black ordering: the plain kernels use half of every coefficient sector they touch (a colour is every other entry of a row); a
// colour-packed copy, made once per operator, halves that traffic (-s3d_sgs_rb_pack 0 keeps the plain kernels; costs 7 more arrays).
void SGS_like_preconditioner::pack_if_redblack()
{
	if (order != S3D_SGS_REDBLACK) return;
	if (s3d::options_has("-s3d_sgs_rb_pack") && !petscoptions_get_bool("-s3d_sgs_rb_pack")) return;
	const Device &d = Device::get();
	if (!rbpack) d.check(s3d_sgs_rb_pack_alloc(d.ctx(), &A.grid(), &rbpack), "s3d_sgs_rb_pack_alloc");
	d.check(s3d_sgs_rb_pack(d.ctx(), &A.grid(), A.dev(), diaginv, rbpack), "s3d_sgs_rb_pack");
}

// Overlapping ordering: the neighbouring slabs' boundary planes of A and of dinv into the slack planes (collective, once per operator)
void SGS_like_preconditioner::overlap_setup()
{
	if (order != S3D_SGS_OVERLAP) return;
	const Device &d = Device::get();
	d.check(s3d_sgs_overlap_setup(d.ctx(), &A.grid(), const_cast<double *>(A.dev()), diaginv), "s3d_sgs_overlap_setup");
}

SolveInfo SGS_like_preconditioner::apply(const SVec &r, SVec &z) const
{
	if (r.m != z.m || r.m != A.m) throw std::runtime_error("apply: Vectors and matrix must be defined over the same mesh!");
	const Device &d = Device::get();
	if (order == S3D_SGS_REDBLACK && rbpack) {
		d.check(s3d_sgs_apply_rb_packed(d.ctx(), &A.grid(), rbpack, r.dev(), z.dev_rw(), scratch.dev_rw()), "s3d_sgs_apply_rb_packed");
		return {false, 1, -1.0, 0.0};
	}
	// the overlapping ordering fills the ghost planes of r (halo exchange) before it sweeps
	d.check(s3d_sgs_apply(d.ctx(), &A.grid(), A.dev(), diaginv, order == S3D_SGS_OVERLAP ? r.dev_halo() : r.dev(), z.dev_rw(), scratch.dev_rw(), order,
	                      std::max(1, params.napplysweeps)), "s3d_sgs_apply");
	return {false, 1, -1.0, 0.0};
}

SGS_preconditioner::SGS_preconditioner(const SMat &lhs, const PreconParams parms) : SGS_like_preconditioner(lhs, parms) {}

void SGS_preconditioner::updateOperator()
{
	const Device &d = Device::get();
	d.check(s3d_sgs_setup(d.ctx(), &A.grid(), A.dev(), diaginv), "s3d_sgs_setup");
	pack_if_redblack();
	overlap_setup();
}

StrILU_preconditioner::StrILU_preconditioner(const SMat &lhs, const PreconParams parms) : SGS_like_preconditioner(lhs, parms)
{
	// the factorisation is built slab by slab (block ILU(0) by slabs): it goes with the slab-local application unless asked otherwise
	if (order == S3D_SGS_OVERLAP && !s3d::options_has("-s3d_sgs_ordering")) order = S3D_SGS_SLABLOCAL;
}

void StrILU_preconditioner::updateOperator()
{
	const Device &d = Device::get();
	// -s3d_sgs_ordering async asks for the reference's threaded behaviour as a whole: the factorisation is then built by asynchronous sweeps
	// too (s3d_ilu.cpp:28-32) unless -s3d_pc_use_threaded_build false keeps the exact build
	if (order == S3D_SGS_ASYNC && params.threadedbuild && params.nbuildsweeps > 0) {
		d.check(s3d_tune_set(d.ctx(), "sgs_async_chunk", std::max(1, params.thread_chunk_size)), "s3d_tune_set");
		d.check(s3d_strilu_setup_async(d.ctx(), &A.grid(), A.dev(), params.nbuildsweeps, diaginv), "s3d_strilu_setup_async");
		return;
	}
	d.check(s3d_strilu_setup(d.ctx(), &A.grid(), A.dev(), params.nbuildsweeps, diaginv), "s3d_strilu_setup");
	pack_if_redblack();
	overlap_setup();
}

// ------------------------------------------------------------------ Richardson (s3d_solverbase.cpp:45-80)

Richardson::Richardson(const SMat &lhs, SolverBase *const precond, const SolveParams params)
	: SolverBase(lhs, precond), sparams(params) {}

void Richardson::updateOperator() { if (prec) prec->updateOperator(); }

SolveInfo Richardson::apply(const SVec &b, SVec &x) const
{
	if (b.m != x.m || b.m != A.m) throw std::runtime_error("All vectors and matrix should be defined over the same mesh!");
	const Device &d = Device::get();
	s3d_ctx *ctx = d.ctx();
	const s3d_grid &g = A.grid();
	SVec &res = work(0), &dx = work(1);
	if (want_small_solve(d, g, prec, 0)) return small_solve(0, A, prec, b, x, res, nullptr, nullptr, sparams, 10, history);
	Scalars sc(d, 4);   // [0] (b,b)  [1] unused  [2] (r,r)
	double *S = sc.S;
	DeviceLoop loop(sparams.maxiter + 1, 10, history);

	d.check(s3d_nrm2sq(ctx, &g, b.dev(), sc.loc(0)), "s3d_nrm2sq");
	sc.allreduce(0, 1);
	auto residual = [&]() {
		d.check(s3d_spmv_exchange(ctx, &g, A.dev(), x.dev_halo(), res.dev_w(), b.dev(), nullptr, 1, sc.loc(1), nullptr), "s3d_spmv_exchange");
		sc.allreduce(2, 1);
	};
	residual();
	loop.test(&S[2], &S[0], sparams.rtol, sparams.maxiter, S3D_CONV_INITIAL);
	loop.gate();
	int iterations = 0, code = 0, cur = 0;
	loop.snapshot(cur);
	const int batch = batch_size(g, 200);
	loop.graphs(g);
	for (;;) {
		loop.batch([&]() {
			for (int i = 0; i < batch; i++) {
				{ TimedPrec t(d); prec->apply(res, dx); }
				d.check(s3d_vecaxpy(ctx, &g, imm(1.0), dx.dev(), x.dev_rw()), "s3d_vecaxpy");
				residual();
				loop.test(&S[2], &S[0], sparams.rtol, sparams.maxiter, 0);
			}
		});
		loop.snapshot(cur ^ 1);
		code = loop.wait(cur, iterations);   // the state before the batch just enqueued
		if (code) break;
		cur ^= 1;
	}
	d.check(s3d_ctx_set_gate(ctx, nullptr), "s3d_ctx_set_gate");
	loop.collect(iterations);
	const SolveInfo info = finish(d, &S[2], &S[0], sparams.rtol, iterations, code, 0, loop.phase_scale(iterations));
	return info;
}

// ------------------------------------------------------------------ GCR (s3d_gcr.cpp:14-86)

GCR::GCR(const SMat &lhs, SolverBase *const precond, const SolveParams params) : SolverBase(lhs, precond), sparams(params) {}

void GCR::updateOperator() { if (prec) prec->updateOperator(); }

SolveInfo GCR::apply(const SVec &b, SVec &x) const
{
	if (b.m != x.m || b.m != A.m) throw std::runtime_error("All vectors and matrix should be defined over the same mesh!");
	const Device &d = Device::get();
	s3d_ctx *ctx = d.ctx();
	const s3d_grid &g = A.grid();
	const int north = std::max(1, sparams.restart);
	SVec &res = work(0), &z = work(1);
	std::vector<SVec *> pp(north), qq(north);
	for (int i = 0; i < north; i++) { pp[i] = &work(2 + 2 * i); qq[i] = &work(3 + 2 * i); }
	struct Ref { std::vector<SVec *> &v; SVec &operator[](int i) const { return *v[i]; } } p{pp}, q{qq};
	if (north <= 32 && want_small_solve(d, g, prec, 2)) {
		std::vector<double *> pq(2 * north);
		for (int i = 0; i < north; i++) { pq[2 * i] = pp[i]->dev_w(); pq[2 * i + 1] = qq[i]->dev_w(); }
		return small_solve(2, A, prec, b, x, res, &z, nullptr, sparams, 5, history, pq.data(), north);
	}

	// scalars: [0] (b,b)  [1..2] spmv reductions  [3] (res,res)  [4..5] (res,q_k),(q_k,q_k)
	//          QQ[i] = (q_i,q_i) cached (the reference recomputes the identical value)   DOT[i] = (q_new,q_i)   BETA[i]
	Scalars sc(d, 8 + 3 * north);
	double *S = sc.S;
	double *QQ = S + 8, *DOT = QQ + north, *BETA = DOT + north;
	DeviceLoop loop(sparams.maxiter + north + 1, 5, history);

	d.check(s3d_nrm2sq(ctx, &g, b.dev(), sc.loc(0)), "s3d_nrm2sq");
	sc.allreduce(0, 1);
	// before the first update the reference's resnorm is 1.0 and is never tested: give the final-result
	// slot a defined value in case maxiter == 0
	{ const double one = 1.0; d.check(s3d_scalars_upload(ctx, &S[3], 1, &one), "s3d_scalars_upload"); }
	loop.gate();

	int iterations = 0, code = 0, cur = 0;
	long steps_enqueued = 0;   // without convergence every cycle adds `north` steps: the host knows when maxiter is reached
	loop.snapshot(cur);
	for (;;) {
		if (steps_enqueued >= sparams.maxiter) { code = loop.wait(cur, iterations); break; }
		steps_enqueued += north;
		d.check(s3d_spmv_exchange(ctx, &g, A.dev(), x.dev_halo(), res.dev_w(), b.dev(), nullptr, 0, nullptr, nullptr), "s3d_spmv_exchange");
		{ TimedPrec t(d); prec->apply(res, p[0]); }
		d.check(s3d_spmv_exchange(ctx, &g, A.dev(), p[0].dev_halo(), q[0].dev_w(), nullptr, nullptr, 0, nullptr, nullptr), "s3d_spmv_exchange");
		for (int k = 0; k < north; k++) {
			// alpha = (res,q_k)/(q_k,q_k): one pass over q_k
			const double *two[2] = {res.dev(), q[k].dev()};
			d.check(s3d_vec_multi_dot(ctx, &g, 2, two, q[k].dev(), sc.loc(4)), "s3d_vec_multi_dot");
			sc.allreduce(4, 2);
			d.check(s3d_vecaxpy(ctx, &g, ratio(1.0, &S[4], &S[5]), p[k].dev(), x.dev_rw()), "s3d_vecaxpy");
			// res -= alpha q_k with ||res||^2 fused
			d.check(s3d_vecwaxpby(ctx, &g, imm(1.0), res.dev(), ratio(-1.0, &S[4], &S[5]), q[k].dev(), res.dev_rw(), sc.loc(3)), "s3d_vecwaxpby");
			sc.allreduce(3, 1);
			loop.test(&S[3], &S[0], sparams.rtol, INT_MAX, S3D_CONV_STRICT);
			if (k == north - 1) break;
			// keep (q_k,q_k) for the Gram-Schmidt coefficients of later directions
			d.check(s3d_scalar_op(ctx, 3, &QQ[k], &S[5], &S[5], nullptr, nullptr, 1, 1.0, nullptr), "s3d_scalar_op");
			{ TimedPrec t(d); prec->apply(res, z); }
			d.check(s3d_spmv_exchange(ctx, &g, A.dev(), z.dev_halo(), q[k + 1].dev_w(), nullptr, nullptr, 0, nullptr, nullptr), "s3d_spmv_exchange");
			d.check(s3d_vecassign(ctx, &g, z.dev(), p[k + 1].dev_rw()), "s3d_vecassign");
			// beta_i = -(q_new,q_i)/(q_i,q_i), i <= k, then p_new += sum beta_i p_i, q_new += sum beta_i q_i
			for (int base = 0; base <= k; base += 32) {
				const int n = std::min(32, k + 1 - base);
				const double *qs[32];
				for (int l = 0; l < n; l++) qs[l] = q[base + l].dev();
				d.check(s3d_vec_multi_dot(ctx, &g, n, qs, q[k + 1].dev(), sc.loc_of(&DOT[base])), "s3d_vec_multi_dot");
				sc.allreduce_at(&DOT[base], n);
			}
			for (int base = 0; base <= k; base += 32) {
				const int n = std::min(32, k + 1 - base);
				const double *qs[32], *ps[32];
				for (int l = 0; l < n; l++) { qs[l] = q[base + l].dev(); ps[l] = p[base + l].dev(); }
				d.check(s3d_scalar_op(ctx, 0, &BETA[base], &DOT[base], &QQ[base], nullptr, nullptr, n, -1.0, nullptr), "s3d_scalar_op");
				d.check(s3d_vec_multi_axpy(ctx, &g, n, &BETA[base], ps, p[k + 1].dev_rw()), "s3d_vec_multi_axpy");
				d.check(s3d_vec_multi_axpy(ctx, &g, n, &BETA[base], qs, q[k + 1].dev_rw()), "s3d_vec_multi_axpy");
			}
		}
		loop.snapshot(cur ^ 1);
		code = loop.wait(cur, iterations);
		if (code) break;
		cur ^= 1;
	}
	d.check(s3d_ctx_set_gate(ctx, nullptr), "s3d_ctx_set_gate");
	loop.collect(iterations);
	const SolveInfo info = finish(d, &S[3], &S[0], sparams.rtol, iterations, code, 0, loop.phase_scale(iterations));
	return info;
}

// ------------------------------------------------------------------ CG (new)

CG::CG(const SMat &lhs, SolverBase *const precond, const SolveParams params) : SolverBase(lhs, precond), sparams(params) {}

void CG::updateOperator() { if (prec) prec->updateOperator(); }

SolveInfo CG::apply(const SVec &b, SVec &x) const
{
	if (b.m != x.m || b.m != A.m) throw std::runtime_error("All vectors and matrix should be defined over the same mesh!");
	const Device &d = Device::get();
	s3d_ctx *ctx = d.ctx();
	const s3d_grid &g = A.grid();
	const bool jacobi = dynamic_cast<const JacobiPreconditioner *>(prec) != nullptr;
	SVec &r = work(0), &p = work(1), &q = work(2);
	if (want_small_solve(d, g, prec, 1)) return small_solve(1, A, prec, b, x, r, &p, &q, sparams, 10, history);
	SVec &z = jacobi ? work(0) : work(3);   // unused with Jacobi (folded into the kernels)
	// scalars: [0] (b,b)  [1..2] SpMV reductions: (p,Ap), -  [3,4] and [5,6]: ping-pong pairs {(r,r), (r,z)}
	Scalars sc(d, 8);
	double *S = sc.S;
	DeviceLoop loop(sparams.maxiter + 1, 10, history);

	d.check(s3d_nrm2sq(ctx, &g, b.dev(), sc.loc(0)), "s3d_nrm2sq");
	sc.allreduce(0, 1);
	d.check(s3d_spmv_exchange(ctx, &g, A.dev(), x.dev_halo(), r.dev_w(), b.dev(), nullptr, 1, sc.loc(4), nullptr), "s3d_spmv_exchange");   // (r,r) -> S[5]
	sc.allreduce(5, 1);
	loop.test(&S[5], &S[0], sparams.rtol, sparams.maxiter, S3D_CONV_INITIAL);
	loop.gate();
	// z = M^-1 r, p = z, (r,z) -> S[6]  (the "previous" pair of iteration 0 is {5,6})
	if (jacobi) {
		{ TimedPrec t(d); prec->apply(r, p); }
		d.check(s3d_dot(ctx, &g, r.dev(), p.dev(), sc.loc(6)), "s3d_dot");
	} else {
		{ TimedPrec t(d); prec->apply(r, z); }
		d.check(s3d_vecassign(ctx, &g, z.dev(), p.dev_rw()), "s3d_vecassign");
		d.check(s3d_dot(ctx, &g, r.dev(), z.dev(), sc.loc(6)), "s3d_dot");
	}
	sc.allreduce(6, 1);

	int iterations = 0, code = 0, cur = 0;
	loop.snapshot(cur);
	const int batch = 2 * ((batch_size(g, 200) + 1) / 2);   // even: the scalar pairs ping-pong, and a captured batch must end where it began
	long it = 0;
	const double *rr_final = &S[5];
	loop.graphs(g);
	for (;;) {
		loop.batch([&]() {
		for (int i = 0; i < batch; i++, it++) {
			double *cur = (it % 2 == 0) ? &S[3] : &S[5];    // this iteration's {(r,r), (r,z)}
			double *prev = (it % 2 == 0) ? &S[5] : &S[3];   // the previous iteration's pair
			d.check(s3d_spmv_exchange(ctx, &g, A.dev(), p.dev_halo(), q.dev_w(), nullptr, p.dev(), 0, sc.loc(1), nullptr), "s3d_spmv_exchange");
			sc.allreduce(1, 1);
			// x += alpha p, r -= alpha q, alpha = (r,z)/(p,Ap); fused (r,r) [and (r, r/diag)]
			d.check(s3d_cg_update(ctx, &g, ratio(1.0, &prev[1], &S[1]), p.dev(), q.dev(), x.dev_rw(), r.dev_rw(),
			                      jacobi ? A.dev() : nullptr, sc.loc_of(cur), nullptr), "s3d_cg_update");
			sc.allreduce_at(cur, 2);
			loop.test(&cur[0], &S[0], sparams.rtol, sparams.maxiter, 0);
			if (jacobi) {
				d.check(s3d_cg_direction_jacobi(ctx, &g, ratio(1.0, &cur[1], &prev[1]), A.dev(), r.dev(), p.dev_rw(), nullptr), "s3d_cg_direction_jacobi");
			} else {
				{ TimedPrec t(d); prec->apply(r, z); }
				d.check(s3d_dot(ctx, &g, r.dev(), z.dev(), sc.loc_of(&cur[1])), "s3d_dot");
				sc.allreduce_at(&cur[1], 1);
				d.check(s3d_vecaxpby(ctx, &g, imm(1.0), z.dev(), ratio(1.0, &cur[1], &prev[1]), p.dev_rw()), "s3d_vecaxpby");
			}
		}
		});
		loop.snapshot(cur ^ 1);
		code = loop.wait(cur, iterations);
		if (code) break;
		cur ^= 1;
	}
	d.check(s3d_ctx_set_gate(ctx, nullptr), "s3d_ctx_set_gate");
	loop.collect(iterations);
	// the residual norm of the last COUNTED iteration lives in the pair that iteration wrote
	if (iterations > 0) rr_final = ((iterations - 1) % 2 == 0) ? &S[3] : &S[5];
	const SolveInfo info = finish(d, rr_final, &S[0], sparams.rtol, iterations, code, 0, loop.phase_scale(iterations));
	return info;
}

// ------------------------------------------------------------------ BiCGSTAB (new)

BiCGSTAB::BiCGSTAB(const SMat &lhs, SolverBase *const precond, const SolveParams params) : SolverBase(lhs, precond), sparams(params) {}

void BiCGSTAB::updateOperator() { if (prec) prec->updateOperator(); }

SolveInfo BiCGSTAB::apply(const SVec &b, SVec &x) const
{
	if (b.m != x.m || b.m != A.m) throw std::runtime_error("All vectors and matrix should be defined over the same mesh!");
	const Device &d = Device::get();
	s3d_ctx *ctx = d.ctx();
	const s3d_grid &g = A.grid();
	SVec &r = work(0), &rhat = work(1), &p = work(2), &v = work(3), &s = work(4), &t = work(5), &ph = work(6), &sh = work(7);
	// scalars: [0] (b,b)  [1,2] (rhat,v),-  [3,4] (t,s),(t,t)  [5,6] / [7,8] ping-pong {(r,r), (rhat,r)}
	//          [9] alpha  [10] omega  [11] beta
	Scalars sc(d, 12);
	double *S = sc.S;
	DeviceLoop loop(sparams.maxiter + 1, 10, history);

	d.check(s3d_nrm2sq(ctx, &g, b.dev(), sc.loc(0)), "s3d_nrm2sq");
	sc.allreduce(0, 1);
	d.check(s3d_spmv_exchange(ctx, &g, A.dev(), x.dev_halo(), r.dev_w(), b.dev(), nullptr, 1, sc.loc(6), nullptr), "s3d_spmv_exchange");   // (r,r) -> S[7]
	sc.allreduce(7, 1);
	d.check(s3d_vecassign(ctx, &g, r.dev(), rhat.dev_rw()), "s3d_vecassign");
	// rho_0 = (rhat, r) = (r,r): place it where iteration 0 expects "the previous pair's (rhat,r)" = S[8]
	d.check(s3d_dot(ctx, &g, rhat.dev(), r.dev(), sc.loc(8)), "s3d_dot");
	sc.allreduce(8, 1);
	loop.test(&S[7], &S[0], sparams.rtol, sparams.maxiter, S3D_CONV_INITIAL);
	loop.gate();

	int iterations = 0, code = 0, cur = 0;
	loop.snapshot(cur);
	const int batch = 2 * ((batch_size(g, 400) + 1) / 2);   // even, see CG
	long it = 0;
	loop.graphs(g);
	for (;;) {
		loop.batch([&]() {
		for (int i = 0; i < batch; i++, it++) {
			double *cur = (it % 2 == 0) ? &S[5] : &S[7];
			double *prev = (it % 2 == 0) ? &S[7] : &S[5];
			if (it == 0) {
				d.check(s3d_vecassign(ctx, &g, r.dev(), p.dev_rw()), "s3d_vecassign");
			} else {
				// beta = (rho_new/rho_old)(alpha/omega); rho_new = prev[1], rho_old = cur[1] (written two iterations ago)
				d.check(s3d_scalar_op(ctx, 1, &S[11], &prev[1], &cur[1], &S[9], &S[10], 1, 1.0, loop.stop()), "s3d_scalar_op");
				d.check(s3d_bicg_direction(ctx, &g, ratio(1.0, &S[11], nullptr), ratio(1.0, &S[10], nullptr), r.dev(), v.dev(), p.dev_rw(), nullptr), "s3d_bicg_direction");
			}
			{ TimedPrec tp(d); prec->apply(p, ph); }
			d.check(s3d_spmv_exchange(ctx, &g, A.dev(), ph.dev_halo(), v.dev_w(), nullptr, rhat.dev(), 0, sc.loc(1), nullptr), "s3d_spmv_exchange");
			sc.allreduce(1, 1);
			d.check(s3d_scalar_op(ctx, 2, &S[9], &prev[1], &S[1], nullptr, nullptr, 1, 1.0, loop.stop()), "s3d_scalar_op");   // alpha = rho/(rhat,v)
			d.check(s3d_vecwaxpby(ctx, &g, imm(1.0), r.dev(), ratio(-1.0, &S[9], nullptr), v.dev(), s.dev_rw(), nullptr), "s3d_vecwaxpby");
			{ TimedPrec tp(d); prec->apply(s, sh); }
			d.check(s3d_spmv_exchange(ctx, &g, A.dev(), sh.dev_halo(), t.dev_w(), nullptr, s.dev(), 1, sc.loc(3), nullptr), "s3d_spmv_exchange");
			sc.allreduce(3, 2);
			d.check(s3d_scalar_op(ctx, 2, &S[10], &S[3], &S[4], nullptr, nullptr, 1, 1.0, loop.stop()), "s3d_scalar_op");    // omega = (t,s)/(t,t)
			d.check(s3d_bicg_update(ctx, &g, ratio(1.0, &S[9], nullptr), ph.dev(), ratio(1.0, &S[10], nullptr), sh.dev(), s.dev(), t.dev(),
			                        rhat.dev(), x.dev_rw(), r.dev_rw(), sc.loc_of(cur), nullptr), "s3d_bicg_update");
			sc.allreduce_at(cur, 2);
			loop.test(&cur[0], &S[0], sparams.rtol, sparams.maxiter, 0);
		}
		});
		loop.snapshot(cur ^ 1);
		code = loop.wait(cur, iterations);
		if (code) break;
		cur ^= 1;
	}
	d.check(s3d_ctx_set_gate(ctx, nullptr), "s3d_ctx_set_gate");
	loop.collect(iterations);
	const double *rr_final = &S[7];
	if (iterations > 0) rr_final = ((iterations - 1) % 2 == 0) ? &S[5] : &S[7];
	const SolveInfo info = finish(d, rr_final, &S[0], sparams.rtol, iterations, code, 0, loop.phase_scale(iterations));
	return info;
}

// ------------------------------------------------------------------ factory (solverfactory.cpp:14-100)

SolverBase *createSolver(const SMat &lhs)
{
	const bool talk = Device::get().rank() == 0;
	const std::string tlsolver = petscoptions_get_string("-s3d_ksp_type", 20);
	std::string precstr = "none";
	if (s3d::options_has("-s3d_pc_type")) precstr = petscoptions_get_string("-s3d_pc_type", 30);
	else if (talk) std::printf("No preconditioner\n");

	SolveParams params;
	params.rtol = petscoptions_get_real("-ksp_rtol");
	params.maxiter = petscoptions_get_int("-ksp_max_it");
	params.restart = 30;
	if (s3d::options_has("-ksp_restart")) params.restart = petscoptions_get_int("-ksp_restart");
	else if (talk && tlsolver == "gcr") std::printf(" No KSP restart specified, using 30.\n");

	SolverBase *prec = nullptr;
	if (precstr == "jacobi") {
		if (talk) std::printf("  Using Jacobi preconditioner\n");
		prec = new JacobiPreconditioner(lhs);
	} else if (precstr == "sgs") {
		if (talk) std::printf("  Using SGS preconditioner\n");
		PreconParams pp;
		pp.nbuildsweeps = 0;
		pp.napplysweeps = petscoptions_get_int("-s3d_pc_apply_sweeps");
		pp.thread_chunk_size = petscoptions_get_int("-s3d_thread_chunk_size");
		pp.threadedbuild = false;
		pp.threadedapply = s3d::options_has("-s3d_pc_use_threaded_apply") ? petscoptions_get_bool("-s3d_pc_use_threaded_apply") : true;
		prec = new SGS_preconditioner(lhs, pp);
	} else if (precstr == "strilu") {
		if (talk) std::printf("  Using StrILU preconditioner\n");
		PreconParams pp;
		pp.nbuildsweeps = petscoptions_get_int("-s3d_pc_build_sweeps");
		pp.napplysweeps = petscoptions_get_int("-s3d_pc_apply_sweeps");
		pp.thread_chunk_size = petscoptions_get_int("-s3d_thread_chunk_size");
		pp.threadedbuild = s3d::options_has("-s3d_pc_use_threaded_build") ? petscoptions_get_bool("-s3d_pc_use_threaded_build") : true;
		pp.threadedapply = s3d::options_has("-s3d_pc_use_threaded_apply") ? petscoptions_get_bool("-s3d_pc_use_threaded_apply") : true;
		prec = new StrILU_preconditioner(lhs, pp);
	} else {
		prec = new NoSolver(lhs);
		if (talk) std::printf("WARNING: createSolver: Using no preconditioner\n");
	}

	SolverBase *solver = nullptr;
	if (tlsolver == "richardson") {
		if (talk) std::printf("  Using Richardson solver\n");
		solver = new Richardson(lhs, prec, params);
	} else if (tlsolver == "gcr") {
		if (talk) std::printf("  Using GCR solver\n");
		solver = new GCR(lhs, prec, params);
	} else if (tlsolver == "cg") {
		if (talk && s3d::options_has("-s3d_sgs_ordering") && petscoptions_get_string("-s3d_sgs_ordering", 40) == "async")
			std::printf(" WARNING: the asynchronous SGS sweeps change from one application to the next; CG and BiCGSTAB assume a fixed "
			            "preconditioner and may stall (measured: BiCGSTAB, convdiff 256^3, no convergence in 5000 iterations). Use gcr or richardson, as the reference does.\n");
		if (talk) std::printf("  Using CG solver\n");
		solver = new CG(lhs, prec, params);
	} else if (tlsolver == "bcgs" || tlsolver == "bicgstab") {
		if (talk && s3d::options_has("-s3d_sgs_ordering") && petscoptions_get_string("-s3d_sgs_ordering", 40) == "async")
			std::printf(" WARNING: the asynchronous SGS sweeps change from one application to the next; CG and BiCGSTAB assume a fixed "
			            "preconditioner and may stall (measured: BiCGSTAB, convdiff 256^3, no convergence in 5000 iterations). Use gcr or richardson, as the reference does.\n");
		if (talk) std::printf("  Using BiCGSTAB solver\n");
		solver = new BiCGSTAB(lhs, prec, params);
	} else {
		delete prec;
		throw std::runtime_error("Need a valid solver option!");
	}
	return solver;
}
