// Whole solves of small grids in ONE kernel launch.
//
// Below ~40^3 an iteration of a host-driven Krylov loop is a handful of launches of a few microseconds each: 25-35 us per iteration
// even when replayed from a CUDA graph, where the reference needs 17 us on 16 host cores (profiles/r1_solves_vs_reference.md rows
// 1-3).  Here ONE THREAD-BLOCK CLUSTER (8 blocks of 1024 threads on 8 SMs) runs the complete loop -- SpMV, reductions, updates,
// convergence test, residual history -- with the hardware cluster barrier between the phases and the partial sums of the dot products
// exchanged through distributed shared memory: no launch, no grid-wide synchronisation, no host in the loop.  (A single block was
// measured first: 24 us per iteration at 24^3, one SM's memory parallelism is not enough.)  The vectors stay in
// global memory (they live in L1/L2 at these sizes); every phase evaluates the same expressions with the same roundings as the
// library's stand-alone kernels (stencil order 0..6, separate multiply and add, true division for Jacobi), so the iterates differ
// from the host-driven loop only through the summation order of the dot products.
//
// Loops (the host-driven versions in host/src/solvers.cpp are the specification; reference: Richardson::apply,
// src/linalg/s3d_solverbase.cpp:45-80; CG has no counterpart in the reference):
//   Richardson:  r = b - A x; test;  repeat { x += M^-1 r;  r = b - A x;  test }
//   CG:          r = b - A x; test;  z = M^-1 r; p = z;  repeat { q = A p; alpha = (r,z)/(p,q); x += alpha p; r -= alpha q; test;
//                                                                  z = M^-1 r; beta = (r,z)/(r,z)_old; p = z + beta p }
//   GCR(north):  the reference's loop, src/linalg/s3d_gcr.cpp:14-86: cycles of `north` directions, each made orthogonal to the earlier
//                ones of its cycle (up to 8 dot products under one cluster barrier), strict inner test, cap checked between cycles
// with M = diag(A) (Jacobi, a true division) or the identity, and the device-side convergence test of s3d_converge_test.
#include <algorithm>
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int SM_NT = 1024;
constexpr int SM_CL = 8;          // blocks of the cluster that runs a solve (the portable maximum): 8192 threads, hardware cluster barrier

struct SmallArgs {
	GridDev g;
	const double *A, *b;
	double *x, *w0, *w1, *w2;
	double rtol;
	int maxiter;
	int *flags;      // [0] stop code, [1] iterations
	double *scal;    // [0] (r,r) at exit, [1] (b,b)
	double *hist;    // relative residual of every counted iteration (may be null)
	int hist_cap;
	double *gpart;   // cooperative-grid team: [4 parities][SM_NV][blocks] partial sums in global memory
	int north;       // GCR: directions per cycle (the restart length), at most SM_NORTH
	double *pq[2 * 32];   // GCR: p_0, q_0, p_1, q_1 ...
};
constexpr int SM_NORTH = 32;
constexpr int SM_NV = 8;          // dot products reduced under one cluster barrier

// Sum over the whole cluster, the same value in every thread: warp shuffle, one partial per block in ITS shared memory, cluster
// barrier, then every block adds the SM_CL partials in rank order through distributed shared memory.  `part` is double-buffered by the
// caller's parity so that one barrier per reduction suffices.
__device__ __forceinline__ double cluster_sum(double v, double *sm, double *part, cg::cluster_group &cl)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	v = warp_sum(v);
	__syncthreads();                   // the previous use of sm is over
	if (lane == 0) sm[warp] = v;
	__syncthreads();
	double t = sm[lane];               // 32 warps: every warp adds the same 32 partials in the same order
	t = warp_sum(t);
	if (threadIdx.x == 0) *part = t;
	cl.sync();
	double s = 0.0;
#pragma unroll
	for (int r = 0; r < SM_CL; r++) s += *cl.map_shared_rank(part, r);
	return s;
}

// NV sums at once under ONE cluster barrier: warp w adds the 32 warp partials of value w.  `part` holds SM_NV doubles.
template <int NV>
__device__ __forceinline__ void cluster_sum_n(double (&v)[NV], double *smn, double *part, cg::cluster_group &cl)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int l = 0; l < NV; l++) v[l] = warp_sum(v[l]);
	__syncthreads();
	if (lane == 0) {
#pragma unroll
		for (int l = 0; l < NV; l++) smn[l * 32 + warp] = v[l];
	}
	__syncthreads();
	if (warp < NV) {
		double t = warp_sum(smn[warp * 32 + lane]);
		if (lane == 0) part[warp] = t;
	}
	cl.sync();
#pragma unroll
	for (int l = 0; l < NV; l++) {
		double s = 0.0;
#pragma unroll
		for (int r = 0; r < SM_CL; r++) s += *cl.map_shared_rank(part + l, r);
		v[l] = s;
	}
}

// The same two reductions over a whole cooperative GRID (blocks of SM_NT threads): one partial per block in global memory, grid
// barrier, then every warp adds the partials of all blocks in the same fixed order.  `gp`: this parity's [NV][nblocks] partials.
__device__ __forceinline__ double grid_sum(double v, double *sm, double *gp, cg::grid_group &gr)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	v = warp_sum(v);
	__syncthreads();
	if (lane == 0) sm[warp] = v;
	__syncthreads();
	double t = warp_sum(sm[lane]);
	if (threadIdx.x == 0) __stcg(gp + blockIdx.x, t);
	gr.sync();
	double s = 0.0;
	for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(gp + b);
	return warp_sum(s);
}
template <int NV>
__device__ __forceinline__ void grid_sum_n(double (&v)[NV], double *smn, double *gp, cg::grid_group &gr)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int l = 0; l < NV; l++) v[l] = warp_sum(v[l]);
	__syncthreads();
	if (lane == 0) {
#pragma unroll
		for (int l = 0; l < NV; l++) smn[l * 32 + warp] = v[l];
	}
	__syncthreads();
	if (warp < NV) {
		double t = warp_sum(smn[warp * 32 + lane]);
		if (lane == 0) __stcg(gp + warp * gridDim.x + blockIdx.x, t);
	}
	gr.sync();
#pragma unroll
	for (int l = 0; l < NV; l++) {
		double s = 0.0;
		for (int b = lane; b < (int)gridDim.x; b += 32) s += __ldcg(gp + l * gridDim.x + b);
		v[l] = warp_sum(s);
	}
}

__device__ __forceinline__ double madd2(double a, double x, double y) { return __dadd_rn(y, __dmul_rn(a, x)); }   // y + a x, two roundings

// METHOD 0: Richardson, 1: CG, 2: GCR.  JAC: Jacobi preconditioner (else the identity).  GRID: the team is a cooperative grid (one
// block per SM, grid barrier, partial sums through global memory) instead of one cluster of SM_CL blocks: for grids too large for
// eight SMs but small enough to live in L2, where an iteration is still bound by its launches.
template <int METHOD, bool JAC, bool GRID>
__device__ __forceinline__ void small_solve_body(const SmallArgs &a)
{
	__shared__ double sm[32];
	__shared__ double part[4];          // this block's partial sums, by reduction parity (two reductions may follow one another)
	__shared__ double smn[METHOD == 2 ? SM_NV * 32 : 1];
	__shared__ double partn[METHOD == 2 ? 4 * SM_NV : 1];
	cg::cluster_group cl = cg::this_cluster();
	cg::grid_group gr = cg::this_grid();
	const GridDev &g = a.g;
	const int N = g.nx * g.ny * g.nz;
	const int tid = GRID ? (int)(blockIdx.x * SM_NT + threadIdx.x) : (int)(cl.block_rank() * SM_NT + threadIdx.x);
	const int NT = GRID ? (int)(gridDim.x * SM_NT) : SM_CL * SM_NT;
	int par = 0;
	auto tsync = [&]() { if (GRID) gr.sync(); else cl.sync(); };
	auto csum = [&](double v) {
		par = (par + 1) & 3;
		return GRID ? grid_sum(v, sm, a.gpart + (size_t)par * SM_NV * gridDim.x, gr) : cluster_sum(v, sm, &part[par], cl);
	};
	const long long cs = g.cstride;
	auto vaddr = [&](int idx, long long &v, long long &c) {
		const int i = idx % g.nx, jk = idx / g.nx, j = jk % g.ny, k = jk / g.ny;
		v = g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + i;
		c = (long long)k * g.cplane + (long long)j * g.cpitch + i;
	};
	// y = A x  (or b - A x) at one point: the library's expression, products rounded separately and added in stencil order
	auto spmv_at = [&](const double *x, long long v, long long c) {
		double s = __dmul_rn(a.A[c], x[v - 1]);
		s = __dadd_rn(s, __dmul_rn(a.A[cs + c], x[v - g.vpitch]));
		s = __dadd_rn(s, __dmul_rn(a.A[2 * cs + c], x[v - g.vplane]));
		s = __dadd_rn(s, __dmul_rn(a.A[3 * cs + c], x[v]));
		s = __dadd_rn(s, __dmul_rn(a.A[4 * cs + c], x[v + 1]));
		s = __dadd_rn(s, __dmul_rn(a.A[5 * cs + c], x[v + g.vpitch]));
		s = __dadd_rn(s, __dmul_rn(a.A[6 * cs + c], x[v + g.vplane]));
		return s;
	};
	int iters = 0, stop = 0;
	// the device-side convergence test (converge_test_kernel): initial = do not count
	auto test = [&](double rr, double bb, bool initial) {
		const double rel = sqrt(rr) / sqrt(bb);
		if (!initial) {
			if (tid == 0 && a.hist && iters < a.hist_cap) a.hist[iters] = rel;
			iters += 1;
		}
		if (METHOD == 2) { if (rel < a.rtol) stop = S3D_STOP_CONVERGED; }   // GCR's inner test is strict and knows no iteration cap
		else if (!(rel > a.rtol)) stop = S3D_STOP_CONVERGED;
		else if (iters >= a.maxiter) stop = S3D_STOP_MAXITER;
	};

	double acc = 0.0;
	for (int idx = tid; idx < N; idx += NT) {
		long long v, c; vaddr(idx, v, c);
		acc += a.b[v] * a.b[v];
	}
	const double bb = csum(acc);
	double *r = a.w0;
	if (METHOD == 2) {
		// GCR (host/src/solvers.cpp GCR::apply; reference s3d_gcr.cpp:14-86): cycles of `north` directions, every direction
		// orthogonalised against the earlier ones of its cycle; the iteration cap is looked at between cycles only
		double *z = a.w1;
		double qq_kept[SM_NORTH];
		double rr_g = 1.0;
		auto csum_n = [&](double (&v)[SM_NV]) {
			par = (par + 1) & 3;
			if (GRID) grid_sum_n<SM_NV>(v, smn, a.gpart + (size_t)par * SM_NV * gridDim.x, gr);
			else cluster_sum_n<SM_NV>(v, smn, &partn[par * SM_NV], cl);
		};
		while (!stop && iters < a.maxiter) {
			// res = b - A x;  p_0 = M^-1 res
			double *p0 = a.pq[0], *q0 = a.pq[1];
			for (int idx = tid; idx < N; idx += NT) {
				long long v, c; vaddr(idx, v, c);
				const double rv = __dsub_rn(a.b[v], spmv_at(a.x, v, c));
				r[v] = rv;
				p0[v] = JAC ? __ddiv_rn(rv, a.A[3 * cs + c]) : rv;
			}
			tsync();
			// q_0 = A p_0, (res,q_0), (q_0,q_0)
			double two[SM_NV] = {};
			for (int idx = tid; idx < N; idx += NT) {
				long long v, c; vaddr(idx, v, c);
				const double qv = spmv_at(p0, v, c);
				q0[v] = qv;
				two[0] += r[v] * qv;
				two[1] += qv * qv;
			}
			csum_n(two);
			double rq = two[0], qq = two[1];
			for (int k = 0; k < a.north; k++) {
				double *pk = a.pq[2 * k], *qk = a.pq[2 * k + 1];
				const bool last = (k == a.north - 1);
				double *pn = last ? z : a.pq[2 * k + 2];
				const double alpha = rq / qq;
				// x += alpha p_k;  res -= alpha q_k, (res,res);  z = p_{k+1} = M^-1 res (used only if the loop goes on)
				acc = 0.0;
				for (int idx = tid; idx < N; idx += NT) {
					long long v, c; vaddr(idx, v, c);
					a.x[v] = madd2(alpha, pk[v], a.x[v]);
					const double rn = __dadd_rn(r[v], __dmul_rn(-alpha, qk[v]));
					r[v] = rn;
					acc += rn * rn;
					const double zv = JAC ? __ddiv_rn(rn, a.A[3 * cs + c]) : rn;
					z[v] = zv;
					pn[v] = zv;
				}
				rr_g = csum(acc);
				test(rr_g, bb, false);
				if (stop || last) break;
				qq_kept[k] = qq;
				double *qn = a.pq[2 * k + 3];
				// q_{k+1} = A z and its products with q_0..q_k, SM_NV at a time
				double beta[SM_NORTH];
				for (int base = 0; base <= k; base += SM_NV) {
					const int n = min(SM_NV, k + 1 - base);
					double dots[SM_NV] = {};
					for (int idx = tid; idx < N; idx += NT) {
						long long v, c; vaddr(idx, v, c);
						double qv;
						if (base == 0) { qv = spmv_at(z, v, c); qn[v] = qv; }
						else qv = qn[v];
						for (int l = 0; l < n; l++) dots[l] += a.pq[2 * (base + l) + 1][v] * qv;
					}
					csum_n(dots);
					for (int l = 0; l < n; l++) beta[base + l] = -1.0 * dots[l] / qq_kept[base + l];
				}
				// p_{k+1} += sum beta_i p_i, q_{k+1} += sum beta_i q_i (in order), then (res,q_{k+1}), (q_{k+1},q_{k+1})
				double nxt[SM_NV] = {};
				for (int idx = tid; idx < N; idx += NT) {
					long long v, c; vaddr(idx, v, c);
					double pv = pn[v], qv = qn[v];
					for (int l = 0; l <= k; l++) {
						pv = madd2(beta[l], a.pq[2 * l][v], pv);
						qv = madd2(beta[l], a.pq[2 * l + 1][v], qv);
					}
					pn[v] = pv;
					qn[v] = qv;
					nxt[0] += r[v] * qv;
					nxt[1] += qv * qv;
				}
				csum_n(nxt);
				rq = nxt[0]; qq = nxt[1];
			}
		}
		tsync();
		if (tid == 0) {
			a.flags[0] = stop;
			a.flags[1] = iters;
			a.scal[0] = rr_g;
			a.scal[1] = bb;
		}
		return;
	}
	// r = b - A x, (r,r)
	acc = 0.0;
	for (int idx = tid; idx < N; idx += NT) {
		long long v, c; vaddr(idx, v, c);
		const double rv = __dsub_rn(a.b[v], spmv_at(a.x, v, c));
		r[v] = rv;
		acc += rv * rv;
	}
	double rr = csum(acc);
	test(rr, bb, true);

	if (METHOD == 0) {
		while (!stop) {
			// x += M^-1 r
			for (int idx = tid; idx < N; idx += NT) {
				long long v, c; vaddr(idx, v, c);
				const double dx = JAC ? __ddiv_rn(r[v], a.A[3 * cs + c]) : r[v];
				a.x[v] = madd2(1.0, dx, a.x[v]);
			}
			tsync();
			acc = 0.0;
			for (int idx = tid; idx < N; idx += NT) {
				long long v, c; vaddr(idx, v, c);
				const double rv = __dsub_rn(a.b[v], spmv_at(a.x, v, c));
				r[v] = rv;
				acc += rv * rv;
			}
			rr = csum(acc);
			test(rr, bb, false);
		}
	} else {
		double *p = a.w1, *q = a.w2;
		// z = M^-1 r, p = z, (r,z)
		acc = 0.0;
		for (int idx = tid; idx < N; idx += NT) {
			long long v, c; vaddr(idx, v, c);
			const double z = JAC ? __ddiv_rn(r[v], a.A[3 * cs + c]) : r[v];
			p[v] = z;
			acc += r[v] * z;
		}
		double rz = csum(acc);   // contains the cluster barrier that makes p visible
		while (!stop) {
			// q = A p, (p,q)
			acc = 0.0;
			for (int idx = tid; idx < N; idx += NT) {
				long long v, c; vaddr(idx, v, c);
				const double qv = spmv_at(p, v, c);
				q[v] = qv;
				acc += p[v] * qv;
			}
			const double pq = csum(acc);
			const double alpha = rz / pq;
			// x += alpha p; r -= alpha q; (r,r), (r, M^-1 r)
			double acc2 = 0.0;
			acc = 0.0;
			for (int idx = tid; idx < N; idx += NT) {
				long long v, c; vaddr(idx, v, c);
				a.x[v] = madd2(alpha, p[v], a.x[v]);
				const double rn = madd2(-alpha, q[v], r[v]);
				r[v] = rn;
				acc += rn * rn;
				acc2 += rn * (JAC ? __ddiv_rn(rn, a.A[3 * cs + c]) : rn);
			}
			rr = csum(acc);
			const double rz_new = csum(acc2);
			test(rr, bb, false);
			if (stop) break;
			const double beta = rz_new / rz;
			rz = rz_new;
			// p = M^-1 r + beta p
			for (int idx = tid; idx < N; idx += NT) {
				long long v, c; vaddr(idx, v, c);
				const double z = JAC ? __ddiv_rn(r[v], a.A[3 * cs + c]) : r[v];
				p[v] = madd2(beta, p[v], z);
			}
			tsync();
		}
	}
	tsync();                          // no block leaves while another may still read its partial sums
	if (tid == 0) {
		a.flags[0] = stop;
		a.flags[1] = iters;
		a.scal[0] = rr;
		a.scal[1] = bb;
	}
}

template <int METHOD, bool JAC>
__global__ void __cluster_dims__(SM_CL, 1, 1) __launch_bounds__(SM_NT, 1) small_solve_kernel(const SmallArgs a)
{
	small_solve_body<METHOD, JAC, false>(a);
}

template <int METHOD, bool JAC>
__global__ void __launch_bounds__(SM_NT, 1) grid_solve_kernel(const SmallArgs a)
{
	small_solve_body<METHOD, JAC, true>(a);
}

// which team runs a solve: the cluster up to S3D_SMALL_CLUSTER_POINTS points, a cooperative grid of one block per SM above
constexpr long long S3D_SMALL_CLUSTER_POINTS = 1 << 15;

template <int METHOD, bool JAC>
int small_launch(s3d_ctx *ctx, SmallArgs &a, long long npoints)
{
	if (npoints <= S3D_SMALL_CLUSTER_POINTS || ctx->small_team == 1) {
		small_solve_kernel<METHOD, JAC><<<SM_CL, SM_NT, 0, ctx->stream>>>(a);
		S3D_LAUNCHED(ctx);
		return S3D_OK;
	}
	const int nblocks = (int)std::max<long long>(1, std::min<long long>(ctx->sm_count, (npoints + SM_NT - 1) / SM_NT));
	const int rc = s3d_ensure_partials(ctx, (size_t)4 * SM_NV * nblocks);
	if (rc != S3D_OK) return rc;
	a.gpart = ctx->partials;
	void *params[] = {&a};
	S3D_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)grid_solve_kernel<METHOD, JAC>, dim3(nblocks), dim3(SM_NT), params, 0, ctx->stream));
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}

}  // namespace

extern "C" {

// method: 0 Richardson, 1 CG;  pc: 0 none (identity), 1 Jacobi.  d_w0..d_w2: work vectors (w1, w2 only for CG).  Results: d_flags[0] =
// stop code (S3D_STOP_*), d_flags[1] = iterations, d_scal[0] = (r,r) of the last residual, d_scal[1] = (b,b), d_hist[it] = relative
// residual of iteration it (cap entries).  Single rank; the grid must have at most S3D_SMALL_MAX_POINTS points.
int s3d_small_solve(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *d_b, double *d_x, double *d_w0, double *d_w1,
                    double *d_w2, int method, int pc, double rtol, int maxiter, int *d_flags, double *d_scal, double *d_hist, int hist_cap)
{
	S3D_REQUIRE(ctx, g && A && d_b && d_x && d_w0 && d_flags && d_scal, "null argument");
	S3D_REQUIRE(ctx, g->nranks == 1, "the single-launch solver runs on one rank");
	S3D_REQUIRE(ctx, (long long)g->nx * g->ny * g->nz <= S3D_SMALL_MAX_POINTS, "grid too large for the single-launch solver");
	S3D_REQUIRE(ctx, method == 0 || (method == 1 && d_w1 && d_w2), "CG needs three work vectors");
	SmallArgs a;
	a.north = 0;
	a.g = to_dev(g); a.A = A; a.b = d_b; a.x = d_x; a.w0 = d_w0; a.w1 = d_w1; a.w2 = d_w2;
	a.rtol = rtol; a.maxiter = maxiter; a.flags = d_flags; a.scal = d_scal; a.hist = d_hist; a.hist_cap = d_hist ? hist_cap : 0;
	a.gpart = nullptr;
	const long long np = (long long)g->nx * g->ny * g->nz;
	if (method == 0 && pc) return small_launch<0, true>(ctx, a, np);
	if (method == 0) return small_launch<0, false>(ctx, a, np);
	if (pc) return small_launch<1, true>(ctx, a, np);
	return small_launch<1, false>(ctx, a, np);
}

// GCR(north) in one launch: d_res, d_z work vectors, d_pq[2 i] = p_i, d_pq[2 i + 1] = q_i (host array of 2 north device pointers).
int s3d_small_solve_gcr(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *d_b, double *d_x, double *d_res, double *d_z,
                        double *const *d_pq, int north, int pc, double rtol, int maxiter, int *d_flags, double *d_scal, double *d_hist,
                        int hist_cap)
{
	S3D_REQUIRE(ctx, g && A && d_b && d_x && d_res && d_z && d_pq && d_flags && d_scal, "null argument");
	S3D_REQUIRE(ctx, g->nranks == 1, "the single-launch solver runs on one rank");
	S3D_REQUIRE(ctx, (long long)g->nx * g->ny * g->nz <= S3D_SMALL_MAX_POINTS, "grid too large for the single-launch solver");
	S3D_REQUIRE(ctx, north >= 1 && north <= SM_NORTH, "restart length out of range (1..32)");
	SmallArgs a;
	a.g = to_dev(g); a.A = A; a.b = d_b; a.x = d_x; a.w0 = d_res; a.w1 = d_z; a.w2 = nullptr;
	a.rtol = rtol; a.maxiter = maxiter; a.flags = d_flags; a.scal = d_scal; a.hist = d_hist; a.hist_cap = d_hist ? hist_cap : 0;
	a.north = north;
	for (int i = 0; i < 2 * SM_NORTH; i++) a.pq[i] = i < 2 * north ? d_pq[i] : nullptr;
	a.gpart = nullptr;
	const long long np = (long long)g->nx * g->ny * g->nz;
	if (pc) return small_launch<2, true>(ctx, a, np);
	return small_launch<2, false>(ctx, a, np);
}

}  // extern "C"
