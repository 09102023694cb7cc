// BLAS-1, reductions, Jacobi and the fused CG / BiCGSTAB steps for sm_100a
// (reference: src/linalg/matvec.cpp:148-288, s3d_jacobi.cpp:14-29; the fused steps are new).
//
// All of these are pure streaming kernels bound by HBM bandwidth.  One skeleton serves them all:
// a grid-stride loop over "items" (one aligned double2 = two consecutive real i-points of one row),
// skipping ghost columns/rows/padding so that only real points are touched, exactly like the
// reference loops.  Operations are small functors; those with reductions finish through the
// deterministic two-stage grid_reduce of common.cuh.
//
// Vector updates use separately rounded multiply and add (__dmul_rn/__dadd_rn), i.e. the same
// roundings as the reference's `y += a*x` compiled without FMA contraction, so given equal scalars
// the updated vectors are bit-identical to the oracle's.
#include "common.cuh"

namespace {

struct Item {
	long long v;   // vector offset of the first point of the pair
	long long c;   // coefficient offset of the same point
	int i, j, k;   // 0-based real indices of the first point
	bool pair;     // second point (i+1) is also a real point
};

__device__ __forceinline__ Item make_item(const GridDev &g, unsigned it, unsigned nx2)
{
	const unsigned row = it / nx2;
	const unsigned i2 = it - row * nx2;
	const unsigned k = row / (unsigned)g.ny;
	const unsigned j = row - k * (unsigned)g.ny;
	Item t;
	t.i = (int)(2 * i2); t.j = (int)j; t.k = (int)k;
	t.v = g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + t.i;
	t.c = (long long)k * g.cplane + (long long)j * g.cpitch + t.i;
	t.pair = (t.i + 1 < g.nx);
	return t;
}

// Ops come in two shapes.  SPLIT ops expose load(item) -> registers and apply(item, registers, acc) and normally run in
// ew_chunk_kernel below; plain ops are called item by item here.
template <class Op>
__global__ void __launch_bounds__(256) ew_kernel(const GridDev g, Op op, const RedOut ro, const int *stop)
{
	if (s3d_stopped(stop)) return;
	constexpr int NR = Op::NR > 0 ? Op::NR : 1;
	double acc[NR];
#pragma unroll
	for (int r = 0; r < NR; r++) acc[r] = 0.0;
	op.prepare();
	const unsigned nx2 = (unsigned)(g.nx + 1) >> 1;
	const unsigned total = nx2 * (unsigned)g.ny * (unsigned)g.nz;
	const unsigned stride = gridDim.x * blockDim.x;
	const unsigned first = blockIdx.x * blockDim.x + threadIdx.x;
	if constexpr (Op::SPLIT) {
		for (unsigned long long it = first; it < total; it += 2ull * stride) {
			const unsigned long long it1 = it + stride;
			const bool two = it1 < total;
			const Item t0 = make_item(g, (unsigned)it, nx2);
			const Item t1 = make_item(g, (unsigned)(two ? it1 : it), nx2);
			const auto r0 = op.load(t0);
			const auto r1 = op.load(t1);   // always a valid address; unused when the second item does not exist
			op.apply(t0, r0, acc);
			if (two) op.apply(t1, r1, acc);
		}
	} else {
		for (unsigned long long it = first; it < total; it += stride) op(make_item(g, (unsigned)it, nx2), acc);
	}
	if (Op::NR > 0) grid_reduce<NR>(acc, ro);
}

// The order the SPLIT ops run in: a block walks contiguous segments of 256*U items (U x 4 KB per stream) and issues the loads of all
// U items of a segment before the first store.  Against the grid-stride order above (512^3, one B200, profiles/r1_ew_unroll.md):
// copy 5749 -> 6258 GB/s, axpy 5945 -> 6545, waxpby 6119 -> 6536, cg_update 6214 -> 6578 with U = 8; the ops that divide
// (Jacobi) lose at U = 8 (the division is a long subroutine) and take U = 4: 6217 -> 6448.  Each op names its U (Op::UNROLL);
// the tuning key ew_unroll overrides it (2, 4, 8; -1 = the grid-stride kernel).
template <class Op, int U>
__global__ void __launch_bounds__(256) ew_chunk_kernel(const GridDev g, Op op, const RedOut ro, const int *stop)
{
	if (s3d_stopped(stop)) return;
	constexpr int NR = Op::NR > 0 ? Op::NR : 1;
	double acc[NR];
#pragma unroll
	for (int r = 0; r < NR; r++) acc[r] = 0.0;
	op.prepare();
	const unsigned nx2 = (unsigned)(g.nx + 1) >> 1;
	const unsigned long long total = (unsigned long long)nx2 * (unsigned)g.ny * (unsigned)g.nz;
	constexpr unsigned long long SEG = 256ull * U;
	for (unsigned long long seg = blockIdx.x; seg * SEG < total; seg += gridDim.x) {
		const unsigned long long base = seg * SEG + threadIdx.x;
		Item t[U];
		decltype(op.load(t[0])) r[U];
#pragma unroll
		for (int u = 0; u < U; u++) {
			const unsigned long long it = base + 256ull * u;
			t[u] = make_item(g, (unsigned)(it < total ? it : seg * SEG), nx2);
			r[u] = op.load(t[u]);
		}
#pragma unroll
		for (int u = 0; u < U; u++)
			if (base + 256ull * u < total) op.apply(t[u], r[u], acc);
	}
	if (Op::NR > 0) grid_reduce<NR>(acc, ro);
}

int g_ew_unroll = 0;

template <class Op>
int run(s3d_ctx *ctx, const s3d_grid *g, const Op &op, double *d_red, const int *stop)
{
	const unsigned long long total = (unsigned long long)((g->nx + 1) / 2) * g->ny * g->nz;
	if (total >= 0xffffffffull) return s3d_fail(ctx, S3D_ERR_ARG, "grid too large for 32-bit item index");
	long long blocks = (long long)((total + 255) / 256);
	const long long cap = (long long)ctx->sm_count * 16;   // 8 resident blocks/SM, 2 waves, then grid-stride
	if (blocks > cap) blocks = cap;
	if (blocks < 1) blocks = 1;
	if (Op::NR > 0) {
		if (!d_red) return s3d_fail(ctx, S3D_ERR_ARG, "reduction output pointer is null");
		const int rc = s3d_ensure_partials(ctx, (size_t)blocks * Op::NR);
		if (rc != S3D_OK) return rc;
	}
	RedOut ro{ctx->partials, ctx->ticket, d_red};
	if constexpr (Op::SPLIT) {
		const int U = (g_ew_unroll > 0) ? g_ew_unroll : (g_ew_unroll == 0 ? Op::UNROLL : 0);
		if (U == 2 || U == 4 || U == 8) {
			const long long segs = (long long)((total + 256ull * U - 1) / (256ull * U));
			const unsigned nb = (unsigned)std::max<long long>(1, std::min<long long>(segs, blocks));
			if (U == 2) ew_chunk_kernel<Op, 2><<<nb, 256, 0, ctx->stream>>>(to_dev(g), op, ro, stop ? stop : ctx->gate);
			else if (U == 4) ew_chunk_kernel<Op, 4><<<nb, 256, 0, ctx->stream>>>(to_dev(g), op, ro, stop ? stop : ctx->gate);
			else ew_chunk_kernel<Op, 8><<<nb, 256, 0, ctx->stream>>>(to_dev(g), op, ro, stop ? stop : ctx->gate);
			S3D_LAUNCHED(ctx);
			return S3D_OK;
		}
	}
	ew_kernel<Op><<<(unsigned)blocks, 256, 0, ctx->stream>>>(to_dev(g), op, ro, stop ? stop : ctx->gate);
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}

__device__ __forceinline__ double2 ldp(const double *p, const Item &t)
{
	return t.pair ? ld2(p + t.v) : make_double2(p[t.v], 0.0);
}
__device__ __forceinline__ double2 ldc(const double *p, const Item &t)   // coefficient-layout load
{
	return t.pair ? ld2(p + t.c) : make_double2(p[t.c], 1.0);
}
__device__ __forceinline__ void stp(double *p, const Item &t, double2 v)
{
	if (t.pair) st2(p + t.v, v);
	else p[t.v] = v.x;
}
__device__ __forceinline__ double madd(double a, double x, double y) { return __dadd_rn(y, __dmul_rn(a, x)); }  // y + a*x, two roundings

// ------------------------------------------------------------------ ops

struct OpSet {
	static constexpr int NR = 0;
	static constexpr bool SPLIT = true;
	static constexpr int UNROLL = 8;
	double a; double *x;
	__device__ void prepare() {}
	__device__ char load(const Item &) const { return 0; }
	__device__ void apply(const Item &t, const char &, double *) const { stp(x, t, make_double2(a, a)); }
};

struct OpCopy {
	static constexpr int NR = 0;
	static constexpr bool SPLIT = true;
	static constexpr int UNROLL = 8;
	const double *x; double *y;
	__device__ void prepare() {}
	__device__ double2 load(const Item &t) const { return ldp(x, t); }
	__device__ void apply(const Item &t, const double2 &r, double *) const { stp(y, t, r); }
};

struct Regs2 { double2 a, b; };
struct Regs3 { double2 a, b, c; };
struct Regs4 { double2 a, b, c, d; };
struct Regs6 { double2 a, b, c, d, e, f; };

struct OpAxpy {   // y += a x
	static constexpr int NR = 0;
	static constexpr bool SPLIT = true;
	static constexpr int UNROLL = 8;
	s3d_scalar sa; const double *x; double *y; double a;
	__device__ void prepare() { a = s3d_sval(sa); }
	__device__ Regs2 load(const Item &t) const { return Regs2{ldp(x, t), ldp(y, t)}; }
	__device__ void apply(const Item &t, const Regs2 &r, double *) const
	{
		stp(y, t, make_double2(madd(a, r.a.x, r.b.x), madd(a, r.a.y, r.b.y)));
	}
};

template <bool READY, bool NORM>
struct OpWaxpby {   // w = a x + b y  [+ (w,w)]
	static constexpr int NR = NORM ? 1 : 0;
	static constexpr bool SPLIT = true;
	static constexpr int UNROLL = 8;
	s3d_scalar sa, sb; const double *x; const double *y; double *w; double a, b;
	__device__ void prepare() { a = s3d_sval(sa); b = READY ? s3d_sval(sb) : 0.0; }
	__device__ Regs2 load(const Item &t) const { return Regs2{ldp(x, t), READY ? ldp(y, t) : make_double2(0.0, 0.0)}; }
	__device__ void apply(const Item &t, const Regs2 &q, double *acc) const
	{
		double2 r = make_double2(__dmul_rn(a, q.a.x), __dmul_rn(a, q.a.y));
		if (READY) {
			r.x = __dadd_rn(r.x, __dmul_rn(b, q.b.x));
			r.y = __dadd_rn(r.y, __dmul_rn(b, q.b.y));
		}
		stp(w, t, r);
		if (NORM) { acc[0] += r.x * r.x; if (t.pair) acc[0] += r.y * r.y; }
	}
};

struct OpXpby {   // y = x + b y   (CG direction update p = z + beta p: z + (beta*p), two roundings)
	static constexpr int NR = 0;
	static constexpr bool SPLIT = true;
	static constexpr int UNROLL = 8;
	s3d_scalar sb; const double *x; double *y; double b;
	__device__ void prepare() { b = s3d_sval(sb); }
	__device__ Regs2 load(const Item &t) const { return Regs2{ldp(x, t), ldp(y, t)}; }
	__device__ void apply(const Item &t, const Regs2 &r, double *) const
	{
		stp(y, t, make_double2(madd(b, r.b.x, r.a.x), madd(b, r.b.y, r.a.y)));
	}
};

constexpr int MAXV = 32;
struct PtrTable { const double *p[MAXV]; };

struct OpMultiAxpy {   // y += sum_l a[l] x_l, applied in order l = 0..n-1 like the reference's n passes
	static constexpr int NR = 0;
	static constexpr bool SPLIT = false;
	int n; const double *a; PtrTable xs; double *y;
	__device__ void prepare() {}
	__device__ void operator()(const Item &t, double *) const
	{
		double2 yv = ldp(y, t);
		for (int l = 0; l < n; l++) {
			const double al = __ldg(a + l);
			const double2 xv = ldp(xs.p[l], t);
			yv.x = madd(al, xv.x, yv.x);
			yv.y = madd(al, xv.y, yv.y);
		}
		stp(y, t, yv);
	}
};

template <int N>
struct OpMultiDot {   // acc[l] = (x_l, y)
	static constexpr int NR = N;
	static constexpr bool SPLIT = false;
	int n; PtrTable xs; const double *y;
	__device__ void prepare() {}
	__device__ void operator()(const Item &t, double *acc) const
	{
		const double2 yv = ldp(y, t);
#pragma unroll
		for (int l = 0; l < N; l++) {
			if (l < n) {
				const double2 xv = ldp(xs.p[l], t);
				acc[l] += xv.x * yv.x + xv.y * yv.y;
			}
		}
	}
};

struct OpDot {
	static constexpr int NR = 1;
	static constexpr bool SPLIT = true;
	static constexpr int UNROLL = 4;
	const double *x, *y;
	__device__ void prepare() {}
	__device__ Regs2 load(const Item &t) const { return Regs2{ldp(x, t), ldp(y, t)}; }
	__device__ void apply(const Item &, const Regs2 &r, double *acc) const { acc[0] += r.a.x * r.b.x + r.a.y * r.b.y; }
};

struct OpNrm2 {
	static constexpr int NR = 1;
	static constexpr bool SPLIT = true;
	static constexpr int UNROLL = 8;
	const double *x;
	__device__ void prepare() {}
	__device__ double2 load(const Item &t) const { return ldp(x, t); }
	__device__ void apply(const Item &, const double2 &r, double *acc) const { acc[0] += r.x * r.x + r.y * r.y; }
};

struct OpNormL2 {   // sum x^2 * vol, vol = 1/8 * wx[i] * wy[j] * wz[k]   (matvec.cpp:221-238)
	static constexpr int NR = 1;
	static constexpr bool SPLIT = false;
	const double *x; const double *wx, *wy, *wz;   // w*[m] = c[m+2] - c[m] for real index m
	__device__ void prepare() {}
	__device__ void operator()(const Item &t, double *acc) const
	{
		const double2 xv = ldp(x, t);
		const double wyz_y = __ldg(wy + t.j), wyz_z = __ldg(wz + t.k);
		const double v0 = __dmul_rn(__dmul_rn(__dmul_rn(0.125, __ldg(wx + t.i)), wyz_y), wyz_z);
		acc[0] += __dmul_rn(__dmul_rn(xv.x, xv.x), v0);
		if (t.pair) {
			const double v1 = __dmul_rn(__dmul_rn(__dmul_rn(0.125, __ldg(wx + t.i + 1)), wyz_y), wyz_z);
			acc[0] += __dmul_rn(__dmul_rn(xv.y, xv.y), v1);
		}
	}
};

struct OpJacobi {   // z = r / diag(A)   (s3d_jacobi.cpp:17-26: a true division)
	static constexpr int NR = 0;
	static constexpr bool SPLIT = true;
	static constexpr int UNROLL = 4;
	const double *diag; const double *r; double *z;
	__device__ void prepare() {}
	__device__ Regs2 load(const Item &t) const { return Regs2{ldp(r, t), ldc(diag, t)}; }
	__device__ void apply(const Item &t, const Regs2 &q, double *) const
	{
		stp(z, t, make_double2(__ddiv_rn(q.a.x, q.b.x), __ddiv_rn(q.a.y, q.b.y)));
	}
};

template <bool JAC>
struct OpCgUpdate {   // x += alpha p; r -= alpha q; (r,r); [ (r, r/diag) ]
	static constexpr int NR = 2;
	static constexpr bool SPLIT = true;
	static constexpr int UNROLL = 8;
	s3d_scalar salpha; const double *p, *q; double *x, *r; const double *diag; double alpha;
	struct R { double2 p, q, x, r, d; };
	__device__ void prepare() { alpha = s3d_sval(salpha); }
	__device__ R load(const Item &t) const
	{
		return R{ldp(p, t), ldp(q, t), ldp(x, t), ldp(r, t), JAC ? ldc(diag, t) : make_double2(1.0, 1.0)};
	}
	__device__ void apply(const Item &t, const R &v, double *acc) const
	{
		const double2 xn = make_double2(madd(alpha, v.p.x, v.x.x), madd(alpha, v.p.y, v.x.y));
		const double2 rn = make_double2(madd(-alpha, v.q.x, v.r.x), madd(-alpha, v.q.y, v.r.y));
		stp(x, t, xn);
		stp(r, t, rn);
		acc[0] += rn.x * rn.x;
		if (t.pair) acc[0] += rn.y * rn.y;
		if (JAC) {
			acc[1] += rn.x * __ddiv_rn(rn.x, v.d.x);
			if (t.pair) acc[1] += rn.y * __ddiv_rn(rn.y, v.d.y);
		}
	}
};

struct OpCgDirJacobi {   // p = r/diag + beta p
	static constexpr int NR = 0;
	static constexpr bool SPLIT = true;
	static constexpr int UNROLL = 4;
	s3d_scalar sbeta; const double *diag; const double *r; double *p; double beta;
	__device__ void prepare() { beta = s3d_sval(sbeta); }
	__device__ Regs3 load(const Item &t) const { return Regs3{ldp(r, t), ldc(diag, t), ldp(p, t)}; }
	__device__ void apply(const Item &t, const Regs3 &v, double *) const
	{
		stp(p, t, make_double2(madd(beta, v.c.x, __ddiv_rn(v.a.x, v.b.x)), madd(beta, v.c.y, __ddiv_rn(v.a.y, v.b.y))));
	}
};

struct OpBicgDir {   // p = r + beta (p - omega v)
	static constexpr int NR = 0;
	static constexpr bool SPLIT = true;
	static constexpr int UNROLL = 4;
	s3d_scalar sbeta, somega; const double *r, *v; double *p; double beta, omega;
	__device__ void prepare() { beta = s3d_sval(sbeta); omega = s3d_sval(somega); }
	__device__ Regs3 load(const Item &t) const { return Regs3{ldp(r, t), ldp(v, t), ldp(p, t)}; }
	__device__ void apply(const Item &t, const Regs3 &q, double *) const
	{
		const double t0 = madd(-omega, q.b.x, q.c.x), t1 = madd(-omega, q.b.y, q.c.y);
		stp(p, t, make_double2(madd(beta, t0, q.a.x), madd(beta, t1, q.a.y)));
	}
};

struct OpBicgUpdate {   // x += alpha ph + omega sh; r = s - omega t; (r,r); (rhat,r)
	static constexpr int NR = 2;
	static constexpr bool SPLIT = true;
	static constexpr int UNROLL = 4;
	s3d_scalar salpha, somega; const double *ph, *sh, *s, *tt, *rhat; double *x, *r; double alpha, omega;
	__device__ void prepare() { alpha = s3d_sval(salpha); omega = s3d_sval(somega); }
	__device__ Regs6 load(const Item &t) const { return Regs6{ldp(ph, t), ldp(sh, t), ldp(s, t), ldp(tt, t), ldp(rhat, t), ldp(x, t)}; }
	__device__ void apply(const Item &t, const Regs6 &q, double *acc) const
	{
		const double2 xn = make_double2(madd(omega, q.b.x, madd(alpha, q.a.x, q.f.x)), madd(omega, q.b.y, madd(alpha, q.a.y, q.f.y)));
		const double2 rn = make_double2(madd(-omega, q.d.x, q.c.x), madd(-omega, q.d.y, q.c.y));
		stp(x, t, xn);
		stp(r, t, rn);
		acc[0] += rn.x * rn.x;
		acc[1] += q.e.x * rn.x;
		if (t.pair) { acc[0] += rn.y * rn.y; acc[1] += q.e.y * rn.y; }
	}
};

struct OpRecip {   // dinv = 1/diag over the coefficient layout (s3d_sgspreconditioners.cpp:126-134)
	static constexpr int NR = 0;
	static constexpr bool SPLIT = false;
	const double *diag; double *dinv;
	__device__ void prepare() {}
	__device__ void operator()(const Item &t, double *) const
	{
		dinv[t.c] = __ddiv_rn(1.0, diag[t.c]);
		if (t.pair) dinv[t.c + 1] = __ddiv_rn(1.0, diag[t.c + 1]);
	}
};

__global__ void converge_test_kernel(const double *rr, const double *bb, double rtol, int maxiter, int flags,
                                     int *stop, int *iters, double *hist)
{
	if (*stop != 0) return;
	const double rel = sqrt(*rr) / sqrt(*bb);
	int it = *iters;
	if (!(flags & 2)) {   // a completed iteration: record and count it
		if (hist) hist[it] = rel;
		it += 1;
		*iters = it;
	}
	const bool done = (flags & 1) ? (rel < rtol) : !(rel > rtol);
	if (done) *stop = 1;
	else if (it >= maxiter) *stop = 3;
}

__global__ void scalar_op_kernel(int op, double *out, const double *a, const double *b, const double *c, const double *d,
                                 int n, double scale, int *stop)
{
	const int l = threadIdx.x;
	if (l >= n) return;
	if (op == 0) {   // out[l] = scale * a[l] / b[l]
		out[l] = scale * a[l] / b[l];
	} else if (op == 1) {   // out = (a/b) * (c/d); a == 0 or a breakdown denominator flags stop = 2
		if (stop && (*a == 0.0 || *b == 0.0 || *d == 0.0) && *stop == 0) *stop = 2;
		out[0] = (*a / *b) * (*c / *d);
	} else if (op == 3) {   // out[l] = a[l]
		out[l] = a[l];
	} else if (op == 2) {   // out = a / b with breakdown check on b
		if (stop && *b == 0.0 && *stop == 0) *stop = 2;
		out[0] = *a / *b;
	}
}

template <int N>
static int multi_dot_n(s3d_ctx *ctx, const s3d_grid *g, int n, const double *const *xs, const double *d_y, double *d_red)
{
	OpMultiDot<N> op;
	op.n = n; op.y = d_y;
	for (int l = 0; l < MAXV; l++) op.xs.p[l] = xs[l < n ? l : 0];
	// the kernel writes N outputs; route them through scratch when the caller's buffer holds only n
	if (n == N) return run(ctx, g, op, d_red, nullptr);
	const int rc = run(ctx, g, op, ctx->red_tmp, nullptr);
	if (rc != S3D_OK) return rc;
	S3D_CUDA(ctx, cudaMemcpyAsync(d_red, ctx->red_tmp, n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
	return S3D_OK;
}

// Stage 2 of the SpMV reductions: adds the per-block partials (up to half a million of them at 1024^3) in a
// fixed order.  64 blocks each add a strided share, then the ticket-elected last block adds the 64 results:
// deterministic, and ~10x faster than one block walking megabytes of partials.
template <int NR>
__global__ void __launch_bounds__(256) final_reduce_kernel(const double *__restrict__ partials, unsigned n, const RedOut ro, const int *stop)
{
	if (s3d_stopped(stop)) return;
	double acc[NR];
#pragma unroll
	for (int r = 0; r < NR; r++) acc[r] = 0.0;
	for (unsigned b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
#pragma unroll
		for (int r = 0; r < NR; r++) acc[r] += partials[(size_t)b * NR + r];
	}
	grid_reduce<NR>(acc, ro);
}

}  // namespace

int s3d_final_reduce(s3d_ctx *ctx, int nr, unsigned nblocks, double *out)
{
	if (nr != 2) return s3d_fail(ctx, S3D_ERR_ARG, "final reduce supports 2 values per block");
	const unsigned nb = nblocks >= 64u * 256u ? 64u : (nblocks + 255u) / 256u;
	// the second-level partials live behind the first-level ones
	const int rc = s3d_ensure_partials(ctx, (size_t)nblocks * nr + (size_t)nb * nr);
	if (rc != S3D_OK) return rc;
	RedOut ro{ctx->partials + (size_t)nblocks * nr, ctx->ticket, out};
	final_reduce_kernel<2><<<nb, 256, 0, ctx->stream>>>(ctx->partials, nblocks, ro, ctx->gate);
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}

extern "C" {

S3D_HIDDEN void s3d_internal_set_ew_unroll(int u) { g_ew_unroll = u; }

int s3d_vecset(s3d_ctx *ctx, const s3d_grid *g, double a, double *d_x)
{
	S3D_REQUIRE(ctx, g && d_x, "null argument");
	return run(ctx, g, OpSet{a, d_x}, nullptr, nullptr);
}

int s3d_vecassign(s3d_ctx *ctx, const s3d_grid *g, const double *d_x, double *d_y)
{
	S3D_REQUIRE(ctx, g && d_x && d_y, "null argument");
	return run(ctx, g, OpCopy{d_x, d_y}, nullptr, nullptr);
}

int s3d_veccopy_all(s3d_ctx *ctx, const s3d_grid *g, const double *d_x, double *d_y)
{
	S3D_REQUIRE(ctx, g && d_x && d_y, "null argument");
	S3D_CUDA(ctx, cudaMemcpyAsync(d_y, d_x, (size_t)g->vlen * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
	return S3D_OK;
}

int s3d_vecaxpy(s3d_ctx *ctx, const s3d_grid *g, s3d_scalar a, const double *d_x, double *d_y)
{
	S3D_REQUIRE(ctx, g && d_x && d_y, "null argument");
	return run(ctx, g, OpAxpy{a, d_x, d_y, 0.0}, nullptr, nullptr);
}

int s3d_vecaxpby(s3d_ctx *ctx, const s3d_grid *g, s3d_scalar a, const double *d_x, s3d_scalar b, double *d_y)
{
	S3D_REQUIRE(ctx, g && d_x && d_y, "null argument");
	const bool bzero = (b.imm == 0.0);
	const bool aone = (a.imm == 1.0 && !a.num && !a.den);
	if (bzero) return run(ctx, g, OpWaxpby<false, false>{a, b, d_x, nullptr, d_y, 0, 0}, nullptr, nullptr);
	if (aone) return run(ctx, g, OpXpby{b, d_x, d_y, 0.0}, nullptr, nullptr);
	return run(ctx, g, OpWaxpby<true, false>{a, b, d_x, d_y, d_y, 0, 0}, nullptr, nullptr);
}

int s3d_vecwaxpby(s3d_ctx *ctx, const s3d_grid *g, s3d_scalar a, const double *d_x, s3d_scalar b, const double *d_y,
                  double *d_w, double *d_red)
{
	S3D_REQUIRE(ctx, g && d_x && d_y && d_w, "null argument");
	if (d_red) return run(ctx, g, OpWaxpby<true, true>{a, b, d_x, d_y, d_w, 0, 0}, d_red, nullptr);
	return run(ctx, g, OpWaxpby<true, false>{a, b, d_x, d_y, d_w, 0, 0}, nullptr, nullptr);
}

int s3d_vec_multi_axpy(s3d_ctx *ctx, const s3d_grid *g, int n, const double *d_a, const double *const *xs, double *d_y)
{
	S3D_REQUIRE(ctx, g && d_y && n >= 0 && n <= MAXV, "bad argument (n <= 32)");
	if (n == 0) return S3D_OK;
	S3D_REQUIRE(ctx, d_a && xs, "null argument");
	OpMultiAxpy op;
	op.n = n; op.a = d_a; op.y = d_y;
	for (int l = 0; l < MAXV; l++) op.xs.p[l] = xs[l < n ? l : 0];
	return run(ctx, g, op, nullptr, nullptr);
}

int s3d_vec_multi_dot(s3d_ctx *ctx, const s3d_grid *g, int n, const double *const *xs, const double *d_y, double *d_red)
{
	S3D_REQUIRE(ctx, g && d_y && xs && d_red && n >= 1 && n <= MAXV, "bad argument (1 <= n <= 32)");
	if (n <= 2) return multi_dot_n<2>(ctx, g, n, xs, d_y, d_red);
	if (n <= 4) return multi_dot_n<4>(ctx, g, n, xs, d_y, d_red);
	if (n <= 8) return multi_dot_n<8>(ctx, g, n, xs, d_y, d_red);
	if (n <= 16) return multi_dot_n<16>(ctx, g, n, xs, d_y, d_red);
	return multi_dot_n<32>(ctx, g, n, xs, d_y, d_red);
}

int s3d_dot(s3d_ctx *ctx, const s3d_grid *g, const double *d_x, const double *d_y, double *d_red)
{
	S3D_REQUIRE(ctx, g && d_x && d_y && d_red, "null argument");
	return run(ctx, g, OpDot{d_x, d_y}, d_red, nullptr);
}

int s3d_nrm2sq(s3d_ctx *ctx, const s3d_grid *g, const double *d_x, double *d_red)
{
	S3D_REQUIRE(ctx, g && d_x && d_red, "null argument");
	return run(ctx, g, OpNrm2{d_x}, d_red, nullptr);
}

int s3d_normL2sq(s3d_ctx *ctx, const s3d_grid *g, const double *h_cx, const double *h_cy, const double *h_cz,
                 const double *d_x, double *d_red)
{
	S3D_REQUIRE(ctx, g && h_cx && h_cy && h_cz && d_x && d_red, "null argument");
	const size_t n = (size_t)g->nx + g->ny + g->nz;
	int rc = s3d_ensure_tab(ctx, n);
	if (rc != S3D_OK) return rc;
	double *h = (double *)malloc(n * sizeof(double));
	size_t o = 0;
	for (int i = 0; i < g->nx; i++) h[o++] = h_cx[i + 2] - h_cx[i];
	for (int j = 0; j < g->ny; j++) h[o++] = h_cy[j + 2] - h_cy[j];
	for (int k = 0; k < g->nz; k++) h[o++] = h_cz[g->k0 + k + 2] - h_cz[g->k0 + k];
	cudaError_t e = cudaMemcpyAsync(ctx->tab, h, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
	free(h);
	S3D_CUDA(ctx, e);
	return run(ctx, g, OpNormL2{d_x, ctx->tab, ctx->tab + g->nx, ctx->tab + g->nx + g->ny}, d_red, nullptr);
}

int s3d_dot_host(s3d_ctx *ctx, const s3d_grid *g, const double *d_x, const double *d_y, double *h_out)
{
	int rc = s3d_dot(ctx, g, d_x, d_y, ctx->red_tmp);
	if (rc == S3D_OK) rc = s3d_allreduce_sum(ctx, ctx->red_tmp, ctx->red_tmp + 32, 1);
	if (rc == S3D_OK) rc = s3d_scalars_download(ctx, ctx->red_tmp + 32, 1, h_out);
	return rc;
}

int s3d_nrm2_host(s3d_ctx *ctx, const s3d_grid *g, const double *d_x, double *h_out)
{
	int rc = s3d_nrm2sq(ctx, g, d_x, ctx->red_tmp);
	if (rc == S3D_OK) rc = s3d_allreduce_sum(ctx, ctx->red_tmp, ctx->red_tmp + 32, 1);
	if (rc == S3D_OK) rc = s3d_scalars_download(ctx, ctx->red_tmp + 32, 1, h_out);
	if (rc == S3D_OK) *h_out = sqrt(*h_out);
	return rc;
}

int s3d_jacobi_apply(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *d_r, double *d_z)
{
	S3D_REQUIRE(ctx, g && A && d_r && d_z, "null argument");
	return run(ctx, g, OpJacobi{A + (size_t)S3D_STENCIL_DIAG * g->cstride, d_r, d_z}, nullptr, nullptr);
}

int s3d_sgs_setup(s3d_ctx *ctx, const s3d_grid *g, const double *A, double *d_dinv)
{
	S3D_REQUIRE(ctx, g && A && d_dinv, "null argument");
	return run(ctx, g, OpRecip{A + (size_t)S3D_STENCIL_DIAG * g->cstride, d_dinv}, nullptr, nullptr);
}

int s3d_cg_update(s3d_ctx *ctx, const s3d_grid *g, s3d_scalar alpha, const double *d_p, const double *d_q, double *d_x,
                  double *d_r, const double *A_or_null, double *d_red, const int *d_stop)
{
	S3D_REQUIRE(ctx, g && d_p && d_q && d_x && d_r && d_red, "null argument");
	if (A_or_null)
		return run(ctx, g, OpCgUpdate<true>{alpha, d_p, d_q, d_x, d_r, A_or_null + (size_t)S3D_STENCIL_DIAG * g->cstride, 0.0}, d_red, d_stop);
	return run(ctx, g, OpCgUpdate<false>{alpha, d_p, d_q, d_x, d_r, nullptr, 0.0}, d_red, d_stop);
}

int s3d_cg_direction_jacobi(s3d_ctx *ctx, const s3d_grid *g, s3d_scalar beta, const double *A, const double *d_r,
                            double *d_p, const int *d_stop)
{
	S3D_REQUIRE(ctx, g && A && d_r && d_p, "null argument");
	return run(ctx, g, OpCgDirJacobi{beta, A + (size_t)S3D_STENCIL_DIAG * g->cstride, d_r, d_p, 0.0}, nullptr, d_stop);
}

int s3d_bicg_direction(s3d_ctx *ctx, const s3d_grid *g, s3d_scalar beta, s3d_scalar omega, const double *d_r,
                       const double *d_v, double *d_p, const int *d_stop)
{
	S3D_REQUIRE(ctx, g && d_r && d_v && d_p, "null argument");
	return run(ctx, g, OpBicgDir{beta, omega, d_r, d_v, d_p, 0.0, 0.0}, nullptr, d_stop);
}

int s3d_bicg_update(s3d_ctx *ctx, const s3d_grid *g, s3d_scalar alpha, const double *d_ph, s3d_scalar omega,
                    const double *d_sh, const double *d_s, const double *d_t, const double *d_rhat, double *d_x,
                    double *d_r, double *d_red, const int *d_stop)
{
	S3D_REQUIRE(ctx, g && d_ph && d_sh && d_s && d_t && d_rhat && d_x && d_r && d_red, "null argument");
	return run(ctx, g, OpBicgUpdate{alpha, omega, d_ph, d_sh, d_s, d_t, d_rhat, d_x, d_r, 0.0, 0.0}, d_red, d_stop);
}

int s3d_converge_test(s3d_ctx *ctx, const double *d_rr, const double *d_bb, double rtol, int maxiter, int flags,
                      int *d_stop, int *d_iters, double *d_hist_or_null)
{
	S3D_REQUIRE(ctx, d_rr && d_bb && d_stop && d_iters, "null argument");
	converge_test_kernel<<<1, 1, 0, ctx->stream>>>(d_rr, d_bb, rtol, maxiter, flags, d_stop, d_iters, d_hist_or_null);
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}

int s3d_scalar_op(s3d_ctx *ctx, int op, double *d_out, const double *a, const double *b, const double *c,
                  const double *d, int n, double scale, int *d_stop)
{
	S3D_REQUIRE(ctx, d_out && a && b && n >= 1 && n <= 64, "bad argument");
	scalar_op_kernel<<<1, 64, 0, ctx->stream>>>(op, d_out, a, b, c, d, n, scale, d_stop);
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}

}  // extern "C"
