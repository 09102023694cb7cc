// Lexicographic Gauss-Seidel sweeps as a ROW-SCAN wavefront for sm_100a
// (reference: SGS_like_preconditioner::apply, src/linalg/s3d_sgspreconditioners.cpp:53-110).
//
// Every other sweep kernel of this library walks along i one column at a time, because the reference's expression makes a point
// depend on its i-neighbour.  But along a row that dependence is a first-order LINEAR recurrence,
//     forward :  y_i = a_i + b_i y_{i-1},   a_i = (r_i - A1 y[j-1] - A2 y[k-1]) dinv_i,   b_i = -A0_i dinv_i
//     backward:  z_i = a_i + b_i z_{i+1},   a_i = y_i - dinv_i (A5 z[j+1] + A6 z[k+1]),   b_i = -dinv_i A4_i
// and a warp solves 32 columns of it at once with an affine prefix scan (five shuffle steps).  The unit of the dataflow becomes a
// SEGMENT of 32 columns of one (j,k) row:
//   * a warp owns a row and walks its segments; lanes are columns, so every access is a coalesced 256-byte row piece (no staging, no
//     shear, 5 doubles of prefetch registers per stage);
//   * segment s of row (j,k) needs segment s of rows (j-1,k) and (j,k-1) and the carry of its own segment s-1: a wavefront over
//     (s, j, k) of depth nx/32 + ny + nz whose step is one scan (~300 cycles) for 32 columns, where the column-by-column kernels pay
//     ~130 cycles for ONE column and nx + ny + nz steps;
//   * a block of WJ x WK warps owns a LINE of WJ x WK rows; inside the block the segments travel through shared-memory rings, between
//     blocks through mailboxes in global memory.  The rings are full / empty mbarrier pipelines (a waiting warp sleeps in hardware;
//     polled rings made the waiting warps of a block slow the working ones down threefold); the mailboxes use the data as its own
//     flag (an empty slot holds a signalling-NaN pattern no arithmetic produces, the reader re-arms it).  Every warp walks the
//     block's line queue on its own; lines are drawn from a ticket counter in (cj + ck) order.
// The scan adds the terms of a row in another order than a sequential walk: the result is the reference's sweep TO ROUNDING (1e-13
// relative on the test grids), not bit for bit -- which is what north_star asks of the SGS (same residual, iteration counts within
// 5 %).  MEASURED SLOWER than the tile kernel (every step along j or k is a hand-off between two warps, ~2500 cycles at 256^3 however
// it is carried; profiles/r2_sgs_rowscan.md): opt-in through the tuning key sgs_variant = 7; the bit-identical kernels stay the
// default and keep running the StrILU build, whose recurrence is not linear.
#include <algorithm>
#include <type_traits>
#include <vector>
#include "common.cuh"

namespace {

constexpr unsigned long long SR_SENT = 0xFFF4A5C3D2E1B497ull;   // "not written yet" (shared with the other mailbox kernels)
constexpr int SR_RING = 8;                                   // segments per shared-memory ring
constexpr int SR_LQ = 8;                                     // entries of the line queue

struct RowArgs {
	int nx, ny, nz, nseg;
	int vpitch, cpitch;
	long long vplane, voff, cplane;
	int ncj, nck;
	const int *order;
	unsigned *ticket;
	unsigned long long *mail_j;    // [line][WK rows][32 nseg]  the line's last-j rows
	unsigned long long *mail_k;    // [line][WJ rows][32 nseg]  the line's last-k rows
	const double *in0, *c0, *c1, *c2, *dv;
	double *out;
	int klo_zero, khi_zero;
	int debug;                     // timing experiments only (results become wrong): 1 = no hand-off between warps and blocks
};

__device__ __forceinline__ unsigned sr_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long lds_u64(unsigned addr)
{
	unsigned long long v;
	asm volatile("ld.volatile.shared.u64 %0, [%1];\n" : "=l"(v) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ void sts_u64(unsigned addr, unsigned long long v) { asm volatile("st.volatile.shared.u64 [%0], %1;\n" ::"r"(addr), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ldcg_u64(const unsigned long long *p)
{
	unsigned long long v;
	asm volatile("ld.global.cg.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void stcg_u64(unsigned long long *p, unsigned long long v) { asm volatile("st.global.cg.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void sr_mbar_init(unsigned a, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory"); }
__device__ __forceinline__ void sr_mbar_arrive(unsigned a) { asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];\n" ::"r"(a) : "memory"); }
// waits in hardware (the thread is suspended up to a system time limit per try): no issue slots burnt, unlike a polling loop --
// with polled rings the waiting warps of a block slowed the working ones down threefold (profiles/r2_sgs_rowscan.md)
__device__ __forceinline__ bool sr_mbar_try(unsigned a, unsigned parity)
{
	unsigned ok;
	asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ int sr_vol_s32(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ void sr_spin_check(long long &t0)
{
	const long long t = clock64();
	if (t0 == 0) t0 = t;
	else if (t - t0 > (8ll << 30)) __trap();   // seconds of waiting: a broken protocol; fail loudly instead of hanging the GPU
}
// a wait that does not end means a broken protocol: fail loudly instead of hanging the GPU (a try suspends the thread for up to
// microseconds, a mailbox poll is an L2 round trip: 2^24 of either are seconds)
__device__ __forceinline__ void sr_tries_check(int &tries) { if (++tries > (1 << 24)) __trap(); }
__device__ __forceinline__ void sr_mbar_wait(unsigned a, unsigned parity)
{
	for (int tries = 0; !sr_mbar_try(a, parity); sr_tries_check(tries)) { }
}

// MODE 0: forward sweep, MODE 1: backward sweep.  Face kinds (warp-uniform): 0 fixed data (ghost row / plane) or zero, 1 shared-memory
// ring, 2 global mailbox.
template <int MODE, int WJ, int WK, int D>
__global__ void __launch_bounds__(32 * WJ * WK)
sgs_rowscan_kernel(const RowArgs a, const int *stop)
{
	constexpr bool FWD = (MODE == 0);
	constexpr int NW = WJ * WK;
	constexpr int RB = SR_RING * 256;                            // bytes of a ring
	constexpr unsigned FULL = 0xffffffffu;
	if (s3d_stopped(stop)) return;

	extern __shared__ __align__(16) unsigned long long sr_smem[];
	unsigned long long *s_rj = sr_smem;                                  // [NW][RING][32]  segments of the j-predecessor row, INTO warp w
	unsigned long long *s_rk = s_rj + (size_t)NW * SR_RING * 32;         // [NW][RING][32]  segments of the k-predecessor row
	unsigned long long *s_bar = s_rk + (size_t)NW * SR_RING * 32;        // [NW][4][RING] mbarriers: full j, full k, empty j, empty k of warp w's rings
	int *s_lineq = reinterpret_cast<int *>(s_bar + (size_t)NW * 4 * SR_RING);   // [LQ]
	int *s_pub = s_lineq + SR_LQ;
	int *s_wseq = s_pub + 1;                                             // [NW]

	for (int t = threadIdx.x; t < NW * 4 * SR_RING; t += blockDim.x) sr_mbar_init(sr_u32(s_bar + t), 1);   // lane 0 arrives for its warp, behind a __syncwarp
	if (threadIdx.x == 0) *s_pub = 0;
	if (threadIdx.x < NW) s_wseq[threadIdx.x] = 0;
	asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
	__syncthreads();

	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int wj = w % WJ, wk = w / WJ;
	const int nlines = a.ncj * a.nck;
	const bool drawer = (FWD ? w == 0 : w == NW - 1);         // the warp that is first in sweep order fills the line queue
	constexpr int step = FWD ? 32 : -32;                         // doubles per segment along a row
	unsigned jin = 0, kin = 0, jout = 0, kout = 0;               // segments that went through the rings before the current line
	int next_line = -1;
	if (drawer) {
		if (lane == 0) next_line = ((int)blockIdx.x < nlines) ? a.order[FWD ? blockIdx.x : nlines - 1 - blockIdx.x] : -1;
		next_line = __shfl_sync(FULL, next_line, 0);
	}

	for (int seq = 0;; seq++) {
		// ------------------------------------------------------------- the next line of the block
		int line, tkt = 0;
		if (drawer) {
			line = next_line;
			if (lane == 0) {
				if (NW > 1) {
					if (seq >= SR_LQ) {
						for (long long t0 = 0;; sr_spin_check(t0)) {
							int lo = seq;
							for (int x = 0; x < NW; x++) if (x != w) lo = min(lo, sr_vol_s32(s_wseq + x));
							if (lo >= seq - SR_LQ + 1) break;
						}
					}
					*reinterpret_cast<volatile int *>(s_lineq + (seq % SR_LQ)) = line;
					__threadfence_block();
					*reinterpret_cast<volatile int *>(s_pub) = seq + 1;
				}
				if (line >= 0) tkt = (int)atomicAdd(a.ticket, 1u) + (int)gridDim.x;
			}
		} else {
			for (long long t0 = 0; sr_vol_s32(s_pub) <= seq; sr_spin_check(t0)) { }
			line = sr_vol_s32(s_lineq + (seq % SR_LQ));
			__syncwarp();
			if (lane == 0) *reinterpret_cast<volatile int *>(s_wseq + w) = seq + 1;
		}
		if (line < 0) break;

		const int cj = line % a.ncj, ck = line / a.ncj;
		const int j = cj * WJ + wj, k = ck * WK + wk;
		if (j < a.ny && k < a.nz) {
			const long long vrow = a.voff + (long long)k * a.vplane + (long long)j * a.vpitch;
			const long long crow = (long long)k * a.cplane + (long long)j * a.cpitch;
			const bool nolink = (a.debug & 1) != 0;
			// this lane's column in the first segment of the sweep; lane 0 is always first in sweep order
			const int col_first = FWD ? lane : (a.nseg - 1) * 32 + (31 - lane);
			// ---- where the neighbour rows come from and where this row goes (kinds are warp-uniform)
			const bool jpred_in = FWD ? (j > 0) : (j + 1 < a.ny), jpred_blk = FWD ? (wj > 0) : (wj < WJ - 1);
			const bool kpred_in = FWD ? (k > 0) : (k + 1 < a.nz), kpred_blk = FWD ? (wk > 0) : (wk < WK - 1);
			const bool jsucc_in = FWD ? (j + 1 < a.ny) : (j > 0), jsucc_blk = FWD ? (wj < WJ - 1) : (wj > 0);
			const bool ksucc_in = FWD ? (k + 1 < a.nz) : (k > 0), ksucc_blk = FWD ? (wk < WK - 1) : (wk > 0);
			const int jsrc_kind = (!jpred_in || nolink) ? 0 : jpred_blk ? 1 : 2;
			const int ksrc_kind = (!kpred_in || nolink) ? 0 : kpred_blk ? 1 : 2;
			const int jdst_kind = (!jsucc_in || nolink) ? 0 : jsucc_blk ? 1 : 2;
			const int kdst_kind = (!ksucc_in || nolink) ? 0 : ksucc_blk ? 1 : 2;
			const bool kzero = (!kpred_in && (FWD ? a.klo_zero : a.khi_zero)) || (nolink && kpred_in);
			const bool jzero = nolink && jpred_in;
			const size_t mrow = (size_t)(32 * a.nseg);
			const unsigned long long *jsrc, *ksrc;                       // kinds 0 and 2: this lane's element of the first segment
			if (jsrc_kind == 2) jsrc = a.mail_j + ((size_t)(FWD ? line - 1 : line + 1) * WK + wk) * mrow + col_first;
			else jsrc = reinterpret_cast<const unsigned long long *>(a.out + a.voff + (long long)k * a.vplane + (long long)(FWD ? j - 1 : j + 1) * a.vpitch) + col_first;
			if (ksrc_kind == 2) ksrc = a.mail_k + ((size_t)(FWD ? line - a.ncj : line + a.ncj) * WJ + wj) * mrow + col_first;
			else ksrc = reinterpret_cast<const unsigned long long *>(a.out + a.voff + (long long)(FWD ? k - 1 : k + 1) * a.vplane + (long long)j * a.vpitch) + col_first;
			unsigned long long *jdst = a.mail_j + ((size_t)line * WK + wk) * mrow + col_first;
			unsigned long long *kdst = a.mail_k + ((size_t)line * WJ + wj) * mrow + col_first;
			const int wsj = (jdst_kind == 1) ? (FWD ? w + 1 : w - 1) : w, wsk = (kdst_kind == 1) ? (FWD ? w + WJ : w - WJ) : w;
			const unsigned jring_in = sr_u32(s_rj) + (unsigned)(w * RB) + lane * 8, kring_in = sr_u32(s_rk) + (unsigned)(w * RB) + lane * 8;
			const unsigned jring_out = sr_u32(s_rj) + (unsigned)(wsj * RB) + lane * 8, kring_out = sr_u32(s_rk) + (unsigned)(wsk * RB) + lane * 8;
			// barriers of the rings this warp reads (its own) and writes (its successors'): [full j, full k, empty j, empty k][RING]
			const unsigned bar_in = sr_u32(s_bar + (size_t)w * 4 * SR_RING);
			const unsigned jbar_out = sr_u32(s_bar + (size_t)wsj * 4 * SR_RING), kbar_out = sr_u32(s_bar + (size_t)wsk * 4 * SR_RING);

			// ---- the register pipeline: inputs and (kinds 0, 2) neighbour values of segment s sit in stage s % D.
			// Addresses are per-lane bases plus one 32-bit element offset (step * s): one IMAD.WIDE per access.
			const double *p_in = a.in0 + vrow + col_first, *p_c0 = a.c0 + crow + col_first, *p_c1 = a.c1 + crow + col_first;
			const double *p_c2 = a.c2 + crow + col_first, *p_dv = a.dv + crow + col_first;
			double *p_out = a.out + vrow + col_first;
			double s_f[D], s_x0[D], s_x1[D], s_x2[D], s_dv[D];
			unsigned long long s_nj[D], s_nk[D];
			// sweep-order segments [full0, full1) lie wholly inside the row: no lane needs a range check there
			const int nfull = a.nx >> 5;
			const int full0 = FWD ? 0 : a.nseg - nfull, full1 = FWD ? nfull : a.nseg;
			const bool jget = (jsrc_kind != 1) && !jzero, kget = (ksrc_kind != 1) && !kzero;   // neighbour values fetched with the inputs
			auto inside = [&](int s) { const int col = col_first + step * s; return (unsigned)col < (unsigned)a.nx; };
			auto fetch = [&](auto u_tag, int s) {
				constexpr int u = decltype(u_tag)::value;
				const int o = step * s;
				if (s >= full0 && s < full1) {                         // warp-uniform
					s_f[u] = __ldg(p_in + o);
					s_x0[u] = ld_stream(p_c0 + o);
					s_x1[u] = ld_stream(p_c1 + o);
					s_x2[u] = ld_stream(p_c2 + o);
					s_dv[u] = __ldg(p_dv + o);
					if (jget) s_nj[u] = ldcg_u64(jsrc + o);
					if (kget) s_nk[u] = ldcg_u64(ksrc + o);
				} else {
					const bool on = s < a.nseg && inside(s);
					s_f[u] = on ? __ldg(p_in + o) : 0.0;
					s_x0[u] = on ? ld_stream(p_c0 + o) : 0.0;
					s_x1[u] = on ? ld_stream(p_c1 + o) : 0.0;
					s_x2[u] = on ? ld_stream(p_c2 + o) : 0.0;
					s_dv[u] = on ? __ldg(p_dv + o) : 0.0;
					s_nj[u] = 0; s_nk[u] = 0;
					if (jget && on) s_nj[u] = ldcg_u64(jsrc + o);
					if (kget && on) s_nk[u] = ldcg_u64(ksrc + o);
				}
			};
#pragma unroll
			for (int u = 0; u < D; u++) { s_nj[u] = 0; s_nk[u] = 0; }
			fetch(std::integral_constant<int, 0>{}, 0);
			if (D > 1) fetch(std::integral_constant<int, (D > 1 ? 1 : 0)>{}, 1);
			if (D > 2) fetch(std::integral_constant<int, (D > 2 ? 2 : 0)>{}, 2);
			double carry = __ldcg(a.out + vrow + (FWD ? -1 : a.nx));   // the ghost column: i-neighbour of the first column
			// ring slots and barrier parities of this line's first segment; they advance by one per segment
			unsigned nji = jin, nki = kin, njo = jout, nko = kout;

			// one segment.  WHOLE: all 32 columns lie inside the row (warp-uniform), nothing is predicated
			auto segment = [&](auto u_tag, auto whole_tag, int s) {
				constexpr int u = decltype(u_tag)::value;
				constexpr bool WHOLE = decltype(whole_tag)::value;
				const bool in = WHOLE ? true : inside(s);
				const int o = step * s;
				// ---- the neighbour rows' segments
				unsigned long long bj = s_nj[u], bk = s_nk[u];
				if (jsrc_kind == 1) {
					const unsigned sl = nji & (SR_RING - 1);
					sr_mbar_wait(bar_in + sl * 8, (nji / SR_RING) & 1);                     // full
					bj = lds_u64(jring_in + sl * 256);
					__syncwarp();
					if (lane == 0) sr_mbar_arrive(bar_in + (2 * SR_RING + sl) * 8);          // empty again
					nji++;
				} else if (jsrc_kind == 2) {
					for (int tries = 0; !__all_sync(FULL, !in || bj != SR_SENT); sr_tries_check(tries)) if (in && bj == SR_SENT) bj = ldcg_u64(jsrc + o);
					if (in) stcg_u64(const_cast<unsigned long long *>(jsrc + o), SR_SENT);   // clean for the next sweep
				}
				if (ksrc_kind == 1) {
					const unsigned sl = nki & (SR_RING - 1);
					sr_mbar_wait(bar_in + (SR_RING + sl) * 8, (nki / SR_RING) & 1);
					bk = lds_u64(kring_in + sl * 256);
					__syncwarp();
					if (lane == 0) sr_mbar_arrive(bar_in + (3 * SR_RING + sl) * 8);
					nki++;
				} else if (ksrc_kind == 2) {
					for (int tries = 0; !__all_sync(FULL, !in || bk != SR_SENT); sr_tries_check(tries)) if (in && bk == SR_SENT) bk = ldcg_u64(ksrc + o);
					if (in) stcg_u64(const_cast<unsigned long long *>(ksrc + o), SR_SENT);
				}
				const double nj = __longlong_as_double((long long)bj), nk = __longlong_as_double((long long)bk);
				// ---- this column's affine map; columns outside the row are the identity, so the carry passes through them
				double ca, cb;
				if (FWD) {
					ca = __dmul_rn(__dsub_rn(__dsub_rn(s_f[u], __dmul_rn(s_x1[u], nj)), __dmul_rn(s_x2[u], nk)), s_dv[u]);
					cb = -__dmul_rn(s_x0[u], s_dv[u]);
				} else {
					ca = __dsub_rn(s_f[u], __dmul_rn(s_dv[u], __dadd_rn(__dmul_rn(s_x1[u], nj), __dmul_rn(s_x2[u], nk))));
					cb = -__dmul_rn(s_dv[u], s_x0[u]);
				}
				if (!WHOLE && !in) { ca = 0.0; cb = 1.0; }
				// the inputs of segment s + D can fly from here on
				fetch(u_tag, s + D);
				// ---- inclusive scan of the maps v -> ca + cb v over the lanes: afterwards value(lane) = ca + cb * carry
#pragma unroll
				for (int d = 1; d < 32; d <<= 1) {
					const double ap = __shfl_up_sync(FULL, ca, d), bp = __shfl_up_sync(FULL, cb, d);
					if (lane >= d) { ca = fma(cb, ap, ca); cb = cb * bp; }
				}
				const double v = fma(cb, carry, ca);
				carry = __shfl_sync(FULL, v, 31);
				// ---- out it goes: the result vector, and the rows behind (in sweep order)
				const unsigned long long bv = (unsigned long long)__double_as_longlong(v);
				if (in) __stcg(p_out + o, v);
				if (jdst_kind == 1) {
					const unsigned sl = njo & (SR_RING - 1);
					sr_mbar_wait(jbar_out + (2 * SR_RING + sl) * 8, ((njo / SR_RING) & 1) ^ 1);   // the slot's last content has been read
					sts_u64(jring_out + sl * 256, bv);
					__syncwarp();
					if (lane == 0) sr_mbar_arrive(jbar_out + sl * 8);
					njo++;
				} else if (jdst_kind == 2) {
					if (in) stcg_u64(jdst + o, bv);
				}
				if (kdst_kind == 1) {
					const unsigned sl = nko & (SR_RING - 1);
					sr_mbar_wait(kbar_out + (3 * SR_RING + sl) * 8, ((nko / SR_RING) & 1) ^ 1);
					sts_u64(kring_out + sl * 256, bv);
					__syncwarp();
					if (lane == 0) sr_mbar_arrive(kbar_out + (SR_RING + sl) * 8);
					nko++;
				} else if (kdst_kind == 2) {
					if (in) stcg_u64(kdst + o, bv);
				}
			};

#pragma unroll 1
			for (int s0 = 0; s0 < a.nseg; s0 += D) {
				if (s0 >= full0 && s0 + D <= full1) {
					segment(std::integral_constant<int, 0>{}, std::true_type{}, s0);
					if (D > 1) segment(std::integral_constant<int, (D > 1 ? 1 : 0)>{}, std::true_type{}, s0 + 1);
					if (D > 2) segment(std::integral_constant<int, (D > 2 ? 2 : 0)>{}, std::true_type{}, s0 + 2);
				} else {
					segment(std::integral_constant<int, 0>{}, std::false_type{}, s0);
					if (D > 1 && s0 + 1 < a.nseg) segment(std::integral_constant<int, (D > 1 ? 1 : 0)>{}, std::false_type{}, s0 + 1);
					if (D > 2 && s0 + 2 < a.nseg) segment(std::integral_constant<int, (D > 2 ? 2 : 0)>{}, std::false_type{}, s0 + 2);
				}
			}
			if (jsrc_kind == 1) jin += (unsigned)a.nseg;
			if (ksrc_kind == 1) kin += (unsigned)a.nseg;
			if (jdst_kind == 1) jout += (unsigned)a.nseg;
			if (kdst_kind == 1) kout += (unsigned)a.nseg;
		}
		if (drawer) {
			if (lane == 0) next_line = (tkt < nlines) ? a.order[FWD ? tkt : nlines - 1 - tkt] : -1;
			next_line = __shfl_sync(FULL, next_line, 0);
		}
	}
}

template <int WJ, int WK> constexpr size_t sr_smem_bytes()
{
	return (size_t)WJ * WK * (2 * SR_RING * 32 + 4 * SR_RING) * sizeof(unsigned long long) + (size_t)(SR_LQ + 1 + WJ * WK) * sizeof(int);
}

__global__ void __launch_bounds__(256) sr_fill_u64_kernel(unsigned long long *p, size_t n, unsigned long long v)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

template <int WJ, int WK, int D>
int sr_launch(s3d_ctx *ctx, const s3d_grid *g, int mode, RowArgs &a, const int *stop)
{
	constexpr int NT = 32 * WJ * WK;
	constexpr size_t SM = sr_smem_bytes<WJ, WK>();
	static bool configured = false;
	static int blocks_per_sm = 1;
	if (!configured) {
		S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_rowscan_kernel<0, WJ, WK, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM));
		S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_rowscan_kernel<1, WJ, WK, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM));
		int x = 1, y = 1;
		S3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&x, sgs_rowscan_kernel<0, WJ, WK, D>, NT, SM));
		S3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&y, sgs_rowscan_kernel<1, WJ, WK, D>, NT, SM));
		blocks_per_sm = std::max(1, std::min(x, y));
		configured = true;
	}
	a.ncj = (g->ny + WJ - 1) / WJ; a.nck = (g->nz + WK - 1) / WK;
	const size_t nlines = (size_t)a.ncj * a.nck;
	const int okey = -(2000 + WJ * 64 + WK);   // no other kernel's dispatch table has this key
	if (ctx->sgs_order_dims[0] != okey || ctx->sgs_order_dims[1] != a.ncj || ctx->sgs_order_dims[2] != a.nck) {
		const int nd = a.ncj + a.nck - 1;
		std::vector<int> start(nd + 1, 0), order(nlines);
		for (int ck = 0; ck < a.nck; ck++)
			for (int cj = 0; cj < a.ncj; cj++) start[cj + ck + 1]++;
		for (int d = 0; d < nd; d++) start[d + 1] += start[d];
		for (int ck = 0; ck < a.nck; ck++)
			for (int cj = 0; cj < a.ncj; cj++) order[start[cj + ck]++] = ck * a.ncj + cj;
		S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		if (ctx->sgs_order) S3D_CUDA(ctx, cudaFree(ctx->sgs_order));
		ctx->sgs_order = nullptr;
		S3D_CUDA(ctx, cudaMalloc(&ctx->sgs_order, order.size() * sizeof(int)));
		S3D_CUDA(ctx, cudaMemcpy(ctx->sgs_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
		ctx->sgs_order_dims[0] = okey; ctx->sgs_order_dims[1] = a.ncj; ctx->sgs_order_dims[2] = a.nck;
	}
	a.order = ctx->sgs_order;
	// mailboxes: WJ + WK rows of 32 nseg values per line, all sentinel outside a sweep (the reader of a value puts the sentinel back)
	const size_t need = nlines * (size_t)(WJ + WK) * (size_t)(32 * a.nseg);
	if (ctx->sgs_mail_cap < need) {
		S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		if (ctx->sgs_mail) S3D_CUDA(ctx, cudaFree(ctx->sgs_mail));
		ctx->sgs_mail = nullptr; ctx->sgs_mail_cap = 0;
		S3D_CUDA(ctx, cudaMalloc(&ctx->sgs_mail, need * sizeof(unsigned long long)));
		sr_fill_u64_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->sgs_mail, need, SR_SENT);
		S3D_LAUNCHED(ctx);
		ctx->sgs_mail_cap = need;
	}
	a.mail_j = ctx->sgs_mail;
	a.mail_k = ctx->sgs_mail + nlines * (size_t)WK * (size_t)(32 * a.nseg);
	int bps = blocks_per_sm;
	if (ctx->sgs_warps_per_sm > 0) bps = std::max(1, std::min(bps, ctx->sgs_warps_per_sm / (WJ * WK)));
	const unsigned nblocks = (unsigned)std::min<size_t>(nlines, (size_t)ctx->sm_count * bps);
	if (mode == 0) sgs_rowscan_kernel<0, WJ, WK, D><<<nblocks, NT, SM, ctx->stream>>>(a, stop);
	else sgs_rowscan_kernel<1, WJ, WK, D><<<nblocks, NT, SM, ctx->stream>>>(a, stop);
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}

}  // namespace

// One sweep (mode 0 forward, 1 backward) of the row-scan kernel.  in0, out: vector layout; c0..c2, dinv: coefficient layout.  The
// caller has rewound the ticket counter (sgs_begin_kernel); sgs_pencil_shape selects the warps of a block (WJ * 100 + WK, 0 = default).
extern "C" S3D_HIDDEN int s3d_internal_sgs_rowscan(s3d_ctx *ctx, const s3d_grid *g, int mode, const double *in0, const double *c0,
                                                   const double *c1, const double *c2, const double *dinv, double *out,
                                                   unsigned *ticket, int klo_zero, int khi_zero, const int *stop)
{
	RowArgs a;
	a.nx = g->nx; a.ny = g->ny; a.nz = g->nz; a.nseg = (g->nx + 31) / 32;
	a.vpitch = g->vpitch; a.cpitch = g->cpitch; a.vplane = g->vplane; a.voff = g->voff; a.cplane = g->cplane;
	a.ticket = ticket;
	a.in0 = in0; a.c0 = c0; a.c1 = c1; a.c2 = c2; a.dv = dinv;
	a.out = out; a.klo_zero = klo_zero; a.khi_zero = khi_zero; a.debug = ctx->sgs_sleep_ns;
	switch (ctx->sgs_pencil_shape) {
	case 101: return sr_launch<1, 1, 2>(ctx, g, mode, a, stop);
	case 204: return sr_launch<2, 4, 2>(ctx, g, mode, a, stop);
	case 404: return sr_launch<4, 4, 2>(ctx, g, mode, a, stop);
	case 408: return sr_launch<4, 8, 2>(ctx, g, mode, a, stop);
	case 804: return sr_launch<8, 4, 2>(ctx, g, mode, a, stop);
	case 4043: return sr_launch<4, 4, 3>(ctx, g, mode, a, stop);
	case 4083: return sr_launch<4, 8, 3>(ctx, g, mode, a, stop);
	default: return sr_launch<4, 8, 2>(ctx, g, mode, a, stop);
	}
}
