// Exact lexicographic Gauss-Seidel sweeps as a STREAMING dataflow for sm_100a
// (reference: SGS_like_preconditioner::apply, src/linalg/s3d_sgspreconditioners.cpp:53-110, and the StrILU diagonal recurrence,
// src/linalg/s3d_ilu.cpp:34-57).  Same expressions, same roundings as the other sweep kernels in sgs.cu: bit-identical results.
//
// Why another kernel (measured in round 1, profiles/r1_sgs_phases.md): the tile-level dataflow hands work from tile to tile through
// global memory -- store rows, fence, flag, poll, fetch halo -- and a tile only starts when its three predecessors are COMPLETE, so the
// dependency chain of a sweep is (#tile hops) x (a whole tile + ~3500 cycles of hand-off): 78 x 7600 cycles at 256^3 against a
// fundamental chain of nx+ny+nz steps.  Here the hand-off is per COLUMN and needs neither fence nor flag:
//
//   * one warp owns a LINE: the 4 x 8 (j,k) pencils of a (cj, ck) cell over ALL nx columns.  Lane (m, c) marches along i one column
//     per step, skewed by m + c steps, its i-neighbour in a register, its j/k-neighbours from lanes -8 / -1 by shuffle.  No
//     i-direction hand-off exists at all;
//   * the five input streams of the line arrive through TMA: per group of 8 columns ONE elected lane issues five
//     cp.async.bulk.tensor.3d loads (box 8 x 4 x 8 doubles, mbarrier complete_tx) into a ring of NG groups; the results leave the
//     same way (cp.async.bulk.tensor store from a small output ring).  No per-lane address arithmetic, no cp.async bursts, and the
//     box lands [k][j][i]-ordered, which is bank-conflict-free for the skewed lanes;
//   * the faces a neighbouring line needs (last-j pencils -> line cj+1, last-k pencils -> line ck+1) go through MAILBOXES in global
//     memory whose empty state is a sentinel bit pattern (a signalling NaN no arithmetic produces): the producer lane stores each
//     value the moment it is computed (8-byte store, no fence), the consumer polls THE DATA ITSELF with L2 loads issued a few steps
//     ahead, and writes the sentinel back once it has the value, so the mailbox is clean again for the next sweep.  The lag between
//     adjacent lines is the lane skew + one group + one L2 round trip instead of a whole tile + fence + flag + fetch;
//   * lines are drawn from a ticket counter in (cj + ck) order, so a line only waits on lines with smaller tickets, which are
//     running or done: no co-residency assumption, no deadlock (same argument as the tile kernel).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int SG_COLS = 8;                                  // columns per group (64-byte row pieces)
constexpr int SG_TK = 8;                                    // k planes of a line (lane & 7)
constexpr unsigned long long SG_SENT = 0xFFF4A5C3D2E1B497ull;   // "not written yet": a signalling-NaN pattern arithmetic never returns

struct StreamArgs {
	int nx, ny, nz;
	int vpitch;
	long long vplane, voff;
	int nty, ntz;                  // lines per plane direction
	int nxp;                       // mailbox row length (columns rounded up to groups)
	const int *order;              // lines sorted by cj + ck
	unsigned *ticket;
	unsigned long long *mail_j;    // [line][8 k-rows][nxp]   last-j pencils of the line
	unsigned long long *mail_k;    // [line][TJ j-rows][nxp]  last-k pencils of the line
	double *out;                   // the output vector: its ghost column / rows / planes are read; the last column of an odd nx is
	                               // stored directly (the TMA store's inner bound is 16-byte granular: measured on B200, a box over
	                               // an odd-length row also writes the element after it -- the ghost column)
	int nx_tma;                    // nx & ~1: columns the out tensor map covers
	int klo_zero, khi_zero;        // slab-local sweeps: what lies below k = 0 / above k = nz-1 is taken as zero
	long long *prof;
};

__device__ __forceinline__ unsigned s_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sg_mbar_init(unsigned a, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory"); }
__device__ __forceinline__ void sg_mbar_expect_tx(unsigned a, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(a), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool sg_mbar_try_wait(unsigned a, unsigned parity)
{
	unsigned ok;
	asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void tma_load3(unsigned dst, const CUtensorMap *tm, int c0, int c1, int c2, unsigned mbar)
{
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
	             ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(mbar) : "memory");
}
__device__ __forceinline__ void tma_store3(const CUtensorMap *tm, int c0, int c1, int c2, unsigned src)
{
	asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];\n"
	             ::"l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(src) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

__device__ __forceinline__ ulonglong2 ld_mail2(const unsigned long long *p)
{
	ulonglong2 v;
	asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];\n" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_mail2(unsigned long long *p, unsigned long long a, unsigned long long b)
{
	asm volatile("st.global.cg.v2.u64 [%0], {%1, %2};\n" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void st_mail_if(unsigned long long *p, double v, bool on)
{
	asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p st.global.cg.f64 [%0], %1;\n}\n" ::"l"(p), "d"(v), "r"((unsigned)on) : "memory");
}

// MODE 0: forward sweep  y = (r - A0 y[i-1] - A1 y[j-1] - A2 y[k-1]) dinv
// MODE 1: backward sweep z = y - dinv (A4 z[i+1] + A5 z[j+1] + A6 z[k+1])
// MODE 2: StrILU diagonal d = ((A3 - P0/d[i-1]) - P1/d[j-1]) - P2/d[k-1]   (forward order; in0 = A3 in the coefficient layout)
template <int MODE, int NG>
__global__ void __launch_bounds__(32)
sgs_stream_kernel(const __grid_constant__ CUtensorMap tm_in0, const __grid_constant__ CUtensorMap tm_c0,
                  const __grid_constant__ CUtensorMap tm_c1, const __grid_constant__ CUtensorMap tm_c2,
                  const __grid_constant__ CUtensorMap tm_dv, const __grid_constant__ CUtensorMap tm_out,
                  const StreamArgs a, const int *stop)
{
	constexpr bool FWD = (MODE != 1);
	constexpr bool ILU = (MODE == 2);
	constexpr int NA = ILU ? 4 : 5;
	constexpr int DIR = FWD ? 1 : -1;
	constexpr int TJ = 4, TK = SG_TK, ROWS = TJ * TK;
	constexpr int GD = ROWS * SG_COLS;                       // doubles of one stream of one group (2 KB)
	constexpr int GS = NA * GD;                              // doubles of one group, all streams
	constexpr int NO = 4;                                    // output ring, groups
	constexpr int HW = NG * SG_COLS;                         // halo ring width, columns
	constexpr int S = TJ + TK - 2;                           // steps the last lane lags the first
	constexpr int QS = (S + 8) & 7, DS = (S + 8) >> 3;       // the last lane leaves group n at step 8 (n + DS) + QS
	static_assert((NG & (NG - 1)) == 0 && NG >= 4, "ring sizes are powers of two");
	constexpr unsigned FULL = 0xffffffffu;
	if (s3d_stopped(stop)) return;

	extern __shared__ __align__(1024) double sg_smem[];      // the only shared memory of the kernel: starts at the CTA's window, 1 KB aligned
	double *s_in = sg_smem;                                  // [NG][NA][GD]
	double *s_out = s_in + NG * GS;                          // [NO][GD]
	double *s_hj = s_out + NO * GD;                          // [8][HW]   j-neighbour rows (physical k) of the line before in j
	double *s_hk = s_hj + 8 * HW;                            // [TJ][HW]  k-neighbour rows (physical j) of the line before in k
	double *s_trash = s_hk + TJ * HW;                        // [32] where out-of-range steps put their (discarded) result
	unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(s_trash + 32);   // [NG]

	const int lane = threadIdx.x;
	const int m = lane >> 3, cc = lane & 7;
	const int s0 = m + cc;
	const int nlines = a.nty * a.ntz;
	const int nwarps = gridDim.x;
	const int gmax = (a.nx - 1) >> 3, itop = 8 * gmax + 7;
	const int ncol = FWD ? a.nx : itop + 1;
	const int ngroups = gmax + 1;
	const int nsteps = ncol + S;

	if (lane == 0) {
#pragma unroll
		for (int s = 0; s < NG; s++) sg_mbar_init(s_u32(s_bar + s), 1);
		asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
	}
	fence_async_smem();
	__syncwarp();
	unsigned phases = 0;                                     // bit s: parity the next wait on ring slot s uses

	int ticket = blockIdx.x;
	int line = (ticket < nlines) ? a.order[FWD ? ticket : nlines - 1 - ticket] : -1;
	while (line >= 0) {
		const int cj = line % a.nty, ck = line / a.nty;
		const int j0 = cj * TJ, k0 = ck * TK;
		const int tj = min(TJ, a.ny - j0), tk = min(TK, a.nz - k0);
		const bool act = (m < tj) && (cc < tk);
		const int pj = act ? (FWD ? m : tj - 1 - m) : 0, pk = act ? (FWD ? cc : tk - 1 - cc) : 0;   // physical pencil inside the line
		const int rowoff = (pk * TJ + pj) * SG_COLS;
		// the lines this one depends on / feeds (sweep order)
		const int pjl = FWD ? (cj > 0 ? line - 1 : -1) : (cj < a.nty - 1 ? line + 1 : -1);
		const int pkl = FWD ? (ck > 0 ? line - a.nty : -1) : (ck < a.ntz - 1 ? line + a.nty : -1);
		const bool jsucc = FWD ? (cj < a.nty - 1) : (cj > 0), ksucc = FWD ? (ck < a.ntz - 1) : (ck > 0);
		const bool jprod = act && jsucc && (m == tj - 1), kprod = act && ksucc && (cc == tk - 1);
		double *odd_dst = a.out + a.voff + (long long)(k0 + pk) * a.vplane + (long long)(j0 + pj) * a.vpitch;   // used for column nx-1 of an odd nx only
		unsigned long long *mj_dst = a.mail_j + ((size_t)line * 8 + pk) * a.nxp;
		unsigned long long *mk_dst = a.mail_k + ((size_t)line * TJ + pj) * a.nxp;
		// face items of this lane: (row, 16-byte chunk) of an 8-column group
		const int frow = lane >> 2, fch = 2 * (lane & 3);
		const bool jitem = frow < tk, kitem = (lane < 4 * TJ) && (frow < tj);
		const bool jpoll = pjl >= 0, kpoll = pkl >= 0;
		const bool kzero = !kpoll && (FWD ? a.klo_zero : a.khi_zero);
		const unsigned long long *jsrc = jpoll
			? a.mail_j + ((size_t)pjl * 8 + frow) * a.nxp + fch
			: reinterpret_cast<const unsigned long long *>(a.out + a.voff + (long long)(k0 + frow) * a.vplane + (long long)(FWD ? -1 : a.ny) * a.vpitch + fch);
		const unsigned long long *ksrc = kpoll
			? a.mail_k + ((size_t)pkl * TJ + frow) * a.nxp + fch
			: reinterpret_cast<const unsigned long long *>(a.out + a.voff + (long long)(FWD ? -1 : a.nz) * a.vplane + (long long)(j0 + frow) * a.vpitch + fch);

		auto stage = [&](int n) {                               // logical group n -> its ring slot (elected lane only)
			const int pg = FWD ? n : gmax - n, slot = pg & (NG - 1);
			const unsigned bar = s_u32(s_bar + slot), dst = s_u32(s_in + slot * GS);
			sg_mbar_expect_tx(bar, (unsigned)(GS * sizeof(double)));
			tma_load3(dst, &tm_in0, 8 * pg, j0, k0, bar);
			tma_load3(dst + GD * 8, &tm_c0, 8 * pg, j0, k0, bar);
			tma_load3(dst + 2 * GD * 8, &tm_c1, 8 * pg, j0, k0, bar);
			tma_load3(dst + 3 * GD * 8, &tm_c2, 8 * pg, j0, k0, bar);
			if (!ILU) tma_load3(dst + 4 * GD * 8, &tm_dv, 8 * pg, j0, k0, bar);
		};
		ulonglong2 fj = make_ulonglong2(0, 0), fk = make_ulonglong2(0, 0);
		auto face_issue = [&](int n) {
			const int i8 = 8 * (FWD ? n : gmax - n);
			if (jitem) fj = ld_mail2(jsrc + i8);
			if (kitem && !kzero) fk = ld_mail2(ksrc + i8);
		};
		auto face_take = [&](int n) {                           // poll until the group's faces are there, move them to shared memory
			const int i8 = 8 * (FWD ? n : gmax - n);
			const int col = i8 + fch;
			const bool wj = jitem && jpoll, wk = kitem && kpoll;
			const long long t0 = clock64();
			for (;;) {
				const bool okj = !wj || ((col >= a.nx || fj.x != SG_SENT) && (col + 1 >= a.nx || fj.y != SG_SENT));
				const bool okk = !wk || ((col >= a.nx || fk.x != SG_SENT) && (col + 1 >= a.nx || fk.y != SG_SENT));
				if (__all_sync(FULL, okj && okk)) break;
				if (!okj) fj = ld_mail2(jsrc + i8);
				if (!okk) fk = ld_mail2(ksrc + i8);
				if (clock64() - t0 > (40ll << 30)) __trap();       // ~20 s without the producing line: fail loudly instead of hanging
			}
			const int hs = i8 & (HW - 1);
			if (jitem) {
				*reinterpret_cast<ulonglong2 *>(s_hj + frow * HW + hs + fch) = fj;
				if (jpoll) st_mail2(const_cast<unsigned long long *>(jsrc) + i8, SG_SENT, SG_SENT);   // clean for the next sweep
			}
			if (kitem) {
				*reinterpret_cast<ulonglong2 *>(s_hk + frow * HW + hs + fch) = fk;
				if (kpoll) st_mail2(const_cast<unsigned long long *>(ksrc) + i8, SG_SENT, SG_SENT);
			}
		};

		// ---- prologue: the first NG groups, the faces of group 0, the ghost column
		__syncwarp();
		if (lane == 0)
			for (int n = 0; n < NG && n < ngroups; n++) stage(n);
		face_issue(0);
		double last0 = 0.0;
		if (act) last0 = __ldcg(a.out + a.voff + (long long)(k0 + pk) * a.vplane + (long long)(j0 + pj) * a.vpitch + (FWD ? -1 : a.nx));
		int nticket = 0, nline = -1;
		if (lane == 0) nticket = (int)atomicAdd(a.ticket, 1u) + nwarps;
		face_take(0);
		{
			const int slot = (FWD ? 0 : gmax) & (NG - 1);
			while (!__all_sync(FULL, sg_mbar_try_wait(s_u32(s_bar + slot), (phases >> slot) & 1u))) { }
			phases ^= 1u << slot;
		}
		__syncwarp();

		// ---- the march.  At step h this lane is at logical column h - s0, physical column icol.
		const bool jhalo = (m == 0), khalo = (cc == 0);
		const double *phj = s_hj + pk * HW, *phk = s_hk + pj * HW;
		int icol = FWD ? -s0 : itop + s0;
		auto inoff = [&](int ic) { return ((ic >> 3) & (NG - 1)) * GS + (ic & 7) + rowoff; };
		auto outoff = [&](int ic) { return ((ic >> 3) & (NO - 1)) * GD + (ic & 7) + rowoff; };
		const int trash = (int)(s_trash - s_out) + lane;
		int io = inoff(icol);
		double a_in = s_in[io], a_c0 = s_in[io + GD], a_c1 = s_in[io + 2 * GD], a_c2 = s_in[io + 3 * GD];
		double a_dv = ILU ? 0.0 : s_in[io + (NA - 1) * GD];
		double hj0 = jhalo ? phj[icol & (HW - 1)] : 0.0, hk0 = khalo ? phk[icol & (HW - 1)] : 0.0;
		int next_store = 0;
		long long tq_face = 0, tq_bar = 0, tq_io = 0, tq0 = 0, tq = 0;
		if (a.prof) tq0 = clock64();
#pragma unroll 1
		for (int h8 = 0; h8 < nsteps; h8 += 8) {
			const int n1 = (h8 >> 3) + 1;                        // the group the first lane enters after this octet
#pragma unroll
			for (int q = 0; q < 8; q++) {
				if (q == 0 && n1 < ngroups) face_issue(n1);
				if (q == QS) {
					// every lane has left logical group nt: its results go out, its ring slot takes group nt + NG
					const int nt = (h8 >> 3) - DS;
					if (nt >= 0 && nt < ngroups) {
						if (a.prof) tq = clock64();
						fence_async_smem();
						__syncwarp();
						if (lane == 0) {
							const int pg = FWD ? nt : gmax - nt;
							if (8 * pg < a.nx_tma) tma_store3(&tm_out, 8 * pg, j0, k0, s_u32(s_out + (pg & (NO - 1)) * GD));
							bulk_commit();
							if (nt + NG < ngroups) stage(nt + NG);
						}
						__syncwarp();
						next_store = nt + 1;
						if (a.prof) tq_io += clock64() - tq;
					}
				}
				if (q == 7 && n1 < ngroups) {
					// the first lane is about to read group n1: its inputs, its faces, and a free slot of the output ring
					if (a.prof) tq = clock64();
					if (lane == 0) bulk_wait_read<1>();
					face_take(n1);
					if (a.prof) { const long long t = clock64(); tq_face += t - tq; tq = t; }
					const int slot = (FWD ? n1 : gmax - n1) & (NG - 1);
					while (!__all_sync(FULL, sg_mbar_try_wait(s_u32(s_bar + slot), (phases >> slot) & 1u))) { }
					phases ^= 1u << slot;
					__syncwarp();
					if (a.prof) tq_bar += clock64() - tq;
				}
				const double shj = __shfl_up_sync(FULL, last0, 8);
				const double shk = __shfl_up_sync(FULL, last0, 1);
				const int ic = icol;
				icol += DIR;
				io = inoff(icol);
				const double n_a_in = s_in[io], n_a_c0 = s_in[io + GD], n_a_c1 = s_in[io + 2 * GD], n_a_c2 = s_in[io + 3 * GD];
				const double n_a_dv = ILU ? 0.0 : s_in[io + (NA - 1) * GD];
				const double n_hj0 = jhalo ? phj[icol & (HW - 1)] : 0.0, n_hk0 = khalo ? phk[icol & (HW - 1)] : 0.0;
				const double wj = jhalo ? hj0 : shj, wk = khalo ? hk0 : shk;
				double v0;
				if (ILU) v0 = __dsub_rn(__dsub_rn(__dsub_rn(a_in, __ddiv_rn(a_c0, last0)), __ddiv_rn(a_c1, wj)), __ddiv_rn(a_c2, wk));
				else if (FWD) v0 = __dmul_rn(__dsub_rn(__dsub_rn(__dsub_rn(a_in, __dmul_rn(a_c0, last0)), __dmul_rn(a_c1, wj)), __dmul_rn(a_c2, wk)), a_dv);
				else v0 = __dsub_rn(a_in, __dmul_rn(a_dv, __dadd_rn(__dadd_rn(__dmul_rn(a_c0, last0), __dmul_rn(a_c1, wj)), __dmul_rn(a_c2, wk))));
				// branch-free commit: out-of-range steps keep the old value and write to the trash slot
				const bool inr = act && (unsigned)ic < (unsigned)a.nx;
				last0 = inr ? v0 : last0;
				s_out[inr ? outoff(ic) : trash] = v0;
				st_mail_if(reinterpret_cast<unsigned long long *>(odd_dst + ic), v0, inr && ic >= a.nx_tma);
				st_mail_if(mj_dst + ic, v0, inr && jprod);
				st_mail_if(mk_dst + ic, v0, inr && kprod);
				a_in = n_a_in; a_c0 = n_a_c0; a_c1 = n_a_c1; a_c2 = n_a_c2; a_dv = n_a_dv; hj0 = n_hj0; hk0 = n_hk0;
			}
			if (h8 == 16 && lane == 0) nline = (nticket < nlines) ? a.order[FWD ? nticket : nlines - 1 - nticket] : -1;
		}
		// ---- epilogue: the groups still in the output ring
		fence_async_smem();
		__syncwarp();
		if (lane == 0) {
			for (int nt = next_store; nt < ngroups; nt++) {
				const int pg = FWD ? nt : gmax - nt;
				if (8 * pg < a.nx_tma) tma_store3(&tm_out, 8 * pg, j0, k0, s_u32(s_out + (pg & (NO - 1)) * GD));
			}
			bulk_commit();
			bulk_wait_read<0>();
			if (nsteps <= 16) nline = (nticket < nlines) ? a.order[FWD ? nticket : nlines - 1 - nticket] : -1;   // short lines never reach the look-up above
			if (a.prof) {
				atomicAdd((unsigned long long *)a.prof + 0, (unsigned long long)tq_face);
				atomicAdd((unsigned long long *)a.prof + 1, (unsigned long long)tq_bar);
				atomicAdd((unsigned long long *)a.prof + 2, (unsigned long long)(clock64() - tq0));
				atomicAdd((unsigned long long *)a.prof + 3, (unsigned long long)tq_io);
				atomicAdd((unsigned long long *)a.prof + 4, 1ull);
				atomicAdd((unsigned long long *)a.prof + 5, (unsigned long long)nsteps);
			}
		}
		__syncwarp();
		line = __shfl_sync(FULL, nline, 0);
	}
	if (lane == 0) bulk_wait_all();                          // the stores have landed before the warp retires
}

template <int NG> constexpr size_t sg_smem(int na) { return (size_t)(NG * na * 256 + 4 * 256 + 12 * NG * 8 + 32) * sizeof(double) + NG * 8; }

__global__ void __launch_bounds__(256) fill_u64_kernel(unsigned long long *p, size_t n, unsigned long long v)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

PFN_cuTensorMapEncodeTiled g_encode = nullptr;

int make_map(s3d_ctx *ctx, CUtensorMap *tm, const double *base, int nx, int ny, int nz, long long pitch, long long plane)
{
	const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
	const cuuint64_t strides[2] = {(cuuint64_t)pitch * 8, (cuuint64_t)plane * 8};
	const cuuint32_t box[3] = {SG_COLS, 4, SG_TK};
	const cuuint32_t es[3] = {1, 1, 1};
	const CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims, strides, box, es,
	                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
	                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) return s3d_fail(ctx, S3D_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %d x %d x %d grid", (int)r, nx, ny, nz);
	return S3D_OK;
}

}  // namespace

// One sweep (mode 0 forward, 1 backward, 2 StrILU recurrence) of the streaming kernel.  in0: vector layout (coefficient layout for
// mode 2); c0..c2, dinv: coefficient layout; out: vector layout.  order/nty/ntz: the line dispatch table of sgs_flow_prepare
// (variant "lines", 4 rows per line).  The caller has rewound the ticket counter (sgs_begin_kernel).
extern "C" S3D_HIDDEN int s3d_internal_sgs_stream(s3d_ctx *ctx, const s3d_grid *g, int mode, const double *in0, const double *c0,
                                                  const double *c1, const double *c2, const double *dinv, double *out,
                                                  const int *order, int nty, int ntz, unsigned *ticket, int klo_zero, int khi_zero,
                                                  const int *stop)
{
	constexpr int NG = 4;
	if (!g_encode) {
		void *fn = nullptr;
		cudaDriverEntryPointQueryResult qres;
		S3D_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
		if (!fn || qres != cudaDriverEntryPointSuccess) return s3d_fail(ctx, S3D_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
		g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
	}
	static bool configured = false;
	static int warps_per_sm = 1;
	if (!configured) {
		S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_stream_kernel<0, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg_smem<NG>(5)));
		S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_stream_kernel<1, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg_smem<NG>(5)));
		S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_stream_kernel<2, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg_smem<NG>(4)));
		int a = 1, b = 1;
		S3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, sgs_stream_kernel<0, NG>, 32, sg_smem<NG>(5)));
		S3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, sgs_stream_kernel<1, NG>, 32, sg_smem<NG>(5)));
		warps_per_sm = std::max(1, std::min(a, b));
		configured = true;
	}
	// mailboxes: 8 + 4 rows of nxp columns per line, all sentinel outside a sweep (the consumer of a value puts the sentinel back)
	const int nxp = ((g->nx + 7) / 8) * 8;
	const size_t nlines = (size_t)nty * ntz;
	const size_t need = nlines * 12 * (size_t)nxp;
	if (ctx->sgs_mail_cap < need) {
		S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		if (ctx->sgs_mail) S3D_CUDA(ctx, cudaFree(ctx->sgs_mail));
		ctx->sgs_mail = nullptr; ctx->sgs_mail_cap = 0;
		S3D_CUDA(ctx, cudaMalloc(&ctx->sgs_mail, need * sizeof(unsigned long long)));
		fill_u64_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->sgs_mail, need, SG_SENT);
		S3D_LAUNCHED(ctx);
		ctx->sgs_mail_cap = need;
	}
	StreamArgs a;
	a.nx = g->nx; a.ny = g->ny; a.nz = g->nz; a.vpitch = g->vpitch; a.vplane = g->vplane; a.voff = g->voff;
	a.nty = nty; a.ntz = ntz; a.nxp = nxp; a.order = order; a.ticket = ticket;
	a.mail_j = ctx->sgs_mail; a.mail_k = ctx->sgs_mail + nlines * 8 * (size_t)nxp;
	a.out = out; a.nx_tma = g->nx & ~1; a.klo_zero = klo_zero; a.khi_zero = khi_zero; a.prof = ctx->sgs_prof;
	CUtensorMap tm[6];
	int rc;
	if (mode == 2) rc = make_map(ctx, &tm[0], in0, g->nx, g->ny, g->nz, g->cpitch, g->cplane);
	else rc = make_map(ctx, &tm[0], in0 + g->voff, g->nx, g->ny, g->nz, g->vpitch, g->vplane);
	if (rc != S3D_OK) return rc;
	const double *cs[4] = {c0, c1, c2, mode == 2 ? c0 : dinv};
	for (int s = 0; s < 4; s++) {
		rc = make_map(ctx, &tm[1 + s], cs[s], g->nx, g->ny, g->nz, g->cpitch, g->cplane);
		if (rc != S3D_OK) return rc;
	}
	rc = make_map(ctx, &tm[5], out + g->voff, std::max(2, g->nx & ~1), g->ny, g->nz, g->vpitch, g->vplane);   // nx == 1: never used
	if (rc != S3D_OK) return rc;
	int wps = warps_per_sm;
	if (ctx->sgs_warps_per_sm > 0) wps = std::min(wps, ctx->sgs_warps_per_sm);
	const unsigned nblocks = (unsigned)std::min<size_t>(nlines, (size_t)ctx->sm_count * wps);
	if (mode == 0) sgs_stream_kernel<0, NG><<<nblocks, 32, sg_smem<NG>(5), ctx->stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], a, stop);
	else if (mode == 1) sgs_stream_kernel<1, NG><<<nblocks, 32, sg_smem<NG>(5), ctx->stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], a, stop);
	else sgs_stream_kernel<2, NG><<<nblocks, 32, sg_smem<NG>(4), ctx->stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], a, stop);
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}
