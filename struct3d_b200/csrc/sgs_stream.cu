// Exact lexicographic Gauss-Seidel sweeps as a STREAMING dataflow for sm_100a
// (reference: SGS_like_preconditioner::apply, src/linalg/s3d_sgspreconditioners.cpp:53-110, and the StrILU diagonal recurrence,
// src/linalg/s3d_ilu.cpp:34-57).  Same expressions, same roundings as the other sweep kernels in sgs.cu: bit-identical results.
//
// Why another kernel (measured in round 1, profiles/r1_sgs_phases.md): the tile-level dataflow hands work from tile to tile through
// global memory -- store rows, fence, flag, poll, fetch halo -- and a tile only starts when its three predecessors are COMPLETE, so the
// dependency chain of a sweep is (#tile hops) x (a whole tile + ~3500 cycles of hand-off): 78 x 7600 cycles at 256^3 against a
// fundamental chain of nx+ny+nz steps.  Here the hand-off is per column pair and needs neither fence nor flag:
//
//   * a block of three warps owns a LINE at a time: the 4 x 8 (j,k) pencils of a (cj, ck) cell over ALL nx columns.
//     The COMPUTE warp: lane (m, c) marches along i one column per step, TWO steps behind lanes (m-1, c) and (m, c-1): its i-neighbour
//     is its own last value, its j/k-neighbours are the values those lanes computed two steps ago (warp shuffle of the value before
//     last, so the shuffle latency is off the recurrence).  No i-direction hand-off exists at all.
//     The LOADER warp stages the inputs and stores the results, the FACE warp polls the neighbouring lines' mailboxes.  (Doing this
//     in the computing warp cost 120 of 265 cycles per step, and one helper warp for both could not keep up: a lone warp issues ~1
//     dependent instruction per 5 cycles; profiles/r2_sgs_stream.md.)
//   * the skew lives in the shared-memory layout: the loader puts element (i, pj, pk) of a stream at SHEARED column u = i + 2 pj + 2 pk
//     of row (pj, pk), so column h of every row holds exactly the element that row's lane needs at step h.  All lanes read the same
//     column at the same step: the inner loop addresses shared memory with immediates only.  Data groups of 16 columns (128-byte row
//     pieces: whole lines for the memory system -- with 64-byte pieces the SM's outstanding requests capped a sweep at 3.8 TB/s)
//     travel through a ring of NG slots as 16-byte cp.async copies, completion counted on an mbarrier; the 16-byte chunks of a row
//     are XOR-swizzled with the row number so that the 16-byte loads of the 32 rows are bank-conflict-free.  (The first working
//     version made the TMA unit do the shear -- tensor maps with row stride = pitch - 2 elements, SWIZZLE_64B -- which was correct but
//     slow: every TMA op on a box of 64-byte rows holds the SM's TMA unit for ~130 cycles.)
//   * results replace the staged input in shared memory (one 16-byte st.shared per lane every other step, same immediate address)
//     and the loader warp stores them with coalesced 16-byte stores, predicated at the ends of the rows;
//   * the faces a neighbouring line needs (last-j pencils -> line cj+1, last-k pencils -> line ck+1) go through MAILBOXES in global
//     memory whose empty state is a sentinel bit pattern (a signalling NaN no arithmetic produces): the producing lane stores each
//     pair the moment it is complete (no fence), the consuming line's FACE warp polls THE DATA ITSELF with L2 loads, hands the values
//     to its compute warp through shared memory in half-groups of 8 columns, and writes the sentinel back, so the mailbox is clean
//     again for the next sweep.  The lag between adjacent lines is the lane skew + 8 columns + one L2 round trip instead of a whole
//     tile + fence + flag + fetch;
//   * lines are drawn from a ticket counter in (cj + ck) order, so a line only waits on lines with smaller tickets, which are
//     running or done: no co-residency assumption, no deadlock (same argument as the tile kernel).
#include <algorithm>
#include "common.cuh"

namespace {

constexpr int SG_COLS = 8;                                  // columns of a half-group: the granularity of the face hand-off
constexpr int SG_DG = 16;                                   // columns of a data group: 128-byte row pieces
constexpr int SG_TJ = 4, SG_TK = 8;                         // pencils of a line: lane = m * 8 + c, m < 4 (j), c < 8 (k)
constexpr int SG_LQ = 8;                                    // entries of the line queue between the warps
constexpr int SG_HW = 64, SG_HP = SG_HW + 2;                // face ring: 8 half-groups wide; row pitch shifts consecutive rows by 16 bytes in the banks
constexpr unsigned long long SG_SENT = 0xFFF4A5C3D2E1B497ull;   // "not written yet": a signalling-NaN pattern arithmetic never returns

struct StreamArgs {
	int nx, ny, nz;
	int vpitch;
	long long vplane, voff;
	int nty, ntz;                  // lines per plane direction
	int nxp;                       // mailbox row length (columns rounded up to even)
	const int *order;              // lines sorted by cj + ck
	unsigned *ticket;
	unsigned long long *mail_j;    // [line][8 k-rows][nxp]   last-j pencils of the line
	unsigned long long *mail_k;    // [line][4 j-rows][nxp]   last-k pencils of the line
	const double *in0, *c0, *c1, *c2, *dv;   // input streams: in0 in the vector layout (coefficient layout for the StrILU recurrence)
	int cpitch;
	long long cplane;
	double *out;                   // the output vector (its ghost column / rows / planes are read, never written)
	int klo_zero, khi_zero;        // slab-local sweeps: what lies below k = 0 / above k = nz-1 is taken as zero
	long long *prof;
	int debug;                     // timing experiments only (results become wrong): 1 = no hand-off between lines (no mailbox traffic at all)
};

__device__ __forceinline__ unsigned s_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sg_mbar_init(unsigned a, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(a), "r"(count) : "memory"); }
__device__ __forceinline__ void sg_mbar_arrive(unsigned a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(a) : "memory"); }
__device__ __forceinline__ bool sg_mbar_try_wait(unsigned a, unsigned parity)
{
	unsigned ok;
	asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(a), "r"(parity) : "memory");
	return ok != 0;
}
__device__ __forceinline__ void cp_async16_if(unsigned dst, const double *src, bool on)
{
	asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p cp.async.cg.shared.global [%0], [%1], 16;\n}\n" ::"r"(dst), "l"(src), "r"((unsigned)on) : "memory");
}
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
// predicated stores of a result pair / of its even-column half (no "memory" clobber: they must not fence the shared-memory loads
// of the following steps; nothing in this kernel reads these locations back through ordinary loads)
__device__ __forceinline__ void st_pair_if(double *p, double lo, double hi, bool both, bool lo_only)
{
	asm volatile("{\n.reg .pred p, q;\nsetp.ne.u32 p, %3, 0;\nsetp.ne.u32 q, %4, 0;\n@p st.global.cg.v2.f64 [%0], {%1, %2};\n@q st.global.cg.f64 [%0], %1;\n}\n"
	             ::"l"(p), "d"(lo), "d"(hi), "r"((unsigned)both), "r"((unsigned)lo_only));
}
__device__ __forceinline__ void st_pair_on(double *p, double lo, double hi, bool on)
{
	asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %3, 0;\n@p st.global.cg.v2.f64 [%0], {%1, %2};\n}\n" ::"l"(p), "d"(lo), "d"(hi), "r"((unsigned)on));
}
__device__ __forceinline__ int ld_volatile_s32(const int *p)
{
	int v;
	asm volatile("ld.volatile.shared.s32 %0, [%1];\n" : "=r"(v) : "r"(s_u32(p)) : "memory");
	return v;
}

// What the warps derive from a line number.
//   sheared column  u = i + 2 pj + 2 pk,  0 <= u < W = nx + 2 (tj-1) + 2 (tk-1): nh half-groups of 8 columns, nd data groups of 16
//   forward sweeps walk the columns upwards, backward sweeps downwards; the lane of logical pencil (m, c) holds the physical pencil
//   (m, c) forward and (tj-1-m, tk-1-c) backward, so the lanes it depends on, (m-1, c) and (m, c-1), are two steps ahead of it.
//   Logical data group d holds the logical half-groups 2 d and 2 d + 1; its 16-column window starts at sheared column ud0(d).
template <bool FWD> struct LineGeo {
	int line, cj, ck, j0, k0, tj, tk, nh, nd, shmax;
	bool act;
	int pj, pk, sh2, row;
	__device__ __forceinline__ void set(const StreamArgs &a, int ln, int lane)
	{
		line = ln;
		cj = ln % a.nty; ck = ln / a.nty;
		j0 = cj * SG_TJ; k0 = ck * SG_TK;
		tj = min(SG_TJ, a.ny - j0); tk = min(SG_TK, a.nz - k0);
		shmax = 2 * (tj - 1) + 2 * (tk - 1);
		nh = (a.nx + shmax + SG_COLS - 1) / SG_COLS;
		nd = (nh + 1) / 2;
		const int m = lane >> 3, c = lane & 7;
		act = (m < tj) && (c < tk);
		pj = act ? (FWD ? m : tj - 1 - m) : 0;
		pk = act ? (FWD ? c : tk - 1 - c) : 0;
		sh2 = 2 * pj + 2 * pk;
		row = act ? pj * SG_TK + pk : lane;   // idle lanes keep a row of their own: it lies outside the grid, nobody reads or stores it
	}
	__device__ __forceinline__ int u0h(int h) const { return SG_COLS * (FWD ? h : nh - 1 - h); }      // first column of logical half-group h
	__device__ __forceinline__ int ud0(int d) const { return FWD ? SG_DG * d : SG_COLS * (nh - 2 - 2 * d); }   // first column of data group d's window
	__device__ __forceinline__ double *orow(const StreamArgs &a) const { return a.out + a.voff + (long long)(k0 + pk) * a.vplane + (long long)(j0 + pj) * a.vpitch; }
};

// MODE 0: forward sweep  y = (r - A0 y[i-1] - A1 y[j-1] - A2 y[k-1]) dinv
// MODE 1: backward sweep z = y - dinv (A4 z[i+1] + A5 z[j+1] + A6 z[k+1])
// MODE 2: StrILU diagonal d = ((A3 - P0/d[i-1]) - P1/d[j-1]) - P2/d[k-1]   (forward order; in0 = A3 in the coefficient layout)
template <int MODE, int NG>
__global__ void __launch_bounds__(96)
sgs_stream_kernel(const StreamArgs a, const int *stop)
{
	constexpr bool FWD = (MODE != 1);
	constexpr bool ILU = (MODE == 2);
	constexpr int NA = ILU ? 4 : 5;
	constexpr int TJ = SG_TJ, TK = SG_TK, ROWS = TJ * TK;
	constexpr int GD = ROWS * SG_DG;                         // doubles of one stream of one data group (4 KB)
	constexpr int GS = NA * GD;                              // doubles of one data group, all streams
	constexpr int HP = SG_HP;
	constexpr unsigned FULL = 0xffffffffu;
	if (s3d_stopped(stop)) return;

	extern __shared__ __align__(1024) double sg_smem[];      // the only shared memory of the kernel: starts at the CTA's window, 1 KB aligned
	double *s_in = sg_smem;                                  // [NG][NA][j 4][k 8][16]: 16-byte chunk q of row r sits at chunk q ^ (r & 7)
	double *s_hj = s_in + NG * GS;                           // [8 k-rows][HP]  j-neighbour values of half-group H at columns (H % 8) * 8 ...
	double *s_hk = s_hj + 8 * HP;                            // [4 j-rows][HP]  k-neighbour values
	double *s_ghost = s_hk + TJ * HP;                        // [SG_LQ][32]     the ghost column of a line, per lane
	unsigned long long *s_full = reinterpret_cast<unsigned long long *>(s_ghost + SG_LQ * 32);   // [NG]   the inputs of a data group have landed
	unsigned long long *s_done = s_full + NG;                // [NG]   the compute warp is done with a data group
	unsigned long long *s_face = s_done + NG;                // [8]    the faces of a half-group are in the face ring
	unsigned long long *s_lbar = s_face + 8;                 // [SG_LQ] a line-queue entry is valid
	int *s_lineq = reinterpret_cast<int *>(s_lbar + SG_LQ);  // [SG_LQ] line numbers, -1 = no more lines
	int *s_prog = s_lineq + SG_LQ;                           // half-groups the compute warp has finished (all lines)

	const int lane = threadIdx.x & 31;
	const int warp = threadIdx.x >> 5;
	const int m = lane >> 3, cc = lane & 7;
	const int nlines = a.nty * a.ntz;

	if (threadIdx.x == 0) {
#pragma unroll
		for (int s = 0; s < NG; s++) { sg_mbar_init(s_u32(s_full + s), 32); sg_mbar_init(s_u32(s_done + s), 1); }
#pragma unroll
		for (int s = 0; s < 8; s++) sg_mbar_init(s_u32(s_face + s), 1);
#pragma unroll
		for (int s = 0; s < SG_LQ; s++) sg_mbar_init(s_u32(s_lbar + s), 1);
		*s_prog = 0;
		asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
	}
	__syncthreads();

	if (warp == 0) {
		// ======================================================================= the COMPUTE warp
		LineGeo<FWD> L;
		int G = 0, H = 0;                                        // data groups / half-groups processed so far (all lines)
		long long tq_wait = 0, tq_fwait = 0, tq0 = 0, nst = 0;
		if (a.prof) tq0 = clock64();
		for (int seq = 0;; seq++) {
			{
				const unsigned bar = s_u32(s_lbar + (seq % SG_LQ));
				while (!__all_sync(FULL, sg_mbar_try_wait(bar, (seq / SG_LQ) & 1))) { }
			}
			const int line = s_lineq[seq % SG_LQ];
			if (line < 0) break;
			L.set(a, line, lane);
			const bool jsucc = FWD ? (L.cj < a.nty - 1) : (L.cj > 0), ksucc = FWD ? (L.ck < a.ntz - 1) : (L.ck > 0);
			const bool jprod = L.act && jsucc && (m == L.tj - 1) && !(a.debug & 1), kprod = L.act && ksucc && (cc == L.tk - 1) && !(a.debug & 1);
			double *mj_dst = reinterpret_cast<double *>(a.mail_j + ((size_t)line * 8 + L.pk) * a.nxp);
			double *mk_dst = reinterpret_cast<double *>(a.mail_k + ((size_t)line * TJ + L.pj) * a.nxp);
			int chunk_off[8];
#pragma unroll
			for (int q = 0; q < 8; q++) chunk_off[q] = L.row * SG_DG + 2 * (q ^ (L.row & 7));
			const bool jhalo = (m == 0), khalo = (cc == 0);
			const double *phj = s_hj + L.pk * HP, *phk = s_hk + L.pj * HP;
			double cur = 0.0, prv = 0.0;                             // the lane's last value (its i-neighbour) and the one before
			long long t_first = 0;
#pragma unroll 1
			for (int d = 0; d < L.nd; d++, G++) {
				const int slot = G % NG;
				{
					long long t = 0;
					if (a.prof) t = clock64();
					const unsigned bar = s_u32(s_full + slot);
					while (!__all_sync(FULL, sg_mbar_try_wait(bar, (G / NG) & 1))) { }
					if (a.prof) tq_wait += clock64() - t;
				}
				double *g_in = s_in + slot * GS;                        // results overwrite the first stream of the slot
				const int ud0 = L.ud0(d);
#pragma unroll
				for (int hh = 0; hh < 2; hh++) {
					if (2 * d + hh >= L.nh) break;
					{
						long long t = 0;
						if (a.prof) t = clock64();
						const unsigned bar = s_u32(s_face + (H & 7));
						while (!__all_sync(FULL, sg_mbar_try_wait(bar, (H >> 3) & 1))) { }
						if (a.prof) tq_fwait += clock64() - t;
					}
					if (d == 0 && hh == 0) {
						cur = s_ghost[(seq % SG_LQ) * 32 + lane];
						if (a.prof && FWD) asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t_first));
					}
					const int pb = FWD ? SG_COLS * hh : SG_COLS * (1 - hh);   // position of the half-group inside the data group's window
					const int u0 = ud0 + pb;
					const int hb = (H & 7) * SG_COLS;
					const int icb = u0 - L.sh2;                         // real column of the half-group's first column
					double2 o_in, o_c0, o_c1, o_c2, o_dv, o_hj, o_hk;
					o_dv = make_double2(0.0, 0.0); o_hj = make_double2(0.0, 0.0); o_hk = make_double2(0.0, 0.0);
					const bool interior = (u0 >= L.shmax) && (u0 + SG_COLS <= a.nx);
					// one column: the reference's expression, the reference's order of roundings
					auto column = [&](int el) -> double {
						const double shj = __shfl_up_sync(FULL, prv, 8);
						const double shk = __shfl_up_sync(FULL, prv, 1);
						const double x_in = el ? o_in.y : o_in.x, x_c0 = el ? o_c0.y : o_c0.x, x_c1 = el ? o_c1.y : o_c1.x, x_c2 = el ? o_c2.y : o_c2.x;
						const double x_dv = el ? o_dv.y : o_dv.x;
						const double wj = jhalo ? (el ? o_hj.y : o_hj.x) : shj, wk = khalo ? (el ? o_hk.y : o_hk.x) : shk;
						if (ILU) return __dsub_rn(__dsub_rn(__dsub_rn(x_in, __ddiv_rn(x_c0, cur)), __ddiv_rn(x_c1, wj)), __ddiv_rn(x_c2, wk));
						if (FWD) return __dmul_rn(__dsub_rn(__dsub_rn(__dsub_rn(x_in, __dmul_rn(x_c0, cur)), __dmul_rn(x_c1, wj)), __dmul_rn(x_c2, wk)), x_dv);
						return __dsub_rn(x_in, __dmul_rn(x_dv, __dadd_rn(__dadd_rn(__dmul_rn(x_c0, cur), __dmul_rn(x_c1, wj)), __dmul_rn(x_c2, wk))));
					};
					auto fetch = [&](int cq, int q) {                   // chunk cq of the window = pair q of the half-group
						o_in = *reinterpret_cast<const double2 *>(g_in + chunk_off[cq]);
						o_c0 = *reinterpret_cast<const double2 *>(g_in + GD + chunk_off[cq]);
						o_c1 = *reinterpret_cast<const double2 *>(g_in + 2 * GD + chunk_off[cq]);
						o_c2 = *reinterpret_cast<const double2 *>(g_in + 3 * GD + chunk_off[cq]);
						if (!ILU) o_dv = *reinterpret_cast<const double2 *>(g_in + 4 * GD + chunk_off[cq]);
						if (jhalo) o_hj = *reinterpret_cast<const double2 *>(phj + hb + 2 * q);
						if (khalo) o_hk = *reinterpret_cast<const double2 *>(phk + hb + 2 * q);
					};
					if (interior) {
						// every lane's eight columns lie inside the grid: no range checks, nothing to select
#pragma unroll
						for (int p = 0; p < 4; p++) {
							const int q = FWD ? p : 3 - p;              // pair of the half-group
							const int cq = (FWD ? 4 * hh : 4 * (1 - hh)) + q;
							fetch(cq, q);
							double v[2];
#pragma unroll
							for (int e = 0; e < 2; e++) {
								const int el = FWD ? e : 1 - e;         // forward sweeps take the even column first
								const double r = column(el);
								prv = cur; cur = r; v[el] = r;
							}
							*reinterpret_cast<double2 *>(g_in + chunk_off[cq]) = make_double2(v[0], v[1]);
							const int iclo = icb + 2 * q;
							st_pair_on(mj_dst + iclo, v[0], v[1], jprod);
							st_pair_on(mk_dst + iclo, v[0], v[1], kprod);
						}
					} else {
#pragma unroll
						for (int p = 0; p < 4; p++) {
							const int q = FWD ? p : 3 - p;
							const int cq = (FWD ? 4 * hh : 4 * (1 - hh)) + q;
							fetch(cq, q);
							double v[2];
							bool in[2];
#pragma unroll
							for (int e = 0; e < 2; e++) {
								const int el = FWD ? e : 1 - e;
								const int ic = icb + 2 * q + el;
								const double r = column(el);
								const bool inr = L.act && (unsigned)ic < (unsigned)a.nx;
								prv = cur; cur = inr ? r : cur; v[el] = r; in[el] = inr;
							}
							*reinterpret_cast<double2 *>(g_in + chunk_off[cq]) = make_double2(v[0], v[1]);
							const int iclo = icb + 2 * q;
							st_pair_if(mj_dst + iclo, v[0], v[1], jprod && in[0] && in[1], jprod && in[0] && !in[1]);
							st_pair_if(mk_dst + iclo, v[0], v[1], kprod && in[0] && in[1], kprod && in[0] && !in[1]);
						}
					}
					H++;
					__syncwarp();
					if (lane == 0) *reinterpret_cast<volatile int *>(s_prog) = H;   // the face ring entry of this half-group is free again
				}
				// the group's results are in shared memory: over to the loader warp, which stores them
				__syncwarp();
				if (lane == 0) sg_mbar_arrive(s_u32(s_done + slot));
			}
			if (a.prof && lane == 0) {
				nst += (long long)L.nh * SG_COLS;
				if (FWD && line < 2048) {                            // timing experiments: when each line's compute started and ended
					long long t_end;
					asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t_end));
					a.prof[8 + 2 * line] = t_first; a.prof[9 + 2 * line] = t_end;
				}
			}
		}
		if (a.prof && lane == 0) {
			atomicAdd((unsigned long long *)a.prof + 0, (unsigned long long)(tq_wait + tq_fwait));
			atomicAdd((unsigned long long *)a.prof + 7, (unsigned long long)tq_fwait);
			atomicAdd((unsigned long long *)a.prof + 2, (unsigned long long)(clock64() - tq0));
			atomicAdd((unsigned long long *)a.prof + 4, 1ull);
			atomicAdd((unsigned long long *)a.prof + 5, (unsigned long long)nst);
		}
		return;
	}

	if (warp == 1) {
		// ======================================================================= the LOADER warp
		// LOAD: the five input streams of a data group into a free ring slot -- 16-byte cp.async, each lane eight row pieces per
		// stream, written where the sheared, swizzled layout wants them; completion counted on the slot's `full` barrier.  RETIRE: the
		// results of the groups the compute warp has finished, shared memory -> global, coalesced 16-byte stores, predicated at the
		// ends of the rows.  It also draws the lines.
		LineGeo<FWD> Ll, Lr;
		int Gl = 0, Gr = 0, nl = 0, nr = 0, seql = 0, seqr = 0;
		bool have_l = false, have_r = false, end_l = false;
		int ticket = blockIdx.x;
		const int nblocks = gridDim.x;
		// this lane's pieces: 16-byte chunk lane & 7 of rows t * 4 + (lane >> 3), t = 0..7: pencil pj = t >> 1, pk = (t & 1) * 4 + (lane >> 3)
		const int pch = lane & 7, prl = lane >> 3;
		long long hp_load = 0, hp_ret = 0, hp_t = 0;
		for (;;) {
			if (a.prof) hp_t = clock64();
			for (;;) {
				if (!have_l && !end_l) {
					// next line of this block: the first comes with the block number, later ones from the ticket counter.  A line-queue
					// entry is only reused once RETIRE is past its old line: seql - seqr < SG_LQ holds because every line has at least
					// one group and LOAD is at most NG <= SG_LQ - 2 groups ahead of RETIRE
					int line = -1;
					if (lane == 0) {
						if (seql > 0) ticket = (int)atomicAdd(a.ticket, 1u) + nblocks;
						line = (ticket < nlines) ? a.order[FWD ? ticket : nlines - 1 - ticket] : -1;
						s_lineq[seql % SG_LQ] = line;
						sg_mbar_arrive(s_u32(s_lbar + (seql % SG_LQ)));
					}
					line = __shfl_sync(FULL, line, 0);
					if (line < 0) end_l = true;
					else { Ll.set(a, line, lane); have_l = true; nl = 0; }
				}
				if (!have_l || Gl >= Gr + NG) break;
				{
					const int slot = Gl % NG;
					const unsigned bar = s_u32(s_full + slot), dst = s_u32(s_in + slot * GS);
					const int ub = Ll.ud0(nl) + 2 * pch;                               // sheared column of this lane's chunk
#pragma unroll
					for (int t = 0; t < 8; t++) {
						const int pj = t >> 1, pk = (t & 1) * 4 + prl, row = pj * TK + pk;
						const int c = ub - 2 * pj - 2 * pk;                            // real column
						const bool on = (pj < Ll.tj) && (pk < Ll.tk) && c >= 0 && c < a.nx;
						const long long co = (long long)(Ll.k0 + pk) * a.cplane + (long long)(Ll.j0 + pj) * a.cpitch + c;
						const long long vo = a.voff + (long long)(Ll.k0 + pk) * a.vplane + (long long)(Ll.j0 + pj) * a.vpitch + c;
						const unsigned d = dst + (row * SG_DG + 2 * (pch ^ (row & 7))) * 8;
						cp_async16_if(d, ILU ? a.in0 + co : a.in0 + vo, on);
						cp_async16_if(d + GD * 8, a.c0 + co, on);
						cp_async16_if(d + 2 * GD * 8, a.c1 + co, on);
						cp_async16_if(d + 3 * GD * 8, a.c2 + co, on);
						if (!ILU) cp_async16_if(d + 4 * GD * 8, a.dv + co, on);
					}
					asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
				}
				Gl++; nl++;
				if (nl == Ll.nd) { have_l = false; seql++; }
			}
			if (a.prof) { const long long t = clock64(); hp_load += t - hp_t; hp_t = t; }
			if (Gr == Gl) break;                                     // nothing loaded that is not retired: the line sequence has ended
			if (!have_r) { Lr.set(a, s_lineq[seqr % SG_LQ], lane); have_r = true; nr = 0; }
			{
				const int slot = Gr % NG;
				const unsigned bar = s_u32(s_done + slot);
				while (!__all_sync(FULL, sg_mbar_try_wait(bar, (Gr / NG) & 1))) { }
				const double *g_out = s_in + slot * GS;
				const int ub = Lr.ud0(nr) + 2 * pch;
#pragma unroll
				for (int t = 0; t < 8; t++) {
					const int pj = t >> 1, pk = (t & 1) * 4 + prl, row = pj * TK + pk;
					const int c = ub - 2 * pj - 2 * pk;
					const double2 v = *reinterpret_cast<const double2 *>(g_out + row * SG_DG + 2 * (pch ^ (row & 7)));
					const bool inlo = (pj < Lr.tj) && (pk < Lr.tk) && c >= 0 && c < a.nx, inhi = inlo && (c + 1 < a.nx);
					st_pair_if(a.out + a.voff + (long long)(Lr.k0 + pk) * a.vplane + (long long)(Lr.j0 + pj) * a.vpitch + c, v.x, v.y, inhi, inlo && !inhi);
				}
				__syncwarp();
				Gr++; nr++;
				if (nr == Lr.nd) { have_r = false; seqr++; }
			}
			if (a.prof) hp_ret += clock64() - hp_t;
		}
		if (a.prof && lane == 0) {
			atomicAdd((unsigned long long *)a.prof + 1, (unsigned long long)hp_load);
			atomicAdd((unsigned long long *)a.prof + 6, (unsigned long long)hp_ret);
		}
		return;
	}

	// =========================================================================== the FACE warp
	// Polls the mailboxes of the two lines this one depends on and hands the values -- and the line's ghost column -- to the compute
	// warp through shared memory, one half-group of 8 columns at a time; puts the sentinel back.  The polls of three consecutive
	// half-groups are in flight (an L2 round trip each); every set of poll registers is written by its own load instruction, never
	// moved, so nothing waits on a poll it does not need yet.
	{
		LineGeo<FWD> Lf;
		int Hf = 0;
		const int frow = lane >> 2, fch = 2 * (lane & 3);
		long long hp_face = 0, hp_t = 0;
		for (int seq = 0;; seq++) {
			{
				const unsigned bar = s_u32(s_lbar + (seq % SG_LQ));
				while (!__all_sync(FULL, sg_mbar_try_wait(bar, (seq / SG_LQ) & 1))) { }
			}
			const int line = s_lineq[seq % SG_LQ];
			if (line < 0) break;
			if (a.prof) hp_t = clock64();
			Lf.set(a, line, lane);
			const int pjl = FWD ? (Lf.cj > 0 ? line - 1 : -1) : (Lf.cj < a.nty - 1 ? line + 1 : -1);
			const int pkl = FWD ? (Lf.ck > 0 ? line - a.nty : -1) : (Lf.ck < a.ntz - 1 ? line + a.nty : -1);
			// face items of this lane: (row, 16-byte chunk) of a half-group.  Row frow of the j-faces is physical k-row frow of the pencils
			// with logical m = 0 (physical pj = fpj); row frow of the k-faces is physical j-row frow of the pencils with logical c = 0
			const int fpj = FWD ? 0 : Lf.tj - 1, fpk = FWD ? 0 : Lf.tk - 1;
			const bool jitem = frow < Lf.tk, kitem = (lane < 4 * TJ) && (frow < Lf.tj);
			const bool jpoll = pjl >= 0 && !(a.debug & 1), kpoll = pkl >= 0 && !(a.debug & 1);
			const bool kzero = !kpoll && (FWD ? a.klo_zero : a.khi_zero);
			const unsigned long long *jsrc = jpoll
				? a.mail_j + ((size_t)pjl * 8 + frow) * a.nxp
				: reinterpret_cast<const unsigned long long *>(a.out + a.voff + (long long)(Lf.k0 + frow) * a.vplane + (long long)(FWD ? -1 : a.ny) * a.vpitch);
			const unsigned long long *ksrc = kpoll
				? a.mail_k + ((size_t)pkl * TJ + frow) * a.nxp
				: reinterpret_cast<const unsigned long long *>(a.out + a.voff + (long long)(FWD ? -1 : a.nz) * a.vplane + (long long)(Lf.j0 + frow) * a.vpitch);
			const int jsh = 2 * fpj + 2 * frow - fch, ksh = 2 * frow + 2 * fpk - fch;   // real column of the item = u0 - shift
			double ghostv = 0.0;
			if (Lf.act) ghostv = __ldcg(Lf.orow(a) + (FWD ? -1 : a.nx));
			ulonglong2 fj0 = make_ulonglong2(0, 0), fk0 = fj0, fj1 = fj0, fk1 = fj0, fj2 = fj0, fk2 = fj0;
			auto ask = [&](int n, ulonglong2 &fj, ulonglong2 &fk) {  // the polls of half-group n of the line
				if (n >= Lf.nh) return;
				if (a.debug & 2) { fj = make_ulonglong2(SG_SENT, SG_SENT); fk = fj; return; }   // experiment: no early polls
				const int u0 = Lf.u0h(n);
				const int cj_ = u0 - jsh, ck_ = u0 - ksh;
				if (jitem && cj_ >= 0 && cj_ < a.nx) fj = ld_mail2(jsrc + cj_);
				if (kitem && !kzero && ck_ >= 0 && ck_ < a.nx) fk = ld_mail2(ksrc + ck_);
			};
			auto take = [&](int n, ulonglong2 &fj, ulonglong2 &fk) { // wait for half-group n's faces, hand them over
				const int u0 = Lf.u0h(n);
				const int cj_ = u0 - jsh, ck_ = u0 - ksh;
				const bool inj = jitem && cj_ >= 0 && cj_ < a.nx, ink = kitem && !kzero && ck_ >= 0 && ck_ < a.nx;
				const bool wj = inj && jpoll, wk = ink && kpoll;
				// the compute warp must be through with the half-group that had this entry of the face ring before
				while (!__all_sync(FULL, ld_volatile_s32(s_prog) >= Hf - 7)) { }
				auto there = [&](const ulonglong2 &pj_, const ulonglong2 &pk_) {
					const bool okj = !wj || (pj_.x != SG_SENT && (cj_ + 1 >= a.nx || pj_.y != SG_SENT));
					const bool okk = !wk || (pk_.x != SG_SENT && (ck_ + 1 >= a.nx || pk_.y != SG_SENT));
					return __all_sync(FULL, okj && okk) != 0;
				};
				if (!there(fj, fk)) {
					// not there yet: keep TWO polls in flight, so that the values are seen within half an L2 round trip of landing
					const long long t0 = clock64();
					ulonglong2 gj = fj, gk = fk;
					if (wj) gj = ld_mail2(jsrc + cj_);
					if (wk) gk = ld_mail2(ksrc + ck_);
					for (;;) {
						if (wj) fj = ld_mail2(jsrc + cj_);
						if (wk) fk = ld_mail2(ksrc + ck_);
						if (there(gj, gk)) { fj = gj; fk = gk; break; }
						if (wj) gj = ld_mail2(jsrc + cj_);
						if (wk) gk = ld_mail2(ksrc + ck_);
						if (there(fj, fk)) break;
						if (clock64() - t0 > (40ll << 30)) __trap(); // ~20 s without the producing line: fail loudly instead of hanging
					}
				}
				const int hs = (Hf & 7) * SG_COLS + fch;
				if (jitem) {
					*reinterpret_cast<ulonglong2 *>(s_hj + frow * HP + hs) = inj ? fj : make_ulonglong2(0, 0);
					if (wj) st_mail2(const_cast<unsigned long long *>(jsrc) + cj_, SG_SENT, SG_SENT);   // clean for the next sweep
				}
				if (kitem) {
					*reinterpret_cast<ulonglong2 *>(s_hk + frow * HP + hs) = ink ? fk : make_ulonglong2(0, 0);
					if (wk) st_mail2(const_cast<unsigned long long *>(ksrc) + ck_, SG_SENT, SG_SENT);
				}
				if (n == 0) s_ghost[(seq % SG_LQ) * 32 + lane] = Lf.act ? ghostv : 0.0;
				__syncwarp();
				if (lane == 0) sg_mbar_arrive(s_u32(s_face + (Hf & 7)));
				Hf++;
			};
			if (a.debug & 4) {
				// experiment: polls two half-groups ahead
				ask(0, fj0, fk0);
				ask(1, fj1, fk1);
#pragma unroll 1
				for (int n = 0; n < Lf.nh; n += 3) {
					ask(n + 2, fj2, fk2);
					take(n, fj0, fk0);
					if (n + 1 >= Lf.nh) break;
					ask(n + 3, fj0, fk0);
					take(n + 1, fj1, fk1);
					if (n + 2 >= Lf.nh) break;
					ask(n + 4, fj1, fk1);
					take(n + 2, fj2, fk2);
				}
			} else {
				// the poll of the next half-group is in flight while this one is handed over: one half-group of lead hides the L2
				// round trip; more lead only makes the line trail its predecessors by more (measured: the lag settles at the lead)
				ask(0, fj0, fk0);
#pragma unroll 1
				for (int n = 0; n < Lf.nh; n += 2) {
					ask(n + 1, fj1, fk1);
					take(n, fj0, fk0);
					if (n + 1 >= Lf.nh) break;
					ask(n + 2, fj0, fk0);
					take(n + 1, fj1, fk1);
				}
			}
			if (a.prof) hp_face += clock64() - hp_t;
		}
		if (a.prof && lane == 0) atomicAdd((unsigned long long *)a.prof + 3, (unsigned long long)hp_face);
	}
}

template <int NG> constexpr size_t sg_smem(int na)
{
	return (size_t)(NG * na * 512 + 12 * SG_HP + SG_LQ * 32) * sizeof(double) + (2 * NG + 8 + SG_LQ) * 8 + (SG_LQ + 1) * 4;
}

__global__ void __launch_bounds__(256) fill_u64_kernel(unsigned long long *p, size_t n, unsigned long long v)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
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
	constexpr int NG = 3;
	static_assert(NG <= SG_LQ - 2, "the line queue must outlast the ring");
	static bool configured = false;
	static int blocks_per_sm = 1;
	if (!configured) {
		S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_stream_kernel<0, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg_smem<NG>(5)));
		S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_stream_kernel<1, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg_smem<NG>(5)));
		S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_stream_kernel<2, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg_smem<NG>(4)));
		int a = 1, b = 1;
		S3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, sgs_stream_kernel<0, NG>, 96, sg_smem<NG>(5)));
		S3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, sgs_stream_kernel<1, NG>, 96, sg_smem<NG>(5)));
		blocks_per_sm = std::max(1, std::min(a, b));
		configured = true;
	}
	// mailboxes: 8 + 4 rows of nxp columns per line, all sentinel outside a sweep (the consumer of a value puts the sentinel back)
	const int nxp = ((g->nx + 1) / 2) * 2;
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
	a.out = out; a.klo_zero = klo_zero; a.khi_zero = khi_zero; a.prof = ctx->sgs_prof; a.debug = ctx->sgs_sleep_ns;
	a.in0 = in0; a.c0 = c0; a.c1 = c1; a.c2 = c2; a.dv = (mode == 2) ? c0 : dinv;
	a.cpitch = g->cpitch; a.cplane = g->cplane;
	int bps = blocks_per_sm;
	if (ctx->sgs_warps_per_sm > 0) bps = std::min(bps, ctx->sgs_warps_per_sm);
	const unsigned nblocks = (unsigned)std::min<size_t>(nlines, (size_t)ctx->sm_count * bps);
	if (mode == 0) sgs_stream_kernel<0, NG><<<nblocks, 96, sg_smem<NG>(5), ctx->stream>>>(a, stop);
	else if (mode == 1) sgs_stream_kernel<1, NG><<<nblocks, 96, sg_smem<NG>(5), ctx->stream>>>(a, stop);
	else sgs_stream_kernel<2, NG><<<nblocks, 96, sg_smem<NG>(4), ctx->stream>>>(a, stop);
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}
