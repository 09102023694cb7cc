// Exact lexicographic Gauss-Seidel sweeps, REGISTER-STREAMING version for sm_100a
// (reference: SGS_like_preconditioner::apply, src/linalg/s3d_sgspreconditioners.cpp:53-110, and the StrILU diagonal recurrence,
// src/linalg/s3d_ilu.cpp:34-57).  Same expressions, same roundings as the other sweep kernels: bit-identical results.
//
// What the earlier kernels taught (profiles/r1_sgs_phases.md, profiles/r2_sgs_stream.md): shared-memory staging of the five input
// streams caps a sweep at 3-4 lines (100-128 pencils) per SM, and every hand-off through global memory costs 2.5 us, of which there
// are 64 + 32 on the critical path of a 256^3 sweep.  This kernel turns both knobs the other way:
//
//   * NO staging.  A lane owns one (j,k) pencil and walks along i in QUADS of four columns: one 256-bit load per input stream and
//     quad (LDG.E.256: a whole 32-byte sector per lane, rows are 128-byte aligned), D quads ahead in registers, one 256-bit store
//     per result quad.  Nothing but 6 KB of face rings per warp lives in shared memory, so registers alone bound the pencils per SM.
//   * a lane runs ONE WHOLE QUAD (one loop iteration) behind its j- and k-neighbours, lanes (m-1, c) and (m, c-1) of the 4 x 8
//     pencils of a warp: the four neighbour values of an iteration are last iteration's results, fetched with eight shuffles at the
//     top of the iteration, off the recurrence.  The chain of a sweep is ny + nz + nx/4 iterations of four dependent columns.
//   * a block of WJ x WK warps owns a LINE, the 4 WJ x 8 WK pencils of a (cj, ck) cell over all nx columns.  Between the warps of a
//     block the faces travel through small shared-memory rings, between blocks through mailboxes in global memory; in both the data
//     is its own flag (an empty slot holds a signalling-NaN pattern no arithmetic produces).  A mailbox slot is re-armed by its reader
//     (clean for the next sweep); a ring slot by its single WRITER, half a ring ahead, which it may do because it never runs half a
//     ring ahead of the reader's published count.  The global hops of a 256^3 sweep drop from 96 to 32 (16 x 16 pencils per block),
//     and a hop costs a store, an L2 round trip and a poll -- no fence, no flag.
//   MEASURED SLOWER than the tile kernel (2.8 vs 0.77 ms at 256^3, 8.8 vs 3.0 ms at 512^3: the quad skew doubles the chain of a sweep
//   and the uncoalesced 256-bit accesses serialise in the SM's load/store path; profiles/r2_sgs_pencil.md): opt-in, sgs_variant 6.
//   * every warp walks the block's line sequence on its own (a small line queue in shared memory, filled by the warp that is
//     first in sweep order): the warps of a block are only coupled through their rings, not through block barriers.
//   * lines are drawn from a ticket counter in (cj + ck) order, so a line only waits on lines with smaller tickets, which are
//     running or done: no co-residency assumption (same argument as the tile kernel).
#include <algorithm>
#include <type_traits>
#include <vector>
#include "common.cuh"

namespace {

constexpr unsigned long long SP_SENT = 0xFFF4A5C3D2E1B497ull;   // "not written yet" (the streaming kernel's sentinel: they share the mailbox buffer)
constexpr int SP_RING = 16;                                  // quads per shared-memory ring row
constexpr int SP_LQ = 8;                                     // entries of the line queue

struct PencilArgs {
	int nx, ny, nz, nq;            // nq = quads per row
	int vpitch, cpitch;
	long long vplane, voff, cplane;
	int ncj, nck;                  // lines per direction
	const int *order;              // lines sorted by cj + ck
	unsigned *ticket;
	unsigned long long *mail_j;    // [line][8 WK k-rows][4 nq]  last-j pencils of the line
	unsigned long long *mail_k;    // [line][4 WJ j-rows][4 nq]  last-k pencils of the line
	const double *in0, *c0, *c1, *c2, *dv;
	double *out;
	int klo_zero, khi_zero;        // slab-local sweeps: what lies below k = 0 / above k = nz-1 is taken as zero
	int debug;                     // timing experiments only (results become wrong): 1 = no hand-off between warps and blocks
};

struct Quad { double v[4]; };
struct Face { unsigned long long v[4]; };

__device__ __forceinline__ void ldg_quad_if(Quad &q, const double *p, bool on)
{
	asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %5, 0;\n@p ld.global.nc.L1::no_allocate.L2::128B.v4.f64 {%0, %1, %2, %3}, [%4];\n}\n"
	             : "+d"(q.v[0]), "+d"(q.v[1]), "+d"(q.v[2]), "+d"(q.v[3]) : "l"(p), "r"((unsigned)on));
}
__device__ __forceinline__ void stg_quad_if(double *p, const double (&r)[4], bool on)
{
	asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %5, 0;\n@p st.global.cg.v4.f64 [%0], {%1, %2, %3, %4};\n}\n"
	             ::"l"(p), "d"(r[0]), "d"(r[1]), "d"(r[2]), "d"(r[3]), "r"((unsigned)on) : "memory");
}
__device__ __forceinline__ void stg_one_if(double *p, double v, bool on)
{
	asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p st.global.cg.f64 [%0], %1;\n}\n" ::"l"(p), "d"(v), "r"((unsigned)on) : "memory");
}
// global mailboxes and fixed face data (ghost rows / planes): L2 accesses, never cached in L1, never hoisted
__device__ __forceinline__ void ldg_face_if(Face &f, const unsigned long long *p, bool on)
{
	asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %5, 0;\n@p ld.global.cg.v2.u64 {%0, %1}, [%4];\n@p ld.global.cg.v2.u64 {%2, %3}, [%4+16];\n}\n"
	             : "+l"(f.v[0]), "+l"(f.v[1]), "+l"(f.v[2]), "+l"(f.v[3]) : "l"(p), "r"((unsigned)on) : "memory");
}
__device__ __forceinline__ void stg_face_if(unsigned long long *p, unsigned long long a, unsigned long long b, unsigned long long c,
                                            unsigned long long d, bool on)
{
	asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %5, 0;\n@p st.global.cg.v2.u64 [%0], {%1, %2};\n@p st.global.cg.v2.u64 [%0+16], {%3, %4};\n}\n"
	             ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d), "r"((unsigned)on) : "memory");
}
// the same for reads that are issued AHEAD of their use: UNCONDITIONAL, pure outputs.  (A predicated vector load keeps the old
// register contents where the predicate is false, so ptxas loads into an aligned temporary quad and moves it out right behind
// the load -- which puts the whole load latency back onto the iteration: half of all stall samples, profiles/r2_sgs_pencil.md.)
// Lanes that have nothing to read are pointed at a harmless address by the caller.
__device__ __forceinline__ void ldg_face_ahead(Face &f, const unsigned long long *p)
{
	asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%4];\nld.global.cg.v2.u64 {%2, %3}, [%4+16];\n"
	             : "=l"(f.v[0]), "=l"(f.v[1]), "=l"(f.v[2]), "=l"(f.v[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void lds_face_ahead(Face &f, unsigned addr)
{
	asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%4];\nld.volatile.shared.v2.u64 {%2, %3}, [%4+16];\n"
	             : "=l"(f.v[0]), "=l"(f.v[1]), "=l"(f.v[2]), "=l"(f.v[3]) : "r"(addr) : "memory");
}
// shared-memory rings
__device__ __forceinline__ void lds_face_if(Face &f, unsigned addr, bool on)
{
	asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %5, 0;\n@p ld.volatile.shared.v2.u64 {%0, %1}, [%4];\n@p ld.volatile.shared.v2.u64 {%2, %3}, [%4+16];\n}\n"
	             : "+l"(f.v[0]), "+l"(f.v[1]), "+l"(f.v[2]), "+l"(f.v[3]) : "r"(addr), "r"((unsigned)on) : "memory");
}
__device__ __forceinline__ void sts_face_if(unsigned addr, unsigned long long a, unsigned long long b, unsigned long long c, unsigned long long d, bool on)
{
	asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %5, 0;\n@p st.volatile.shared.v2.u64 [%0], {%1, %2};\n@p st.volatile.shared.v2.u64 [%0+16], {%3, %4};\n}\n"
	             ::"r"(addr), "l"(a), "l"(b), "l"(c), "l"(d), "r"((unsigned)on) : "memory");
}
__device__ __forceinline__ void sts_u32_if(unsigned addr, unsigned v, bool on)
{
	asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p st.volatile.shared.u32 [%0], %1;\n}\n" ::"r"(addr), "r"(v), "r"((unsigned)on) : "memory");
}
__device__ __forceinline__ void lds_u32_if(unsigned &v, unsigned addr, bool on)
{
	asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p ld.volatile.shared.u32 %0, [%1];\n}\n" : "+r"(v) : "r"(addr), "r"((unsigned)on) : "memory");
}
__device__ __forceinline__ bool face_there(const Face &f)
{
	return f.v[0] != SP_SENT && f.v[1] != SP_SENT && f.v[2] != SP_SENT && f.v[3] != SP_SENT;
}
__device__ __forceinline__ int ld_vol_s32(const int *p) { return *reinterpret_cast<const volatile int *>(p); }
__device__ __forceinline__ unsigned sp_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// a wait that lasts seconds means a broken protocol: fail loudly instead of hanging the GPU (only looked at while spinning)
__device__ __forceinline__ void sp_spin_check(long long &t0)
{
	const long long t = clock64();
	if (t0 == 0) t0 = t;
	else if (t - t0 > (8ll << 30)) __trap();
}

// MODE 0: forward sweep  y = (r - A0 y[i-1] - A1 y[j-1] - A2 y[k-1]) dinv
// MODE 1: backward sweep z = y - dinv (A4 z[i+1] + A5 z[j+1] + A6 z[k+1])
// MODE 2: StrILU diagonal d = ((A3 - P0/d[i-1]) - P1/d[j-1]) - P2/d[k-1]   (forward order; in0 = A3 in the coefficient layout)
// Face source / destination kinds (warp-uniform): 0 none or fixed data (ghost row / plane, zero), 1 shared-memory ring, 2 global mailbox.
// Shared-memory rings: a slot holds the sentinel until the producing lane writes the quad; the producer also puts the sentinel
// into the slot RING/2 ahead (single writer per slot, nothing to order), which it may do because it never runs RING/2 quads ahead
// of the consumer's count (published per ring row).
template <int MODE, int WJ, int WK, int D>
__global__ void __launch_bounds__(32 * WJ * WK)
sgs_pencil_kernel(const PencilArgs a, const int *stop)
{
	constexpr bool FWD = (MODE != 1);
	constexpr bool ILU = (MODE == 2);
	constexpr int NW = WJ * WK, BJ = 4 * WJ, BK = 8 * WK;
	constexpr int RB = SP_RING * 32;                             // bytes of a ring row
	constexpr unsigned FULL = 0xffffffffu;
	if (s3d_stopped(stop)) return;

	extern __shared__ __align__(32) unsigned long long sp_smem[];
	unsigned long long *s_mj = sp_smem;                                  // [NW][8 c][RING][4]   j-faces INTO warp w
	unsigned long long *s_mk = s_mj + (size_t)NW * 8 * SP_RING * 4;      // [NW][4 m][RING][4]   k-faces INTO warp w
	unsigned *s_cnt = reinterpret_cast<unsigned *>(s_mk + (size_t)NW * 4 * SP_RING * 4);   // [NW][12] quads consumed per ring row
	int *s_lineq = reinterpret_cast<int *>(s_cnt + NW * 12);             // [LQ]
	int *s_pub = s_lineq + SP_LQ;                                        // queue entries published so far
	int *s_wseq = s_pub + 1;                                             // [NW] queue entries each warp has read

	for (int t = threadIdx.x; t < NW * 12 * SP_RING * 4; t += blockDim.x) sp_smem[t] = SP_SENT;
	for (int t = threadIdx.x; t < NW * 12; t += blockDim.x) s_cnt[t] = 0;
	if (threadIdx.x == 0) *s_pub = 0;
	if (threadIdx.x < NW) s_wseq[threadIdx.x] = 0;
	__syncthreads();

	const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
	const int wj = w % WJ, wk = w / WJ;
	const int m = lane >> 3, c = lane & 7;
	const int skew = m + c;
	const int nlines = a.ncj * a.nck;
	const bool drawer = (FWD ? w == 0 : w == NW - 1);         // the warp that is first in sweep order fills the line queue
	constexpr int step = FWD ? 4 : -4;                           // doubles per iteration along a row
	unsigned jin = 0, kin = 0, jout = 0, kout = 0;               // quads that went through this lane's ring rows before the current line
	unsigned jseen = 0, kseen = 0;                               // what this lane last read of its consumers' counts
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
					if (seq >= SP_LQ) {                                 // the entry about to be overwritten has been read by every warp
						for (long long t0 = 0;; sp_spin_check(t0)) {
							int lo = seq;
							for (int x = 0; x < NW; x++) if (x != w) lo = min(lo, ld_vol_s32(s_wseq + x));
							if (lo >= seq - SP_LQ + 1) break;
						}
					}
					*reinterpret_cast<volatile int *>(s_lineq + (seq % SP_LQ)) = line;
					__threadfence_block();
					*reinterpret_cast<volatile int *>(s_pub) = seq + 1;
				}
				if (line >= 0) tkt = (int)atomicAdd(a.ticket, 1u) + (int)gridDim.x;   // looked up when this line is done
			}
		} else {
			for (long long t0 = 0; ld_vol_s32(s_pub) <= seq; sp_spin_check(t0)) { }
			line = ld_vol_s32(s_lineq + (seq % SP_LQ));
			__syncwarp();
			if (lane == 0) *reinterpret_cast<volatile int *>(s_wseq + w) = seq + 1;
		}
		if (line < 0) break;

		const int cj = line % a.ncj, ck = line / a.ncj;
		const int j0 = cj * BJ + wj * 4, k0 = ck * BK + wk * 8;
		const int tj = min(4, a.ny - j0), tk = min(8, a.nz - k0);
		if (tj > 0 && tk > 0) {
			const bool act = (m < tj) && (c < tk);
			const int pj = act ? (FWD ? m : tj - 1 - m) : 0, pk = act ? (FWD ? c : tk - 1 - c) : 0;
			const long long vrow = a.voff + (long long)(k0 + pk) * a.vplane + (long long)(j0 + pj) * a.vpitch;
			const long long crow = (long long)(k0 + pk) * a.cplane + (long long)(j0 + pj) * a.cpitch;
			const int smax = tj + tk - 2;
			const int niter = a.nq + smax;
			const int q00 = FWD ? 0 : 4 * (a.nq - 1);            // column offset of the quad a lane starts with
			const bool nolink = (a.debug & 1) != 0;

			// ---- where the faces come from and go to (kinds are warp-uniform)
			const bool jface = act && (m == 0), kface = act && (c == 0);
			const bool jlast = act && (m == tj - 1), klast = act && (c == tk - 1);
			const bool jpred_in = FWD ? (j0 > 0) : (j0 + tj < a.ny), jpred_blk = FWD ? (wj > 0) : (wj < WJ - 1);
			const bool kpred_in = FWD ? (k0 > 0) : (k0 + tk < a.nz), kpred_blk = FWD ? (wk > 0) : (wk < WK - 1);
			const bool jsucc_in = FWD ? (j0 + tj < a.ny) : (j0 > 0), jsucc_blk = FWD ? (wj < WJ - 1) : (wj > 0);
			const bool ksucc_in = FWD ? (k0 + tk < a.nz) : (k0 > 0), ksucc_blk = FWD ? (wk < WK - 1) : (wk > 0);
			const int jsrc_kind = (!jpred_in || nolink) ? 0 : jpred_blk ? 1 : 2;
			const int ksrc_kind = (!kpred_in || nolink) ? 0 : kpred_blk ? 1 : 2;
			const int jdst_kind = (!jsucc_in || nolink) ? 0 : jsucc_blk ? 1 : 2;
			const int kdst_kind = (!ksucc_in || nolink) ? 0 : ksucc_blk ? 1 : 2;
			const bool kzero = !kpred_in && (FWD ? a.klo_zero : a.khi_zero);
			// running pointers: at iteration h a lane is at quad Q = h - skew, i.e. at columns q00 + step * Q ...
			const long long qstart = (long long)q00 - (long long)step * skew;
			const unsigned long long *jsrc = nullptr, *ksrc = nullptr;   // kinds 0 and 2
			unsigned long long *jdst = nullptr, *kdst = nullptr;         // kind 2
			if (jsrc_kind == 2) jsrc = a.mail_j + ((size_t)(FWD ? line - 1 : line + 1) * BK + (wk * 8 + pk)) * (size_t)(4 * a.nq) + qstart;
			else if (jsrc_kind == 0) jsrc = reinterpret_cast<const unsigned long long *>(a.out + a.voff + (long long)(k0 + pk) * a.vplane + (long long)(FWD ? -1 : a.ny) * a.vpitch) + qstart;
			if (ksrc_kind == 2) ksrc = a.mail_k + ((size_t)(FWD ? line - a.ncj : line + a.ncj) * BJ + (wj * 4 + pj)) * (size_t)(4 * a.nq) + qstart;
			else if (ksrc_kind == 0) ksrc = reinterpret_cast<const unsigned long long *>(a.out + a.voff + (long long)(FWD ? -1 : a.nz) * a.vplane + (long long)(j0 + pj) * a.vpitch) + qstart;
			if (jdst_kind == 2) jdst = a.mail_j + ((size_t)line * BK + (wk * 8 + pk)) * (size_t)(4 * a.nq) + qstart;
			if (kdst_kind == 2) kdst = a.mail_k + ((size_t)line * BJ + (wj * 4 + pj)) * (size_t)(4 * a.nq) + qstart;
			const int wsj = (jdst_kind == 1) ? (FWD ? w + 1 : w - 1) : w, wsk = (kdst_kind == 1) ? (FWD ? w + WJ : w - WJ) : w;
			const unsigned jring_in = sp_u32(s_mj) + (unsigned)((w * 8 + c) * RB), kring_in = sp_u32(s_mk) + (unsigned)((w * 4 + m) * RB);
			const unsigned jring_out = sp_u32(s_mj) + (unsigned)((wsj * 8 + c) * RB), kring_out = sp_u32(s_mk) + (unsigned)((wsk * 4 + m) * RB);
			const unsigned jcnt_in = sp_u32(s_cnt + w * 12 + c), kcnt_in = sp_u32(s_cnt + w * 12 + 8 + m);
			const unsigned jcnt_out = sp_u32(s_cnt + wsj * 12 + c), kcnt_out = sp_u32(s_cnt + wsk * 12 + 8 + m);
			// fixed data (ghost rows and planes) is read like a mailbox that is always full; nothing is read where the face is zero
			const bool jask = jface && !(nolink && jpred_in), kask = kface && !kzero && !(nolink && kpred_in);
			// lanes without a face to read stay on a harmless address (the reads ahead are unconditional)
			const unsigned long long *const dummy = a.mail_j;
			if (!jask || jsrc_kind == 1) jsrc = dummy;
			if (!kask || ksrc_kind == 1) ksrc = dummy;
			const int jstep = (jask && jsrc_kind != 1) ? step : 0, kstep = (kask && ksrc_kind != 1) ? step : 0;

			// ---- input streams: running pointers and the register pipeline
			const double *p_in = a.in0 + (ILU ? crow : vrow) + qstart;
			const double *p_c0 = a.c0 + crow + qstart, *p_c1 = a.c1 + crow + qstart, *p_c2 = a.c2 + crow + qstart;
			const double *p_dv = a.dv + crow + qstart;
			double *p_out = a.out + vrow + qstart;
			Quad s_in[D], s_c0[D], s_c1[D], s_c2[D], s_dv[D];
#pragma unroll
			for (int u = 0; u < D; u++) {
#pragma unroll
				for (int e = 0; e < 4; e++) { s_in[u].v[e] = 0.0; s_c0[u].v[e] = 0.0; s_c1[u].v[e] = 0.0; s_c2[u].v[e] = 0.0; s_dv[u].v[e] = 0.0; }
				const bool on = act && (unsigned)(u - skew) < (unsigned)a.nq;
				ldg_quad_if(s_in[u], p_in + step * u, on);
				ldg_quad_if(s_c0[u], p_c0 + step * u, on);
				ldg_quad_if(s_c1[u], p_c1 + step * u, on);
				ldg_quad_if(s_c2[u], p_c2 + step * u, on);
				if (!ILU) ldg_quad_if(s_dv[u], p_dv + step * u, on);
			}
			double cur = act ? __ldcg(a.out + vrow + (FWD ? -1 : a.nx)) : 0.0;   // the ghost column: i-neighbour of the first column
			double rp[4] = {0.0, 0.0, 0.0, 0.0};                                  // last iteration's results
			// the faces travel through the same D-deep register pipeline as the input streams (they stay zero where nothing is ever read)
			Face fj[D], fk[D];
#pragma unroll
			for (int u = 0; u < D; u++) {
#pragma unroll
				for (int e = 0; e < 4; e++) { fj[u].v[e] = 0; fk[u].v[e] = 0; }
				const bool on = (unsigned)(u - skew) < (unsigned)a.nq;
				if (jsrc_kind == 1) lds_face_if(fj[u], jring_in + ((jin + (unsigned)(u - skew)) % SP_RING) * 32, jask && on);
				else ldg_face_if(fj[u], jsrc + jstep * u, jask && on);
				if (ksrc_kind == 1) lds_face_if(fk[u], kring_in + ((kin + (unsigned)(u - skew)) % SP_RING) * 32, kask && on);
				else ldg_face_if(fk[u], ksrc + kstep * u, kask && on);
			}

			// one iteration; FAST: every lane of the warp is inside its row with a whole quad, and so are the quads asked for ahead
			auto iterate = [&](auto fast_tag, auto u_tag, int h) {
				constexpr bool FAST = decltype(fast_tag)::value;
				constexpr int u = decltype(u_tag)::value;
				const int Q = h - skew;
				const bool valid = FAST ? act : (act && (unsigned)Q < (unsigned)a.nq);
				const int col0 = q00 + step * Q;                       // first (lowest) column of the quad
				// ---- neighbours inside the warp: last iteration's results of lanes (m-1, c) and (m, c-1)
				double nj[4], nk[4];
#pragma unroll
				for (int e = 0; e < 4; e++) { nj[e] = __shfl_up_sync(FULL, rp[e], 8); nk[e] = __shfl_up_sync(FULL, rp[e], 1); }
				// ---- faces from outside the warp: asked for D iterations ago, waited for here
				const bool jw = jask && valid, kw = kask && valid;
				if (jsrc_kind == 1) {
					for (long long t0 = 0; !__all_sync(FULL, !jw || face_there(fj[u])); sp_spin_check(t0)) lds_face_if(fj[u], jring_in + ((jin + (unsigned)Q) % SP_RING) * 32, jw && !face_there(fj[u]));
				} else if (jsrc_kind == 2) {
					for (long long t0 = 0; !__all_sync(FULL, !jw || face_there(fj[u])); sp_spin_check(t0)) ldg_face_if(fj[u], jsrc, jw && !face_there(fj[u]));
					stg_face_if(const_cast<unsigned long long *>(jsrc), SP_SENT, SP_SENT, SP_SENT, SP_SENT, jw);   // clean for the next sweep
				}
				if (ksrc_kind == 1) {
					for (long long t0 = 0; !__all_sync(FULL, !kw || face_there(fk[u])); sp_spin_check(t0)) lds_face_if(fk[u], kring_in + ((kin + (unsigned)Q) % SP_RING) * 32, kw && !face_there(fk[u]));
				} else if (ksrc_kind == 2) {
					for (long long t0 = 0; !__all_sync(FULL, !kw || face_there(fk[u])); sp_spin_check(t0)) ldg_face_if(fk[u], ksrc, kw && !face_there(fk[u]));
					stg_face_if(const_cast<unsigned long long *>(ksrc), SP_SENT, SP_SENT, SP_SENT, SP_SENT, kw);
				}
				if (m == 0) {
#pragma unroll
					for (int e = 0; e < 4; e++) nj[e] = __longlong_as_double((long long)fj[u].v[e]);
				}
				if (c == 0) {
#pragma unroll
					for (int e = 0; e < 4; e++) nk[e] = __longlong_as_double((long long)fk[u].v[e]);
				}
				// the faces D iterations ahead: their reads fly during the arithmetic of D iterations
				{
					const bool nv = FAST ? act : (act && (unsigned)(Q + D) < (unsigned)a.nq);
					if (jsrc_kind == 1) {
						lds_face_ahead(fj[u], jring_in + ((jin + (unsigned)(Q + D)) % SP_RING) * 32);
						sts_u32_if(jcnt_in, jin + (unsigned)(Q + 1), jw);
					} else { ldg_face_ahead(fj[u], (FAST || nv) ? jsrc + jstep * D : dummy); jsrc += jstep; }
					if (ksrc_kind == 1) {
						lds_face_ahead(fk[u], kring_in + ((kin + (unsigned)(Q + D)) % SP_RING) * 32);
						sts_u32_if(kcnt_in, kin + (unsigned)(Q + 1), kw);
					} else { ldg_face_ahead(fk[u], (FAST || nv) ? ksrc + kstep * D : dummy); ksrc += kstep; }
				}
				// ---- four columns: the reference's expression, the reference's order of roundings
				double r[4];
#pragma unroll
				for (int ee = 0; ee < 4; ee++) {
					const int e = FWD ? ee : 3 - ee;
					const double x_in = s_in[u].v[e], x_c0 = s_c0[u].v[e], x_c1 = s_c1[u].v[e], x_c2 = s_c2[u].v[e], x_dv = s_dv[u].v[e];
					double v;
					if (ILU) v = __dsub_rn(__dsub_rn(__dsub_rn(x_in, __ddiv_rn(x_c0, cur)), __ddiv_rn(x_c1, nj[e])), __ddiv_rn(x_c2, nk[e]));
					else if (FWD) v = __dmul_rn(__dsub_rn(__dsub_rn(__dsub_rn(x_in, __dmul_rn(x_c0, cur)), __dmul_rn(x_c1, nj[e])), __dmul_rn(x_c2, nk[e])), x_dv);
					else v = __dsub_rn(x_in, __dmul_rn(x_dv, __dadd_rn(__dadd_rn(__dmul_rn(x_c0, cur), __dmul_rn(x_c1, nj[e])), __dmul_rn(x_c2, nk[e]))));
					r[e] = v;
					if (FAST) cur = v;
					else cur = (valid && col0 + e < a.nx) ? v : cur;
				}
#pragma unroll
				for (int e = 0; e < 4; e++) rp[e] = r[e];
				// ---- results: one 256-bit store per quad (element by element where the row ends inside the quad)
				if (FAST) {
					stg_quad_if(p_out, r, valid);
				} else {
					const bool whole = valid && (col0 + 3 < a.nx);
					stg_quad_if(p_out, r, whole);
					if (__any_sync(FULL, valid && !whole)) {
#pragma unroll
						for (int e = 0; e < 3; e++) stg_one_if(p_out + e, r[e], valid && !whole && (col0 + e < a.nx));
					}
				}
				// ---- faces for the warps and lines behind (in sweep order)
				const unsigned long long b0 = (unsigned long long)__double_as_longlong(r[0]), b1 = (unsigned long long)__double_as_longlong(r[1]);
				const unsigned long long b2 = (unsigned long long)__double_as_longlong(r[2]), b3 = (unsigned long long)__double_as_longlong(r[3]);
				if (jdst_kind == 1) {
					const bool on = jlast && valid;
					const unsigned n = jout + (unsigned)Q;                   // quads this lane has put into the ring before this one
					for (long long t0 = 0; !__all_sync(FULL, !on || (n - jseen) < (unsigned)(SP_RING / 2)); sp_spin_check(t0)) lds_u32_if(jseen, jcnt_out, on);
					sts_face_if(jring_out + (n % SP_RING) * 32, b0, b1, b2, b3, on);
					sts_face_if(jring_out + ((n + SP_RING / 2) % SP_RING) * 32, SP_SENT, SP_SENT, SP_SENT, SP_SENT, on);
				} else if (jdst_kind == 2) {
					stg_face_if(jdst, b0, b1, b2, b3, jlast && valid);
					jdst += step;
				}
				if (kdst_kind == 1) {
					const bool on = klast && valid;
					const unsigned n = kout + (unsigned)Q;
					for (long long t0 = 0; !__all_sync(FULL, !on || (n - kseen) < (unsigned)(SP_RING / 2)); sp_spin_check(t0)) lds_u32_if(kseen, kcnt_out, on);
					sts_face_if(kring_out + (n % SP_RING) * 32, b0, b1, b2, b3, on);
					sts_face_if(kring_out + ((n + SP_RING / 2) % SP_RING) * 32, SP_SENT, SP_SENT, SP_SENT, SP_SENT, on);
				} else if (kdst_kind == 2) {
					stg_face_if(kdst, b0, b1, b2, b3, klast && valid);
					kdst += step;
				}
				// ---- refill this stage with the quad D iterations ahead
				{
					const bool on = FAST ? act : (act && (unsigned)(Q + D) < (unsigned)a.nq);
					ldg_quad_if(s_in[u], p_in + step * D, on);
					ldg_quad_if(s_c0[u], p_c0 + step * D, on);
					ldg_quad_if(s_c1[u], p_c1 + step * D, on);
					ldg_quad_if(s_c2[u], p_c2 + step * D, on);
					if (!ILU) ldg_quad_if(s_dv[u], p_dv + step * D, on);
				}
				p_in += step; p_c0 += step; p_c1 += step; p_c2 += step; p_dv += step; p_out += step;
			};

#pragma unroll 1
			for (int h0 = 0; h0 < niter; h0 += D) {
				// all D iterations of the group are "fast" ones: h > smax (everybody has started, nobody is at a first quad) and
				// h + D + 1 < nq (nobody is at a last quad, every quad asked for ahead exists)
				const bool fast = (h0 > smax) && (h0 + 2 * D < a.nq);
				if (fast) {
					iterate(std::true_type{}, std::integral_constant<int, 0>{}, h0);
					if (D > 1) iterate(std::true_type{}, std::integral_constant<int, (D > 1 ? 1 : 0)>{}, h0 + 1);
					if (D > 2) iterate(std::true_type{}, std::integral_constant<int, (D > 2 ? 2 : 0)>{}, h0 + 2);
				} else {
					iterate(std::false_type{}, std::integral_constant<int, 0>{}, h0);
					if (D > 1 && h0 + 1 < niter) iterate(std::false_type{}, std::integral_constant<int, (D > 1 ? 1 : 0)>{}, h0 + 1);
					if (D > 2 && h0 + 2 < niter) iterate(std::false_type{}, std::integral_constant<int, (D > 2 ? 2 : 0)>{}, h0 + 2);
				}
			}
			// every ring row that was in use has carried nq quads
			if (jsrc_kind == 1) jin += (unsigned)a.nq;
			if (ksrc_kind == 1) kin += (unsigned)a.nq;
			if (jdst_kind == 1) jout += (unsigned)a.nq;
			if (kdst_kind == 1) kout += (unsigned)a.nq;
		}
		// ------------------------------------------------------------- the drawer looks its next line up
		if (drawer) {
			if (lane == 0) next_line = (tkt < nlines) ? a.order[FWD ? tkt : nlines - 1 - tkt] : -1;
			next_line = __shfl_sync(FULL, next_line, 0);
		}
	}
}

template <int WJ, int WK> constexpr size_t sp_smem_bytes()
{
	return (size_t)WJ * WK * 12 * SP_RING * 4 * sizeof(unsigned long long) + (size_t)(WJ * WK * 12 + SP_LQ + 1 + WJ * WK) * sizeof(int);
}

__global__ void __launch_bounds__(256) sp_fill_u64_kernel(unsigned long long *p, size_t n, unsigned long long v)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

template <int WJ, int WK, int D>
int sp_launch(s3d_ctx *ctx, const s3d_grid *g, int mode, PencilArgs &a, const int *stop)
{
	constexpr int NT = 32 * WJ * WK;
	constexpr size_t SM = sp_smem_bytes<WJ, WK>();
	static bool configured = false;
	static int blocks_per_sm = 1;
	if (!configured) {
		S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_pencil_kernel<0, WJ, WK, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM));
		S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_pencil_kernel<1, WJ, WK, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM));
		S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_pencil_kernel<2, WJ, WK, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM));
		int x = 1, y = 1;
		S3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&x, sgs_pencil_kernel<0, WJ, WK, D>, NT, SM));
		S3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&y, sgs_pencil_kernel<1, WJ, WK, D>, NT, SM));
		blocks_per_sm = std::max(1, std::min(x, y));
		configured = true;
	}
	// lines: (ny / 4 WJ) x (nz / 8 WK), dispatched in (cj + ck) order
	a.ncj = (g->ny + 4 * WJ - 1) / (4 * WJ); a.nck = (g->nz + 8 * WK - 1) / (8 * WK);
	const size_t nlines = (size_t)a.ncj * a.nck;
	const int okey = -(1000 + WJ * 16 + WK);   // no other kernel's table has this key
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
	// mailboxes: 4 WJ + 8 WK rows of 4 nq values per line, all sentinel outside a sweep (the consumer of a value puts the sentinel back)
	const size_t need = nlines * (size_t)(4 * WJ + 8 * WK) * (size_t)(4 * a.nq);
	if (ctx->sgs_mail_cap < need) {
		S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		if (ctx->sgs_mail) S3D_CUDA(ctx, cudaFree(ctx->sgs_mail));
		ctx->sgs_mail = nullptr; ctx->sgs_mail_cap = 0;
		S3D_CUDA(ctx, cudaMalloc(&ctx->sgs_mail, need * sizeof(unsigned long long)));
		sp_fill_u64_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->sgs_mail, need, SP_SENT);
		S3D_LAUNCHED(ctx);
		ctx->sgs_mail_cap = need;
	}
	a.mail_j = ctx->sgs_mail;
	a.mail_k = ctx->sgs_mail + nlines * (size_t)(8 * WK) * (size_t)(4 * a.nq);
	int bps = blocks_per_sm;
	if (ctx->sgs_warps_per_sm > 0) bps = std::max(1, std::min(bps, ctx->sgs_warps_per_sm / (WJ * WK)));
	const unsigned nblocks = (unsigned)std::min<size_t>(nlines, (size_t)ctx->sm_count * bps);
	if (mode == 0) sgs_pencil_kernel<0, WJ, WK, D><<<nblocks, NT, SM, ctx->stream>>>(a, stop);
	else if (mode == 1) sgs_pencil_kernel<1, WJ, WK, D><<<nblocks, NT, SM, ctx->stream>>>(a, stop);
	else sgs_pencil_kernel<2, WJ, WK, D><<<nblocks, NT, SM, ctx->stream>>>(a, stop);
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}

}  // namespace

// One sweep (mode 0 forward, 1 backward, 2 StrILU recurrence) of the register-streaming kernel.  in0: vector layout (coefficient
// layout for mode 2); c0..c2, dinv: coefficient layout; out: vector layout.  The caller has rewound the ticket counter
// (sgs_begin_kernel); shape selects the warps of a block (tuning key sgs_pencil_shape: 0 = by grid size).
extern "C" S3D_HIDDEN int s3d_internal_sgs_pencil(s3d_ctx *ctx, const s3d_grid *g, int mode, const double *in0, const double *c0,
                                                  const double *c1, const double *c2, const double *dinv, double *out,
                                                  unsigned *ticket, int klo_zero, int khi_zero, const int *stop)
{
	PencilArgs a;
	a.nx = g->nx; a.ny = g->ny; a.nz = g->nz; a.nq = (g->nx + 3) / 4;
	a.vpitch = g->vpitch; a.cpitch = g->cpitch; a.vplane = g->vplane; a.voff = g->voff; a.cplane = g->cplane;
	a.ticket = ticket;
	a.in0 = in0; a.c0 = c0; a.c1 = c1; a.c2 = c2; a.dv = (mode == 2) ? c0 : dinv;
	a.out = out; a.klo_zero = klo_zero; a.khi_zero = khi_zero; a.debug = ctx->sgs_sleep_ns;
	switch (ctx->sgs_pencil_shape) {
	case 11: return sp_launch<1, 1, 2>(ctx, g, mode, a, stop);
	case 21: return sp_launch<2, 1, 2>(ctx, g, mode, a, stop);
	case 22: return sp_launch<2, 2, 2>(ctx, g, mode, a, stop);
	case 41: return sp_launch<4, 1, 2>(ctx, g, mode, a, stop);
	case 42: return sp_launch<4, 2, 2>(ctx, g, mode, a, stop);
	case 44: return sp_launch<4, 4, 2>(ctx, g, mode, a, stop);
	case 111: return sp_launch<1, 1, 1>(ctx, g, mode, a, stop);
	case 113: return sp_launch<1, 1, 3>(ctx, g, mode, a, stop);
	case 223: return sp_launch<2, 2, 3>(ctx, g, mode, a, stop);
	case 423: return sp_launch<4, 2, 3>(ctx, g, mode, a, stop);
	case 221: return sp_launch<2, 2, 1>(ctx, g, mode, a, stop);
	case 421: return sp_launch<4, 2, 1>(ctx, g, mode, a, stop);
	default: return sp_launch<4, 2, 2>(ctx, g, mode, a, stop);
	}
}
