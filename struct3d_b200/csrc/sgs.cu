// Symmetric Gauss-Seidel preconditioner application  z = (D+U)^-1 D (D+L)^-1 r  for sm_100a
// (reference: SGS_like_preconditioner::apply, src/linalg/s3d_sgspreconditioners.cpp:14-115).
//
// The reference sweeps the grid lexicographically: forward  y = dinv (r - A0 y[i-1] - A1 y[j-1] - A2 y[k-1]),
// backward z = y - dinv (A4 z[i+1] + A5 z[j+1] + A6 z[k+1]).  Two GPU orderings are offered:
//
//  S3D_SGS_LEXICOGRAPHIC  the reference's ordering, executed as a wavefront: all points on a hyperplane
//      i+j+k = const are independent once the previous hyperplane is done, and each point evaluates the
//      very same expression, so the result equals the sequential sweep bit for bit (same roundings:
//      __dmul_rn/__dsub_rn, no FMA contraction).  The iteration count of any solver using it therefore
//      equals the reference's.  Implemented as a tile-level DATAFLOW kernel (below); the one-launch-per-
//      hyperplane version is kept as S3D_SGS_LEXICOGRAPHIC_PLANES for cross-checking.
//
//  S3D_SGS_REDBLACK       the same product M = (D+L) D^-1 (D+U) with the unknowns ordered red-then-black
//      ((i+j+k) parity).  Every half-sweep is fully parallel: three streaming launches, no dependencies.
//      It changes the iteration (different M), which is why it is an option and not the default.
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace {

// ---------------------------------------------------------------- lexicographic wavefront (exact)

// forward hyperplane h: points with i+j+k == h; threads enumerate (j,k), i = h-j-k
__global__ void __launch_bounds__(256)
sgs_fwd_plane_kernel(const GridDev g, const double *__restrict__ A, const double *__restrict__ dinv,
                     const double *__restrict__ r, double *y, int h, const int *stop)
{
	if (s3d_stopped(stop)) return;
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	const int k = blockIdx.y * blockDim.y + threadIdx.y;
	if (j >= g.ny || k >= g.nz) return;
	const int i = h - j - k;
	if (i < 0 || i >= g.nx) return;
	const long long v = g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + i;
	const long long c = (long long)k * g.cplane + (long long)j * g.cpitch + i;
	double t = __dsub_rn(r[v], __dmul_rn(A[c], y[v - 1]));
	t = __dsub_rn(t, __dmul_rn(A[c + g.cstride], y[v - g.vpitch]));
	t = __dsub_rn(t, __dmul_rn(A[c + 2 * g.cstride], y[v - g.vplane]));
	y[v] = __dmul_rn(t, dinv[c]);
}

__global__ void __launch_bounds__(256)
sgs_bwd_plane_kernel(const GridDev g, const double *__restrict__ A, const double *__restrict__ dinv,
                     const double *__restrict__ y, double *z, int h, const int *stop)
{
	if (s3d_stopped(stop)) return;
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	const int k = blockIdx.y * blockDim.y + threadIdx.y;
	if (j >= g.ny || k >= g.nz) return;
	const int i = h - j - k;
	if (i < 0 || i >= g.nx) return;
	const long long v = g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + i;
	const long long c = (long long)k * g.cplane + (long long)j * g.cpitch + i;
	double s = __dmul_rn(A[c + 4 * g.cstride], z[v + 1]);
	s = __dadd_rn(s, __dmul_rn(A[c + 5 * g.cstride], z[v + g.vpitch]));
	s = __dadd_rn(s, __dmul_rn(A[c + 6 * g.cstride], z[v + g.vplane]));
	z[v] = __dsub_rn(y[v], __dmul_rn(dinv[c], s));
}


// ---------------------------------------------------------------- lexicographic, tile-level dataflow (exact, fast)
//
// The grid is cut into tiles of 32 x 8 x 8 points.  A tile can be swept as soon as its three predecessor
// tiles (i-1, j-1, k-1 for the forward sweep; +1 for the backward sweep) are complete, so tiles on a tile-
// hyperplane run concurrently: at 256^3 that is 32768 tiles, ~120 in flight on average, ~460 at 512^3.
//   - persistent blocks draw tickets from a counter; ticket t maps to the t-th tile in HYPERPLANE order
//     (sorted by ci+cj+ck; reversed for the backward sweep), so the ~300 resident blocks always hold a
//     whole wavefront of mutually independent tiles (lexicographic dispatch was measured 35x slower: its
//     window spans one k-plane of tiles, of which only a diagonal is independent).  A tile only ever waits
//     on tiles with SMALLER tickets, whose owners are already running: no co-residency assumption, no deadlock;
//   - completion flags carry an epoch number (no clearing between sweeps); a block publishes its tile with
//     store -> __syncthreads -> __threadfence -> flag, a waiter spins on the flag and fences before loading
//     the halo values with ld.cg (L2): y lines never straddle tiles (rows are 128-byte aligned, tiles are 32
//     points wide), so a line read after the flag is complete and final;
//   - inside a tile the local hyperplanes h = li + lj + lk are swept by 64 threads, one per (j,k) pencil (see
//     the kernel); the five per-point inputs of the tile are staged through shared memory with cp.async
//     and the results leave through it, so every global access is a coalesced 16-byte transfer;
//   - every point evaluates the reference's expression with the reference's roundings, so the result is
//     bit-identical to the sequential sweep.
constexpr int SGS_TI = 32, SGS_TJ = 8, SGS_TK = 8;
constexpr int SGS_TILE = SGS_TI * SGS_TJ * SGS_TK;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
	const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
// cp.async groups are tracked per WARP (one scoreboard): commit and wait must be warp-uniform
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

struct SgsFlow {
	int *flags;        // one per tile: epoch of the last sweep that completed it
	double *iface;     // warp kernel: [tile][64] the tile's last i-column, indexed by logical pencil b*8+c
	const int *order;  // dispatch order: tile numbers sorted by tile hyperplane ci+cj+ck (then lexicographically)
	unsigned *ticket;  // work counter, zeroed before every launch
	const int *epoch_p; // device word holding the epoch of this sweep (bumped on the device, so a captured graph can be replayed)
	int ntx, nty, ntz;
	int sleep_ns;      // 0: busy-spin on the flag
	// z-slab runs (warp-per-tile kernel): the k-face rows of the rank before this one in sweep order arrive in x_in_plane (a plane of
	// this rank's peer-memory window, flag per (ci, cj) in x_in_flags); this rank's own go to x_out_plane / x_out_flags of the next rank
	const double *x_in_plane; const int *x_in_flags; double *x_out_plane; int *x_out_flags;
	int kzero_lo, kzero_hi;   // slab-local sweeps (warp kernel): what lies below k = 0 / above k = nz-1 is taken as zero, not read
	long long *prof;   // optional per-phase cycle accumulators (warp kernel; timing experiments only)
	int debug;         // timing experiments only (results become wrong): 1 skip wavefront, 2 skip staging, 4 skip fences, 8 skip waits
	// warp kernel, MAILBOX hand-off (one rank or slab-local): per tile the rows of its last j ([8 c][TI]), of its last k ([TJ b][TI]) and
	// its last i-column ([64]), all holding a sentinel outside a sweep; null: hand-off through `out`, fences and the epoch flags
	unsigned long long *mail;
	int mail_stride;   // doubles per tile: 8 TI + TJ TI + 64
	int probe;         // poll one value per predecessor in a tight loop before the full fetch (pays on latency-bound grids only)
	// warp kernel, z-BLOCKS on one device (S3D_SGS_BLOCKS, mailbox hand-off only): the planes of tile layer ck are not [8 ck, 8 ck + 8) but
	// layers[ck] = {first plane, planes, bit 0: a layer of the same block precedes it in sweep order | bit 1: one follows, plane whose
	// results are NOT stored (the block's overlap plane: it belongs to the neighbouring block) or a negative number}; null: the plain tiling
	const int4 *layers;
	int dbuf;          // warp kernel: two staging areas -- the copies of the next tile are issued BEFORE the sweep of this one (throughput-bound grids)
	const int4 *jlayers;   // the same along j (blocks of rows): rows of tile row cj = jlayers[cj] = {first row, rows, pred | succ << 1, row not stored}; null: plain
};
constexpr unsigned long long SGW_SENT = 0xFFF4A5C3D2E1B497ull;   // "not written yet": a signalling-NaN pattern no arithmetic produces

// FWD: in0 = r,  c0,c1,c2 = A0,A1,A2, out = y      y = (r - A0 y[i-1] - A1 y[j-1] - A2 y[k-1]) * dinv
// BWD: in0 = y,  c0,c1,c2 = A4,A5,A6, out = z      z = y - dinv * (A4 z[i+1] + A5 z[j+1] + A6 z[k+1])
//
// Thread mapping (measured: the first version gave every POINT COLUMN a thread, 256 threads per tile, and was
// issue-bound -- 8 warps executing the step body 46 times with <= 8 lanes active, 13.8 ms at 512^3).
// Now the 64 threads of a block are the 8 x 8 (j,k) PENCILS of the tile and each marches along i:
// at local step h the thread of pencil (lj,lk) computes point li = h - lj - lk.  Its i-neighbour is its own
// previous value (a register), its j- and k-neighbours are the previous step's values of the adjacent
// pencils (a double-buffered 10 x 10 array in shared memory); one 64-thread barrier per step, every lane
// busy for 32 of the 46 steps.
constexpr int SGS_NT = SGS_TJ * SGS_TK;   // 64 threads

// MODE 0: forward SGS sweep, 1: backward SGS sweep, 2: StrILU diagonal recurrence (forward order)
//   d = ((A3 - P0/d[i-1]) - P1/d[j-1]) - P2/d[k-1]   with P0 = A0*A4[i-1] etc. precomputed (zero where there is no neighbour),
//   reference: StrILU_preconditioner::updateOperatorWithBranchingInLoop, src/linalg/s3d_ilu.cpp:34-57.
template <int MODE>
__global__ void __launch_bounds__(SGS_NT)
sgs_dataflow_kernel(const GridDev g, const double *__restrict__ in0, const double *__restrict__ c0,
                    const double *__restrict__ c1, const double *__restrict__ c2, const double *__restrict__ dinv,
                    double *out, const SgsFlow f, const int *stop)
{
	constexpr bool FWD = (MODE != 1);
	constexpr bool ILU = (MODE == 2);
	if (s3d_stopped(stop)) return;
	const int epoch = __ldcg(f.epoch_p);
	extern __shared__ __align__(16) double smem[];
	double *s_in = smem;                       // [TK][TJ][TI] each; s_in is overwritten with the results
	double *s_c0 = s_in + SGS_TILE;
	double *s_c1 = s_c0 + SGS_TILE;
	double *s_c2 = s_c1 + SGS_TILE;
	double *s_dv = s_c2 + SGS_TILE;
	double *s_hj = s_dv + SGS_TILE;            // [TK][TI] the j-neighbour row of the adjacent tile (or ghost row)
	double *s_hk = s_hj + SGS_TK * SGS_TI;     // [TJ][TI] the k-neighbour row of the adjacent tile (or ghost plane)
	double *s_w = s_hk + SGS_TJ * SGS_TI;      // [2][TK+2][TJ+2] last value of every pencil, double-buffered
	__shared__ int s_tile;
	constexpr int WP = SGS_TJ + 2, WPL = (SGS_TK + 2) * WP;

	const int tid = threadIdx.x;
	const int lj = tid >> 3, lk = tid & 7;
	const int ntiles = f.ntx * f.nty * f.ntz;

	for (;;) {
		__syncthreads();                        // previous tile fully consumed before s_tile / smem are reused
		if (tid == 0) s_tile = (int)atomicAdd(f.ticket, 1u);
		__syncthreads();
		const int ticket = s_tile;
		if (ticket >= ntiles) return;
		const int tile = f.order[FWD ? ticket : ntiles - 1 - ticket];
		const int ci = tile % f.ntx, cj = (tile / f.ntx) % f.nty, ck = tile / (f.ntx * f.nty);
		const int i0 = ci * SGS_TI, j0 = cj * SGS_TJ, k0 = ck * SGS_TK;
		const int ti = min(SGS_TI, g.nx - i0), tj = min(SGS_TJ, g.ny - j0), tk = min(SGS_TK, g.nz - k0);
		const long long vtile = g.voff + (long long)k0 * g.vplane + (long long)j0 * g.vpitch + i0;
		const long long ctile = (long long)k0 * g.cplane + (long long)j0 * g.cpitch + i0;

		// stage the tile's five inputs (coalesced 16-byte cp.async; rows are 128-byte aligned)
		for (int c = tid; c < SGS_TK * SGS_TJ * (SGS_TI / 2) && !(f.debug & 2); c += SGS_NT) {
			const int i2 = c % (SGS_TI / 2), row = c / (SGS_TI / 2);
			const int rj = row % SGS_TJ, rk = row / SGS_TJ;
			if (rk < tk && rj < tj && i0 + 2 * i2 < g.cpitch) {
				const long long cv = ctile + (long long)rk * g.cplane + (long long)rj * g.cpitch + 2 * i2;
				const long long vv = vtile + (long long)rk * g.vplane + (long long)rj * g.vpitch + 2 * i2;
				const int so = (rk * SGS_TJ + rj) * SGS_TI + 2 * i2;
				cp_async16(s_in + so, in0 + (ILU ? cv : vv));   // the ILU recurrence starts from the diagonal (coefficient layout)
				cp_async16(s_c0 + so, c0 + cv);
				cp_async16(s_c1 + so, c1 + cv);
				cp_async16(s_c2 + so, c2 + cv);
				if (!ILU) cp_async16(s_dv + so, dinv + cv);
			}
		}
		// wait for the three predecessor tiles (the loads above do not depend on them)
		if (tid < 3) {
			int pred = -1;
			if (FWD) {
				if (tid == 0 && ci > 0) pred = tile - 1;
				if (tid == 1 && cj > 0) pred = tile - f.ntx;
				if (tid == 2 && ck > 0) pred = tile - f.ntx * f.nty;
			} else {
				if (tid == 0 && ci < f.ntx - 1) pred = tile + 1;
				if (tid == 1 && cj < f.nty - 1) pred = tile + f.ntx;
				if (tid == 2 && ck < f.ntz - 1) pred = tile + f.ntx * f.nty;
			}
			if (pred >= 0 && !(f.debug & 8)) {
				const volatile int *fl = f.flags + pred;
				while (*fl != epoch) { }
			}
			if (!(f.debug & 4)) __threadfence();
		}
		__syncthreads();
		// halo rows of `out` from the neighbouring tiles / ghost layer, straight from L2 (ld.cg):
		//   j-halo: row j0-1 (FWD) or j0+tj (BWD) for every level;  k-halo: plane k0-1 (FWD) or k0+tk (BWD) for every row
		const int dn = FWD ? -1 : 1;
		for (int c = tid; c < SGS_TK * SGS_TI; c += SGS_NT) {
			const int i = c % SGS_TI, q = c / SGS_TI;
			if (q < tk && i < ti) s_hj[c] = __ldcg(out + vtile + (long long)q * g.vplane + (long long)(FWD ? -1 : tj) * g.vpitch + i);
		}
		for (int c = tid; c < SGS_TJ * SGS_TI; c += SGS_NT) {
			const int i = c % SGS_TI, q = c / SGS_TI;
			if (q < tj && i < ti) s_hk[c] = __ldcg(out + vtile + (long long)(FWD ? -1 : tk) * g.vplane + (long long)q * g.vpitch + i);
		}
		const bool active = (lj < tj) && (lk < tk);
		// logical position of this pencil in the sweep order
		const int b = FWD ? lj : tj - 1 - lj, cc = FWD ? lk : tk - 1 - lk;
		double wi = 0.0;                         // the i-neighbour: first the adjacent tile's / ghost value, then our own last value
		if (active) wi = __ldcg(out + vtile + (long long)lk * g.vplane + (long long)lj * g.vpitch + (FWD ? -1 : ti));
		cp_async_wait_all();
		__syncthreads();

		const int nsteps = (f.debug & 1) ? 0 : ti + tj + tk - 2;
		const int rowbase = (lk * SGS_TJ + lj) * SGS_TI;
		const int wslot = (lk + 1) * WP + (lj + 1);
		for (int h = 0; h < nsteps; h++) {
			const int c = h - b - cc;                           // logical i-position of this pencil at step h
			if (active && c >= 0 && c < ti) {
				const int li = FWD ? c : ti - 1 - c;
				const double *prev = s_w + ((h + 1) & 1) * WPL;
				const double wj = (b == 0) ? s_hj[lk * SGS_TI + li] : prev[wslot + dn];
				const double wk = (cc == 0) ? s_hk[lj * SGS_TI + li] : prev[wslot + dn * WP];
				const int so = rowbase + li;
				double val;
				if (ILU) {
					double t = __dsub_rn(s_in[so], __ddiv_rn(s_c0[so], wi));
					t = __dsub_rn(t, __ddiv_rn(s_c1[so], wj));
					val = __dsub_rn(t, __ddiv_rn(s_c2[so], wk));
				} else if (FWD) {
					double t = __dsub_rn(s_in[so], __dmul_rn(s_c0[so], wi));
					t = __dsub_rn(t, __dmul_rn(s_c1[so], wj));
					t = __dsub_rn(t, __dmul_rn(s_c2[so], wk));
					val = __dmul_rn(t, s_dv[so]);
				} else {
					double t = __dmul_rn(s_c0[so], wi);
					t = __dadd_rn(t, __dmul_rn(s_c1[so], wj));
					t = __dadd_rn(t, __dmul_rn(s_c2[so], wk));
					val = __dsub_rn(s_in[so], __dmul_rn(s_dv[so], t));
				}
				wi = val;
				s_w[(h & 1) * WPL + wslot] = val;
				s_in[so] = val;                                   // results replace the staged input
			}
			__syncthreads();
		}
		// results back to global, coalesced 16-byte stores (odd row lengths: the last element alone)
		for (int c = tid; c < SGS_TK * SGS_TJ * (SGS_TI / 2); c += SGS_NT) {
			const int i2 = c % (SGS_TI / 2), row = c / (SGS_TI / 2);
			const int rj = row % SGS_TJ, rk = row / SGS_TJ;
			if (rk < tk && rj < tj && 2 * i2 < ti) {
				const int so = (rk * SGS_TJ + rj) * SGS_TI + 2 * i2;
				double *dst = out + vtile + (long long)rk * g.vplane + (long long)rj * g.vpitch + 2 * i2;
				if (2 * i2 + 1 < ti) st2(dst, make_double2(s_in[so], s_in[so + 1]));
				else *dst = s_in[so];
			}
		}
		if (!(f.debug & 4)) __threadfence();
		__syncthreads();
		if (tid == 0) {
			if (!(f.debug & 4)) __threadfence();
			*(volatile int *)(f.flags + tile) = epoch;
		}
	}
}

constexpr size_t SGS_SMEM = (5 * SGS_TILE + (SGS_TK + SGS_TJ) * SGS_TI + 2 * (SGS_TK + 2) * (SGS_TJ + 2)) * sizeof(double);

// ---------------------------------------------------------------- warp-per-tile dataflow (the default)
//
// Same flags, dispatch order and arithmetic as sgs_dataflow_kernel, but ONE WARP owns a tile and nothing is block-wide.
// What was measured on the way (profiles/r1_sgs_phases.md):
//   * the block version runs load -> wait -> sweep -> store one after the other behind block barriers, ~16 us per tile;
//   * a lone warp is bound by instruction ISSUE, not by the fp64 chain (dmul/dadd 8.5 cycles, 64-bit shuffle 28 cycles): a first
//     warp version that streamed its operands through per-lane cp.async rings spent ~130 instructions = 650-800 cycles per
//     step on ring indexing, divergent bodies and copy issue; the step loop below is branch-free (~66 instructions for two points
//     per lane, 135-150 cycles);
//   * a poll loop run by three lanes leaves the warp split and every later shuffle takes the divergent slow path (1640 instead
//     of 154 cycles per step): the poll is a warp-uniform vote;
//   * cp.async.bulk per 256-byte row costs ~55 cycles of issue each (17600 cycles per tile); 16-byte cp.async by all lanes: 2500;
//   * an 8-byte store per point makes L2 fetch the rest of each 32-byte sector and the fence that follows waits for it (5000
//     cycles under load): results leave as whole rows.
// Design (P = pencils per lane, tile = TI x 4P x 8 points):
//   - lane (m, c) = (lane >> 3, lane & 7) owns the (j,k) pencils b = P m .. P m + P-1, c (logical = sweep-order positions) and
//     marches along i; with P = 2 the two points it computes in a step lie on one hyperplane, so two chains interleave;
//   - neighbour values travel by warp shuffle (j-neighbour: lane-8, or with P = 2 the lane's own other pencil; k-neighbour:
//     lane-1); the tile's halo rows and its i-neighbour column come from L2 after the predecessors' flags;
//   - the inputs of the tile are staged in shared memory with cp.async (issued for the NEXT tile right after the flag of the
//     current one, so they fly during the wait).  Row rho(pencil) of a stream sits at rho*(TI+2) doubles with rho chosen so
//     that the lanes of a half-warp, each at its own skewed i, hit 16 different banks; with P = 2 the second pencil is row rho+8,
//     so ONE shared pointer, advanced by one element per step, addresses all operands with immediate offsets;
//   - every lane evaluates every step (out-of-range steps read padding and are not committed); results overwrite the staged
//     input and leave as whole rows: first what other tiles of this sweep need (the last i-column through a per-tile mailbox,
//     the rows of the last j and the last k), then fence + flag, then the rest.
//   - r2, the default on one rank and for slab-local sweeps (SgsFlow::mail): the hand-off needs neither fence nor flag.  Every tile
//     owns three boxes (last-j rows, last-k rows, last i-column) in a buffer that holds a sentinel outside a sweep; the warp stores
//     its faces there the moment its step loop ends -- only where a tile of this sweep will read them -- and its successors poll
//     the data itself, re-arm the boxes and sweep; all rows then go to `out`, which only later kernels read.  12-16 % faster at
//     every size, bit-identical (profiles/r2_sgs_tile_mail.md).  A sweep THROUGH z-slabs keeps the flag protocol (the cross-rank
//     window is flag-based).
//   - r2 re-entry, BLOCKS (SgsFlow::layers / jlayers; S3D_SGS_BLOCKS on one device, inside every slab of S3D_SGS_OVERLAP): what bounds the
//     exact sweep up to ~256^3 is its chain of nx/TI + ny/TJ + nz/8 tile hops, and no build of the hop got it under ~3.6 us.  So the
//     CHAIN is cut: the grid is cut into blocks of planes and of rows that sweep concurrently, each preceded (in sweep order) by ONE
//     plane / row of the neighbouring block with zero inflow beyond -- computed inside the tile, never stored.  For the kernel a tile
//     row / layer is then not [TJ cj, TJ cj + TJ) / [8 ck, 8 ck + 8) but an entry of a small table {first, count, predecessor |
//     successor inside the block, plane not stored}; tiles are dispatched by their wavefront number inside their block.  Another
//     operator than the reference's sweep (iteration counts within a few per cent: north_star's gate for a parallel SGS), stated in
//     the oracle on top of the reference's half-sweeps and matched bit for bit.  256^3: 0.72 -> 0.43 ms (profiles/r2_sgs_blocks.md).
//   - measured and off: a second staging area (SgsFlow::dbuf) -- time follows the warps per SM, not the overlap of copies and sweep.
constexpr int SGW_PAD = 32;                // slack around the staged block: out-of-range steps read (and discard) it
constexpr int SGW_TK = 8;
// db: a second staging area, so that the next tile's copies fly during the sweep of this one (SgsFlow::dbuf)
constexpr size_t sgw_smem(int na, int ti, int p, int db = 0)
{
	return (size_t)(3 * SGW_PAD + na * 32 * p * (ti + 2) + (8 + 4 * p) * (ti + 2) + (db ? na * 32 * p * (ti + 2) + SGW_PAD : 0)) * sizeof(double) + 16;
}

__device__ __forceinline__ int ld_acquire(const int *p)
{
	int v;
	asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ int ld_acquire_sys(const int *p)
{
	int v;
	asm volatile("ld.acquire.sys.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ double2 ldcg2(const double *p)
{
	double2 v;
	asm volatile("ld.global.cg.v2.f64 {%0, %1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
	return v;
}

// bank-conflict-free row numbering for a row stride == 2 (mod 16) doubles: lane (m, c) at step h reads element h - s of its row,
// s = b + c, so the bank (8-byte units) is (2 rho - s + h) mod 16; rho is chosen so that this is a bijection on each half-warp.
// The backward sweep reads element ti-1 - (h - s): bank (2 rho + s - h + const), so it takes the negated low part.
template <int P, bool FWD> __device__ __forceinline__ int sgw_row(int b, int c)
{
	const int lo = (P == 2) ? 5 * (b >> 1) + 2 * (c >> 1) : ((b + c + 1) >> 1) + (c >> 1) + 4 * (b & 1);
	const int hi = (P == 2) ? (b >> 2) * 4 + (c & 1) * 2 + (b & 1) : 2 * (b & 1) + (c & 1);
	return 8 * hi + ((FWD ? lo : -lo) & 7);
}

template <int MODE, int TI, int P>
__global__ void __launch_bounds__(32)
sgs_warpflow_kernel(const GridDev g, const double *__restrict__ in0, const double *__restrict__ c0,
                    const double *__restrict__ c1, const double *__restrict__ c2, const double *__restrict__ dinv,
                    double *out, const SgsFlow f, const int *stop)
{
	constexpr bool FWD = (MODE != 1);
	constexpr bool ILU = (MODE == 2);
	constexpr int NA = ILU ? 4 : 5;
	constexpr int DIR = FWD ? 1 : -1;
	constexpr int TJ = 4 * P, TK = SGW_TK, NROWS = TJ * TK;
	constexpr int R = TI + 2, ASZ = NROWS * R;
	constexpr int CH = TI / 2;                               // 16-byte chunks per row
	static_assert(R % 16 == 2, "row stride must be 2 modulo 16 (see sgw_row)");
	constexpr unsigned FULL = 0xffffffffu;
	if (s3d_stopped(stop)) return;
	const int epoch = __ldcg(f.epoch_p);
	extern __shared__ __align__(16) double smem[];
	double *s_a = smem + SGW_PAD;                             // [stream][row rho][R]
	double *s_hj = s_a + NA * ASZ + SGW_PAD;                  // [c][R]  j-neighbour rows of the adjacent tile / ghost row
	double *s_hk = s_hj + 8 * R;                              // [b][R]  k-neighbour rows of the adjacent tile / ghost plane
	double *s_other = s_hk + TJ * R + SGW_PAD;                // f.dbuf: the second staging area (the two swap roles after every tile)

	const int lane = threadIdx.x;
	const int m = lane >> 3, cc = lane & 7;
	const int b0 = P * m, b1 = P * m + 1;                     // logical (sweep-order) j positions of the lane's pencils (b1: P = 2)
	const int s0 = b0 + cc;                                   // step at which pencil 0 starts; pencil 1 starts one later
	const int rho0 = sgw_row<P, FWD>(b0, cc);                      // P = 2: pencil 1 is row rho0 + 8
	const int ntiles = f.ntx * f.nty * f.ntz;
	const int nwarps = gridDim.x;

	// stage a tile's input rows: 16-byte cp.async, CH lanes per row.  Issued for the NEXT tile as soon as the flag of the current
	// one is out, so the copies fly while the warp waits for the predecessors.
	// blocks: the layer entries of a tile (planes, rows) are fetched the moment the tile is known and travel in registers from then on
	auto layers_of = [&](int t, int4 &Lk, int4 &Lj) {
		if (f.layers) { Lk = __ldg(f.layers + t / (f.ntx * f.nty)); Lj = __ldg(f.jlayers + (t / f.ntx) % f.nty); }
	};
	auto stage = [&](int t, double *s_dst, const int4 &Lk, const int4 &Lj) {
		const int ci = t % f.ntx, cj = (t / f.ntx) % f.nty, ck = t / (f.ntx * f.nty);
		const int i0 = ci * TI;
		int j0 = cj * TJ, tj = min(TJ, g.ny - j0);
		int k0 = ck * TK, tk = min(TK, g.nz - k0);
		if (f.layers) { k0 = Lk.x; tk = Lk.y; j0 = Lj.x; tj = Lj.y; }
		const long long vtile = g.voff + (long long)k0 * g.vplane + (long long)j0 * g.vpitch + i0;
		const long long ctile = (long long)k0 * g.cplane + (long long)j0 * g.cpitch + i0;
		__syncwarp();                                           // the last tile's reads of the staging area come first
#pragma unroll
		for (int u = 0; u < NROWS * CH / 32; u++) {
			const int q = lane + 32 * u, i2 = q % CH, row = q / CH, rb = row >> 3, rc = row & 7;   // row of logical pencil (rb, rc), chunk i2
			if (rb < tj && rc < tk && i0 + 2 * i2 < g.cpitch) {
				const int pj = FWD ? rb : tj - 1 - rb, pk = FWD ? rc : tk - 1 - rc;
				const long long cv = ctile + (long long)pk * g.cplane + (long long)pj * g.cpitch + 2 * i2;
				const long long vv = vtile + (long long)pk * g.vplane + (long long)pj * g.vpitch + 2 * i2;
				double *dst = s_dst + sgw_row<P, FWD>(rb, rc) * R + 2 * i2;
				cp_async16(dst, in0 + (ILU ? cv : vv));
				cp_async16(dst + ASZ, c0 + cv);
				cp_async16(dst + 2 * ASZ, c1 + cv);
				cp_async16(dst + 3 * ASZ, c2 + cv);
				if (!ILU) cp_async16(dst + 4 * ASZ, dinv + cv);
			}
		}
	};

	int ticket = blockIdx.x;                                  // the first ticket is the warp's own number, later ones come from the counter
	int tile = (ticket < ntiles) ? f.order[FWD ? ticket : ntiles - 1 - ticket] : -1;
	int4 Lk = make_int4(0, 0, 0, 0), Lj = Lk, nLk = Lk, nLj = Lk;   // layer entries of this tile / of the next one
	if (tile >= 0) { layers_of(tile, Lk, Lj); stage(tile, s_a, Lk, Lj); }
	while (tile >= 0) {
		long long tp0 = 0, tp1 = 0, tp2 = 0, tp3 = 0;
		if (f.prof) tp0 = clock64();
		long long gt0 = 0, gt1 = 0, gt2 = 0, gt3 = 0, gt4 = 0;   // global-timer stamps of this tile (timing experiments: S3D_SGS_TILELOG)
		if (f.prof) asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(gt0));
		const int ci = tile % f.ntx, cj = (tile / f.ntx) % f.nty, ck = tile / (f.ntx * f.nty);
		const int i0 = ci * TI;
		int j0 = cj * TJ, tj = min(TJ, g.ny - j0);
		int k0 = ck * TK, tk = min(TK, g.nz - k0);
		bool kpred = FWD ? (ck > 0) : (ck < f.ntz - 1), ksucc = FWD ? (ck < f.ntz - 1) : (ck > 0);
		bool jpred = FWD ? (cj > 0) : (cj < f.nty - 1), jsucc = FWD ? (cj < f.nty - 1) : (cj > 0);
		int skipk = -(1 << 30), skipj = -(1 << 30);               // blocks: the plane / row of this tile whose results stay in the tile
		if (f.layers) {
			k0 = Lk.x; tk = Lk.y; kpred = Lk.z & 1; ksucc = (Lk.z >> 1) & 1; skipk = Lk.w;
			j0 = Lj.x; tj = Lj.y; jpred = Lj.z & 1; jsucc = (Lj.z >> 1) & 1; skipj = Lj.w;
		}
		const int ti = min(TI, g.nx - i0);
		const long long vtile = g.voff + (long long)k0 * g.vplane + (long long)j0 * g.vpitch + i0;
		const bool act0 = (b0 < tj) && (cc < tk), act1 = (P == 2) && (b1 < tj) && (cc < tk);
		// z-slabs: the first plane of tiles (in sweep order) takes its k-halo from the window, the last one feeds the next rank's
		const bool xin = f.x_in_plane && (FWD ? ck == 0 : ck == f.ntz - 1);
		const bool xout = f.x_out_plane && (FWD ? ck == f.ntz - 1 : ck == 0);
		const int xidx = cj * f.ntx + ci;
		const long long inpl = (g.voff - g.vplane) + (long long)j0 * g.vpitch + i0;   // the tile's first column inside a plane
		const int lk = FWD ? cc : tk - 1 - cc;
		const int lj0 = FWD ? b0 : tj - 1 - b0, lj1 = FWD ? b1 : tj - 1 - b1;
		const long long vv0 = vtile + (long long)lk * g.vplane + (long long)lj0 * g.vpitch;
		const long long vv1 = vtile + (long long)lk * g.vplane + (long long)lj1 * g.vpitch;

		// where the halo comes from (addresses only; the loads follow the wait): rows 0..7 are the j-halo (row j0-1 / j0+tj at
		// level c), rows 8..8+TJ-1 the k-halo (plane k0-1 / k0+tk at row b), then the i-neighbours of the lane's pencils
		constexpr int NH = (8 + TJ) * CH / 32;                  // halo chunks per lane
		const double *hp[NH];
		// mailbox hand-off: the faces come from the predecessor tiles' mailboxes (data is its own flag); bit u of hmail: hp[u] is one
		const bool mailmode = (f.mail != nullptr);
		unsigned hmail = 0;
		const double *mjp = reinterpret_cast<const double *>(f.mail) + (long long)(FWD ? tile - f.ntx : tile + f.ntx) * f.mail_stride;
		const double *mkp = reinterpret_cast<const double *>(f.mail) + (long long)(FWD ? tile - f.ntx * f.nty : tile + f.ntx * f.nty) * f.mail_stride + 8 * TI;
#pragma unroll
		for (int u = 0; u < NH; u++) {
			const int q = lane + 32 * u, row = q / CH, i2 = q % CH;
			hp[u] = nullptr;
			if (2 * i2 < ti) {
				if (row < 8) {
					if (row < tk) {
						if (mailmode && jpred) { hp[u] = mjp + row * TI + 2 * i2; hmail |= 1u << u; }
						else if (!(f.jlayers && !jpred)) hp[u] = out + vtile + (long long)(FWD ? row : tk - 1 - row) * g.vplane + (long long)(FWD ? -1 : tj) * g.vpitch + 2 * i2;
					}
				} else {
					const int r = row - 8;
					const bool kz = f.layers ? !kpred : FWD ? (f.kzero_lo && ck == 0) : (f.kzero_hi && ck == f.ntz - 1);   // z-blocks: zero inflow into a block
					if (r < tj && !kz) {
						if (mailmode && kpred) { hp[u] = mkp + r * TI + 2 * i2; hmail |= 1u << u; }
						else hp[u] = (xin ? f.x_in_plane + inpl : out + vtile + (long long)(FWD ? -1 : tk) * g.vplane) + (long long)(FWD ? r : tj - 1 - r) * g.vpitch + 2 * i2;
					}
				}
			}
		}
		const bool iedge = FWD ? (ci == 0) : (ci == f.ntx - 1);    // no tile there: the ghost column of `out`
		const double *box = mailmode ? reinterpret_cast<const double *>(f.mail) + (long long)(FWD ? tile - 1 : tile + 1) * f.mail_stride + (8 + TJ) * TI
		                             : f.iface + (long long)(FWD ? tile - 1 : tile + 1) * 64;
		const double *wp0 = !act0 ? nullptr : iedge ? out + vv0 + (FWD ? -1 : ti) : box + b0 * 8 + cc;
		const double *wp1 = !act1 ? nullptr : iedge ? out + vv1 + (FWD ? -1 : ti) : box + b1 * 8 + cc;
		// ---- wait for the three predecessor tiles (warp-uniform vote, see above); with mailboxes the data below is the wait
		if (!mailmode) {
			const int *fp = nullptr;
			const int sk = f.ntx * f.nty;
			if (FWD) {
				if (lane == 0 && ci > 0) fp = f.flags + tile - 1;
				if (lane == 1 && cj > 0) fp = f.flags + tile - f.ntx;
				if (lane == 2) fp = (ck > 0) ? f.flags + tile - sk : xin ? f.x_in_flags + xidx : nullptr;
			} else {
				if (lane == 0 && ci < f.ntx - 1) fp = f.flags + tile + 1;
				if (lane == 1 && cj < f.nty - 1) fp = f.flags + tile + f.ntx;
				if (lane == 2) fp = (ck < f.ntz - 1) ? f.flags + tile + sk : xin ? f.x_in_flags + xidx : nullptr;
			}
			bool done = (fp == nullptr) || (f.debug & 1);
			const long long tw0 = xin ? clock64() : 0;
			for (;;) {
				if (!done) done = ((xin ? ld_acquire_sys(fp) : ld_acquire(fp)) == epoch);
				if (__all_sync(FULL, done)) break;
				if (xin && clock64() - tw0 > (40ll << 30)) __trap();   // ~20 s without the neighbouring rank: fail loudly instead of hanging
			}
		}
		__syncwarp();
		if (f.prof) tp1 = clock64();
		// we are about to be busy for a few microseconds: claim the next ticket now, look its tile up before the sweep
		int nticket = 0, ntile = -1;   // (nticket: the RAW counter value until the look-up; an add right behind the atomic would wait for its round trip)
		double last0 = 0.0, last1 = 0.0;                      // the lane's newest value of either pencil
		if (lane == 0) nticket = (int)atomicAdd(f.ticket, 1u);
		// ---- halo values (from L2): all loads first, then the stores
		{
			if (mailmode && f.probe) {
				// a TIGHT probe first: one value per predecessor (of the chunk each of them stores last), one load per round, so that the
				// arrival is noticed within an L2 round trip; the full fetch-and-check below (~150 instructions per round) then usually
				// passes at once.  Correctness does not rest on the probe: whatever is still missing below is polled there.
				const double *pr = nullptr;
				if (lane == 0 && !iedge) pr = box + (tj - 1) * 8 + (tk - 1);
				if (lane == 1 && jpred) pr = mjp + (tk - 1) * TI + 2 * ((ti - 1) >> 1);
				if (lane == 2 && kpred && !xin) pr = mkp + (tj - 1) * TI + 2 * ((ti - 1) >> 1);
				const long long tq0 = clock64();
				for (;;) {
					const bool here = (pr == nullptr) || ((unsigned long long)__double_as_longlong(__ldcg(pr)) != SGW_SENT);
					if (__all_sync(FULL, here)) break;
					if (clock64() - tq0 > (8ll << 30)) __trap();
				}
			}
			double2 hv[NH];
#pragma unroll
			for (int u = 0; u < NH; u++) hv[u] = hp[u] ? ldcg2(hp[u]) : make_double2(0.0, 0.0);
			double wi0 = wp0 ? __ldcg(wp0) : 0.0, wi1 = wp1 ? __ldcg(wp1) : 0.0;   // i-neighbours: the adjacent tile's last column / the ghost column
			if (mailmode) {
				// every value that comes from a mailbox is polled until it is there (the producing warp stores it the moment its sweep is
				// over: no fence, no flag), then the sentinel goes back so that the box is clean for the next sweep
				const double sent = __longlong_as_double((long long)SGW_SENT);
				auto is_sent = [](double v) { return (unsigned long long)__double_as_longlong(v) == SGW_SENT; };
				const bool mi0 = wp0 && !iedge, mi1 = wp1 && !iedge;
				const long long tw0 = clock64();
				for (;;) {
					bool ok = !(mi0 && is_sent(wi0)) && !(mi1 && is_sent(wi1));
#pragma unroll
					for (int u = 0; u < NH; u++) {
						const int i2 = (lane + 32 * u) % CH;
						if ((hmail >> u) & 1u) ok = ok && !is_sent(hv[u].x) && (2 * i2 + 1 >= ti || !is_sent(hv[u].y));
					}
					if (__all_sync(FULL, ok)) break;
#pragma unroll
					for (int u = 0; u < NH; u++) {
						const int i2 = (lane + 32 * u) % CH;
						if (((hmail >> u) & 1u) && (is_sent(hv[u].x) || (2 * i2 + 1 < ti && is_sent(hv[u].y)))) hv[u] = ldcg2(hp[u]);
					}
					if (mi0 && is_sent(wi0)) wi0 = __ldcg(wp0);
					if (mi1 && is_sent(wi1)) wi1 = __ldcg(wp1);
					if (clock64() - tw0 > (8ll << 30)) __trap();   // seconds without the producing tile: fail loudly instead of hanging
				}
				if (f.prof) asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(gt1));
#pragma unroll
				for (int u = 0; u < NH; u++)
					if ((hmail >> u) & 1u) __stcg(reinterpret_cast<double2 *>(const_cast<double *>(hp[u])), make_double2(sent, sent));
				if (mi0) __stcg(const_cast<double *>(wp0), sent);
				if (mi1) __stcg(const_cast<double *>(wp1), sent);
			}
#pragma unroll
			for (int u = 0; u < NH; u++) {
				const int q = lane + 32 * u, row = q / CH, i2 = q % CH;
				*reinterpret_cast<double2 *>(s_hj + row * R + 2 * i2) = hv[u];
			}
			// the ticket is back by now: look the next tile up (the load completes during the sweep)
			if (lane == 0) { nticket += nwarps; ntile = (nticket < ntiles) ? f.order[FWD ? nticket : ntiles - 1 - nticket] : -1; }
			cp_async_wait_all();
			__syncwarp();
			if (f.dbuf) {
				// two staging areas: the next tile's copies start now and fly during this sweep, the publish and the copy-out
				ntile = __shfl_sync(FULL, ntile, 0);
				if (ntile >= 0) { layers_of(ntile, nLk, nLj); stage(ntile, s_other, nLk, nLj); }
			}
			if (f.prof) { tp2 = clock64(); asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(gt2)); }

			// ---- the sweep: pencil 0 is at i-position u0 = h - s0 (li = u0 forward, ti-1-u0 backward), pencil 1 at u0 - 1
			const int start = FWD ? -s0 : ti - 1 + s0;
			double *pa = s_a + rho0 * R + start;                 // pencil 0 operand; pencil 1: + 8R - DIR
			const double *phj = s_hj + cc * R + start;           // used by lanes with b0 == 0
			const double *phk = s_hk + b0 * R + start;           // used by lanes with c == 0; pencil 1: + R - DIR
			last0 = wi0; last1 = wi1;
			const int nsteps = ti + tj + tk - 2 + (P - 1);       // P = 2: pencil 1 of a lane runs one step behind pencil 0
			constexpr int E1 = 8 * R - DIR;
			const bool jhalo = (m == 0), khalo = (cc == 0);
			// operands of the step are fetched one step ahead so that no shared-memory latency sits on the fp64 chain
			double a_in = pa[0], a_c0 = pa[ASZ], a_c1 = pa[2 * ASZ], a_c2 = pa[3 * ASZ], a_dv = ILU ? 0.0 : pa[(NA - 1) * ASZ];
			double hj0 = jhalo ? phj[0] : 0.0, hk0 = khalo ? phk[0] : 0.0;
			double b_in = 0.0, b_c0 = 0.0, b_c1 = 0.0, b_c2 = 0.0, b_dv = 0.0, hk1 = 0.0;
			if (P == 2) {
				b_in = pa[E1]; b_c0 = pa[E1 + ASZ]; b_c1 = pa[E1 + 2 * ASZ]; b_c2 = pa[E1 + 3 * ASZ]; b_dv = ILU ? 0.0 : pa[E1 + (NA - 1) * ASZ];
				hk1 = khalo ? phk[R - DIR] : 0.0;
			}
#pragma unroll 2
			for (int h = 0; h < nsteps; h++) {
				const double shj = __shfl_up_sync(FULL, P == 2 ? last1 : last0, 8);
				const double shk0 = __shfl_up_sync(FULL, last0, 1);
				double shk1 = 0.0;
				if (P == 2) shk1 = __shfl_up_sync(FULL, last1, 1);
				pa += DIR; phj += DIR; phk += DIR;
				const double n_a_in = pa[0], n_a_c0 = pa[ASZ], n_a_c1 = pa[2 * ASZ], n_a_c2 = pa[3 * ASZ], n_a_dv = ILU ? 0.0 : pa[(NA - 1) * ASZ];
				const double n_hj0 = jhalo ? phj[0] : 0.0, n_hk0 = khalo ? phk[0] : 0.0;
				double n_b_in = 0.0, n_b_c0 = 0.0, n_b_c1 = 0.0, n_b_c2 = 0.0, n_b_dv = 0.0, n_hk1 = 0.0;
				if (P == 2) {
					n_b_in = pa[E1]; n_b_c0 = pa[E1 + ASZ]; n_b_c1 = pa[E1 + 2 * ASZ]; n_b_c2 = pa[E1 + 3 * ASZ]; n_b_dv = ILU ? 0.0 : pa[E1 + (NA - 1) * ASZ];
					n_hk1 = khalo ? phk[R - DIR] : 0.0;
				}
				const int u0 = h - s0;
				const bool in0 = act0 && (unsigned)u0 < (unsigned)ti;
				const bool in1 = act1 && (unsigned)(u0 - 1) < (unsigned)ti;
				const double wj0 = jhalo ? hj0 : shj, wk0 = khalo ? hk0 : shk0, wk1 = khalo ? hk1 : shk1;
				const double wj1 = last0;                          // pencil 2m of this lane, one step ago, same i
				double v0, v1 = 0.0;
				if (ILU) {
					v0 = __dsub_rn(__dsub_rn(__dsub_rn(a_in, __ddiv_rn(a_c0, last0)), __ddiv_rn(a_c1, wj0)), __ddiv_rn(a_c2, wk0));
					if (P == 2) v1 = __dsub_rn(__dsub_rn(__dsub_rn(b_in, __ddiv_rn(b_c0, last1)), __ddiv_rn(b_c1, wj1)), __ddiv_rn(b_c2, wk1));
				} else if (FWD) {
					v0 = __dmul_rn(__dsub_rn(__dsub_rn(__dsub_rn(a_in, __dmul_rn(a_c0, last0)), __dmul_rn(a_c1, wj0)), __dmul_rn(a_c2, wk0)), a_dv);
					if (P == 2) v1 = __dmul_rn(__dsub_rn(__dsub_rn(__dsub_rn(b_in, __dmul_rn(b_c0, last1)), __dmul_rn(b_c1, wj1)), __dmul_rn(b_c2, wk1)), b_dv);
				} else {
					v0 = __dsub_rn(a_in, __dmul_rn(a_dv, __dadd_rn(__dadd_rn(__dmul_rn(a_c0, last0), __dmul_rn(a_c1, wj0)), __dmul_rn(a_c2, wk0))));
					if (P == 2) v1 = __dsub_rn(b_in, __dmul_rn(b_dv, __dadd_rn(__dadd_rn(__dmul_rn(b_c0, last1), __dmul_rn(b_c1, wj1)), __dmul_rn(b_c2, wk1))));
				}
				// results replace the staged input (pa already points one step ahead)
				if (in0) { last0 = v0; pa[-DIR] = v0; }
				if (P == 2 && in1) { last1 = v1; pa[E1 - DIR] = v1; }
				a_in = n_a_in; a_c0 = n_a_c0; a_c1 = n_a_c1; a_c2 = n_a_c2; a_dv = n_a_dv; hj0 = n_hj0; hk0 = n_hk0;
				if (P == 2) { b_in = n_b_in; b_c0 = n_b_c0; b_c1 = n_b_c1; b_c2 = n_b_c2; b_dv = n_b_dv; hk1 = n_hk1; }
			}
		}
		// ---- publish.  What other tiles of this sweep need goes out first: the last i-column (to the tile's mailbox, straight from
		//      the registers) and the rows of the last j and the last k (whole rows of `out`); then the fence and the flag.  The other
		//      rows follow after the flag -- only later kernels read them.
		long long tpf = 0;
		if (f.prof) { tpf = clock64(); asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(gt3)); }
		ntile = __shfl_sync(FULL, ntile, 0);
		if (!f.dbuf && ntile >= 0) layers_of(ntile, nLk, nLj);     // in flight during the publish and the copy-out
		__syncwarp();
		auto copy_row_chunk = [&](int rb, int rc, int i2) {
			const int pj = FWD ? rb : tj - 1 - rb, pk = FWD ? rc : tk - 1 - rc;
			if (k0 + pk == skipk || j0 + pj == skipj) return;       // blocks: the overlap plane / row belongs to the neighbouring block
			double *dst = out + vtile + (long long)pk * g.vplane + (long long)pj * g.vpitch + 2 * i2;
			const double2 v = *reinterpret_cast<const double2 *>(s_a + sgw_row<P, FWD>(rb, rc) * R + 2 * i2);
			if (2 * i2 + 1 < ti) *reinterpret_cast<double2 *>(dst) = v;
			else *dst = v.x;
		};
		if (mailmode) {
			// MAILBOX hand-off: the faces go into this tile's boxes -- only where a tile of this sweep will read (and re-arm) them --
			// the moment the sweep is over; no fence, no flag.  Then every row goes to `out`, which only later kernels read.
			const bool isucc = FWD ? (ci < f.ntx - 1) : (ci > 0);
			double *mb = reinterpret_cast<double *>(f.mail) + (long long)tile * f.mail_stride;
			if (isucc) {
				if (act0) __stcg(mb + (8 + TJ) * TI + b0 * 8 + cc, last0);
				if (act1) __stcg(mb + (8 + TJ) * TI + b1 * 8 + cc, last1);
			}
#pragma unroll
			for (int u = 0; u < (8 + TJ) * CH / 32; u++) {
				const int q = lane + 32 * u, i2 = q % CH, fr = q / CH;
				const int rb = (fr < 8) ? tj - 1 : fr - 8, rc = (fr < 8) ? fr : tk - 1;
				if (rc < tk && rb < tj && 2 * i2 < ti && ((fr < 8) ? jsucc : ksucc)) {
					const double2 v = *reinterpret_cast<const double2 *>(s_a + sgw_row<P, FWD>(rb, rc) * R + 2 * i2);
					double *dst = mb + fr * TI + 2 * i2;               // rows 0..7: last j at level c = fr; rows 8..: last k at position b = fr - 8
					if (2 * i2 + 1 < ti) __stcg(reinterpret_cast<double2 *>(dst), v);
					else __stcg(dst, v.x);
				}
			}
			if (f.prof) tp3 = clock64();
			if (f.prof && lane == 0) {
				const long long tp4 = clock64();
				asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(gt4));
				if (FWD && tile < 800) {                       // per-tile log of the forward sweep (small grids)
					long long *tl = f.prof + 8 + 5 * tile;
					tl[0] = gt0; tl[1] = gt1; tl[2] = gt2; tl[3] = gt3; tl[4] = gt4;
				}
				atomicAdd((unsigned long long *)f.prof + 0, (unsigned long long)(tp1 - tp0));
				atomicAdd((unsigned long long *)f.prof + 1, (unsigned long long)(tp2 - tp1));   // ticket + polling the boxes + copy completion
				atomicAdd((unsigned long long *)f.prof + 2, (unsigned long long)(tpf - tp2));   // sweep loop
				atomicAdd((unsigned long long *)f.prof + 3, (unsigned long long)(tp4 - tpf));   // boxes out
				atomicAdd((unsigned long long *)f.prof + 4, 1ull);
				atomicAdd((unsigned long long *)f.prof + 5, (unsigned long long)(ti + tj + tk - 2 + (P - 1)));
			}
#pragma unroll
			for (int u = 0; u < NROWS * CH / 32; u++) {
				const int q = lane + 32 * u, i2 = q % CH, row = q / CH, rb = row >> 3, rc = row & 7;
				if (rb < tj && rc < tk && 2 * i2 < ti) copy_row_chunk(rb, rc, i2);
			}
		} else {
		if (act0) f.iface[(long long)tile * 64 + b0 * 8 + cc] = last0;
		if (act1) f.iface[(long long)tile * 64 + b1 * 8 + cc] = last1;
		// face rows: 0..7 the last j (b = tj-1) at every c, then the last k (c = tk-1) at every b < tj-1
#pragma unroll
		for (int u = 0; u < (8 + TJ) * CH / 32; u++) {
			const int q = lane + 32 * u, i2 = q % CH, fr = q / CH;
			const int rb = (fr < 8) ? tj - 1 : fr - 8, rc = (fr < 8) ? fr : tk - 1;
			if (rc < tk && rb < tj && (fr < 8 || rb < tj - 1) && 2 * i2 < ti) copy_row_chunk(rb, rc, i2);
		}
		if (xout) {
			// the last-k rows also go to the next rank's window, straight over NVLink
#pragma unroll
			for (int u = 0; u < (TJ * CH + 31) / 32; u++) {
				const int q = lane + 32 * u, i2 = q % CH, rb = q / CH;
				if (rb < tj && 2 * i2 < ti) {
					const double2 v = *reinterpret_cast<const double2 *>(s_a + sgw_row<P, FWD>(rb, tk - 1) * R + 2 * i2);
					double *dst = f.x_out_plane + inpl + (long long)(FWD ? rb : tj - 1 - rb) * g.vpitch + 2 * i2;
					if (2 * i2 + 1 < ti) *reinterpret_cast<double2 *>(dst) = v;
					else *dst = v.x;
				}
			}
			__threadfence_system();
		} else {
			asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
		}
		__syncwarp();
		if (f.prof) tp3 = clock64();
		if (lane == 0) {
			*(volatile int *)(f.flags + tile) = epoch;
			if (xout) *(volatile int *)(f.x_out_flags + xidx) = epoch;
			if (f.prof) {
				const long long tp4 = clock64();
				atomicAdd((unsigned long long *)f.prof + 0, (unsigned long long)(tp1 - tp0));   // flag wait
				atomicAdd((unsigned long long *)f.prof + 1, (unsigned long long)(tp2 - tp1));   // ticket + halo rows + copy completion
				atomicAdd((unsigned long long *)f.prof + 2, (unsigned long long)(tpf - tp2));   // sweep loop
				atomicAdd((unsigned long long *)f.prof + 3, (unsigned long long)(tp4 - tpf));   // face rows + fence + flag
				atomicAdd((unsigned long long *)f.prof + 4, 1ull);
				atomicAdd((unsigned long long *)f.prof + 5, (unsigned long long)(ti + tj + tk - 2 + (P - 1)));
			}
		}
#pragma unroll
		for (int u = 0; u < NROWS * CH / 32; u++) {
			const int q = lane + 32 * u, i2 = q % CH, row = q / CH, rb = row >> 3, rc = row & 7;
			if (rb < tj - 1 && rc < tk - 1 && 2 * i2 < ti) copy_row_chunk(rb, rc, i2);
		}
		}
		if (f.dbuf) { double *t_ = s_a; s_a = s_other; s_other = t_; }
		else if (ntile >= 0) stage(ntile, s_a, nLk, nLj);
		tile = ntile; Lk = nLk; Lj = nLj;
	}
}

// ---------------------------------------------------------------- warp-per-LINE dataflow
//
// The warp-per-tile kernel pays the skew of its pencils (tj + tk - 2 steps) once per tile: 26 steps for 16 columns.  Here a warp
// owns a whole LINE of the grid -- the 4 x 8 (j,k) pencils of a (cj, ck) cell over ALL nx columns -- and marches through it once:
// the skew is paid once per line, every step computes 32 points, the i-neighbour stays in a register, and there are no flags,
// mailboxes or pipeline drains along i.  The inputs stream through a ring of 32 columns in shared memory:
//   - columns come in groups of 8 (physical, 16-byte aligned); at every step h = 4 (mod 8) the group all lanes have left
//     (logical columns [h-20, h-12)) is copied out to `out` as 64-byte row pieces and the group [h+12, h+20) is staged into the
//     same ring slots with cp.async -- 12 steps before its first use;
//   - dependencies are per (line, block of 16 columns): a line needs the last-j rows of line (cj-1, ck) and the last-k rows of line
//     (cj, ck-1); their flags are polled 8 steps before the block is entered and the halo rows then stream into two small rings;
//   - a block's flag is published 36 steps after its first column was entered (its stores have drained by then, the fence is cheap).
// Same arithmetic, same roundings, same lane mapping and bank map as the tile kernel; lines are dispatched in (cj + ck) order.
constexpr int SGL_W = 32;                  // ring width in columns
constexpr int SGL_R = SGL_W + 2;           // row stride in doubles (== 2 mod 16, see sgw_row)
constexpr size_t sgl_smem(int na, int p) { return (size_t)(na * 32 * p + 8 + 4 * p) * SGL_R * sizeof(double); }

template <int MODE, int P>
__global__ void __launch_bounds__(32)
sgs_lineflow_kernel(const GridDev g, const double *__restrict__ in0, const double *__restrict__ c0,
                    const double *__restrict__ c1, const double *__restrict__ c2, const double *__restrict__ dinv,
                    double *out, const SgsFlow f, const int *stop)
{
	constexpr bool FWD = (MODE != 1);
	constexpr bool ILU = (MODE == 2);
	constexpr int NA = ILU ? 4 : 5;
	constexpr int DIR = FWD ? 1 : -1;
	constexpr int TJ = 4 * P, TK = SGW_TK, NROWS = TJ * TK;
	constexpr int R = SGL_R, ASZ = NROWS * R;
	constexpr int S = TJ + TK - 2 + (P - 1);                 // steps the last pencil of the warp lags the first
	static_assert(S <= 12 && P == 1, "the copy-out / staging schedule below assumes a skew of at most 12 steps and one pencil per lane");
	constexpr unsigned FULL = 0xffffffffu;
	if (s3d_stopped(stop)) return;
	const int epoch = __ldcg(f.epoch_p);
	extern __shared__ __align__(16) double smem[];
	double *s_a = smem;                                       // [stream][row rho][R], column i lives in slot i & 31
	double *s_hj = s_a + NA * ASZ;                            // [c][R]  rows j0-1 / j0+tj of the neighbouring line
	double *s_hk = s_hj + 8 * R;                              // [b][R]  rows of plane k0-1 / k0+tk

	const int lane = threadIdx.x;
	const int m = lane >> 3, cc = lane & 7;
	const int b0 = m;
	const int s0 = b0 + cc;
	const int rho0 = sgw_row<P, FWD>(b0, cc);
	const int nlines = f.nty * f.ntz;
	const int nwarps = gridDim.x;
	// logical columns c = 0 .. ncol-1 in sweep order; backward sweeps start at the top of the last 8-column group, so that
	// logical groups [8n, 8n+8) are physical, 16-byte aligned groups in either direction
	const int gmax = (g.nx - 1) >> 3, itop = 8 * gmax + 7;
	const int ncol = FWD ? g.nx : itop + 1;
	const int ngroups = gmax + 1, nblocks = (ncol + 15) >> 4;
	const int fstride = f.ntx;                                // flags per line (>= nblocks of either direction)
	const int nsteps = ncol + S;

	int ticket = blockIdx.x;
	int line = (ticket < nlines) ? f.order[FWD ? ticket : nlines - 1 - ticket] : -1;
	while (line >= 0) {
		const int cj = line % f.nty, ck = line / f.nty;
		const int j0 = cj * TJ, k0 = ck * TK;
		const int tj = min(TJ, g.ny - j0), tk = min(TK, g.nz - k0);
		const long long vline = g.voff + (long long)k0 * g.vplane + (long long)j0 * g.vpitch;
		const long long cline = (long long)k0 * g.cplane + (long long)j0 * g.cpitch;
		const bool act0 = (b0 < tj) && (cc < tk);
		const int pj = FWD ? (cj > 0 ? line - 1 : -1) : (cj < f.nty - 1 ? line + 1 : -1);
		const int pk = FWD ? (ck > 0 ? line - f.nty : -1) : (ck < f.ntz - 1 ? line + f.nty : -1);

		// what this lane moves in a burst: NROWS / 8 (row, 16-byte chunk) items; everything that does not depend on the group
		// is computed once per line
		constexpr int NU = NROWS / 8;
		long long it_cv[NU], it_vv[NU];       // global offsets of the item's row in the coefficient / vector layout, + 2*chunk
		int it_s[NU];                         // shared-memory offset of the item's row (doubles), or -1 when the row is outside the grid
#pragma unroll
		for (int u = 0; u < NU; u++) {
			const int q = lane + 32 * u, ch = q & 3, row = q >> 2, rb = row >> 3, rc = row & 7;
			const int jp = FWD ? rb : tj - 1 - rb, kp = FWD ? rc : tk - 1 - rc;
			it_cv[u] = cline + (long long)kp * g.cplane + (long long)jp * g.cpitch + 2 * ch;
			it_vv[u] = vline + (long long)kp * g.vplane + (long long)jp * g.vpitch + 2 * ch;
			it_s[u] = (rb < tj && rc < tk) ? sgw_row<P, FWD>(rb, rc) * R + 2 * ch : -1;
		}
		auto stage_group = [&](int n) {                         // logical group n of every row, all streams
			const int i8 = 8 * (FWD ? n : gmax - n), sl = i8 & (SGL_W - 1);
#pragma unroll
			for (int u = 0; u < NU; u++)
				if (it_s[u] >= 0) {
					double *dst = s_a + it_s[u] + sl;
					const long long cv = it_cv[u] + i8;
					cp_async16(dst, in0 + (ILU ? cv : it_vv[u] + i8));
					cp_async16(dst + ASZ, c0 + cv);
					cp_async16(dst + 2 * ASZ, c1 + cv);
					cp_async16(dst + 3 * ASZ, c2 + cv);
					if (!ILU) cp_async16(dst + 4 * ASZ, dinv + cv);
				}
		};
		auto copy_group = [&](int n) {                          // results of logical group n: whole 64-byte row pieces
			const int i8 = 8 * (FWD ? n : gmax - n), sl = i8 & (SGL_W - 1);
#pragma unroll
			for (int u = 0; u < NU; u++) {
				const int i = i8 + 2 * ((lane + 32 * u) & 3);
				if (it_s[u] >= 0 && i < g.nx) {
					const double2 v = *reinterpret_cast<const double2 *>(s_a + it_s[u] + sl);
					double *dst = out + it_vv[u] + i8;
					if (i + 1 < g.nx) *reinterpret_cast<double2 *>(dst) = v;
					else *dst = v.x;
				}
			}
		};
		auto stage_halo = [&](int bl) {                         // logical block bl (16 columns) of the 8 + TJ halo rows
			const int i16 = FWD ? 16 * bl : itop - 16 * bl - 15;
#pragma unroll
			for (int u = 0; u < (8 + TJ) * 8 / 32; u++) {
				const int q = lane + 32 * u, ch = q & 7, row = q >> 3;
				const int i = i16 + 2 * ch;
				if (i >= 0 && i < g.nx) {
					if (row < 8) {
						if (row < tk)
							cp_async16(s_hj + row * R + (i & (SGL_W - 1)),
							           out + vline + (long long)(FWD ? row : tk - 1 - row) * g.vplane + (long long)(FWD ? -1 : tj) * g.vpitch + i);
					} else {
						const int r = row - 8;
						if (r < tj)
							cp_async16(s_hk + r * R + (i & (SGL_W - 1)),
							           out + vline + (long long)(FWD ? -1 : tk) * g.vplane + (long long)(FWD ? r : tj - 1 - r) * g.vpitch + i);
					}
				}
			}
		};
		// flags of the two predecessor lines: lane 0 watches (cj-1, ck), lane 1 (cj, ck-1).  The load is issued early (flag_peek) and
		// looked at 8 steps later (flag_wait, a warp-uniform vote -- a partial-warp poll loop leaves the warp split)
		const int mypred = (lane == 0) ? pj : (lane == 1) ? pk : -1;
		int peeked = 0;
		auto flag_peek = [&](int bl) { if (mypred >= 0) peeked = ld_acquire(f.flags + (long long)mypred * fstride + bl); };
		auto flag_wait = [&](int bl) {
			bool done = (mypred < 0) || (peeked == epoch);
			for (;;) {
				if (__all_sync(FULL, done)) break;
				if (!done) done = (ld_acquire(f.flags + (long long)mypred * fstride + bl) == epoch);
			}
		};

		// ---- prologue: two groups of inputs, the first block of halo rows, the ghost column
		__syncwarp();
		flag_peek(0);
		stage_group(0); cp_async_commit();
		if (ngroups > 1) stage_group(1);
		cp_async_commit();
		const int jp0 = FWD ? b0 : tj - 1 - b0, kp0 = FWD ? cc : tk - 1 - cc;
		const long long vrow0 = vline + (long long)kp0 * g.vplane + (long long)jp0 * g.vpitch;
		double last0 = act0 ? __ldcg(out + vrow0 + (FWD ? -1 : g.nx)) : 0.0;   // the ghost column
		int nticket = 0, nline = -1;
		if (lane == 0) nticket = (int)atomicAdd(f.ticket, 1u) + nwarps;
		flag_wait(0);
		stage_halo(0); cp_async_commit();
		cp_async_wait_all();
		__syncwarp();

		// ---- the march.  Lane at logical column u0 = h - s0, physical column FWD ? u0 : itop - u0.
		double *pa_base = s_a + rho0 * R;
		const double *phj_base = s_hj + cc * R, *phk_base = s_hk + b0 * R;
		const bool jhalo = (m == 0), khalo = (cc == 0);
		int icol = FWD ? -s0 : itop + s0;                       // physical column of this lane at step h
		int slot = icol & (SGL_W - 1);
		double a_in = pa_base[slot], a_c0 = pa_base[slot + ASZ], a_c1 = pa_base[slot + 2 * ASZ], a_c2 = pa_base[slot + 3 * ASZ];
		double a_dv = ILU ? 0.0 : pa_base[slot + (NA - 1) * ASZ];
		double hj0 = jhalo ? phj_base[slot] : 0.0, hk0 = khalo ? phk_base[slot] : 0.0;
		int next_copy = 0, next_pub = 0;
		long long tq0 = 0, tq_wait = 0, tq_flag = 0, tq_trig = 0, tq = 0;
		if (f.prof) tq0 = clock64();
		// steps come in octets so that the housekeeping sits at compile-time positions of a fully unrolled inner loop
#pragma unroll 1
		for (int h8 = 0; h8 < nsteps; h8 += 8) {
			const bool odd = (h8 & 8) != 0;
#pragma unroll
			for (int q = 0; q < 8; q++) {
				const int h = h8 + q;
				if (q == 0 && !odd) flag_peek(min((h + 16) >> 4, nblocks - 1));            // looked at 8 steps from now
				if (q == 0 && odd) {                                                          // h = 8 (mod 16): halo rows of the next block
					const int bl = (h + 8) >> 4;
					if (f.prof) tq = clock64();
					if (bl < nblocks) { flag_wait(bl); stage_halo(bl); }
					cp_async_commit();
					if (f.prof) tq_flag += clock64() - tq;
				}
				if (q == 4) {                                                                 // h = 4 (mod 8)
					if (f.prof) tq = clock64();
					if (!odd && h >= 36 && next_pub < nblocks) {                                // h = 4 (mod 16): the stores of block next_pub left 8+ steps ago
						asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
						__syncwarp();
						if (lane == 0) *(volatile int *)(f.flags + (long long)line * fstride + next_pub) = epoch;
						next_pub++;
					}
					__syncwarp();
					if (h >= 20 && next_copy < ngroups) copy_group(next_copy++);
					__syncwarp();
					const int nin = (h + 12) >> 3;
					if (nin < ngroups) stage_group(nin);
					cp_async_commit();
					if (f.prof) tq_trig += clock64() - tq;
				}
				if (q == 7) {                                                                 // the group (and halo block) entered at step h+1 has landed
					if (f.prof) tq = clock64();
					cp_async_wait_group<1>(); __syncwarp();
					if (f.prof) tq_wait += clock64() - tq;
				}
				const double shj = __shfl_up_sync(FULL, last0, 8);
				const double shk0 = __shfl_up_sync(FULL, last0, 1);
				const int islot = slot;                              // where this step's result goes
				const bool in0 = act0 && (unsigned)icol < (unsigned)g.nx;
				icol += DIR;
				slot = icol & (SGL_W - 1);
				const double n_a_in = pa_base[slot], n_a_c0 = pa_base[slot + ASZ], n_a_c1 = pa_base[slot + 2 * ASZ], n_a_c2 = pa_base[slot + 3 * ASZ];
				const double n_a_dv = ILU ? 0.0 : pa_base[slot + (NA - 1) * ASZ];
				const double n_hj0 = jhalo ? phj_base[slot] : 0.0, n_hk0 = khalo ? phk_base[slot] : 0.0;
				const double wj0 = jhalo ? hj0 : shj, wk0 = khalo ? hk0 : shk0;
				double v0;
				if (ILU) v0 = __dsub_rn(__dsub_rn(__dsub_rn(a_in, __ddiv_rn(a_c0, last0)), __ddiv_rn(a_c1, wj0)), __ddiv_rn(a_c2, wk0));
				else if (FWD) v0 = __dmul_rn(__dsub_rn(__dsub_rn(__dsub_rn(a_in, __dmul_rn(a_c0, last0)), __dmul_rn(a_c1, wj0)), __dmul_rn(a_c2, wk0)), a_dv);
				else v0 = __dsub_rn(a_in, __dmul_rn(a_dv, __dadd_rn(__dadd_rn(__dmul_rn(a_c0, last0), __dmul_rn(a_c1, wj0)), __dmul_rn(a_c2, wk0))));
				if (in0) { last0 = v0; pa_base[islot] = v0; }          // results replace the staged input
				a_in = n_a_in; a_c0 = n_a_c0; a_c1 = n_a_c1; a_c2 = n_a_c2; a_dv = n_a_dv; hj0 = n_hj0; hk0 = n_hk0;
			}
			if (h8 == 16 && lane == 0) nline = (nticket < nlines) ? f.order[FWD ? nticket : nlines - 1 - nticket] : -1;
		}
		if (f.prof && lane == 0) {
			atomicAdd((unsigned long long *)f.prof + 0, (unsigned long long)tq_flag);              // flag waits + halo issue
			atomicAdd((unsigned long long *)f.prof + 1, (unsigned long long)tq_wait);              // cp.async waits
			atomicAdd((unsigned long long *)f.prof + 2, (unsigned long long)(clock64() - tq0));    // whole march
			atomicAdd((unsigned long long *)f.prof + 3, (unsigned long long)tq_trig);              // copy-out + staging + publish
			atomicAdd((unsigned long long *)f.prof + 4, 1ull);
			atomicAdd((unsigned long long *)f.prof + 5, (unsigned long long)nsteps);
		}
		// ---- epilogue: the groups still in the ring, then the remaining flags
		cp_async_wait_all();
		__syncwarp();
		while (next_copy < ngroups) copy_group(next_copy++);
		asm volatile("fence.acq_rel.gpu;\n" ::: "memory");
		__syncwarp();
		if (lane == 0)
			for (int bl = next_pub; bl < nblocks; bl++) *(volatile int *)(f.flags + (long long)line * fstride + bl) = epoch;
		if (nsteps <= 16 && lane == 0) nline = (nticket < nlines) ? f.order[FWD ? nticket : nlines - 1 - nticket] : -1;   // short lines never reach the look-up above
		line = __shfl_sync(FULL, nline, 0);
	}
}

// ---------------------------------------------------------------- red-black

// phase 0: y_R = r_R dinv                       (red:   (i+j+k) even in the reference's 1-based indices = odd+... see parity())
// phase 1: y_B = (r_B - sum_6 A_s y_s) dinv ;  z_B = y_B
// phase 2: z_R = y_R - dinv sum_6 A_s z_s
__device__ __forceinline__ int parity(int i, int j, int k, int k0) { return (i + j + k + k0 + 3) & 1; }   // 1-based (i+1)+(j+1)+(k+1)

template <int PHASE>
__global__ void __launch_bounds__(256)
sgs_rb_kernel(const GridDev g, int k0, const double *__restrict__ A, const double *__restrict__ dinv,
              const double *__restrict__ r, double *y, double *z, const int *stop)
{
	if (s3d_stopped(stop)) return;
	const unsigned nx2 = (unsigned)(g.nx + 1) >> 1;
	const unsigned total = nx2 * (unsigned)g.ny * (unsigned)g.nz;
	constexpr int COLOUR = (PHASE == 1) ? 1 : 0;   // 0 = red (even), 1 = black
	for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
		const unsigned row = it / nx2;
		const int k = (int)(row / (unsigned)g.ny);
		const int j = (int)(row - (unsigned)k * (unsigned)g.ny);
		int i = 2 * (int)(it - row * nx2);
		if (parity(i, j, k, k0) != COLOUR) i += 1;   // exactly one point of each aligned pair has the colour
		if (i >= g.nx) continue;
		const long long v = g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + i;
		const long long c = (long long)k * g.cplane + (long long)j * g.cpitch + i;
		const long long cs = g.cstride;
		if (PHASE == 0) {
			y[v] = __dmul_rn(r[v], dinv[c]);
		} else if (PHASE == 1) {
			double t = __dsub_rn(r[v], __dmul_rn(A[c], y[v - 1]));
			t = __dsub_rn(t, __dmul_rn(A[c + cs], y[v - g.vpitch]));
			t = __dsub_rn(t, __dmul_rn(A[c + 2 * cs], y[v - g.vplane]));
			t = __dsub_rn(t, __dmul_rn(A[c + 4 * cs], y[v + 1]));
			t = __dsub_rn(t, __dmul_rn(A[c + 5 * cs], y[v + g.vpitch]));
			t = __dsub_rn(t, __dmul_rn(A[c + 6 * cs], y[v + g.vplane]));
			t = __dmul_rn(t, dinv[c]);
			y[v] = t;
			z[v] = t;
		} else {
			double s = __dmul_rn(A[c], z[v - 1]);
			s = __dadd_rn(s, __dmul_rn(A[c + cs], z[v - g.vpitch]));
			s = __dadd_rn(s, __dmul_rn(A[c + 2 * cs], z[v - g.vplane]));
			s = __dadd_rn(s, __dmul_rn(A[c + 4 * cs], z[v + 1]));
			s = __dadd_rn(s, __dmul_rn(A[c + 5 * cs], z[v + g.vpitch]));
			s = __dadd_rn(s, __dmul_rn(A[c + 6 * cs], z[v + g.vplane]));
			z[v] = __dsub_rn(y[v], __dmul_rn(dinv[c], s));
		}
	}
}

// ---- red-black on COLOUR-PACKED coefficients.  A colour occupies every other entry of a row, so the kernel above uses half of every
// 32-byte sector of the six off-diagonal arrays and of dinv it touches (r1: 5.2 ms at 512^3, 0.38 of the copy peak, ~190 B of DRAM traffic
// per point against 96 algorithmic).  rb_pack_kernel copies the seven arrays once per operator into a layout where the entries of one
// colour are contiguous: packed[colour][array a = 0..6][(k ny + j) cp2 + i / 2], arrays in the order A0 A1 A2 A4 A5 A6 dinv,
// cp2 = roundup((nx + 1) / 2, 16).  sgs_rb_packed_kernel is sgs_rb_kernel reading its coefficients from there (same expressions, same
// roundings, bit-identical results); the vectors keep their layout (their sectors are shared by the two colours' neighbour reads).
__global__ void __launch_bounds__(256)
rb_pack_kernel(const GridDev g, int k0, const double *__restrict__ A, const double *__restrict__ dinv, double *__restrict__ packed, int cp2, long long astride)
{
	const long long total = (long long)g.nx * g.ny * g.nz;
	for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
		const int i = (int)(t % g.nx);
		const long long row = t / g.nx;
		const int j = (int)(row % g.ny), k = (int)(row / g.ny);
		const long long c = (long long)k * g.cplane + (long long)j * g.cpitch + i;
		const int colour = parity(i, j, k, k0);
		double *dst = packed + (long long)colour * 7 * astride + row * cp2 + (i >> 1);
		const long long cs = g.cstride;
		dst[0] = A[c]; dst[astride] = A[c + cs]; dst[2 * astride] = A[c + 2 * cs];
		dst[3 * astride] = A[c + 4 * cs]; dst[4 * astride] = A[c + 5 * cs]; dst[5 * astride] = A[c + 6 * cs];
		dst[6 * astride] = dinv[c];
	}
}

template <int PHASE>
__global__ void __launch_bounds__(256)
sgs_rb_packed_kernel(const GridDev g, int k0, const double *__restrict__ packed, int cp2, long long astride,
                     const double *__restrict__ r, double *y, double *z, const int *stop)
{
	if (s3d_stopped(stop)) return;
	const unsigned nx2 = (unsigned)(g.nx + 1) >> 1;
	const unsigned total = nx2 * (unsigned)g.ny * (unsigned)g.nz;
	constexpr int COLOUR = (PHASE == 1) ? 1 : 0;   // 0 = red (even), 1 = black
	const double *P = packed + (long long)COLOUR * 7 * astride;
	for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
		const unsigned row = it / nx2;
		const int k = (int)(row / (unsigned)g.ny);
		const int j = (int)(row - (unsigned)k * (unsigned)g.ny);
		const int ip = (int)(it - row * nx2);
		int i = 2 * ip;
		if (parity(i, j, k, k0) != COLOUR) i += 1;   // exactly one point of each aligned pair has the colour
		if (i >= g.nx) continue;
		const long long v = g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + i;
		const long long c = (long long)row * cp2 + ip;
		if (PHASE == 0) {
			y[v] = __dmul_rn(r[v], ld_stream(P + 6 * astride + c));
		} else if (PHASE == 1) {
			double t = __dsub_rn(r[v], __dmul_rn(ld_stream(P + c), y[v - 1]));
			t = __dsub_rn(t, __dmul_rn(ld_stream(P + astride + c), y[v - g.vpitch]));
			t = __dsub_rn(t, __dmul_rn(ld_stream(P + 2 * astride + c), y[v - g.vplane]));
			t = __dsub_rn(t, __dmul_rn(ld_stream(P + 3 * astride + c), y[v + 1]));
			t = __dsub_rn(t, __dmul_rn(ld_stream(P + 4 * astride + c), y[v + g.vpitch]));
			t = __dsub_rn(t, __dmul_rn(ld_stream(P + 5 * astride + c), y[v + g.vplane]));
			t = __dmul_rn(t, ld_stream(P + 6 * astride + c));
			y[v] = t;
			z[v] = t;
		} else {
			double s = __dmul_rn(ld_stream(P + c), z[v - 1]);
			s = __dadd_rn(s, __dmul_rn(ld_stream(P + astride + c), z[v - g.vpitch]));
			s = __dadd_rn(s, __dmul_rn(ld_stream(P + 2 * astride + c), z[v - g.vplane]));
			s = __dadd_rn(s, __dmul_rn(ld_stream(P + 3 * astride + c), z[v + 1]));
			s = __dadd_rn(s, __dmul_rn(ld_stream(P + 4 * astride + c), z[v + g.vpitch]));
			s = __dadd_rn(s, __dmul_rn(ld_stream(P + 5 * astride + c), z[v + g.vplane]));
			z[v] = __dsub_rn(y[v], __dmul_rn(ld_stream(P + 6 * astride + c), s));
		}
	}
}

// StrILU helpers.  products: P_s[c] = A_s[c] * A_{s+4}[neighbour], zero where the neighbour does not exist (the reference's
// `if(i > 0)` branches, s3d_ilu.cpp:45-56); fill: every entry (ghosts included) of a vector-layout array;
// invert: dinv[c] = 1 / d[v].
__global__ void __launch_bounds__(256)
ilu_products_kernel(const GridDev g, const double *__restrict__ A, double *__restrict__ P)
{
	const unsigned total = (unsigned)g.nx * (unsigned)g.ny * (unsigned)g.nz;
	for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
		const int i = it % g.nx, j = (it / g.nx) % g.ny, k = it / (g.nx * g.ny);
		const long long c = (long long)k * g.cplane + (long long)j * g.cpitch + i;
		const long long cs = g.cstride;
		P[c] = (i > 0) ? __dmul_rn(A[c], A[4 * cs + c - 1]) : 0.0;
		P[cs + c] = (j > 0) ? __dmul_rn(A[cs + c], A[5 * cs + c - g.cpitch]) : 0.0;
		P[2 * cs + c] = (k > 0) ? __dmul_rn(A[2 * cs + c], A[6 * cs + c - g.cplane]) : 0.0;
	}
}

__global__ void __launch_bounds__(256) fill_all_kernel(double *p, long long n, double v)
{
	for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void __launch_bounds__(256)
ilu_invert_kernel(const GridDev g, const double *__restrict__ d, double *__restrict__ dinv)
{
	const unsigned total = (unsigned)g.nx * (unsigned)g.ny * (unsigned)g.nz;
	for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
		const int i = it % g.nx, j = (it / g.nx) % g.ny, k = it / (g.nx * g.ny);
		dinv[(long long)k * g.cplane + (long long)j * g.cpitch + i] =
		    __ddiv_rn(1.0, d[g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + i]);
	}
}

// ---------------------------------------------------------------- synchronous Jacobi sweeps on the triangular systems
//
// The reference's threaded apply (threadedapply = true, napplysweeps = n) runs n asynchronous, unordered in-place sweeps per
// triangular factor.  Its deterministic GPU counterpart: n SYNCHRONOUS sweeps, each a fully parallel streaming pass that
// recomputes every point from the previous sweep's values (ping-pong buffers), starting from zero.  One sweep is Jacobi; the
// result converges to the exact lexicographic solve as n grows (bit-identical from n = nx+ny+nz-2 on).  48 B/point per sweep.
// PHASE 0: out = dinv (r - A0 in[i-1] - A1 in[j-1] - A2 in[k-1])      PHASE 1: out = y - dinv (A4 in[i+1] + A5 in[j+1] + A6 in[k+1])
// first: the previous iterate is zero (skip its loads).
template <int PHASE>
__global__ void __launch_bounds__(256)
sgs_sweep_kernel(const GridDev g, const double *__restrict__ A, const double *__restrict__ dinv, const double *__restrict__ rhs,
                 const double *__restrict__ in, double *__restrict__ out, int first, const int *stop)
{
	if (s3d_stopped(stop)) return;
	const unsigned nx2 = (unsigned)(g.nx + 1) >> 1;
	const unsigned total = nx2 * (unsigned)g.ny * (unsigned)g.nz;
	const long long cs = g.cstride;
	const int o = PHASE == 0 ? 0 : 4;
	for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < total; it += gridDim.x * blockDim.x) {
		const unsigned row = it / nx2;
		const int i = 2 * (int)(it - row * nx2);
		const int k = (int)(row / (unsigned)g.ny);
		const int j = (int)(row - (unsigned)k * (unsigned)g.ny);
		const long long v = g.voff + (long long)k * g.vplane + (long long)j * g.vpitch + i;
		const long long c = (long long)k * g.cplane + (long long)j * g.cpitch + i;
		const int np = (i + 1 < g.nx) ? 2 : 1;
		const long long di = PHASE == 0 ? -1 : 1, dj = PHASE == 0 ? -(long long)g.vpitch : g.vpitch, dk = PHASE == 0 ? -g.vplane : g.vplane;
		double res[2];
		for (int e = 0; e < np; e++) {
			const double ni = first ? 0.0 : in[v + e + di], nj = first ? 0.0 : in[v + e + dj], nk = first ? 0.0 : in[v + e + dk];
			const double a0 = ld_stream(A + (o + 0) * cs + c + e), a1 = ld_stream(A + (o + 1) * cs + c + e), a2 = ld_stream(A + (o + 2) * cs + c + e);
			const double dv = dinv[c + e], b = rhs[v + e];
			if (PHASE == 0) {
				double t = __dsub_rn(b, __dmul_rn(a0, ni));
				t = __dsub_rn(t, __dmul_rn(a1, nj));
				t = __dsub_rn(t, __dmul_rn(a2, nk));
				res[e] = __dmul_rn(t, dv);
			} else {
				double t = __dmul_rn(a0, ni);
				t = __dadd_rn(t, __dmul_rn(a1, nj));
				t = __dadd_rn(t, __dmul_rn(a2, nk));
				res[e] = __dsub_rn(b, __dmul_rn(dv, t));
			}
		}
		if (np == 2) st2(out + v, make_double2(res[0], res[1]));
		else out[v] = res[0];
	}
}

unsigned ew_blocks(const s3d_ctx *ctx, const s3d_grid *g)
{
	const unsigned long long total = (unsigned long long)((g->nx + 1) / 2) * g->ny * g->nz;
	unsigned long long b = (total + 255) / 256;
	const unsigned long long cap = (unsigned long long)ctx->sm_count * 16;
	if (b > cap) b = cap;
	return (unsigned)(b < 1 ? 1 : b);
}

inline int sgw_tile_i(const s3d_ctx *ctx) { return ctx->sgs_tile_i == 32 ? 32 : 16; }
// pencils per lane: 2 (tile TI x 8 x 8) has the shorter dependency chain across the grid, 1 (tile TI x 4 x 8) twice as many tiles in
// flight per SM.  Measured on B200 with the mailbox hand-off (profiles/r2_sgs_tile_mail.md): 256^3 0.770 vs 0.799 ms, 320^3 1.176 vs
// 1.120, 384^3 1.813 vs 1.521, 512^3 3.859 vs 3.022 ms: the switch sits at 24 Mi points (r1, flags and fences: 40 Mi).
inline int sgw_pencils(const s3d_ctx *ctx, const s3d_grid *g)
{
	if (ctx->sgs_pencils == 1 || ctx->sgs_pencils == 2) return ctx->sgs_pencils;
	// every rank of a z-slab run must pick the same tile shape (the cross-rank flags are per (ci, cj)): decide on the mean slab
	const long long nzl = (g->nranks > 1) ? ((long long)g->nz_global + g->nranks - 1) / g->nranks : g->nz;
	return ((long long)g->nx * g->ny * nzl >= (24ll << 20)) ? 1 : 2;
}

// (the SgsFlow among the arguments says whether the launch uses the second staging area)
template <int MODE, typename... Args> void sgw_launch_db(s3d_ctx *ctx, const s3d_grid *g, unsigned nblocks, int db, Args... args)
{
	constexpr int NA = (MODE == 2) ? 4 : 5;
	const int ti = sgw_tile_i(ctx), p = sgw_pencils(ctx, g);
	if (ti == 32 && p == 2) sgs_warpflow_kernel<MODE, 32, 2><<<nblocks, 32, sgw_smem(NA, 32, 2, db), ctx->stream>>>(args...);
	else if (ti == 32) sgs_warpflow_kernel<MODE, 32, 1><<<nblocks, 32, sgw_smem(NA, 32, 1, db), ctx->stream>>>(args...);
	else if (p == 2) sgs_warpflow_kernel<MODE, 16, 2><<<nblocks, 32, sgw_smem(NA, 16, 2, db), ctx->stream>>>(args...);
	else sgs_warpflow_kernel<MODE, 16, 1><<<nblocks, 32, sgw_smem(NA, 16, 1, db), ctx->stream>>>(args...);
}
template <int MODE, typename A0, typename A1, typename A2, typename A3, typename A4, typename A5, typename A6>
void sgw_launch(s3d_ctx *ctx, const s3d_grid *g, unsigned nblocks, GridDev gd, A0 a0, A1 a1, A2 a2, A3 a3, A4 a4, A5 a5, const SgsFlow &f, A6 a6)
{
	sgw_launch_db<MODE>(ctx, g, nblocks, f.dbuf, gd, a0, a1, a2, a3, a4, a5, f, a6);
}

template <int TI, int P> int sgw_configure(int db)
{
	// (the attribute is set for the larger, double-buffered size; it covers both)
	cudaFuncSetAttribute(sgs_warpflow_kernel<0, TI, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sgw_smem(5, TI, P, 1));
	cudaFuncSetAttribute(sgs_warpflow_kernel<1, TI, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sgw_smem(5, TI, P, 1));
	cudaFuncSetAttribute(sgs_warpflow_kernel<2, TI, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sgw_smem(4, TI, P, 1));
	int a = 1, b = 1;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, sgs_warpflow_kernel<0, TI, P>, 32, sgw_smem(5, TI, P, db));
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, sgs_warpflow_kernel<1, TI, P>, 32, sgw_smem(5, TI, P, db));
	return std::max(1, std::min(a, b));
}

__global__ void __launch_bounds__(256) fill_u64_sgs_kernel(unsigned long long *p, size_t n, unsigned long long v)
{
	for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}

// start of a dataflow sweep: rewind the ticket counter and move to the next epoch -- on the device, so that a captured CUDA graph
// of a solver iteration can be replayed (a host-side epoch would be frozen into the graph)
__global__ void sgs_begin_kernel(int *ctl, const int *stop)
{
	if (s3d_stopped(stop)) return;
	ctl[0] = 0;
	ctl[1] += 1;
}

// which build of the exact sweep runs: the streaming kernel (4) works inside one slab; a sweep that has to run THROUGH the z-slabs
// (exact lexicographic ordering on a decomposed grid) is the warp-per-tile kernel's (0)
inline int sgs_effective_variant(const s3d_ctx *ctx, const s3d_grid *g, bool through_slabs)
{
	int v = ctx->sgs_variant;
	if ((v == 4 || v == 5 || v == 6 || v == 7) && through_slabs && g->nranks > 1) return 0;
	if (v == 6) return 6;   // the register-streaming kernel whatever the size
	if (v == 7) return 7;   // the row-scan kernel whatever the size (the reference's sweep to rounding, not bit for bit)
	// 4 = by grid size (measured crossover, profiles/r2_sgs_stream.md), 5 = the streaming kernel whatever the size
	if (v == 4 && (long long)g->nx * g->ny * g->nz > ctx->sgs_stream_max_points) return 0;
	return v == 5 ? 4 : v;
}

// flags / dispatch table / occupancy for the dataflow kernels; returns the number of persistent blocks
int sgs_flow_prepare(s3d_ctx *ctx, const s3d_grid *g, SgsFlow &f, unsigned &nblocks, int variant)
{
	const bool lines = (variant == 3 || variant == 4);
	const int tile_i = lines ? 16 : (variant == 0) ? sgw_tile_i(ctx) : SGS_TI;
	const int tile_j = lines ? 4 : (variant == 0) ? 4 * sgw_pencils(ctx, g) : SGS_TJ;
	// lines: "ntx" is the number of 16-column flag blocks of a line (backward sweeps start at the top of the last 8-column group)
	f.ntx = lines ? (((g->nx + 7) / 8) * 8 + 15) / 16 : (g->nx + tile_i - 1) / tile_i;
	f.nty = (g->ny + tile_j - 1) / tile_j; f.ntz = (g->nz + SGS_TK - 1) / SGS_TK;
	const size_t ntiles = (size_t)f.ntx * f.nty * f.ntz;
	if (!ctx->sgs_ctl) {
		S3D_CUDA(ctx, cudaMalloc(&ctx->sgs_ctl, 2 * sizeof(int)));
		S3D_CUDA(ctx, cudaMemsetAsync(ctx->sgs_ctl, 0, 2 * sizeof(int), ctx->stream));
	}
	if (ctx->sgs_flags_cap < ntiles + 1) {
		S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		if (ctx->sgs_flags) S3D_CUDA(ctx, cudaFree(ctx->sgs_flags));
		ctx->sgs_flags = nullptr; ctx->sgs_flags_cap = 0;
		S3D_CUDA(ctx, cudaMalloc(&ctx->sgs_flags, (ntiles + 1) * sizeof(int)));
		S3D_CUDA(ctx, cudaMemsetAsync(ctx->sgs_flags, 0, (ntiles + 1) * sizeof(int), ctx->stream));
		ctx->sgs_flags_cap = ntiles + 1;
	}
	// The epoch is never rewound: it only grows, so whatever an old sweep (of this or another tiling) left in a flag is smaller than
	// every later epoch and cannot satisfy a wait.  On z-slabs this also keeps the ranks' epochs equal whatever their local tile counts.
	static bool configured = false;
	static int blocks_per_sm = 1;
	if (!configured) {
		S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_dataflow_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SGS_SMEM));
		S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_dataflow_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SGS_SMEM));
		S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_dataflow_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SGS_SMEM));
		S3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, sgs_dataflow_kernel<0>, SGS_NT, SGS_SMEM));
		if (blocks_per_sm < 1) blocks_per_sm = 1;
		configured = true;
	}
	const int okey = lines ? -f.ntx : f.ntx;   // lines and tiles never share a dispatch table
	if (ctx->sgs_order_dims[0] != okey || ctx->sgs_order_dims[1] != f.nty || ctx->sgs_order_dims[2] != f.ntz) {
		// counting sort of the tiles (or lines) by hyperplane
		const int ox = lines ? 1 : f.ntx;
		const int nd = ox + f.nty + f.ntz - 2;
		std::vector<int> start(nd + 1, 0), order((size_t)ox * f.nty * f.ntz);
		for (int ck = 0; ck < f.ntz; ck++)
			for (int cj = 0; cj < f.nty; cj++)
				for (int ci = 0; ci < ox; ci++) start[ci + cj + ck + 1]++;
		for (int d = 0; d < nd; d++) start[d + 1] += start[d];
		for (int ck = 0; ck < f.ntz; ck++)
			for (int cj = 0; cj < f.nty; cj++)
				for (int ci = 0; ci < ox; ci++) order[start[ci + cj + ck]++] = (ck * f.nty + cj) * ox + ci;
		S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		if (ctx->sgs_order) S3D_CUDA(ctx, cudaFree(ctx->sgs_order));
		ctx->sgs_order = nullptr;
		S3D_CUDA(ctx, cudaMalloc(&ctx->sgs_order, order.size() * sizeof(int)));
		S3D_CUDA(ctx, cudaMemcpy(ctx->sgs_order, order.data(), order.size() * sizeof(int), cudaMemcpyHostToDevice));
		ctx->sgs_order_dims[0] = okey; ctx->sgs_order_dims[1] = f.nty; ctx->sgs_order_dims[2] = f.ntz;
	}
	f.flags = ctx->sgs_flags;
	f.iface = ctx->sgs_iface;
	f.mail = nullptr; f.mail_stride = 0; f.probe = 0; f.layers = nullptr; f.jlayers = nullptr; f.dbuf = 0;
	f.x_in_plane = nullptr; f.x_in_flags = nullptr; f.x_out_plane = nullptr; f.x_out_flags = nullptr;
	f.kzero_lo = 0; f.kzero_hi = 0;
	f.order = ctx->sgs_order;
	f.sleep_ns = 0;
	f.debug = ctx->sgs_sleep_ns;   // reuses the tuning slot: debug bit mask
	f.prof = ctx->sgs_prof;
	f.ticket = reinterpret_cast<unsigned *>(ctx->sgs_ctl);
	f.epoch_p = ctx->sgs_ctl + 1;
	if (variant == 4) { nblocks = 0; return S3D_OK; }   // the streaming kernel sizes its own launch (sgs_stream.cu)
	if (lines) {
		static bool lconfigured = false;
		static int lines_per_sm = 1;
		if (!lconfigured) {
			S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_lineflow_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sgl_smem(5, 1)));
			S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_lineflow_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sgl_smem(5, 1)));
			S3D_CUDA(ctx, cudaFuncSetAttribute(sgs_lineflow_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sgl_smem(4, 1)));
			int a = 1, b = 1;
			S3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, sgs_lineflow_kernel<0, 1>, 32, sgl_smem(5, 1)));
			S3D_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, sgs_lineflow_kernel<1, 1>, 32, sgl_smem(5, 1)));
			lines_per_sm = std::max(1, std::min(a, b));
			lconfigured = true;
		}
		int wps = lines_per_sm;
		if (ctx->sgs_warps_per_sm > 0) wps = std::min(wps, ctx->sgs_warps_per_sm);
		nblocks = (unsigned)std::min<size_t>((size_t)f.nty * f.ntz, (size_t)ctx->sm_count * wps);
		return S3D_OK;
	}
	if (variant == 0) {
		if (ctx->sgs_iface_cap < ntiles * 64) {
			S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
			if (ctx->sgs_iface) S3D_CUDA(ctx, cudaFree(ctx->sgs_iface));
			ctx->sgs_iface = nullptr; ctx->sgs_iface_cap = 0;
			S3D_CUDA(ctx, cudaMalloc(&ctx->sgs_iface, ntiles * 64 * sizeof(double)));
			ctx->sgs_iface_cap = ntiles * 64;
		}
		f.iface = ctx->sgs_iface;
		if (ctx->sgs_tile_mail) {
			// mailbox hand-off (the caller drops it again for a sweep that runs THROUGH z-slabs): all sentinel outside a sweep
			const size_t stride = (size_t)(8 + tile_j) * tile_i + 64, need = ntiles * stride;
			if (ctx->sgs_mail_cap < need) {
				S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
				if (ctx->sgs_mail) S3D_CUDA(ctx, cudaFree(ctx->sgs_mail));
				ctx->sgs_mail = nullptr; ctx->sgs_mail_cap = 0;
				S3D_CUDA(ctx, cudaMalloc(&ctx->sgs_mail, need * sizeof(unsigned long long)));
				fill_u64_sgs_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->sgs_mail, need, SGW_SENT);
				S3D_LAUNCHED(ctx);
				ctx->sgs_mail_cap = need;
			}
			f.mail = ctx->sgs_mail; f.mail_stride = (int)stride;
			// measured (profiles/r2_sgs_tile_mail.md): the tight probe takes 0.3 us off a hop (64^3: 0.162 -> 0.150 ms) but its polling traffic
			// costs throughput on large grids (512^3: 3.02 -> 3.25 ms): up to 2 Mi points only
			f.probe = ((long long)g->nx * g->ny * g->nz <= (2ll << 20)) ? 1 : 0;
		}
		static bool wconfigured = false;
		static int warps_per_sm[2][4] = {{1, 1, 1, 1}, {1, 1, 1, 1}};   // [second staging area][tile_i == 32][pencils == 2]
		if (!wconfigured) {
			for (int db = 0; db < 2; db++) {
				warps_per_sm[db][0] = sgw_configure<16, 1>(db); warps_per_sm[db][1] = sgw_configure<16, 2>(db);
				warps_per_sm[db][2] = sgw_configure<32, 1>(db); warps_per_sm[db][3] = sgw_configure<32, 2>(db);
			}
			S3D_CUDA(ctx, cudaGetLastError());
			wconfigured = true;
		}
		// second staging area (tuning key sgs_dbuf: 1 on, 0 off, -1 by size): pays where the sweep is bound by tiles in flight, not by its chain
		f.dbuf = (ctx->sgs_dbuf == 1 || (ctx->sgs_dbuf < 0 && (long long)g->nx * g->ny * g->nz >= ctx->sgs_dbuf_min_points)) ? 1 : 0;
		int wps = warps_per_sm[f.dbuf][(sgw_tile_i(ctx) == 32 ? 2 : 0) + (sgw_pencils(ctx, g) == 2 ? 1 : 0)];
		if (ctx->sgs_warps_per_sm > 0) wps = std::min(wps, ctx->sgs_warps_per_sm);
		if (wps < 1) wps = 1;
		nblocks = (unsigned)std::min<size_t>(ntiles, (size_t)ctx->sm_count * wps);
		return S3D_OK;
	}
	nblocks = (unsigned)std::min<size_t>(ntiles, (size_t)ctx->sm_count * blocks_per_sm);
	return S3D_OK;
}

}  // namespace

extern "C" S3D_HIDDEN int s3d_internal_dev_alloc(s3d_ctx *ctx, void **d_out, size_t bytes);
extern "C" S3D_HIDDEN bool s3d_internal_has_slack(const s3d_ctx *ctx, const void *d_ptr);
extern "C" S3D_HIDDEN int s3d_internal_sgs_window(s3d_ctx *ctx, const s3d_grid *g, int ntx, int nty, s3d_sgs_window *out);
extern "C" S3D_HIDDEN int s3d_internal_sgs_stream(s3d_ctx *ctx, const s3d_grid *g, int mode, const double *in0, const double *c0,
                                                  const double *c1, const double *c2, const double *dinv, double *out,
                                                  const int *order, int nty, int ntz, unsigned *ticket, int klo_zero, int khi_zero,
                                                  const int *stop);

extern "C" S3D_HIDDEN int s3d_internal_sgs_pencil(s3d_ctx *ctx, const s3d_grid *g, int mode, const double *in0, const double *c0,
                                                  const double *c1, const double *c2, const double *dinv, double *out,
                                                  unsigned *ticket, int klo_zero, int khi_zero, const int *stop);

extern "C" S3D_HIDDEN int s3d_internal_sgs_rowscan(s3d_ctx *ctx, const s3d_grid *g, int mode, const double *in0, const double *c0,
                                                   const double *c1, const double *c2, const double *dinv, double *out,
                                                   unsigned *ticket, int klo_zero, int khi_zero, const int *stop);
extern "C" S3D_HIDDEN int s3d_internal_strilu_async(s3d_ctx *ctx, const s3d_grid *g, const double *a3, const double *P, double *D, int nsweeps);
extern "C" S3D_HIDDEN int s3d_internal_sgs_async(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *d_dinv, const double *d_r,
                                                 double *d_z, double *d_y, int nsweeps);

namespace {
// the ticket counter of the kernels that keep their own dispatch tables (sgs_pencil.cu)
int sgs_ctl_ensure(s3d_ctx *ctx)
{
	if (!ctx->sgs_ctl) {
		S3D_CUDA(ctx, cudaMalloc(&ctx->sgs_ctl, 2 * sizeof(int)));
		S3D_CUDA(ctx, cudaMemsetAsync(ctx->sgs_ctl, 0, 2 * sizeof(int), ctx->stream));
	}
	return S3D_OK;
}
}  // namespace

namespace {
// Blocks on one device (S3D_SGS_BLOCKS): layer tables along k and j and the dispatch tables of both sweep directions, cached per
// (grid, blocks, tile shape).  Along an axis of n points cut into B blocks, a block owns a run of planes (rows); forward, it sweeps them
// preceded by ONE plane of the block below (zero inflow under it) and does not store that plane; backward, followed by one plane of the
// block above -- whose forward values y the block above has stored by then: the launches are ordered -- with zero inflow over it,
// and does not store it either.
struct AxisLayers { std::vector<int4> L[2]; std::vector<int> depth[2]; };   // [sweep direction]; depth: position of a layer inside its block, in sweep order

void axis_blocks(int n, int B, int TL, AxisLayers &out)
{
	constexpr int T = 8;   // the partition does not depend on the tile shape (TL = 4 or 8 rows per tile layer): 7 modulo 8 is also 3 modulo 4
	// block thickness: T - 1 modulo T, so that a block plus its overlap plane fills whole tile layers in either direction (a layer that
	// holds one plane costs a full tile sweep); the first blocks take T more where that evens out the last one, which takes what is
	// left (at least T).  cport.block_bounds states the same.
	B = std::max(1, std::min(B, n / T));
	const int thick = T * std::max(1, (n + B) / (T * B)) - 1;
	while (B > 1 && n - (B - 1) * thick < T) B--;
	const int rem = n - (B - 1) * thick;
	int nthicker = std::max(0, std::min((2 * (rem - n / B) + T) / (2 * T), B - 1));
	while (nthicker > 0 && rem - T * nthicker < T) nthicker--;
	int knext = 0;
	for (int p = 0; p < B; p++) {
		const int k0 = knext, k1 = (p == B - 1) ? n : k0 + thick + (p < nthicker ? T : 0);
		knext = k1;
		for (int dir = 0; dir < 2; dir++) {
			const int lo = (dir == 0 && p > 0) ? k0 - 1 : k0, hi = (dir == 1 && p < B - 1) ? k1 + 1 : k1;
			const int skip = (dir == 0 && p > 0) ? k0 - 1 : (dir == 1 && p < B - 1) ? k1 : -(1 << 30);
			const int nl = (hi - lo + TL - 1) / TL;
			for (int l = 0; l < nl; l++) {
				const int first = lo + l * TL, cnt = std::min(TL, hi - first);
				const bool below = l > 0, above = l < nl - 1;   // a layer of the same block below / above this one
				const int pred = dir == 0 ? below : above, succ = dir == 0 ? above : below;
				out.L[dir].push_back(make_int4(first, cnt, pred | (succ << 1), skip));
				out.depth[dir].push_back(dir == 0 ? l : nl - 1 - l);
			}
		}
	}
}

// host-side plan of one sweep direction: the k layers, the j layers, and the dispatch order of the tiles
struct BlocksPlan { std::vector<int4> kl, jl; std::vector<int> order; int ntx = 0, nty = 0, ntz = 0; };

void sgs_blocks_plan(int nx, int ny, int nz, int Bz, int By, int tile_i, int tile_j, BlocksPlan plan[2])
{
	const int ntx = (nx + tile_i - 1) / tile_i;
	AxisLayers az, ay;
	axis_blocks(nz, Bz, SGS_TK, az);
	axis_blocks(ny, By, tile_j, ay);
	for (int dir = 0; dir < 2; dir++) {
		const int ntz = (int)az.L[dir].size(), nty = (int)ay.L[dir].size();
		// counting sort by the tile's wavefront number inside its block: every block ramps up at once.  The kernel takes
		// order[ticket] forward and order[ntiles - 1 - ticket] backward, so the backward table ascends in the REVERSED wavefront number.
		const int nd = ntx + nty + ntz;
		std::vector<int> start(nd + 1, 0);
		plan[dir].order.assign((size_t)ntx * nty * ntz, 0);
		auto wave = [&](int ci, int cj, int ck) {
			return dir == 0 ? ci + ay.depth[0][cj] + az.depth[0][ck] : nd - 1 - ((ntx - 1 - ci) + ay.depth[1][cj] + az.depth[1][ck]);
		};
		for (int ck = 0; ck < ntz; ck++)
			for (int cj = 0; cj < nty; cj++)
				for (int ci = 0; ci < ntx; ci++) start[wave(ci, cj, ck) + 1]++;
		for (int d = 0; d < nd; d++) start[d + 1] += start[d];
		for (int ck = 0; ck < ntz; ck++)
			for (int cj = 0; cj < nty; cj++)
				for (int ci = 0; ci < ntx; ci++) plan[dir].order[start[wave(ci, cj, ck)]++] = (ck * nty + cj) * ntx + ci;
		plan[dir].kl = az.L[dir]; plan[dir].jl = ay.L[dir];
		plan[dir].ntx = ntx; plan[dir].nty = nty; plan[dir].ntz = ntz;
	}
}

int sgs_blocks_prepare(s3d_ctx *ctx, const s3d_grid *g, int Bz, int By, int tile_i, int tile_j)
{
	s3d_ctx::SgsBlocks &t = ctx->sgsblocks;
	const int key[7] = {g->nx, g->ny, g->nz, Bz, By, tile_i, tile_j};
	if (t.order[0] && !memcmp(key, t.key, sizeof key)) return S3D_OK;
	BlocksPlan plan[2];
	sgs_blocks_plan(g->nx, g->ny, g->nz, Bz, By, tile_i, tile_j, plan);
	S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	for (int dir = 0; dir < 2; dir++) {
		const BlocksPlan &q = plan[dir];
		if (t.order[dir]) S3D_CUDA(ctx, cudaFree(t.order[dir]));
		if (t.layers[dir]) S3D_CUDA(ctx, cudaFree(t.layers[dir]));
		t.order[dir] = nullptr; t.layers[dir] = nullptr;
		S3D_CUDA(ctx, cudaMalloc(&t.order[dir], q.order.size() * sizeof(int)));
		S3D_CUDA(ctx, cudaMalloc(&t.layers[dir], (size_t)(q.ntz + q.nty) * sizeof(int4)));        // [k layers][j layers]
		S3D_CUDA(ctx, cudaMemcpy(t.order[dir], q.order.data(), q.order.size() * sizeof(int), cudaMemcpyHostToDevice));
		S3D_CUDA(ctx, cudaMemcpy(t.layers[dir], q.kl.data(), (size_t)q.ntz * sizeof(int4), cudaMemcpyHostToDevice));
		S3D_CUDA(ctx, cudaMemcpy(t.layers[dir] + q.ntz, q.jl.data(), (size_t)q.nty * sizeof(int4), cudaMemcpyHostToDevice));
		t.ntz[dir] = q.ntz; t.nty[dir] = q.nty;
	}
	memcpy(t.key, key, sizeof key);
	return S3D_OK;
}

// The blocks ordering over grid `g` (one device's grid, or a rank's slab extended by its overlap planes; `gd` its device view, the array
// pointers shifted to match): the grid is cut into blocks of planes (and of rows) that sweep concurrently, each forward sweep preceded by
// one plane (row) of the block below, each backward sweep by one of the block above (sgs_blocks_prepare).  The dependency chain of a
// sweep shrinks from nx/TI + ny/TJ + nz/8 tile hops to about nx/TI + ny/(TJ By) + nz/(8 Bz) + 2, which is what bounds the exact sweep up
// to ~256^3; another operator than the reference's sweep (iteration counts within a few per cent).  *ran = false: one block (or no
// mailboxes) -- the caller runs the plain sweep.
int sgs_blocks_run(s3d_ctx *ctx, const s3d_grid *g, const GridDev &gd, const double *A, const double *d_dinv, const double *d_r, double *d_y,
                   double *d_z, bool *ran)
{
	*ran = false;
	// blocks asked for, or (0) chosen by size: about 32 planes / rows per block (measured, profiles/r2_sgs_blocks.md: 256^3 0.72 -> 0.43 ms,
	// 128^3 0.32 -> 0.15, 64^3 0.143 -> 0.105); none above 96 Mi points, where the sweep is bound by tiles in flight and blocks only add work
	const long long npts = (long long)g->nx * g->ny * g->nz;
	const int Bz_ask = ctx->sgs_blocks > 0 ? ctx->sgs_blocks : npts > (96ll << 20) ? 1 : std::min(g->nz / 32, 16);
	const int By_ask = ctx->sgs_blocks_y > 0 ? ctx->sgs_blocks_y : npts > (96ll << 20) ? 1 : std::min(g->ny / 32, 8);
	const int Bz = std::max(1, std::min(Bz_ask, g->nz / SGS_TK)), By = std::max(1, std::min(By_ask, g->ny / 8));
	if ((Bz == 1 && By == 1) || !ctx->sgs_tile_mail) return S3D_OK;
	// one pencil per lane (tiles of 4 rows, twice as many in flight) once rows are cut into blocks: measured faster there at every size
	const int tile_i = sgw_tile_i(ctx), tile_j = 4 * ((ctx->sgs_pencils == 1 || ctx->sgs_pencils == 2) ? ctx->sgs_pencils : By >= 2 ? 1 : sgw_pencils(ctx, g));
	const int rcb = sgs_blocks_prepare(ctx, g, Bz, By, tile_i, tile_j);
	if (rcb != S3D_OK) return rcb;
	const s3d_ctx::SgsBlocks &t = ctx->sgsblocks;
	s3d_grid gx = *g;                                // sizes the mailboxes / the launch for the larger of the two tilings
	gx.nz = SGS_TK * std::max(t.ntz[0], t.ntz[1]);
	gx.ny = tile_j * std::max(t.nty[0], t.nty[1]);
	const int pencils_asked = ctx->sgs_pencils;
	ctx->sgs_pencils = tile_j / 4;                   // the tile shape is chosen by size: g's choice holds for gx too
	SgsFlow f;
	unsigned nblocks = 0;
	const int rcp = sgs_flow_prepare(ctx, &gx, f, nblocks, 0);
	if (rcp != S3D_OK) { ctx->sgs_pencils = pencils_asked; return rcp; }
	if (!f.mail || (size_t)std::max(t.ntz[0], t.ntz[1]) * std::max(t.nty[0], t.nty[1]) * f.ntx * f.mail_stride > ctx->sgs_mail_cap) {
		ctx->sgs_pencils = pencils_asked;
		return s3d_fail(ctx, S3D_ERR_STATE, "S3D_SGS_BLOCKS: mailbox set-up does not match the block tiling");
	}
	const long long cs = g->cstride;
	for (int dir = 0; dir < 2; dir++) {
		f.ntz = t.ntz[dir]; f.nty = t.nty[dir]; f.order = t.order[dir]; f.layers = t.layers[dir]; f.jlayers = t.layers[dir] + t.ntz[dir];
		const unsigned nb = (unsigned)std::min<size_t>((size_t)f.ntx * f.nty * f.ntz, (size_t)nblocks);
		sgs_begin_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sgs_ctl, ctx->gate);
		ctx->launches++;
		if (dir == 0) sgw_launch<0>(ctx, &gx, nb, gd, d_r, A, A + cs, A + 2 * cs, d_dinv, d_y, f, (const int *)ctx->gate);
		else sgw_launch<1>(ctx, &gx, nb, gd, (const double *)d_y, A + 4 * cs, A + 5 * cs, A + 6 * cs, d_dinv, d_z, f, (const int *)ctx->gate);
		ctx->launches++;
	}
	ctx->sgs_pencils = pencils_asked;
	S3D_CUDA(ctx, cudaGetLastError());
	*ran = true;
	return S3D_OK;
}
}  // namespace

extern "C" {

// The blocks ordering's host-side plan for inspection (no device needed; tests/test_cpu.py walks it): direction 0 = forward, 1 = backward.
// dims_out = {ntx, nty, ntz}; layers_out: 4 ints per layer {first, count, pred | succ << 1, index not stored}, the ntz k layers then the nty
// j layers; order_out: the ntx*nty*ntz tile numbers ((ck * nty + cj) * ntx + ci) as the kernel takes them (forward order[ticket], backward
// order[ntiles - 1 - ticket]).  Null output pointers are skipped (call once for the sizes).
int s3d_sgs_blocks_plan(int nx, int ny, int nz, int blocks, int blocks_y, int tile_i, int pencils, int direction, int *dims_out, int *layers_out,
                        int *order_out)
{
	if (nx < 1 || ny < 1 || nz < 1 || (tile_i != 16 && tile_i != 32) || (pencils != 1 && pencils != 2) || (direction != 0 && direction != 1) || !dims_out)
		return S3D_ERR_ARG;
	const int Bz = std::max(1, std::min(blocks, nz / SGS_TK)), By = std::max(1, std::min(blocks_y, ny / 8));
	BlocksPlan plan[2];
	sgs_blocks_plan(nx, ny, nz, Bz, By, tile_i, 4 * pencils, plan);
	const BlocksPlan &q = plan[direction];
	dims_out[0] = q.ntx; dims_out[1] = q.nty; dims_out[2] = q.ntz;
	if (layers_out) {
		for (int l = 0; l < q.ntz; l++) { const int4 v = q.kl[l]; int *o = layers_out + 4 * l; o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
		for (int l = 0; l < q.nty; l++) { const int4 v = q.jl[l]; int *o = layers_out + 4 * (q.ntz + l); o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
	}
	if (order_out) memcpy(order_out, q.order.data(), q.order.size() * sizeof(int));
	return S3D_OK;
}

// can the exact sweep run through the z-slabs of this grid?  (collective: creates the peer-memory window)
int s3d_sgs_slab_sweeps_available(s3d_ctx *ctx, const s3d_grid *g, int *yes)
{
	S3D_REQUIRE(ctx, g && yes, "null argument");
	*yes = 0;
	if (g->nranks == 1) { *yes = 1; return S3D_OK; }
	if (sgs_effective_variant(ctx, g, true) != 0) return S3D_OK;
	SgsFlow f;
	unsigned nblocks = 0;
	int rc = sgs_flow_prepare(ctx, g, f, nblocks, 0);
	if (rc != S3D_OK) return rc;
	s3d_sgs_window xw;
	rc = s3d_internal_sgs_window(ctx, g, f.ntx, f.nty, &xw);
	if (rc != S3D_OK) return rc;
	*yes = xw.ok;
	return S3D_OK;
}

// S3D_SGS_OVERLAP, once per operator (after s3d_sgs_setup / s3d_strilu_setup): the neighbouring slabs' boundary planes of the seven
// coefficient arrays and of dinv into the slack planes.  Collective.
int s3d_sgs_overlap_setup(s3d_ctx *ctx, const s3d_grid *g, double *A, double *d_dinv)
{
	S3D_REQUIRE(ctx, g && A && d_dinv, "null argument");
	if (g->nranks == 1) return S3D_OK;
	int rc = s3d_coef_halo_exchange(ctx, g, A, S3D_NSTENCIL);
	if (rc != S3D_OK) return rc;
	return s3d_coef_halo_exchange(ctx, g, d_dinv, 1);
}

// Red-black SGS on colour-packed coefficients (see rb_pack_kernel): size of the packed buffer in doubles, the packing (once per operator,
// after s3d_sgs_setup / s3d_strilu_setup), and the application -- bit-identical to s3d_sgs_apply(..., S3D_SGS_REDBLACK, ...).
static inline int rb_cp2(const s3d_grid *g) { return (((g->nx + 1) / 2 + 15) / 16) * 16; }
static inline long long rb_astride(const s3d_grid *g) { return (long long)rb_cp2(g) * g->ny * g->nz; }

int s3d_sgs_rb_pack_size(const s3d_grid *g, size_t *ndoubles)
{
	if (!g || !ndoubles) return S3D_ERR_ARG;
	*ndoubles = (size_t)14 * (size_t)rb_astride(g);
	return S3D_OK;
}

int s3d_sgs_rb_pack_alloc(s3d_ctx *ctx, const s3d_grid *g, double **d_out)
{
	S3D_REQUIRE(ctx, g && d_out, "null argument");
	size_t n = 0;
	s3d_sgs_rb_pack_size(g, &n);
	{ const int rc = s3d_internal_dev_alloc(ctx, (void **)d_out, n * sizeof(double)); if (rc != S3D_OK) return rc; }   // released with s3d_free
	S3D_CUDA(ctx, cudaMemsetAsync(*d_out, 0, n * sizeof(double), ctx->stream));   // the padding is read by nobody, but keep it defined
	return S3D_OK;
}

int s3d_sgs_rb_pack(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *d_dinv, double *d_packed)
{
	S3D_REQUIRE(ctx, g && A && d_dinv && d_packed, "null argument");
	const GridDev gd = to_dev(g);
	rb_pack_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(gd, g->k0, A, d_dinv, d_packed, rb_cp2(g), rb_astride(g));
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}

int s3d_sgs_apply_rb_packed(s3d_ctx *ctx, const s3d_grid *g, const double *d_packed, const double *d_r, double *d_z, double *d_y)
{
	S3D_REQUIRE(ctx, g && d_packed && d_r && d_z && d_y, "null argument");
	S3D_REQUIRE(ctx, d_r != d_z && d_r != d_y && d_z != d_y, "r, z and the scratch vector must be distinct");
	const GridDev gd = to_dev(g);
	const unsigned nb = ew_blocks(ctx, g);
	const int cp2 = rb_cp2(g);
	const long long as = rb_astride(g);
	sgs_rb_packed_kernel<0><<<nb, 256, 0, ctx->stream>>>(gd, g->k0, d_packed, cp2, as, d_r, d_y, d_z, ctx->gate);
	S3D_LAUNCHED(ctx);
	int rc = s3d_halo_exchange(ctx, g, d_y);
	if (rc != S3D_OK) return rc;
	sgs_rb_packed_kernel<1><<<nb, 256, 0, ctx->stream>>>(gd, g->k0, d_packed, cp2, as, d_r, d_y, d_z, ctx->gate);
	S3D_LAUNCHED(ctx);
	rc = s3d_halo_exchange(ctx, g, d_z);
	if (rc != S3D_OK) return rc;
	sgs_rb_packed_kernel<2><<<nb, 256, 0, ctx->stream>>>(gd, g->k0, d_packed, cp2, as, d_r, d_y, d_z, ctx->gate);
	S3D_LAUNCHED(ctx);
	return S3D_OK;
}

int s3d_sgs_apply(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *d_dinv, const double *d_r,
                  double *d_z, double *d_y, int ordering, int nsweeps)
{
	S3D_REQUIRE(ctx, g && A && d_dinv && d_r && d_z && d_y, "null argument");
	S3D_REQUIRE(ctx, d_r != d_z && d_r != d_y && d_z != d_y, "r, z and the scratch vector must be distinct");
	S3D_REQUIRE(ctx, nsweeps >= 1, "need at least one sweep");
	const GridDev gd = to_dev(g);
	if (ordering == S3D_SGS_REDBLACK) {
		const unsigned nb = ew_blocks(ctx, g);
		sgs_rb_kernel<0><<<nb, 256, 0, ctx->stream>>>(gd, g->k0, A, d_dinv, d_r, d_y, d_z, ctx->gate);
		S3D_LAUNCHED(ctx);
		int rc = s3d_halo_exchange(ctx, g, d_y);
		if (rc != S3D_OK) return rc;
		sgs_rb_kernel<1><<<nb, 256, 0, ctx->stream>>>(gd, g->k0, A, d_dinv, d_r, d_y, d_z, ctx->gate);
		S3D_LAUNCHED(ctx);
		rc = s3d_halo_exchange(ctx, g, d_z);
		if (rc != S3D_OK) return rc;
		sgs_rb_kernel<2><<<nb, 256, 0, ctx->stream>>>(gd, g->k0, A, d_dinv, d_r, d_y, d_z, ctx->gate);
		S3D_LAUNCHED(ctx);
		return S3D_OK;
	}
	if (ordering == S3D_SGS_JACOBI_SWEEPS) {
		// forward: ping-pong between d_y and d_z, ending in d_y; backward: ping-pong between d_z and the context's sweep buffer
		const unsigned nb = ew_blocks(ctx, g);
		if (ctx->sweep_buf_cap < (size_t)g->vlen) {
			S3D_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
			if (ctx->sweep_buf) S3D_CUDA(ctx, cudaFree(ctx->sweep_buf));
			ctx->sweep_buf = nullptr; ctx->sweep_buf_cap = 0;
			S3D_CUDA(ctx, cudaMalloc(&ctx->sweep_buf, (size_t)g->vlen * sizeof(double)));
			S3D_CUDA(ctx, cudaMemsetAsync(ctx->sweep_buf, 0, (size_t)g->vlen * sizeof(double), ctx->stream));   // ghosts stay zero
			ctx->sweep_buf_cap = (size_t)g->vlen;
			ctx->sweep_buf_dims[0] = g->nx; ctx->sweep_buf_dims[1] = g->ny; ctx->sweep_buf_dims[2] = g->nz;
		} else if (ctx->sweep_buf_dims[0] != g->nx || ctx->sweep_buf_dims[1] != g->ny || ctx->sweep_buf_dims[2] != g->nz) {
			// another grid shape: what are ghost positions now may hold old real values
			S3D_CUDA(ctx, cudaMemsetAsync(ctx->sweep_buf, 0, (size_t)g->vlen * sizeof(double), ctx->stream));
			ctx->sweep_buf_dims[0] = g->nx; ctx->sweep_buf_dims[1] = g->ny; ctx->sweep_buf_dims[2] = g->nz;
		}
		double *w = ctx->sweep_buf;
		// choose the starting buffer so that the last forward sweep lands in d_y
		double *cur = (nsweeps % 2 == 1) ? d_y : w, *oth = (nsweeps % 2 == 1) ? w : d_y;
		for (int swp = 0; swp < nsweeps; swp++) {
			if (swp > 0) { const int rc = s3d_halo_exchange(ctx, g, oth); if (rc != S3D_OK) return rc; }
			sgs_sweep_kernel<0><<<nb, 256, 0, ctx->stream>>>(gd, A, d_dinv, d_r, oth, cur, swp == 0, ctx->gate);
			S3D_LAUNCHED(ctx);
			double *t = cur; cur = oth; oth = t;
		}
		// now `oth` == d_y holds the forward result
		cur = (nsweeps % 2 == 1) ? d_z : w; oth = (nsweeps % 2 == 1) ? w : d_z;
		for (int swp = 0; swp < nsweeps; swp++) {
			if (swp > 0) { const int rc = s3d_halo_exchange(ctx, g, oth); if (rc != S3D_OK) return rc; }
			sgs_sweep_kernel<1><<<nb, 256, 0, ctx->stream>>>(gd, A, d_dinv, d_y, oth, cur, swp == 0, ctx->gate);
			S3D_LAUNCHED(ctx);
			double *t = cur; cur = oth; oth = t;
		}
		return S3D_OK;
	}
	if (ordering == S3D_SGS_ASYNC) return s3d_internal_sgs_async(ctx, g, A, d_dinv, d_r, d_z, d_y, nsweeps);
	if (ordering == S3D_SGS_BLOCKS) {
		// Blocks on ONE device (sgs_blocks_run); across z-slabs the ranks are the blocks of planes: S3D_SGS_OVERLAP, which cuts every slab further
		S3D_REQUIRE(ctx, g->nranks == 1, "S3D_SGS_BLOCKS is a one-device ordering (z-slabs: S3D_SGS_OVERLAP)");
		bool ran = false;
		const int rcb = sgs_blocks_run(ctx, g, gd, A, d_dinv, d_r, d_y, d_z, &ran);
		if (rcb != S3D_OK || ran) return rcb;
		ordering = S3D_SGS_LEXICOGRAPHIC;
	}
	if (ordering == S3D_SGS_OVERLAP && g->nranks == 1) ordering = S3D_SGS_LEXICOGRAPHIC;
	if (ordering == S3D_SGS_OVERLAP) {
		// Overlapping slabs (restricted additive Schwarz with the exact sweep as the subdomain solver): every rank sweeps its own planes
		// PLUS one plane of either neighbour -- planes -1 .. nz of its ghosted arrays, zero inflow beyond them -- and keeps its own.
		// The neighbours' planes of r arrive by the ordinary one-plane halo exchange, those of the coefficients and of dinv sit in
		// the slack planes (s3d_sgs_overlap_setup).  No cross-rank wait inside the sweeps: the cost per application is the slab-local
		// one plus the exchange of r and 2 / nz more planes, and the iteration count stays within a few per cent of the exact sweep
		// over the whole grid (slab-local: +5 .. +36 %; tools/sgs_overlap_study.py, profiles/r2_sgs_overlap.md).
		// The kernel is the warp-per-tile one on a grid descriptor that starts one plane early; it does not know.
		if (!s3d_internal_has_slack(ctx, A) || !s3d_internal_has_slack(ctx, d_dinv))
			return s3d_fail(ctx, S3D_ERR_ARG, "S3D_SGS_OVERLAP needs A from s3d_mat_alloc and dinv from s3d_coef_alloc on the z-slab grid (slack planes)");
		if (ctx->comm) {
			// (with no communicator the caller has filled the ghost planes of r and the slack planes itself: another transport, tests)
			const int rch = s3d_halo_exchange(ctx, g, const_cast<double *>(d_r));
			if (rch != S3D_OK) return rch;
		}
		const int lo = g->rank > 0 ? 1 : 0, hi = g->rank < g->nranks - 1 ? 1 : 0;
		s3d_grid gx = *g;
		gx.nz = g->nz + lo + hi;
		gx.voff = g->voff - lo * g->vplane;
		const GridDev gdx = to_dev(&gx);
		const long long cs = g->cstride;
		const double *Ax = A - lo * g->cplane, *dx = d_dinv - lo * g->cplane;
		{
			// the extended slab is cut further into concurrent blocks of planes and rows where that pays (by size, or as the tuning keys say;
			// sgs_blocks = sgs_blocks_y = 1: one sweep per slab): the chain of a slab's sweep is what bounds it up to ~256^3 points per rank
			bool ran = false;
			const int rcb = sgs_blocks_run(ctx, &gx, gdx, Ax, dx, d_r, d_y, d_z, &ran);
			if (rcb != S3D_OK || ran) return rcb;
		}
		SgsFlow f;
		unsigned nblocks = 0;
		const int rcp = sgs_flow_prepare(ctx, &gx, f, nblocks, 0);
		if (rcp != S3D_OK) return rcp;
		f.kzero_lo = lo; f.kzero_hi = hi;
		sgs_begin_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sgs_ctl, ctx->gate);
		S3D_LAUNCHED(ctx);
		sgw_launch<0>(ctx, &gx, nblocks, gdx, d_r, Ax, Ax + cs, Ax + 2 * cs, dx, d_y, f, (const int *)ctx->gate);
		S3D_LAUNCHED(ctx);
		sgs_begin_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sgs_ctl, ctx->gate);
		S3D_LAUNCHED(ctx);
		sgw_launch<1>(ctx, &gx, nblocks, gdx, (const double *)d_y, Ax + 4 * cs, Ax + 5 * cs, Ax + 6 * cs, dx, d_z, f, (const int *)ctx->gate);
		S3D_LAUNCHED(ctx);
		return S3D_OK;
	}
	if (ordering != S3D_SGS_LEXICOGRAPHIC && ordering != S3D_SGS_LEXICOGRAPHIC_PLANES && ordering != S3D_SGS_SLABLOCAL)
		return s3d_fail(ctx, S3D_ERR_ARG, "unknown SGS ordering %d", ordering);
	// slab-local: the exact lexicographic sweep inside every z-slab, nothing taken from the neighbouring slabs (block Jacobi across the
	// slab faces): no exchange, no cross-rank wait.  On one rank it IS the lexicographic sweep.
	const bool slablocal = (ordering == S3D_SGS_SLABLOCAL);
	if (slablocal) ordering = S3D_SGS_LEXICOGRAPHIC;
	const int variant = sgs_effective_variant(ctx, g, !slablocal);
	if (g->nranks > 1 && !slablocal && (ordering != S3D_SGS_LEXICOGRAPHIC || variant != 0))
		return s3d_fail(ctx, S3D_ERR_ARG, "across z-slabs only the warp-per-tile lexicographic sweep (or slab-local / red-black / sweeps) is available");
	if (g->nranks > 1 && slablocal && variant != 4 && variant != 0 && variant != 6 && variant != 7)
		return s3d_fail(ctx, S3D_ERR_ARG, "the slab-local ordering runs on the streaming, register-streaming or warp-per-tile kernel");
	// the scratch vector's real entries are all overwritten and its ghosts stay zero (they are never
	// written), so the per-call zero fill of the reference (s3d_sgspreconditioners.cpp:25-28) is implicit.
	// Sequential sweeps are idempotent, so nsweeps > 1 repeats identical work; do one.
	if (ordering == S3D_SGS_LEXICOGRAPHIC && variant == 7) {
		const long long cs = g->cstride;
		const int klo = slablocal && g->rank > 0, khi = slablocal && g->rank < g->nranks - 1;
		const int rcc = sgs_ctl_ensure(ctx);
		if (rcc != S3D_OK) return rcc;
		unsigned *ticket = reinterpret_cast<unsigned *>(ctx->sgs_ctl);
		sgs_begin_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sgs_ctl, ctx->gate);
		S3D_LAUNCHED(ctx);
		const int rcs = s3d_internal_sgs_rowscan(ctx, g, 0, d_r, A, A + cs, A + 2 * cs, d_dinv, d_y, ticket, klo, khi, ctx->gate);
		if (rcs != S3D_OK) return rcs;
		sgs_begin_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sgs_ctl, ctx->gate);
		S3D_LAUNCHED(ctx);
		return s3d_internal_sgs_rowscan(ctx, g, 1, d_y, A + 4 * cs, A + 5 * cs, A + 6 * cs, d_dinv, d_z, ticket, klo, khi, ctx->gate);
	}
	if (ordering == S3D_SGS_LEXICOGRAPHIC && variant == 6) {
		const long long cs = g->cstride;
		const int klo = slablocal && g->rank > 0, khi = slablocal && g->rank < g->nranks - 1;
		const int rcc = sgs_ctl_ensure(ctx);
		if (rcc != S3D_OK) return rcc;
		unsigned *ticket = reinterpret_cast<unsigned *>(ctx->sgs_ctl);
		sgs_begin_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sgs_ctl, ctx->gate);
		S3D_LAUNCHED(ctx);
		const int rcs = s3d_internal_sgs_pencil(ctx, g, 0, d_r, A, A + cs, A + 2 * cs, d_dinv, d_y, ticket, klo, khi, ctx->gate);
		if (rcs != S3D_OK) return rcs;
		sgs_begin_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sgs_ctl, ctx->gate);
		S3D_LAUNCHED(ctx);
		return s3d_internal_sgs_pencil(ctx, g, 1, d_y, A + 4 * cs, A + 5 * cs, A + 6 * cs, d_dinv, d_z, ticket, klo, khi, ctx->gate);
	}
	if (ordering == S3D_SGS_LEXICOGRAPHIC) {
		SgsFlow f;
		unsigned nblocks = 0;
		const int rcp = sgs_flow_prepare(ctx, g, f, nblocks, variant);
		if (rcp != S3D_OK) return rcp;
		dim3 block(SGS_NT);
		const long long cs = g->cstride;
		if (variant == 4) {
			const int klo = slablocal && g->rank > 0, khi = slablocal && g->rank < g->nranks - 1;
			sgs_begin_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sgs_ctl, ctx->gate);
			S3D_LAUNCHED(ctx);
			int rcs = s3d_internal_sgs_stream(ctx, g, 0, d_r, A, A + cs, A + 2 * cs, d_dinv, d_y, f.order, f.nty, f.ntz, f.ticket, klo, khi, ctx->gate);
			if (rcs != S3D_OK) return rcs;
			sgs_begin_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sgs_ctl, ctx->gate);
			S3D_LAUNCHED(ctx);
			return s3d_internal_sgs_stream(ctx, g, 1, d_y, A + 4 * cs, A + 5 * cs, A + 6 * cs, d_dinv, d_z, f.order, f.nty, f.ntz, f.ticket, klo, khi, ctx->gate);
		}
		if (slablocal) { f.kzero_lo = g->rank > 0; f.kzero_hi = g->rank < g->nranks - 1; }
		s3d_sgs_window xw;
		if (g->nranks > 1 && !slablocal) {
			f.mail = nullptr;   // the cross-rank hand-off is the flag protocol's
			// the sweep runs THROUGH the slabs: rank p's first plane of tiles waits for rank p-1's last one (peer-memory window)
			const int rcw = s3d_internal_sgs_window(ctx, g, f.ntx, f.nty, &xw);
			if (rcw != S3D_OK) return rcw;
			if (!xw.ok) return s3d_fail(ctx, S3D_ERR_STATE, "lexicographic SGS across z-slabs needs the peer-memory windows (CUDA IPC); use S3D_SGS_REDBLACK");
			f.x_in_plane = xw.fwd_in_plane; f.x_in_flags = xw.fwd_in_flags; f.x_out_plane = xw.fwd_out_plane; f.x_out_flags = xw.fwd_out_flags;
		}
		sgs_begin_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sgs_ctl, ctx->gate);
		S3D_LAUNCHED(ctx);
		if (variant == 3)
			sgs_lineflow_kernel<0, 1><<<nblocks, 32, sgl_smem(5, 1), ctx->stream>>>(gd, d_r, A, A + cs, A + 2 * cs, d_dinv, d_y, f, ctx->gate);
		else if (variant == 0)
			sgw_launch<0>(ctx, g, nblocks, gd, d_r, A, A + cs, A + 2 * cs, d_dinv, d_y, f, (const int *)ctx->gate);
		else
			sgs_dataflow_kernel<0><<<nblocks, block, SGS_SMEM, ctx->stream>>>(gd, d_r, A, A + cs, A + 2 * cs, d_dinv, d_y, f, ctx->gate);
		S3D_LAUNCHED(ctx);
		if (g->nranks > 1 && !slablocal) { f.x_in_plane = xw.bwd_in_plane; f.x_in_flags = xw.bwd_in_flags; f.x_out_plane = xw.bwd_out_plane; f.x_out_flags = xw.bwd_out_flags; }
		sgs_begin_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sgs_ctl, ctx->gate);
		S3D_LAUNCHED(ctx);
		if (variant == 3)
			sgs_lineflow_kernel<1, 1><<<nblocks, 32, sgl_smem(5, 1), ctx->stream>>>(gd, d_y, A + 4 * cs, A + 5 * cs, A + 6 * cs, d_dinv, d_z, f, ctx->gate);
		else if (variant == 0)
			sgw_launch<1>(ctx, g, nblocks, gd, (const double *)d_y, A + 4 * cs, A + 5 * cs, A + 6 * cs, d_dinv, d_z, f, (const int *)ctx->gate);
		else
			sgs_dataflow_kernel<1><<<nblocks, block, SGS_SMEM, ctx->stream>>>(gd, d_y, A + 4 * cs, A + 5 * cs, A + 6 * cs, d_dinv, d_z, f, ctx->gate);
		S3D_LAUNCHED(ctx);
		return S3D_OK;
	}
	// S3D_SGS_LEXICOGRAPHIC_PLANES: one launch per global hyperplane (the first correct version)
	dim3 block(32, 8);
	dim3 grid((g->ny + 31) / 32, (g->nz + 7) / 8);
	const int hmax = g->nx + g->ny + g->nz - 3;
	for (int h = 0; h <= hmax; h++) {
		sgs_fwd_plane_kernel<<<grid, block, 0, ctx->stream>>>(gd, A, d_dinv, d_r, d_y, h, ctx->gate);
		ctx->launches++;
	}
	for (int h = hmax; h >= 0; h--) {
		sgs_bwd_plane_kernel<<<grid, block, 0, ctx->stream>>>(gd, A, d_dinv, d_y, d_z, h, ctx->gate);
		ctx->launches++;
	}
	S3D_CUDA(ctx, cudaGetLastError());
	return S3D_OK;
}

// StrILU_preconditioner::updateOperator (src/linalg/s3d_ilu.cpp:17-64), sequential build: the ILU(0) diagonal of the 7-point
// matrix, d_ijk = A3 - A0*A4[i-1]/d[i-1] - A1*A5[j-1]/d[j-1] - A2*A6[k-1]/d[k-1], then dinv = 1/d.  A lexicographic sweep computes
// the exact fixed point, so further build sweeps reproduce it (they are skipped); 0 sweeps gives dinv = 1/diag like the reference.
// The result feeds s3d_sgs_apply (same (D+L) D^-1 (D+U) application, SGS_like_preconditioner::apply).
static int strilu_setup_impl(s3d_ctx *ctx, const s3d_grid *g, const double *A, int nbuildsweeps, double *d_dinv, bool async);

int s3d_strilu_setup(s3d_ctx *ctx, const s3d_grid *g, const double *A, int nbuildsweeps, double *d_dinv)
{
	return strilu_setup_impl(ctx, g, A, nbuildsweeps, d_dinv, false);
}

// The reference's THREADED build (threadedbuild = true, s3d_ilu.cpp:28-32): nbuildsweeps asynchronous sweeps from d = diag(A), rows dealt to
// warps in chunks (tuning key sgs_async_chunk = -s3d_thread_chunk_size) that never wait for one another.  Timing-dependent like the
// reference's; it reaches the exact diagonal after at most nx + ny + nz sweeps.
int s3d_strilu_setup_async(s3d_ctx *ctx, const s3d_grid *g, const double *A, int nbuildsweeps, double *d_dinv)
{
	return strilu_setup_impl(ctx, g, A, nbuildsweeps, d_dinv, true);
}

static int strilu_setup_impl(s3d_ctx *ctx, const s3d_grid *g, const double *A, int nbuildsweeps, double *d_dinv, bool async)
{
	S3D_REQUIRE(ctx, g && A && d_dinv && nbuildsweeps >= 0, "bad argument");
	// On a z-slab decomposition the recurrence is run inside every slab (nothing taken across the slab faces: the products below are
	// zero where the k-1 neighbour lives on another rank), the factorisation that goes with the slab-local application.
	if (nbuildsweeps == 0) return s3d_sgs_setup(ctx, g, A, d_dinv);
	const GridDev gd = to_dev(g);
	double *P = nullptr, *D = nullptr;
	S3D_CUDA(ctx, cudaMalloc(&P, 3 * (size_t)g->cstride * sizeof(double)));
	cudaError_t e = cudaMalloc(&D, (size_t)g->vlen * sizeof(double));
	if (e != cudaSuccess) { cudaFree(P); S3D_CUDA(ctx, e); }
	const unsigned nb = (unsigned)std::min<long long>(((long long)g->nx * g->ny * g->nz + 255) / 256, (long long)ctx->sm_count * 16);
	ilu_products_kernel<<<nb, 256, 0, ctx->stream>>>(gd, A, P);
	ctx->launches++;
	fill_all_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(D, g->vlen, 1.0);   // ghosts = 1 so that 0/d stays 0 at the boundary
	ctx->launches++;
	SgsFlow f;
	unsigned nblocks = 0;
	int variant = sgs_effective_variant(ctx, g, false);
	if (variant == 7) variant = 0;   // the ILU recurrence is not linear in d[i-1]: no row scan; the bit-exact tile kernel builds it
	if (async) {
		int rca = s3d_internal_strilu_async(ctx, g, A + 3 * (size_t)g->cstride, P, D, nbuildsweeps);
		if (rca == S3D_OK) {
			ilu_invert_kernel<<<nb, 256, 0, ctx->stream>>>(gd, D, d_dinv);
			ctx->launches++;
			e = cudaGetLastError();
			if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
			if (e != cudaSuccess) rca = s3d_fail(ctx, S3D_ERR_CUDA, "s3d_strilu_setup_async: %s", cudaGetErrorString(e));
		}
		cudaFree(P);
		cudaFree(D);
		return rca;
	}
	int rc = (variant == 6) ? sgs_ctl_ensure(ctx) : sgs_flow_prepare(ctx, g, f, nblocks, variant);
	if (rc == S3D_OK) {
		sgs_begin_kernel<<<1, 1, 0, ctx->stream>>>(ctx->sgs_ctl, nullptr);
		ctx->launches++;
		e = cudaGetLastError();
		if (e == cudaSuccess) {
			if (variant == 4) {
				rc = s3d_internal_sgs_stream(ctx, g, 2, A + 3 * (size_t)g->cstride, P, P + g->cstride, P + 2 * (size_t)g->cstride, nullptr, D,
				                             f.order, f.nty, f.ntz, f.ticket, 0, 0, nullptr);
				ctx->launches--;   // counted below
			} else if (variant == 6) {
				rc = s3d_internal_sgs_pencil(ctx, g, 2, A + 3 * (size_t)g->cstride, P, P + g->cstride, P + 2 * (size_t)g->cstride, nullptr, D,
				                             reinterpret_cast<unsigned *>(ctx->sgs_ctl), 0, 0, nullptr);
				ctx->launches--;   // counted below
			} else if (variant == 3)
				sgs_lineflow_kernel<2, 1><<<nblocks, 32, sgl_smem(4, 1), ctx->stream>>>(gd, A + 3 * (size_t)g->cstride, P, P + g->cstride, P + 2 * (size_t)g->cstride,
				                                                                         nullptr, D, f, nullptr);
			else if (variant == 0)
				sgw_launch<2>(ctx, g, nblocks, gd, A + 3 * (size_t)g->cstride, (const double *)P, (const double *)(P + g->cstride),
				              (const double *)(P + 2 * (size_t)g->cstride), (const double *)nullptr, D, f, (const int *)nullptr);
			else
				sgs_dataflow_kernel<2><<<nblocks, SGS_NT, SGS_SMEM, ctx->stream>>>(gd, A + 3 * (size_t)g->cstride, P, P + g->cstride, P + 2 * (size_t)g->cstride,
				                                                                      nullptr, D, f, nullptr);
			ctx->launches++;
			ilu_invert_kernel<<<nb, 256, 0, ctx->stream>>>(gd, D, d_dinv);
			ctx->launches++;
			e = cudaGetLastError();
		}
		if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
		if (e != cudaSuccess) rc = s3d_fail(ctx, S3D_ERR_CUDA, "s3d_strilu_setup: %s", cudaGetErrorString(e));
	}
	cudaFree(P);
	cudaFree(D);
	return rc;
}

}  // extern "C"
