// Internal header of libs3d_cuda.so (sm_100a only).  Nothing here is part of the ABI.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <unordered_map>
#include <nccl.h>

#include "s3d_cuda.h"

// what a z-slab SpMV reads instead of x's two ghost planes: the slots of this rank's peer-memory window (context.cu)
// cross-file helpers of the library that are not part of the C-ABI (include/s3d_cuda.h)
#define S3D_HIDDEN __attribute__((visibility("hidden")))

struct s3d_halo_view {
	const double *lo = nullptr, *hi = nullptr;         // plane-shaped slots (null at the ends of the global grid)
	const int *flag_lo = nullptr, *flag_hi = nullptr;  // hold `epoch` once the neighbour's plane has landed
	int epoch = 0;
};

// cross-rank (z-slab) extension of the exact SGS sweeps: where a slab's first plane of tiles finds the k-face rows of the rank before
// it in sweep order, and where its last plane of tiles puts its own (context.cu: s3d_internal_sgs_window)
struct s3d_sgs_window {
	int ok = 0;
	const double *fwd_in_plane = nullptr, *bwd_in_plane = nullptr;
	const int *fwd_in_flags = nullptr, *bwd_in_flags = nullptr;
	double *fwd_out_plane = nullptr, *bwd_out_plane = nullptr;
	int *fwd_out_flags = nullptr, *bwd_out_flags = nullptr;
};

struct s3d_ctx {
	int device = 0;
	int sm_count = 148;
	cudaStream_t stream = nullptr;
	bool own_stream = false;
	// two-stage deterministic reductions: per-block partial sums + a ticket that elects the last block
	double *partials = nullptr;
	size_t partials_cap = 0;      // doubles
	unsigned *ticket = nullptr;
	// small device scratch: 1-D tables for assembly / weighted norms, pointer tables for multi-vector ops
	double *tab = nullptr;
	size_t tab_cap = 0;           // doubles
	double *red_tmp = nullptr;    // 64 device doubles for the *_host conveniences
	double *h_pinned = nullptr;   // 64 pinned doubles
	// SGS dataflow bookkeeping (tile completion flags + work ticket)
	int *sgs_flags = nullptr;
	size_t sgs_flags_cap = 0;
	double *sgs_iface = nullptr;  // per tile: the 64 values of its last i-column (what the next tile along i needs)
	size_t sgs_iface_cap = 0;
	double *sweep_buf = nullptr;   // second ping-pong vector of the Jacobi-sweeps SGS
	size_t sweep_buf_cap = 0;
	int sweep_buf_dims[3] = {0, 0, 0};
	int *sgs_ctl = nullptr;       // device: [0] ticket counter, [1] epoch of the running dataflow sweep
	bool capturing = false;       // between s3d_graph_begin and s3d_graph_end
	int64_t capture_launches0 = 0;
	int sgs_sleep_ns = 0;
	int sgs_variant = 4;        // 4: by size (streaming kernel below sgs_stream_max_points, warp-per-tile above; the default), 5: streaming, 6: register-streaming
	                            // (sgs_pencil.cu, opt-in: measured slower, profiles/r2_sgs_pencil.md), 7: row-scan wavefront (sgs_rowscan.cu, opt-in: the
	                            // reference's sweep to rounding, measured slower, profiles/r2_sgs_rowscan.md), 0: warp-per-tile dataflow
	                            // kernel (also what the exact sweep THROUGH z-slabs uses), 1: block-per-tile kernel, 3: r1's warp-per-line experiment
	long long sgs_stream_max_points = 0;   // variant 4 uses the streaming kernel up to this many points per rank, the tile kernel above.  0 since the
	                                       // tile kernel hands its faces over through mailboxes: it is the faster one at every size (profiles/r2_sgs_tile_mail.md)
	unsigned long long *sgs_mail = nullptr;   // streaming kernel: face mailboxes (sentinel-filled outside a sweep)
	size_t sgs_mail_cap = 0;
	int sgs_warps_per_sm = 0;   // 0: as many as fit
	int small_team = 0;         // single-launch solver: 0 = cluster up to 32 Ki points, cooperative grid above; 1 = always the cluster (A/B)
	int sgs_tile_mail = 1;      // warp-per-tile kernel: hand the faces over through sentinel mailboxes (no fence, no flag) instead of `out` + epoch flags
	int sgs_async_chunk = 4;    // rows per chunk of the asynchronous sweeps (S3D_SGS_ASYNC); the host passes -s3d_thread_chunk_size
	int sgs_pencil_shape = 0;   // register-streaming kernel (variant 6): warps of a block as WJ*10 + WK (+ prefetch depth as a third digit); 0 = default
	long long *sgs_prof = nullptr;   // timing experiments: per-phase cycle sums of the warp kernel (tuning key sgs_profile)
	int sgs_tile_i = 0;         // warp kernel: tile length along i, 16 (default) or 32
	int sgs_pencils = 0;        // warp kernel: (j,k) pencils per lane, 1 (tile TI x 4 x 8) or 2 (tile TI x 8 x 8); 0: by grid size
	int sgs_dbuf = 0;           // warp kernel: second staging area (1 on, 0 off, -1: from sgs_dbuf_min_points points on)
	long long sgs_dbuf_min_points = 1ll << 62;
	int sgs_blocks = 0;         // S3D_SGS_BLOCKS: blocks of planes the grid is cut into on one device (tuning key sgs_blocks, host option -s3d_sgs_blocks; 0: by size)
	int sgs_blocks_y = 0;       // ... and blocks of rows (tuning key sgs_blocks_y, host option -s3d_sgs_blocks_y; 0: by size)
	struct SgsBlocks { int key[7] = {0, 0, 0, 0, 0, 0, 0}; int *order[2] = {nullptr, nullptr}; int4 *layers[2] = {nullptr, nullptr}; int ntz[2] = {0, 0}, nty[2] = {0, 0}; } sgsblocks;
	int *sgs_order = nullptr;     // tile dispatch order (hyperplane-sorted), cached per tile-grid shape
	int sgs_order_dims[3] = {0, 0, 0};
	// device-side gate applied to every launch that does not bring its own stop flag
	const int *gate = nullptr;
	// timing
	cudaEvent_t ev0 = nullptr, ev1 = nullptr;
	// pipelined host-buffer SpMV: copy streams and per-chunk events
	cudaStream_t up_stream = nullptr, down_stream = nullptr;
	// split SpMV launches that share one reduction (overlapped halo exchange): partials are appended, summed once at the end
	bool no_overlap = true;        // measured on 2 and 8 B200s: overlapping the NCCL halo exchange with the interior SpMV gains nothing (r1), so it is opt-in
	bool red_defer = false;
	size_t red_blocks = 0;
	cudaStream_t comm_stream = nullptr;   // NCCL halo traffic, concurrent with the interior SpMV
	// peer-memory halo: the push of large planes runs on a high-priority side stream, beside the multiply that follows it on the
	// compute stream (tuning key halo_push_side, default 1); the compute stream re-joins it before anything may touch x again
	cudaStream_t push_stream = nullptr;
	cudaEvent_t push_ev[2] = {};
	bool push_pending = false;
	int push_side = 1;
	cudaEvent_t comm_ev[2] = {};
	double *stage = nullptr;       // device staging for contiguous host uploads (repacked into the padded layout on the device)
	size_t stage_cap = 0;
	double *stage_y = nullptr;     // device staging for contiguous host downloads (results packed into the host layout on the device)
	size_t stage_y_cap = 0;
	int *h_slots = nullptr;        // 8 pinned slots of 8 ints for pipelined flag polling
	cudaEvent_t slot_ev[8] = {};
	cudaEvent_t *chunk_ev = nullptr;
	int chunk_ev_cap = 0;
	struct Phase { cudaEvent_t *ev = nullptr; int used = 0, cap = 0; };
	Phase phase[4];
	int64_t launches = 0;
	// coefficient storage handed out with slack planes (z-slab grids, see s3d_mat_alloc): returned pointer -> start of the allocation
	std::unordered_map<const void *, void *> padded;
	// z-slab communicator
	ncclComm_t comm = nullptr;
	// peer-memory halo window (one process per GPU, neighbours' windows mapped through CUDA IPC); see context.cu
	struct P2P {
		int mode = 1;                 // tuning key halo_p2p: 1 = use peer memory when it can be set up, 0 = NCCL send/recv
		bool ok = false;              // windows of the neighbours are mapped on every rank
		double *win = nullptr;        // [32 doubles of header: 4 int flags][4 planes: (from below, from above) x parity]
		size_t plane_cap = 0;         // doubles per plane slot
		size_t tried_cap = 0;         // plane size the window set-up was last attempted for (the attempt is collective: never repeat it on one rank only)
		double *peer_lo = nullptr, *peer_hi = nullptr;   // the windows of rank-1 / rank+1 as mapped here
		unsigned *ticket = nullptr;   // last-block election of the push kernel
		int epoch = 0;
	} p2p;
	struct SgsWin {
		bool tried = false, ok = false;
		double *win = nullptr, *peer_lo = nullptr, *peer_hi = nullptr;
		int ntx = 0, nty = 0;
		size_t plane = 0;
	} sgswin;
	// one-shot all-reduce of a few scalars through peer memory (context.cu): every rank's window is mapped by all the others
	struct ArWin {
		int mode = 1;                 // tuning key allreduce_p2p: 1 = peer memory when it can be set up, 0 = ncclAllReduce
		bool tried = false, ok = false;
		double *win = nullptr;        // [2 parities][nranks][S3D_AR_STRIDE doubles: 32 values + the epoch]
		double *peer[16] = {};
		int epoch = 0;
	} arwin;
	int rank = 0, nranks = 1;
	char err[512] = "";
};

int s3d_fail(s3d_ctx *ctx, int code, const char *fmt, ...);
int s3d_ensure_partials(s3d_ctx *ctx, size_t ndoubles);
int s3d_ensure_tab(s3d_ctx *ctx, size_t ndoubles);

#define S3D_CUDA(ctx, call)                                                                     \
	do {                                                                                        \
		cudaError_t e__ = (call);                                                               \
		if (e__ != cudaSuccess)                                                                 \
			return s3d_fail((ctx), S3D_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,   \
			                cudaGetErrorString(e__));                                           \
	} while (0)

#define S3D_NCCL(ctx, call)                                                                     \
	do {                                                                                        \
		ncclResult_t r__ = (call);                                                              \
		if (r__ != ncclSuccess)                                                                 \
			return s3d_fail((ctx), S3D_ERR_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #call,   \
			                ncclGetErrorString(r__));                                           \
	} while (0)

#define S3D_LAUNCHED(ctx)                                                                       \
	do {                                                                                        \
		(ctx)->launches++;                                                                      \
		S3D_CUDA((ctx), cudaGetLastError());                                                    \
	} while (0)

#define S3D_REQUIRE(ctx, cond, msg)                                                             \
	do {                                                                                        \
		if (!(cond)) return s3d_fail((ctx), S3D_ERR_ARG, "%s: %s", __func__, (msg));            \
	} while (0)

// ---------------------------------------------------------------------------------------------
// device-side view of the grid (passed by value to kernels)
struct GridDev {
	int nx, ny, nz;
	int vpitch, cpitch;
	long long vplane, voff, cplane, cstride;
};

static inline GridDev to_dev(const s3d_grid *g)
{
	GridDev d;
	d.nx = g->nx; d.ny = g->ny; d.nz = g->nz;
	d.vpitch = g->vpitch; d.cpitch = g->cpitch;
	d.vplane = g->vplane; d.voff = g->voff; d.cplane = g->cplane; d.cstride = g->cstride;
	return d;
}

struct RedOut {
	double *partials;
	unsigned *ticket;   // null: write the per-block partials only; a separate final_reduce launch adds them up
	double *out;        // device scalars receiving the final sums (may be null: no reduction requested)
};

// second stage as its own launch, for kernels with very many blocks (a per-block ticket atomic on one
// address serialises in L2: 262k blocks cost ~1 ms).  Same fixed summation order: deterministic.
int s3d_final_reduce(s3d_ctx *ctx, int nr, unsigned nblocks, double *out);

#ifdef __CUDACC__

__device__ __forceinline__ double s3d_sval(const s3d_scalar &s)
{
	double v = s.imm;
	if (s.num) v *= __ldg(s.num);
	if (s.den) v /= __ldg(s.den);
	return v;
}

__device__ __forceinline__ bool s3d_stopped(const int *stop) { return stop != nullptr && __ldg(stop) != 0; }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}

// Sums NR per-thread accumulators over the whole grid into ro.out[0..NR).
// Stage 1: warp shuffle + shared memory inside the block, one partial per block.
// Stage 2: the block that draws the last ticket adds the partials in a fixed order.
// The result depends only on the launch geometry, never on scheduling: deterministic.
// Every thread of every block must call this (it contains __syncthreads).
template <int NR, bool PARTIALS_ONLY = false>
__device__ __forceinline__ void grid_reduce(double (&acc)[NR], const RedOut &ro)
{
	__shared__ double sm[NR][32];
	__shared__ bool is_last;
	const int tid = threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z);
	const int nthreads = blockDim.x * blockDim.y * blockDim.z;
	const int lane = tid & 31, warp = tid >> 5, nwarps = (nthreads + 31) >> 5;
	const unsigned nblocks = gridDim.x * gridDim.y * gridDim.z;
	const unsigned bid = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);

#pragma unroll
	for (int r = 0; r < NR; r++) {
		const double w = warp_sum(acc[r]);
		if (lane == 0) sm[r][warp] = w;
	}
	__syncthreads();
	if (warp == 0) {
#pragma unroll
		for (int r = 0; r < NR; r++) {
			double v = lane < nwarps ? sm[r][lane] : 0.0;
			v = warp_sum(v);
			if (lane == 0) ro.partials[(size_t)bid * NR + r] = v;
		}
	}
	if (PARTIALS_ONLY || ro.ticket == nullptr) return;   // stage 2 is a separate launch (s3d_final_reduce)
	if (tid == 0) {
		__threadfence();
		const unsigned t = atomicAdd(ro.ticket, 1u);
		is_last = (t == nblocks - 1);
	}
	__syncthreads();
	if (!is_last) return;
	__threadfence();
#pragma unroll
	for (int r = 0; r < NR; r++) {
		double s = 0.0;
		for (unsigned b = tid; b < nblocks; b += nthreads) s += __ldcg(&ro.partials[(size_t)b * NR + r]);
		s = warp_sum(s);
		__syncthreads();
		if (lane == 0) sm[r][warp] = s;
		__syncthreads();
		if (warp == 0) {
			double v = lane < nwarps ? sm[r][lane] : 0.0;
			v = warp_sum(v);
			if (lane == 0) ro.out[r] = v;
		}
	}
	if (tid == 0) *ro.ticket = 0u;
}

// streaming (read-once / write-once) accesses: evict-first in L1 and L2 so the coefficient streams
// do not push the x-planes out of cache
__device__ __forceinline__ double2 ld_stream2(const double *p) { return __ldcs(reinterpret_cast<const double2 *>(p)); }
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream2(double *p, double2 v) { __stcs(reinterpret_cast<double2 *>(p), v); }
__device__ __forceinline__ double2 ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
__device__ __forceinline__ void st2(double *p, double2 v) { *reinterpret_cast<double2 *>(p) = v; }

#endif  // __CUDACC__
