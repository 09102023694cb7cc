/* s3d_cuda.h -- C ABI of libs3d_cuda.so, the B200 (sm_100a) implementation of Struct3d's native
 * structured-grid linear-algebra path.
 *
 * This is the only boundary between host code and the GPU.  The C++ host classes in
 * struct3d_b200/host/ (SVec, SMat, SolverBase, PDEBase ... -- the reference's own interface,
 * /root/reference/src/linalg/matvec.hpp:23-115, s3d_solverbase.hpp:34-54, pde/pdebase.hpp:14-62)
 * call nothing but these functions; so do the Python tests and bench.py (through ctypes).
 * Plain pointers and sizes only: no C++ types, no torch types.
 *
 * Conventions
 *   - every function returns an int status: S3D_OK (0) or an S3D_ERR_* code; s3d_error_string()
 *     gives the message of the last failure on that context.  There is NO CPU fallback: without a
 *     CUDA device s3d_ctx_create fails and nothing else can be called.
 *   - `double *` arguments named d_* / A / x / y ... are DEVICE pointers obtained from
 *     s3d_vec_alloc / s3d_mat_alloc / s3d_scalars_alloc (or any 256-byte aligned device memory of
 *     the documented length, e.g. a torch CUDA tensor's data_ptr()).  Arguments named h_* are HOST
 *     pointers in the REFERENCE's host layouts (see below), so reference data moves over unchanged.
 *   - all work is enqueued on the context's stream and is asynchronous; only functions that
 *     return numbers to the host (`*_host`, downloads, s3d_ctx_sync) block.
 *   - reductions are deterministic: fixed grid, fixed-order two-stage tree, no floating-point atomics.
 *
 * Device layout (DESIGN.md "Data layout in HBM")
 *   vector : ghost-padded like the reference's SVec (one ghost layer per side, real points 1..n),
 *            but each i-row is padded so that real point i=0 sits on a 128-byte boundary:
 *            row = [vlead-1 pad][ghost][nx real][ghost][pad], vpitch = roundup(vlead+nx+1, 16) doubles;
 *            plane = (ny+2) rows; (nz+2) planes.  Element (i,j,k) of the real grid (0-based) lives at
 *            voff + k*vplane + j*vpitch + i.  Ghost entries are stored and read exactly like the
 *            reference reads them (they impose the homogeneous Dirichlet condition, matvec.cpp:59-101)
 *            and are never written by any operation except uploads, copies-with-ghosts and halo exchange.
 *   matrix : 7 coefficient arrays, stencil-major {i-1, j-1, k-1, diag, i+1, j+1, k+1} like SMat::vals,
 *            no ghosts, cpitch = roundup(nx, 16); element (i,j,k) of array s at
 *            s*cstride + k*cplane + j*cpitch + i.  Every row starts on a 128-byte boundary.
 * Host layouts (the reference's, src/cartmesh.hpp:141-154)
 *   vector : (nx+2)(ny+2)(nz+2) doubles, index k*(ny+2)(nx+2) + j*(nx+2) + i, ghosts included;
 *   matrix : per stencil array nx*ny*nz doubles, index k*ny*nx + j*nx + i.
 */
#ifndef S3D_CUDA_H
#define S3D_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define S3D_OK            0
#define S3D_ERR_CUDA      1   /* a CUDA runtime call or kernel failed */
#define S3D_ERR_ARG       2   /* bad argument (null pointer, size mismatch, unsupported option) */
#define S3D_ERR_NCCL      3   /* an NCCL call failed */
#define S3D_ERR_NODEVICE  4   /* no usable CUDA device: there is no CPU fallback */
#define S3D_ERR_STATE     5   /* call not valid in the current state (e.g. halo exchange without a communicator) */

#define S3D_NSTENCIL      7
#define S3D_STENCIL_DIAG  3

typedef struct s3d_ctx s3d_ctx;

/* Local (per-rank z-slab) grid descriptor.  Fill it with s3d_grid_init(); read-only afterwards. */
typedef struct s3d_grid {
	int32_t nx, ny, nz;        /* real points owned by this rank (nz = local slab thickness) */
	int32_t vlead;             /* doubles in front of real i=0 in a vector row (16: 128-byte aligned) */
	int32_t vpitch;            /* vector row pitch in doubles */
	int32_t cpitch;            /* coefficient row pitch in doubles */
	int64_t vplane;            /* vpitch * (ny+2) */
	int64_t voff;              /* offset of real point (0,0,0) = vplane + vpitch + vlead */
	int64_t vlen;              /* doubles per vector = vplane * (nz+2) */
	int64_t cplane;            /* cpitch * ny */
	int64_t cstride;           /* doubles between consecutive stencil arrays */
	int32_t nz_global;         /* real points along z of the whole grid */
	int32_t k0;                /* global index of this rank's first real plane */
	int32_t rank, nranks;      /* position in the z-slab decomposition (0,1 when not decomposed) */
} s3d_grid;

/* A scalar operand for the solver kernels: value = imm * (*num if num) / (*den if den).
 * num/den are device pointers (typically reduction outputs), so Krylov coefficients such as
 * alpha = (r,z)/(p,Ap) never round-trip through the host. */
typedef struct s3d_scalar {
	double        imm;
	const double *num;
	const double *den;
} s3d_scalar;

/* ------------------------------------------------------------------ context */

/* device < 0 selects $LOCAL_RANK (or 0).  stream: an existing cudaStream_t to enqueue on (e.g.
 * torch's current stream) or NULL to create a private non-blocking stream. */
int  s3d_ctx_create(int device, void *stream, s3d_ctx **out);
int  s3d_ctx_destroy(s3d_ctx *ctx);
int  s3d_ctx_sync(s3d_ctx *ctx);
int  s3d_ctx_device(const s3d_ctx *ctx);
void *s3d_ctx_stream(const s3d_ctx *ctx);
const char *s3d_error_string(const s3d_ctx *ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
int64_t s3d_ctx_launch_count(const s3d_ctx *ctx);
/* CUDA-event timing on the context's stream: start/stop bracket, elapsed in milliseconds */
int  s3d_timer_start(s3d_ctx *ctx);
int  s3d_timer_stop_ms(s3d_ctx *ctx, double *ms);

/* kernel tuning knobs for sweeps: "spmv_kchunk" (planes per march, 0 = automatic), "spmv_minb" (resident blocks per
 * SM the march kernel is compiled for: 0 auto, 3, 4, 6), "spmv_hints" (evict-first hints on/off for the PAIR/NAIVE
 * variants), "spmv_variant" (what S3D_SPMV_AUTO resolves to), "halo_overlap" (0 = exchange first, then multiply), "halo_p2p" (1 = halo
 * planes through peer-memory windows, 0 = NCCL), "ew_unroll" (items a thread of the element-wise kernels keeps in flight: 0 = each
 * kernel's own choice, 2/4/8, -1 = grid-stride order), "sgs_variant" (4 = by grid size (default): the streaming warp-per-line kernel up to "sgs_stream_max_points" points per rank, the warp-per-tile dataflow kernel above;
 * 5 = streaming kernel, 0 = warp-per-tile, 1 = block-per-tile, 3 = r1's warp-per-line experiment),
 * "sgs_tile_i" (16/32), "sgs_pencils" (1/2/0 = by grid size), "sgs_warps_per_sm", "sgs_profile" (timing experiments),
 * "sgs_blocks" / "sgs_blocks_y" (blocks of planes / rows of S3D_SGS_BLOCKS and inside the slabs of S3D_SGS_OVERLAP: 0 = by size, 1 = none), "sgs_tile_mail"
 * (1 = tile faces through sentinel mailboxes, 0 = flags and fences), "sgs_dbuf" (1 = a second staging area per warp: measured slower, off). */
int  s3d_tune_set(s3d_ctx *ctx, const char *key, int value);

/* Replaces the size arithmetic of SVec::SVec / SMat::SMat (matvec.cpp:8-57).  nz is the local slab. */
int  s3d_grid_init(s3d_grid *g, int nx, int ny, int nz);
/* z-slab decomposition of a global grid: rank p of P owns planes [p*nz/P, (p+1)*nz/P). */
int  s3d_grid_init_slab(s3d_grid *g, int nx, int ny, int nz_global, int rank, int nranks);

/* ------------------------------------------------------------------ storage (SVec / SMat, matvec.hpp:23-89) */

int  s3d_vec_alloc(s3d_ctx *ctx, const s3d_grid *g, double **d_out);          /* zero-filled, ghosts included */
int  s3d_mat_alloc(s3d_ctx *ctx, const s3d_grid *g, double **d_out);          /* 7*cstride doubles, zero-filled (z-slabs: with slack planes, see s3d_coef_alloc) */
int  s3d_scalars_alloc(s3d_ctx *ctx, int n, double **d_out);                  /* n device doubles, zeroed */
/* ONE array in the coefficient layout (cstride doubles, zero-filled), e.g. the dinv of the SGS / StrILU preconditioners.  On a z-slab grid
 * (s3d_grid_init_slab, nranks > 1) it, like every array of s3d_mat_alloc, carries two SLACK planes -- plane -1 and plane nz -- that hold the
 * neighbouring slabs' boundary planes for S3D_SGS_OVERLAP (cstride covers nz+2 planes there, and the allocation starts one plane before
 * the pointer handed out; release with s3d_free).  Nothing else reads them. */
int  s3d_coef_alloc(s3d_ctx *ctx, const s3d_grid *g, double **d_out);
/* slack planes by hand (another transport than NCCL, tests): host plane in the reference layout (nx*ny doubles) -> plane -1 (side 0) or
 * plane nz (side 1) of array `stencil` of d_C; s3d_coef_halo_exchange does it between the ranks of the communicator (collective). */
int  s3d_coef_slack_upload(s3d_ctx *ctx, const s3d_grid *g, double *d_C, int stencil, int side, const double *h_plane);
int  s3d_coef_halo_exchange(s3d_ctx *ctx, const s3d_grid *g, double *d_C, int narrays);
int  s3d_free(s3d_ctx *ctx, void *d_ptr);
int  s3d_vec_upload(s3d_ctx *ctx, const s3d_grid *g, double *d_x, const double *h_x);     /* host ghosted layout -> device */
int  s3d_vec_download(s3d_ctx *ctx, const s3d_grid *g, const double *d_x, double *h_x);   /* blocks */
int  s3d_mat_upload(s3d_ctx *ctx, const s3d_grid *g, double *d_A, int stencil, const double *h_vals);
int  s3d_mat_download(s3d_ctx *ctx, const s3d_grid *g, const double *d_A, int stencil, double *h_vals);
int  s3d_scalars_download(s3d_ctx *ctx, const double *d_s, int n, double *h_out);          /* blocks */
int  s3d_scalars_upload(s3d_ctx *ctx, double *d_s, int n, const double *h_in);
/* pinned host memory for callers that want overlapped transfers */
int  s3d_host_alloc(s3d_ctx *ctx, size_t bytes, void **h_out);
int  s3d_host_free(s3d_ctx *ctx, void *h_ptr);

/* ------------------------------------------------------------------ SpMV (SMat::apply / apply_res, matvec.cpp:59-146) */

#define S3D_SPMV_AUTO      0   /* pick the fastest variant for the grid */
#define S3D_SPMV_MARCH     1   /* register z-marching, i-neighbours by warp shuffle, j-neighbours via L1 */
#define S3D_SPMV_TMA       2   /* z-planes of x staged through a shared-memory ring by TMA bulk row copies (cp.async.bulk, mbarrier completion) */
#define S3D_SPMV_NAIVE     3   /* one thread per point, all seven neighbours from global (the first correct version) */
#define S3D_SPMV_PAIR      4   /* two points per thread, no marching: 128-bit loads, x-reuse through L1/L2 (AUTO resolves to MARCH: measured faster) */

/* y = A x (d_b == NULL) or y = b - A x.  The seven products are formed and summed in stencil order
 * 0..6 with separately rounded multiplies and adds, exactly the reference's expression.
 * Optional fused reductions into device scalars d_red (may be NULL):
 *   d_w != NULL : d_red[0] = (w, y) over real points      (CG's (p, Ap): pass d_w = d_x)
 *   want_yy     : d_red[1] = (y, y)                        (residual norm^2 for Richardson / BiCGSTAB)
 * d_stop (may be NULL): device int; when *d_stop != 0 the launch does nothing (device-side convergence gate).
 * In a z-slab decomposition the caller must have exchanged x's halo planes first (s3d_halo_exchange). */
int  s3d_spmv(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *d_x, double *d_y,
              const double *d_b, const double *d_w, int want_yy, double *d_red, const int *d_stop, int variant);

/* SMat::apply / apply_res on a z-slab in ONE call, fused reductions as in s3d_spmv.  With peer-memory windows (the default, see
 * s3d_halo_exchange) a push kernel stores this rank's two boundary planes of x into the neighbours' windows and the multiply reads the
 * neighbours' planes from its own window: the two k-chunks that need them are dispatched last and only their blocks wait for the flags,
 * so the transfer overlaps the interior planes (the ghost planes of d_x are not written in this mode).  Without peer memory: halo
 * exchange by NCCL, then the multiply; with "halo_overlap" the NCCL exchange runs on a second stream beside the interior planes and the
 * two boundary planes follow.  One rank: identical to s3d_spmv. */
int  s3d_spmv_exchange(s3d_ctx *ctx, const s3d_grid *g, const double *A, double *d_x, double *d_y, const double *d_b,
                       const double *d_w, int want_yy, double *d_red, const int *d_stop);

/* SMat::apply / apply_res for HOST vectors (the reference's SVec lives in host memory): y = A x, or b - A x with a device-resident
 * d_b.  Pipelined in nchunks z-chunks (<= 0: 16; bench.py uses 32) over three streams, so the upload of x, the kernel and the download of y overlap and
 * the call runs at PCIe speed.  h_x / h_y: host SVec layout, pinned (s3d_host_alloc) for full speed; d_x / d_y: device vectors used
 * as staging (d_y as allocated by s3d_vec_alloc: zero ghosts).  Single GPU; blocks until h_y is complete. */
int  s3d_spmv_host(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *h_x, double *h_y, double *d_x, double *d_y,
                   const double *d_b, int nchunks);

/* ------------------------------------------------------------------ BLAS-1 (matvec.cpp:148-288), real points only */

int  s3d_vecset(s3d_ctx *ctx, const s3d_grid *g, double a, double *d_x);                        /* vecset     :188 */
int  s3d_vecassign(s3d_ctx *ctx, const s3d_grid *g, const double *d_x, double *d_y);            /* vecassign  :203 */
int  s3d_veccopy_all(s3d_ctx *ctx, const s3d_grid *g, const double *d_x, double *d_y);          /* ghosts too: NoSolver::apply, s3d_solverbase.cpp:21-33 */
int  s3d_vecaxpy(s3d_ctx *ctx, const s3d_grid *g, s3d_scalar a, const double *d_x, double *d_y);/* vecaxpy    :148  y += a x */
/* y = a x + b y  (p = z + beta p of CG; b = 0 never reads y) */
int  s3d_vecaxpby(s3d_ctx *ctx, const s3d_grid *g, s3d_scalar a, const double *d_x, s3d_scalar b, double *d_y);
/* w = a x + b y, optionally d_red[0] = (w,w) */
int  s3d_vecwaxpby(s3d_ctx *ctx, const s3d_grid *g, s3d_scalar a, const double *d_x, s3d_scalar b,
                   const double *d_y, double *d_w, double *d_red);
/* vec_multi_axpy :166  y += sum_l a[l] x_l in ONE pass (the reference makes n passes); a: n device doubles,
 * xs: n device vector pointers (host array of device pointers), n <= 32 */
int  s3d_vec_multi_axpy(s3d_ctx *ctx, const s3d_grid *g, int n, const double *d_a, const double *const *xs, double *d_y);
/* d_red[l] = (x_l, y), l < n, one pass over y; n <= 32 */
int  s3d_vec_multi_dot(s3d_ctx *ctx, const s3d_grid *g, int n, const double *const *xs, const double *d_y, double *d_red);
int  s3d_dot(s3d_ctx *ctx, const s3d_grid *g, const double *d_x, const double *d_y, double *d_red);  /* inner_vector_l2 :257 */
int  s3d_nrm2sq(s3d_ctx *ctx, const s3d_grid *g, const double *d_x, double *d_red);                   /* norm_vector_l2^2 :240 */
/* norm_L2^2 :221 -- weights 1/8 (cx[i+1]-cx[i-1])(cy..)(cz..); h_c*: host 1-D coordinate arrays of n+2 entries */
int  s3d_normL2sq(s3d_ctx *ctx, const s3d_grid *g, const double *h_cx, const double *h_cy, const double *h_cz,
                  const double *d_x, double *d_red);
/* blocking conveniences returning the reference's values (sqrt applied where the reference applies it) */
int  s3d_dot_host(s3d_ctx *ctx, const s3d_grid *g, const double *d_x, const double *d_y, double *h_out);
int  s3d_nrm2_host(s3d_ctx *ctx, const s3d_grid *g, const double *d_x, double *h_out);
/* sum-allreduce n device scalars across the z-slab ranks: d_global[i] = sum over ranks of d_local[i].
 * Out of place on purpose: kernels write d_local, consumers read d_global, so a repeated call with an unchanged
 * d_local (what happens once the device-side gate has stopped the kernels) is idempotent.  One rank: a copy
 * (nothing at all when d_local == d_global). */
int  s3d_allreduce_sum(s3d_ctx *ctx, const double *d_local, double *d_global, int n);

/* ------------------------------------------------------------------ preconditioners */

/* JacobiPreconditioner::apply (s3d_jacobi.cpp:14-29): z = r / diag(A), a true division */
int  s3d_jacobi_apply(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *d_r, double *d_z);

/* SGS_preconditioner::updateOperator (s3d_sgspreconditioners.cpp:122-135): dinv = 1/diag, matrix-compact layout (cstride doubles) */
int  s3d_sgs_setup(s3d_ctx *ctx, const s3d_grid *g, const double *A, double *d_dinv);
/* StrILU_preconditioner::updateOperator, sequential build (s3d_ilu.cpp:17-64): dinv = 1 / (ILU(0) diagonal); use with s3d_sgs_apply.
 * nbuildsweeps > 1 is idempotent for the sequential build and is treated as 1; 0 gives 1/diag.  On a z-slab decomposition the
 * recurrence runs inside every slab (block ILU(0) by slabs: what the slab-local application, S3D_SGS_SLABLOCAL, goes with). */
int  s3d_strilu_setup(s3d_ctx *ctx, const s3d_grid *g, const double *A, int nbuildsweeps, double *d_dinv);
/* The reference's THREADED build (threadedbuild = true, s3d_ilu.cpp:28-32): nbuildsweeps >= 1 ASYNCHRONOUS sweeps from d = diag(A), chunks of rows
 * (tuning key sgs_async_chunk) dealt to warps that never wait for one another.  Timing-dependent; exact after at most nx+ny+nz sweeps. */
int  s3d_strilu_setup_async(s3d_ctx *ctx, const s3d_grid *g, const double *A, int nbuildsweeps, double *d_dinv);
#define S3D_SGS_LEXICOGRAPHIC 0  /* the reference's ordering, executed as a tile-level dataflow wavefront: same arithmetic */
#define S3D_SGS_LEXICOGRAPHIC_PLANES 2  /* same result, one launch per hyperplane i+j+k (slow; kept for cross-checks) */
#define S3D_SGS_REDBLACK      1  /* two-colour ordering: fully parallel half-sweeps (another operator: changes the iteration counts) */
#define S3D_SGS_SLABLOCAL     4  /* the exact lexicographic sweep INSIDE every z-slab, nothing taken from the neighbouring slabs (block Jacobi across
                                  * the slab faces): no exchange and no cross-rank wait, so it scales; one rank: identical to LEXICOGRAPHIC */
#define S3D_SGS_JACOBI_SWEEPS 3  /* nsweeps synchronous Jacobi sweeps per triangular factor: the deterministic counterpart of the
                                  * reference's asynchronous threaded sweeps (threadedapply = true); exact for nsweeps >= nx+ny+nz-2 */
#define S3D_SGS_ASYNC         5  /* nsweeps ASYNCHRONOUS sweeps per triangular factor, the reference's threaded apply itself
                                  * (s3d_sgspreconditioners.cpp:30-112 with threadedapply = true): chunks of rows dealt to persistent warps that
                                  * never wait for one another (tuning key sgs_async_chunk = -s3d_thread_chunk_size), exact recurrence along a
                                  * row.  Timing-dependent like the reference's; across z-slabs it runs slab by slab. */
#define S3D_SGS_OVERLAP       6  /* z-slabs: the exact lexicographic sweep over every slab EXTENDED by one plane of either neighbour, own planes kept
                                  * (restricted additive Schwarz, SGS as the subdomain solver): no cross-rank wait like SLABLOCAL, one halo
                                  * exchange of r per application, and iteration counts within a few per cent of the exact sweep over the whole
                                  * grid.  Needs the slack planes of s3d_mat_alloc / s3d_coef_alloc and s3d_sgs_overlap_setup; one rank: LEXICOGRAPHIC.
                                  * The call fills the ghost planes of d_r (halo exchange) and leaves intermediate values in those of d_z / d_y. */
#define S3D_SGS_BLOCKS        7  /* ONE device: the grid cut into blocks of planes and of rows (tuning keys sgs_blocks, sgs_blocks_y; 0 = by size, about 32
                                  * planes / rows each, none above 96 Mi points) that sweep CONCURRENTLY -- forward, each block is preceded by one plane
                                  * and one row of the blocks below it (zero inflow under them); backward, by one plane and row of the blocks above, whose
                                  * forward values their owners have stored by then.  The dependency chain of a sweep, which bounds the exact sweep up to
                                  * ~256^3, shrinks from nx/16 + ny/8 + nz/8 tile hops to about nx/16 + ny/(4 By) + nz/(8 Bz) + 2 (256^3: 0.79 -> 0.43 ms per
                                  * application).  Another operator than the reference's sweep: deterministic, iteration counts within a few per cent
                                  * of it (north_star's gate for a parallel SGS); opt-in, LEXICOGRAPHIC stays the default. */
/* The blocks ordering's host-side plan, for inspection and tests (no device, no context): what tile layers the kernel works through and in
 * which order it takes the tiles.  direction 0 = forward sweep, 1 = backward.  dims_out = {ntx, nty, ntz}; layers_out (may be NULL): 4 ints per
 * layer {first plane / row, count, predecessor-in-block | successor-in-block << 1, plane / row not stored or a negative number}, the ntz layers
 * of planes then the nty layers of rows; order_out (may be NULL): ntx*nty*ntz tile numbers (ck * nty + cj) * ntx + ci in the order of the tickets
 * (forward order[ticket], backward order[ntiles - 1 - ticket]).  pencils = 1 or 2 (tiles of 4 or 8 rows), tile_i = 16 or 32. */
int  s3d_sgs_blocks_plan(int nx, int ny, int nz, int blocks, int blocks_y, int tile_i, int pencils, int direction, int *dims_out, int *layers_out,
                         int *order_out);
/* Once per operator, after s3d_sgs_setup / s3d_strilu_setup, for S3D_SGS_OVERLAP: the neighbouring slabs' boundary planes of the seven
 * coefficient arrays and of dinv go into the slack planes (grouped ncclSend / ncclRecv).  Collective; a no-op on one rank. */
int  s3d_sgs_overlap_setup(s3d_ctx *ctx, const s3d_grid *g, double *A, double *d_dinv);
/* Exact (lexicographic) SGS on a z-slab decomposition: the sweep runs THROUGH the slabs -- the last plane of tiles of rank p pushes its
 * k-face rows into rank p+1's peer-memory window (backward: p-1) and raises per-tile flags there, so results and iteration counts are
 * those of the single-process reference (s3d_sgspreconditioners.cpp:14-115) at any GPU count; the ranks work as a pipeline.
 * *yes = 1 when that is possible for this grid (one rank: always; several: the CUDA-IPC windows could be mapped on every rank).
 * Collective over the ranks of the communicator. */
int  s3d_sgs_slab_sweeps_available(s3d_ctx *ctx, const s3d_grid *g, int *yes);

/* SGS_like_preconditioner::apply (s3d_sgspreconditioners.cpp:14-115): z = (D+U)^-1 D (D+L)^-1 r.
 * d_y: scratch vector (vlen doubles); it is re-zeroed by the call like the reference's temporary. */
int  s3d_sgs_apply(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *d_dinv,
                   const double *d_r, double *d_z, double *d_y, int ordering, int nsweeps);
/* Red-black SGS on COLOUR-PACKED coefficients: a colour occupies every other entry of a row, so the plain red-black kernels use half of every
 * sector of the coefficient arrays they touch.  s3d_sgs_rb_pack copies A0 A1 A2 A4 A5 A6 and dinv once per operator (after s3d_sgs_setup /
 * s3d_strilu_setup) into a buffer of s3d_sgs_rb_pack_size doubles in which the entries of one colour are contiguous; s3d_sgs_apply_rb_packed
 * is s3d_sgs_apply(..., S3D_SGS_REDBLACK, ...) reading from it: same expressions, bit-identical results (no counterpart in the reference). */
int  s3d_sgs_rb_pack_size(const s3d_grid *g, size_t *ndoubles);
int  s3d_sgs_rb_pack_alloc(s3d_ctx *ctx, const s3d_grid *g, double **d_out);   /* a buffer of that size; release it with s3d_free */
int  s3d_sgs_rb_pack(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *d_dinv, double *d_packed);
int  s3d_sgs_apply_rb_packed(s3d_ctx *ctx, const s3d_grid *g, const double *d_packed, const double *d_r, double *d_z, double *d_y);

/* ------------------------------------------------------------------ fused solver steps (new: CG / BiCGSTAB) */

/* CG step 1:  x += alpha p;  r -= alpha q;  d_red[0] = (r,r);  if A != NULL also d_red[1] = (r, r/diag(A))
 * (the Jacobi-preconditioned (r,z) without storing z). */
int  s3d_cg_update(s3d_ctx *ctx, const s3d_grid *g, s3d_scalar alpha, const double *d_p, const double *d_q,
                   double *d_x, double *d_r, const double *A_or_null, double *d_red, const int *d_stop);
/* CG step 2 with Jacobi folded in:  p = r/diag(A) + beta p */
int  s3d_cg_direction_jacobi(s3d_ctx *ctx, const s3d_grid *g, s3d_scalar beta, const double *A, const double *d_r,
                             double *d_p, const int *d_stop);
/* BiCGSTAB:  p = r + beta (p - omega v) */
int  s3d_bicg_direction(s3d_ctx *ctx, const s3d_grid *g, s3d_scalar beta, s3d_scalar omega, const double *d_r,
                        const double *d_v, double *d_p, const int *d_stop);
/* BiCGSTAB:  x += alpha ph + omega sh;  r = s - omega t;  d_red[0] = (r,r);  d_red[1] = (rhat, r) */
int  s3d_bicg_update(s3d_ctx *ctx, const s3d_grid *g, s3d_scalar alpha, const double *d_ph, s3d_scalar omega,
                     const double *d_sh, const double *d_s, const double *d_t, const double *d_rhat,
                     double *d_x, double *d_r, double *d_red, const int *d_stop);
/* Device-side convergence gate: if sqrt(*d_rr) <= rtol * sqrt(*d_bb) set *d_stop = 1, else ++*d_iters.
 * Kernels given d_stop become no-ops once it is set, so the host may enqueue iterations ahead and
 * poll, and x is exactly the iterate at which the test first passed. */
#define S3D_CONV_STRICT   1   /* stop when rel <  rtol (GCR's inner test, s3d_gcr.cpp:57) instead of !(rel > rtol) */
#define S3D_CONV_INITIAL  2   /* test the initial residual: do not count or record an iteration */
#define S3D_STOP_CONVERGED 1
#define S3D_STOP_BREAKDOWN 2
#define S3D_STOP_MAXITER   3
int  s3d_converge_test(s3d_ctx *ctx, const double *d_rr, const double *d_bb, double rtol, int maxiter, int flags,
                       int *d_stop, int *d_iters, double *d_hist_or_null);
/* tiny device-side scalar arithmetic so Krylov coefficients stay on the GPU:
 *   op 0: out[l] = scale * a[l] / b[l], l < n       (GCR's beta_i = -(q_new,q_i)/(q_i,q_i))
 *   op 1: out[0] = (a/b) * (c/d)                     (BiCGSTAB's beta; a, b or d == 0 sets *d_stop = 2)
 *   op 2: out[0] = a / b                             (alpha, omega; b == 0 sets *d_stop = 2)
 *   op 3: out[l] = a[l], l < n                       (keep a reduction result for later) */
int  s3d_scalar_op(s3d_ctx *ctx, int op, double *d_out, const double *a, const double *b, const double *c,
                   const double *d, int n, double scale, int *d_stop);
/* A whole solve of a SMALL grid in one kernel launch (no counterpart in the reference, whose loops call their kernels one by one,
 * src/linalg/s3d_solverbase.cpp:45-80): one block runs the complete loop with block barriers between the phases, so an iteration
 * costs no launch at all.  method: 0 Richardson, 1 CG (GCR: s3d_small_solve_gcr); pc: 0 identity, 1 Jacobi; same expressions and roundings as the stand-alone
 * kernels, same convergence test as s3d_converge_test.  d_flags[0] = S3D_STOP_* code, d_flags[1] = iterations, d_scal[0] = (r,r) at
 * exit, d_scal[1] = (b,b), d_hist_or_null[it] = relative residual after iteration it (hist_cap entries).  One rank, at most
 * S3D_SMALL_MAX_POINTS points. */
#define S3D_SMALL_MAX_POINTS (1 << 22)
int  s3d_small_solve(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *d_b, double *d_x, double *d_w0, double *d_w1,
                     double *d_w2, int method, int pc, double rtol, int maxiter, int *d_flags, double *d_scal, double *d_hist_or_null,
                     int hist_cap);
/* The same for GCR(north) (reference loop: src/linalg/s3d_gcr.cpp:14-86): d_res, d_z work vectors; pq = HOST array of 2*north device
 * pointers, pq[2i] = p_i, pq[2i+1] = q_i; north <= 32.  The iteration cap is looked at between cycles only and the inner test is the
 * strict one, as in the reference.  d_scal[0] = 1 when no iteration ran. */
int  s3d_small_solve_gcr(s3d_ctx *ctx, const s3d_grid *g, const double *A, const double *d_b, double *d_x, double *d_res, double *d_z,
                         double *const *pq, int north, int pc, double rtol, int maxiter, int *d_flags, double *d_scal,
                         double *d_hist_or_null, int hist_cap);
/* Context-wide gate: while set, EVERY kernel this context launches (SpMV, BLAS-1, preconditioners ...) first
 * reads *d_stop and does nothing when it is non-zero, unless the call passes its own d_stop.  A host-driven
 * solver sets the gate to its stop flag for the duration of the loop; pass NULL to clear it. */
int  s3d_ctx_set_gate(s3d_ctx *ctx, const int *d_stop);
/* Accumulating GPU stopwatch (CUDA events on the stream) for phase timing such as SolveInfo::precapplywtime:
 * begin/end may be called any number of times; total_ms synchronises and returns the sum, then resets the slot. */
int  s3d_phase_begin(s3d_ctx *ctx, int slot);
int  s3d_phase_end(s3d_ctx *ctx, int slot);
int  s3d_phase_total_ms(s3d_ctx *ctx, int slot, double *ms);
/* CUDA graphs for the host-driven loops (no counterpart in the reference, whose loops call their kernels directly,
 * src/linalg/s3d_solverbase.cpp:45-80, s3d_gcr.cpp:14-86).  Between graph_begin and graph_end every s3d_* call that enqueues work on
 * the context's stream is RECORDED instead of run; graph_launch replays the recording.  A batch of solver iterations qualifies because
 * its coefficients, iteration counter, residual history index and stop flag all live in device memory.  Calls that allocate or
 * synchronise (first-use scratch, downloads, s3d_phase_total_ms) must not be made inside a capture: run one batch eagerly first.
 * s3d_phase_begin/end inside a capture are ignored.  s3d_ctx_launch_count counts a replay as the kernels it contains. */
typedef struct s3d_graph s3d_graph;
int  s3d_graph_begin(s3d_ctx *ctx);
int  s3d_graph_end(s3d_ctx *ctx, s3d_graph **out);
int  s3d_graph_launch(s3d_ctx *ctx, s3d_graph *g);
int  s3d_graph_destroy(s3d_ctx *ctx, s3d_graph *g);
int  s3d_ints_alloc(s3d_ctx *ctx, int n, int **d_out);
int  s3d_ints_download(s3d_ctx *ctx, const int *d_s, int n, int *h_out);
int  s3d_ints_zero(s3d_ctx *ctx, int *d_s, int n);
/* Pipelined polling of device flags: `snapshot` enqueues an asynchronous copy of n (<= 8) device ints into pinned slot
 * `slot` (0..7) and records an event; `snapshot_wait` spins on that event (no blocking driver wait) and returns the values.
 * A host-driven loop snapshots after batch k, enqueues batch k+1, and only then waits for snapshot k. */
int  s3d_ints_snapshot(s3d_ctx *ctx, const int *d_s, int n, int slot);
int  s3d_ints_snapshot_wait(s3d_ctx *ctx, int slot, int n, int *h_out);

/* ------------------------------------------------------------------ assembly (src/pde) */

/* PDEBase::computeLHS + Poisson::lhsmat_kernel (pdebase.cpp:123-150, poisson.cpp:28-52).
 * h_cx/h_cy/h_cz: host 1-D coordinates incl. boundary points (nx+2, ny+2, nz_global+2 entries). */
int  s3d_assemble_poisson(s3d_ctx *ctx, const s3d_grid *g, const double *h_cx, const double *h_cy,
                          const double *h_cz, double *A);
/* PDEBase::computeLHS + ConvDiff::lhsmat_kernel (convdiff.cpp:48-88): mu * diffusion, then first-order upwind */
int  s3d_assemble_convdiff(s3d_ctx *ctx, const s3d_grid *g, const double *h_cx, const double *h_cy,
                           const double *h_cz, const double vel[3], double mu, double *A);
/* PDEBase::computeLHS + ConvDiffCirc::lhsmat_kernel (convdiff_circular.cpp:59-102): mu * diffusion plus upwinding
 * with the rotational velocity (2y(1-x^2), -2x(1-y^2), sin(pi z)), reproduced as the reference has it. */
int  s3d_assemble_convdiff_circ(s3d_ctx *ctx, const s3d_grid *g, const double *h_cx, const double *h_cy,
                                const double *h_cz, double mu, double *A);
/* PDEBase::computeVector for separable functions f = sum_t X_t(x_i) Y_t(y_j) Z_t(z_k)  (pdebase.cpp:50-72):
 * h_tx: nterms*(nx+2) host table, h_ty: nterms*(ny+2), h_tz: nterms*(nz_global+2), row t = term t, evaluated
 * at the grid coordinates (boundary entries unused).  Each term is ((X*Y)*Z), terms are added in order.
 * h_face (may be NULL): 6 per-face additive terms for the boundary-adjacent layers, in the reference's face
 * order i-,i+,j-,j+,k-,k+ (Poisson::rhs_kernel, poisson.cpp:54-122), already divided by their spacing factor. */
int  s3d_assemble_rhs_separable(s3d_ctx *ctx, const s3d_grid *g, int nterms, const double *h_tx, const double *h_ty,
                                const double *h_tz, const double *h_face, double *d_f);

/* ------------------------------------------------------------------ multi-GPU: z-slabs, one process per GPU */

#define S3D_COMM_ID_BYTES 128
int  s3d_comm_unique_id(void *id_out /* S3D_COMM_ID_BYTES */);       /* rank 0 creates, caller broadcasts */
int  s3d_comm_init(s3d_ctx *ctx, int rank, int nranks, const void *id);
int  s3d_comm_destroy(s3d_ctx *ctx);
/* fills ghost planes k=-1 and k=nz of x with the neighbours' boundary planes; the global ends keep their Dirichlet zeros.
 * No-op for one rank.  The planes go through peer memory: each rank's window is mapped by its two neighbours (CUDA IPC, set up at
 * the first exchange of a plane size -- a collective step), a push kernel stores my planes into their windows over NVLink and raises
 * their flags, a pull kernel waits for mine and fills the ghost planes.  s3d_tune_set("halo_p2p", 0), or a failed IPC setup on any
 * rank, selects grouped ncclSend/ncclRecv instead.  s3d_spmv_exchange goes one step further: its SpMV reads the window itself. */
int  s3d_halo_exchange(s3d_ctx *ctx, const s3d_grid *g, double *d_x);

#ifdef __cplusplus
}
#endif
#endif /* S3D_CUDA_H */
