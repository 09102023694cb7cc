// Grid vectors and 7-point stencil matrices, device-resident.
// Interface of the reference's src/linalg/matvec.hpp:23-115 (same type, member and function names,
// same argument meaning, same exceptions); storage and arithmetic live on the GPU behind the C ABI.
#ifndef STRUCT3D_B200_MATVEC_HPP
#define STRUCT3D_B200_MATVEC_HPP

#include <array>
#include <vector>
#include "cartmesh.hpp"
#include "s3d_device.hpp"

/// Vector over a 3-D structured grid with one ghost layer; 'i' is the fastest index.
/// Device layout: ghost-padded rows, real i=0 aligned to 128 bytes (include/s3d_cuda.h).
/// `vals` is a coherent host view in the reference layout: vals[m->localFlattenedIndexAll(k,j,i)].
struct SVec
{
	explicit SVec(const CartMesh *const mesh);
	SVec();
	SVec(const SVec &other);              ///< device-to-device copy, ghosts included
	SVec(SVec &&other) noexcept;
	SVec &operator=(const SVec &) = delete;
	~SVec();

	/// (re)allocate over a mesh; zero-filled, ghosts included
	void init(const CartMesh *const mesh);

	const CartMesh *m;
	const sint start;                     ///< first real index per axis (1)
	const int nghost;                     ///< ghost layers per side (1)
	std::array<sint, NDIM> sz;            ///< real points per axis owned by this rank (sz[2] = z-slab thickness)
	s3d::HostMirror vals;

	// ---- device side (B200 additions)
	const double *dev() const { vals.device_will_read(); return d_; }
	double *dev_rw() { vals.device_will_read(); vals.device_wrote(); return d_; }
	double *dev_w() { vals.device_will_read(); vals.device_wrote(); return d_; }
	/// device pointer for a kernel that reads the vector AND refreshes its z-ghost planes (halo exchange)
	double *dev_halo() const { vals.device_will_read(); if (s3d::Device::get().nranks() > 1) vals.device_wrote(); return d_; }
	const s3d_grid &grid() const { return m->device_grid(); }
	/// make ghost planes k=-1 / k=nz current with the z-neighbours' data (no-op on one rank)
	void exchange_halo() const;

private:
	void allocate();
	void release();
	double *d_ = nullptr;
};

/// Matrix for a 7-point stencil on a 3-D structured grid; stencil position is the slowest index:
/// vals[s][m->localFlattenedIndexReal(k,j,i)], s in {i-1, j-1, k-1, diag, i+1, j+1, k+1}.
struct SMat
{
	explicit SMat(const CartMesh *const mesh);
	SMat(const SMat &other);
	SMat(SMat &&other) noexcept;
	SMat &operator=(const SMat &) = delete;
	~SMat();

	/// y = A x (y is overwritten).  Throws std::runtime_error unless x and y share a mesh.
	void apply(const SVec &x, SVec &y) const;
	/// y = b - A x.  Throws unless all vectors and the matrix share a mesh.
	void apply_res(const SVec &b, const SVec &x, SVec &y) const;

	const CartMesh *const m;
	const sint start;                     ///< 0
	const int nghost;                     ///< 1
	const std::array<sint, NDIM> sz;
	s3d::HostMirror vals[NSTENCIL];

	// ---- device side
	const double *dev() const;
	double *dev_w();
	const s3d_grid &grid() const { return m->device_grid(); }

private:
	void bind_views();
	double *d_ = nullptr;
};

/// x[:] <- a over real points
void vecset(const sreal a, SVec &x);
/// y <- x over real points
void vecassign(const SVec &x, SVec &y);
/// y <- y + a x
void vecaxpy(const sreal a, const SVec &x, SVec &y);
/// y <- y + sum_l a[l] x[l]   (one fused pass on the GPU, same rounding order as numvecs passes)
void vec_multi_axpy(const int numvecs, const sreal *const a, const SVec *const x, SVec &y);
/// sqrt(sum x^2 vol) with the dual-cell volumes of the mesh; real points only
sreal norm_L2(const SVec &x);
/// sqrt(sum x^2) over real points (all ranks)
sreal norm_vector_l2(const SVec &x);
/// sum x y over real points (all ranks); no square root
sreal inner_vector_l2(const SVec &x, const SVec &y);
/// norm_L2(y - x)
sreal compute_error_L2(const SVec &x, const SVec &y);

#endif
