// Non-uniform Cartesian grid: the three 1-D coordinate arrays whose tensor product is the mesh.
// Interface of the reference's native CartMesh (src/cartmesh.hpp:67-154); the PETSc/DMDA half is out
// of scope.  New here: the mesh also carries the device grid descriptor (s3d_grid) of this rank's
// z-slab, which every SVec/SMat defined over the mesh uses.
#ifndef STRUCT3D_B200_CARTMESH_HPP
#define STRUCT3D_B200_CARTMESH_HPP

#include "s3d_cuda.h"
#include "struct3d_config.h"

class CartMesh
{
public:
	CartMesh();
	~CartMesh();
	CartMesh(const CartMesh &) = delete;
	CartMesh &operator=(const CartMesh &) = delete;

	/// Allocates the coordinate arrays for npdim[] points per axis (boundary points included) and
	/// derives this rank's z-slab (all of z when the process runs alone).  Returns 0.
	int createMesh(const sint npdim[NDIM]);

	/// x_i = (a+b)/2 + (b-a)/2 cos(pi - i*theta), theta = pi/(N-1)   (reference cartmesh.cpp:159-178)
	void generateMesh_ChebyshevDistribution(const sreal rmin[NDIM], const sreal rmax[NDIM]);
	/// x_i = a + (b-a) i/(N-1); throws if the spacing deviates by more than epsilon (cartmesh.cpp:181-209)
	void generateMesh_UniformDistribution(const sreal rmin[NDIM], const sreal rmax[NDIM]);

	sint gnpoind(const int idim) const { return npoind[idim]; }
	sint gnghost() const { return nghost; }
	sreal gcoords(const int idim, const sint ipoin) const { return coords[idim][ipoin]; }
	sint gnpointotal() const { return npointotal; }
	sint gninpoin() const { return ninpoin; }
	sreal gh() const { return h; }
	const sint *pointer_npoind() const { return npoind; }
	const sreal *const *pointer_coords() const { return coords; }

	/// Flat index over real points only, k slowest / i fastest.  NOTE the (k,j,i) argument order.
	sint localFlattenedIndexReal(const sint k, const sint j, const sint i) const
	{
		return k * (npoind[1] - 2) * (npoind[0] - 2) + j * (npoind[0] - 2) + i;
	}
	/// Flat index including the ghost layer.  NOTE the (k,j,i) argument order.
	sint localFlattenedIndexAll(const sint k, const sint j, const sint i) const
	{
		return k * npoind[1] * npoind[0] + j * npoind[0] + i;
	}

	// ---- B200 additions
	/// Device layout of this rank's slab (local nz, pitches, offsets).
	const s3d_grid &device_grid() const { return dgrid; }
	/// Real points along z owned by this rank and the global index of the first one.
	sint local_nz() const { return dgrid.nz; }
	sint z_offset() const { return dgrid.k0; }

protected:
	void computeMeshSize();

	sint npoind[NDIM];
	sreal **coords;
	sint npointotal;
	sint ninpoin;
	sreal h;
	int nprocs[NDIM];
	int ntprocs;
	sint nghost;
	s3d_grid dgrid;
};

#endif
