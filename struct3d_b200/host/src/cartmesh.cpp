// Cartesian mesh: coordinate generation and sizes (reference: src/cartmesh.cpp:102-209, native half).
#include <limits>
#include "cartmesh.hpp"
#include "common_utils.hpp"
#include "s3d_device.hpp"

CartMesh::CartMesh() : coords{nullptr}, npointotal{0}, ninpoin{0}, h{0}, ntprocs{1}, nghost{1}
{
	for (int d = 0; d < NDIM; d++) { npoind[d] = 0; nprocs[d] = 1; }
	std::memset(&dgrid, 0, sizeof dgrid);
}

CartMesh::~CartMesh()
{
	if (!coords) return;
	for (int d = 0; d < NDIM; d++) std::free(coords[d]);
	std::free(coords);
}

int CartMesh::createMesh(const sint npdim[NDIM])
{
	const s3d::Device &dev = s3d::Device::get();
	const bool talk = dev.rank() == 0;
	for (int d = 0; d < NDIM; d++) {
		if (npdim[d] < 3) throw std::runtime_error("CartMesh: need at least one interior point per axis");
		npoind[d] = npdim[d];
	}
	npointotal = npoind[0] * npoind[1] * npoind[2];
	// boundary points: two full k-faces, two j-faces without their k-edges, two i-faces without j/k edges
	const sint nbpoints = 2 * npoind[0] * npoind[1] + 2 * (npoind[2] - 2) * npoind[0] + 2 * (npoind[1] - 2) * (npoind[2] - 2);
	ninpoin = npointotal - nbpoints;

	// z-slab decomposition: one process per GPU, all of x and y, a contiguous range of z-planes
	nprocs[0] = nprocs[1] = 1;
	nprocs[2] = dev.nranks();
	ntprocs = dev.nranks();
	int rc;
	if (dev.nranks() == 1) rc = s3d_grid_init(&dgrid, npoind[0] - 2, npoind[1] - 2, npoind[2] - 2);
	else rc = s3d_grid_init_slab(&dgrid, npoind[0] - 2, npoind[1] - 2, npoind[2] - 2, dev.rank(), dev.nranks());
	if (rc != S3D_OK) throw std::runtime_error("CartMesh: cannot lay the grid out on the device (too few z-planes for the number of GPUs?)");

	coords = (sreal **)std::malloc(NDIM * sizeof(sreal *));
	for (int d = 0; d < NDIM; d++) coords[d] = (sreal *)std::malloc(npoind[d] * sizeof(sreal));

	if (talk) {
		std::printf("CartMesh: Number of points in each direction: %d,%d,%d.\n", npoind[0], npoind[1], npoind[2]);
		std::printf("CartMesh: Number of GPUs (z-slabs): %d; this rank owns z-planes [%d,%d).\n", ntprocs, dgrid.k0, dgrid.k0 + dgrid.nz);
		std::printf("CartMesh: Total points = %d, interior points = %d, partitions = %d\n", npointotal, ninpoin, ntprocs);
	}
	return 0;
}

// Longest cell diagonal.  The reference scans every cell (O(N), cartmesh.cpp:10-33); the diagonal is
// monotone in each spacing, so the maximum sits at the per-axis maxima: same value, O(n).
void CartMesh::computeMeshSize()
{
	sreal sum = 0;
	for (int d = 0; d < NDIM; d++) {
		sreal hd = 0;
		for (sint i = 0; i < npoind[d] - 1; i++) {
			const sreal s = coords[d][i + 1] - coords[d][i];
			if (s > hd) hd = s;
		}
		sum += hd * hd;
	}
	h = std::sqrt(sum);
}

void CartMesh::generateMesh_ChebyshevDistribution(const sreal rmin[NDIM], const sreal rmax[NDIM])
{
	for (int d = 0; d < NDIM; d++) {
		const sreal theta = PI / (npoind[d] - 1);
		for (sint i = 0; i < npoind[d]; i++)
			coords[d][i] = (rmax[d] + rmin[d]) * 0.5 + (rmax[d] - rmin[d]) * 0.5 * std::cos(PI - i * theta);
	}
	computeMeshSize();
	if (s3d::Device::get().rank() == 0) std::printf("CartMesh: Chebyshev grid, h = %f\n", h);
}

void CartMesh::generateMesh_UniformDistribution(const sreal rmin[NDIM], const sreal rmax[NDIM])
{
	for (int d = 0; d < NDIM; d++)
		for (sint i = 0; i < npoind[d]; i++)
			coords[d][i] = rmin[d] + (rmax[d] - rmin[d]) * i / (npoind[d] - 1);
	computeMeshSize();
	if (s3d::Device::get().rank() == 0) std::printf("CartMesh: uniform grid, h = %f\n", h);

	for (int d = 0; d < NDIM; d++) {
		const sreal dr = coords[d][1] - coords[d][0];
		for (sint j = 0; j < npoind[d] - 1; j++)
			if (std::abs(coords[d][j + 1] - coords[d][j] - dr) > std::numeric_limits<sreal>::epsilon())
				throw std::runtime_error("Uniform mesh is not actually uniform!");
	}
}
