// SVec / SMat and the BLAS-1 free functions: host objects whose data and arithmetic live on the GPU.
// Behaviour contract: reference src/linalg/matvec.cpp (constructors :8-57, apply :59-101,
// apply_res :103-146, vecaxpy :148, vec_multi_axpy :166, vecset :188, vecassign :203, norm_L2 :221,
// norm_vector_l2 :240, inner_vector_l2 :257, compute_error_L2 :281) -- same checks, same exceptions.
#include <cassert>
#include "linalg/matvec.hpp"

using s3d::Device;

namespace {
size_t ghosted_len(const s3d_grid &g) { return (size_t)(g.nx + 2) * (g.ny + 2) * (g.nz + 2); }
size_t compact_len(const s3d_grid &g) { return (size_t)g.nx * g.ny * g.nz; }
s3d_scalar imm(double a) { return s3d_scalar{a, nullptr, nullptr}; }
}

// ------------------------------------------------------------------ SVec

SVec::SVec(const CartMesh *const mesh) : m{mesh}, start{1}, nghost{1}
{
	allocate();
}

SVec::SVec() : m{nullptr}, start{1}, nghost{1}, sz{{0, 0, 0}} {}

SVec::SVec(const SVec &o) : m{o.m}, start{1}, nghost{1}
{
	if (!m) { sz = {0, 0, 0}; return; }
	allocate();
	const Device &d = Device::get();
	d.check(s3d_veccopy_all(d.ctx(), &grid(), o.dev(), d_), "s3d_veccopy_all");
}

SVec::SVec(SVec &&o) noexcept : m{o.m}, start{1}, nghost{1}, sz(o.sz), d_{o.d_}
{
	if (d_) {
		// pending host edits of the source must reach the device before the view is re-bound
		try { o.vals.device_will_read(); } catch (...) {}
		vals.bind(s3d::HostMirror::VECTOR, &m->device_grid(), d_, 0, ghosted_len(m->device_grid()));
	}
	o.d_ = nullptr;
	o.vals.unbind();
}

SVec::~SVec() { release(); }

void SVec::init(const CartMesh *const mesh)
{
	release();
	m = mesh;
	allocate();
}

void SVec::allocate()
{
	if (m->gnghost() != nghost) throw std::runtime_error("Num ghost points don't match between SVec and mesh!");
	const s3d_grid &g = m->device_grid();
	sz = {g.nx, g.ny, g.nz};
	const Device &d = Device::get();
	d.check(s3d_vec_alloc(d.ctx(), &g, &d_), "s3d_vec_alloc");   // zero-filled, ghosts included (matvec.cpp:19-22)
	vals.bind(s3d::HostMirror::VECTOR, &g, d_, 0, ghosted_len(g));
}

void SVec::release()
{
	if (d_) {
		const Device &d = Device::get();
		s3d_free(d.ctx(), d_);
		d_ = nullptr;
	}
	vals.unbind();
}

void SVec::exchange_halo() const
{
	const Device &d = Device::get();
	if (d.nranks() == 1) return;
	vals.device_will_read();
	d.check(s3d_halo_exchange(d.ctx(), &grid(), d_), "s3d_halo_exchange");
	vals.device_wrote();
}

// ------------------------------------------------------------------ SMat

SMat::SMat(const CartMesh *const mesh)
	: m{mesh}, start{0}, nghost{1}, sz{{mesh->device_grid().nx, mesh->device_grid().ny, mesh->device_grid().nz}}
{
	if (m->gnghost() != nghost) throw std::runtime_error("Num ghost points don't match between SMat and mesh!");
	const Device &d = Device::get();
	d.check(s3d_mat_alloc(d.ctx(), &grid(), &d_), "s3d_mat_alloc");
	bind_views();
}

SMat::SMat(const SMat &o) : m{o.m}, start{0}, nghost{1}, sz(o.sz)
{
	const Device &d = Device::get();
	d.check(s3d_mat_alloc(d.ctx(), &grid(), &d_), "s3d_mat_alloc");
	bind_views();
	// matrices are copied rarely: go stencil by stencil through the host mirror
	std::vector<sreal> tmp(compact_len(grid()));
	for (int s = 0; s < NSTENCIL; s++) {
		o.vals[s].copy_to(tmp.data());
		vals[s].assign(tmp.data());
	}
}

SMat::SMat(SMat &&o) noexcept : m{o.m}, start{0}, nghost{1}, sz(o.sz), d_{o.d_}
{
	for (int s = 0; s < NSTENCIL; s++) {
		try { o.vals[s].device_will_read(); } catch (...) {}
		o.vals[s].unbind();
	}
	o.d_ = nullptr;
	bind_views();
}

SMat::~SMat()
{
	if (d_) s3d_free(Device::get().ctx(), d_);
}

void SMat::bind_views()
{
	for (int s = 0; s < NSTENCIL; s++) vals[s].bind(s3d::HostMirror::STENCIL, &m->device_grid(), d_, s, compact_len(grid()));
}

const double *SMat::dev() const
{
	for (int s = 0; s < NSTENCIL; s++) vals[s].device_will_read();
	return d_;
}

double *SMat::dev_w()
{
	for (int s = 0; s < NSTENCIL; s++) { vals[s].device_will_read(); vals[s].device_wrote(); }
	return d_;
}

void SMat::apply(const SVec &x, SVec &y) const
{
	if (x.m != y.m) throw std::runtime_error("Both vectors should be defined over the same mesh!");
	assert(x.nghost == 1);
	const Device &d = Device::get();
	d.check(s3d_spmv_exchange(d.ctx(), &grid(), dev(), x.dev_halo(), y.dev_w(), nullptr, nullptr, 0, nullptr, nullptr), "s3d_spmv_exchange");
}

void SMat::apply_res(const SVec &b, const SVec &x, SVec &y) const
{
	if (x.m != y.m || x.m != b.m || x.m != m)
		throw std::runtime_error("All vectors and matrix should be defined over the same mesh!");
	assert(x.nghost == 1);
	const Device &d = Device::get();
	d.check(s3d_spmv_exchange(d.ctx(), &grid(), dev(), x.dev_halo(), y.dev_w(), b.dev(), nullptr, 0, nullptr, nullptr), "s3d_spmv_exchange");
}

// ------------------------------------------------------------------ free functions

void vecset(const sreal a, SVec &x)
{
	const Device &d = Device::get();
	d.check(s3d_vecset(d.ctx(), &x.grid(), a, x.dev_rw()), "s3d_vecset");
}

void vecassign(const SVec &x, SVec &y)
{
	if (x.m != y.m) throw std::runtime_error("Both vectors should be defined over the same mesh!");
	const Device &d = Device::get();
	d.check(s3d_vecassign(d.ctx(), &x.grid(), x.dev(), y.dev_rw()), "s3d_vecassign");
}

void vecaxpy(const sreal a, const SVec &x, SVec &y)
{
	if (x.m != y.m) throw std::runtime_error("Both vectors should be defined over the same mesh!");
	const Device &d = Device::get();
	d.check(s3d_vecaxpy(d.ctx(), &x.grid(), imm(a), x.dev(), y.dev_rw()), "s3d_vecaxpy");
}

void vec_multi_axpy(const int numvecs, const sreal *const a, const SVec *const x, SVec &y)
{
	if (numvecs == 0) return;
	if (x[0].m != y.m) throw std::runtime_error("Both vectors should be defined over the same mesh!");
	const Device &d = Device::get();
	// one fused pass per group of 32 vectors; coefficients go through a small device buffer
	for (int base = 0; base < numvecs; base += 32) {
		const int n = std::min(32, numvecs - base);
		double *da = d.scalars(n);
		d.check(s3d_scalars_upload(d.ctx(), da, n, a + base), "s3d_scalars_upload");
		const double *xs[32];
		for (int l = 0; l < n; l++) xs[l] = x[base + l].dev();
		d.check(s3d_vec_multi_axpy(d.ctx(), &y.grid(), n, da, xs, y.dev_rw()), "s3d_vec_multi_axpy");
		d.release(da);
	}
}

sreal norm_L2(const SVec &x)
{
	const Device &d = Device::get();
	double *red = d.scalars(2);
	const sreal *const *c = x.m->pointer_coords();
	d.check(s3d_normL2sq(d.ctx(), &x.grid(), c[0], c[1], c[2], x.dev(), red), "s3d_normL2sq");
	d.check(s3d_allreduce_sum(d.ctx(), red, red + 1, 1), "s3d_allreduce_sum");
	sreal v = 0;
	d.check(s3d_scalars_download(d.ctx(), red + 1, 1, &v), "s3d_scalars_download");
	d.release(red);
	return std::sqrt(v);
}

sreal norm_vector_l2(const SVec &x)
{
	const Device &d = Device::get();
	sreal v = 0;
	d.check(s3d_nrm2_host(d.ctx(), &x.grid(), x.dev(), &v), "s3d_nrm2_host");
	return v;
}

sreal inner_vector_l2(const SVec &x, const SVec &y)
{
	if (x.m != y.m) throw std::runtime_error("Both vectors should be defined over the same mesh!");
	if (x.sz[0] != y.sz[0] || x.sz[1] != y.sz[1] || x.sz[2] != y.sz[2]) throw std::runtime_error("Sizes don't match!");
	const Device &d = Device::get();
	sreal v = 0;
	d.check(s3d_dot_host(d.ctx(), &x.grid(), x.dev(), y.dev(), &v), "s3d_dot_host");
	return v;
}

sreal compute_error_L2(const SVec &x, const SVec &y)
{
	assert(x.m == y.m);
	SVec diff = y;
	vecaxpy(-1.0, x, diff);
	return norm_L2(diff);
}
