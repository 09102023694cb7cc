// Process-wide GPU context for the host classes (one process per GPU).
// The host layer reaches CUDA only through the C ABI of include/s3d_cuda.h; this class owns the
// s3d_ctx, translates status codes into std::runtime_error (the reference's error style,
// SURVEY.md 8b), and provides the coherent host mirror behind the public `vals` members.
#ifndef STRUCT3D_B200_DEVICE_HPP
#define STRUCT3D_B200_DEVICE_HPP

#include <cstddef>
#include <stdexcept>
#include <string>
#include <vector>

#include "s3d_cuda.h"
#include "struct3d_config.h"

namespace s3d {

class Device
{
public:
	/// The context of this process; created on first use on device $S3D_DEVICE / $LOCAL_RANK / 0.
	/// Throws std::runtime_error when no GPU is usable -- there is no CPU fallback.
	static Device& get();

	/// Explicit initialisation (before first use) for a z-slab run: one process per GPU.
	/// nccl_id: S3D_COMM_ID_BYTES from s3d_comm_unique_id() of rank 0, identical on all ranks.
	static void initialize(int device, int rank, int nranks, const void *nccl_id);
	/// torchrun-style bootstrap from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT with a file rendezvous
	/// for the NCCL id (single node).  No-op when WORLD_SIZE is unset or 1.
	static void initialize_from_env();
	static void shutdown();

	s3d_ctx *ctx() const { return ctx_; }
	int rank() const { return rank_; }
	int nranks() const { return nranks_; }

	/// Throws std::runtime_error carrying the library's message when rc != S3D_OK.
	void check(int rc, const char *what) const;
	void sync() const;

	/// n zero-initialised device doubles / ints, freed with release()
	double *scalars(int n) const;
	int *ints(int n) const;
	void release(void *p) const;

private:
	Device() = default;
	s3d_ctx *ctx_ = nullptr;
	int rank_ = 0, nranks_ = 1;
};

/// Device-resident array with a lazily synchronised host copy in the REFERENCE's host layout.
/// The reference exposes `SVec::vals` / `SMat::vals[s]` as public std::vectors that callers index
/// directly (pdebase.cpp:65,143; s3d_jacobi.cpp:25; the tests).  Here the data lives in HBM; the first
/// host access downloads it, a non-const access marks the device copy stale, and the next kernel that
/// needs it uploads it again.  Host code that never touches vals never causes a transfer.
class HostMirror
{
public:
	enum Kind { VECTOR, STENCIL };

	HostMirror() = default;
	void bind(Kind kind, const s3d_grid *grid, double *dev, int stencil, size_t host_len);
	void unbind();

	size_t size() const { return host_len_; }
	sreal &operator[](size_t i) { host_rw(); return host_[i]; }
	const sreal &operator[](size_t i) const { host_ro(); return host_[i]; }
	sreal *data() { host_rw(); return host_.data(); }
	const sreal *data() const { host_ro(); return host_.data(); }
	/// replace the whole content from a host array in the reference layout
	void assign(const sreal *src);
	/// copy the whole content into a host array in the reference layout
	void copy_to(sreal *dst) const;

	// device side (used by the host classes, not by callers)
	void device_will_read() const;     ///< upload pending host writes
	void device_wrote() const;         ///< the host copy is stale now
	bool bound() const { return dev_ != nullptr; }

private:
	void host_ro() const;
	void host_rw();
	Kind kind_ = VECTOR;
	const s3d_grid *grid_ = nullptr;
	double *dev_ = nullptr;
	int stencil_ = 0;
	size_t host_len_ = 0;
	mutable std::vector<sreal> host_;
	mutable bool host_valid_ = false;   // host_ holds the current data
	mutable bool dev_valid_ = true;     // the device buffer holds the current data
};

}  // namespace s3d

#endif
