// Process-wide GPU context and the coherent host mirror (see s3d_device.hpp).
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include "s3d_device.hpp"

namespace s3d {

static Device *g_device = nullptr;

static int env_int(const char *name, int dflt)
{
	const char *v = std::getenv(name);
	return v ? std::atoi(v) : dflt;
}

Device &Device::get()
{
	if (!g_device) {
		const int dev = env_int("S3D_DEVICE", env_int("LOCAL_RANK", 0));
		initialize(dev, 0, 1, nullptr);
	}
	return *g_device;
}

void Device::initialize(int device, int rank, int nranks, const void *nccl_id)
{
	if (g_device) throw std::runtime_error("s3d::Device is already initialised");
	Device *d = new Device();
	const int rc = s3d_ctx_create(device, nullptr, &d->ctx_);
	if (rc != S3D_OK) {
		const std::string msg = std::string("struct3d_b200: cannot create the GPU context (there is no CPU fallback): ")
		                        + s3d_error_string(nullptr);
		delete d;
		throw std::runtime_error(msg);
	}
	d->rank_ = rank;
	d->nranks_ = nranks;
	g_device = d;
	if (nranks > 1) d->check(s3d_comm_init(d->ctx_, rank, nranks, nccl_id), "s3d_comm_init");
}

void Device::initialize_from_env()
{
	const int world = env_int("WORLD_SIZE", 1);
	if (world <= 1) { get(); return; }
	const int rank = env_int("RANK", 0);
	const int local = env_int("LOCAL_RANK", rank);
	// The NCCL id travels from rank 0 to the other ranks of THIS launch through a small file.  Its name carries the launcher's pid
	// (all ranks are children of one torchrun agent), so a file left by an earlier launch on the same port can never be picked up;
	// it is created exclusively with mode 0600 (no following of planted links) and removed once every rank has joined.
	const char *tmpdir = std::getenv("TMPDIR");
	const std::string path = std::string(tmpdir && *tmpdir ? tmpdir : "/tmp") + "/s3d_nccl_id_" + std::to_string((long)getuid()) + "_"
	                         + (std::getenv("MASTER_PORT") ? std::getenv("MASTER_PORT") : "0") + "_" + std::to_string((long)getppid());
	struct Record { char magic[8]; long ppid; char id[S3D_COMM_ID_BYTES]; } rec;
	char id[S3D_COMM_ID_BYTES];
	if (rank == 0) {
		if (s3d_comm_unique_id(id) != S3D_OK) throw std::runtime_error("s3d_comm_unique_id failed");
		std::memcpy(rec.magic, "S3DNCCL", 8); rec.ppid = (long)getppid(); std::memcpy(rec.id, id, sizeof id);
		const std::string tmp = path + ".tmp";
		::unlink(tmp.c_str());
		::unlink(path.c_str());                     // same launcher pid, same port: can only be a leftover of a crashed attempt of ours
		const int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
		if (fd < 0 || ::write(fd, &rec, sizeof rec) != (ssize_t)sizeof rec) { if (fd >= 0) ::close(fd); throw std::runtime_error("cannot write the NCCL id file " + tmp); }
		::close(fd);
		if (std::rename(tmp.c_str(), path.c_str()) != 0) throw std::runtime_error("cannot publish the NCCL id file " + path);
	} else {
		for (int tries = 0;; tries++) {
			const int fd = ::open(path.c_str(), O_RDONLY | O_NOFOLLOW);
			if (fd >= 0) {
				struct stat st;
				const bool mine = ::fstat(fd, &st) == 0 && st.st_uid == getuid() && (st.st_mode & 077) == 0;
				const bool got = mine && ::read(fd, &rec, sizeof rec) == (ssize_t)sizeof rec && !std::memcmp(rec.magic, "S3DNCCL", 8) && rec.ppid == (long)getppid();
				::close(fd);
				if (got) { std::memcpy(id, rec.id, sizeof id); break; }
			}
			if (tries > 6000) throw std::runtime_error("timed out waiting for the NCCL id file " + path);
			std::this_thread::sleep_for(std::chrono::milliseconds(10));
		}
	}
	initialize(local, rank, world, id);            // ncclCommInitRank is collective: every rank has read the id when it returns
	if (rank == 0) ::unlink(path.c_str());
}

void Device::shutdown()
{
	if (!g_device) return;
	s3d_ctx_destroy(g_device->ctx_);
	delete g_device;
	g_device = nullptr;
}

void Device::check(int rc, const char *what) const
{
	if (rc == S3D_OK) return;
	throw std::runtime_error(std::string(what) + ": " + s3d_error_string(ctx_));
}

void Device::sync() const { check(s3d_ctx_sync(ctx_), "s3d_ctx_sync"); }

double *Device::scalars(int n) const
{
	double *p = nullptr;
	check(s3d_scalars_alloc(ctx_, n, &p), "s3d_scalars_alloc");
	return p;
}

int *Device::ints(int n) const
{
	int *p = nullptr;
	check(s3d_ints_alloc(ctx_, n, &p), "s3d_ints_alloc");
	return p;
}

void Device::release(void *p) const { check(s3d_free(ctx_, p), "s3d_free"); }

// ------------------------------------------------------------------ HostMirror

void HostMirror::bind(Kind kind, const s3d_grid *grid, double *dev, int stencil, size_t host_len)
{
	kind_ = kind; grid_ = grid; dev_ = dev; stencil_ = stencil; host_len_ = host_len;
	host_.clear(); host_.shrink_to_fit();
	host_valid_ = false; dev_valid_ = true;
}

void HostMirror::unbind()
{
	dev_ = nullptr; grid_ = nullptr; host_len_ = 0;
	host_.clear(); host_.shrink_to_fit();
	host_valid_ = false; dev_valid_ = true;
}

void HostMirror::host_ro() const
{
	if (host_valid_) return;
	if (!dev_) throw std::runtime_error("HostMirror: access to an unallocated vector/matrix");
	host_.resize(host_len_);
	const Device &d = Device::get();
	if (kind_ == VECTOR) d.check(s3d_vec_download(d.ctx(), grid_, dev_, host_.data()), "s3d_vec_download");
	else d.check(s3d_mat_download(d.ctx(), grid_, dev_, stencil_, host_.data()), "s3d_mat_download");
	host_valid_ = true;
}

void HostMirror::host_rw()
{
	host_ro();
	dev_valid_ = false;
}

void HostMirror::device_will_read() const
{
	if (dev_valid_) return;
	const Device &d = Device::get();
	if (kind_ == VECTOR) d.check(s3d_vec_upload(d.ctx(), grid_, dev_, host_.data()), "s3d_vec_upload");
	else d.check(s3d_mat_upload(d.ctx(), grid_, dev_, stencil_, host_.data()), "s3d_mat_upload");
	d.sync();   // host_ may be modified again right after this returns
	dev_valid_ = true;
}

void HostMirror::device_wrote() const
{
	host_valid_ = false;
}

void HostMirror::assign(const sreal *src)
{
	host_.assign(src, src + host_len_);
	host_valid_ = true;
	dev_valid_ = false;
	device_will_read();
}

void HostMirror::copy_to(sreal *dst) const
{
	host_ro();
	std::copy(host_.begin(), host_.end(), dst);
}

}  // namespace s3d
