/* TEST INFRASTRUCTURE ONLY (oracle build): the one Boost template the reference uses. */
#ifndef S3D_ORACLE_ALIGNED_ALLOCATOR_STUB_H
#define S3D_ORACLE_ALIGNED_ALLOCATOR_STUB_H
#include <cstddef>
#include <cstdlib>
#include <new>
namespace boost { namespace alignment {
template <class T, std::size_t Align>
struct aligned_allocator {
	typedef T value_type;
	aligned_allocator() noexcept {}
	template <class U> aligned_allocator(const aligned_allocator<U, Align> &) noexcept {}
	template <class U> struct rebind { typedef aligned_allocator<U, Align> other; };
	T *allocate(std::size_t n)
	{
		void *p = nullptr;
		if (n == 0) n = 1;
		if (posix_memalign(&p, Align < sizeof(void *) ? sizeof(void *) : Align, n * sizeof(T)))
			throw std::bad_alloc();
		return static_cast<T *>(p);
	}
	void deallocate(T *p, std::size_t) noexcept { std::free(p); }
	template <class U> bool operator==(const aligned_allocator<U, Align> &) const noexcept { return true; }
	template <class U> bool operator!=(const aligned_allocator<U, Align> &) const noexcept { return false; }
};
}}
#endif
