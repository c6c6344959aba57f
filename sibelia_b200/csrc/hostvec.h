// Host vectors whose resize() leaves new elements uninitialised: the per-element arrays of a simplification stage are
// filled by parallel_ranges() right after sizing, and a value-initialising resize would first touch every page from one
// thread (at 500 Mbases that alone took 7 s).  push_back / assign / swap behave as usual.
// Large blocks ask for transparent huge pages (MADV_HUGEPAGE; a no-op where THP is off): the ordered commit of a stage
// jumps between loci of a 15 GB working set (29 B per element at 500 Mbases), one TLB miss per 4 KB page it touches.
#pragma once
#include <sys/mman.h>

#include <cstdint>
#include <memory>
#include <utility>
#include <vector>

namespace sibgpu {

template<class T>
struct NoInitAlloc : std::allocator<T> {
	template<class U> struct rebind { typedef NoInitAlloc<U> other; };
	NoInitAlloc() = default;
	template<class U> NoInitAlloc(const NoInitAlloc<U>&) {}
	T *allocate(std::size_t n)
	{
		T *p = std::allocator<T>::allocate(n);
		const std::size_t bytes = n * sizeof(T);
		if(bytes >= (std::size_t(8) << 20))
		{
			const std::uintptr_t a = (reinterpret_cast<std::uintptr_t>(p) + 4095) & ~std::uintptr_t(4095);
			const std::uintptr_t e = (reinterpret_cast<std::uintptr_t>(p) + bytes) & ~std::uintptr_t(4095);
			if(e > a) madvise(reinterpret_cast<void*>(a), e - a, MADV_HUGEPAGE);
		}
		return p;
	}
	template<class U, class... A>
	void construct(U *p, A&&... a)
	{
		if constexpr(sizeof...(A) == 0) ::new(static_cast<void*>(p)) U;          // default-init: no zeroing for PODs
		else ::new(static_cast<void*>(p)) U(std::forward<A>(a)...);
	}
};

typedef std::vector<char, NoInitAlloc<char> > HostChars;
typedef std::vector<uint32_t, NoInitAlloc<uint32_t> > HostU32;
typedef std::vector<int32_t, NoInitAlloc<int32_t> > HostI32;

} // namespace sibgpu
