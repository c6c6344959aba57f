// Host vectors whose resize() leaves new elements uninitialised: the per-element arrays of a simplification stage are
// filled by parallel_ranges() right after sizing, and a value-initialising resize would first touch every page from one
// thread (at 500 Mbases that alone took 7 s).  push_back / assign / swap behave as usual.
#pragma once
#include <memory>
#include <utility>
#include <vector>

namespace sibgpu {

template<class T>
struct NoInitAlloc : std::allocator<T> {
	template<class U> struct rebind { typedef NoInitAlloc<U> other; };
	NoInitAlloc() = default;
	template<class U> NoInitAlloc(const NoInitAlloc<U>&) {}
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
