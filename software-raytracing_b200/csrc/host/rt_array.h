// rt_array.h -- std::vector whose resize() leaves plain records uninitialised.
#pragma once
#include <memory>
#include <utility>
#include <vector>

// Allocator whose construct() default-initialises: resize() of a vector of plain records does not write the new elements.
// The big per-triangle arrays of a flattened scene are sized by the serial walk and then filled by worker threads, so the
// first touch of their pages -- the kernel zeroing 2 GB of fresh memory for a 10 M-triangle scene -- is spread over the
// cores instead of sitting in one thread's memset.
template<typename T>
struct RtNoInitAllocator : std::allocator<T>
{
	template<typename U> struct rebind { typedef RtNoInitAllocator<U> other; };
	RtNoInitAllocator() = default;
	template<typename U> RtNoInitAllocator(const RtNoInitAllocator<U>&) {}
	template<typename U> void construct(U* p) { ::new (static_cast<void*>(p)) U; }
	template<typename U, typename... Args> void construct(U* p, Args&&... args) { ::new (static_cast<void*>(p)) U(std::forward<Args>(args)...); }
};
template<typename T> using RtArray = std::vector<T, RtNoInitAllocator<T>>;
