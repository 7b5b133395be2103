// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- not part of the product path.
// Minimal stand-in for core/include/gnuradio-4.0/MemoryAllocators.hpp (needs <print>, Logger, meta
// formatter -- unavailable with g++-13). Only what the reference FFT sources use:
// gr::allocator::Aligned<T,A> (MemoryAllocators.hpp:39) and gr::allocator::isAligned (:26-34).
#ifndef GR4B200_ORACLE_SHIM_MEMORYALLOCATORS_HPP
#define GR4B200_ORACLE_SHIM_MEMORYALLOCATORS_HPP
#include <bit>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <new>
#include <vector>

#include <gnuradio-4.0/meta/CacheLineSize.hpp>

namespace gr::allocator {
template<std::size_t Align, typename T>
[[nodiscard]] inline bool isAligned(const T* p) noexcept { return reinterpret_cast<std::uintptr_t>(p) % Align == 0; }
template<typename T>
[[nodiscard]] inline bool isAligned(const T* p, std::size_t alignment) noexcept { return reinterpret_cast<std::uintptr_t>(p) % alignment == 0; }

template<typename T, std::size_t alignment = gr::kCacheLine>
struct Aligned {
    using value_type = T;
    Aligned() noexcept = default;
    template<typename U>
    Aligned(const Aligned<U, alignment>&) noexcept {}
    [[nodiscard]] T* allocate(std::size_t n) { return n == 0 ? nullptr : static_cast<T*>(::operator new(n * sizeof(T), std::align_val_t{alignment})); }
    void             deallocate(T* p, std::size_t) noexcept { ::operator delete(p, std::align_val_t{alignment}); }
    template<typename U>
    struct rebind { using other = Aligned<U, alignment>; };
    friend bool operator==(const Aligned&, const Aligned&) noexcept { return true; }
};

namespace detail {
template<typename Container, typename TOut>
struct deduce_output_allocator { using type = Aligned<TOut>; };
template<typename Container, typename TOut>
using deduce_output_allocator_t = typename deduce_output_allocator<Container, TOut>::type;
} // namespace detail
} // namespace gr::allocator
#endif
