// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- not part of the product path.
//
// Stand-in for the un-vendored vir-simd v0.4.4 dependency (reference DEPENDENCIES.md:12) so that the
// reference's own FFT sources can be compiled *where they lie* under /root/reference by oracle/Makefile.
// vir-simd is a thin vehicle over std::experimental::simd (which libstdc++-13 ships); the only extras
// the reference FFT path touches are vir::stdx, vir::simd_permute and a tuple-returning split
// (call sites: algorithm/include/gnuradio-4.0/algorithm/fourier/SimdFFT.hpp:93-133).
// Everything below is written from the call-site contract, not from vir-simd's sources.
#ifndef GR4B200_ORACLE_SHIM_VIR_SIMD_H
#define GR4B200_ORACLE_SHIM_VIR_SIMD_H

#include <array>
#include <cstddef>
#include <experimental/simd>
#include <tuple>
#include <utility>

namespace vir {
namespace stdx {
using namespace std::experimental::parallelism_v2;

namespace shim_detail {
template<typename V, typename T, typename Abi, std::size_t... Part>
inline auto splitToTuple(const simd<T, Abi>& x, std::index_sequence<Part...>) {
    alignas(64) T buffer[simd<T, Abi>::size()];
    x.copy_to(buffer, element_aligned);
    return std::tuple{V(buffer + Part * V::size(), element_aligned)...};
}
} // namespace shim_detail

// SimdFFT.hpp assigns the result of split<V>() through std::tie (C++23 tuple-like assignment from
// std::array needs libstdc++-14); returning a real std::tuple keeps the call sites valid on GCC 13.
template<typename V, typename T, typename Abi>
inline auto split(const simd<T, Abi>& x) {
    static_assert(simd<T, Abi>::size() % V::size() == 0);
    return shim_detail::splitToTuple<V>(x, std::make_index_sequence<simd<T, Abi>::size() / V::size()>{});
}
} // namespace stdx

// out[i] = in[indexMap(i)], i in [0, NOut)
template<int NOut = 0, typename T, typename Abi, typename F>
inline auto simd_permute(const stdx::simd<T, Abi>& in, F indexMap) {
    constexpr std::size_t nIn  = stdx::simd<T, Abi>::size();
    constexpr std::size_t nOut = NOut == 0 ? nIn : static_cast<std::size_t>(NOut);
    alignas(64) T src[nIn];
    alignas(64) T dst[nOut];
    in.copy_to(src, stdx::element_aligned);
    for (std::size_t i = 0; i < nOut; ++i) {
        dst[i] = src[static_cast<std::size_t>(indexMap(static_cast<int>(i)))];
    }
    return stdx::simd<T, stdx::simd_abi::deduce_t<T, nOut>>(dst, stdx::element_aligned);
}
} // namespace vir

#endif
