// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- not part of the product path.
// Minimal stand-in for meta/include/gnuradio-4.0/meta/utils.hpp (needs <print> + vir-simd). Provides only the
// few names the reference's fourier/window.hpp, fourier/fft.hpp use: gr::meta::fixed_string,
// array_or_vector_type, complex_like, always_false, type_name.
#ifndef GR4B200_ORACLE_SHIM_META_UTILS_HPP
#define GR4B200_ORACLE_SHIM_META_UTILS_HPP
#include <algorithm>
#include <array>
#include <complex>
#include <cstddef>
#include <string>
#include <string_view>
#include <type_traits>
#include <typeinfo>
#include <vector>

namespace gr::meta {
template<typename CharT, std::size_t SIZE>
struct fixed_string {
    CharT _data[SIZE + 1]{};
    constexpr fixed_string(const CharT (&str)[SIZE + 1]) noexcept { std::copy_n(str, SIZE + 1, _data); }
    [[nodiscard]] constexpr operator std::basic_string_view<CharT>() const noexcept { return {_data, SIZE}; }
};
template<typename CharT, std::size_t N>
fixed_string(const CharT (&str)[N]) -> fixed_string<CharT, N - 1>;

template<typename T>
struct is_std_array : std::false_type {};
template<typename T, std::size_t N>
struct is_std_array<std::array<T, N>> : std::true_type {};
template<typename T>
struct is_std_vector : std::false_type {};
template<typename T, typename A>
struct is_std_vector<std::vector<T, A>> : std::true_type {};
template<typename T>
concept array_or_vector_type = is_std_array<std::remove_cvref_t<T>>::value || is_std_vector<std::remove_cvref_t<T>>::value;

template<typename T>
struct is_complex : std::false_type {};
template<typename T>
struct is_complex<std::complex<T>> : std::true_type {};
template<typename T>
concept complex_like = is_complex<std::remove_cvref_t<T>>::value;

template<typename...>
inline constexpr bool always_false = false;

template<typename T>
std::string type_name() { return typeid(T).name(); }
} // namespace gr::meta
#endif
