// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- not part of the product path.
// Minimal stand-in for meta/include/gnuradio-4.0/meta/UncertainValue.hpp: the reference's FilterTool.hpp only
// needs the type to exist so that its UncertainValue specialisations parse; the oracle/_ref harness never
// instantiates them (the hot path is float / double only).
#ifndef GR4B200_ORACLE_SHIM_UNCERTAINVALUE_HPP
#define GR4B200_ORACLE_SHIM_UNCERTAINVALUE_HPP
#include <type_traits>

#include <gnuradio-4.0/meta/utils.hpp>

#include <cmath>

namespace gr {
template<typename T>
concept arithmetic_or_complex_like = std::is_arithmetic_v<T> || meta::complex_like<T>;

namespace math { // the reference wraps <cmath> for UncertainValue arguments; plain floating point only here
template<typename T>
inline T cos(T x) noexcept { return std::cos(x); }
template<typename T>
inline T sin(T x) noexcept { return std::sin(x); }
} // namespace math

template<typename T>
struct UncertainValue {
    using value_type = T;
    T value{};
    T uncertainty{};
    constexpr UncertainValue() = default;
    constexpr UncertainValue(T v, T u = T{}) : value(v), uncertainty(u) {}
};

template<typename T>
struct is_uncertain_value : std::false_type {};
template<typename T>
struct is_uncertain_value<UncertainValue<T>> : std::true_type {};
template<typename T>
concept UncertainValueLike = is_uncertain_value<std::remove_cvref_t<T>>::value;

template<typename T>
constexpr auto value(const T& v) noexcept {
    if constexpr (UncertainValueLike<T>) {
        return v.value;
    } else {
        return v;
    }
}
template<typename T>
constexpr auto uncertainty(const T& v) noexcept {
    if constexpr (UncertainValueLike<T>) {
        return v.uncertainty;
    } else {
        return T{};
    }
}

namespace meta {
template<typename T>
struct fundamental_base_value_type { using type = T; };
template<typename T>
struct fundamental_base_value_type<UncertainValue<T>> { using type = T; };
template<typename T>
using fundamental_base_value_type_t = typename fundamental_base_value_type<T>::type;
} // namespace meta
} // namespace gr
#endif
