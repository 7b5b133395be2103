// gr4b200 host layer -- property_map: string -> Value, the shape in which settings reach a block
// (reference: gr::property_map / gr::Value, core/include/gnuradio-4.0/Value.hpp; here a small closed variant).
#pragma once

#include <complex>
#include <cstdint>
#include <map>
#include <string>
#include <type_traits>
#include <variant>
#include <vector>

namespace gr {

using Size_t = std::uint32_t;

struct Value {
    using Storage = std::variant<std::monostate, bool, std::int64_t, std::uint64_t, double, std::complex<double>, std::string, std::vector<double>>;
    Storage data;

    Value() = default;
    Value(bool v) : data(v) {}
    template<typename T>
    requires(std::is_integral_v<T> && std::is_signed_v<T> && !std::is_same_v<T, bool>)
    Value(T v) : data(static_cast<std::int64_t>(v)) {}
    template<typename T>
    requires(std::is_integral_v<T> && std::is_unsigned_v<T> && !std::is_same_v<T, bool>)
    Value(T v) : data(static_cast<std::uint64_t>(v)) {}
    template<typename T>
    requires std::is_floating_point_v<T>
    Value(T v) : data(static_cast<double>(v)) {}
    template<typename T>
    Value(std::complex<T> v) : data(std::complex<double>(v.real(), v.imag())) {}
    Value(const char* v) : data(std::string(v)) {}
    Value(std::string v) : data(std::move(v)) {}
    template<typename T>
    requires std::is_arithmetic_v<T>
    Value(const std::vector<T>& v) : data(std::vector<double>(v.begin(), v.end())) {}

    [[nodiscard]] bool holdsNumber() const { return std::holds_alternative<bool>(data) || std::holds_alternative<std::int64_t>(data) || std::holds_alternative<std::uint64_t>(data) || std::holds_alternative<double>(data); }
    [[nodiscard]] double asDouble() const {
        if (auto* b = std::get_if<bool>(&data)) return *b ? 1.0 : 0.0;
        if (auto* i = std::get_if<std::int64_t>(&data)) return static_cast<double>(*i);
        if (auto* u = std::get_if<std::uint64_t>(&data)) return static_cast<double>(*u);
        if (auto* d = std::get_if<double>(&data)) return *d;
        return 0.0;
    }
    friend bool operator==(const Value&, const Value&) = default;
};

using property_map = std::map<std::string, Value, std::less<>>;

// assigns `value` to a block member of type T; false if the kinds do not match (reference: setting rejected)
template<typename T>
bool assignFromValue(T& member, const Value& value) {
    if constexpr (std::is_same_v<T, bool>) {
        if (!value.holdsNumber()) return false;
        member = value.asDouble() != 0.0;
    } else if constexpr (std::is_arithmetic_v<T>) {
        if (!value.holdsNumber()) return false;
        member = static_cast<T>(value.asDouble());
    } else if constexpr (std::is_enum_v<T>) {
        if (!value.holdsNumber()) return false;
        member = static_cast<T>(static_cast<std::underlying_type_t<T>>(value.asDouble()));
    } else if constexpr (std::is_same_v<T, std::string>) {
        auto* s = std::get_if<std::string>(&value.data);
        if (s == nullptr) return false;
        member = *s;
    } else if constexpr (std::is_same_v<T, std::complex<float>> || std::is_same_v<T, std::complex<double>>) {
        if (auto* c = std::get_if<std::complex<double>>(&value.data)) {
            member = T(static_cast<typename T::value_type>(c->real()), static_cast<typename T::value_type>(c->imag()));
        } else if (value.holdsNumber()) {
            member = T(static_cast<typename T::value_type>(value.asDouble()), 0);
        } else {
            return false;
        }
    } else if constexpr (requires { typename T::value_type; member.assign(std::declval<const double*>(), std::declval<const double*>()); }) {
        auto* v = std::get_if<std::vector<double>>(&value.data);
        if (v == nullptr) return false;
        member.clear();
        for (double d : *v) member.push_back(static_cast<typename T::value_type>(d));
    } else {
        return false;
    }
    return true;
}

template<typename T>
Value valueOf(const T& member) {
    if constexpr (std::is_enum_v<T>) {
        return Value(static_cast<std::int64_t>(member));
    } else if constexpr (std::is_constructible_v<Value, const T&>) {
        return Value(member);
    } else {
        return Value{};
    }
}

} // namespace gr
