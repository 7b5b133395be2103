// TEST INFRASTRUCTURE ONLY (oracle/_ref build) -- not part of the product path.
// Minimal stand-in for core/include/gnuradio-4.0/Message.hpp: the reference FFT/filter sources only need
// gr::exception(message, source_location) (Logger.hpp:140-150) and std::format support for their enums and
// for the radix-plan std::array (the reference gets those from meta/formatter.hpp, which needs GCC >= 14).
#ifndef GR4B200_ORACLE_SHIM_MESSAGE_HPP
#define GR4B200_ORACLE_SHIM_MESSAGE_HPP
#include <array>
#include <exception>
#include <format>
#include <source_location>
#include <string>
#include <string_view>
#include <type_traits>

namespace gr {
struct exception : std::exception {
    std::string          message;
    std::source_location sourceLocation;
    exception(std::string_view msg = "unknown exception", std::source_location location = std::source_location::current()) noexcept : message(msg), sourceLocation(location) {}
    [[nodiscard]] const char* what() const noexcept override { return message.c_str(); }
};
} // namespace gr

template<typename E>
requires std::is_enum_v<E>
struct std::formatter<E, char> {
    constexpr auto parse(std::format_parse_context& ctx) { return ctx.begin(); }
    auto           format(E value, std::format_context& ctx) const { return std::format_to(ctx.out(), "{}", static_cast<long long>(value)); }
};

template<typename T, std::size_t N>
struct std::formatter<std::array<T, N>, char> {
    constexpr auto parse(std::format_parse_context& ctx) { return ctx.begin(); }
    auto           format(const std::array<T, N>& values, std::format_context& ctx) const {
        auto out = std::format_to(ctx.out(), "[");
        for (std::size_t i = 0; i < N; ++i) {
            out = i == 0 ? std::format_to(out, "{}", values[i]) : std::format_to(out, ", {}", values[i]);
        }
        return std::format_to(out, "]");
    }
};
#endif
