// gr4b200 host layer -- compile-time helpers: fixed_string (for connect<"out","in">), member reflection.
// Our own minimal take on what the reference spreads over meta/ (meta/utils.hpp fixed_string, reflection.hpp
// GR_MAKE_REFLECTABLE); only the spelling a block author sees is kept.
#pragma once

#include <algorithm>
#include <array>
#include <cstddef>
#include <string_view>
#include <tuple>
#include <utility>

namespace gr::meta {

template<std::size_t N>
struct fixed_string {
    char value[N]{};
    constexpr fixed_string(const char (&text)[N]) noexcept { std::copy_n(text, N, value); }
    [[nodiscard]] constexpr std::string_view view() const noexcept { return {value, N - 1}; }
    [[nodiscard]] constexpr operator std::string_view() const noexcept { return view(); }
};

// "a, b, c" -> {"a","b","c"} at compile time
template<std::size_t Count>
constexpr std::array<std::string_view, Count> splitNames(std::string_view list) {
    std::array<std::string_view, Count> names{};
    std::size_t                         index = 0, begin = 0;
    for (std::size_t i = 0; i <= list.size() && index < Count; ++i) {
        if (i == list.size() || list[i] == ',') {
            std::size_t b = begin, e = i;
            while (b < e && (list[b] == ' ' || list[b] == '\t' || list[b] == '\n')) {
                ++b;
            }
            while (e > b && (list[e - 1] == ' ' || list[e - 1] == '\t' || list[e - 1] == '\n')) {
                --e;
            }
            names[index++] = list.substr(b, e - b);
            begin          = i + 1;
        }
    }
    return names;
}

template<typename... Ts>
constexpr std::size_t countArgs(const Ts&...) {
    return sizeof...(Ts);
}

} // namespace gr::meta

// Lists the ports and settings of a block, in declaration order (same call as the reference's macro):
//   GR_MAKE_REFLECTABLE(MyBlock, in, out, gain);
#define GR_MAKE_REFLECTABLE(TypeName, ...)                                                                                          \
    using gr_reflected_type = TypeName;                                                                                             \
    static constexpr std::string_view gr_type_name() { return #TypeName; }                                                          \
    auto                              gr_members() { return std::tie(__VA_ARGS__); }                                                \
    static constexpr auto             gr_member_names() {                                                                           \
        constexpr std::size_t count = std::tuple_size_v<decltype(std::declval<TypeName&>().gr_members())>;              \
        return gr::meta::splitNames<count>(#__VA_ARGS__);                                                                           \
    }
