// Tiny stand-in for boost-ut's `"name"_test = [] { expect(...) };` spelling so that the C++ tests read like the
// reference's qa_*.cpp files (boost-ut itself is not available offline).
#pragma once
#include <cstdio>
#include <functional>
#include <source_location>
#include <string>

namespace ut {
inline int failures = 0, checks = 0;
struct Test {
    std::string name;
    void operator=(const std::function<void()>& body) const {
        const int before = failures;
        body();
        std::printf("[%s] %s\n", failures == before ? " ok " : "FAIL", name.c_str());
    }
};
inline Test operator""_test(const char* name, std::size_t) { return Test{name}; }
inline bool expect(bool condition, const char* what = "", std::source_location loc = std::source_location::current()) {
    ++checks;
    if (!condition) {
        ++failures;
        std::printf("    expect failed at %s:%u %s\n", loc.file_name(), loc.line(), what);
    }
    return condition;
}
inline int summary() {
    std::printf("%d checks, %d failures\n", checks, failures);
    return failures == 0 ? 0 : 1;
}
} // namespace ut
