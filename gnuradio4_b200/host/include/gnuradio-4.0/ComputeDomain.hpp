// gr4b200 host layer -- the per-block `compute_domain` setting: "kind[:backend[:deviceIndex]]".
// Same grammar and defaults as the reference parser (core/include/gnuradio-4.0/ComputeDomain.hpp:47-100), restated.
#pragma once

#include <charconv>
#include <string>
#include <string_view>

namespace gr {

struct ComputeDomain {
    std::string kind{"host"};
    std::string backend{"none"};
    int         deviceIndex{-1};

    [[nodiscard]] bool isHost() const noexcept { return kind == "host"; }
    [[nodiscard]] bool isCuda() const noexcept { return kind == "gpu" && backend == "cuda"; }
    [[nodiscard]] int  cudaDevice() const noexcept { return deviceIndex < 0 ? 0 : deviceIndex; }

    static ComputeDomain parse(std::string_view text) {
        ComputeDomain d;
        if (text.empty() || text == "host" || text == "default_cpu" || text == "default_io") {
            return d;
        }
        const auto first = text.find(':');
        const auto kind  = text.substr(0, first);
        if (kind != "gpu" && kind != "fpga" && kind != "tpu") {
            return d; // unknown kinds are host
        }
        d.kind    = std::string(kind);
        d.backend = kind == "gpu" ? "sycl" : "none";
        if (first == std::string_view::npos) {
            return d;
        }
        const auto rest   = text.substr(first + 1);
        const auto second = rest.find(':');
        if (const auto backend = rest.substr(0, second); !backend.empty()) {
            d.backend = std::string(backend);
        }
        if (second != std::string_view::npos) {
            const auto index = rest.substr(second + 1);
            int        parsed = -1;
            if (std::from_chars(index.data(), index.data() + index.size(), parsed).ec == std::errc{}) {
                d.deviceIndex = parsed;
            }
        }
        return d;
    }
};

} // namespace gr
