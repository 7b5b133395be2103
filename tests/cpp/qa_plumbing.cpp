// Host-side plumbing of the gr4b200 host layer, no GPU needed (pytest -m "not gpu" runs it).
// Mirrors: core/test/qa_Block.cpp device-seam tests (:1309-1410), qa_ComputeDomain.cpp parse tests (:186-189),
// qa_Scheduler.cpp linear graphs, blocks/math/test/qa_Math.cpp graph test (:16-41), and BASELINE config #1
// (NullSource -> MultiplyConst -> CountingSink, 1 000 448 complex<float> samples, scheduler::Simple, host only).
#include <chrono>
#include <complex>

#include <gnuradio-4.0/Scheduler.hpp>
#include <gnuradio-4.0/filter/time_domain_filter.hpp>
#include <gnuradio-4.0/math/Math.hpp>
#include <gnuradio-4.0/testing/NullSources.hpp>

#include "mini_ut.hpp"

using namespace ut;
using cf32 = std::complex<float>;

// a user block in the reference's spelling: one processOne, settings in the reflection list, settingsChanged hook
template<typename T>
struct ScaleAndOffset : gr::Block<ScaleAndOffset<T>> {
    using gr::Block<ScaleAndOffset<T>>::Block;
    gr::PortIn<T>                   in;
    gr::PortOut<T>                  out;
    gr::Annotated<T, "scale factor"> scale  = T(1);
    T                               offset = T(0);
    int                             changes = 0;
    GR_MAKE_REFLECTABLE(ScaleAndOffset, in, out, scale, offset);
    void settingsChanged(const gr::property_map& /*old*/, const gr::property_map& updated) { changes += static_cast<int>(updated.size()); }
    [[nodiscard]] constexpr T processOne(const T& v) const noexcept { return v * scale.value + offset; }
};

// has a device body but is asked to run on an unknown backend: must keep working on the host (warn-once seam)
template<typename T>
struct HostOrDevice : gr::Block<HostOrDevice<T>> {
    using gr::Block<HostOrDevice<T>>::Block;
    gr::PortIn<T>  in;
    gr::PortOut<T> out;
    GR_MAKE_REFLECTABLE(HostOrDevice, in, out);
    [[nodiscard]] constexpr T processOne(const T& v) const noexcept { return v + T(1); }
};

int main() {
    "ComputeDomain::parse grammar"_test = [] {
        auto d = gr::ComputeDomain::parse("gpu:cuda:3");
        expect(d.kind == "gpu" && d.backend == "cuda" && d.deviceIndex == 3 && d.isCuda());
        d = gr::ComputeDomain::parse("gpu");
        expect(d.kind == "gpu" && d.backend == "sycl" && d.deviceIndex == -1 && !d.isCuda());
        d = gr::ComputeDomain::parse("gpu:cuda:x");
        expect(d.isCuda() && d.deviceIndex == -1 && d.cudaDevice() == 0);
        for (const char* alias : {"", "host", "default_cpu", "default_io", "quantum:foo"}) {
            expect(gr::ComputeDomain::parse(alias).isHost(), alias);
        }
        expect(gr::ComputeDomain::parse("fpga").backend == "none");
    };

    "qa_Math: source -> AddConst -> sink gives exact results"_test = [] {
        gr::Graph g;
        auto&     src  = g.emplaceBlock<gr::testing::VectorSource<cf32>>();
        src.values     = {{1, 0}, {2, 0}, {8, 0}, {17, 0}};
        auto& add      = g.emplaceBlock<gr::blocks::math::AddConst<cf32>>({{"value", cf32(2, 0)}});
        auto& mul      = g.emplaceBlock<gr::blocks::math::MultiplyConst<cf32>>({{"value", cf32(0, 1)}});
        auto& sink     = g.emplaceBlock<gr::testing::VectorSink<cf32>>();
        expect(g.connect<"out", "in">(src, add).has_value());
        expect(g.connect<"out", "in">(add, mul).has_value());
        expect(g.connect<"out", "in">(mul, sink).has_value());
        gr::scheduler::Simple<> sched;
        expect(sched.exchange(std::move(g)).has_value());
        expect(sched.runAndWait().has_value());
        expect(sink._samples.size() == 4);
        expect(sink._samples == std::vector<cf32>{{0, 3}, {0, 4}, {0, 10}, {0, 19}});
    };

    "settings: property_map init, staging, settingsChanged, Annotated"_test = [] {
        gr::Graph g;
        auto&     src   = g.emplaceBlock<gr::testing::CountingSource<float>>({{"n_samples_max", 10}});
        auto&     block = g.emplaceBlock<ScaleAndOffset<float>>({{"scale", 2.f}, {"offset", 1}, {"name", "my scaler"}, {"unknown_key", 5}});
        auto&     sink  = g.emplaceBlock<gr::testing::VectorSink<float>>();
        expect(g.connect<"out", "in">(src, block).has_value() && g.connect<"out", "in">(block, sink).has_value());
        gr::scheduler::Simple<> sched(std::move(g));
        expect(sched.runAndWait().has_value());
        expect(block.name == "my scaler" && block.changes == 3); // name + scale + offset; unknown keys are not applied
        expect(sink._samples.size() == 10 && sink._samples[0] == 1.f && sink._samples[9] == 19.f);
        auto settings = block.currentSettings();
        expect(settings.contains("scale") && settings.at("scale") == gr::Value(2.f) && settings.contains("compute_domain"));
    };

    "connect: errors are values, not exceptions"_test = [] {
        gr::Graph g;
        auto&     a = g.emplaceBlock<gr::testing::NullSource<float>>();
        auto&     b = g.emplaceBlock<gr::testing::NullSink<float>>();
        auto&     c = g.emplaceBlock<gr::testing::NullSink<cf32>>();
        expect(!g.connect<"nope", "in">(a, b).has_value());
        expect(!g.connect<"out", "in">(a, c).has_value()); // type size mismatch
        expect(g.connect<"out", "in">(a, b).has_value());
        expect(!g.connect<"out", "in">(a, b).has_value()); // already connected
        gr::Graph other;
        auto&     stranger = other.emplaceBlock<gr::testing::NullSink<float>>();
        expect(!g.connect<"out", "in">(a, stranger).has_value());
    };

    "device seam: unknown backend falls back to the host body once, device-only blocks refuse the host"_test = [] {
        { // reference behaviour pinned by core/test/qa_Block.cpp:1315-1343: compute_domain set, no device body => CPU
            gr::Graph g;
            auto&     src   = g.emplaceBlock<gr::testing::CountingSource<float>>({{"n_samples_max", 8}});
            auto&     block = g.emplaceBlock<HostOrDevice<float>>({{"compute_domain", "gpu:cuda:0"}});
            auto&     sink  = g.emplaceBlock<gr::testing::VectorSink<float>>();
            expect(g.connect<"out", "in">(src, block).has_value() && g.connect<"out", "in">(block, sink).has_value());
            gr::scheduler::Simple<> sched(std::move(g));
            expect(sched.runAndWait().has_value());
            expect(block.warnedDeviceFallback() && !block.runsOnDevice());
            expect(sink._samples.size() == 8 && sink._samples[7] == 8.f);
        }
        { // the accelerated blocks have no host body: a host compute_domain is an error at init, never a silent CPU run
            gr::Graph g;
            auto&     src  = g.emplaceBlock<gr::testing::NullSource<cf32>>();
            auto&     fir  = g.emplaceBlock<gr::filter::fir_filter<cf32>>({{"b", std::vector<float>{0.5f, 0.5f}}});
            auto&     sink = g.emplaceBlock<gr::testing::CountingSink<cf32>>({{"n_samples_max", 100}});
            expect(g.connect<"out", "in">(src, fir).has_value() && g.connect<"out", "in">(fir, sink).has_value());
            gr::scheduler::Simple<> sched(std::move(g));
            auto                    result = sched.runAndWait();
            expect(!result.has_value() && result.error().message.find("no host implementation") != std::string::npos);
        }
        { // host block wired straight to a device block: refused, transitions are explicit (H2D / D2H)
            gr::Graph g;
            auto&     src  = g.emplaceBlock<gr::testing::NullSource<cf32>>();
            auto&     mul  = g.emplaceBlock<gr::blocks::math::MultiplyConst<cf32>>({{"compute_domain", "gpu:cuda:0"}});
            auto&     sink = g.emplaceBlock<gr::testing::CountingSink<cf32>>({{"n_samples_max", 100}});
            expect(g.connect<"out", "in">(src, mul).has_value() && g.connect<"out", "in">(mul, sink).has_value());
            gr::scheduler::Simple<> sched(std::move(g));
            auto                    result = sched.runAndWait();
            expect(!result.has_value());
            expect(result.error().message.find("H2D") != std::string::npos || result.error().message.find("CUDA") != std::string::npos);
        }
    };

    "resampling: chunks are whole multiples of input_chunk_size"_test = [] {
        struct Pairs : gr::Block<Pairs, gr::Resampling<2, 1>> {
            using gr::Block<Pairs, gr::Resampling<2, 1>>::Block;
            gr::PortIn<float>  in;
            gr::PortOut<float> out;
            GR_MAKE_REFLECTABLE(Pairs, in, out);
            gr::work::Status processBulk(std::span<const float> input, std::span<float> output) {
                for (std::size_t i = 0; i < output.size(); ++i) {
                    output[i] = input[2 * i] + input[2 * i + 1];
                }
                return input.size() == 2 * output.size() ? gr::work::Status::OK : gr::work::Status::ERROR;
            }
        };
        gr::Graph g;
        auto&     src  = g.emplaceBlock<gr::testing::CountingSource<float>>({{"n_samples_max", 11}}); // odd: last sample stays unconsumed
        auto&     sum  = g.emplaceBlock<Pairs>();
        auto&     sink = g.emplaceBlock<gr::testing::VectorSink<float>>();
        expect(g.connect<"out", "in">(src, sum).has_value() && g.connect<"out", "in">(sum, sink).has_value());
        gr::scheduler::Simple<> sched(std::move(g));
        expect(sched.runAndWait().has_value());
        expect(sink._samples == std::vector<float>{1, 5, 9, 13, 17});
    };

    "BASELINE config #1: NullSource -> MultiplyConst -> CountingSink, 1 000 448 complex<float>, host only"_test = [] {
        constexpr gr::Size_t kSamples = 1'000'448; // core/benchmarks/bm_MergeApi.cpp:20
        double               best     = 1e9;
        for (int repeat = 0; repeat < 10; ++repeat) {
            gr::Graph g;
            auto&     src  = g.emplaceBlock<gr::testing::NullSource<cf32>>();
            auto&     mul  = g.emplaceBlock<gr::blocks::math::MultiplyConst<cf32>>({{"value", cf32(2, 0)}});
            auto&     sink = g.emplaceBlock<gr::testing::CountingSink<cf32>>({{"n_samples_max", kSamples}});
            expect(g.connect<"out", "in">(src, mul).has_value() && g.connect<"out", "in">(mul, sink).has_value());
            gr::scheduler::Simple<> sched(std::move(g));
            const auto              t0 = std::chrono::steady_clock::now();
            expect(sched.runAndWait().has_value());
            best = std::min(best, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
            expect(sink._count >= kSamples);
        }
        std::printf("    config1_plumbing_msamples_per_s=%.1f (best of 10; reference publishes 87-162 MS/s for float chains, docs/USER_API_Connecting_Blocks.md:208-209)\n", kSamples / best / 1e6);
    };

    return summary();
}
