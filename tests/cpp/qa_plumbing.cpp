// Host-side plumbing of the gr4b200 host layer, no GPU needed (pytest -m "not gpu" runs it).
// Mirrors: core/test/qa_Block.cpp device-seam tests (:1309-1410), qa_ComputeDomain.cpp parse tests (:186-189),
// qa_Scheduler.cpp linear graphs, blocks/math/test/qa_Math.cpp graph test (:16-41), and BASELINE config #1
// (NullSource -> MultiplyConst -> CountingSink, 1 000 448 complex<float> samples, scheduler::Simple, host only).
#include <chrono>
#include <complex>

#include <gnuradio-4.0/Scheduler.hpp>
#include <gnuradio-4.0/basic/ConverterBlocks.hpp>
#include <gnuradio-4.0/filter/time_domain_filter.hpp>
#include <gnuradio-4.0/math/Math.hpp>
#include <gnuradio-4.0/testing/NullSources.hpp>
#include <gnuradio-4.0/testing/TagMonitors.hpp>

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

    "qa_Math: Add / Multiply with n_inputs ports fold left to right (Math.hpp:73-108, qa_Math.cpp:93-151)"_test = [] {
        gr::Graph g;
        auto&     a   = g.emplaceBlock<gr::testing::VectorSource<float>>();
        auto&     b   = g.emplaceBlock<gr::testing::VectorSource<float>>();
        auto&     c   = g.emplaceBlock<gr::testing::VectorSource<float>>();
        a.values      = {1, 2, 3, 4, 5};
        b.values      = {10, 20, 30, 40, 50};
        c.values      = {100, 200, 300, 400, 500};
        auto& add     = g.emplaceBlock<gr::blocks::math::Add<float>>({{"n_inputs", 3}});
        auto& sink    = g.emplaceBlock<gr::testing::VectorSink<float>>();
        expect(g.connect(a, "out", add, "in#0").has_value() && g.connect(b, "out", add, "in#1").has_value() && g.connect(c, "out", add, "in#2").has_value());
        expect(!g.connect(c, "out", add, "in#3").has_value(), "there is no fourth input");
        expect(g.connect<"out", "in">(add, sink).has_value());
        gr::scheduler::Simple<> sched(std::move(g));
        expect(sched.runAndWait().has_value());
        expect(sink._samples == std::vector<float>{111, 222, 333, 444, 555});
    };

    "one output feeds two inputs: every reader sees the whole stream at its own pace (CircularBuffer is SPMC)"_test = [] {
        struct Pairs : gr::Block<Pairs, gr::Resampling<2, 1>> {
            using gr::Block<Pairs, gr::Resampling<2, 1>>::Block;
            gr::PortIn<float>  in;
            gr::PortOut<float> out;
            GR_MAKE_REFLECTABLE(Pairs, in, out);
            gr::work::Status processBulk(std::span<const float> input, std::span<float> output) {
                for (std::size_t i = 0; i < output.size(); ++i) {
                    output[i] = input[2 * i] + input[2 * i + 1];
                }
                return gr::work::Status::OK;
            }
        };
        constexpr gr::Size_t kSamples = 300'000; // several turns of the shared 65536-item edge
        gr::Graph g;
        auto&     src   = g.emplaceBlock<gr::testing::TagSource<float>>({{"n_samples_max", kSamples}, {"sample_rate", 48'000.f}});
        auto&     sums  = g.emplaceBlock<Pairs>();
        auto&     sinkA = g.emplaceBlock<gr::testing::TagSink<float>>();
        auto&     sinkB = g.emplaceBlock<gr::testing::TagSink<float>>();
        src._tags       = {gr::Tag{100'000, {{"custom", 1}}}};
        expect(g.connect<"out", "in">(src, sinkA).has_value() && g.connect<"out", "in">(src, sums).has_value() && g.connect<"out", "in">(sums, sinkB).has_value());
        expect(!g.connect<"out", "in">(sums, sinkA).has_value(), "an input has one source");
        gr::scheduler::Simple<> sched(std::move(g));
        expect(sched.runAndWait().has_value());
        expect(sinkA._samples.size() == kSamples && sinkB._samples.size() == kSamples / 2);
        bool ok = true;
        for (std::size_t i = 0; ok && i < kSamples / 2; ++i) {
            ok = sinkA._samples[2 * i] == static_cast<float>(2 * i) && sinkB._samples[i] == static_cast<float>(4 * i + 1);
        }
        expect(ok, "both readers got every sample");
        expect(sinkA.sample_rate == 48'000.f && sinkB.sample_rate == 24'000.f, "each branch sees its own rate");
        auto tagAt = [](const auto& sink, std::size_t index) {
            for (const auto& t : sink._tags) {
                if (t.map.contains("custom") && t.index == index) {
                    return true;
                }
            }
            return false;
        };
        expect(tagAt(sinkA, 100'000) && tagAt(sinkB, 50'000), "the mid-stream tag reaches both readers exactly once, at its sample");
        std::size_t customA = 0;
        for (const auto& t : sinkA._tags) {
            customA += t.map.contains("custom") ? 1 : 0;
        }
        expect(customA == 1);
    };

    "tags: sample_rate is rescaled by a decimating block, tags ride on chunk starts, settings follow tags"_test = [] {
        // blocks/filter/test/qa_filter.cpp:267-293 ("Decimator - Low-pass Filter Test") with a host decimator
        struct KeepEveryNth : gr::Block<KeepEveryNth, gr::Resampling<1, 1, false>> {
            using gr::Block<KeepEveryNth, gr::Resampling<1, 1, false>>::Block;
            gr::PortIn<float>  in;
            gr::PortOut<float> out;
            gr::Size_t         decim = 1;
            GR_MAKE_REFLECTABLE(KeepEveryNth, in, out, decim);
            void             settingsChanged(const gr::property_map&, const gr::property_map&) { this->input_chunk_size = decim; }
            gr::work::Status processBulk(std::span<const float> input, std::span<float> output) {
                for (std::size_t i = 0; i < output.size(); ++i) {
                    output[i] = input[i * decim];
                }
                return gr::work::Status::OK;
            }
        };
        constexpr float      kInputRate = 10'000.f;
        constexpr gr::Size_t kDecim = 10, kSamples = 200'000; // several turns of the 65536-item edge: its capacity is rounded to the chunk size
        gr::Graph g;
        auto&     source = g.emplaceBlock<gr::testing::TagSource<float>>({{"sample_rate", kInputRate}, {"n_samples_max", kSamples}});
        auto&     decim  = g.emplaceBlock<KeepEveryNth>({{"decim", kDecim}});
        auto&     sink   = g.emplaceBlock<gr::testing::TagSink<float>>({{"n_samples_expected", kSamples / kDecim}});
        source._tags     = {gr::Tag{40, {{"gr:trigger_name", "mark"}}}, gr::Tag{45, {{"custom", 7}}}};
        expect(g.connect<"out", "in">(source, decim).has_value() && g.connect<"out", "in">(decim, sink).has_value());
        gr::scheduler::Simple<> sched(std::move(g));
        expect(sched.runAndWait().has_value());
        expect(decim.input_chunk_size == kDecim && decim.output_chunk_size == 1);
        expect(sink._nSamplesProduced == kSamples / kDecim);
        expect(sink.sample_rate == kInputRate / static_cast<float>(kDecim), "rate seen downstream");
        expect(source.sample_rate == kInputRate, "the source keeps its own rate");
        // first tag: the source's sample_rate on sample 0; the tag on input sample 40 arrives with output sample 4; the one
        // on input sample 45 sits inside a decimation chunk and moves to that chunk's first output (sample 4 as well, or
        // 5 when the chunking split there): never later than its sample, never in the middle of a chunk
        expect(!sink._tags.empty() && sink._tags.front().index == 0 && sink._tags.front().map.contains("sample_rate"));
        bool sawMark = false, sawCustom = false;
        for (const auto& t : sink._tags) {
            if (t.map.contains("gr:trigger_name")) {
                sawMark = t.index == 4;
            }
            if (t.map.contains("custom")) {
                sawCustom = t.index == 4;
            }
        }
        expect(sawMark && sawCustom, "tags keep their (decimated) positions");
        std::vector<float> want;
        for (gr::Size_t i = 0; i < kSamples; i += kDecim) {
            want.push_back(static_cast<float>(i));
        }
        expect(sink._samples == want);
    };

    "schedulers: Simple, BreadthFirst and DepthFirst visit the same graph in their own order (Scheduler.hpp:1944-2125)"_test = [] {
        using gr::blocks::math::AddConst;
        using gr::blocks::math::MultiplyConst;
        auto build = [](gr::Graph& g) { // S feeds A and B; A -> C -> sinkC, B -> D -> sinkD; emplaced in a scrambled order
            auto& sinkD = g.emplaceBlock<gr::testing::VectorSink<float>>({{"name", "sinkD"}});
            auto& c     = g.emplaceBlock<AddConst<float>>({{"name", "C"}, {"value", 1.f}});
            auto& s     = g.emplaceBlock<gr::testing::VectorSource<float>>({{"name", "S"}});
            auto& b     = g.emplaceBlock<MultiplyConst<float>>({{"name", "B"}, {"value", 3.f}});
            auto& sinkC = g.emplaceBlock<gr::testing::VectorSink<float>>({{"name", "sinkC"}});
            auto& a     = g.emplaceBlock<MultiplyConst<float>>({{"name", "A"}, {"value", 2.f}});
            auto& d     = g.emplaceBlock<AddConst<float>>({{"name", "D"}, {"value", -1.f}});
            s.values.resize(100'000);
            for (std::size_t i = 0; i < s.values.size(); ++i) {
                s.values[i] = static_cast<float>(i % 1000);
            }
            bool ok = g.connect<"out", "in">(s, a).has_value() && g.connect<"out", "in">(s, b).has_value();
            ok      = ok && g.connect<"out", "in">(b, d).has_value() && g.connect<"out", "in">(a, c).has_value();
            ok      = ok && g.connect<"out", "in">(d, sinkD).has_value() && g.connect<"out", "in">(c, sinkC).has_value();
            expect(ok);
            return std::pair<gr::testing::VectorSink<float>*, gr::testing::VectorSink<float>*>{&sinkC, &sinkD};
        };
        auto check = [&](auto&& sched, std::vector<std::string_view> wantOrder) {
            gr::Graph g;
            auto [sinkC, sinkD] = build(g);
            expect(sched.exchange(std::move(g)).has_value());
            expect(sched.runAndWait().has_value());
            std::vector<std::string_view> order;
            for (const gr::BlockModel* block : sched.executionOrder()) {
                order.push_back(block->name());
            }
            expect(order == wantOrder, "execution order");
            bool ok = sinkC->_samples.size() == 100'000 && sinkD->_samples.size() == 100'000;
            for (std::size_t i = 0; ok && i < 100'000; ++i) {
                const float v = static_cast<float>(i % 1000);
                ok            = sinkC->_samples[i] == v * 2.f + 1.f && sinkD->_samples[i] == v * 3.f - 1.f;
            }
            expect(ok, "same samples whatever the order");
        };
        check(gr::scheduler::Simple<>{}, {"sinkD", "C", "S", "B", "sinkC", "A", "D"});
        check(gr::scheduler::BreadthFirst<>{}, {"S", "A", "B", "C", "D", "sinkC", "sinkD"});
        check(gr::scheduler::DepthFirst<>{}, {"S", "A", "C", "sinkC", "B", "D", "sinkD"});
    };

    "qa_Converter: complex <-> interleaved for float, int16 and int8 items (qa_Converter.cpp:242-268)"_test = [] {
        auto roundTrip = []<typename R>(R) {
            using namespace gr::blocks::type::converter;
            gr::Graph g;
            auto&     src = g.emplaceBlock<gr::testing::VectorSource<cf32>>();
            src.values    = {{1, 2}, {3, 4}, {5, 6}};
            auto& toItems = g.emplaceBlock<ComplexToInterleaved<cf32, R>>();
            auto& items   = g.emplaceBlock<gr::testing::VectorSink<R>>();
            auto& back    = g.emplaceBlock<InterleavedToComplex<R, cf32>>();
            auto& sink    = g.emplaceBlock<gr::testing::VectorSink<cf32>>();
            expect(g.connect<"out", "in">(src, toItems).has_value() && g.connect<"interleaved", "in">(toItems, items).has_value());
            expect(g.connect<"interleaved", "interleaved">(toItems, back).has_value() && g.connect<"out", "in">(back, sink).has_value());
            gr::scheduler::Simple<> sched(std::move(g));
            expect(sched.runAndWait().has_value());
            expect(items._samples == std::vector<R>{R(1), R(2), R(3), R(4), R(5), R(6)}, "two items per complex sample, re first");
            expect(sink._samples == std::vector<cf32>{{1, 2}, {3, 4}, {5, 6}}, "and back");
        };
        roundTrip(float{});
        roundTrip(std::int16_t{});
        roundTrip(std::int8_t{});
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

    "Stride<>: overlap, skip and default, the reference's test vectors (qa_Block.cpp:587-650, 757-777)"_test = [] {
        struct Recorder : gr::Block<Recorder, gr::Resampling<1, 1, false>, gr::Stride<0, false>> { // the reference test's Resampler<int>
            using gr::Block<Recorder, gr::Resampling<1, 1, false>, gr::Stride<0, false>>::Block;
            gr::PortIn<int>  in;
            gr::PortOut<int> out;
            GR_MAKE_REFLECTABLE(Recorder, in, out);
            std::size_t      counter = 0, lastIn = 0, lastOut = 0, totalIn = 0, totalOut = 0;
            std::vector<int> seen;
            gr::work::Status processBulk(std::span<const int> input, std::span<int> output) {
                ++counter;
                lastIn = input.size(), lastOut = output.size();
                totalIn += input.size(), totalOut += output.size();
                seen.insert(seen.end(), input.begin(), input.end());
                std::fill(output.begin(), output.end(), 0);
                return gr::work::Status::OK;
            }
        };
        struct Case {
            gr::Size_t       n, outChunk, inChunk, stride;
            std::size_t      expIn, expOut, expCounter, expTotalIn, expTotalOut;
            std::vector<int> expSeen;
        };
        const std::vector<Case> cases{
            {1000, 50, 50, 100, 50, 50, 10, 500, 500, {}},
            {1000, 50, 50, 133, 50, 50, 8, 400, 400, {}},
            {1000, 100, 100, 50, 100, 100, 19, 1900, 1900, {}},
            {1000, 100, 100, 33, 100, 100, 28, 2800, 2800, {}},
            {1000, 50, 100, 50, 100, 50, 19, 1900, 950, {}},
            {1000, 24, 48, 50, 48, 24, 20, 960, 480, {}},
            {15, 5, 5, 3, 5, 5, 4, 20, 20, {0, 1, 2, 3, 4, 3, 4, 5, 6, 7, 6, 7, 8, 9, 10, 9, 10, 11, 12, 13}},
            {15, 3, 3, 5, 3, 3, 3, 9, 9, {0, 1, 2, 5, 6, 7, 10, 11, 12}},
            {1000000, 100, 100, 250000, 100, 100, 4, 400, 400, {}},
            {1000000, 100, 100, 249900, 100, 100, 5, 500, 500, {}},
        };
        for (const Case& c : cases) {
            gr::Graph g;
            auto&     src  = g.emplaceBlock<gr::testing::CountingSource<int>>({{"n_samples_max", c.n}});
            auto&     rec  = g.emplaceBlock<Recorder>({{"output_chunk_size", c.outChunk}, {"input_chunk_size", c.inChunk}, {"stride", c.stride}});
            auto&     sink = g.emplaceBlock<gr::testing::NullSink<int>>();
            // a small odd ring in front of the block: chunks run across its end again and again (the wrap-safe span path)
            expect(g.connect<"out", "in">(src, rec, {.minBufferSize = c.n >= 100000 ? 65536u : 3 * c.inChunk + 1}).has_value() && g.connect<"out", "in">(rec, sink).has_value());
            gr::scheduler::Simple<> sched(std::move(g));
            expect(sched.runAndWait().has_value());
            const bool ok = rec.counter == c.expCounter && rec.lastIn == c.expIn && rec.lastOut == c.expOut && rec.totalIn == c.expTotalIn && rec.totalOut == c.expTotalOut && (c.expSeen.empty() || rec.seen == c.expSeen);
            if (!ok) {
                std::printf("    stride case n=%u in=%u out=%u stride=%u: counter %zu (want %zu) totalIn %zu (want %zu)\n", c.n, c.inChunk, c.outChunk, c.stride, rec.counter, c.expCounter, rec.totalIn, c.expTotalIn);
            }
            expect(ok, "stride case");
        }
        // stride == input_chunk_size and stride == 0 are the plain resampling path: every sample is seen exactly once
        for (gr::Size_t stride : {gr::Size_t{0}, gr::Size_t{50}}) {
            gr::Graph g;
            auto&     src  = g.emplaceBlock<gr::testing::CountingSource<int>>({{"n_samples_max", gr::Size_t{1000}}});
            auto&     rec  = g.emplaceBlock<Recorder>({{"output_chunk_size", gr::Size_t{25}}, {"input_chunk_size", gr::Size_t{50}}, {"stride", stride}});
            auto&     sink = g.emplaceBlock<gr::testing::NullSink<int>>();
            expect(g.connect<"out", "in">(src, rec).has_value() && g.connect<"out", "in">(rec, sink).has_value());
            gr::scheduler::Simple<> sched(std::move(g));
            expect(sched.runAndWait().has_value());
            expect(rec.totalIn == 1000 && rec.totalOut == 500);
        }
    };

    "multiThreaded: blocks on several launcher threads, edges as hand-off queues (qa_Scheduler.cpp *_multi_threaded cases)"_test = [] {
        // CountingSource -> MultiplyConst -> AddConst -> VectorSink with every block on its own thread: the values must arrive
        // complete and in order, through edges of 4096 items (many wrap-arounds and full / empty conditions)
        constexpr gr::Size_t kSamples = 300'000;
        for (std::size_t threads : {std::size_t{1}, std::size_t{2}, std::size_t{4}}) {
            gr::Graph g;
            auto&     src  = g.emplaceBlock<gr::testing::CountingSource<float>>({{"n_samples_max", kSamples}});
            auto&     mul  = g.emplaceBlock<gr::blocks::math::MultiplyConst<float>>({{"value", 2.f}});
            auto&     add  = g.emplaceBlock<gr::blocks::math::AddConst<float>>({{"value", 1.f}});
            auto&     sink = g.emplaceBlock<gr::testing::VectorSink<float>>();
            expect(g.connect<"out", "in">(src, mul, {.minBufferSize = 4096}).has_value() && g.connect<"out", "in">(mul, add, {.minBufferSize = 4096}).has_value() && g.connect<"out", "in">(add, sink, {.minBufferSize = 4096}).has_value());
            gr::scheduler::Simple<gr::scheduler::ExecutionPolicy::multiThreaded> sched(std::move(g));
            sched.host_threads = threads;
            const auto result  = sched.runAndWait();
            expect(result.has_value(), result ? "" : result.error().message.c_str());
            expect(sink._samples.size() == kSamples);
            bool ordered = sink._samples.size() == kSamples;
            for (std::size_t i = 0; ordered && i < kSamples; ++i) {
                ordered = sink._samples[i] == 2.f * static_cast<float>(i) + 1.f;
            }
            expect(ordered, "values complete and in order");
        }
        // the end of a stream: the source publishes its last span and is done while the next block's thread is in the middle
        // of work() -- many short runs, every one must deliver every sample (a block that reads the item count first and the
        // "producer done" flag second finishes with the last span unread once in a few hundred runs)
        std::size_t shortRuns = 0;
        for (int run = 0; run < 1500; ++run) {
            constexpr gr::Size_t kShort = 3'000;
            gr::Graph            g;
            auto&                src  = g.emplaceBlock<gr::testing::CountingSource<float>>({{"n_samples_max", kShort}});
            auto&                mul  = g.emplaceBlock<gr::blocks::math::MultiplyConst<float>>({{"value", 2.f}});
            auto&                sink = g.emplaceBlock<gr::testing::VectorSink<float>>();
            expect(g.connect<"out", "in">(src, mul, {.minBufferSize = 512}).has_value() && g.connect<"out", "in">(mul, sink, {.minBufferSize = 512}).has_value());
            gr::scheduler::Simple<gr::scheduler::ExecutionPolicy::multiThreaded> sched(std::move(g));
            sched.host_threads = 3;
            shortRuns += sched.runAndWait().has_value() && sink._samples.size() == kShort ? 0 : 1;
        }
        expect(shortRuns == 0, "every short multi-threaded run delivers every sample");
        // a sink that stops the graph (CountingSink) ends every thread
        gr::Graph g;
        auto&     src  = g.emplaceBlock<gr::testing::NullSource<cf32>>();
        auto&     mul  = g.emplaceBlock<gr::blocks::math::MultiplyConst<cf32>>({{"value", cf32(2, 0)}});
        auto&     sink = g.emplaceBlock<gr::testing::CountingSink<cf32>>({{"n_samples_max", gr::Size_t{1'000'448}}});
        expect(g.connect<"out", "in">(src, mul).has_value() && g.connect<"out", "in">(mul, sink).has_value());
        gr::scheduler::Simple<gr::scheduler::ExecutionPolicy::multiThreaded> sched(std::move(g));
        sched.host_threads = 3;
        expect(sched.runAndWait().has_value());
        expect(sink._count >= 1'000'448);
    };

    return summary();
}
