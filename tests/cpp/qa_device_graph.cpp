// Device flowgraphs through the gr4b200 host layer, checked against the CPU oracle. Needs a B200 (pytest -m gpu).
// Mirrors the reference's integration-test shape: TagSource(values) -> block under test -> TagSink, compare _samples
// (blocks/math/test/qa_Math.cpp:16-41, blocks/filter/test/qa_filter.cpp:267-321, blocks/fourier/test/qa_fourier.cpp:120-150).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstring>
#include <random>

#include <gnuradio-4.0/Scheduler.hpp>
#include <gnuradio-4.0/basic/ConverterBlocks.hpp>
#include <gnuradio-4.0/cuda/Transfer.hpp>
#include <gnuradio-4.0/filter/time_domain_filter.hpp>
#include <gnuradio-4.0/fourier/fft.hpp>
#include <gnuradio-4.0/math/Math.hpp>
#include <gnuradio-4.0/math/Rotator.hpp>
#include <gnuradio-4.0/testing/NullSources.hpp>
#include <gnuradio-4.0/testing/TagMonitors.hpp>

#include "mini_ut.hpp"

using namespace ut;
using cf32 = std::complex<float>;

extern "C" { // the CPU oracle (oracle/oracle.cpp): test infrastructure
int oracle_fir_generate_f32(std::size_t nTaps, int windowType, float fc, float beta, int normaliseDc, float* out);
int oracle_window_f32(int type, std::size_t n, float beta, float* out);
int oracle_fir_cf32(const float* taps, std::size_t nTaps, const float* in, float* out, std::size_t n, float* state);
int oracle_fir_decim_cf32(const float* taps, std::size_t nTaps, std::size_t decimate, const float* in, float* out, std::size_t n, float* state);
int oracle_fft_block_cf32(const float* in, std::size_t nfft, std::size_t batch, const float* window, int db, int deg, int unwrap, float* signals, float* ranges);
int oracle_mathop_const_cf32(int op, const float* in, float* out, std::size_t n, float re, float im);
int oracle_rotator_cf32(const float* in, float* out, std::size_t n, float phaseIncrement, float* accumulatedPhase);
}

static std::vector<cf32> randomSignal(std::size_t n, unsigned seed) {
    std::mt19937                          rng(seed);
    std::uniform_real_distribution<float> dist(-1.f, 1.f);
    std::vector<cf32>                     x(n);
    for (auto& v : x) {
        v = {dist(rng), dist(rng)};
    }
    return x;
}

struct TaggedVectorSource : gr::Block<TaggedVectorSource> { // VectorSource that also publishes tags (ascending index)
    using gr::Block<TaggedVectorSource>::Block;
    gr::PortOut<cf32> out;
    GR_MAKE_REFLECTABLE(TaggedVectorSource, out);
    std::vector<cf32>    values;
    std::vector<gr::Tag> _tags;
    std::size_t          _position = 0, _nextTag = 0;
    gr::work::Status processBulk(std::span<cf32> output) {
        const std::size_t n = std::min(output.size(), values.size() - _position);
        std::copy_n(values.begin() + static_cast<std::ptrdiff_t>(_position), n, output.begin());
        while (_nextTag < _tags.size() && _tags[_nextTag].index < _position + n) {
            this->publishTag(_tags[_nextTag].map, _tags[_nextTag].index - _position);
            ++_nextTag;
        }
        _position += n;
        this->publishOnly(n);
        return _position >= values.size() ? gr::work::Status::DONE : gr::work::Status::OK;
    }
};

static bool bitEqual(const std::vector<cf32>& a, const std::vector<cf32>& b) { return a.size() == b.size() && std::memcmp(a.data(), b.data(), a.size() * sizeof(cf32)) == 0; }

int main() {
    if (gr4b200_device_count() < 1) {
        std::printf("no CUDA device: nothing to test\n");
        return 77;
    }
    constexpr std::size_t kFft = 4096;
    const std::string     gpu  = "gpu:cuda:0";

    "MultiplyConst on the device is bit-identical to the reference operator"_test = [&] {
        const auto x = randomSignal(100'000, 1);
        gr::Graph  g;
        auto&      src  = g.emplaceBlock<gr::testing::VectorSource<cf32>>();
        src.values      = x;
        auto& up        = g.emplaceBlock<gr::cuda::H2D<cf32>>();
        auto& mul       = g.emplaceBlock<gr::blocks::math::MultiplyConst<cf32>>({{"value", cf32(0.37f, -1.91f)}, {"compute_domain", gpu}});
        auto& div       = g.emplaceBlock<gr::blocks::math::DivideConst<cf32>>({{"value", cf32(1.5f, 0.25f)}, {"compute_domain", gpu}});
        auto& down      = g.emplaceBlock<gr::cuda::D2H<cf32>>();
        auto& sink      = g.emplaceBlock<gr::testing::VectorSink<cf32>>();
        expect(g.connect<"out", "in">(src, up).has_value() && g.connect<"out", "in">(up, mul).has_value() && g.connect<"out", "in">(mul, div).has_value());
        expect(g.connect<"out", "in">(div, down).has_value() && g.connect<"out", "in">(down, sink).has_value());
        gr::scheduler::Simple<> sched(std::move(g));
        auto                    result = sched.runAndWait();
        expect(result.has_value(), result ? "" : result.error().message.c_str());
        std::vector<cf32> tmp(x.size()), want(x.size());
        oracle_mathop_const_cf32(2, reinterpret_cast<const float*>(x.data()), reinterpret_cast<float*>(tmp.data()), x.size(), 0.37f, -1.91f);
        oracle_mathop_const_cf32(3, reinterpret_cast<const float*>(tmp.data()), reinterpret_cast<float*>(want.data()), x.size(), 1.5f, 0.25f);
        expect(mul.runsOnDevice() && div.runsOnDevice());
        expect(bitEqual(sink._samples, want));
    };

    "multiThreaded policy: host source, device chain and host sink on three launcher threads"_test = [&] {
        const auto         x = randomSignal(400'000, 11);
        std::vector<float> taps(127);
        oracle_fir_generate_f32(127, 2 /*Hamming*/, 0.1f, 1.6f, 1, taps.data());
        gr::Graph g;
        auto&     src = g.emplaceBlock<gr::testing::VectorSource<cf32>>();
        src.values    = x;
        auto& up      = g.emplaceBlock<gr::cuda::H2D<cf32>>();
        auto& mul     = g.emplaceBlock<gr::blocks::math::MultiplyConst<cf32>>({{"value", cf32(0.37f, -1.91f)}, {"compute_domain", gpu}});
        auto& fir     = g.emplaceBlock<gr::filter::fir_filter<cf32>>({{"b", taps}, {"compute_domain", gpu}});
        auto& down    = g.emplaceBlock<gr::cuda::D2H<cf32>>();
        auto& sink    = g.emplaceBlock<gr::testing::VectorSink<cf32>>();
        expect(g.connect<"out", "in">(src, up, {.minBufferSize = 30000}).has_value() && g.connect<"out", "in">(up, mul, {.minBufferSize = 30000}).has_value() && g.connect<"out", "in">(mul, fir, {.minBufferSize = 30000}).has_value());
        expect(g.connect<"out", "in">(fir, down, {.minBufferSize = 30000}).has_value() && g.connect<"out", "in">(down, sink, {.minBufferSize = 30000}).has_value());
        gr::scheduler::Simple<gr::scheduler::ExecutionPolicy::multiThreaded> sched(std::move(g));
        sched.host_threads = 2; // source and sink on different threads, the four device blocks on the device's thread
        auto result        = sched.runAndWait();
        expect(result.has_value(), result ? "" : result.error().message.c_str());
        std::vector<cf32> tmp(x.size()), want(x.size());
        oracle_mathop_const_cf32(2, reinterpret_cast<const float*>(x.data()), reinterpret_cast<float*>(tmp.data()), x.size(), 0.37f, -1.91f);
        oracle_fir_cf32(taps.data(), 127, reinterpret_cast<const float*>(tmp.data()), reinterpret_cast<float*>(want.data()), x.size(), nullptr);
        expect(bitEqual(sink._samples, want));
    };

    "Stride<> on a device edge: overlapping chunks across the end of the HBM ring"_test = [&] {
        struct StridedGain : gr::Block<StridedGain, gr::Resampling<96, 96, false>, gr::Stride<0, false>> {
            using gr::Block<StridedGain, gr::Resampling<96, 96, false>, gr::Stride<0, false>>::Block;
            gr::PortIn<cf32>  in;
            gr::PortOut<cf32> out;
            GR_MAKE_REFLECTABLE(StridedGain, in, out);
            gr::work::Status processBulk_cuda(void* stream, const cf32* input, cf32* output, std::size_t nIn, std::size_t) {
                return gr4b200_mathop_const_cf32(stream, GR4B200_OP_MULTIPLY, reinterpret_cast<const float*>(input), reinterpret_cast<float*>(output), nIn, 2.f, 0.f) == GR4B200_OK ? gr::work::Status::OK : gr::work::Status::ERROR;
            }
        };
        for (const gr::Size_t stride : {gr::Size_t{40}, gr::Size_t{250}}) { // overlap (40 < 96) and skip (250 > 96)
            const auto x = randomSignal(20'000, 21);
            gr::Graph  g;
            auto&      src = g.emplaceBlock<gr::testing::VectorSource<cf32>>();
            src.values     = x;
            auto& up       = g.emplaceBlock<gr::cuda::H2D<cf32>>();
            auto& gain     = g.emplaceBlock<StridedGain>({{"stride", stride}, {"compute_domain", gpu}});
            auto& down     = g.emplaceBlock<gr::cuda::D2H<cf32>>();
            auto& sink     = g.emplaceBlock<gr::testing::VectorSink<cf32>>();
            expect(g.connect<"out", "in">(src, up, {.minBufferSize = 1000}).has_value() && g.connect<"out", "in">(up, gain, {.minBufferSize = 1000}).has_value()); // 1000 is no multiple of 40 or 250: chunks wrap
            expect(g.connect<"out", "in">(gain, down, {.minBufferSize = 960}).has_value() && g.connect<"out", "in">(down, sink, {.minBufferSize = 960}).has_value());
            gr::scheduler::Simple<> sched(std::move(g));
            auto                    result = sched.runAndWait();
            expect(result.has_value(), result ? "" : result.error().message.c_str());
            std::vector<cf32> want;
            for (std::size_t first = 0; first + 96 <= x.size(); first += stride) {
                for (std::size_t i = 0; i < 96; ++i) {
                    want.push_back(x[first + i] * 2.f);
                }
            }
            expect(bitEqual(sink._samples, want), "strided chunks");
        }
    };

    "FIR -> FFT flowgraph (the north-star path): FIR bit-exact, spectrum within tolerance"_test = [&] {
        const std::size_t  n = kFft * 40;
        const auto         x = randomSignal(n, 2);
        std::vector<float> taps(127), window(kFft);
        oracle_fir_generate_f32(127, 2 /*Hamming*/, 0.1f, 1.6f, 1, taps.data());
        oracle_window_f32(3 /*Hann*/, kFft, 1.6f, window.data());
        using Frame = gr::blocks::fft::SpectrumFrame<kFft>;
        gr::Graph g;
        auto&     src = g.emplaceBlock<gr::testing::VectorSource<cf32>>();
        src.values    = x;
        auto& up      = g.emplaceBlock<gr::cuda::H2D<cf32>>();
        auto& fir     = g.emplaceBlock<gr::filter::fir_filter<cf32>>({{"b", taps}, {"compute_domain", gpu}});
        auto& fft     = g.emplaceBlock<gr::blocks::fft::FFT<cf32, kFft>>({{"window", "Hann"}, {"compute_domain", gpu}});
        auto& down    = g.emplaceBlock<gr::cuda::D2H<Frame>>();
        auto& sink    = g.emplaceBlock<gr::testing::VectorSink<Frame>>();
        // 20000-sample edges: deliberately not a multiple of the FFT size, chunks get cut to whole 4096 multiples
        expect(g.connect<"out", "in">(src, up, {.minBufferSize = 20000}).has_value() && g.connect<"out", "in">(up, fir, {.minBufferSize = 20000}).has_value());
        expect(g.connect<"out", "in">(fir, fft, {.minBufferSize = 5 * kFft}).has_value() && g.connect<"out", "in">(fft, down, {.minBufferSize = 8}).has_value() && g.connect<"out", "in">(down, sink, {.minBufferSize = 8}).has_value());
        gr::scheduler::Simple<> sched(std::move(g));
        auto                    result = sched.runAndWait();
        expect(result.has_value(), result ? "" : result.error().message.c_str());
        expect(sink._samples.size() == n / kFft);
        std::vector<cf32>  y(n);
        std::vector<float> want(n / kFft * 4 * kFft);
        oracle_fir_cf32(taps.data(), 127, reinterpret_cast<const float*>(x.data()), reinterpret_cast<float*>(y.data()), n, nullptr);
        oracle_fft_block_cf32(reinterpret_cast<const float*>(y.data()), kFft, n / kFft, window.data(), 0, 0, 0, want.data(), nullptr);
        float maxSpectrum = 0.f, errSpectrum = 0.f, errMagnitude = 0.f, maxMagnitude = 0.f;
        for (std::size_t c = 0; c < sink._samples.size(); ++c) {
            const float* w = want.data() + c * 4 * kFft;
            for (std::size_t k = 0; k < kFft; ++k) {
                maxSpectrum  = std::max({maxSpectrum, std::abs(w[2 * kFft + k]), std::abs(w[3 * kFft + k])});
                errSpectrum  = std::max({errSpectrum, std::abs(sink._samples[c].re[k] - w[2 * kFft + k]), std::abs(sink._samples[c].im[k] - w[3 * kFft + k])});
                maxMagnitude = std::max(maxMagnitude, w[k]);
                errMagnitude = std::max(errMagnitude, std::abs(sink._samples[c].magnitude[k] - w[k]));
            }
        }
        expect(errSpectrum <= 2.0e-6f * 64.f * maxSpectrum, "spectrum tolerance");
        expect(errMagnitude <= 1.0e-5f * maxMagnitude + 1e-7f, "magnitude tolerance");
        const auto ds = fft.materialise(sink._samples[0]);
        expect(ds.signal_values.size() == 4 * kFft && ds.signal_ranges.size() == 4 && ds.axis_values[kFft / 2] == 0.f && ds.signal_names[0] == "Magnitude(unknown signal)");
    };

    "DDC chain: Rotator -> BasicDecimatingFilter(x8): phase recurrence and FIR order as the reference"_test = [&] {
        const std::size_t n = 8 * 30000;
        const auto        x = randomSignal(n, 3);
        gr::Graph         g;
        auto&             src = g.emplaceBlock<gr::testing::VectorSource<cf32>>();
        src.values            = x;
        auto& up              = g.emplaceBlock<gr::cuda::H2D<cf32>>();
        auto& rot             = g.emplaceBlock<gr::blocks::math::Rotator<cf32>>({{"frequency_shift", 0.1f}, {"compute_domain", gpu}});
        auto& filt            = g.emplaceBlock<gr::filter::BasicDecimatingFilter<cf32>>({{"filter_order", 4}, {"f_low", 0.05f}, {"decimate", 8}, {"fir_design_method", 2}, {"compute_domain", gpu}});
        auto& down            = g.emplaceBlock<gr::cuda::D2H<cf32>>();
        auto& sink            = g.emplaceBlock<gr::testing::VectorSink<cf32>>();
        expect(g.connect<"out", "in">(src, up).has_value() && g.connect<"out", "in">(up, rot).has_value() && g.connect<"out", "in">(rot, filt).has_value());
        expect(g.connect<"out", "in">(filt, down).has_value() && g.connect<"out", "in">(down, sink).has_value());
        gr::scheduler::Simple<> sched(std::move(g));
        auto                    result = sched.runAndWait();
        expect(result.has_value(), result ? "" : result.error().message.c_str());
        expect(sink._samples.size() == n / 8);
        std::vector<cf32> mixed(n), want(n / 8);
        float             phase = 0.f;
        oracle_rotator_cf32(reinterpret_cast<const float*>(x.data()), reinterpret_cast<float*>(mixed.data()), n, gr4b200_rotator_phase_increment(0.1f, 1.f), &phase);
        expect(rot.accumulatedPhase() == phase, "accumulated phase is bit-identical");
        oracle_fir_decim_cf32(filt._taps.data(), filt._taps.size(), 8, reinterpret_cast<const float*>(mixed.data()), reinterpret_cast<float*>(want.data()), n, nullptr);
        float err = 0.f, sumTaps = 0.f;
        for (float t : filt._taps) {
            sumTaps += std::abs(t);
        }
        for (std::size_t i = 0; i < want.size(); ++i) {
            err = std::max(err, std::abs(sink._samples[i] - want[i]));
        }
        expect(err <= 6.f * 5.96e-8f * 1.42f * sumTaps * 1.42f, "round-1 bound (cos/sin <= 2 ulp propagated through the taps)");
        expect(bitEqual(sink._samples, want), "bit-identical since the mixer's cos / sin are the reference's (sincos_core.cuh)");
    };

    "Multiply with three device inputs folds left to right, bit-identical to std::complex (Math.hpp:100-107)"_test = [&] {
        const std::size_t n = 50'000;
        const auto        a = randomSignal(n, 11), b = randomSignal(n, 12), c = randomSignal(n, 13);
        gr::Graph         g;
        auto&             mul  = g.emplaceBlock<gr::blocks::math::Multiply<cf32>>({{"n_inputs", 3}, {"compute_domain", gpu}});
        auto&             down = g.emplaceBlock<gr::cuda::D2H<cf32>>();
        auto&             sink = g.emplaceBlock<gr::testing::VectorSink<cf32>>();
        const std::vector<cf32>* inputs[3] = {&a, &b, &c};
        for (int k = 0; k < 3; ++k) {
            auto& src  = g.emplaceBlock<gr::testing::VectorSource<cf32>>();
            src.values = *inputs[k];
            auto& up   = g.emplaceBlock<gr::cuda::H2D<cf32>>();
            expect(g.connect<"out", "in">(src, up).has_value() && g.connect(up, "out", mul, "in#" + std::to_string(k)).has_value());
        }
        expect(g.connect<"out", "in">(mul, down).has_value() && g.connect<"out", "in">(down, sink).has_value());
        gr::scheduler::Simple<> sched(std::move(g));
        auto                    result = sched.runAndWait();
        expect(result.has_value(), result ? "" : result.error().message.c_str());
        std::vector<cf32> want(n);
        for (std::size_t i = 0; i < n; ++i) {
            want[i] = (a[i] * b[i]) * c[i];
        }
        expect(bitEqual(sink._samples, want), "(a * b) * c with the reference's operator*");
    };

    "FFT block on a real stream (FFT<float>): half-spectrum frames, sine peak in its bin (qa_fourier.cpp:53-109)"_test = [&] {
        constexpr std::size_t kN = 1024;
        const std::size_t     n  = kN * 9;
        std::vector<float>    x(n);
        for (std::size_t i = 0; i < n; ++i) {
            x[i] = std::sin(2.f * 3.14159265f * 0.1f * static_cast<float>(i % kN)); // 0.1 fs: bin 102.4
        }
        using Block = gr::blocks::fft::FFT<float, kN>;
        using Frame = Block::Frame;
        static_assert(sizeof(Frame) == 4 * (kN / 2) * sizeof(float));
        gr::Graph g;
        auto&     src = g.emplaceBlock<gr::testing::VectorSource<float>>();
        src.values    = x;
        auto& up      = g.emplaceBlock<gr::cuda::H2D<float>>();
        auto& fft     = g.emplaceBlock<Block>({{"window", "Hann"}, {"sample_rate", 1000.f}, {"compute_domain", gpu}});
        auto& down    = g.emplaceBlock<gr::cuda::D2H<Frame>>();
        auto& sink    = g.emplaceBlock<gr::testing::VectorSink<Frame>>();
        expect(g.connect<"out", "in">(src, up).has_value() && g.connect<"out", "in">(up, fft).has_value() && g.connect<"out", "in">(fft, down, {.minBufferSize = 8}).has_value() && g.connect<"out", "in">(down, sink, {.minBufferSize = 8}).has_value());
        gr::scheduler::Simple<> sched(std::move(g));
        auto                    result = sched.runAndWait();
        expect(result.has_value(), result ? "" : result.error().message.c_str());
        expect(sink._samples.size() == n / kN);
        bool peaksOk = !sink._samples.empty();
        for (const auto& frame : sink._samples) {
            const auto peak = static_cast<std::size_t>(std::max_element(frame.magnitude.begin(), frame.magnitude.end()) - frame.magnitude.begin());
            peaksOk         = peaksOk && (peak == 102 || peak == 103);
        }
        expect(peaksOk, "the sine shows up at 0.1 fs in every frame");
        if (!sink._samples.empty()) {
            const auto ds = fft.materialise(sink._samples[0]);
            expect(ds.signal_values.size() == 4 * (kN / 2) && ds.axis_values.front() == 0.f && std::abs(ds.axis_values[102] - 99.609375f) < 1e-3f, "half-spectrum axis [DC, fs/2)");
        }
    };

    "int16 I/Q in, int16 I/Q out: InterleavedToComplex -> MultiplyConst -> ComplexToInterleaved on the device"_test = [&] {
        using namespace gr::blocks::type::converter;
        const std::size_t         n = 200'001; // odd: the last sample takes the item-wise kernel
        std::mt19937              rng(77);
        std::uniform_int_distribution<int> dist(-32768, 32767);
        std::vector<std::int16_t> iq(2 * n);
        for (auto& v : iq) {
            v = static_cast<std::int16_t>(dist(rng));
        }
        gr::Graph g;
        auto&     src = g.emplaceBlock<gr::testing::VectorSource<std::int16_t>>();
        src.values    = iq;
        auto& up      = g.emplaceBlock<gr::cuda::H2D<std::int16_t>>();
        auto& widen   = g.emplaceBlock<InterleavedToComplex<std::int16_t, cf32>>({{"compute_domain", gpu}});
        auto& gain    = g.emplaceBlock<gr::blocks::math::MultiplyConst<cf32>>({{"value", cf32(0.4375f, 0.25f)}, {"compute_domain", gpu}});
        auto& narrow  = g.emplaceBlock<ComplexToInterleaved<cf32, std::int16_t>>({{"compute_domain", gpu}});
        auto& down    = g.emplaceBlock<gr::cuda::D2H<std::int16_t>>();
        auto& sink    = g.emplaceBlock<gr::testing::VectorSink<std::int16_t>>();
        expect(g.connect<"out", "in">(src, up).has_value() && g.connect<"out", "interleaved">(up, widen).has_value() && g.connect<"out", "in">(widen, gain).has_value());
        expect(g.connect<"out", "in">(gain, narrow).has_value() && g.connect<"interleaved", "in">(narrow, down).has_value() && g.connect<"out", "in">(down, sink).has_value());
        gr::scheduler::Simple<> sched(std::move(g));
        auto                    result = sched.runAndWait();
        expect(result.has_value(), result ? "" : result.error().message.c_str());
        expect(widen.runsOnDevice() && narrow.runsOnDevice());
        std::vector<cf32> x(n), y(n);
        for (std::size_t i = 0; i < n; ++i) {
            x[i] = {static_cast<float>(iq[2 * i]), static_cast<float>(iq[2 * i + 1])};
        }
        oracle_mathop_const_cf32(2, reinterpret_cast<const float*>(x.data()), reinterpret_cast<float*>(y.data()), n, 0.4375f, 0.25f);
        std::vector<std::int16_t> want(2 * n);
        for (std::size_t i = 0; i < n; ++i) { // |y| < 32768 * 0.6875: in range, the cast truncates toward zero
            want[2 * i]     = static_cast<std::int16_t>(y[i].real());
            want[2 * i + 1] = static_cast<std::int16_t>(y[i].imag());
        }
        expect(sink._samples == want, "bit-identical int16 stream");
    };

    "fan-out on a device edge: the FIR output feeds a gain block and a decimator, each at its own pace"_test = [&] {
        const std::size_t n = 8 * 40'000;
        const auto        x = randomSignal(n, 21);
        std::vector<float> taps(127);
        oracle_fir_generate_f32(taps.size(), 2 /*Hamming*/, 0.1f, 1.6f, 1, taps.data());
        gr::Graph g;
        auto&     src   = g.emplaceBlock<gr::testing::VectorSource<cf32>>();
        src.values      = x;
        auto& up        = g.emplaceBlock<gr::cuda::H2D<cf32>>();
        auto& fir       = g.emplaceBlock<gr::filter::fir_filter<cf32>>({{"b", taps}, {"compute_domain", gpu}});
        auto& gain      = g.emplaceBlock<gr::blocks::math::MultiplyConst<cf32>>({{"value", cf32(2.f, 0.f)}, {"compute_domain", gpu}});
        auto& decim     = g.emplaceBlock<gr::filter::Decimator<cf32>>({{"decim", 8}, {"compute_domain", gpu}});
        auto& downA     = g.emplaceBlock<gr::cuda::D2H<cf32>>();
        auto& downB     = g.emplaceBlock<gr::cuda::D2H<cf32>>();
        auto& sinkA     = g.emplaceBlock<gr::testing::VectorSink<cf32>>();
        auto& sinkB     = g.emplaceBlock<gr::testing::VectorSink<cf32>>();
        expect(g.connect<"out", "in">(src, up).has_value() && g.connect<"out", "in">(up, fir).has_value());
        expect(g.connect<"out", "in">(fir, gain).has_value() && g.connect<"out", "in">(fir, decim).has_value(), "two readers on the FIR's HBM edge");
        expect(g.connect<"out", "in">(gain, downA).has_value() && g.connect<"out", "in">(downA, sinkA).has_value());
        expect(g.connect<"out", "in">(decim, downB).has_value() && g.connect<"out", "in">(downB, sinkB).has_value());
        gr::scheduler::BreadthFirst<> sched(std::move(g)); // any of the three schedulers drives a device graph
        auto                    result = sched.runAndWait();
        expect(result.has_value(), result ? "" : result.error().message.c_str());
        std::vector<cf32> y(n), wantA(n), wantB(n / 8);
        oracle_fir_cf32(taps.data(), taps.size(), reinterpret_cast<const float*>(x.data()), reinterpret_cast<float*>(y.data()), n, nullptr);
        for (std::size_t i = 0; i < n; ++i) {
            wantA[i] = y[i] * cf32(2.f, 0.f);
        }
        for (std::size_t i = 0; i < n / 8; ++i) {
            wantB[i] = y[8 * i];
        }
        expect(bitEqual(sinkA._samples, wantA), "branch A: FIR then gain, bit-identical");
        expect(bitEqual(sinkB._samples, wantB), "branch B: FIR then every 8th sample, bit-identical");
    };

    "tags across device chunks: Decimator and BasicDecimatingFilter rescale sample_rate (qa_filter.cpp:267-320)"_test = [&] {
        {
            constexpr float      kInputRate = 10'000.f;
            constexpr gr::Size_t kDecim = 10, kSamples = 100'000;
            gr::Graph g;
            auto&     source = g.emplaceBlock<gr::testing::TagSource<cf32>>({{"sample_rate", kInputRate}, {"n_samples_max", kSamples}});
            auto&     up     = g.emplaceBlock<gr::cuda::H2D<cf32>>();
            auto&     decim  = g.emplaceBlock<gr::filter::Decimator<cf32>>({{"decim", kDecim}, {"compute_domain", gpu}});
            auto&     down   = g.emplaceBlock<gr::cuda::D2H<cf32>>();
            auto&     sink   = g.emplaceBlock<gr::testing::TagSink<cf32>>({{"n_samples_expected", kSamples / kDecim}});
            source._tags     = {gr::Tag{50'000, {{"gr:trigger_name", "mark"}}}};
            expect(g.connect<"out", "in">(source, up).has_value() && g.connect<"out", "in">(up, decim).has_value() && g.connect<"out", "in">(decim, down).has_value() && g.connect<"out", "in">(down, sink).has_value());
            gr::scheduler::Simple<> sched(std::move(g));
            auto                    result = sched.runAndWait();
            expect(result.has_value(), result ? "" : result.error().message.c_str());
            expect(decim.input_chunk_size == kDecim && decim.output_chunk_size == 1);
            expect(sink._nSamplesProduced == kSamples / kDecim);
            expect(sink.sample_rate == kInputRate / static_cast<float>(kDecim), "rate seen downstream of the device decimator");
            bool sawMark = false;
            for (const auto& t : sink._tags) {
                sawMark = sawMark || (t.map.contains("gr:trigger_name") && t.index == 5'000);
            }
            expect(sawMark, "a mid-stream tag forces a chunk boundary on the device edge and keeps its (decimated) position");
            bool samplesOk = sink._samples.size() == kSamples / kDecim;
            for (std::size_t i = 0; samplesOk && i < sink._samples.size(); ++i) {
                samplesOk = sink._samples[i] == cf32(static_cast<float>(i * kDecim), 0.f);
            }
            expect(samplesOk, "decimated samples");
        }
        {
            constexpr float      kInputRate = 32'000.f;
            constexpr gr::Size_t kDecimate = 10, kSamples = 4'000;
            gr::Graph g;
            auto&     source = g.emplaceBlock<gr::testing::TagSource<cf32>>({{"sample_rate", kInputRate}, {"n_samples_max", kSamples}});
            auto&     up     = g.emplaceBlock<gr::cuda::H2D<cf32>>();
            auto&     filter = g.emplaceBlock<gr::filter::BasicDecimatingFilter<cf32>>({{"sample_rate", kInputRate}, {"f_low", 400.f}, {"decimate", kDecimate}, {"compute_domain", gpu}});
            auto&     down   = g.emplaceBlock<gr::cuda::D2H<cf32>>();
            auto&     sink   = g.emplaceBlock<gr::testing::TagSink<cf32>>({{"n_samples_expected", kSamples / kDecimate}});
            expect(g.connect<"out", "in">(source, up).has_value() && g.connect<"out", "in">(up, filter).has_value() && g.connect<"out", "in">(filter, down).has_value() && g.connect<"out", "in">(down, sink).has_value());
            gr::scheduler::Simple<> sched(std::move(g));
            auto                    result = sched.runAndWait();
            expect(result.has_value(), result ? "" : result.error().message.c_str());
            expect(filter.sample_rate == kInputRate, "filter member holds its input rate");
            expect(sink.sample_rate == kInputRate / static_cast<float>(kDecimate), "rate seen downstream");
        }
    };

    "new coefficients by tag on a running fir_filter keep the past samples exactly when the reference's HistoryBuffer would (time_domain_filter.hpp:39-43)"_test = [&] {
        struct Case {
            std::size_t first, second;
            bool        keeps;
        };
        constexpr std::size_t kSamples = 12'000, kChange = 6'100;
        for (const Case c : {Case{100, 127, true}, Case{127, 33, true}, Case{20, 32, true}, Case{20, 40, false}, Case{127, 129, false}, Case{5, 300, false}}) {
            for (const std::size_t edgeItems : {std::size_t{65536}, std::size_t{16}}) { // default edges: past samples in the input ring; 16-item edges: in the plan's state
                const auto                            x = randomSignal(kSamples, static_cast<unsigned>(c.first * 1000 + c.second));
                std::mt19937                          rng(static_cast<unsigned>(c.first + c.second));
                std::uniform_real_distribution<float> dist(-1.f, 1.f);
                std::vector<float>                    a(c.first), b(c.second);
                std::ranges::generate(a, [&] { return dist(rng); });
                std::ranges::generate(b, [&] { return dist(rng); });
                gr::Graph g;
                auto&     src = g.emplaceBlock<TaggedVectorSource>();
                src.values    = x;
                src._tags     = {gr::Tag{kChange, {{"b", b}}}};
                auto& up      = g.emplaceBlock<gr::cuda::H2D<cf32>>();
                auto& fir     = g.emplaceBlock<gr::filter::fir_filter<cf32>>({{"b", a}, {"compute_domain", gpu}});
                auto& down    = g.emplaceBlock<gr::cuda::D2H<cf32>>();
                auto& sink    = g.emplaceBlock<gr::testing::VectorSink<cf32>>();
                expect(g.connect<"out", "in">(src, up, {.minBufferSize = edgeItems}).has_value() && g.connect<"out", "in">(up, fir, {.minBufferSize = edgeItems}).has_value());
                expect(g.connect<"out", "in">(fir, down, {.minBufferSize = edgeItems}).has_value() && g.connect<"out", "in">(down, sink, {.minBufferSize = edgeItems}).has_value());
                gr::scheduler::Simple<> sched(std::move(g));
                auto                    result = sched.runAndWait();
                expect(result.has_value(), result ? "" : result.error().message.c_str());
                expect(fir.b == b, "the tag replaced the coefficients");
                std::vector<cf32> before(kSamples), after(kSamples), want(kSamples);
                oracle_fir_cf32(a.data(), a.size(), reinterpret_cast<const float*>(x.data()), reinterpret_cast<float*>(before.data()), kSamples, nullptr);
                if (c.keeps) { // the whole stream under the new coefficients, seen from the change on
                    oracle_fir_cf32(b.data(), b.size(), reinterpret_cast<const float*>(x.data()), reinterpret_cast<float*>(after.data()), kSamples, nullptr);
                } else { // a fresh, zeroed history buffer in front of the change
                    oracle_fir_cf32(b.data(), b.size(), reinterpret_cast<const float*>(x.data() + kChange), reinterpret_cast<float*>(after.data() + kChange), kSamples - kChange, nullptr);
                }
                std::copy_n(before.begin(), kChange, want.begin());
                std::copy(after.begin() + kChange, after.end(), want.begin() + kChange);
                char what[256];
                std::size_t firstBad = kSamples, nBad = 0;
                for (std::size_t i = 0; i < std::min(kSamples, sink._samples.size()); ++i) {
                    if (std::memcmp(&sink._samples[i], &want[i], sizeof(cf32)) != 0) {
                        firstBad = std::min(firstBad, i);
                        ++nBad;
                    }
                }
                std::snprintf(what, sizeof(what), "%zu -> %zu coefficients, history kept: %d, edge items %zu: %zu samples differ, the first at %zu of %zu", c.first, c.second, c.keeps ? 1 : 0, edgeItems, nBad, firstBad, sink._samples.size());
                expect(bitEqual(sink._samples, want), what);
            }
        }
    };

    "a setting that arrives by tag re-designs a running BasicDecimatingFilter: new taps over a zeroed history (time_domain_filter.hpp:160-181)"_test = [&] {
        constexpr std::size_t kDecimate = 8, kSamples = 8 * 1500, kChange = 8 * 500;
        auto                  design    = [](float fLow) {
            std::vector<float> taps(1u << 16);
            const long         n = gr4b200_fir_design_f32_host(0 /*LOWPASS*/, 4, fLow, 0.2, 1.0, 1.0, 40.0, 1.6, 2 /*Hamming*/, taps.data(), taps.size());
            taps.resize(n > 0 ? static_cast<std::size_t>(n) : 0);
            return taps;
        };
        const auto first = design(0.05f), second = design(0.08f);
        expect(!first.empty() && !second.empty() && first != second);
        for (const std::size_t edgeItems : {std::size_t{65536}, std::size_t{64}}) { // past samples in the input ring / in the plan's state
            const auto x = randomSignal(kSamples, 77);
            gr::Graph  g;
            auto&      src = g.emplaceBlock<TaggedVectorSource>();
            src.values     = x;
            src._tags      = {gr::Tag{kChange, {{"f_low", 0.08f}}}};
            auto& up       = g.emplaceBlock<gr::cuda::H2D<cf32>>();
            auto& filt     = g.emplaceBlock<gr::filter::BasicDecimatingFilter<cf32>>({{"filter_order", 4}, {"f_low", 0.05f}, {"decimate", 8}, {"fir_design_method", 2}, {"compute_domain", gpu}});
            auto& down     = g.emplaceBlock<gr::cuda::D2H<cf32>>();
            auto& sink     = g.emplaceBlock<gr::testing::VectorSink<cf32>>();
            expect(g.connect<"out", "in">(src, up, {.minBufferSize = edgeItems}).has_value() && g.connect<"out", "in">(up, filt, {.minBufferSize = edgeItems}).has_value());
            expect(g.connect<"out", "in">(filt, down, {.minBufferSize = edgeItems}).has_value() && g.connect<"out", "in">(down, sink, {.minBufferSize = edgeItems}).has_value());
            gr::scheduler::Simple<> sched(std::move(g));
            auto                    result = sched.runAndWait();
            expect(result.has_value(), result ? "" : result.error().message.c_str());
            expect(filt._taps == second, "re-designed from the tag");
            std::vector<cf32> want(kSamples / kDecimate);
            oracle_fir_decim_cf32(first.data(), first.size(), kDecimate, reinterpret_cast<const float*>(x.data()), reinterpret_cast<float*>(want.data()), kChange, nullptr);
            oracle_fir_decim_cf32(second.data(), second.size(), kDecimate, reinterpret_cast<const float*>(x.data() + kChange), reinterpret_cast<float*>(want.data() + kChange / kDecimate), kSamples - kChange, nullptr);
            expect(bitEqual(sink._samples, want), edgeItems > 64 ? "history in the input ring" : "history in the plan's state");
        }
    };

    if (gr4b200_device_count() >= 2) {
        "pipelined over two GPUs in one process: FIR on cuda:0 -> PeerCopy -> FFT block on cuda:1, same bits as on one GPU"_test = [&] {
            const std::size_t  n = kFft * 24;
            const auto         x = randomSignal(n, 31);
            std::vector<float> taps(127);
            oracle_fir_generate_f32(taps.size(), 2 /*Hamming*/, 0.1f, 1.6f, 1, taps.data());
            using Frame = gr::blocks::fft::SpectrumFrame<kFft>;
            auto run    = [&](bool twoDevices, std::vector<Frame>& frames) {
                gr::Graph g;
                auto&     src = g.emplaceBlock<gr::testing::VectorSource<cf32>>();
                src.values    = x;
                auto& up      = g.emplaceBlock<gr::cuda::H2D<cf32>>({{"device", 0}});
                auto& fir     = g.emplaceBlock<gr::filter::fir_filter<cf32>>({{"b", taps}, {"compute_domain", "gpu:cuda:0"}});
                auto& fft     = g.emplaceBlock<gr::blocks::fft::FFT<cf32, kFft>>({{"window", "Hann"}, {"compute_domain", twoDevices ? "gpu:cuda:1" : "gpu:cuda:0"}});
                auto& down    = g.emplaceBlock<gr::cuda::D2H<Frame>>({{"device", twoDevices ? 1 : 0}});
                auto& sink    = g.emplaceBlock<gr::testing::VectorSink<Frame>>();
                bool  wired   = g.connect<"out", "in">(src, up, {.minBufferSize = 20000}).has_value() && g.connect<"out", "in">(up, fir, {.minBufferSize = 20000}).has_value();
                if (twoDevices) {
                    auto& hop = g.emplaceBlock<gr::cuda::PeerCopy<cf32>>({{"source_device", 0}, {"device", 1}});
                    wired     = wired && g.connect<"out", "in">(fir, hop, {.minBufferSize = 5 * kFft}).has_value() && g.connect<"out", "in">(hop, fft, {.minBufferSize = 5 * kFft}).has_value();
                } else {
                    wired = wired && g.connect<"out", "in">(fir, fft, {.minBufferSize = 5 * kFft}).has_value();
                }
                wired = wired && g.connect<"out", "in">(fft, down, {.minBufferSize = 8}).has_value() && g.connect<"out", "in">(down, sink, {.minBufferSize = 8}).has_value();
                expect(wired);
                gr::scheduler::Simple<> sched(std::move(g));
                auto                    result = sched.runAndWait();
                expect(result.has_value(), result ? "" : result.error().message.c_str());
                frames = sink._samples;
            };
            std::vector<Frame> one, two;
            run(false, one);
            run(true, two);
            expect(one.size() == n / kFft && two.size() == one.size());
            expect(!one.empty() && std::memcmp(one.data(), two.data(), one.size() * sizeof(Frame)) == 0, "the inter-GPU edge changes no bit");
            { // a direct edge between blocks on different devices is refused with a pointer to the bridge block
                gr::Graph g;
                auto&     src  = g.emplaceBlock<gr::testing::VectorSource<cf32>>();
                src.values     = x;
                auto& up       = g.emplaceBlock<gr::cuda::H2D<cf32>>({{"device", 0}});
                auto& a        = g.emplaceBlock<gr::blocks::math::MultiplyConst<cf32>>({{"compute_domain", "gpu:cuda:0"}});
                auto& b        = g.emplaceBlock<gr::blocks::math::MultiplyConst<cf32>>({{"compute_domain", "gpu:cuda:1"}});
                auto& down     = g.emplaceBlock<gr::cuda::D2H<cf32>>({{"device", 1}});
                auto& sink     = g.emplaceBlock<gr::testing::VectorSink<cf32>>();
                expect(g.connect<"out", "in">(src, up).has_value() && g.connect<"out", "in">(up, a).has_value() && g.connect<"out", "in">(a, b).has_value() && g.connect<"out", "in">(b, down).has_value() && g.connect<"out", "in">(down, sink).has_value());
                gr::scheduler::Simple<> sched(std::move(g));
                auto                    result = sched.runAndWait();
                expect(!result.has_value() && result.error().message.find("PeerCopy") != std::string::npos);
            }
        };

        "an edge placed on the consumer's GPU: the producer's kernel stores into the peer ring over NVLink, no copy block"_test = [&] {
            const std::size_t  n = 8 * 300'000; // many turns of the ring: the cross-device events pace producer and consumer
            const auto         x = randomSignal(n, 41);
            std::vector<float> taps(127);
            oracle_fir_generate_f32(taps.size(), 2 /*Hamming*/, 0.1f, 1.6f, 1, taps.data());
            gr::Graph g;
            auto&     src = g.emplaceBlock<gr::testing::VectorSource<cf32>>();
            src.values    = x;
            auto& up      = g.emplaceBlock<gr::cuda::H2D<cf32>>({{"device", 0}});
            auto& gain    = g.emplaceBlock<gr::blocks::math::MultiplyConst<cf32>>({{"value", cf32(0.5f, -0.25f)}, {"compute_domain", "gpu:cuda:0"}});
            auto& fir     = g.emplaceBlock<gr::filter::fir_filter<cf32>>({{"b", taps}, {"compute_domain", "gpu:cuda:1"}});
            auto& decim   = g.emplaceBlock<gr::filter::Decimator<cf32>>({{"decim", 8}, {"compute_domain", "gpu:cuda:1"}});
            auto& down    = g.emplaceBlock<gr::cuda::D2H<cf32>>({{"device", 1}});
            auto& sink    = g.emplaceBlock<gr::testing::VectorSink<cf32>>();
            expect(g.connect<"out", "in">(src, up).has_value() && g.connect<"out", "in">(up, gain).has_value());
            expect(g.connect<"out", "in">(gain, fir, {.domain = "gpu:cuda:1"}).has_value(), "ring in GPU 1's HBM, written by GPU 0");
            expect(g.connect<"out", "in">(fir, decim).has_value() && g.connect<"out", "in">(decim, down).has_value() && g.connect<"out", "in">(down, sink).has_value());
            gr::scheduler::Simple<> sched(std::move(g));
            auto                    result = sched.runAndWait();
            expect(result.has_value(), result ? "" : result.error().message.c_str());
            std::vector<cf32> scaled(n), y(n), want(n / 8);
            oracle_mathop_const_cf32(2, reinterpret_cast<const float*>(x.data()), reinterpret_cast<float*>(scaled.data()), n, 0.5f, -0.25f);
            oracle_fir_cf32(taps.data(), taps.size(), reinterpret_cast<const float*>(scaled.data()), reinterpret_cast<float*>(y.data()), n, nullptr);
            for (std::size_t i = 0; i < n / 8; ++i) {
                want[i] = y[8 * i];
            }
            expect(bitEqual(sink._samples, want), "bit-identical through the peer edge");
        };
    } else {
        std::printf("    (one CUDA device: the two-GPU PeerCopy test is skipped)\n");
    }

    return summary();
}
