// Throughput of the FIR(127, exact) -> FFT block(4096, Hann) flowgraph through the C++ host layer: gr::Graph +
// gr::scheduler::Simple, the surface BASELINE.json's north_star names (reference: core/benchmarks/bm_Scheduler.cpp builds
// its graphs the same way and times runAndWait). Two shapes (GPU box only):
//   host   : pinned host array -> gr::cuda::HostSource [-> InterleavedToComplex] -> fir_filter -> FFT -> gr::cuda::HostSink
//            -> pinned host array; copies in both directions inside the timed region, uploads / kernels / downloads on
//            three streams. Variants: complex<float> or interleaved int16 I/Q in; the four DataSet planes or only the
//            magnitude plane out.
//   device : a capture resident in HBM -> fir_filter -> FFT -> device sink, for a sweep of work-chunk sizes: what the
//            scheduler, the rings and the per-chunk launches cost next to one launch over the whole stream.
// Built twice (tests/cpp/Makefile): the executable, and libbm_flowgraph.so whose extern "C" entry points bench.py calls
// in-process for its `e2e` figure (its own pinned buffers, its own barrier between ranks).
// usage: bm_flowgraph [--samples N] [--chunk C] [--sweep] [--host-only | --device-only] [--device D] [--repeats R] [--variant V]
#include <chrono>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include <gnuradio-4.0/Scheduler.hpp>
#include <gnuradio-4.0/basic/ConverterBlocks.hpp>
#include <gnuradio-4.0/cuda/Transfer.hpp>
#include <gnuradio-4.0/filter/time_domain_filter.hpp>
#include <gnuradio-4.0/fourier/fft.hpp>

namespace {

using cf32                 = std::complex<float>;
constexpr std::size_t kFft = 4096;
using Frame                = gr::blocks::fft::SpectrumFrame<kFft>;

enum Variant : int { kComplexInAllPlanes = 0, kInt16InAllPlanes = 1, kComplexInMagnitude = 2, kInt16InMagnitude = 3 };

std::vector<float> lowPassTaps() {
    std::vector<float> taps(127);
    gr4b200_fir_generate_f32_host(127, 2 /*Hamming*/, 0.1f, 1.6f, 1, taps.data());
    return taps;
}

struct Result {
    double      setupSeconds = 0.0; // init(): edges, streams, plans -- before the timed region
    double      seconds      = 0.0;
    std::size_t frames       = 0;
    std::string error;
};

template<typename Scheduler>
void timedRun(Scheduler& sched, Result& r) {
    const auto tSetup = std::chrono::steady_clock::now();
    if (const auto ready = sched.init(); !ready) { // allocations (rings, plans) and stream creation: not part of the stream rate
        r.error = ready.error().message;
        return;
    }
    gr4b200_stream_synchronize(nullptr);
    const auto t0   = std::chrono::steady_clock::now();
    r.setupSeconds  = std::chrono::duration<double>(t0 - tSetup).count();
    const auto done = sched.runAndWait();
    r.seconds       = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!done) {
        r.error = done.error().message;
    }
}

// host array -> device chain -> host array. hostIn: nSamples complex<float> (variants 0, 2) or 2 * nSamples int16 (1, 3);
// hostOut: nSamples / 4096 frames of 4 x 4096 floats (0, 1) or of 4096 floats, the magnitude plane (2, 3)
Result runHost(int device, int variant, const void* hostIn, std::size_t nSamples, void* hostOut, std::size_t chunk) {
    const std::string gpu      = "gpu:cuda:" + std::to_string(device);
    const auto        dev      = static_cast<gr::Size_t>(device);
    const bool        int16In  = variant == kInt16InAllPlanes || variant == kInt16InMagnitude;
    const bool        magnitudeOnly = variant == kComplexInMagnitude || variant == kInt16InMagnitude;
    gr::Graph         g;
    auto&             fir  = g.emplaceBlock<gr::filter::fir_filter<cf32>>({{"b", lowPassTaps()}, {"compute_domain", gpu}});
    auto&             fft  = g.emplaceBlock<gr::blocks::fft::FFT<cf32, kFft>>({{"window", "Hann"}, {"compute_domain", gpu}});
    auto&             sink = g.emplaceBlock<gr::cuda::HostSink<Frame>>(magnitudeOnly ? gr::property_map{{"device", dev}, {"copy_offset", gr::Size_t{0}}, {"copy_bytes", static_cast<gr::Size_t>(kFft * sizeof(float))}} : gr::property_map{{"device", dev}});
    sink.setBuffer(hostOut, nSamples / kFft);
    Result r;
    bool   connected = true;
    if (int16In) {
        auto& src   = g.emplaceBlock<gr::cuda::HostSource<std::int16_t>>({{"device", dev}});
        auto& widen = g.emplaceBlock<gr::blocks::type::converter::InterleavedToComplex<std::int16_t, cf32>>({{"compute_domain", gpu}});
        src.setData(static_cast<const std::int16_t*>(hostIn), 2 * nSamples);
        connected = g.connect<"out", "interleaved">(src, widen, {.minBufferSize = 4 * chunk}).has_value() && g.connect<"out", "in">(widen, fir, {.minBufferSize = 2 * chunk}).has_value();
    } else {
        auto& src = g.emplaceBlock<gr::cuda::HostSource<cf32>>({{"device", dev}});
        src.setData(static_cast<const cf32*>(hostIn), nSamples);
        connected = g.connect<"out", "in">(src, fir, {.minBufferSize = 2 * chunk}).has_value();
    }
    connected = connected && g.connect<"out", "in">(fir, fft, {.minBufferSize = 2 * chunk}).has_value() && g.connect<"out", "in">(fft, sink, {.minBufferSize = 2 * chunk / kFft}).has_value();
    if (!connected) {
        r.error = "connect failed";
        return r;
    }
    gr::scheduler::Simple<> sched(std::move(g));
    sched.max_work_items = int16In ? 2 * chunk : chunk; // the int16 source counts items, two per sample
    timedRun(sched, r);
    r.frames = sink.itemsReceived();
    return r;
}

// capture in HBM -> device chain -> device sink
Result runDevice(int device, const cf32* deviceCapture, std::size_t captureSize, std::size_t nSamples, std::size_t chunk) {
    const std::string gpu = "gpu:cuda:" + std::to_string(device);
    const auto        dev = static_cast<gr::Size_t>(device);
    gr::Graph         g;
    auto&             src  = g.emplaceBlock<gr::cuda::DeviceReplaySource<cf32>>({{"device", dev}, {"n_samples_max", static_cast<gr::Size_t>(nSamples)}});
    auto&             fir  = g.emplaceBlock<gr::filter::fir_filter<cf32>>({{"b", lowPassTaps()}, {"compute_domain", gpu}});
    auto&             fft  = g.emplaceBlock<gr::blocks::fft::FFT<cf32, kFft>>({{"window", "Hann"}, {"compute_domain", gpu}});
    auto&             sink = g.emplaceBlock<gr::cuda::DeviceNullSink<Frame>>({{"device", dev}});
    src.setCapture(deviceCapture, captureSize);
    Result r;
    if (!g.connect<"out", "in">(src, fir, {.minBufferSize = 2 * chunk}) || !g.connect<"out", "in">(fir, fft, {.minBufferSize = 2 * chunk}) || !g.connect<"out", "in">(fft, sink, {.minBufferSize = 2 * chunk / kFft})) {
        r.error = "connect failed";
        return r;
    }
    gr::scheduler::Simple<> sched(std::move(g));
    sched.max_work_items = chunk;
    if (const char* env = std::getenv("GR4B200_COMPUTE_STREAMS"); env != nullptr) { // A/B: kernel streams the independent chunks rotate over
        sched.compute_streams = static_cast<std::size_t>(std::max(1, std::atoi(env)));
    }
    timedRun(sched, r);
    r.frames = sink._count;
    return r;
}

void copyError(const std::string& message, char* error, std::size_t errorLength) {
    if (error != nullptr && errorLength > 0) {
        std::snprintf(error, errorLength, "%s", message.c_str());
    }
}

} // namespace

extern "C" {

// One pass of the host-to-host flowgraph on `device` (the calling thread is bound to it). Returns 0 and fills
// seconds / setupSeconds / frames, or -1 with a message in `error`.
int bm_flowgraph_host(int device, int variant, const void* hostIn, std::size_t nSamples, void* hostOut, std::size_t chunk, double* seconds, double* setupSeconds, std::size_t* frames, char* error, std::size_t errorLength) {
    if (gr4b200_init(device) != GR4B200_OK) {
        copyError(gr4b200_last_error(), error, errorLength);
        return -1;
    }
    const Result r = runHost(device, variant, hostIn, nSamples / kFft * kFft, hostOut, std::max(kFft, chunk / kFft * kFft));
    if (!r.error.empty()) {
        copyError(r.error, error, errorLength);
        return -1;
    }
    *seconds      = r.seconds;
    *setupSeconds = r.setupSeconds;
    *frames       = r.frames;
    return 0;
}

// One pass of the device-resident flowgraph; deviceCapture holds captureSize >= 2 * chunk samples in HBM.
int bm_flowgraph_device(int device, const void* deviceCapture, std::size_t captureSize, std::size_t nSamples, std::size_t chunk, double* seconds, double* setupSeconds, std::size_t* frames, char* error, std::size_t errorLength) {
    if (gr4b200_init(device) != GR4B200_OK) {
        copyError(gr4b200_last_error(), error, errorLength);
        return -1;
    }
    const Result r = runDevice(device, static_cast<const cf32*>(deviceCapture), captureSize, nSamples / kFft * kFft, std::max(kFft, chunk / kFft * kFft));
    if (!r.error.empty()) {
        copyError(r.error, error, errorLength);
        return -1;
    }
    *seconds      = r.seconds;
    *setupSeconds = r.setupSeconds;
    *frames       = r.frames;
    return 0;
}

} // extern "C"

#ifndef BM_FLOWGRAPH_LIBRARY
int main(int argc, char** argv) {
    std::size_t nSamples = std::size_t{1} << 27, chunk = std::size_t{1} << 22;
    int         device = 0, repeats = 3, variant = 0;
    bool        sweep = false, hostLeg = true, deviceLeg = true;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--samples" && i + 1 < argc) {
            nSamples = std::strtoull(argv[++i], nullptr, 0);
        } else if (a == "--chunk" && i + 1 < argc) {
            chunk = std::strtoull(argv[++i], nullptr, 0);
        } else if (a == "--device" && i + 1 < argc) {
            device = std::atoi(argv[++i]);
        } else if (a == "--repeats" && i + 1 < argc) {
            repeats = std::atoi(argv[++i]);
        } else if (a == "--variant" && i + 1 < argc) {
            variant = std::atoi(argv[++i]);
        } else if (a == "--sweep") {
            sweep = true;
        } else if (a == "--host-only") {
            deviceLeg = false;
        } else if (a == "--device-only") {
            hostLeg = false;
        }
    }
    if (gr4b200_device_count() <= device || gr4b200_init(device) != GR4B200_OK) {
        std::printf("{\"error\": \"no CUDA device %d\"}\n", device);
        return 77;
    }
    nSamples = nSamples / kFft * kFft;
    chunk    = std::max(kFft, chunk / kFft * kFft);
    char error[512] = {};

    if (hostLeg) {
        const bool        int16In  = variant == kInt16InAllPlanes || variant == kInt16InMagnitude;
        const bool        magnitudeOnly = variant == kComplexInMagnitude || variant == kInt16InMagnitude;
        const std::size_t inBytes  = nSamples * (int16In ? 2 * sizeof(std::int16_t) : sizeof(cf32));
        const std::size_t outBytes = nSamples / kFft * (magnitudeOnly ? kFft * sizeof(float) : sizeof(Frame));
        void*             hostIn   = gr4b200_malloc_host(inBytes);
        void*             hostOut  = gr4b200_malloc_host(outBytes);
        if (hostIn == nullptr || hostOut == nullptr) {
            std::printf("{\"error\": \"pinned allocation failed: %s\"}\n", gr4b200_last_error());
            return 1;
        }
        std::mt19937                          rng(device + 1);
        std::uniform_real_distribution<float> dist(-1.f, 1.f);
        const std::size_t                     period = std::min<std::size_t>(nSamples, 1u << 22); // the rest repeats the first 4 Mi samples
        for (std::size_t i = 0; i < nSamples; ++i) {
            if (int16In) {
                auto* p      = static_cast<std::int16_t*>(hostIn);
                p[2 * i]     = i < period ? static_cast<std::int16_t>(32767.f * dist(rng)) : p[2 * (i - period)];
                p[2 * i + 1] = i < period ? static_cast<std::int16_t>(32767.f * dist(rng)) : p[2 * (i - period) + 1];
            } else {
                auto* p = static_cast<cf32*>(hostIn);
                p[i]    = i < period ? cf32{dist(rng), dist(rng)} : p[i - period];
            }
        }
        std::memset(hostOut, 0, outBytes);
        double      best = 0.0, bestSetup = 0.0;
        std::size_t frames = 0;
        for (int rep = 0; rep <= repeats; ++rep) { // rep 0 is the warm-up (module load, first touch)
            double seconds = 0.0, setup = 0.0;
            if (bm_flowgraph_host(device, variant, hostIn, nSamples, hostOut, chunk, &seconds, &setup, &frames, error, sizeof error) != 0) {
                std::printf("{\"leg\": \"host\", \"error\": \"%s\"}\n", error);
                return 1;
            }
            if (rep > 0 && (best == 0.0 || seconds < best)) {
                best      = seconds;
                bestSetup = setup;
            }
        }
        double checksum = 0.0;
        for (std::size_t k = 0; k < kFft; k += 64) {
            checksum += std::abs(static_cast<const float*>(hostOut)[k]) + std::abs(static_cast<const float*>(hostOut)[outBytes / sizeof(float) - 1 - k]);
        }
        std::printf("{\"leg\": \"host\", \"variant\": %d, \"api\": \"c++ gr::Graph / gr::scheduler::Simple::runAndWait, HostSource -> %sfir_filter -> FFT -> HostSink%s, pinned host arrays, 3 streams\", \"samples\": %zu, \"chunk\": %zu, \"seconds\": %.6f, \"msamples_per_s\": %.1f, \"frames\": %zu, \"h2d_bytes\": %zu, \"d2h_bytes\": %zu, \"checksum\": %.4f, \"repeats\": %d, \"setup_seconds\": %.4f}\n", variant,
            int16In ? "InterleavedToComplex<int16> -> " : "", magnitudeOnly ? " (magnitude plane only)" : "", nSamples, chunk, best, static_cast<double>(nSamples) / best / 1e6, frames, inBytes, outBytes, checksum, repeats, bestSetup);
        gr4b200_free_host(hostIn);
        gr4b200_free_host(hostOut);
    }
    if (deviceLeg) {
        std::vector<std::size_t> chunks;
        if (sweep) {
            for (std::size_t c = std::size_t{1} << 16; c <= (std::size_t{1} << 24); c <<= 2) {
                chunks.push_back(c);
            }
        } else {
            chunks.push_back(chunk);
        }
        const std::size_t captureSize = 2 * chunks.back();
        auto*             capture     = static_cast<cf32*>(gr4b200_malloc(captureSize * sizeof(cf32)));
        std::vector<cf32> host(captureSize);
        std::mt19937      rng(99);
        std::uniform_real_distribution<float> dist(-1.f, 1.f);
        for (auto& v : host) {
            v = {dist(rng), dist(rng)};
        }
        if (capture == nullptr || gr4b200_copy_h2d(capture, host.data(), captureSize * sizeof(cf32), nullptr) != GR4B200_OK || gr4b200_stream_synchronize(nullptr) != GR4B200_OK) {
            std::printf("{\"error\": \"capture upload failed: %s\"}\n", gr4b200_last_error());
            return 1;
        }
        for (const std::size_t c : chunks) {
            // small chunks are bound by the launcher thread: keep the run short enough to finish in seconds
            const std::size_t n = std::min(nSamples, std::max<std::size_t>(c * 4096, std::size_t{1} << 24)) / kFft * kFft;
            double            best = 0.0, bestSetup = 0.0;
            std::size_t       frames = 0;
            for (int rep = 0; rep <= repeats; ++rep) {
                double seconds = 0.0, setup = 0.0;
                if (bm_flowgraph_device(device, capture, std::min(captureSize, 2 * c), n, c, &seconds, &setup, &frames, error, sizeof error) != 0) { // the source fills its two-chunk ring once
                    std::printf("{\"leg\": \"device\", \"chunk\": %zu, \"error\": \"%s\"}\n", c, error);
                    return 1;
                }
                if (rep > 0 && (best == 0.0 || seconds < best)) {
                    best      = seconds;
                    bestSetup = setup;
                }
            }
            std::printf("{\"leg\": \"device\", \"api\": \"c++ gr::Graph / gr::scheduler::Simple::runAndWait, capture in HBM -> fir_filter -> FFT -> device sink\", \"samples\": %zu, \"chunk\": %zu, \"seconds\": %.6f, \"msamples_per_s\": %.1f, \"frames\": %zu, \"us_per_chunk\": %.2f, \"setup_seconds\": %.4f}\n", n, c, best, static_cast<double>(n) / best / 1e6, frames,
                best * 1e6 / (static_cast<double>(n) / static_cast<double>(c)), bestSetup);
        }
        gr4b200_free(capture);
    }
    return 0;
}
#endif
