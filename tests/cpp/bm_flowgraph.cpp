// Throughput of the FIR(127, exact) -> FFT block(4096, Hann) flowgraph through the C++ host layer: gr::Graph +
// gr::scheduler::Simple, the surface BASELINE.json's north_star names (reference: core/benchmarks/bm_Scheduler.cpp builds
// its graphs the same way and times runAndWait). Two shapes, one JSON line each (GPU box only):
//   host   : pinned host array -> gr::cuda::HostSource -> fir_filter -> FFT -> gr::cuda::HostSink -> pinned host array;
//            copies in both directions inside the timed region, uploads / kernels / downloads on three streams;
//   device : a capture resident in HBM -> fir_filter -> FFT -> device sink, for a sweep of work-chunk sizes: what the
//            scheduler, the rings and the per-chunk launches cost next to one launch over the whole stream.
// usage: bm_flowgraph [--samples N] [--chunk C] [--sweep] [--host-only | --device-only] [--device D] [--repeats R]
#include <chrono>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include <gnuradio-4.0/Scheduler.hpp>
#include <gnuradio-4.0/cuda/Transfer.hpp>
#include <gnuradio-4.0/filter/time_domain_filter.hpp>
#include <gnuradio-4.0/fourier/fft.hpp>

using cf32                 = std::complex<float>;
constexpr std::size_t kFft = 4096;
using Frame                = gr::blocks::fft::SpectrumFrame<kFft>;

static std::vector<float> lowPassTaps() {
    std::vector<float> taps(127);
    gr4b200_fir_generate_f32_host(127, 2 /*Hamming*/, 0.1f, 1.6f, 1, taps.data());
    return taps;
}

struct Result {
    double      setupSeconds = 0.0; // init(): edges, streams, plans -- before the timed region
    double      seconds  = 0.0;
    std::size_t frames   = 0;
    double      checksum = 0.0;
    std::string error;
};

// host array -> device chain -> host array
static Result runHost(int device, const cf32* hostIn, std::size_t nSamples, Frame* hostOut, std::size_t chunk) {
    const std::string gpu = "gpu:cuda:" + std::to_string(device);
    gr::Graph         g;
    auto&             src  = g.emplaceBlock<gr::cuda::HostSource<cf32>>({{"device", static_cast<gr::Size_t>(device)}});
    auto&             fir  = g.emplaceBlock<gr::filter::fir_filter<cf32>>({{"b", lowPassTaps()}, {"compute_domain", gpu}});
    auto&             fft  = g.emplaceBlock<gr::blocks::fft::FFT<cf32, kFft>>({{"window", "Hann"}, {"compute_domain", gpu}});
    auto&             sink = g.emplaceBlock<gr::cuda::HostSink<Frame>>({{"device", static_cast<gr::Size_t>(device)}});
    src.setData(hostIn, nSamples);
    sink.setBuffer(hostOut, nSamples / kFft);
    Result r;
    if (!g.connect<"out", "in">(src, fir, {.minBufferSize = 2 * chunk}) || !g.connect<"out", "in">(fir, fft, {.minBufferSize = 2 * chunk}) || !g.connect<"out", "in">(fft, sink, {.minBufferSize = 2 * chunk / kFft})) {
        r.error = "connect failed";
        return r;
    }
    gr::scheduler::Simple<> sched(std::move(g));
    sched.max_work_items = chunk;
    const auto tSetup    = std::chrono::steady_clock::now();
    if (const auto ready = sched.init(); !ready) { // allocations (rings, plans) and stream creation: not part of the stream rate
        r.error = ready.error().message;
        return r;
    }
    gr4b200_stream_synchronize(nullptr);
    const auto t0   = std::chrono::steady_clock::now();
    r.setupSeconds  = std::chrono::duration<double>(t0 - tSetup).count();
    const auto done = sched.runAndWait();
    r.seconds       = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!done) {
        r.error = done.error().message;
        return r;
    }
    r.frames = sink.itemsReceived();
    for (std::size_t k = 0; k < kFft; k += 64) {
        r.checksum += std::abs(hostOut[0].re[k]) + std::abs(hostOut[r.frames - 1].im[k]);
    }
    return r;
}

// capture in HBM -> device chain -> device sink
static Result runDevice(int device, const cf32* deviceCapture, std::size_t captureSize, std::size_t nSamples, std::size_t chunk) {
    const std::string gpu = "gpu:cuda:" + std::to_string(device);
    gr::Graph         g;
    auto&             src  = g.emplaceBlock<gr::cuda::DeviceReplaySource<cf32>>({{"device", static_cast<gr::Size_t>(device)}, {"n_samples_max", static_cast<gr::Size_t>(nSamples)}});
    auto&             fir  = g.emplaceBlock<gr::filter::fir_filter<cf32>>({{"b", lowPassTaps()}, {"compute_domain", gpu}});
    auto&             fft  = g.emplaceBlock<gr::blocks::fft::FFT<cf32, kFft>>({{"window", "Hann"}, {"compute_domain", gpu}});
    auto&             sink = g.emplaceBlock<gr::cuda::DeviceNullSink<Frame>>({{"device", static_cast<gr::Size_t>(device)}});
    src.setCapture(deviceCapture, captureSize);
    Result r;
    if (!g.connect<"out", "in">(src, fir, {.minBufferSize = 2 * chunk}) || !g.connect<"out", "in">(fir, fft, {.minBufferSize = 2 * chunk}) || !g.connect<"out", "in">(fft, sink, {.minBufferSize = 2 * chunk / kFft})) {
        r.error = "connect failed";
        return r;
    }
    gr::scheduler::Simple<> sched(std::move(g));
    sched.max_work_items = chunk;
    const auto tSetup    = std::chrono::steady_clock::now();
    if (const auto ready = sched.init(); !ready) { // allocations (rings, plans) and stream creation: not part of the stream rate
        r.error = ready.error().message;
        return r;
    }
    gr4b200_stream_synchronize(nullptr);
    const auto t0   = std::chrono::steady_clock::now();
    r.setupSeconds  = std::chrono::duration<double>(t0 - tSetup).count();
    const auto done = sched.runAndWait();
    r.seconds       = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!done) {
        r.error = done.error().message;
        return r;
    }
    r.frames = sink._count;
    return r;
}

int main(int argc, char** argv) {
    std::size_t nSamples = std::size_t{1} << 27, chunk = std::size_t{1} << 22;
    int         device = 0, repeats = 3;
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

    if (hostLeg) {
        auto* hostIn  = static_cast<cf32*>(gr4b200_malloc_host(nSamples * sizeof(cf32)));
        auto* hostOut = static_cast<Frame*>(gr4b200_malloc_host(nSamples / kFft * sizeof(Frame)));
        if (hostIn == nullptr || hostOut == nullptr) {
            std::printf("{\"error\": \"pinned allocation failed: %s\"}\n", gr4b200_last_error());
            return 1;
        }
        std::mt19937                          rng(device + 1);
        std::uniform_real_distribution<float> dist(-1.f, 1.f);
        for (std::size_t i = 0; i < std::min<std::size_t>(nSamples, 1u << 22); ++i) {
            hostIn[i] = {dist(rng), dist(rng)};
        }
        for (std::size_t i = 1u << 22; i < nSamples; ++i) { // the rest repeats the first 4 Mi samples (cheap to generate)
            hostIn[i] = hostIn[i & ((1u << 22) - 1)];
        }
        std::memset(static_cast<void*>(hostOut), 0, nSamples / kFft * sizeof(Frame));
        Result best;
        for (int rep = 0; rep <= repeats; ++rep) { // rep 0 is the warm-up (plans, first touch)
            const Result r = runHost(device, hostIn, nSamples, hostOut, chunk);
            if (!r.error.empty()) {
                std::printf("{\"leg\": \"host\", \"error\": \"%s\"}\n", r.error.c_str());
                return 1;
            }
            if (rep > 0 && (best.seconds == 0.0 || r.seconds < best.seconds)) {
                best = r;
            }
        }
        std::printf("{\"leg\": \"host\", \"api\": \"c++ gr::Graph / gr::scheduler::Simple::runAndWait, HostSource -> fir_filter -> FFT -> HostSink, pinned host arrays, 3 streams\", \"samples\": %zu, \"chunk\": %zu, \"seconds\": %.6f, \"msamples_per_s\": %.1f, \"frames\": %zu, \"h2d_bytes\": %zu, \"d2h_bytes\": %zu, \"checksum\": %.4f, \"repeats\": %d, \"setup_seconds\": %.4f}\n", nSamples, chunk, best.seconds,
            static_cast<double>(nSamples) / best.seconds / 1e6, best.frames, nSamples * sizeof(cf32), best.frames * sizeof(Frame), best.checksum, repeats, best.setupSeconds);
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
            Result            best;
            for (int rep = 0; rep <= repeats; ++rep) {
                const Result r = runDevice(device, capture, captureSize, n, c);
                if (!r.error.empty()) {
                    std::printf("{\"leg\": \"device\", \"chunk\": %zu, \"error\": \"%s\"}\n", c, r.error.c_str());
                    return 1;
                }
                if (rep > 0 && (best.seconds == 0.0 || r.seconds < best.seconds)) {
                    best = r;
                }
            }
            std::printf("{\"leg\": \"device\", \"api\": \"c++ gr::Graph / gr::scheduler::Simple::runAndWait, capture in HBM -> fir_filter -> FFT -> device sink\", \"samples\": %zu, \"chunk\": %zu, \"seconds\": %.6f, \"msamples_per_s\": %.1f, \"frames\": %zu, \"us_per_chunk\": %.2f, \"setup_seconds\": %.4f}\n", n, c, best.seconds, static_cast<double>(n) / best.seconds / 1e6, best.frames,
                best.seconds * 1e6 / (static_cast<double>(n) / static_cast<double>(c)), best.setupSeconds);
        }
        gr4b200_free(capture);
    }
    return 0;
}
