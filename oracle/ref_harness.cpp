// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or called from the product path.
//
// oracle/_ref/libgr4ref.so: the REFERENCE's own algorithm sources for the hot path, compiled where they lie under
// /root/reference (nothing is copied into this repo) behind a small C ABI so that tests can pin oracle/oracle.cpp
// against the real thing and bench.py can time it as the CPU baseline (`cpu_baseline.kind == "reference"`).
//
// What is the reference's code here (compiled in place, see oracle/Makefile for the include paths):
//   * gr::algorithm::FFT<std::complex<float>> -> SimdFFT           algorithm/.../fourier/fft.hpp:113, SimdFFT.hpp:491
//   * gr::algorithm::window::create                                 algorithm/.../fourier/window.hpp:71
//   * gr::algorithm::fft::computeMagnitudeSpectrum / PhaseSpectrum  algorithm/.../fourier/fft_common.hpp:22,93
//   * gr::filter::fir::generateCoefficients / designFilter          algorithm/.../filter/FilterTool.hpp:964,1007
//   * gr::filter::Filter<T>::processOne (FIR branch)                algorithm/.../filter/FilterTool.hpp:137-139,244
//   * gr::HistoryBuffer<T>                                          core/.../HistoryBuffer.hpp
// What is NOT compilable with g++-13 (needs Block.hpp -> <print>, reflection, GCC>=14) and is therefore re-stated
// here as the same one- or two-line expression the block body holds, on top of the real pieces above:
//   * fir_filter<T>::processOne            blocks/filter/.../time_domain_filter.hpp:44-47
//   * BasicFilterProto::processBulk        blocks/filter/.../time_domain_filter.hpp:190-204
//   * MathOpImpl<T,op>::processOne         blocks/math/.../Math.hpp:38-56           (std::complex operators)
//   * Rotator<T>::processOne               blocks/math/.../Rotator.hpp:51-61
//   * FFT block processBulk/createDataset  blocks/fourier/.../fft.hpp:147-250
// The un-vendored vir-simd dependency is replaced by oracle/ref_shim/vir/simd.h (std::experimental::simd).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstring>
#include <execution>
#include <functional>
#include <numbers>
#include <numeric>
#include <vector>

#include <gnuradio-4.0/HistoryBuffer.hpp>
#include <gnuradio-4.0/algorithm/filter/FilterTool.hpp>
#include <gnuradio-4.0/algorithm/fourier/fft.hpp>
#include <gnuradio-4.0/algorithm/fourier/fft_common.hpp>
#include <gnuradio-4.0/algorithm/fourier/window.hpp>

using cf32 = std::complex<float>;

namespace {
template<typename T>
struct FirFilterBody { // state + body of fir_filter<T> (time_domain_filter.hpp:36-47)
    std::vector<T>       b;
    gr::HistoryBuffer<T> inputHistory{32};

    explicit FirFilterBody(const T* taps, std::size_t nTaps) : b(taps, taps + nTaps) {
        if (b.size() > inputHistory.capacity()) {
            inputHistory = gr::HistoryBuffer<T>(std::bit_ceil(b.size()));
        }
    }
    T processOne(T input) noexcept {
        inputHistory.push_front(input);
        return std::transform_reduce(std::execution::unseq, b.cbegin(), b.cend(), inputHistory.cbegin(), T{0}, std::plus<>{}, std::multiplies<>{});
    }
};

gr::algorithm::FFT<cf32>& threadLocalFft() {
    thread_local gr::algorithm::FFT<cf32> fft;
    return fft;
}
gr::algorithm::FFT<float, cf32>& threadLocalRealFft() {
    thread_local gr::algorithm::FFT<float, cf32> fft;
    return fft;
}
} // namespace

extern "C" {

int gr4ref_abi_version() { return 1; }

// ---- FFT -------------------------------------------------------------------------------------------------------
// out = FFT(in), n complex<float> interleaved; returns 0, or -1 if the reference threw
int gr4ref_fft_c2c_f32(const float* in, float* out, std::size_t n) {
    try {
        std::vector<cf32, gr::allocator::Aligned<cf32>> input(n), output(n);
        std::memcpy(input.data(), in, n * sizeof(cf32));
        threadLocalFft().compute(input, output);
        std::memcpy(out, output.data(), n * sizeof(cf32));
        return 0;
    } catch (...) {
        return -1;
    }
}

// batch of back-to-back transforms (used as CPU baseline; aligned zero-copy path of fft.hpp:182-189)
int gr4ref_fft_c2c_f32_batch(const float* in, float* out, std::size_t n, std::size_t batch) {
    try {
        std::vector<cf32, gr::allocator::Aligned<cf32>> input(n), output(n);
        auto&                                           fft = threadLocalFft();
        for (std::size_t i = 0; i < batch; ++i) {
            std::memcpy(input.data(), in + 2 * i * n, n * sizeof(cf32));
            fft.compute(input, output);
            std::memcpy(out + 2 * i * n, output.data(), n * sizeof(cf32));
        }
        return 0;
    } catch (...) {
        return -1;
    }
}

int gr4ref_window_f32(int type, std::size_t n, float beta, float* out) {
    try {
        std::vector<float> w(n);
        gr::algorithm::window::create(w, static_cast<gr::algorithm::window::Type>(type), beta);
        std::copy(w.begin(), w.end(), out);
        return 0;
    } catch (...) {
        return -1;
    }
}

int gr4ref_window_f64(int type, std::size_t n, double beta, double* out) {
    try {
        std::vector<double> w(n);
        gr::algorithm::window::create(w, static_cast<gr::algorithm::window::Type>(type), beta);
        std::copy(w.begin(), w.end(), out);
        return 0;
    } catch (...) {
        return -1;
    }
}

int gr4ref_magnitude_f32(const float* spectrum, std::size_t n, int outputInDb, int shift, float* out) {
    try {
        std::vector<cf32> X(reinterpret_cast<const cf32*>(spectrum), reinterpret_cast<const cf32*>(spectrum) + n);
        auto              mag = gr::algorithm::fft::computeMagnitudeSpectrum(X, std::vector<float>(n), gr::algorithm::fft::ConfigMagnitude{.computeHalfSpectrum = false, .outputInDb = outputInDb != 0, .shiftSpectrum = shift != 0});
        std::copy(mag.begin(), mag.end(), out);
        return 0;
    } catch (...) {
        return -1;
    }
}

int gr4ref_phase_f32(const float* spectrum, std::size_t n, int outputInDeg, int unwrap, int shift, float* out) {
    try {
        std::vector<cf32> X(reinterpret_cast<const cf32*>(spectrum), reinterpret_cast<const cf32*>(spectrum) + n);
        auto              ph = gr::algorithm::fft::computePhaseSpectrum(X, std::vector<float>(n), gr::algorithm::fft::ConfigPhase{.computeHalfSpectrum = false, .outputInDeg = outputInDeg != 0, .unwrapPhase = unwrap != 0, .shiftSpectrum = shift != 0});
        std::copy(ph.begin(), ph.end(), out);
        return 0;
    } catch (...) {
        return -1;
    }
}

int gr4ref_unwrap_phase_f64(double* phase, std::size_t n) {
    std::vector<double> p(phase, phase + n);
    gr::algorithm::fft::unwrapPhase(p);
    std::copy(p.begin(), p.end(), phase);
    return 0;
}

// the same block for T = float (computeFullSpectrum == false): the statements of processBulk / createDataset with the
// reference's own FFT<float>, computeMagnitudeSpectrum and computePhaseSpectrum; signals[c][0..3][nfft/2]
int gr4ref_fft_block_f32(const float* in, std::size_t nfft, std::size_t batch, const float* window, int outputInDb, int outputInDeg, int unwrapPhase, float* signals, float* ranges) {
    try {
        std::vector<float> inData(nfft);
        std::vector<cf32>  outData;
        std::vector<float> magnitude, phase;
        auto&              fft = threadLocalRealFft();
        for (std::size_t c = 0; c < batch; ++c) {
            std::copy_n(in + c * nfft, nfft, inData.begin());
            for (std::size_t i = 0; i < nfft; ++i) { // fft.hpp:155-162
                inData[i] *= window[i];
            }
            const auto spectrum = fft.compute(inData); // fft.hpp:164
            outData.assign(spectrum.begin(), spectrum.end());
            magnitude = gr::algorithm::fft::computeMagnitudeSpectrum(outData, magnitude, gr::algorithm::fft::ConfigMagnitude{.computeHalfSpectrum = true, .outputInDb = outputInDb != 0, .shiftSpectrum = true});
            phase     = gr::algorithm::fft::computePhaseSpectrum(outData, phase, gr::algorithm::fft::ConfigPhase{.computeHalfSpectrum = true, .outputInDeg = outputInDeg != 0, .unwrapPhase = unwrapPhase != 0, .shiftSpectrum = true});
            const std::size_t N = magnitude.size(); // fft.hpp:175
            if (N != nfft / 2 || phase.size() != N || outData.size() < N) {
                return -2;
            }
            float* sig = signals + c * 4 * N;
            std::copy(magnitude.begin(), magnitude.end(), sig);
            std::copy(phase.begin(), phase.end(), sig + N);
            const auto upper = std::span{outData}.last(N); // fft.hpp:213
            for (std::size_t i = 0; i < N; ++i) {
                sig[2 * N + i] = upper[i].real();
                sig[3 * N + i] = upper[i].imag();
            }
            if (ranges != nullptr) {
                for (std::size_t s = 0; s < 4; ++s) {
                    const auto mm               = std::minmax_element(sig + s * N, sig + (s + 1) * N);
                    ranges[(c * 4 + s) * 2 + 0] = *mm.first;
                    ranges[(c * 4 + s) * 2 + 1] = *mm.second;
                }
            }
        }
        return 0;
    } catch (...) {
        return -1;
    }
}

// FFT block (blocks/fourier/.../fft.hpp:147-171 + createDataset :173-250) for T = complex<float>, batch chunks of nfft:
// signals[c][0..3][nfft] = {magnitude (shifted), phase (shifted), Re, Im}; ranges[c][0..3][2] = {min,max} per signal
int gr4ref_fft_block_cf32(const float* in, std::size_t nfft, std::size_t batch, const float* window, int outputInDb, int outputInDeg, int unwrapPhase, float* signals, float* ranges) {
    try {
        std::vector<cf32, gr::allocator::Aligned<cf32>> inData(nfft), outData(nfft);
        std::vector<float>                              magnitude(nfft), phase(nfft);
        auto&                                           fft = threadLocalFft();
        for (std::size_t c = 0; c < batch; ++c) {
            std::memcpy(inData.data(), in + 2 * c * nfft, nfft * sizeof(cf32));
            for (std::size_t i = 0; i < nfft; ++i) { // fft.hpp:155-162
                inData[i].real(inData[i].real() * window[i]);
                inData[i].imag(inData[i].imag() * window[i]);
            }
            fft.compute(inData, outData);
            magnitude = gr::algorithm::fft::computeMagnitudeSpectrum(outData, magnitude, gr::algorithm::fft::ConfigMagnitude{.computeHalfSpectrum = false, .outputInDb = outputInDb != 0, .shiftSpectrum = true});
            phase     = gr::algorithm::fft::computePhaseSpectrum(outData, phase, gr::algorithm::fft::ConfigPhase{.computeHalfSpectrum = false, .outputInDeg = outputInDeg != 0, .unwrapPhase = unwrapPhase != 0, .shiftSpectrum = true});
            float* sig = signals + c * 4 * nfft;
            std::copy(magnitude.begin(), magnitude.end(), sig);
            std::copy(phase.begin(), phase.end(), sig + nfft);
            for (std::size_t i = 0; i < nfft; ++i) {
                sig[2 * nfft + i] = outData[i].real();
                sig[3 * nfft + i] = outData[i].imag();
            }
            if (ranges != nullptr) {
                for (std::size_t s = 0; s < 4; ++s) { // fft.hpp:222-225
                    const auto mm               = std::minmax_element(sig + s * nfft, sig + (s + 1) * nfft);
                    ranges[(c * 4 + s) * 2 + 0] = *mm.first;
                    ranges[(c * 4 + s) * 2 + 1] = *mm.second;
                }
            }
        }
        return 0;
    } catch (...) {
        return -1;
    }
}

// ---- FIR design ------------------------------------------------------------------------------------------------
int gr4ref_fir_generate_f32(std::size_t nTaps, int windowType, float fc, float beta, int normaliseDc, float* out) {
    try {
        auto coefficients = gr::filter::fir::generateCoefficients<float>(nTaps, static_cast<gr::algorithm::window::Type>(windowType), fc, beta);
        if (normaliseDc != 0) {
            const auto [ok, gain] = gr::filter::normaliseFilterCoefficients(coefficients, 0.f, 1.f);
            if (!ok) {
                return -2;
            }
        }
        std::copy(coefficients.b.begin(), coefficients.b.end(), out);
        return 0;
    } catch (...) {
        return -1;
    }
}

// fir::designFilter<T>(type, params, window); returns number of taps written (<= capacity) or negative on error
long gr4ref_fir_design_f32(int filterType, std::size_t order, double fLow, double fHigh, double fs, double gain, double attenuationDb, double beta, int windowType, float* out, std::size_t capacity) {
    try {
        gr::filter::FilterParameters params{.order = order, .fLow = fLow, .fHigh = fHigh, .gain = gain, .attenuationDb = attenuationDb, .beta = beta, .fs = fs};
        const auto coefficients = gr::filter::fir::designFilter<float>(static_cast<gr::filter::Type>(filterType), params, static_cast<gr::algorithm::window::Type>(windowType));
        if (coefficients.b.size() > capacity) {
            return -static_cast<long>(coefficients.b.size());
        }
        std::copy(coefficients.b.begin(), coefficients.b.end(), out);
        return static_cast<long>(coefficients.b.size());
    } catch (...) {
        return -1;
    }
}

long gr4ref_fir_design_f64(int filterType, std::size_t order, double fLow, double fHigh, double fs, double gain, double attenuationDb, double beta, int windowType, double* out, std::size_t capacity) {
    try {
        gr::filter::FilterParameters params{.order = order, .fLow = fLow, .fHigh = fHigh, .gain = gain, .attenuationDb = attenuationDb, .beta = beta, .fs = fs};
        const auto coefficients = gr::filter::fir::designFilter<double>(static_cast<gr::filter::Type>(filterType), params, static_cast<gr::algorithm::window::Type>(windowType));
        if (coefficients.b.size() > capacity) {
            return -static_cast<long>(coefficients.b.size());
        }
        std::copy(coefficients.b.begin(), coefficients.b.end(), out);
        return static_cast<long>(coefficients.b.size());
    } catch (...) {
        return -1;
    }
}

double gr4ref_fir_magnitude_response_f64(const double* b, std::size_t nTaps, double normalisedFrequency) {
    gr::filter::FilterCoefficients<double> coefficients{.b = std::vector<double>(b, b + nTaps)};
    return gr::filter::calculateResponse<gr::filter::Frequency::Normalised, gr::filter::ResponseType::Magnitude>(normalisedFrequency, coefficients);
}

// ---- FIR ----------------------------------------------------------------------------------------------------------
// fir_filter<float>: zero initial state, n samples
int gr4ref_fir_f32(const float* taps, std::size_t nTaps, const float* in, float* out, std::size_t n) {
    FirFilterBody<float> filter(taps, nTaps);
    for (std::size_t i = 0; i < n; ++i) {
        out[i] = filter.processOne(in[i]);
    }
    return 0;
}

int gr4ref_fir_f64(const double* taps, std::size_t nTaps, const double* in, double* out, std::size_t n) {
    FirFilterBody<double> filter(taps, nTaps);
    for (std::size_t i = 0; i < n; ++i) {
        out[i] = filter.processOne(in[i]);
    }
    return 0;
}

// the build target: complex<float> stream = two independent real fir_filter<float> on re and im
int gr4ref_fir_cf32(const float* taps, std::size_t nTaps, const float* in, float* out, std::size_t n) {
    FirFilterBody<float> re(taps, nTaps), im(taps, nTaps);
    for (std::size_t i = 0; i < n; ++i) {
        out[2 * i]     = re.processOne(in[2 * i]);
        out[2 * i + 1] = im.processOne(in[2 * i + 1]);
    }
    return 0;
}

// BasicDecimatingFilter body (time_domain_filter.hpp:190-204) over gr::filter::Filter<float> built from FIR taps;
// complex stream = re/im independently. n must be a multiple of decimate; writes n/decimate samples.
int gr4ref_fir_decim_cf32(const float* taps, std::size_t nTaps, std::size_t decimate, const float* in, float* out, std::size_t n) {
    gr::filter::FilterCoefficients<float> coefficients{.b = std::vector<float>(taps, taps + nTaps)};
    gr::filter::Filter<float>             re(coefficients), im(coefficients);
    std::size_t                           outIdx = 0;
    for (std::size_t i = 0; i < n; ++i) {
        const float yr = re.processOne(in[2 * i]);
        const float yi = im.processOne(in[2 * i + 1]);
        if (i % decimate == 0) {
            out[2 * outIdx]     = yr;
            out[2 * outIdx + 1] = yi;
            ++outIdx;
        }
    }
    return 0;
}

int gr4ref_fir_decim_f32(const float* taps, std::size_t nTaps, std::size_t decimate, const float* in, float* out, std::size_t n) {
    gr::filter::FilterCoefficients<float> coefficients{.b = std::vector<float>(taps, taps + nTaps)};
    gr::filter::Filter<float>             filter(coefficients);
    std::size_t                           outIdx = 0;
    for (std::size_t i = 0; i < n; ++i) {
        const float y = filter.processOne(in[i]);
        if (i % decimate == 0) {
            out[outIdx++] = y;
        }
    }
    return 0;
}

// ---- math / mixer (block bodies restated over std::complex; see header) ----------------------------------------------
// op: 0 add, 1 subtract, 2 multiply, 3 divide   (Math.hpp:38-56, non-SIMD branch `op()(a, value)`)
int gr4ref_mathop_const_cf32(int op, const float* in, float* out, std::size_t n, float valueRe, float valueIm) {
    const cf32  value(valueRe, valueIm);
    const cf32* a = reinterpret_cast<const cf32*>(in);
    cf32*       y = reinterpret_cast<cf32*>(out);
    switch (op) {
    case 0: std::transform(a, a + n, y, [value](cf32 v) { return std::plus<cf32>()(v, value); }); break;
    case 1: std::transform(a, a + n, y, [value](cf32 v) { return std::minus<cf32>()(v, value); }); break;
    case 2: std::transform(a, a + n, y, [value](cf32 v) { return std::multiplies<cf32>()(v, value); }); break;
    case 3: std::transform(a, a + n, y, [value](cf32 v) { return std::divides<cf32>()(v, value); }); break;
    default: return -1;
    }
    return 0;
}

// MathOpMultiPortImpl::processBulk (Math.hpp:100-107): copy in[0], then fold remaining inputs
int gr4ref_mathop_multi_cf32(int op, const float* const* ins, std::size_t nInputs, float* out, std::size_t n) {
    cf32* y = reinterpret_cast<cf32*>(out);
    std::copy(reinterpret_cast<const cf32*>(ins[0]), reinterpret_cast<const cf32*>(ins[0]) + n, y);
    for (std::size_t k = 1; k < nInputs; ++k) {
        const cf32* b = reinterpret_cast<const cf32*>(ins[k]);
        switch (op) {
        case 0: std::transform(y, y + n, b, y, std::plus<cf32>()); break;
        case 1: std::transform(y, y + n, b, y, std::minus<cf32>()); break;
        case 2: std::transform(y, y + n, b, y, std::multiplies<cf32>()); break;
        case 3: std::transform(y, y + n, b, y, std::divides<cf32>()); break;
        default: return -1;
        }
    }
    return 0;
}

// Rotator<complex<float>>::processOne (Rotator.hpp:51-61); phase is in/out state
int gr4ref_rotator_cf32(const float* in, float* out, std::size_t n, float phaseIncrement, float* accumulatedPhase) {
    float       phase = *accumulatedPhase;
    const cf32* x     = reinterpret_cast<const cf32*>(in);
    cf32*       y     = reinterpret_cast<cf32*>(out);
    for (std::size_t i = 0; i < n; ++i) {
        phase += phaseIncrement;
        if (phase > 2.f * std::numbers::pi_v<float>) {
            phase -= 2.f * std::numbers::pi_v<float>;
        } else if (phase < 0.f) {
            phase += 2.f * std::numbers::pi_v<float>;
        }
        y[i] = x[i] * cf32(std::cos(phase), std::sin(phase));
    }
    *accumulatedPhase = phase;
    return 0;
}

} // extern "C"
