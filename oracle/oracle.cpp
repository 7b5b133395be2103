// TEST INFRASTRUCTURE ONLY -- CPU oracle for the gnuradio4_b200 hot path.
//
// A CPU restatement, in our own words, of the arithmetic the fair-acc/gnuradio4 reference performs on the
// FIR -> FFT streaming path. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library, and only as the checker / reported baseline; the product
// (gnuradio4_b200/csrc + include/gr4b200.h) never links, loads or calls it.
//
// Pinning: every function here is checked in tests/test_oracle.py against (i) the reference's own known-answer
// tests and (ii) oracle/_ref/libgr4ref.so, which is the reference's own source compiled in place (ref_harness.cpp).
// Exceptions (parity unpinned, no reference implementation exists): oracle_pfb_channelizer_cf32.
//
// All paths are relative to /root/reference.
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numbers>
#include <vector>

using cf32 = std::complex<float>;
using cf64 = std::complex<double>;

namespace {

// ------------------------------------------------------------------------------------------------------------------
// windows: algorithm/include/gnuradio-4.0/algorithm/fourier/window.hpp:71-183 (enum order :35)
// ------------------------------------------------------------------------------------------------------------------
enum WindowType { None = 0, Rectangular, Hamming, Hann, HannExp, Blackman, Nuttall, BlackmanHarris, BlackmanNuttall, FlatTop, Exponential, Kaiser };

template<typename T>
T besselI0(T x) { // window.hpp:42-56: power series, stop when term^2 <= sum * eps
    T   sum = 1, term = 1;
    int k   = 1;
    const T half = x / 2;
    do {
        term *= half / static_cast<T>(k);
        sum += term * term;
        ++k;
    } while (term * term > sum * std::numeric_limits<T>::epsilon());
    return sum;
}

template<typename T>
int makeWindow(int type, std::size_t n, T beta, T* w) {
    if (n == 0) {
        return 0;
    }
    const T step = (2 * std::numbers::pi_v<T>) / static_cast<T>(n - 1); // all cosine windows use the N-1 denominator
    switch (type) {
    case None:
    case Rectangular: std::fill(w, w + n, T(1)); return 0;
    case Hamming: // :86-91 (0.53836 / 0.46164 variant)
        for (std::size_t i = 0; i < n; ++i) {
            w[i] = static_cast<T>(0.53836) - static_cast<T>(0.46164) * std::cos(step * static_cast<T>(i));
        }
        return 0;
    case Hann: // :93-98
        for (std::size_t i = 0; i < n; ++i) {
            w[i] = static_cast<T>(.5) - static_cast<T>(.5) * std::cos(step * static_cast<T>(i));
        }
        return 0;
    case HannExp: // :100-104  sin^2 via pow
        for (std::size_t i = 0; i < n; ++i) {
            w[i] = std::pow(std::sin(step * static_cast<T>(i)), static_cast<T>(2.));
        }
        return 0;
    case Blackman: // :106-114
        for (std::size_t i = 0; i < n; ++i) {
            const T ai = step * static_cast<T>(i);
            w[i]       = static_cast<T>(0.42) - static_cast<T>(0.5) * std::cos(ai) + static_cast<T>(0.08) * std::cos(static_cast<T>(2.) * ai);
        }
        return 0;
    case Nuttall: // :116-126
        for (std::size_t i = 0; i < n; ++i) {
            const T ai = step * static_cast<T>(i);
            w[i]       = static_cast<T>(0.355768) - static_cast<T>(0.487396) * std::cos(ai) + static_cast<T>(0.144232) * std::cos(2 * ai) - static_cast<T>(0.012604) * std::cos(3 * ai);
        }
        return 0;
    case BlackmanHarris: // :128-138
        for (std::size_t i = 0; i < n; ++i) {
            const T ai = step * static_cast<T>(i);
            w[i]       = static_cast<T>(0.35875) - static_cast<T>(0.48829) * std::cos(ai) + static_cast<T>(0.14128) * std::cos(2 * ai) - static_cast<T>(0.01168) * std::cos(3 * ai);
        }
        return 0;
    case BlackmanNuttall: // :140-148
        for (std::size_t i = 0; i < n; ++i) {
            const T ai = step * static_cast<T>(i);
            w[i]       = static_cast<T>(0.3635819) - static_cast<T>(0.4891775) * std::cos(ai) + static_cast<T>(0.1365995) * std::cos(static_cast<T>(2.) * ai) - static_cast<T>(0.0106411) * std::cos(static_cast<T>(3.) * ai);
        }
        return 0;
    case FlatTop: // :150-160
        for (std::size_t i = 0; i < n; ++i) {
            const T ai = step * static_cast<T>(i);
            w[i]       = static_cast<T>(1.0) - static_cast<T>(1.93) * std::cos(ai) + static_cast<T>(1.29) * std::cos(2 * ai) - static_cast<T>(0.388) * std::cos(3 * ai) + static_cast<T>(0.032) * std::cos(4 * ai);
        }
        return 0;
    case Exponential: { // :162-168  exp(i / (3 n)) / exp(0)
        const T e0 = std::exp(static_cast<T>(0.));
        const T a  = static_cast<T>(3.) * static_cast<T>(n);
        for (std::size_t i = 0; i < n; ++i) {
            w[i] = std::exp(static_cast<T>(i) / a) / e0;
        }
        return 0;
    }
    case Kaiser: { // :170-186
        if (beta < 0 || n <= 1) {
            return -1; // reference throws std::invalid_argument
        }
        const T factor = static_cast<T>(1) / static_cast<T>(n - 1);
        const T i0Beta = besselI0(beta);
        for (std::size_t i = 0; i < n; ++i) {
            const T term = (static_cast<T>(2 * i) * factor) - static_cast<T>(1);
            w[i]         = besselI0(beta * std::sqrt(std::abs(static_cast<T>(1) - term * term))) / i0Beta;
        }
        return 0;
    }
    default: return -1;
    }
}

// ------------------------------------------------------------------------------------------------------------------
// FIR design: algorithm/include/gnuradio-4.0/algorithm/filter/FilterTool.hpp:964-976 (generateCoefficients),
// :376-404 (calculateResponse), :415-423 (normaliseFilterCoefficients), :985-1004 (tap-count estimate), :1007-1071
// ------------------------------------------------------------------------------------------------------------------
template<typename T>
int firGenerate(std::size_t n, int window, T fc, T beta, T* b) {
    const T M = static_cast<T>(n - 1) / static_cast<T>(2);
    if (makeWindow<T>(window, n, beta, b) != 0) {
        return -1;
    }
    for (std::size_t i = 0; i < n; ++i) {
        const T x    = static_cast<T>(2) * fc * (static_cast<T>(i) - M);
        const T a    = std::numbers::pi_v<T>;
        const T sinc = x == static_cast<T>(0) ? static_cast<T>(1) : std::sin(a * x) / (a * x);
        b[i]         = b[i] * static_cast<T>(2) * fc * sinc;
    }
    return 0;
}

template<typename T>
T firMagnitudeResponse(const T* b, std::size_t n, T normalisedFrequency) { // |B(e^{jw})| / |A| with a = {1}
    using C = std::complex<T>;
    const C iOmega = std::polar(static_cast<T>(1), static_cast<T>(2) * std::numbers::pi_v<T> * normalisedFrequency);
    C       numerator(0);
    for (std::size_t k = 0; k < n; ++k) {
        numerator = numerator + b[k] * static_cast<C>(std::pow(iOmega, -static_cast<int>(k)));
    }
    const C denominator = C(0) + static_cast<T>(1) * static_cast<C>(std::pow(iOmega, 0));
    return static_cast<T>(1) * std::abs(numerator / denominator);
}

template<typename T>
bool firNormalise(T* b, std::size_t n, T normalisedFrequency, T targetGain) {
    const T magnitude = firMagnitudeResponse(b, n, normalisedFrequency);
    if (magnitude == 0) {
        return false;
    }
    for (std::size_t i = 0; i < n; ++i) {
        b[i] = b[i] * targetGain / magnitude;
    }
    return true;
}

std::size_t kaiserTapEstimate(double attenuationDb, double transitionWidth) { // :985-991, always odd
    auto n = static_cast<std::size_t>(std::ceil((attenuationDb - 8.0) / (2.285 * transitionWidth)));
    return n % 2 == 0 ? n + 1 : n;
}

double requiredTransitionWidth(int type, std::size_t order, double fLow, double fHigh, double fs) { // :993-1004
    const double w = 0.1 / static_cast<double>(order);
    switch (type) {
    case 0: return std::min(w, std::min(std::abs(fLow / fs), std::abs(0.5 - fLow / fs)));
    case 1: return std::min(w, std::abs(fHigh / fs));
    case 2: return std::min(w, std::min(std::abs(fLow / fs), std::abs(0.5 - fHigh / fs)));
    case 3: return std::min(w, std::min(std::abs(0.5 - fHigh / fs), std::min(fLow, 0.5 * std::abs(fHigh - fLow)) / fs));
    default: return 0.;
    }
}

template<typename T>
long firDesign(int type, std::size_t order, double fLow, double fHigh, double fs, double gain, double attenuationDb, double beta, int window, T* out, std::size_t capacity) {
    const std::size_t n = kaiserTapEstimate(attenuationDb, 2. * std::numbers::pi * requiredTransitionWidth(type, order, fLow, fHigh, fs));
    if (n > capacity) {
        return -static_cast<long>(n);
    }
    std::vector<T> b(n), hp(n);
    bool           ok = false;
    switch (type) {
    case 0: // LOWPASS :1012-1020
        if (firGenerate<T>(n, window, static_cast<T>(fLow / fs), static_cast<T>(beta), b.data()) != 0) {
            return -1;
        }
        ok = firNormalise<T>(b.data(), n, static_cast<T>(0), static_cast<T>(gain));
        break;
    case 1: // HIGHPASS :1022-1037: mirrored low-pass, spectral inversion, normalise at 0.48
        if (firGenerate<T>(n, window, static_cast<T>(0.5 - fHigh / fs), static_cast<T>(beta), b.data()) != 0) {
            return -1;
        }
        for (std::size_t i = 0; i < n; ++i) {
            b[i] *= (i % 2 == 0 ? 1 : -1);
        }
        ok = firNormalise<T>(b.data(), n, static_cast<T>(0.48), static_cast<T>(gain));
        break;
    case 2: // BANDPASS :1039-1050
        if (firGenerate<T>(n, window, static_cast<T>(fLow / fs), static_cast<T>(beta), b.data()) != 0 || firGenerate<T>(n, window, static_cast<T>(fHigh / fs), static_cast<T>(beta), hp.data()) != 0) {
            return -1;
        }
        for (std::size_t i = 0; i < n; ++i) {
            b[i] = b[i] - hp[i];
        }
        ok = firNormalise<T>(b.data(), n, static_cast<T>(std::sqrt(fHigh * fLow) / fs), static_cast<T>(gain));
        break;
    case 3: // BANDSTOP :1052-1069
        if (firGenerate<T>(n, window, static_cast<T>(fLow / fs), static_cast<T>(beta), b.data()) != 0 || firGenerate<T>(n, window, static_cast<T>(fHigh / fs), static_cast<T>(beta), hp.data()) != 0) {
            return -1;
        }
        for (std::size_t i = 0; i < n; ++i) {
            b[i] -= hp[i];
            if (n % 2 != 0 && i == (n - 1) / 2) {
                b[i] = 1 - b[i];
            }
        }
        ok = firNormalise<T>(b.data(), n, static_cast<T>(0), static_cast<T>(gain));
        break;
    default: return -1;
    }
    if (!ok) {
        return -2;
    }
    std::copy(b.begin(), b.end(), out);
    return static_cast<long>(n);
}

// ------------------------------------------------------------------------------------------------------------------
// FIR: blocks/filter/include/gnuradio-4.0/filter/time_domain_filter.hpp:44-47
//   y[n] = transform_reduce(unseq, b, history, 0, plus<>, multiplies<>) with history[k] = x[n-k], x[<0] = 0.
// Summation order is fixed by libstdc++'s __simd_transform_reduce for a non-std::plus<T> functor
// (/usr/include/c++/13/pstl/unseq_backend_simd.h:455-505): 64-byte lane block => L = 64/sizeof(T) lanes;
//   n <= 2L : plain left fold  init + f(0) + f(1) + ...
//   n  > 2L : lane[j] = f(j) + f(L+j);  lane[j] += f(i+j) for i = 2L, 3L, ... < L*(n/L);  remainder lane[j] += f(L*(n/L)+j);
//             result = ((init + lane[0]) + lane[1]) + ... + lane[L-1]
// products and sums are separately rounded (reference release flags have no -march => no FMA contraction).
// `hist` points at x[n] with x[n-k] = hist[-k].
// ------------------------------------------------------------------------------------------------------------------
template<typename T>
inline T firDot(const T* b, std::size_t nTaps, const T* hist, std::ptrdiff_t stride) {
    constexpr std::size_t L = 64 / sizeof(T);
    auto                  f = [&](std::size_t k) -> T { return b[k] * hist[-static_cast<std::ptrdiff_t>(k) * stride]; };
    T                     init{0};
    if (nTaps > 2 * L) {
        T lane[L];
        for (std::size_t j = 0; j < L; ++j) {
            lane[j] = f(j) + f(L + j);
        }
        const std::size_t lastBlock = L * (nTaps / L);
        for (std::size_t i = 2 * L; i < lastBlock; i += L) {
            for (std::size_t j = 0; j < L; ++j) {
                lane[j] = lane[j] + f(i + j);
            }
        }
        for (std::size_t j = 0; j < nTaps - lastBlock; ++j) {
            lane[j] = lane[j] + f(lastBlock + j);
        }
        for (std::size_t j = 0; j < L; ++j) {
            init = init + lane[j];
        }
    } else {
        for (std::size_t k = 0; k < nTaps; ++k) {
            init = init + f(k);
        }
    }
    return init;
}

// streaming FIR over `channels` interleaved real sub-streams (1 = real stream, 2 = complex<T> as re/im pair).
// state: (nTaps-1)*channels values = the last nTaps-1 input samples of the previous chunk (oldest first), zeros at
// stream start; updated on return. decimate: keep outputs with (i % decimate == 0), i counted from chunk start
// (time_domain_filter.hpp:190-204).
template<typename T>
void firStream(const T* taps, std::size_t nTaps, std::size_t channels, std::size_t decimate, const T* in, T* out, std::size_t n, T* state) {
    const std::size_t halo = nTaps - 1;
    std::vector<T>    work((halo + n) * channels);
    if (state != nullptr) {
        std::copy(state, state + halo * channels, work.begin());
    } else {
        std::fill(work.begin(), work.begin() + static_cast<std::ptrdiff_t>(halo * channels), T(0));
    }
    std::copy(in, in + n * channels, work.begin() + static_cast<std::ptrdiff_t>(halo * channels));
    std::size_t o = 0;
    for (std::size_t i = 0; i < n; ++i) {
        if (i % decimate != 0) {
            continue; // the reference computes and discards these (same retained values)
        }
        for (std::size_t c = 0; c < channels; ++c) {
            out[o * channels + c] = firDot<T>(taps, nTaps, work.data() + (halo + i) * channels + c, static_cast<std::ptrdiff_t>(channels));
        }
        ++o;
    }
    if (state != nullptr) {
        std::copy(work.end() - static_cast<std::ptrdiff_t>(halo * channels), work.end(), state);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// FFT: algorithm/include/gnuradio-4.0/algorithm/fourier/fft.hpp:113-153 -> SimdFFT.hpp:491-690.
// Contract restated: unnormalised forward DFT X[k] = sum_n x[n] exp(-j 2 pi k n / N), natural order, computed in T.
// The restatement is a Stockham autosort radix-4 (+ one radix-2 pass when log2 N is odd) FFT with twiddles generated
// in T via cos/sin like SimdFFT.hpp:385-461; it is NOT bit-identical to SimdFFT's 4-lane pass order -- FFT parity is a
// spectral tolerance (tests state it) and the restatement is pinned against _ref within that tolerance.
// Non power-of-two sizes: plain O(N^2) DFT in double (small N only; reference uses mixed radix / Bluestein).
// ------------------------------------------------------------------------------------------------------------------
template<typename T>
void fftPow2(const std::complex<T>* in, std::complex<T>* out, std::size_t N) {
    using C = std::complex<T>;
    std::vector<C> a(in, in + N), bbuf(N);
    C*             x = a.data();
    C*             y = bbuf.data();
    std::size_t    n = N, s = 1; // n: current sub-transform length, s: stride
    const T        theta0 = -(2 * std::numbers::pi_v<T>) / static_cast<T>(N);
    while (n > 1) {
        if (n % 4 == 0) {
            const std::size_t n1 = n / 4;
            for (std::size_t p = 0; p < n1; ++p) {
                const T arg = theta0 * static_cast<T>(s * p) ; // exp(-j 2 pi p / n) = exp(theta0 * s * p) because n * s == N
                const C w1(std::cos(arg), std::sin(arg));
                const C w2(std::cos(2 * arg), std::sin(2 * arg));
                const C w3(std::cos(3 * arg), std::sin(3 * arg));
                for (std::size_t q = 0; q < s; ++q) {
                    const C a0 = x[q + s * (p + 0 * n1)];
                    const C a1 = x[q + s * (p + 1 * n1)];
                    const C a2 = x[q + s * (p + 2 * n1)];
                    const C a3 = x[q + s * (p + 3 * n1)];
                    const C t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3;
                    const C d  = a1 - a3;
                    const C t3(d.imag(), -d.real()); // -j * (a1 - a3)
                    y[q + s * (4 * p + 0)] = t0 + t2;
                    y[q + s * (4 * p + 1)] = (t1 + t3) * w1;
                    y[q + s * (4 * p + 2)] = (t0 - t2) * w2;
                    y[q + s * (4 * p + 3)] = (t1 - t3) * w3;
                }
            }
            n /= 4;
            s *= 4;
        } else {
            const std::size_t n1 = n / 2;
            for (std::size_t p = 0; p < n1; ++p) {
                const T arg = theta0 * static_cast<T>(s * p);
                const C w(std::cos(arg), std::sin(arg));
                for (std::size_t q = 0; q < s; ++q) {
                    const C a0             = x[q + s * p];
                    const C a1             = x[q + s * (p + n1)];
                    y[q + s * (2 * p + 0)] = a0 + a1;
                    y[q + s * (2 * p + 1)] = (a0 - a1) * w;
                }
            }
            n /= 2;
            s *= 2;
        }
        std::swap(x, y);
    }
    std::copy(x, x + N, out);
}

template<typename T>
void fftAny(const std::complex<T>* in, std::complex<T>* out, std::size_t N) {
    if (N == 0) {
        return;
    }
    if ((N & (N - 1)) == 0) {
        fftPow2<T>(in, out, N);
        return;
    }
    for (std::size_t k = 0; k < N; ++k) {
        cf64 acc(0, 0);
        for (std::size_t n = 0; n < N; ++n) {
            const double arg = -2.0 * std::numbers::pi * static_cast<double>((k * n) % N) / static_cast<double>(N);
            acc += cf64(in[n]) * cf64(std::cos(arg), std::sin(arg));
        }
        out[k] = std::complex<T>(static_cast<T>(acc.real()), static_cast<T>(acc.imag()));
    }
}

// fft_common.hpp:22-56
template<typename T>
void magnitudeSpectrum(const std::complex<T>* X, std::size_t N, bool dB, bool shift, T* out) {
    for (std::size_t i = 0; i < N; ++i) {
        const T mag = std::hypot(X[i].real(), X[i].imag()) * T(2.) / static_cast<T>(N);
        if (dB && mag > T(0)) {
            out[i] = T(20.) * std::log10(mag);
        } else if (dB) {
            out[i] = std::numeric_limits<T>::lowest();
        } else {
            out[i] = mag;
        }
    }
    if (shift) {
        std::rotate(out, out + static_cast<std::ptrdiff_t>(N) / 2, out + N);
    }
}

template<typename T>
void unwrapInPlace(T* phase, std::size_t n) { // fft_common.hpp:72-90
    if (n == 0) {
        return;
    }
    const T pi   = std::numbers::pi_v<T>;
    T       prev = phase[0];
    for (std::size_t i = 1; i < n; ++i) {
        T cur  = phase[i];
        T diff = cur - prev;
        while (diff > pi) {
            cur -= 2 * pi;
            diff = cur - prev;
        }
        while (diff < -pi) {
            cur += 2 * pi;
            diff = cur - prev;
        }
        prev     = cur;
        phase[i] = cur;
    }
}

// fft_common.hpp:93-123
template<typename T>
void phaseSpectrum(const std::complex<T>* X, std::size_t N, bool deg, bool unwrap, bool shift, T* out) {
    for (std::size_t i = 0; i < N; ++i) {
        out[i] = std::atan2(X[i].imag(), X[i].real());
    }
    if (unwrap) {
        unwrapInPlace(out, N);
    }
    if (deg) {
        for (std::size_t i = 0; i < N; ++i) {
            out[i] = out[i] * static_cast<T>(180.) * std::numbers::inv_pi_v<T>;
        }
    }
    if (shift) {
        std::rotate(out, out + static_cast<std::ptrdiff_t>(N) / 2, out + N);
    }
}

} // namespace

// element loops of the sample-format converters (ConverterBlocks.hpp:233-277), used by the entry points below
template<typename R>
static void interleavedToComplex(const R* in, cf32* out, std::size_t n) {
    for (std::size_t i = 0; i < n; ++i) {
        out[i] = cf32{static_cast<float>(in[2 * i]), static_cast<float>(in[2 * i + 1])};
    }
}
template<typename R>
static void complexToInterleaved(const cf32* in, R* out, std::size_t n) {
    for (std::size_t i = 0; i < n; ++i) {
        out[2 * i]     = static_cast<R>(in[i].real());
        out[2 * i + 1] = static_cast<R>(in[i].imag());
    }
}

extern "C" {

int oracle_abi_version() { return 1; }

int oracle_window_f32(int type, std::size_t n, float beta, float* out) { return makeWindow<float>(type, n, beta, out); }
int oracle_window_f64(int type, std::size_t n, double beta, double* out) { return makeWindow<double>(type, n, beta, out); }

int oracle_fir_generate_f32(std::size_t nTaps, int windowType, float fc, float beta, int normaliseDc, float* out) {
    if (firGenerate<float>(nTaps, windowType, fc, beta, out) != 0) {
        return -1;
    }
    if (normaliseDc != 0 && !firNormalise<float>(out, nTaps, 0.f, 1.f)) {
        return -2;
    }
    return 0;
}
long oracle_fir_design_f32(int filterType, std::size_t order, double fLow, double fHigh, double fs, double gain, double attenuationDb, double beta, int windowType, float* out, std::size_t capacity) { return firDesign<float>(filterType, order, fLow, fHigh, fs, gain, attenuationDb, beta, windowType, out, capacity); }
long oracle_fir_design_f64(int filterType, std::size_t order, double fLow, double fHigh, double fs, double gain, double attenuationDb, double beta, int windowType, double* out, std::size_t capacity) { return firDesign<double>(filterType, order, fLow, fHigh, fs, gain, attenuationDb, beta, windowType, out, capacity); }
double oracle_fir_magnitude_response_f64(const double* b, std::size_t nTaps, double normalisedFrequency) { return firMagnitudeResponse<double>(b, nTaps, normalisedFrequency); }

// state may be NULL (zero history, not written back)
int oracle_fir_f32(const float* taps, std::size_t nTaps, const float* in, float* out, std::size_t n, float* state) {
    firStream<float>(taps, nTaps, 1, 1, in, out, n, state);
    return 0;
}
int oracle_fir_f64(const double* taps, std::size_t nTaps, const double* in, double* out, std::size_t n, double* state) {
    firStream<double>(taps, nTaps, 1, 1, in, out, n, state);
    return 0;
}
int oracle_fir_cf32(const float* taps, std::size_t nTaps, const float* in, float* out, std::size_t n, float* state) {
    firStream<float>(taps, nTaps, 2, 1, in, out, n, state);
    return 0;
}
int oracle_fir_decim_cf32(const float* taps, std::size_t nTaps, std::size_t decimate, const float* in, float* out, std::size_t n, float* state) {
    if (decimate == 0 || n % decimate != 0) {
        return -1;
    }
    firStream<float>(taps, nTaps, 2, decimate, in, out, n, state);
    return 0;
}
int oracle_fir_decim_f32(const float* taps, std::size_t nTaps, std::size_t decimate, const float* in, float* out, std::size_t n, float* state) {
    if (decimate == 0 || n % decimate != 0) {
        return -1;
    }
    firStream<float>(taps, nTaps, 1, decimate, in, out, n, state);
    return 0;
}
// Decimator<T> (time_domain_filter.hpp:234-244): keep i % decim == 0
int oracle_decimate_cf32(const float* in, float* out, std::size_t n, std::size_t decim) {
    std::size_t o = 0;
    for (std::size_t i = 0; i < n; ++i) {
        if (i % decim == 0) {
            out[2 * o]     = in[2 * i];
            out[2 * o + 1] = in[2 * i + 1];
            ++o;
        }
    }
    return 0;
}

int oracle_fft_c2c_f32(const float* in, float* out, std::size_t n, std::size_t batch) {
    for (std::size_t b = 0; b < batch; ++b) {
        fftAny<float>(reinterpret_cast<const cf32*>(in) + b * n, reinterpret_cast<cf32*>(out) + b * n, n);
    }
    return 0;
}
// float input, transform carried out in double: the error-norm yardstick for the float paths
int oracle_fft_c2c_f32_via_f64(const float* in, double* out, std::size_t n, std::size_t batch) {
    std::vector<cf64> x(n), X(n);
    for (std::size_t b = 0; b < batch; ++b) {
        for (std::size_t i = 0; i < n; ++i) {
            x[i] = cf64(in[2 * (b * n + i)], in[2 * (b * n + i) + 1]);
        }
        fftAny<double>(x.data(), X.data(), n);
        std::memcpy(out + 2 * b * n, X.data(), n * sizeof(cf64));
    }
    return 0;
}

int oracle_magnitude_f32(const float* spectrum, std::size_t n, int outputInDb, int shift, float* out) {
    magnitudeSpectrum<float>(reinterpret_cast<const cf32*>(spectrum), n, outputInDb != 0, shift != 0, out);
    return 0;
}
int oracle_phase_f32(const float* spectrum, std::size_t n, int outputInDeg, int unwrap, int shift, float* out) {
    phaseSpectrum<float>(reinterpret_cast<const cf32*>(spectrum), n, outputInDeg != 0, unwrap != 0, shift != 0, out);
    return 0;
}
int oracle_unwrap_phase_f64(double* phase, std::size_t n) {
    unwrapInPlace<double>(phase, n);
    return 0;
}

// FFT block, blocks/fourier/include/gnuradio-4.0/fourier/fft.hpp:147-171 (+ createDataset :173-250), T = complex<float>:
// per chunk of nfft samples: x*w (re and im separately), FFT, magnitude (2/N, shifted), phase (shifted), Re, Im (unshifted);
// signals[c][4][nfft], ranges[c][4][2] = per-signal {min,max} (ranges may be NULL)
int oracle_fft_block_cf32(const float* in, std::size_t nfft, std::size_t batch, const float* window, int outputInDb, int outputInDeg, int unwrapPhase, float* signals, float* ranges) {
    std::vector<cf32> x(nfft), X(nfft);
    for (std::size_t c = 0; c < batch; ++c) {
        for (std::size_t i = 0; i < nfft; ++i) {
            x[i] = cf32(in[2 * (c * nfft + i)] * window[i], in[2 * (c * nfft + i) + 1] * window[i]);
        }
        fftAny<float>(x.data(), X.data(), nfft);
        float* sig = signals + c * 4 * nfft;
        magnitudeSpectrum<float>(X.data(), nfft, outputInDb != 0, true, sig);
        phaseSpectrum<float>(X.data(), nfft, outputInDeg != 0, unwrapPhase != 0, true, sig + nfft);
        for (std::size_t i = 0; i < nfft; ++i) {
            sig[2 * nfft + i] = X[i].real();
            sig[3 * nfft + i] = X[i].imag();
        }
        if (ranges != nullptr) {
            for (std::size_t s = 0; s < 4; ++s) {
                const auto mm               = std::minmax_element(sig + s * nfft, sig + (s + 1) * nfft);
                ranges[(c * 4 + s) * 2 + 0] = *mm.first;
                ranges[(c * 4 + s) * 2 + 1] = *mm.second;
            }
        }
    }
    return 0;
}

// FFT block on REAL input, T = float (blocks/fourier/include/gnuradio-4.0/fourier/fft.hpp:147-250 with
// computeFullSpectrum == false): the window multiplies the real samples (:155-162), compute() returns the full N-bin
// spectrum (algorithm/.../fourier/fft.hpp:214-258), magnitude and phase are taken from bins [0, N/2) with the 2/N scaling
// and WITHOUT the fft-shift (fft_common.hpp:30,52,99,118), and createDataset copies Re / Im from the LAST N/2 bins of the
// spectrum, i.e. X[N/2 .. N-1] (fft.hpp:212-217, `std::span{_outData}.last(N)`). signals[c][0..3][nfft/2].
int oracle_fft_block_f32(const float* in, std::size_t nfft, std::size_t batch, const float* window, int outputInDb, int outputInDeg, int unwrapPhase, float* signals, float* ranges) {
    const std::size_t  half = nfft / 2;
    std::vector<cf32>  x(nfft), X(nfft);
    std::vector<float> full(nfft);
    for (std::size_t c = 0; c < batch; ++c) {
        for (std::size_t i = 0; i < nfft; ++i) {
            x[i] = cf32(in[c * nfft + i] * window[i], 0.f);
        }
        fftAny<float>(x.data(), X.data(), nfft);
        X[0]    = cf32(X[0].real(), 0.f); // the packed real transform has no imaginary part in DC and Nyquist (fft.hpp:245,249)
        X[half] = cf32(X[half].real(), 0.f);
        float* sig = signals + c * 4 * half;
        magnitudeSpectrum<float>(X.data(), nfft, outputInDb != 0, false, full.data());
        std::copy_n(full.begin(), half, sig);
        phaseSpectrum<float>(X.data(), nfft, outputInDeg != 0, unwrapPhase != 0, false, full.data()); // unwrapping is a prefix operation
        std::copy_n(full.begin(), half, sig + half);
        for (std::size_t i = 0; i < half; ++i) {
            sig[2 * half + i] = X[half + i].real();
            sig[3 * half + i] = X[half + i].imag();
        }
        if (ranges != nullptr) {
            for (std::size_t s = 0; s < 4; ++s) {
                const auto mm               = std::minmax_element(sig + s * half, sig + (s + 1) * half);
                ranges[(c * 4 + s) * 2 + 0] = *mm.first;
                ranges[(c * 4 + s) * 2 + 1] = *mm.second;
            }
        }
    }
    return 0;
}

// blocks/math/include/gnuradio-4.0/math/Math.hpp:38-56 (scalar branch for complex<float>: `op()(a, value)` with the
// std::complex operators, i.e. libgcc's Annex-G multiply / divide). op: 0 add, 1 subtract, 2 multiply, 3 divide
int oracle_mathop_const_cf32(int op, const float* in, float* out, std::size_t n, float valueRe, float valueIm) {
    const cf32  value(valueRe, valueIm);
    const cf32* a = reinterpret_cast<const cf32*>(in);
    cf32*       y = reinterpret_cast<cf32*>(out);
    for (std::size_t i = 0; i < n; ++i) {
        switch (op) {
        case 0: y[i] = a[i] + value; break;
        case 1: y[i] = a[i] - value; break;
        case 2: y[i] = a[i] * value; break;
        case 3: y[i] = a[i] / value; break;
        default: return -1;
        }
    }
    return 0;
}

// Math.hpp:100-107: out = ins[0]; out = out op ins[k] for k = 1..
int oracle_mathop_multi_cf32(int op, const float* const* ins, std::size_t nInputs, float* out, std::size_t n) {
    cf32* y = reinterpret_cast<cf32*>(out);
    std::memcpy(out, ins[0], n * sizeof(cf32));
    for (std::size_t k = 1; k < nInputs; ++k) {
        const cf32* b = reinterpret_cast<const cf32*>(ins[k]);
        for (std::size_t i = 0; i < n; ++i) {
            switch (op) {
            case 0: y[i] = y[i] + b[i]; break;
            case 1: y[i] = y[i] - b[i]; break;
            case 2: y[i] = y[i] * b[i]; break;
            case 3: y[i] = y[i] / b[i]; break;
            default: return -1;
            }
        }
    }
    return 0;
}

// blocks/basic/include/gnuradio-4.0/basic/ConverterBlocks.hpp:258-277 (InterleavedToComplex<R, complex<float>>::processBulk):
// out[i] = {float(in[2i]), float(in[2i+1])}. itemType: 0 = float, 1 = int16_t, 2 = int8_t.
int oracle_interleaved_to_complex_cf32(int itemType, const void* in, float* out, std::size_t n) {
    cf32* y = reinterpret_cast<cf32*>(out);
    switch (itemType) {
    case 0: interleavedToComplex(static_cast<const float*>(in), y, n); return 0;
    case 1: interleavedToComplex(static_cast<const std::int16_t*>(in), y, n); return 0;
    case 2: interleavedToComplex(static_cast<const std::int8_t*>(in), y, n); return 0;
    default: return -1;
    }
}

// ConverterBlocks.hpp:235-256 (ComplexToInterleaved<complex<float>, R>::processBulk): out[2i] = static_cast<R>(re),
// out[2i+1] = static_cast<R>(im) -- truncation toward zero; what an out-of-range value becomes is whatever this
// compiler's cast does on this host (x86-64: cvttss2si, then the low bits), as in the compiled reference.
int oracle_complex_to_interleaved_cf32(int itemType, const float* in, void* out, std::size_t n) {
    const cf32* x = reinterpret_cast<const cf32*>(in);
    switch (itemType) {
    case 0: complexToInterleaved(x, static_cast<float*>(out), n); return 0;
    case 1: complexToInterleaved(x, static_cast<std::int16_t*>(out), n); return 0;
    case 2: complexToInterleaved(x, static_cast<std::int8_t*>(out), n); return 0;
    default: return -1;
    }
}

// blocks/math/include/gnuradio-4.0/math/Rotator.hpp:51-61: float phase accumulator, wrap to [0, 2pi], out = in * e^{j phase}
int oracle_rotator_cf32(const float* in, float* out, std::size_t n, float phaseIncrement, float* accumulatedPhase) {
    constexpr float twoPi = 2.f * std::numbers::pi_v<float>;
    float           phase = *accumulatedPhase;
    const cf32*     x     = reinterpret_cast<const cf32*>(in);
    cf32*           y     = reinterpret_cast<cf32*>(out);
    for (std::size_t i = 0; i < n; ++i) {
        phase += phaseIncrement;
        if (phase > twoPi) {
            phase -= twoPi;
        } else if (phase < 0.f) {
            phase += twoPi;
        }
        y[i] = x[i] * cf32(std::cos(phase), std::sin(phase));
    }
    *accumulatedPhase = phase;
    return 0;
}
// phase sequence only (no samples): phases[i] = accumulated phase used for sample i
int oracle_rotator_phases_f32(std::size_t n, float phaseIncrement, float* accumulatedPhase, float* phases) {
    constexpr float twoPi = 2.f * std::numbers::pi_v<float>;
    float           phase = *accumulatedPhase;
    for (std::size_t i = 0; i < n; ++i) {
        phase += phaseIncrement;
        if (phase > twoPi) {
            phase -= twoPi;
        } else if (phase < 0.f) {
            phase += twoPi;
        }
        if (phases != nullptr) {
            phases[i] = phase;
        }
    }
    *accumulatedPhase = phase;
    return 0;
}
// Rotator.hpp:40-49: phase_increment derived from frequency_shift / sample_rate (float arithmetic as written there)
float oracle_rotator_phase_increment_f32(float frequencyShift, float sampleRate) { return 2.f * static_cast<float>(std::numbers::pi_v<float> * frequencyShift / sampleRate); }

// ------------------------------------------------------------------------------------------------------------------
// Polyphase channelizer -- PARITY UNPINNED: the reference has no polyphase resampler / channelizer.
// Own definition (critically sampled, M channels, prototype h of length M*P, taps real):
//   for output frame t (consuming inputs x[tM .. tM+M-1], history zero before stream start):
//     v[p] = sum_{q=0}^{P-1} h[p + qM] * x[tM + (M-1-p)... ]  -- see below, written as the standard commutator form
//     u[r] = sum_{q} h[r + qM] * x[(t - q) M + (M - 1 - r)]      r = 0..M-1      (branch r sees every M-th sample)
//     y[t][k] = sum_{r} u[r] exp(-j 2 pi k r / M)                 k = 0..M-1      (M-point forward DFT, float)
// float arithmetic, q ascending, acc = fma(h, x, acc) (single rounding per tap). state: (P-1)*M complex samples of history
// (oldest first).
// ------------------------------------------------------------------------------------------------------------------
// stage 1 alone: u[t][r], r = 0..M-1, for nFrames frames
int oracle_pfb_filter_cf32(const float* proto, std::size_t nChannels, std::size_t tapsPerBranch, const float* in, float* out, std::size_t nFrames, float* state) {
    const std::size_t M = nChannels, P = tapsPerBranch, halo = (P - 1) * M;
    std::vector<cf32> work(halo + nFrames * M);
    if (state != nullptr) {
        std::memcpy(work.data(), state, halo * sizeof(cf32));
    }
    std::memcpy(work.data() + halo, in, nFrames * M * sizeof(cf32));
    cf32* u = reinterpret_cast<cf32*>(out);
    for (std::size_t t = 0; t < nFrames; ++t) {
        const cf32* frame = work.data() + halo + t * M; // x[(t)M + i] = frame[i]; x[(t-q)M + i] = frame[i - qM]
        for (std::size_t r = 0; r < M; ++r) {
            float accRe = 0.f, accIm = 0.f;
            for (std::size_t q = 0; q < P; ++q) {
                const cf32  x = frame[static_cast<std::ptrdiff_t>(M - 1 - r) - static_cast<std::ptrdiff_t>(q * M)];
                const float h = proto[r + q * M];
                accRe         = std::fma(h, x.real(), accRe); // one rounding per tap: our own definition, chosen for the device
                accIm         = std::fma(h, x.imag(), accIm);
            }
            u[t * M + r] = cf32(accRe, accIm);
        }
    }
    if (state != nullptr) {
        std::memcpy(state, work.data() + nFrames * M, halo * sizeof(cf32));
    }
    return 0;
}

int oracle_pfb_channelizer_cf32(const float* proto, std::size_t nChannels, std::size_t tapsPerBranch, const float* in, float* out, std::size_t nFrames, float* state) {
    const std::size_t M = nChannels;
    std::vector<cf32> u(nFrames * M), Y(M);
    oracle_pfb_filter_cf32(proto, nChannels, tapsPerBranch, in, reinterpret_cast<float*>(u.data()), nFrames, state);
    for (std::size_t t = 0; t < nFrames; ++t) {
        fftAny<float>(u.data() + t * M, Y.data(), M);
        std::memcpy(out + 2 * t * M, Y.data(), M * sizeof(cf32));
    }
    return 0;
}

} // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Polyphase rational resampler -- PARITY UNPINNED: like the channelizer there is no implementation in the reference.
// Own definition (interpolate by L, FIR h of length K, decimate by M, computed without the zero-stuffed samples):
//   y[m] = sum_{k=0}^{P-1} h[p_m + k L] * x[q_m - k],   p_m = (m M) mod L,  q_m = floor(m M / L),  P = ceil(K / L),
//   taps beyond K are zero, x[< 0] comes from `state` ((P-1) complex samples, oldest first; zero at stream start),
//   float arithmetic, k ascending, acc = fma(h, x, acc). nIn must be a multiple of M; nIn * L / M outputs.
// ------------------------------------------------------------------------------------------------------------------
extern "C" int oracle_resampler_cf32(const float* taps, std::size_t nTaps, std::size_t interpolation, std::size_t decimation, const float* in, float* out, std::size_t nIn, float* state) {
    const std::size_t L = interpolation, M = decimation;
    if (L == 0 || M == 0 || nTaps == 0 || nIn % M != 0) {
        return -1;
    }
    const std::size_t P = (nTaps + L - 1) / L, halo = P - 1, nOut = nIn / M * L;
    std::vector<cf32> work(halo + nIn);
    if (state != nullptr) {
        std::memcpy(work.data(), state, halo * sizeof(cf32));
    }
    std::memcpy(work.data() + halo, in, nIn * sizeof(cf32));
    cf32* y = reinterpret_cast<cf32*>(out);
    for (std::size_t m = 0; m < nOut; ++m) {
        const std::size_t p = (m * M) % L, q = (m * M) / L;
        float             accRe = 0.f, accIm = 0.f;
        for (std::size_t k = 0; k < P; ++k) {
            const std::size_t tapIndex = p + k * L;
            const float       h        = tapIndex < nTaps ? taps[tapIndex] : 0.f;
            const cf32        x        = work[halo + q - k];
            accRe                      = std::fma(h, x.real(), accRe);
            accIm                      = std::fma(h, x.imag(), accIm);
        }
        y[m] = cf32(accRe, accIm);
    }
    if (state != nullptr) {
        std::memcpy(state, work.data() + nIn, halo * sizeof(cf32));
    }
    return 0;
}
