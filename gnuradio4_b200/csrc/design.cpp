// Host-side coefficient design of the product: window tables and windowed-sinc FIR taps. These run once per settings
// change on the host (the reference does the same in settingsChanged) and their results are uploaded into the plans.
//   windows : algorithm/include/gnuradio-4.0/algorithm/fourier/window.hpp:71-183 (type numbering :35)
//   taps    : algorithm/include/gnuradio-4.0/algorithm/filter/FilterTool.hpp:964-976, DC/centre normalisation :415-423,
//             tap-count rule :985-1004, response types :1007-1071
// Expressions are evaluated in the element type and in the reference's operand order so that the tables come out
// bit-identical (tests/test_cabi.py compares them with the oracle, tests/test_gpu_golden.py with the reference fixtures).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstddef>
#include <numbers>
#include <vector>

#include "../../include/gr4b200.h"

namespace {

template<typename T>
T i0Series(T x) { // modified Bessel I0 by its power series, terminated at relative term^2 <= eps (window.hpp:42-56)
    const T h   = x / 2;
    T       acc = 1, term = 1;
    for (int k = 1;; ++k) {
        term *= h / static_cast<T>(k);
        acc += term * term;
        if (!(term * term > acc * std::numeric_limits<T>::epsilon())) {
            return acc;
        }
    }
}

struct CosineSum {
    int    terms;
    double a[5];
};
// generalised cosine windows a0 - a1 cos(x) + a2 cos(2x) - a3 cos(3x) + a4 cos(4x), x = 2 pi i / (N-1)
constexpr CosineSum kCosineSums[] = {
    /* Hamming        */ {2, {0.53836, 0.46164}},
    /* Hann           */ {2, {.5, .5}},
    /* Blackman       */ {3, {0.42, 0.5, 0.08}},
    /* Nuttall        */ {4, {0.355768, 0.487396, 0.144232, 0.012604}},
    /* BlackmanHarris */ {4, {0.35875, 0.48829, 0.14128, 0.01168}},
    /* BlackmanNuttall*/ {4, {0.3635819, 0.4891775, 0.1365995, 0.0106411}},
    /* FlatTop        */ {5, {1.0, 1.93, 1.29, 0.388, 0.032}},
};

template<typename T>
int buildWindow(int type, std::size_t n, T beta, T* w) {
    if (n == 0) {
        return 0;
    }
    const T   dx       = (2 * std::numbers::pi_v<T>) / static_cast<T>(n - 1);
    int       cosineId = -1;
    switch (type) {
    case 0: // None
    case 1: std::fill_n(w, n, T(1)); return 0; // Rectangular
    case 2: cosineId = 0; break;
    case 3: cosineId = 1; break;
    case 4: // HannExp
        for (std::size_t i = 0; i < n; ++i) {
            w[i] = std::pow(std::sin(dx * static_cast<T>(i)), static_cast<T>(2.));
        }
        return 0;
    case 5: cosineId = 2; break;
    case 6: cosineId = 3; break;
    case 7: cosineId = 4; break;
    case 8: cosineId = 5; break;
    case 9: cosineId = 6; break;
    case 10: { // Exponential
        const T unit = std::exp(static_cast<T>(0.));
        const T tau  = static_cast<T>(3.) * static_cast<T>(n);
        for (std::size_t i = 0; i < n; ++i) {
            w[i] = std::exp(static_cast<T>(i) / tau) / unit;
        }
        return 0;
    }
    case 11: { // Kaiser
        if (beta < 0 || n <= 1) {
            return GR4B200_ERROR;
        }
        const T inv   = static_cast<T>(1) / static_cast<T>(n - 1);
        const T denom = i0Series(beta);
        for (std::size_t i = 0; i < n; ++i) {
            const T u = (static_cast<T>(2 * i) * inv) - static_cast<T>(1);
            w[i]      = i0Series(beta * std::sqrt(std::abs(static_cast<T>(1) - u * u))) / denom;
        }
        return 0;
    }
    default: return GR4B200_ERROR;
    }
    const CosineSum& cs = kCosineSums[cosineId];
    for (std::size_t i = 0; i < n; ++i) {
        const T x   = dx * static_cast<T>(i);
        T       acc = static_cast<T>(cs.a[0]);
        for (int h = 1; h < cs.terms; ++h) {
            const T c = static_cast<T>(cs.a[h]) * std::cos(h == 1 ? x : static_cast<T>(h) * x);
            acc       = (h % 2 == 1) ? acc - c : acc + c;
        }
        w[i] = acc;
    }
    return 0;
}

template<typename T>
int windowedSinc(std::size_t n, int window, T fc, T beta, T* b) {
    if (buildWindow<T>(window, n, beta, b) != 0) {
        return GR4B200_ERROR;
    }
    const T centre = static_cast<T>(n - 1) / static_cast<T>(2);
    for (std::size_t i = 0; i < n; ++i) {
        const T x = static_cast<T>(2) * fc * (static_cast<T>(i) - centre);
        const T px = std::numbers::pi_v<T> * x;
        b[i]      = b[i] * static_cast<T>(2) * fc * (x == static_cast<T>(0) ? static_cast<T>(1) : std::sin(px) / px);
    }
    return 0;
}

template<typename T>
T magnitudeAt(const T* b, std::size_t n, T fNorm) { // |sum b_k e^{-j w k}| via integer powers of e^{jw} (FilterTool.hpp:376-404)
    const std::complex<T> z = std::polar(static_cast<T>(1), static_cast<T>(2) * std::numbers::pi_v<T> * fNorm);
    std::complex<T>       num(0);
    for (std::size_t k = 0; k < n; ++k) {
        num = num + b[k] * std::pow(z, -static_cast<int>(k));
    }
    const std::complex<T> den = std::complex<T>(0) + static_cast<T>(1) * std::pow(z, 0);
    return static_cast<T>(1) * std::abs(num / den);
}

template<typename T>
bool scaleToGain(T* b, std::size_t n, T fNorm, T gain) {
    const T m = magnitudeAt(b, n, fNorm);
    if (m == 0) {
        return false;
    }
    for (std::size_t i = 0; i < n; ++i) {
        b[i] = b[i] * gain / m;
    }
    return true;
}

std::size_t tapCount(int type, std::size_t order, double fLow, double fHigh, double fs, double attenuationDb) {
    double width = 0.1 / static_cast<double>(order);
    switch (type) {
    case 0: width = std::min(width, std::min(std::abs(fLow / fs), std::abs(0.5 - fLow / fs))); break;
    case 1: width = std::min(width, std::abs(fHigh / fs)); break;
    case 2: width = std::min(width, std::min(std::abs(fLow / fs), std::abs(0.5 - fHigh / fs))); break;
    case 3: width = std::min(width, std::min(std::abs(0.5 - fHigh / fs), std::min(fLow, 0.5 * std::abs(fHigh - fLow)) / fs)); break;
    default: return 0;
    }
    auto n = static_cast<std::size_t>(std::ceil((attenuationDb - 8.0) / (2.285 * (2. * std::numbers::pi * width))));
    return n | 1u; // odd
}

} // namespace

extern "C" {

int gr4b200_window_f32_host(int windowType, size_t n, float beta, float* out_host) { return buildWindow<float>(windowType, n, beta, out_host); }

int gr4b200_fir_generate_f32_host(size_t nTaps, int windowType, float fc, float beta, int normaliseDc, float* out_host) {
    if (nTaps == 0 || windowedSinc<float>(nTaps, windowType, fc, beta, out_host) != 0) {
        return GR4B200_ERROR;
    }
    if (normaliseDc != 0 && !scaleToGain<float>(out_host, nTaps, 0.f, 1.f)) {
        return GR4B200_ERROR;
    }
    return GR4B200_OK;
}

long gr4b200_fir_design_f32_host(int filterType, size_t order, double fLow, double fHigh, double fs, double gain, double attenuationDb, double beta, int windowType, float* out_host, size_t capacity) {
    using T             = float;
    const std::size_t n = tapCount(filterType, order, fLow, fHigh, fs, attenuationDb);
    if (n == 0) {
        return GR4B200_ERROR;
    }
    if (n > capacity) {
        return -static_cast<long>(n);
    }
    std::vector<T> b(n), other(n);
    const T        g = static_cast<T>(gain), bt = static_cast<T>(beta);
    bool           ok = false;
    switch (filterType) {
    case 0: ok = windowedSinc<T>(n, windowType, static_cast<T>(fLow / fs), bt, b.data()) == 0 && scaleToGain<T>(b.data(), n, T(0), g); break;
    case 1:
        ok = windowedSinc<T>(n, windowType, static_cast<T>(0.5 - fHigh / fs), bt, b.data()) == 0;
        for (std::size_t i = 0; ok && i < n; ++i) {
            b[i] *= (i % 2 == 0 ? 1 : -1);
        }
        ok = ok && scaleToGain<T>(b.data(), n, static_cast<T>(0.48), g);
        break;
    case 2:
        ok = windowedSinc<T>(n, windowType, static_cast<T>(fLow / fs), bt, b.data()) == 0 && windowedSinc<T>(n, windowType, static_cast<T>(fHigh / fs), bt, other.data()) == 0;
        for (std::size_t i = 0; ok && i < n; ++i) {
            b[i] = b[i] - other[i];
        }
        ok = ok && scaleToGain<T>(b.data(), n, static_cast<T>(std::sqrt(fHigh * fLow) / fs), g);
        break;
    case 3:
        ok = windowedSinc<T>(n, windowType, static_cast<T>(fLow / fs), bt, b.data()) == 0 && windowedSinc<T>(n, windowType, static_cast<T>(fHigh / fs), bt, other.data()) == 0;
        for (std::size_t i = 0; ok && i < n; ++i) {
            b[i] -= other[i];
            if (n % 2 != 0 && i == (n - 1) / 2) {
                b[i] = 1 - b[i];
            }
        }
        ok = ok && scaleToGain<T>(b.data(), n, T(0), g);
        break;
    default: return GR4B200_ERROR;
    }
    if (!ok) {
        return GR4B200_ERROR;
    }
    std::copy(b.begin(), b.end(), out_host);
    return static_cast<long>(n);
}

} // extern "C"
